"""Per-kernel SASS opcode histogram of libidb200.so (no GPU needed): what proves a Blackwell-native kernel -- UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA tensor load / store / reduce), UBLKCP (cp.async.bulk), SYNCS
(mbarrier) -- next to the legacy HMMA (mma.sync) count.    python tools/sass_histogram.py [lib.so] > profiles/r02_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "interpolated_diffusion_b200", "libidb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "LDSM", "FFMA", "MUFU", "REDUX"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        per[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in line:
            per[cur]["UTCHMMA.2CTA"] += 1
        per[cur]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode histogram of {os.path.basename(lib)} (cuobjdump -sass), kernels with tensor-core / TMA / TMEM instructions first")
print("# " + " ".join(f"{k:>12}" for k in KEYS) + "        total  kernel")
tot = collections.Counter()
rows = []
for (name, c), dn in zip(per.items(), demangle):
    for k in KEYS:
        tot[k] += c[k]
    short = re.sub(r"\(.*", "", dn)[:100]
    rows.append((-(c["UTCHMMA"] + c["LDTM"] + c["UTMALDG"] + c["UBLKCP"]), -c["HMMA"], "  " + " ".join(f"{c[k]:>12}" for k in KEYS) + f" {c['_total']:>12}  {short}"))
for _, _, r in sorted(rows):
    print(r)
print("# totals: " + ", ".join(f"{k}={tot[k]}" for k in KEYS))
