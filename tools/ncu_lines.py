"""Aggregate an .ncu-rep source page by CUDA source line: warp instructions executed and stall samples per file:line."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict(); cur = None; fname = "?"; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ia = r.index("Address"); iex = r.index("Instructions Executed"); ist = r.index("Warp Stall Sampling (All Samples)"); stc = [(i, c[6:]) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0]:
        cur = (fname, int(r[0]), r[1].strip()[:90]); agg.setdefault(cur, [0, 0, collections.Counter()]); continue
    if cur is None: continue
    try:
        agg[cur][0] += int(r[iex]); agg[cur][1] += int(r[ist])
        for i, nm in stc:
            if r[i] not in ("", "0"): agg[cur][2][nm] += int(r[i])
    except ValueError: pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot_i} ({tot_i/per:.0f} per unit), stall samples {tot_s}")
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
for (f, ln, src), (i, s, c) in agg.items():
    if i / max(tot_i, 1) > thr or s / max(tot_s, 1) > thr:
        top = " ".join(f"{k}:{v}" for k, v in c.most_common(3))
        print(f"{f}:{ln:4d} inst/unit {i/per:8.1f} ({100*i/tot_i:4.1f}%) stalls {100*s/max(tot_s,1):4.1f}% [{top}]  {src}")
