"""Summarise an .ncu-rep: headline metrics per kernel + hot SASS blocks (run here, no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"{k} = {r[i]} {rows[1][i]}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Kernel Name": break
    try: data.append((r[ia], r[isrc], int(r[iex]), int(r[ist])))
    except ValueError: pass
tot = sum(d[2] for d in data)
print(f"warp-instructions {tot}  per unit {tot/per:.1f}  stall samples {sum(d[3] for d in data)}")
if '--top' in sys.argv:
    for d in sorted(data, key=lambda x: -x[3])[:30]: print(d[3], f"{d[2]/per:.2f}", d[1][:90])
if '--dump' in sys.argv:
    for a, s, e, st in data:
        if e / per >= 0.3: print(f"{e/per:6.2f} {st:6d} {s[:100]}")
