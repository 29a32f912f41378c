"""Text summary of one kernel from an .ncu-rep: `python profiles/ncu_summary.py X.ncu-rep > X.txt`
(runs `ncu -i X --page raw --csv` and `--page source --csv --print-source sass`)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed.sum", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
print(f"# {rep}")
for k, uu, x in zip(h, u, v):
    if k in want or ("issue_stalled" in k and "per_issue_active" in k and x and float(x) > 0.25):
        print(f"{k:95s} {x} {uu}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]
si, ei, st = hh.index("Source"), hh.index("Instructions Executed"), hh.index("# Samples")
data = [(r[si], int(r[ei]), int(r[st])) for r in rows[2:] if len(r) > st and r[ei].isdigit() and r[st].isdigit()]
tot = sum(n for _, n, _ in data)
stot = sum(s for _, _, s in data) or 1
print(f"\nwarp instructions executed {tot}, SASS lines {len(data)}, stall samples {stot}")
hist = collections.Counter()
for s_, n, _ in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", s_)
    hist[m.group(2) if m else s_[:10]] += n
print("opcode mix: " + ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in hist.most_common(14)))
print("\nmost-sampled SASS lines (index, samples, executed, instruction):")
for s__, i in sorted(((s, i) for i, (_, _, s) in enumerate(data)), reverse=True)[:14]:
    print(f"  {i:5d} {s__:6d} {data[i][1]:10d}  {data[i][0].strip()[:70]}")
