"""Per-launch metric summary of EVERY kernel in an .ncu-rep: `python profiles/ncu_kernels.py X.ncu-rep > X.txt`
(runs `ncu -i X --page raw --csv`)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
ki = h.index("Kernel Name")
for r in rows[2:]:
    print(f"---- {r[ki]}")
    for k in want:
        if k in h:
            i = h.index(k)
            print(f"  {k:85s} {r[i]} {u[i]}")
