"""Headline metrics + stall samples of the first kernel in an ncu report (development aid)."""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
for w in """gpu__time_duration.sum launch__registers_per_thread launch__waves_per_multiprocessor launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum sm__cycles_elapsed.avg sm__cycles_active.avg dram__bytes_read.sum dram__bytes_write.sum
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct""".split():
    if w in d:
        print(f"{w:75s} {d[w][0]:>16s} {d[w][1]}")
for k in sorted((h for h in d if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h), key=lambda h: -float(d[h][0] or 0)):
    print("  ", k.replace("smsp__pcsamp_warps_issue_stalled_", ""), d[k][0])
