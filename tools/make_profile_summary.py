"""profiles/<name>.txt from an ncu report: the raw-page metrics the roofline cites, stall samples and a region breakdown.

usage: python tools/make_profile_summary.py gpurun_out/prof_full.ncu-rep profiles/r1_run_kernel_ncu_full.txt "<command that was profiled>"
"""
import csv
import io
import subprocess
import sys

rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = """gpu__time_duration.sum launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum dram__bytes_write.sum
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active
l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum lts__t_sector_hit_rate.pct""".split()
lines = [f"# {cmd}", f"# kernel: {d['Kernel Name'][0]}, grid {d['Grid Size'][0]} x block {d['Block Size'][0]}"]
for w in want:
    if w in d:
        lines.append(f"{w:85s} {d[w][0]:>18s} {d[w][1]}")
lines.append("# warp stall samples (smsp__pcsamp_warps_issue_stalled_*)")
for k in sorted(h for h in d if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
    lines.append(f"  {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):24s} {d[k][0]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
sh, data = srows[1], srows[2:]
ci, csamp = sh.index("Instructions Executed"), sh.index("# Samples")
grid = int(d["Grid Size"][0].strip("()").split(",")[0])
warps = grid * int(d["Block Size"][0].strip("()").split(",")[0]) // 32
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
per = warps * steps
tot_i = sum(int(r[ci]) for r in data)
tot_s = sum(int(r[csamp]) for r in data)
lines.append(f"# SASS page: {tot_i / per:.0f} instructions per warp per MD step ({warps} warps x {steps} steps), {tot_s} stall samples")
lines.append("# instruction / sample share by 250-row SASS bucket (row range: instructions per warp-step, % of samples)")
for a in range(0, len(data), 250):
    i = sum(int(r[ci]) for r in data[a:a + 250]) / per
    s = sum(int(r[csamp]) for r in data[a:a + 250])
    if i > 5 or s > 0.01 * tot_s:
        lines.append(f"#   rows {a:5d}-{a + 249:5d}: {i:7.1f}  {100.0 * s / tot_s:5.1f} %")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
