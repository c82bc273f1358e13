"""Turns an .ncu-rep into the small CSV summaries kept under profiles/ (run in the build container)."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__shared_mem_per_block', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__average_warps_issue_stalled', 'sm__cycles_elapsed.max',
        'sass__inst_executed_local', 'smsp__thread_inst_executed_per_inst_executed']
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["launch", "metric", "unit", "value"])
    for li, vals in enumerate(rows[2:]):
        for h, u, v in zip(hdr, units, vals):
            if any(k in h for k in keep) and "not_issued" not in h:
                w.writerow([li, h, u, v])
print("wrote", out)
