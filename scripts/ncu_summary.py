"""Extract a compact per-kernel summary from an .ncu-rep (read here on CPU): python scripts/ncu_summary.py rep out.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected"]
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (per launch; cold-cache, serialised)\n")
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        f.write(f"\n== {d.get('Kernel Name')}  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}\n")
        for k in h:
            if any(k.endswith(x) or k == x for x in KEYS):
                f.write(f"  {k:90s} {d[k]:>16s} {u.get(k, '')}\n")
print(open(out).read()[:3000])
