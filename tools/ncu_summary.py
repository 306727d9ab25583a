"""Prints the key metrics of every launch in an .ncu-rep (ncu --set full) — time, DRAM bytes, stalls, pipes."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
for r in rows[2:]:
    for w in want:
        if w in H:
            print(f"{w:70s} {r[H.index(w)]}")
    st = [(float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per_")[0]) for i, h in enumerate(H)
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    print("stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    print("-" * 100)
