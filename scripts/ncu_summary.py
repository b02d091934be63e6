"""Summarise an .ncu-rep (ncu --set full) into the few counters DESIGN.md / bench.py cite.

    python scripts/ncu_summary.py gpurun_out/prof_tc.ncu-rep "header line" > profiles/rNN_ncu_full_<tag>_summary.txt
"""
import csv
import subprocess
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(w) for w in WANT if w in hdr]
    for r in rows[2:]:
        print("-----")
        for i in cols:
            print(f"{hdr[i]:<75} {r[i][:200]} {units[i]}")


if __name__ == "__main__":
    main()
