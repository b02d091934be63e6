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
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_ldgsts.sum", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def traffic_json(rows, hdr, config: str, out_path: str):
    """profiles/r02_ncu_traffic.json: dram read + write bytes per launch of the three hot kernels of one bench step, keyed by
    config; bench.py reads `roofline.traffic` from it.  Capture order of a step: forward, wgrad, dgrad (wgrad runs first in
    the backward so that its all-reduce overlaps dgrad)."""
    import json
    import os

    name_i, rd_i, wr_i = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units_row = rows[1]
    fwd_like = [r for r in rows[2:] if "conv_tc_fwd" in r[name_i]]
    wgrad = [r for r in rows[2:] if "conv_tc_wgrad" in r[name_i]]
    val = lambda r: float(r[rd_i].replace(",", "")) * unit.get(units_row[rd_i], 1.0) + float(r[wr_i].replace(",", "")) * unit.get(units_row[wr_i], 1.0)  # noqa: E731
    rec = {}
    if fwd_like:
        rec["fwd"] = val(fwd_like[0])
    if len(fwd_like) > 1:
        rec["dgrad"] = val(fwd_like[1])
    if wgrad:
        rec["wgrad"] = val(wgrad[0])
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data[config] = rec
    data["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump(data, open(out_path, "w"), indent=1)


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if len(r) > 8]
    hdr, units = rows[0], rows[1]
    if len(sys.argv) > 4:  # ncu_summary.py rep "header" <config> <traffic.json>
        traffic_json(rows, hdr, sys.argv[3], sys.argv[4])
    cols = [hdr.index(w) for w in WANT if w in hdr]
    for r in rows[2:]:
        print("-----")
        for i in cols:
            print(f"{hdr[i]:<75} {r[i][:200]} {units[i]}")


if __name__ == "__main__":
    main()
