#!/usr/bin/env python3
"""Turns an Nsight Compute report (.ncu-rep) into the short text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp16.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, unit, r)}
        print("kernel:", d.get("Kernel Name", ("", ""))[1])
        for k in KEYS:
            for h in hdr:
                if h == k or h.startswith(k + " "):
                    print(f"  {h:75s} {d[h][1]:>18s} {d[h][0]}")
        print("  warp stall reasons (cycles per issued instruction):")
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[h][1])
                except ValueError:
                    continue
                if v >= 0.05:
                    name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                    print(f"    {name:30s} {v:8.3f}")
        print()


if __name__ == "__main__":
    main()
