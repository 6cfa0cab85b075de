#!/usr/bin/env python
"""Condensed view of an `ncu --set full` report: `python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]`.
Prints the metrics DESIGN.md and profiles/ quote (duration, DRAM bytes, pipe utilisation, stall reasons)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "sm__cycles_elapsed.max"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"== {path}: {d.get('Kernel Name', '?')[:100]}")
            for k in KEYS:
                if k in d:
                    print(f"  {k:72s} {d[k]:>18s} {u[k]}")
            stalls = sorted(((float(v), k) for k, v in d.items()
                             if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v),
                            reverse=True)
            if not stalls:
                stalls = sorted(((float(v), k) for k, v in d.items()
                                 if k.startswith("smsp__average_warp_latency_issue_stalled") and v), reverse=True)
            for v, k in stalls[:8]:
                print(f"  stall {k.split('issue_stalled_')[1]:60s} {v:10.3f}")


if __name__ == "__main__":
    main()
