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


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale[unit]


def traffic(args):
    """--traffic key=report.ncu-rep:cells [...] : dram__bytes_read.sum + dram__bytes_write.sum of every launch in the report,
    per interior cell the launch processed (mean over the launches), merged into profiles/ncu_traffic.json, which bench.py
    reads for `roofline.traffic`."""
    import json
    import os
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    rec = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for a in args:
        key, rest = a.split("=", 1)
        path, cells = rest.rsplit(":", 1)
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        per = []
        for vals in rows[2:]:
            d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
            rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
            wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            per.append({"kernel": d.get("Kernel Name", "?")[:80], "dram_read": rd, "dram_write": wr,
                        "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[u["gpu__time_duration.sum"]]})
        tot = sum(x["dram_read"] + x["dram_write"] for x in per)
        rec[key] = {"dram_bytes_per_cell": tot / len(per) / float(cells), "launches": len(per), "cells_per_launch": int(cells),
                    "source": os.path.basename(path) + " (ncu --set full --clock-control none; summary profiles/" + os.path.basename(path).replace(".ncu-rep", ".txt") + ")",
                    "per_launch": per}
        print(key, rec[key]["dram_bytes_per_cell"], "B/cell over", len(per), "launches")
    json.dump(rec, open(out_path, "w"), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        return traffic(sys.argv[2:])
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
