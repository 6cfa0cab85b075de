#!/usr/bin/env python
"""Per-SASS-instruction stall samples of one kernel from an .ncu-rep (ncu --set full --import-source on):
prints the instructions that collect the most warp-stall samples, with the dominant stall reason, and the
totals per reason at the BAR.SYNC / SYNCS instructions.  usage: ncu_source_stalls.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    reasons = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    total = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"{len(rows)} SASS instructions, {total} samples")
    per_reason = {c: sum(int(r[c] or 0) for r in rows) for c in reasons}
    print("per reason:", ", ".join(f"{c[6:]}={v} ({100*v/total:.1f}%)" for c, v in sorted(per_reason.items(), key=lambda x: -x[1]) if v))
    ranked = sorted(enumerate(rows), key=lambda x: -int(x[1]["# Samples"] or 0))[:top]
    for idx, r in sorted(ranked, key=lambda x: x[0]):
        n = int(r["# Samples"] or 0)
        dom = max(reasons, key=lambda c: int(r[c] or 0))
        print(f"{idx:5d} {100*n/total:5.1f}%  {dom[6:]:14s} {r['Source'].strip()[:90]}")


if __name__ == "__main__":
    main()
