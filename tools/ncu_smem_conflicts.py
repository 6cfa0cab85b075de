#!/usr/bin/env python
"""Shared-memory wavefronts per SASS instruction of the FIRST kernel of an .ncu-rep (source page): the instructions with
excessive (bank-conflict) wavefronts.  usage: ncu_smem_conflicts.py report.ncu-rep"""
import csv
import io
import subprocess
import sys


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
    end = next((i for i in range(starts[0] + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[starts[0]:end]))))
    num = lambda r, k: int(r[k] or 0)
    tw, te, ti = (sum(num(r, k) for r in rows) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared Ideal"))
    print(f"{len(rows)} SASS instructions: shared wavefronts {tw}, ideal {ti}, excessive {te} ({100.0*te/max(tw,1):.1f}%)")
    for i, r in enumerate(rows):
        e = num(r, "L1 Wavefronts Shared Excessive")
        if e > 0.02 * max(te, 1):
            print(f"{i:5d} {r['Source'].strip()[:60]:60s} executed {r['Instructions Executed']:>9s} wavefronts {r['L1 Wavefronts Shared']:>10s} ideal {r['L1 Wavefronts Shared Ideal']:>10s} excessive {e}")


if __name__ == "__main__":
    main()
