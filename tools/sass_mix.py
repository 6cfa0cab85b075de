#!/usr/bin/env python
"""Static SASS instruction mix per kernel: `cuobjdump -sass spade_b200/libspade_b200.so | python tools/sass_mix.py`.
No GPU needed. UTMALDG / UTMASTG / UTMAPF are the TMA tensor load / store / prefetch, SYNCS the mbarrier operations."""
import collections
import re
import subprocess
import sys

KEYS = ["DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "SHFL",
        "IMAD", "MOV"]


def main():
    cur, mix = None, collections.OrderedDict()
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            mix[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            mix[cur][m.group(1).split(".")[0]] += 1
    names = list(mix)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("static SASS instruction counts per kernel (sm_100a); loop bodies count once")
    print(f"{'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS) + "  kernel")
    for n, d in sorted(zip(names, dem), key=lambda x: x[1]):
        c = mix[n]
        short = re.sub(r"\(.*", "", d).replace("spb::", "").replace("void ", "")
        print(f"{sum(c.values()):7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEYS) + "  " + short[:120])


if __name__ == "__main__":
    main()
