#!/usr/bin/env python
"""Per source line: which stall reason the samples fall under.  python tools/ncu_stalls.py src.csv [stall_long_sb] [top]"""
import collections
import csv
import sys


def main(path, reason="stall_long_sb", top=25):
    rows = list(csv.reader(open(path)))
    hdr, fname = None, None
    agg = collections.Counter()
    tot = collections.Counter()
    text = {}
    # the CUDA-source view: sections start with "File Path"
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit() and fname:
            i = hdr.index(reason)
            try:
                v = int(r[i] or 0)
            except ValueError:
                continue
            agg[(fname, int(r[0]))] += v
            text[(fname, int(r[0]))] = r[1].strip()[:110]
            for k in range(hdr.index("stall_barrier"), hdr.index("stall_wait") + 1):
                try:
                    tot[hdr[k]] += int(r[k] or 0)
                except ValueError:
                    pass
    print("totals:", dict(tot.most_common()))
    s = sum(agg.values())
    for (f, ln), v in agg.most_common(int(top)):
        print(f"{v:6d} {100.0 * v / max(1, s):5.1f}%  {f}:{ln}  {text[(f, ln)]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
