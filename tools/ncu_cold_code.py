#!/usr/bin/env python
"""Never-executed SASS inside the step kernel, as address runs and per source line (cold code inside the lockstep
instruction stream costs fetch slots even when it is branched over).

    ncu -i prof.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_cold_code.py src.csv [min_run]
"""
import collections
import csv
import sys


def main(path, min_run=24):
    rows = list(csv.reader(open(path)))
    cur, curline, seq, text = None, None, [], {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        try:
            curline = (cur, int(r[0]))
            text[curline] = r[1]
            continue
        except ValueError:
            pass
        if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
            seq.append((int(r[2], 16), int(r[7] or 0), curline))
    seq.sort()
    base = seq[0][0]
    print(f"{len(seq)} SASS instructions, {sum(1 for s in seq if s[1] == 0)} never executed in this launch")
    i = 0
    while i < len(seq):
        if seq[i][1] == 0:
            j = i
            while j < len(seq) and seq[j][1] == 0:
                j += 1
            if j - i >= int(min_run):
                lines = collections.Counter(s[2] for s in seq[i:j])
                top = ", ".join(f"{k[0][:14]}:{k[1]}x{v}" for k, v in lines.most_common(5))
                print(f"  +{seq[i][0] - base:#08x} {j - i:5d} instr | {top}")
            i = j
        else:
            i += 1


if __name__ == "__main__":
    main(*sys.argv[1:])
