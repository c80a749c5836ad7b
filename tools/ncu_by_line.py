#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export per source line / per pipeline stage.

Usage: ncu -i prof.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_by_line.py src.csv
"""
import collections
import csv
import sys

STAGES = [  # (first line marker in b2h_physics.cuh, name)
    ("position stage", "kinematics+com+crb"), ("collision (lane", "collision"), ("constraint rows", "constraint rows + J"),
    ("velocity stage", "comvel+rne+passive"), ("acceleration: qacc_smooth", "factor M + solve"),
    ("mj_fwdConstraint: Newton", "newton"), ("mj_Euler", "euler"), ("env layer", "env epilogue"),
]
SUB = [("warmstart(): the cheaper", "newton: warmstart"), ("PrimalUpdateConstraint", "newton: update constraint + J^T f"),
       ("Hessian H = M", "newton: hessian + factor"), ("PrimalUpdateGradient", "newton: gradient solve"),
       ("PrimalSearch: exact", "newton: line search"), ("if (alpha == T(0)) break", "newton: move")]


def main(path, src_path="mujocoposelearning_b200/csrc/b2h_physics.cuh"):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            sections.append(cur)
        elif r and r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and r and r[0] not in ("Function Name",):
            cur["rows"].append(r)
    src = open(src_path).read().split("\n")
    marks = []
    first = next(i for i, l in enumerate(src) if " physics_step(const DevModel" in l)
    marks.append((1, "helpers (inlined math, cholesky, mat_vec)"))
    for key, name in STAGES + SUB:
        for i, l in enumerate(src):
            if i >= first and key in l:
                marks.append((i + 1, name))
                break
    marks.sort()
    done = set()
    total = collections.Counter()
    stall = collections.Counter()
    byline = collections.Counter()
    byline_stall = collections.Counter()
    other = collections.Counter()
    for s in sections:
        key = (s["file"],)
        if key in done:
            continue  # the export repeats per profiled launch; take the first
        done.add(key)
        h = s["hdr"]
        iL, iI, iS = h.index("Line No"), h.index("Instructions Executed"), h.index("# Samples")
        for r in s["rows"]:
            try:
                ln, n, sm = int(r[iL]), int(r[iI] or 0), int(r[iS] or 0)
            except ValueError:
                continue
            if s["file"].endswith("b2h_physics.cuh"):
                byline[ln] += n
                byline_stall[ln] += sm
                name = "prologue/helpers"
                for m, nm in marks:
                    if ln >= m:
                        name = nm
                total[name] += n
                stall[name] += sm
            else:
                other[s["file"].split("/")[-1]] += n
                total["<" + s["file"].split("/")[-1] + ">"] += n
                stall["<" + s["file"].split("/")[-1] + ">"] += sm
    tot = sum(total.values())
    ts = sum(stall.values())
    print(f"total warp instructions {tot}, stall samples {ts}")
    for k, v in sorted(total.items(), key=lambda x: -x[1]):
        print(f"{k:40s} inst {v:12d} {100*v/tot:5.1f}%   samples {stall[k]:8d} {100*stall[k]/max(1,ts):5.1f}%")
    print("--- top lines")
    for ln, v in byline.most_common(40):
        print(f"{ln:5d} {100*v/tot:5.1f}% s{100*byline_stall[ln]/max(1,ts):5.1f}%  {src[ln-1].strip()[:110]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
