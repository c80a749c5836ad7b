#!/usr/bin/env python
"""Static SASS instruction count per source region of b2h_physics.cuh.

Usage: cuobjdump -xelf all libb2h.so; nvdisasm -g -c b2h_api.sm_100a.cubin > dis.txt; python tools/sass_by_stage.py dis.txt [kernel-substring]
"""
import collections
import re
import sys

KEYS = [("position stage", "position"), ("collision (lane", "collision"), ("constraint rows", "rows+J"), ("velocity stage", "velocity"),
        ("acceleration: qacc_smooth", "accel"), ("mj_fwdConstraint: Newton", "newton"), ("warmstart(): the cheaper", "n:warmstart"),
        ("PrimalUpdateConstraint", "n:update"), ("Hessian H = M", "n:hessian"), ("PrimalUpdateGradient", "n:grad"),
        ("PrimalSearch: exact", "n:linesearch"), ("if (alpha == T(0)) break", "n:move"), ("mj_Euler", "euler"), ("env layer", "env")]


def main(path, kernel="step_kernelIf", src_path="mujocoposelearning_b200/csrc/b2h_physics.cuh"):
    lines = open(path).read().split("\n")
    heads = [i for i, l in enumerate(lines) if l.startswith(".text.")]
    start = next(i for i in heads if kernel in lines[i])
    end = min([i for i in heads if i > start] + [len(lines)])
    cur, cnt, total = None, collections.Counter(), 0
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            cnt[cur] += 1
            total += 1
    src = open(src_path).read().split("\n")
    first = next(i for i, l in enumerate(src) if "bool physics_step(" in l)
    marks = [(1, "helpers")]
    for key, name in KEYS:
        for i, l in enumerate(src):
            if i >= first and key in l:
                marks.append((i + 1, name))
                break
    marks.sort()
    b = collections.Counter()
    for (f, ln), c in cnt.items():
        if f == "b2h_physics.cuh":
            nm = "helpers"
            for m_, n_ in marks:
                if ln >= m_:
                    nm = n_
            b[nm] += c
        else:
            b["<" + str(f) + ">"] += c
    print(f"{kernel}: {total} static SASS instructions = {total * 16 / 1024:.0f} KB")
    for k, v in b.most_common():
        print(f"  {k:28s} {v:6d}")
    print("top lines")
    for (f, ln), c in sorted(cnt.items(), key=lambda x: -x[1])[:30]:
        print(f"  {c:6d} {f}:{ln}  {src[ln - 1].strip()[:100] if f == 'b2h_physics.cuh' else ''}")


if __name__ == "__main__":
    main(*sys.argv[1:])
