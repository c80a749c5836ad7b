#!/bin/bash
# A/B of tuning builds on the GPU box: tools/ab_libs.sh "<env sizes>" lib1.so lib2.so ...  ("default" = the in-tree library)
sizes="$1"; shift
mkdir -p gpurun_out
for lib in "$@"; do
  for E in $sizes; do
    steps=1400; [ "$E" -gt 4096 ] && steps=400
    if [ "$lib" = default ]; then unset B2H_LIB; else export B2H_LIB=$PWD/$lib; fi
    python bench.py --quick --steps $steps --warmup 100 --n-envs $E 2>&1 | tail -1 | python -c "
import json,sys
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('$lib', d['n_envs'], '%.3fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'iter %.3f' % d['newton_iter'], d['launch'])
except Exception as e:
    print('$lib', 'FAILED', l[-300:])
"
  done
done
