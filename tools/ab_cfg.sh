#!/bin/bash
# A/B of launch configurations on the GPU box: tools/ab_cfg.sh "<env sizes>" "label;lib;ENV=V ENV=V" ...   (lib "default" = in-tree)
sizes="$1"; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  IFS=';' read -r label lib envs <<< "$cfg"
  for E in $sizes; do
    steps=600; [ "$E" -gt 4096 ] && steps=200; [ "$E" -gt 16384 ] && steps=60
    (
      [ "$lib" != default ] && export B2H_LIB=$PWD/$lib
      for kv in $envs; do export "$kv"; done
      python bench.py --quick --steps $steps --warmup 20 --n-envs $E 2>&1 | tail -1 | python -c "
import json,sys
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('$label', d['n_envs'], '%.3fM' % (d['value']/1e6), 'ms %.4f' % d['ms_per_step'], 'iter %.3f' % d['newton_iter'], d['launch'])
except Exception as e:
    print('$label', 'FAILED', l[-300:])
"
    )
  done
done
