#!/bin/bash
# A/B tuning runs on the GPU box: tools/ab.sh "<name>|<env sizes>|<env assignments>" ...   -> gpurun_out/ab_<name>_<E>.log
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%|*}"; rest="${spec#*|}"; sizes="${rest%%|*}"; envs="${rest#*|}"
  for E in $sizes; do
    env $envs python bench.py --quick --steps 1400 --warmup 100 --n-envs $E > gpurun_out/ab_${name}_$E.log 2>&1
    echo "$name $E: $(tail -1 gpurun_out/ab_${name}_$E.log | cut -c1-260)"
  done
done
