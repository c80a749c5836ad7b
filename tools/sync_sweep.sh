#!/bin/bash
# CTA lockstep tuning sweep (B2H_SYNC_MODE) for the step kernel
for m in 0 1 2; do for e in 4096 16384; do
  B2H_SYNC_MODE=$m python bench.py --steps 60 --warmup 10 --n-envs $e --no-cpu-baseline > gpurun_out/sync_${m}_${e}.log 2>&1
  echo "mode $m envs $e rc=$? $(tail -1 gpurun_out/sync_${m}_${e}.log | cut -c1-140)"
done; done
