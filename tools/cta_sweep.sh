#!/bin/bash
# lockstep-group size sweep (B2H_WARPS_PER_CTA) for the step kernel
for w in 14 7 4 3 2; do for e in 4096 16384; do
  B2H_WARPS_PER_CTA=$w python bench.py --steps 60 --warmup 10 --n-envs $e --no-cpu-baseline > gpurun_out/cta_${w}_${e}.log 2>&1
  echo "warps $w envs $e rc=$? $(tail -1 gpurun_out/cta_${w}_${e}.log | cut -c1-120)"
done; done
