#!/bin/bash
# A/B of the rank-local global-InfoNCE implementations (symmetric CUDA-core tiles vs the tensor-core chain) per batch size
for b in 64 128 512; do for m in 1 0; do
  echo -n "B=$b prefer_sym=$m: "
  CFA_GLOBAL_PREFER_SYM=$m python bench.py --no-adamspd --no-cpu --batch $b --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
done; done
