#!/bin/bash
# sweep the FLY kernel's store/claim batch on the default bench step
for b in 1 4 6 8 12 16; do
  MCB_FLY_BATCH=$b python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fly_batch', $b, d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"
done
