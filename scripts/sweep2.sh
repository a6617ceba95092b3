#!/bin/bash
run() { env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['waves_per_step'])"; }
run MCB_WAVE0_BLOCKS=3
run MCB_WAVE0_BLOCKS=6
run MCB_STEP_BUDGET=64
run MCB_STEP_BUDGET=128
run MCB_STEP_BUDGET=192
run MCB_FLY_BATCH=12
run MCB_FLY_BATCH=6
run MCB_TAIL=8192
run MCB_TAIL=131072
