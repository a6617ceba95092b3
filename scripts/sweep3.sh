#!/bin/bash
run() { env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['waves_per_step'])"; }
run MCB_AGG_STEPS=0
run MCB_AGG_STEPS=4
run MCB_AGG_STEPS=12
run MCB_WAVE0_EXACT=1 MCB_AGG_STEPS=6
run MCB_WAVE0_EXACT=1 MCB_AGG_STEPS=16
run MCB_WAVE0_EXACT=1 MCB_AGG_STEPS=40
