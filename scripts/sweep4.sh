#!/bin/bash
run() { env "$@" python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks'])"; }
run MCB_WAVE0_EXACT=0
run MCB_WAVE0_EXACT=1
run MCB_WAVE0_EXACT=0
run MCB_WAVE0_EXACT=1
run MCB_WAVE0_ORDER=0
