#!/bin/bash
# Knob sweep on the default bench step, one bench run per argument; an argument is a
# space-separated list of MCB_* settings (bench.py maps MCB_<OPTION> to mcb200_set_option).
#   scripts/sweep.sh "MCB_FLY_BATCH=6" "MCB_FLY_BATCH=12" "MCB_STEP_BUDGET=64 MCB_TAIL=8192"
# The sweeps of round 1 (profiles/r01_sweep_*.txt, r01_wave0_order_ab.txt) were:
#   MCB_AGG_STEPS=0|2|4|6|12, MCB_BATCH=4|12|16, MCB_ORDER=0, MCB_WAVE0_BLOCKS=3|6, MCB_STEP_BUDGET=64|128|192,
#   MCB_FLY_BATCH=1|4|6|8|12|16, MCB_TAIL=8192|131072, MCB_WAVE0_EXACT=0|1 (x MCB_AGG_STEPS=6|16|40), MCB_WAVE0_ORDER=0
run() { env $1 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['waves_per_step'], d['clocks']['sm_mhz'])"; }
[ $# -eq 0 ] && set -- "X=defaults"
for s in "$@"; do run "$s"; done
