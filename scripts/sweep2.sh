run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --packets ${PK:-16000000} --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['segments_per_s'], d['kernel_ms_per_step'])"; }
run MCB_WAVEFRONT=1 MCB_STEP_BUDGET=96
run MCB_WAVEFRONT=1 MCB_STEP_BUDGET=48
run MCB_WAVEFRONT=1 MCB_STEP_BUDGET=192
run MCB_WAVEFRONT=1 MCB_STEP_BUDGET=100000
run MCB_WAVEFRONT=0
PK=125000000 run MCB_WAVEFRONT=1 MCB_STEP_BUDGET=96
PK=125000000 run MCB_WAVEFRONT=0
