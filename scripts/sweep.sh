run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 2 --packets 16000000 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['segments_per_s'], d['kernel_ms_per_step'])"; }
run MCB_AGG_STEPS=6
run MCB_AGG_STEPS=0
run MCB_AGG_STEPS=2
run MCB_AGG_STEPS=0 MCB_BATCH=4
run MCB_AGG_STEPS=0 MCB_BATCH=12
run MCB_AGG_STEPS=0 MCB_BATCH=16
run MCB_AGG_STEPS=0 MCB_ORDER=0
