#!/bin/bash
# same-box A/B of two builds: scripts/_build/ab/lib_a.so (baseline) vs the in-tree library
run() { env "$@" python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['kernel_ms_per_step'])"; }
for i in 1 2; do
run MCB200_LIB=$PWD/scripts/_build/ab/lib_a.so
run X=in-tree
done
