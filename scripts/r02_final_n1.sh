# round 2, last GPU call: the one-GPU paths touched since the last full suite run (sparse escape fetch, one-rank exchange), smoke(), default bench
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_zy_gpu_native_comm.py -m gpu -q --timeout 120 -x -k "sparse or one_rank or exchange_without" > gpurun_out/r02_final_n1_tests.log 2>&1
tail -2 gpurun_out/r02_final_n1_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --cpu-seconds 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","e2e","roofline","cpu_baseline","clocks","gpu_launches")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_n1_final.err").read()[-2500:])
PY
