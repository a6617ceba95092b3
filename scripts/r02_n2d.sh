# round 2: two-GPU check after reordering the exchange: packed-mode tests, then the bench with the exchange timeline
timeout 200 python -m pytest tests/test_zy_gpu_native_comm.py -m gpu -q -rA --timeout 150 -x -k "p2p] or p2p+slabs or checksum or fetch_cells" > gpurun_out/r02_n2e_tests.log 2>&1
tail -2 gpurun_out/r02_n2e_tests.log
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --no-e2e --steps 5 --warmup 3"
MCB200_TRACE_EXCHANGE=1 $TR > gpurun_out/r02_bench_n2_flagstage.json 2> gpurun_out/r02_bench_n2_flagstage.err
grep "mcb200 trace" gpurun_out/r02_bench_n2_flagstage.err | tail -2 | tee gpurun_out/r02_exchange_trace_n2_flagstage.txt
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n2_flagstage.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved")}, d["exchange"]["detail"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_n2_flagstage.err").read()[-2500:])
PY
