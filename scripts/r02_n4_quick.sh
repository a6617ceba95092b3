# round 2: packed peer-memory merge at N=4 (short: parity check + 3 timed steps)
TR="timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --no-cpu --no-e2e --steps 3 --warmup 3"
MCB200_TRACE_EXCHANGE=1 $TR > gpurun_out/r02_bench_n4_final.json 2> gpurun_out/r02_bench_n4_final.err
grep "mcb200 trace" gpurun_out/r02_bench_n4_final.err | tail -1 | tee gpurun_out/r02_exchange_trace_n4_final.txt
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n4_final.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved")}, d["exchange"]["detail"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_n4_final.err").read()[-2500:])
PY
