# round 2: the bench at N GPUs (argument: N): fused peer-memory merge (default, with e2e), then the NCCL reduce-scatter path
N=${1:-8}
nvidia-smi -L | head -8
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu"
$TR --steps 8 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
if [ -n "$NCCL_LEG" ]; then MCB_EXCHANGE_P2P=0 $TR --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/r02_bench_n${N}_nccl.err; fi
for f in n$N ${NCCL_LEG:+n${N}_nccl}; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved","exchange","e2e")})
except Exception as e:
    print("$f ERR", e); print(open("gpurun_out/r02_bench_$f.err").read()[-2500:])
PY
done
