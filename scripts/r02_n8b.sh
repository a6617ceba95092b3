# round 2: variants of the peer-memory merge at N GPUs
N=${1:-8}
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu --no-e2e --steps 4 --warmup 3"
for v in 2 1 0; do
MCB_EXCHANGE_PUSH=$v $TR > gpurun_out/r02_bench_n${N}_push$v.json 2> gpurun_out/r02_bench_n${N}_push$v.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n${N}_push$v.json").read().strip().splitlines()[-1])
    print("push=$v", {k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity")}, d["exchange"]["detail"])
except Exception as e:
    print("push=$v ERR", e); print(open("gpurun_out/r02_bench_n${N}_push$v.err").read()[-2500:])
PY
done
