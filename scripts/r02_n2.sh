# round 2: two-GPU evidence -- N-rank == 1-rank tests, then the bench at N=2
nvidia-smi -L
timeout 420 python -m pytest tests/test_gpu_multi.py tests/test_zy_gpu_native_comm.py -m gpu -q -rA --timeout 200 -x > gpurun_out/r02_n2_tests.log 2>&1
tail -4 gpurun_out/r02_n2_tests.log
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu"
$TR > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
    print("n2", {k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved","exchange","e2e")})
except Exception as e:
    print("n2 ERR", e); print(open("gpurun_out/r02_bench_n2.err").read()[-2500:])
PY
