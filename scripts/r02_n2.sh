# round 2: two-GPU evidence -- N-rank == 1-rank tests, then the exchange variants timed at N=2
nvidia-smi -L
timeout 420 python -m pytest tests/test_gpu_multi.py tests/test_zy_gpu_native_comm.py -m gpu -q -rA --timeout 200 -x > gpurun_out/r02_n2_tests.log 2>&1
tail -25 gpurun_out/r02_n2_tests.log
TR="timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu"
$TR > gpurun_out/r02_bench_n2_native_rs.json 2> gpurun_out/r02_bench_n2_native_rs.err
MCB_EXCHANGE_ALLREDUCE=1 $TR > gpurun_out/r02_bench_n2_native_ar.json 2> gpurun_out/r02_bench_n2_native_ar.err
MCB_NATIVE_COMM=0 $TR > gpurun_out/r02_bench_n2_torch.json 2> gpurun_out/r02_bench_n2_torch.err
for f in native_rs native_ar torch; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n2_$f.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved","exchange")})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_n2_$f.err").read()[-1500:])
PY
done
