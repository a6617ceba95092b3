# round 2: two-GPU check of the packed push -- native-exchange tests, bench with the packed push (default) and with the 64-bit push
nvidia-smi -L
timeout 300 python -m pytest tests/test_zy_gpu_native_comm.py -m gpu -q -rA --timeout 200 -x > gpurun_out/r02_n2b_tests.log 2>&1
tail -3 gpurun_out/r02_n2b_tests.log
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu"
$TR --steps 6 --warmup 3 > gpurun_out/r02_bench_n2_packed.json 2> gpurun_out/r02_bench_n2_packed.err
if [ -n "$UNPACKED" ]; then MCB_EXCHANGE_PACK=0 $TR --steps 5 --warmup 3 --no-e2e > gpurun_out/r02_bench_n2_unpacked.json 2> gpurun_out/r02_bench_n2_unpacked.err; fi
for f in n2_packed ${UNPACKED:+n2_unpacked}; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved","exchange","e2e")})
except Exception as e:
    print("$f ERR", e); print(open("gpurun_out/r02_bench_$f.err").read()[-2500:])
PY
done
