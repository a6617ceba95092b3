# round 2: two-GPU check of the packed push with zero-block skipping: the packed-mode tests, then the bench (16 and 4 push CTAs per SM)
nvidia-smi -L
timeout 200 python -m pytest tests/test_zy_gpu_native_comm.py -m gpu -q -rA --timeout 150 -x -k "p2p] or p2p+slabs or checksum or fetch_cells" > gpurun_out/r02_n2c_tests.log 2>&1
tail -3 gpurun_out/r02_n2c_tests.log
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --no-e2e --steps 5 --warmup 3"
$TR > gpurun_out/r02_bench_n2_zskip.json 2> gpurun_out/r02_bench_n2_zskip.err
MCB_EXCHANGE_PUSH_BLOCKS=4 $TR > gpurun_out/r02_bench_n2_zskip_b4.json 2> gpurun_out/r02_bench_n2_zskip_b4.err
for f in n2_zskip n2_zskip_b4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","nrank_parity","packets_conserved")}, d["exchange"]["detail"])
except Exception as e:
    print("$f ERR", e); print(open("gpurun_out/r02_bench_$f.err").read()[-2500:])
PY
done
