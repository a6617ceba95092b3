# round 2: host/device timeline of one exchange + merge at N=2 (MCB200_TRACE_EXCHANGE=1, rank 0)
TR="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu --no-e2e --steps 3 --warmup 3"
MCB200_TRACE_EXCHANGE=1 $TR > gpurun_out/r02_bench_n2_trace.json 2> gpurun_out/r02_bench_n2_trace.err
grep "mcb200 trace" gpurun_out/r02_bench_n2_trace.err | tail -3 | tee gpurun_out/r02_exchange_trace_n2.txt
tail -c 600 gpurun_out/r02_bench_n2_trace.json
