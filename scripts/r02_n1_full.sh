# round 2, one GPU: new tests, converged decks, ncu evidence, A/B, the default bench line
nvidia-smi -L
timeout 500 python -m pytest tests/test_zw_gpu_multideck.py tests/test_zx_gpu_gas_deck.py tests/test_zz_gpu_deck.py tests/test_gpu_reference_golden.py -m gpu -q --timeout 300 > gpurun_out/r02_n1_newtests.log 2>&1
tail -6 gpurun_out/r02_n1_newtests.log
bash scripts/r02_decks.sh
bash scripts/r02_ab.sh default occ3:MCB_BLOCKS_PER_SM=3
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_traffic.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_step.log 2>&1
tail -2 gpurun_out/r02_ncu_step.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wf_fly -s 10 -c 1 -o gpurun_out/r02_wf_fly_wave0 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_fly.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 300 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 1500 gpurun_out/r02_bench_n1.json
