nvidia-smi -L
timeout 700 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r02_n1_tests.log 2>&1
tail -4 gpurun_out/r02_n1_tests.log
bash scripts/r02_ab.sh default prev
bash scripts/r02_decks.sh
timeout 300 python scripts/small_configs.py 1e7 > gpurun_out/r02_small_configs.jsonl 2> gpurun_out/r02_small_configs.err
cat gpurun_out/r02_small_configs.jsonl | cut -c1-260
