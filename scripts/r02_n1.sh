# round 2, one GPU: the whole -m gpu suite, then A/B of the FLY kernel variants on the headline workload
nvidia-smi -L
timeout 700 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r02_n1_tests.log 2>&1
tail -8 gpurun_out/r02_n1_tests.log
B="timeout 240 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/r02_ab_new.json 2> gpurun_out/r02_ab_new.err
MCB_INLINE_ESCAPE=0 $B > gpurun_out/r02_ab_noesc.json 2> gpurun_out/r02_ab_noesc.err
MCB_FLY_STEPS=2 $B > gpurun_out/r02_ab_steps2.json 2> gpurun_out/r02_ab_steps2.err
MCB_FLY_STEPS=4 $B > gpurun_out/r02_ab_steps4.json 2> gpurun_out/r02_ab_steps4.err
MCB200_LIB=$PWD/mocassin_b200/ab/libmcb_base.so $B > gpurun_out/r02_ab_base.json 2> gpurun_out/r02_ab_base.err
for f in new noesc steps2 steps4 base; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_ab_$f.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","kernel_ms_per_step","gpu_launches")}, d["roofline"]["frac"], (d.get("access_roofline") or {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_ab_$f.err").read()[-1500:])
PY
done
