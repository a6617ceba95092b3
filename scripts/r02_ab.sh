# A/B of library builds on the headline workload: scripts/r02_ab.sh name[:ENV=VAL,...] ...   (mocassin_b200/ab/libmcb_<name>.so)
B="timeout 240 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu"
for spec in "$@"; do
  n=${spec%%:*}; envs=""
  if [ "$spec" != "$n" ]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=$PWD/mocassin_b200/ab/libmcb_$n.so
  [ "$n" = "default" ] && lib=$PWD/mocassin_b200/libmocassin_b200.so
  tag=$(echo "$spec" | tr ':,=' '___')
  env MCB200_LIB=$lib $envs $B > gpurun_out/r02_ab_$tag.json 2> gpurun_out/r02_ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_ab_$tag.json").read().strip().splitlines()[-1])
    print("$spec", {k:round(d.get(k),2) for k in ("ms_per_step","kernel_ms_per_step")}, d.get("gpu_launches"))
except Exception as e:
    print("$spec", "ERR", e); print(open("gpurun_out/r02_ab_$tag.err").read()[-800:])
PY
done
