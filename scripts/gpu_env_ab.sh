# Generic A/B of bench.py under sets of environment overrides.
# Usage: gpurun -- 'bash scripts/gpu_env_ab.sh TAG "A=1 B=2" "A=2" ...'   (each argument after TAG is one configuration)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_$i.json"))
print("[$cfg] ms/step", round(d["ms_per_step"], 2), "genes/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 1) for k, x in d["stage_ms_per_step"].items()})
PY
done
