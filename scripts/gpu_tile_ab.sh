# A/B of the tile plan and block order of the bootstrap: "LPT:WORKSPACE_GB:BALANCE" triples
# Usage: gpurun -- 'COMBOS="0:6:0 1:6:0" bash scripts/gpu_tile_ab.sh TAG'
TAG=${1:-tile}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -3 gpurun_out/pytest_$TAG.log
for c in ${COMBOS:-0:6:0 1:6:0 1:12:0 1:24:0 1:6:1}; do
  IFS=: read lpt ws bal <<< "$c"
  MM_BOOT_LPT=$lpt MM_WORKSPACE_GB=$ws MM_TILE_BALANCE=$bal timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$c.json 2> gpurun_out/bench_${TAG}_$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_$c.json"))
print("lpt $lpt ws $ws balance $bal ms/step", round(d["ms_per_step"], 2), "genes/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 1) for k, x in d["stage_ms_per_step"].items()})
PY
done
