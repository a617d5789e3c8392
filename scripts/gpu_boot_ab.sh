# A/B of the Poissonised bootstrap kernel variants: "VARIANT:SLOTS" pairs
# (MM_BOOT_VARIANT: 0 plain, 1 Philox-7 + direct log rows, 2 = 1 + 2^52 conversion; MM_BOOT_SLOTS: replicates per lane)
# Usage: gpurun -- 'COMBOS="1:1 1:2" bash scripts/gpu_boot_ab.sh TAG'
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log
for c in ${COMBOS:-0:1 1:1 1:2 1:3 2:2}; do
  v=${c%%:*}; sl=${c##*:}
  MM_BOOT_VARIANT=$v MM_BOOT_SLOTS=$sl timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_v${v}s$sl.json 2> gpurun_out/bench_${TAG}_v${v}s$sl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_v${v}s$sl.json"))
print("variant $v slots $sl ms/step", round(d["ms_per_step"], 2), "genes/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(x, 1) for k, x in d["stage_ms_per_step"].items()})
PY
done
