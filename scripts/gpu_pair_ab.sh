# A/B of the pair bootstrap slot count (MM_PAIR_SLOTS) on the C3 ht_2d workload.  Usage: gpurun -- 'bash scripts/gpu_pair_ab.sh TAG'
TAG=${1:-pair}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair or ht_2d or 2d" 2>&1 | tail -5
for s in ${SLOTS:-1 2 3}; do
  MM_PAIR_SLOTS=$s timeout 600 python scripts/bench_ht2d.py > gpurun_out/ht2d_${TAG}_s$s.json 2> gpurun_out/ht2d_${TAG}_s$s.err
  echo "slots $s: $(tail -1 gpurun_out/ht2d_${TAG}_s$s.json | cut -c1-900)"
done
