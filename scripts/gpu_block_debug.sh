# block GEMM timing experiments: MM_BLOCK_DEBUG values given as arguments (default 0)
for d in ${@:-0}; do
echo "MM_BLOCK_DEBUG=$d"; MM_BLOCK_DEBUG=$d timeout 120 python scripts/bench_block.py 2>/dev/null | cut -c60-200
done
