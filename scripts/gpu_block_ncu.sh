mkdir -p gpurun_out
python scripts/bench_block.py --reps 2 > gpurun_out/block_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'block_gemm|block_panels' -s 40 -c 4 -o gpurun_out/prof_block python scripts/bench_block.py --reps 2 > gpurun_out/ncu_block.log 2>&1
tail -2 gpurun_out/ncu_block.log
