# ncu --set full of the dense-block kernels (panels, scaling, persistent tcgen05 GEMM) of one group of the configs[2] block.
# Under ncu kernels are serialised, so the panels of the next group no longer overlap the GEMM.
mkdir -p gpurun_out
python scripts/bench_block.py --reps 2 > gpurun_out/block_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'block_gemm|block_panels|block_scaling' -s 40 -c 4 -f -o gpurun_out/r02_prof_block python scripts/bench_block.py --reps 2 > gpurun_out/ncu_block.log 2>&1
tail -2 gpurun_out/ncu_block.log
