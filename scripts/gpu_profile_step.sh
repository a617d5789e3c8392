# Launch list of one bench step (device time per kernel, ncu-serialised) + full captures of the top kernels.
# Usage: gpurun -- 'bash scripts/gpu_profile_step.sh TAG'
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
tail -2 gpurun_out/ncu_l_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:'bootstrap_1d_poisson|bootstrap_1d_kernel' -s 4 -c 2 -o gpurun_out/prof_boot_$TAG $CMD > gpurun_out/ncu_b_$TAG.log 2>&1
tail -2 gpurun_out/ncu_b_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:'seg_moments_tile|regress_asl|fill_log' -s 3 -c 3 -o gpurun_out/prof_misc_$TAG $CMD > gpurun_out/ncu_m_$TAG.log 2>&1
tail -2 gpurun_out/ncu_m_$TAG.log
ls -la gpurun_out | tail -8
