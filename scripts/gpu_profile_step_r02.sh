# Round-2 evidence: launch list of one bench step (library kernels only) + full captures of the bootstrap kernel
# inside the step and of the moment kernel (stream + edge) on the bench matrix.
# Usage: gpurun -- 'bash scripts/gpu_profile_step_r02.sh'
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-shapes"
K='regex:seg_|bootstrap|boot_prepare|unique|fill_|wls_|regress|gev|relayout|csr_row|validate|resample|pair_|block_'
$CMD > gpurun_out/plain_r02.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_r02.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/ncu_l_r02.log 2>&1
tail -1 gpurun_out/ncu_l_r02.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:bootstrap_1d_poisson -s 4 -c 1 -o gpurun_out/r02_prof_boot $CMD > gpurun_out/ncu_b_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'seg_moments_stream|seg_moments_edge' -s 6 -c 2 -o gpurun_out/r02_prof_moments $CMD > gpurun_out/ncu_m_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'relayout_tile|relayout_rowscan|csr_row_sums' -c 4 -o gpurun_out/r02_prof_ingest $CMD > gpurun_out/ncu_i_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep
