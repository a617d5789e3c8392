for d in 0; do
MM_BLOCK_DEBUG=$d MM_BLOCK_CLUSTER=1 timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:block_gemm -c 3 --csv --log-file gpurun_out/blk_dbg$d.csv python scripts/bench_block.py --reps 1 > /dev/null 2>&1
echo debug $d; grep -v "^==" gpurun_out/blk_dbg$d.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -6
done
