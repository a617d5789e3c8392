set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc
python bench.py --cells 3000 --genes 600 --num-boot 1000 --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 2500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3500 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
K='regex:(row_sums|seg_moments|unique_|bootstrap_1d|fill_log|wls_functional|regress_asl|pair_products)'
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:bootstrap_1d -s 1 -c 1 -o gpurun_out/prof_boot python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ncu --set full --clock-control none --import-source on -k 'regex:(seg_moments_warp|unique_warp|csr_row_sums)' -s 2 -c 4 -o gpurun_out/prof_hbm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
ls -la gpurun_out
