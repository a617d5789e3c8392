# GPU tests + one full bench line (plain, no profiler).  Usage: gpurun -- 'bash scripts/gpu_check.sh TAG'
TAG=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
