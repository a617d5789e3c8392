# tensor-core block kernel: correctness first (short timeout: a wrong descriptor can hang the kernel), then timing
mkdir -p gpurun_out
timeout 300 python scripts/diag_block.py 2>&1 | grep -v Warn | tail -12
timeout 300 python -m pytest tests/test_gpu_block.py tests/test_gpu_parity.py -k "block or corr_matrix or 2d_moments" -m gpu -x -q 2>&1 | tail -8
timeout 600 python scripts/bench_block.py 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/bench_block.json
