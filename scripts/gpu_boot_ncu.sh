# ncu --set full capture of the Poissonised bootstrap kernel on one C2 gene tile (after a plain run has exited 0)
TAG=${1:-boot}
mkdir -p gpurun_out
python scripts/diag_boot.py > gpurun_out/diag_boot_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bootstrap_1d_poisson -c 1 -o gpurun_out/prof_$TAG python scripts/diag_boot.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/diag_boot_plain_$TAG.log; tail -3 gpurun_out/ncu_$TAG.log
