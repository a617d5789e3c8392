mkdir -p gpurun_out
python scripts/diag_boot.py > gpurun_out/diag_boot_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bootstrap_1d_poisson -c 2 -o gpurun_out/prof_bootfast python scripts/diag_boot.py > gpurun_out/ncu_bootfast.log 2>&1
tail -3 gpurun_out/ncu_bootfast.log
