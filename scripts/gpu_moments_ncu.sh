# one full ncu capture of the moment kernel on the C2 matrix (run scripts/gpu_moments_tile.sh first: it writes /tmp/mom.py)
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/mom.py/,/^PY$/p' scripts/gpu_moments_tile.sh > /tmp/mk.sh; bash /tmp/mk.sh
export MM_MOMENTS_THREADS=${1:-640}
python /tmp/mom.py > gpurun_out/mom_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:seg_moments -s 10 -c 2 -o gpurun_out/prof_momstream python /tmp/mom.py > gpurun_out/ncu_momtile.log 2>&1
tail -3 gpurun_out/ncu_momtile.log
