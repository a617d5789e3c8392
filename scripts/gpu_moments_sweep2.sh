head -25 scripts/gpu_moments_sweep.sh | sed -n '/^cat > \/tmp\/mom.py/,/^PY$/p' > /tmp/mk.sh; bash /tmp/mk.sh
for mode in flat noflat; do
  if [ $mode = noflat ]; then export MM_MOMENTS_NOFLAT=1 MM_MOMENTS_NOSMEM=1; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:seg_moments -s 6 -c 3 --csv --log-file gpurun_out/mom_$mode.csv python /tmp/mom.py > /dev/null 2>&1
  python - <<PY
import csv
lines=[l for l in open('gpurun_out/mom_$mode.csv') if not l.startswith('==')]
rows={}
for r in csv.DictReader(lines):
    rows.setdefault(r['ID'],{'k':r['Kernel Name'][:40]})[r['Metric Name'].split('.')[0][-28:]]=r['Metric Value']
for v in rows.values(): print('$mode', v)
PY
done
