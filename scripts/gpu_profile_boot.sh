ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum --clock-control none -k regex:bootstrap_1d --csv --log-file gpurun_out/boot_launches.csv bash scripts/gpu_sampler_compare.sh > gpurun_out/ncu_p.log 2>&1; tail -2 gpurun_out/ncu_p.log
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/boot_launches.csv') if not l.startswith('==')]
rd=csv.DictReader(lines)
rows={}
for r in rd:
    rows.setdefault(r['ID'],{'k':r['Kernel Name'].split('(')[0]})[r['Metric Name']]=r['Metric Value']
for i,(k,v) in enumerate(rows.items()):
    print(i, v)
PY
