mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 80 --csv --log-file gpurun_out/launches_s1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_s1.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hd=rows[h]; ik=hd.index('Kernel Name'); iv=hd.index('Metric Value')
for r in rows[h+1:]:
    if len(r)>iv: print(r[ik][:60].ljust(60), r[iv])
PY
