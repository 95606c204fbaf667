B="python bench.py --config cfg4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subconfigs --graph off"
$B > gpurun_out/plain_cfg4.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 150 --csv --log-file gpurun_out/cfg4_l.csv $B > gpurun_out/ncu_l_cfg4.log 2>&1; echo rc=$?
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/cfg4_l.csv') if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    a=agg.setdefault((r[ki][:60], r[mi]), [0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
for k,a in agg.items(): print(k, a[0], round(a[1]/a[0],1))
PY
