set -x
python scripts/perf_gcfm.py 12500 100000 2>&1 | tail -3
python scripts/perf_gcfm.py 12500 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep|prepare|setup|scan|scatter|noise|exit_compact|state_pack" -c 120 --csv --log-file gpurun_out/r2_gcfm_launches.csv python scripts/perf_gcfm.py 12500 > gpurun_out/ncu_gcfm.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2_gcfm_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ki=h.index('Kernel Name'); mi=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    try: v=float(r[mi].replace(',',''))
    except: continue
    a=agg.setdefault(r[ki].split('(')[0],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{k:30s} n={n:3d} avg {t/n/1e3:8.1f} us")
PY
