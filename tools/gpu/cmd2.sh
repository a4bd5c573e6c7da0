ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches10.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo rc=$?
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches10.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+2:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value')
# last full step: find last k_encode with big grid
names=[(r[ki].split('(')[0].replace('<unnamed>::','').replace('void ',''), float(r[vi].replace(',',''))) for r in data]
idx=[i for i,(n,v) in enumerate(names) if n.startswith('k_encode')]
start=idx[-1] if idx else 0
# device-resident step = the one before the e2e steps; print the last 40 launches
for n,v in names[-45:]:
    print(f"{n[:40]:40s} {v/1000:9.1f} us")
PY
