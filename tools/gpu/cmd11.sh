ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 30 -c 40 --csv --log-file gpurun_out/launches_c3.csv python tools/prof_one.py 3 > gpurun_out/ncu_c3.log 2>&1; echo rc=$?
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_c3.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+2:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value'); mi=H.index('Metric Name')
cur={}
order=[]
for r in data:
    key=(r[0], r[ki].split('(')[0].replace('<unnamed>::','').replace('void ',''))
    if key not in cur: cur[key]={}; order.append(key)
    cur[key][r[mi]]=float(r[vi].replace(',',''))
for k in order:
    d=cur[k]
    print(f"{k[1][:30]:30s} {d.get('gpu__time_duration.sum',0)/1000:9.1f} us  rd {d.get('dram__bytes_read.sum',0)/1e6:8.1f} MB wr {d.get('dram__bytes_write.sum',0)/1e6:8.1f} MB")
PY
