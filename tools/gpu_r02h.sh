#!/bin/bash
O=gpurun_out/r02h
mkdir -p $O
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" > $O/status.txt
tail -5 $O/bench_n1.err
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"cddt|RadixSort|Onesweep|Histogram" --csv --log-file $O/c3_launches.csv python tools/prof_r02.py c3_cddt 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02h/c3_launches.csv')))
# find header
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[start:]:
    if len(r)<=iv: continue
    d.setdefault((int(r[iid]),r[ik][:60]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items()): print(i,k,m)
PY
cat $O/status.txt
