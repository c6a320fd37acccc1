#!/bin/bash
# DRAM bytes and duration of the LLM GEMM shapes for several rasterisation group sizes (A3_GEMM_GROUP_M).
cd "$(dirname "$0")/.."
for g in 4 8 16 32; do
  echo "== group_m $g"
  A3_GEMM_GROUP_M=$g python scripts/gpu_llm_gemm_bench.py 2>&1 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | tr '\n' ';'; echo
  A3_GEMM_GROUP_M=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm -s 60 -c 385 --csv python scripts/gpu_llm_gemm_bench.py 2>/dev/null \
    | python -c "
import csv,sys,collections
rows=[r for r in csv.reader(l for l in sys.stdin if l.startswith('\"'))]
h=rows[0]; ix={n:i for i,n in enumerate(h)}
per=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    v*={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
    per.setdefault(int(r[ix['ID']]),[0,0])[0 if 'read' in r[ix['Metric Name']] else 1]+=v
ids=sorted(per)
# 35 launches per shape (5 warm-up + 30 timed), 11 shapes in order
for k in range(0,len(ids),35):
    grp=[per[i] for i in ids[k:k+35]]
    print('shape',k//35,'read MB %.0f write MB %.0f'%(sum(g[0] for g in grp)/len(grp)/1e6,sum(g[1] for g in grp)/len(grp)/1e6))
"
done
