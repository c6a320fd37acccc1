#!/bin/bash
# DRAM bytes and duration of the LLM GEMM shapes (scripts/gpu_llm_gemm_bench.py) under ncu for
#   A3_GEMM_GROUP_M (rasterisation group size) and A3_GEMM_HINT (L2 cache policies of the operand loads).
#   usage: scripts/gpu_gemm_traffic.sh "group_m=16 hint=0" "group_m=16 hint=al_bf" ...
cd "$(dirname "$0")/.."
[ $# -eq 0 ] && set -- "group_m=16 hint=0" "group_m=16 hint=al_bf" "group_m=16 hint=af_bl" "group_m=8 hint=0" "group_m=8 hint=al_bf"
for cfg in "$@"; do
  g=$(echo "$cfg" | sed 's/.*group_m=\([0-9]*\).*/\1/'); h=$(echo "$cfg" | sed 's/.*hint=\([a-z0-9_]*\).*/\1/')
  echo "== group_m $g hint $h"
  A3_GEMM_GROUP_M=$g A3_GEMM_HINT=$h python scripts/gpu_llm_gemm_bench.py 2>&1 | awk '{printf "%s %s %s %s us %s TF; ", $1,$2,$3,$(NF-3),$(NF-1)}'; echo
  A3_GEMM_GROUP_M=$g A3_GEMM_HINT=$h ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm2 -c 385 --csv python scripts/gpu_llm_gemm_bench.py 2>/dev/null \
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
M,D,F=6680,4096,11008
e=2e-6
names=['qkv plain','qkv +rope','wo plain','wo +res','w13 plain','w13 fused swiglu','w2 plain','w2 +res','w2 dgrad plain','w2 dgrad fused dswiglu','wgrad w13']
alg=[(M*D+3*D*D, M*3*D),(M*D+3*D*D, M*3*D),(M*D+D*D, M*D),(2*M*D+D*D, M*D),(M*D+2*F*D, M*2*F),(M*D+2*F*D, M*3*F),(M*F+D*F, M*D),(M*F+D*F+M*D, M*D),(M*D+D*F, M*F),(M*D+D*F+M*2*F, M*2*F),(M*2*F+M*D, 2*F*D)]
tr=ta=0
for k in range(0,len(ids),35):
    grp=[per[i] for i in ids[k:k+35]][5:]
    i=k//35
    rd=sum(g[0] for g in grp)/len(grp)/1e6; wr=sum(g[1] for g in grp)/len(grp)/1e6
    ar,aw=alg[i][0]*e,alg[i][1]*e
    tr+=rd+wr; ta+=ar+aw
    print('   %-24s read %6.0f MB (algorithmic %5.0f)  write %5.0f MB (%5.0f)  total x%.2f'%(names[i],rd,ar,wr,aw,(rd+wr)/(ar+aw)))
print('   all shapes: %.0f MB vs algorithmic %.0f MB = x%.2f'%(tr,ta,tr/ta))
"
done
