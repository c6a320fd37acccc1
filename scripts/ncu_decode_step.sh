#!/bin/bash
# ncu launch list of ONE eager decode step at batch $1 (default 16) in the tree $2 (default .): per-kernel totals.
B=${1:-16}; T=${2:-.}
cd "$(dirname "$0")/.."; cd $T
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv python scripts/gpu_decode_ncu.py $B 1 2>/dev/null | python -c "
import csv,sys,collections
rows=[r for r in csv.reader(l for l in sys.stdin if l.startswith('\"'))]
h=rows[0]; ix={n:i for i,n in enumerate(h)}
t=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    k=r[ix['Kernel Name']][:60]+' grid '+r[ix['Grid Size']]
    d=t.setdefault(k,[0,0.0,0.0])
    if 'time' in r[ix['Metric Name']]:
        d[0]+=1; d[1]+=v*{'ns':1e-3,'us':1,'ms':1e3,'nsecond':1e-3,'usecond':1,'msecond':1e3}.get(u,1)
    else:
        d[2]+=v*{'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u,1)
tot=sum(d[1] for d in t.values())
for k,d in sorted(t.items(), key=lambda x:-x[1][1])[:14]:
    print('%-90s n=%4d total %9.1f us  avg %8.1f us  read %9.1f MB'%(k,d[0],d[1],d[1]/d[0],d[2]))
print('step total %.1f us'%tot)
"
