import csv, sys
from collections import OrderedDict
c=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/l_{c}.csv")) if len(r)>10 and r[0].isdigit()]
names=[(r[4].split("(")[0][:44], float(r[-1])) for r in rows]
tail=names[-31:]
agg=OrderedDict()
for n,t in tail: agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=t
print(c, len(names), {k:(v[0], round(v[1]/v[0]/1000,2)) for k,v in agg.items()}, "sum/6 =", round(sum(v[1] for v in agg.values())/6000,1))
