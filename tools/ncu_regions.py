import sys, csv
sys.path.insert(0,'/root/repo/tools')
from ncu_hot import kernels
path, pat, occ = sys.argv[1], sys.argv[2], int(sys.argv[3])
ks=[k for k in kernels(path) if pat in k["name"] and "Address" in k["hdr"]]
k=ks[occ]; h=k["hdr"]
src, ie, smp, tie = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
rows=k["rows"]
tot=sum(int(r[ie]) for r in rows); tots=sum(int(r[smp]) for r in rows)
print(k["name"][:60], "total warp-inst", tot)
# segment into runs of similar exec count & thr
seg=[]; 
for i,r in enumerate(rows):
    e=int(r[ie]); t=int(r[tie])
    seg.append((i,e,t,int(r[smp])))
# print cumulative per block of N lines
B=int(sys.argv[4]) if len(sys.argv)>4 else 50
for b in range(0,len(rows),B):
    e=sum(s[1] for s in seg[b:b+B]); t=sum(s[2] for s in seg[b:b+B]); sm=sum(s[3] for s in seg[b:b+B])
    if e/tot>0.003: print("%5d-%5d exec %5.1f%% samples %5.1f%% thr %4.1f" % (b,b+B-1,100*e/tot,100*sm/tots,t/max(1,e)))
