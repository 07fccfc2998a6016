import sys
sys.path.insert(0,'/root/repo/tools')
from ncu_hot import kernels
path, pat, occ, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ks=[k for k in kernels(path) if pat in k["name"] and "Address" in k["hdr"]]
k=ks[occ]; h=k["hdr"]
src, ie, smp, tie = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
rows=k["rows"]
tot=sum(int(r[ie]) for r in rows); tots=sum(int(r[smp]) for r in rows)
for i in range(a,b):
    r=rows[i]; e=int(r[ie])
    print("%5d %5.2f%% s%5.2f%% %4.1f  %s" % (i, 100*e/tot, 100*int(r[smp])/tots, int(r[tie])/max(1,e), r[src].strip()[:100]))
