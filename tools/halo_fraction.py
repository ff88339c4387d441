"""CPU (numpy): remote elements a rank needs for one H.v under two static row partitions of the closed chain -- a contiguous LEX
slice, and a contiguous suffix-rank range of every sector of the split layout (hv_split_tables.h) -- as a fraction of D.
usage: halo_fraction.py m   (n = m, cut p = m // 2, W = 2, 4, 8).  Numbers quoted in DESIGN.md section 9."""
import numpy as np, sys
from math import comb
def gen(m,n):
    if m==1: return np.array([[n]],dtype=np.int8)
    blocks=[]
    for k in range(n,-1,-1):
        sub=gen(m-1,n-k)
        blocks.append(np.hstack([np.full((len(sub),1),k,dtype=np.int8),sub]))
    return np.vstack(blocks)
m=int(sys.argv[1]); n=m; p=m//2; s=m-p
f=np.zeros((m,n+2),dtype=np.int64)
for q in range(m-1):
    for R in range(1,n+2): f[q][R]=comb(R-1+m-1-q,m-1-q)
def rank(S):
    after=n-np.cumsum(S.astype(np.int64),axis=1)
    r=np.zeros(len(S),dtype=np.int64)
    for q in range(m-1): r+=f[q][after[:,q]]
    return r
def sufinfo(S):
    after=n-np.cumsum(S.astype(np.int64),axis=1)
    R=after[:,p-1]
    sr=np.zeros(len(S),dtype=np.int64)
    for q in range(p,m-1): sr+=f[q][after[:,q]]
    return R,sr
S=gen(m,n); D=len(S)
R,sr=sufinfo(S)
nS=np.array([comb(r+s-1,s-1) for r in range(n+1)])
for W in (2,4,8):
    # owner by suffix range within the sector (even split of sufrank range)
    own=(sr*W)//nS[R]
    # owner by contiguous LEX slice
    nloc=(D+W-1)//W
    own_lex=np.arange(D)//nloc
    res={}
    for name,ow in (("suffix-range",own),("LEX-slice",own_lex)):
        need=[set() for _ in range(W)]
        cnt=np.zeros(W)
        needmask=np.zeros((W,D),dtype=bool)
        for q in range(m):
            a,b=q,(q+1)%m
            for (src,dst) in ((a,b),(b,a)):
                ok=S[:,src]>0
                rows=np.nonzero(ok)[0]
                T=S[ok].copy(); T[:,src]-=1; T[:,dst]+=1
                t=rank(T)
                remote=ow[rows]!=ow[t]
                for r in range(W):
                    sel=remote&(ow[rows]==r)
                    needmask[r,t[sel]]=True
        fr=needmask.sum(axis=1)/D
        sizes=np.bincount(ow,minlength=W)/D
        res[name]=(fr,sizes)
        print(m,W,name,"remote elements needed / D per rank:",np.round(fr,3),"max",round(fr.max(),3),"own share",np.round(sizes,3))
