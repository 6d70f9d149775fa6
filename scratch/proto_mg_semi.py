"""numpy preview of the semi-coarsening rule (coarsen a direction only while its cell size is within 1.5x of
the smallest): PCG iteration counts on non-cubic grids.  Uses proto_mg_general's pieces."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, proto_mg_general as G
from keff_cfd_b200 import structures as S
def coarsen_rule(L, h, first):
    nz,ny,nx = L.shape
    cand=[nz>1, ny>1, nx>1]
    hmin=min(hh for hh,c in zip(h,cand) if c)
    if first: f=[cand[0] and h[0] < 1.5*min(h[1],h[2]), True, True]
    else: f=[c and hh<1.5*hmin for hh,c in zip(h,cand)]
    return f
def coarsen(L, f):
    def reduce_all(a, skip):
        for ax,n in enumerate(a.shape):
            if ax!=skip and f[ax]: a=G.agg(a,ax,n)
        return a
    sc=lambda ax:0.5 if f[ax] else 1.0
    ccx = sc(2)*reduce_all(L.cx[:,:,1::2],2) if f[2] else reduce_all(L.cx,2)
    ccy = sc(1)*reduce_all(L.cy[:,1::2,:],1) if f[1] else reduce_all(L.cy,1)
    ccz = sc(0)*reduce_all(L.cz[1::2],0) if f[0] else reduce_all(L.cz,0)
    def r2(a):
        if f[0]: a=G.agg(a,0,a.shape[0])
        if f[1]: a=G.agg(a,1,a.shape[1])
        return a
    wl,wr=sc(2)*r2(L.wl),sc(2)*r2(L.wr)
    shape=tuple((n+1)//2 if ff else n for n,ff in zip(L.shape,f))
    ccx=ccx[:,:,:shape[2]-1]; ccy=ccy[:,:shape[1]-1,:]; ccz=ccz[:shape[0]-1]
    return G.Level(ccx,ccy,ccz,wl,wr,shape)
def run(shape, semi=True):
    solid=S.bernoulli_volume(shape,0.5,1234)
    L0=G.fine_level(solid); nz,ny,nx=shape
    b=np.zeros(shape); b[:,:,-1]=L0.wr
    x0=np.broadcast_to((np.arange(nx)+0.5)/nx,shape).copy()
    levels,flags=[L0],[]; h=[1/nz if nz>1 else 1e9,1/ny,1/nx]
    while np.prod(levels[-1].shape)>64:
        f=coarsen_rule(levels[-1],h,len(levels)==1) if semi else [n>1 for n in levels[-1].shape]
        levels.append(coarsen(levels[-1],f)); flags.append(f)
        h=[hh*2 if ff else hh for hh,ff in zip(h,f)]
    dense=G.dense_inverse(levels[-1])
    x,it=G.pcg(L0,b,x0,lambda r:G.vcycle(levels,flags,0,r,0.9,dense,True))
    print(shape,"semi" if semi else "full",[l.shape for l in levels][:4],"iters",it, flush=True)
if __name__ == "__main__":
    for shp in ((16,64,64),(64,64,16),(7,30,50),(20,33,64),(1,100,60),(1,60,100),(3,16,200),(32,32,32)):
        run(shp, True); run(shp, False)
