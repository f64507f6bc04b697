"""LINE-map culling study (lives under tests/ because it takes its exit rays from the oracle): hits, per-bin sphere-criterion
passes and lane-tests per escaping ray for the candidate tile shapes.  Run from the repository root."""
import sys
sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import numpy as np, pyoracle as O
n=60000
rec,st=O.trace(O.scene(theta_max=170.0),O.source(),n,seed=4357,prec=O.F32)
m=(rec['status']==1)&(rec['pos'][:,2]<-100)
L=rec['pos'][m].astype(np.float64); v=rec['dir'][m].astype(np.float64)
L=L[:1500]; v=v[:1500]
nt,npb=180,90; R=100.0; w=20.0
th=(np.arange(nt)+0.5)*0.5*np.pi/180; ph=(np.arange(npb)+0.5)*4*np.pi/180
P=np.stack([R*np.sin(th)[:,None]*np.cos(ph)[None,:], R*np.sin(th)[:,None]*np.sin(ph)[None,:], (-100-R*np.cos(th))[:,None]*np.ones((1,npb))],-1)  # nt,np,3
d=P-np.array([0,0,-100.0]); N=np.stack([-d[...,1],d[...,0],d[...,2]],-1); N/=np.linalg.norm(N,axis=-1,keepdims=True)
def dist2(C):   # C: (...,3) -> (rays, ...) squared distance to lines
    mvec=C[None]-L.reshape((-1,)+(1,)*(C.ndim-1)+(3,))
    mv=(mvec*v.reshape((-1,)+(1,)*(C.ndim-1)+(3,))).sum(-1)
    return (mvec**2).sum(-1)-mv**2
# exact hits
dot=(N[None]*v[:,None,None,:]).sum(-1)
dd=L[:,None,None,:]-P[None]
num=(dd*N[None]).sum(-1)
q=dot[...,None]*dd-num[...,None]*v[:,None,None,:]
hit=((q**2).sum(-1)<=w*w*dot**2)&(np.abs(dot)>=1e-10)
print('hits/ray',hit.sum()/len(L))
sph=dist2(P)<=w*w
print('per-bin sphere passes/ray',sph.sum()/len(L), 'hits covered', (hit&~sph).sum())
for tt,tp in [(32,1),(16,2),(8,4),(4,8),(16,1),(8,2),(4,4),(8,1),(4,2),(2,4)]:
    ntt,ntp=-(-nt//tt),-(-npb//tp)
    tot=0
    for a in range(ntt):
        for b in range(ntp):
            blk=P[a*tt:(a+1)*tt,b*tp:(b+1)*tp].reshape(-1,3)
            c=blk.mean(0); r=np.sqrt(((blk-c)**2).sum(-1).max())
            ps=dist2(c[None])[:,0]<=(w+r+0.5)**2
            tot+=ps.sum()*tt*tp
    print(f'tile {tt}x{tp}: lane-tests/ray {tot/len(L):.0f}  (tiles {ntt*ntp})')
