"""CPU experiment behind the column pre-sort of the block SVD: a real theta-prime of a Heisenberg
DMRG run (numpy oracle), its largest charge groups fed to the numpy model of the block-Jacobi iteration under
different preconditionings (QR of the transpose as is / after a column sort by norm / with column pivoting / + LQ).
   python profiles/r2/precond_experiment.py"""
import sys, time, copy
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'profiles', 'r2'))
import numpy as np, scipy.linalg as sl
import qtb_oracle as orc
from quantit_b200 import workloads as wl
import jacobi_model as jm

L, D = 24, 400
mpo=[orc.BT(**h) for h in wl.heisenberg_mpo(L)]
mps=[orc.BT(**p) for p in wl.random_mps(L, 4, L%2, seed=0)]
t0=time.time()
orc.dmrg(mps, mpo, 0, 1e-20, 0.0, D, 4, max_iter=3)
print('dmrg done', time.time()-t0, flush=True)
# theta' at the middle bond
env={}
env[-1],env[L]=orc.trivial_edges(mps,mpo)
oc=0
# after dmrg with even number of half sweeps oc is back at 0: move through left envs to the middle
for i in range(L//2-1):
    # canonical move: use two-site step to advance
    pass
# simpler: rebuild envs for oc=0 and run a forward half sweep capturing theta' at the middle
for i in range(L-1,0,-1):
    env[i]=orc.compute_right_env(mpo[i],mps[i],env[i+1])
h2=[orc.compute_2sites_hamil(mpo[i],mpo[i+1]) for i in range(L-1)]
cap=None
oc=0
for s in range(L//2):
    theta=orc.tensordot(mps[oc],mps[oc+1],[2],[0])
    E,theta2=orc.two_sites_update(theta,h2[oc],env[oc-1],env[oc+2])
    if oc==L//2-1:
        cap=theta2; break
    u,d,v=orc.svd_trunc(theta2,2,1e-20,4,D)
    nrm=np.sqrt(sum(float(np.sum(x*x)) for x in d.blocks.values()))
    for k in d.blocks: d.blocks[k]=d.blocks[k]/nrm
    mps[oc]=u; mps[oc+1]=orc.permute(orc.conj(orc.mul_bcast(v,d)),[2,0,1])
    env[oc]=orc.compute_left_env(mpo[oc],mps[oc],env[oc-1]); oc+=1
groups=orc.svd_groups(orc.reshape_split(cap,2))
groups=sorted(groups,key=lambda g:-min(g[0].shape))[:3]
for dense,rows,cols in groups:
    A=dense if dense.shape[0]>=dense.shape[1] else dense.T
    m,n=A.shape
    sv=np.linalg.svd(A,compute_uv=False)
    print('group',A.shape,'sv range %.1e..%.1e'%(sv[0],sv[-1]), 'col norm range %.1e..%.1e'%(np.linalg.norm(A,axis=0).max(),np.linalg.norm(A,axis=0).min()), flush=True)
    variants={}
    variants['none']=A
    R=np.linalg.qr(A)[1]; variants['qrT']=R.T.copy()
    o=np.argsort(-np.linalg.norm(A,axis=0)); R=np.linalg.qr(A[:,o])[1]; variants['sort+qrT']=R.T.copy()
    R=sl.qr(A,mode='economic',pivoting=True)[1]; variants['qrcp T']=R.T.copy()
    R1=np.linalg.qr(A)[1]; L2=np.linalg.qr(R1.T)[1].T; variants['qr+lq (L)']=L2.copy()   # R1 = L2 Q2 -> Jacobi on L2
    R1=sl.qr(A,mode='economic',pivoting=True)[1]; L2=np.linalg.qr(R1.T)[1].T; variants['qrcp+lq (L)']=L2.copy()
    o=np.argsort(-np.linalg.norm(A,axis=0)); R1=np.linalg.qr(A[:,o])[1]; L2=np.linalg.qr(R1.T)[1].T; variants['sort+qr+lq (L)']=L2.copy()
    del variants['none']; del variants['qrcp T']; del variants['qr+lq (L)']
    for name,B in variants.items():
        t0=time.time()
        ns,hist,visits,X=jm.block_jacobi(B,16,2,adaptive=True)
        s2=np.sort(np.linalg.norm(X,axis=0))[::-1]
        print('  %-14s sweeps %2d visits %5d err %.1e (%.0fs) %s'%(name,ns,sum(visits),np.abs(s2-sv).max()/sv[0],time.time()-t0,' '.join('%.0e'%h for h in hist)),flush=True)
