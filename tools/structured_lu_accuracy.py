#!/usr/bin/env python
"""Componentwise accuracy of the linear solves along a trajectory (CPU study, not product): Newton matrices assembled by
the oracle at every <stride>-th accepted step, solved by LAPACK (natural order, partial pivoting), by the structured
elimination with ordinary partial pivoting inside the light-particle border, without any interchange, with the
threshold rule the kernel uses (diagonal kept unless tau * max > |diagonal|), and by an interchange-free elimination
in natural order -- each compared with an 80-bit partial-pivoting solve.  Reported: max over species with Y > 1e-20 of
|dx - dx_ref| / Y.    usage: tools/structured_lu_accuracy.py xrb304 100 > profiles/r02_structured_lu_accuracy.txt"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import xnet_b200 as X
from oracle import xnet_oracle as O
from conftest import DATA, PROBLEMS, control_text
from test_slu_cpu import Slu
name=sys.argv[1]; stride=int(sys.argv[2])
txt = control_text(*PROBLEMS[name])
op = O.Problem.from_control(txt, DATA)
ctl = X.read_controls(txt)
net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
names=[s.strip() for s in net.names()]
h = op.run(nzone=1, hist_cap=6000)
slu = Slu(net, 0); P=slu.perm; nb=slu.nb; n=slu.n; nh=n-nb
def ge(A,b,dt,piv_from=None,tau=None):
    """GE; rows >= piv_from use partial pivoting (threshold tau keeps the diagonal when |d| >= tau*max)"""
    A=A.astype(dt).copy(); x=b.astype(dt).copy(); n=len(b)
    for k in range(n):
        if piv_from is not None and k>=piv_from:
            col=np.abs(A[k:,k]); p=k+int(np.argmax(col))
            if tau is not None and col[0]>=tau*col[p-k]: p=k
            if p!=k: A[[k,p]]=A[[p,k]]; x[[k,p]]=x[[p,k]]
        l=A[k+1:,k]/A[k,k]
        A[k+1:,k+1:]-=np.outer(l,A[k,k+1:]); x[k+1:]-=l*x[k]
    for k in range(n-1,-1,-1):
        x[k]=(x[k]-A[k,k+1:]@x[k+1:])/A[k,k]
    return x
N=int(h['hist_n'][0])
worst={}
for k in range(2,N-1,stride):
    rdt = 1.0 / h['hist_tdel'][0, k]; y = h['hist_y'][0, k-1]
    cs, yd, jac, ye = op.eval_point(h['hist_t9'][0, k], h['hist_rho'][0, k], y, rdt, -1.0)
    xr = ge(jac, yd, np.longdouble, 0)
    ynew = y + xr.astype(float); m = ynew>1e-20
    Ap=jac[np.ix_(P,P)]; bp=yd[P]
    res={}
    res['dense_pp']=np.linalg.solve(jac,yd)
    res['slu_borderpp']=ge(Ap,bp,np.float64,nh)[slu.pos]
    res['slu_nopiv']=ge(Ap,bp,np.float64,None)[slu.pos]
    res['slu_border_tau.01']=ge(Ap,bp,np.float64,nh,0.01)[slu.pos]
    res['nat_nopiv']=ge(jac,yd,np.float64,None)
    line='%5d t9 %.3f'%(k,h['hist_t9'][0,k])
    for key,x in res.items():
        e=np.abs(x-xr).astype(float)/np.maximum(ynew,1e-300); e=np.where(m,e,0)
        line+='  %s %.1e(%s)'%(key,e.max(),names[int(np.argmax(e))])
        if e.max()>worst.get(key,(0,))[0]: worst[key]=(e.max(),k,names[int(np.argmax(e))])
    print(line)
print('WORST',worst)
