#!/usr/bin/env python
"""Long run of BASELINE config 1 (periodic channel Re_b = 5640, 64^3, dynamic Smagorinsky) on the C/OpenMP restatement of the
reference (oracle/c): the flow transitions from the deterministic Poiseuille + vortex-pair start and the friction Reynolds
number settles near the DNS value (180; a 64^3 LES sits 5-10 % below).  Physics sanity of the oracle, ~8 minutes on 8 cores;
profiles/r2t_cpu_les_channel64_dsmag.log is its output.  TEST INFRASTRUCTURE (runs the oracle, not the product)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle.param as op
from oracle.cport import CSim
from oracle.initgrid import initgrid
kw = dict(ng=(64, 64, 64), sgstype="dsmag")
d=op.deck_channel(**kw)
c=CSim(d)
dzc, dzf, zc, zf = initgrid(d.gtype, 64, d.gr, d.l[2])
nu=d.visc; t=0.; t0=time.time()
def retau():
    u=c.f["u"]
    tw_lo=nu*(u[1:-1,1:-1,1]-u[1:-1,1:-1,0]).mean()/dzc[0]
    tw_hi=nu*(u[1:-1,1:-1,-2]-u[1:-1,1:-1,-1]).mean()/dzc[64]
    tw=0.5*(tw_lo+tw_hi)
    return np.sqrt(abs(tw))*0.5/nu, tw_lo, tw_hi
for it in range(1,30001):
    c.step(icheck=10); t+=c.dt
    if it%500==0:
        r,a,b=retau()
        ke=0.5*float((c.f["v"][1:-1,1:-1,1:-1]**2).mean()+(c.f["w"][1:-1,1:-1,1:-1]**2).mean())
        print("step %6d t %8.2f dt %.4f Re_tau %.1f  tw %.3e %.3e  <v2+w2>/2 %.3e  ub %.4f  (%.0f s)"%(it,t,c.dt,r,a,b,ke,float((c.f["u"][1:-1,1:-1,1:-1]*dzf[1:-1][None,None,:]).sum()/(64*64)), time.time()-t0), flush=True)
