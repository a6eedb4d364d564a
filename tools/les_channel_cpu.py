#!/usr/bin/env python
"""Long channel runs on the C/OpenMP restatement of the reference (oracle/c) as PHYSICS sanity of the oracle -- the restatement
the CUDA path is held to.  TEST INFRASTRUCTURE: runs the oracle, not the product.

  python tools/les_channel_cpu.py          BASELINE config 1: periodic channel Re_b = 5640, 64^3, dynamic Smagorinsky, 30000 steps
                                           (~8 min on 8 cores).  The flow transitions from the deterministic Poiseuille + vortex-pair
                                           start and the friction Reynolds number settles near the DNS value (180; a 64^3 LES sits
                                           5-10 % below).  Output: profiles/r2t_cpu_les_channel64_dsmag.log
  python tools/les_channel_cpu.py --wm     the wall-modelled channel of configs 3 / 5 (Re_b = 87000, log-law wall model, static
                                           Smagorinsky + van Driest) on a coarse 96 x 48 x 32 grid, 12000 steps (~2 min): Re_tau
                                           ~ 0.09 Re_b^0.88 ~ 2000 expected.  Output: profiles/r2t_cpu_wmles_channel_96x48x32.log"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.param as op  # noqa: E402
from oracle.cport import CSim  # noqa: E402
from oracle.initgrid import initgrid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wm", action="store_true")
    ap.add_argument("--steps", type=int, default=None)
    a = ap.parse_args()
    if a.wm:
        ng = (96, 48, 32)
        d = op.deck_channel(ng=ng, sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.)
        nsteps, every = a.steps or 12000, 1000
    else:
        ng = (64, 64, 64)
        d = op.deck_channel(ng=ng, sgstype="dsmag")
        nsteps, every = a.steps or 30000, 500
    c = CSim(d)
    dzc, dzf, zc, zf = initgrid(d.gtype, ng[2], d.gr, d.l[2])
    nu, h = d.visc, 0.5 * d.l[2]
    t, t0 = 0., time.time()
    I = (slice(1, -1),) * 3
    for it in range(1, nsteps + 1):
        c.step(icheck=10)
        t += c.dt
        if it % every:
            continue
        if a.wm:          # wall stress from the momentum balance: the bulk forcing of the last substep per unit time, times h
            tw = c.forcing()[0] / ((45. / 60. - 25. / 60.) * c.dt) * h
        else:             # resolved wall: first-cell velocity difference, both walls
            u = c.f["u"]
            tw = 0.5 * nu * ((u[1:-1, 1:-1, 1] - u[1:-1, 1:-1, 0]).mean() / dzc[0] + (u[1:-1, 1:-1, -2] - u[1:-1, 1:-1, -1]).mean() / dzc[ng[2]])
        ke = 0.5 * float((c.f["v"][I] ** 2).mean() + (c.f["w"][I] ** 2).mean())
        print("step %6d t %8.2f dt %.4f  tau_w %.3e  Re_tau %.1f  <v2+w2>/2 %.3e  (%.0f s)" % (it, t, c.dt, tw, np.sqrt(abs(tw)) * h / nu, ke, time.time() - t0),
              flush=True)
    c.close()


if __name__ == "__main__":
    main()
