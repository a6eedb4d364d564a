#!/usr/bin/env python
"""Off-box pin of the oracle: compare a restart file written by this code with one written by a
site-built CaLES (the Fortran/MPI/FFTW or nvfortran/cuDecomp reference, which cannot be built in this image) for the SAME
input deck and number of steps.  Both are `fld.bin` in the reference's format (src/load.f90:20-187: u,v,w,p global,
halo-free, Fortran order, then [time, istep]).

  python tools/compare_fld.py OURS.bin REFERENCE.bin --ng NX NY NZ [--tol 1e-10]

Prints the relative L-inf difference per field (velocities normalised by the largest component, pressure after removing
the volume mean -- its additive constant is round-off of the singular Poisson mode) and exits non-zero above --tol.
Recipe ("Pinning the oracle off this box"): run the reference with `nstep = 100, isave = 100,
is_overwrite_save = T` on a deterministic deck (e.g. examples/dns/_manuscript_taylor_green_vortex, or the channel deck
with is_wallturb = T), run `python -m cales_b200.run input.nml` on the same deck, and compare the two data/fld.bin."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cales_b200 import checkpoint as ck  # noqa: E402


def read(fn, ng):
    f = [np.zeros((ng[0] + 2, ng[1] + 2, ng[2] + 2), order="F") for _ in range(4)]
    time, istep = ck.load_all("r", fn, ng, (1, 1, 1), ng, *f)
    return [a[1:-1, 1:-1, 1:-1] for a in f], time, istep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ours"); ap.add_argument("reference")
    ap.add_argument("--ng", type=int, nargs=3, required=True)
    ap.add_argument("--tol", type=float, default=1e-10)
    a = ap.parse_args()
    fa, ta, ia = read(a.ours, a.ng); fb, tb, ib = read(a.reference, a.ng)
    print("ours: time %r istep %d | reference: time %r istep %d" % (ta, ia, tb, ib))
    vs = max(float(np.abs(x).max()) for x in fb[:3])
    worst = 0.
    for nm, x, y in zip("uvwp", fa, fb):
        if nm == "p":
            x = x - x.mean(); y = y - y.mean()
            scale = max(float(np.abs(y).max()), vs * vs)
        else:
            scale = vs
        e = float(np.abs(x - y).max()) / max(scale, 1e-300)
        worst = max(worst, e)
        print("%s: relative L-inf difference %.3e" % (nm, e))
    ok = worst <= a.tol and ia == ib and abs(ta - tb) <= 1e-10 * max(abs(tb), 1e-300)
    print("PASS" if ok else "FAIL", "(tolerance %.1e)" % a.tol)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
