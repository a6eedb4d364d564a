"""The C/OpenMP restatement (oracle/c, the multi-threaded CPU arm of bench.py) against the numpy oracle: two
independently written restatements of the same reference lines must agree -- bit for bit wherever no
transform is involved (ghost fill, strain rate / Smagorinsky, chkdt: the library is built with
-ffp-contract=off), to round-off after full RK3 steps (own Stockham FFT vs pocketfft)."""
import numpy as np
import pytest

import oracle.param as op
from oracle.cport import CSim
from oracle.main import Sim

PAIRS = (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("visct", "VISCT"))


@pytest.mark.parametrize("ng", [(32, 24, 16), (30, 20, 12), (16, 18, 7)])
def test_c_port_matches_numpy_oracle(ng):
    d = op.deck_tgv(ng=ng)
    o, c = Sim(d), CSim(d)
    try:
        assert c.dt == o.dt                                   # chkdt bit-exact on the initial field
        for nm, on in PAIRS:                                  # initial ghost fill + first eddy viscosity: bit-exact
            assert np.array_equal(c.f[nm], getattr(o, on)[0]), nm
        for _ in range(3):
            divo = o.step(icheck=1)
            divc = c.step(icheck=1)
        assert divc < 1e-12 and divo[1] < 1e-12
        for nm, on in PAIRS:
            a = c.f[nm][1:-1, 1:-1, 1:-1]
            b = getattr(o, on)[0][1:-1, 1:-1, 1:-1]
            if nm == "p":
                a = a - a.mean(); b = b - b.mean()
            err = np.abs(a - b).max() / np.abs(b).max()
            assert err < 1e-12, (nm, err)                     # north-star tolerance for the Poisson chain
        assert abs(c.dt - o.dt) <= 1e-13 * o.dt
    finally:
        c.close()


def test_c_port_poisson_residual():
    """The discrete Laplacian of the C solver's solution reproduces the right-hand side (de-meaned)."""
    ng = (24, 20, 10)
    d = op.deck_tgv(ng=ng)
    c = CSim(d)
    try:
        rng = np.random.default_rng(1234)
        rhs = rng.standard_normal(ng)
        rhs -= rhs.mean()
        c.f["pp"][1:-1, 1:-1, 1:-1] = rhs
        c.lib.cales_cpu_solver(c.h)
        c.lib.cales_cpu_boundp(c.h, 4)
        p = c.f["pp"]
        dli = d.dli
        lap = ((p[2:, 1:-1, 1:-1] - 2 * p[1:-1, 1:-1, 1:-1] + p[:-2, 1:-1, 1:-1]) * dli[0] ** 2 +
               (p[1:-1, 2:, 1:-1] - 2 * p[1:-1, 1:-1, 1:-1] + p[1:-1, :-2, 1:-1]) * dli[1] ** 2 +
               (p[1:-1, 1:-1, 2:] - 2 * p[1:-1, 1:-1, 1:-1] + p[1:-1, 1:-1, :-2]) * dli[2] ** 2)
        # all-periodic: the singular (0,0) mode is only regularised by the +eps pivots (solver.f90:165-170), so the
        # solution carries a huge round-off-determined constant; its ulp bounds the residual
        floor = np.finfo(float).eps * np.abs(p).max() * 4. * (dli[0] ** 2 + dli[1] ** 2 + dli[2] ** 2)
        assert np.abs(lap - rhs).max() < max(1e-11 * np.abs(rhs).max(), 4. * floor)
    finally:
        c.close()
