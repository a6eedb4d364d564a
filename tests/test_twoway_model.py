"""Host-side model of the two-way (twisted-factorisation) z solve of the product build (cales_b200/csrc/gaussel_tab.cu:
gauss2_build_k / gauss2_k), statement for statement in numpy, against dense solves and against the reference's own
elimination order (dgtsv_homebrewed / gaussel_periodic, src/solver.f90:109-179): the upper half of a column is eliminated
top-down, the lower half bottom-up, the halves meet in a 2 x 2 system and are substituted outwards.  The kernels are checked
against the oracle on the GPU (tests/test_gpu_kernels.py); this checks the algebra without one."""
import numpy as np
import pytest

from test_zdist_model import EPS, dense, system, thomas


def meeting_row(nlev):
    return (((nlev + 15) // 16) // 2) * 16                      # first level of the lower half (level groups of 16)


def twoway(a, b, c, lam, r):
    """non-periodic system of nlev = len(r) rows"""
    nlev = len(r); s = meeting_row(nlev); bb = b + lam
    if s == 0 or s >= nlev:
        return thomas(a[:nlev], bb[:nlev], c[:nlev], r)        # one level group: the kernel is not used
    z = np.zeros(nlev); y = np.zeros(nlev)
    d = 0.; yy = 0.
    for l in range(s):                                          # warp 0: top-down
        z[l] = 1. / (bb[l] - a[l] * d + EPS); d = c[l] * z[l]; yy = (r[l] - a[l] * yy) * z[l]; y[l] = yy
    df, yf = d, yy
    d = 0.; yy = 0.
    for l in range(nlev - 1, s - 1, -1):                        # warp 1: bottom-up
        z[l] = 1. / (bb[l] - c[l] * d + EPS); d = a[l] * z[l]; yy = (r[l] - c[l] * yy) * z[l]; y[l] = yy
    db, yb = d, yy
    w = 1. / (1. - db * df + EPS)
    xs = (yb - db * yf) * w                                     # the 2 x 2 meeting system
    x = np.zeros(nlev)
    xx = xs
    for l in range(s - 1, -1, -1):
        xx = y[l] - (c[l] * z[l]) * xx; x[l] = xx
    xx = yf - df * xs
    for l in range(s, nlev):
        xx = y[l] - (a[l] * z[l]) * xx; x[l] = xx
    return x


def twoway_periodic(a, b, c, lam, r):
    """gaussel_periodic (solver.f90:109-151) with both sub-solves done two-way"""
    n = len(r)
    p1 = twoway(a[:n - 1], b[:n - 1], c[:n - 1], lam, r[:n - 1])
    r2 = np.zeros(n - 1); r2[0] = -a[0]; r2[n - 2] += -c[n - 2]
    p2 = twoway(a[:n - 1], b[:n - 1], c[:n - 1], lam, r2)
    pn = (r[n - 1] - c[n - 1] * p1[0] - a[n - 1] * p1[n - 2]) / ((b[n - 1] + lam) + c[n - 1] * p2[0] + a[n - 1] * p2[n - 2] + EPS)
    return np.concatenate([p1 + p2 * pn, [pn]])


@pytest.mark.parametrize("n", [32, 33, 48, 100, 256])
@pytest.mark.parametrize("periodic", [0, 1])
def test_twoway_vs_dense_and_vs_reference_order(n, periodic):
    rng = np.random.default_rng(10 * n + periodic)
    a, b, c = system(n, periodic, rng)
    for lam in (-3.7, -1e-2, -1e-5):
        r = rng.standard_normal(n)
        xr = np.linalg.solve(dense(a, b, c, lam, periodic), r)
        x = twoway_periodic(a, b, c, lam, r) if periodic else twoway(a, b, c, lam, r)
        tol = 1e-10 * max(1., 1e-4 / abs(lam)) * np.abs(xr).max()
        assert np.abs(x - xr).max() <= tol
        if not periodic:                                        # and the reference's one-way order gives the same to round-off
            assert np.abs(x - thomas(a, b + lam, c, r)).max() <= tol


def test_meeting_row_is_a_level_group_boundary_inside_the_column():
    for nlev in range(17, 1100):
        s = meeting_row(nlev)
        assert s % 16 == 0 and 0 < s < nlev
