"""Known-answer checks that pin the oracle to PHYSICS rather than to the text of the reference: closed-form results any
correct implementation of the reference's scheme must reproduce.  The reference ships no tests for this path (SURVEY.md
section 8c), so these -- with tests/test_oracle_identities.py -- are what stands between the restatement and a shared
misreading of the Fortran.  CPU only, small grids."""
import numpy as np
import pytest

from oracle.main import Sim
from oracle.param import Deck, c_smag, deck_channel, deck_duct, deck_tgv


def _run_to_steady(s, t_end):
    while s.time < t_end:
        s.step()


def discrete_poiseuille(dzc, dzf, visc, l3):
    """Dense solution of the discrete steady channel problem nu D2 u = -G, centred Dirichlet ghosts u_0 = -u_1 and
    u_{n+1} = -u_n, bulk velocity 1; dzc[k] = zc[k+1] - zc[k], dzf[k] = zf[k] - zf[k-1], k = 0..n3+1.  -> (u(1:n3), G)"""
    n3 = dzf.size - 2
    A = np.zeros((n3, n3))
    for k in range(1, n3 + 1):
        up, lo = 1. / (dzc[k] * dzf[k]), 1. / (dzc[k - 1] * dzf[k])
        A[k - 1, k - 1] = -(up + lo)
        if k < n3:
            A[k - 1, k] = up
        else:
            A[k - 1, k - 1] -= up
        if k > 1:
            A[k - 1, k - 2] = lo
        else:
            A[k - 1, k - 1] -= lo
    shape = np.linalg.solve(visc * A, -np.ones(n3))      # u for G = 1
    G = 1. / float((shape * dzf[1:n3 + 1] / l3).sum())
    return shape * G, G


def test_laminar_channel_reaches_the_discrete_poiseuille_solution_and_its_pressure_gradient():
    """Forced laminar channel (bulk velocity held at 1): the steady state of the scheme is the solution of the DISCRETE
    problem nu * D2 u = -G with the centred Dirichlet ghost u_0 = -u_1, mean(u) = 1 -- solved here with a dense matrix --
    and the forcing the code applies per unit time (sum of f over the substeps / dt, rk.f90:197-222) is G.  Pins the viscous
    term of mom_xyz_ad, set_bc('D'), the bulk forcing and the RK weights; on a stretched grid also dzci/dzfi."""
    for gr in (0., 2.):
        n3 = 12
        d = deck_channel(ng=(4, 4, n3), sgstype="none", visci=2., gr=gr)
        d.is_wallturb = False
        s = Sim(d)
        _run_to_steady(s, 3.0)
        u_ref, G = discrete_poiseuille(s.dzc_g, s.dzf_g, d.visc, d.l[2])
        u = s.U[0][1:-1, 1:-1, 1:-1]
        assert np.abs(u - u_ref[None, None, :]).max() < 1e-9, gr
        assert np.abs(s.V[0]).max() < 1e-12 and np.abs(s.W[0]).max() < 1e-12
        # forcing: one more step, accumulating f over the substeps
        ftot = 0.
        s.istep += 1
        for irk in range(3):
            s.substep(irk)
            ftot += s.f[0]
        assert abs(ftot / s.dt - G) < 1e-8 * G, (gr, ftot / s.dt, G)
        # continuum value 12 nu U_b / h^2 within the O(dz^2) of the grid
        assert abs(G - 12. * d.visc) < 0.05 * 12. * d.visc


def _duct_series(y, z, a=1., b=1., nterm=60):
    """Laminar flow in the rectangular duct |y| <= a, |z| <= b (White, Viscous Fluid Flow, eq. 3-48), unnormalised."""
    u = np.zeros((y.size, z.size))
    for m in range(nterm):
        n = 2 * m + 1
        u += (-1.) ** m / n ** 3 * (1. - np.cosh(n * np.pi * z[None, :] / (2. * a)) / np.cosh(n * np.pi * b / (2. * a))) * \
            np.cos(n * np.pi * y[:, None] / (2. * a))
    return u


def test_laminar_duct_converges_to_the_series_solution_at_second_order():
    """Forced laminar square duct (walls in y and z, DCT pressure solve in y, Neumann pressure walls): the steady state
    approaches the classical series solution with error O(h^2).  Independent of initflow's own 'duc' series."""
    errs = []
    for n in (8, 16):
        d = deck_duct(ng=(2, n, n), sgstype="none", visci=1.)
        s = Sim(d)
        _run_to_steady(s, 1.5)
        yc = (np.arange(1, n + 1) - .5) * d.dl[1] - 1.
        zc = s.zc_g[1:n + 1] - 1.
        ref = _duct_series(yc, zc)
        ref = ref / ref.mean()                           # bulk velocity 1 on the uniform grid
        u = s.U[0][1, 1:-1, 1:-1]
        assert np.abs(s.U[0][1:-1, 1:-1, 1:-1] - u[None]).max() < 1e-12      # x-invariant
        errs.append(np.abs(u - ref).max())
    assert errs[1] < 1.5e-2 and 3. < errs[0] / errs[1] < 5.5, errs


def _linear_shear(s, S):
    """u = S z on the oracle's arrays, ghost cells included (zc carries the ghost centres); v = w = 0"""
    r = s.world.ranks[0]
    u = s.U[0]
    u[:] = (S * s.st[0].zc)[None, None, :]
    s.V[0][:] = 0.
    s.W[0][:] = 0.
    return r


def test_smagorinsky_of_a_linear_shear_is_cs_delta_squared_times_the_shear_rate():
    """No walls (tri-periodic deck, ghosts filled linearly by hand): s0 = |S| and nu_t = (c_s Delta)^2 |S| with
    Delta = (dx dy dz)^(1/3), c_s = 0.11 (sgs.f90:103-150, 1088-1106)."""
    S = 3.7
    s = Sim(deck_tgv(ng=(6, 8, 10), sgstype="smag"))
    _linear_shear(s, S)
    s.cmpt_sgs()
    d = s.deck
    delta = (d.dl[0] * d.dl[1] * s.dzf_g[1:-1]) ** (1. / 3.)
    expect = (c_smag * delta) ** 2 * abs(S)
    v = s.VISCT[0][1:-1, 1:-1, 1:-1]
    assert np.abs(v - expect[None, None, :]).max() < 1e-14 * expect.max()


def test_van_driest_damping_of_a_linear_shear_closed_form():
    """Channel walls: with the Dirichlet ghost u_0 = -u_1 the wall shear of u = S z at the lower wall is nu S, so in the
    lower half d+ = z sqrt(nu S)/nu and nu_t = (c_s Delta (1 - exp(-d+/25)))^2 |S| (sgs.f90:106-150)."""
    S, n3 = 40., 16
    d = deck_channel(ng=(4, 4, n3), sgstype="smag", visci=500., gr=0.)
    d.is_wallturb = False
    s = Sim(d)
    _linear_shear(s, S)
    u = s.U[0]
    u[:, :, 0] = -u[:, :, 1]
    u[:, :, n3 + 1] = -u[:, :, n3]
    s.cmpt_sgs()
    zc = s.zc_g[1:-1]
    delta = (d.dl[0] * d.dl[1] * s.dzf_g[1:-1]) ** (1. / 3.)
    dplus = zc * np.sqrt(d.visc * S) / d.visc
    expect = (c_smag * delta * (1. - np.exp(-dplus / 25.))) ** 2 * S
    v = s.VISCT[0][1:-1, 1:-1, 1:-1]
    K = slice(1, n3 // 2)                                # away from the wall cell (ghost changes its strain rate) and the upper half
    assert np.abs(v[:, :, K] - expect[None, None, K]).max() < 1e-12 * expect[K].max()
    assert 0.05 < (1. - np.exp(-dplus[1] / 25.)) < 0.95  # the damping is active in the tested range


def test_dynamic_smagorinsky_switches_off_in_a_laminar_shear():
    """Germano-Lilly on u = S z: S_ij has only the 13 component, L_13 = filt(u w) - filt(u) filt(w) = 0 (w = 0) and
    M_11 = M_22 = M_33 = 0, so M:L = 0 and the eddy viscosity vanishes away from the walls -- the defining property of
    the dynamic model (sgs.f90:153-380)."""
    S, n3 = 5., 16
    d = deck_channel(ng=(6, 6, n3), sgstype="dsmag", visci=100., gr=0.)
    d.is_wallturb = False
    s = Sim(d)
    _linear_shear(s, S)
    u = s.U[0]
    u[:, :, 0] = -u[:, :, 1]
    u[:, :, n3 + 1] = -u[:, :, n3]
    s.cmpt_sgs()
    v = s.VISCT[0][1:-1, 1:-1, 1:-1]
    smag = (c_smag * (d.dl[0] * d.dl[1] * d.dl[2]) ** (1. / 3.)) ** 2 * S      # what the static model would give
    assert np.abs(v[:, :, 3:n3 - 3]).max() < 1e-10 * smag


def test_chkdt_closed_form_for_a_uniform_flow():
    """chkdt.f90:40-97: dtmax = min(0.4125 / dtid, 1.732 / dti) with dti = |u|/dx + |v|/dy + |w|/dz and
    dtid = (nu + nu_t)(1/dx^2 + 1/dy^2 + 1/dz^2) on a uniform grid."""
    from oracle import ops
    n = (5, 6, 7)
    dl = np.array([.1, .2, .05])
    dz = np.full(n[2] + 2, dl[2])
    a, b, c, nu, nut = .7, -1.3, .4, 1e-2, 3e-2
    shp = tuple(x + 2 for x in n)
    u, v, w, vt = (np.full(shp, x, order="F") for x in (a, b, c, nut))
    dt = ops.chkdt_local(n, dl, 1. / dz, 1. / dz, nu, vt, u, v, w)
    dti = abs(a) / dl[0] + abs(b) / dl[1] + abs(c) / dl[2]
    dtid = (nu + nut) * (1. / dl[0] ** 2 + 1. / dl[1] ** 2 + 1. / dl[2] ** 2)
    assert dt == pytest.approx(min(0.4125 / dtid, 1.732 / dti), rel=1e-13)
