"""CPU tests that pin the oracle itself (the reference ships no tests for this path, so the oracle is
"parity unpinned"; these are the analytical identities and known-answer checks named in SURVEY.md section 8(c))."""
import numpy as np
import pytest
import scipy.fft as sfft

from oracle import solver as osl
from oracle.main import Sim
from oracle.param import Deck, deck_cavity, deck_channel, deck_duct, deck_tgv, eps, rkcoeff, small


def test_constants():
    assert small == eps * 1e7 and abs(sum(rkcoeff[0]) + sum(rkcoeff[1]) + sum(rkcoeff[2]) - 1.0) < 1e-15


@pytest.mark.parametrize("kinds", [("R2HC", "HC2R", 1.0, 0), ("REDFT10", "REDFT01", 2.0, 0), ("RODFT10", "RODFT01", 2.0, 0),
                                   ("REDFT11", "REDFT11", 2.0, 0), ("RODFT11", "RODFT11", 2.0, 0), ("REDFT00", "REDFT00", 2.0, -1),
                                   ("RODFT00", "RODFT00", 2.0, 1)])
@pytest.mark.parametrize("n", [8, 15, 64, 96])
def test_transform_round_trip(kinds, n):
    """bwd(fwd(x)) * normfft = x with the norm table of find_fft (fft.f90:192-245)."""
    kf, kb, n1, n2 = kinds
    x = np.random.default_rng(1).standard_normal((n, 3))
    y = x.copy(); osl.fft(kf, n, y, 0); osl.fft(kb, n, y, 0)
    assert np.allclose(y / (n1 * (n + n2)), x, atol=1e-13)


def test_halfcomplex_layout_matches_fftw_definition():
    n = 10
    x = np.random.default_rng(2).standard_normal(n)
    y = x.copy()[:, None]; osl.fft("R2HC", n, y, 0)
    X = np.fft.fft(x)
    ref = np.concatenate([X.real[:n // 2 + 1], X.imag[1:(n + 1) // 2][::-1]])
    assert np.allclose(y[:, 0], ref, atol=1e-13)


@pytest.mark.parametrize("bc,cf", [("PP", "c"), ("NN", "c"), ("DD", "c"), ("ND", "c"), ("DN", "c")])
def test_eigenvalues_diagonalise_second_difference(bc, cf):
    """fwd transform of the discrete second difference of x == lambda * fwd(x) for the BC-consistent ghost cells."""
    n = 24
    x = np.random.default_rng(3).standard_normal(n)
    g = np.zeros(n + 2); g[1:-1] = x
    g[0] = {"P": x[-1], "N": x[0], "D": -x[0]}[bc[0]]
    g[-1] = {"P": x[0], "N": x[-1], "D": -x[-1]}[bc[1]]
    d2 = g[2:] - 2 * g[1:-1] + g[:-2]
    kf, kb, norm = osl.find_fft(np.array(list(bc)), cf)
    a = x.copy()[:, None]; b = d2.copy()[:, None]
    osl.fft(kf, n, a, 0); osl.fft(kf, n, b, 0)
    lam = osl.eigenvalues(n, np.array(list(bc)), cf)
    assert np.allclose(b[:, 0], lam * a[:, 0], atol=1e-11)


def test_thomas_vs_dense():
    rng = np.random.default_rng(4)
    n = 40
    a = 0.5 + rng.random(n); c = 0.5 + rng.random(n); b = -(a + c) - 0.3
    p = rng.standard_normal((2, 3, n))
    ref = p.copy(); osl.gaussel(2, 3, n, a, b, c, ref)
    A = np.diag(b) + np.diag(a[1:], -1) + np.diag(c[:-1], 1)
    assert np.allclose(ref[1, 2], np.linalg.solve(A, p[1, 2]), rtol=1e-10)
    refp = p.copy(); osl.gaussel_periodic(2, 3, n, a, b, c, refp)
    A[0, -1] = a[0]; A[-1, 0] = c[-1]
    assert np.allclose(refp[1, 2], np.linalg.solve(A, p[1, 2]), rtol=1e-9)


@pytest.mark.parametrize("deck", [deck_channel(ng=(16, 12, 14), sgstype="none"), deck_tgv(ng=(12, 12, 12), sgstype="none"),
                                  deck_duct(ng=(8, 12, 12), sgstype="none"), deck_cavity(ng=(12, 10, 8), sgstype="none")])
def test_poisson_residual_and_projection(deck):
    """Laplacian(solver(rhs)) = rhs and div(u) after correction at round-off (main.f90:538 demands < small)."""
    s = Sim(deck)
    for _ in range(2):
        divtot, divmax = s.step(icheck=1)
        assert divmax < 1e-11
    assert divmax < small


def test_taylor_green_2d_decay():
    """2-D Taylor-Green vortex (inivel 'tgw', utils/useful_fortran_blocks/tgv_validation-inc.f90:21-37): the
    kinetic energy decays like exp(-4 nu t); second-order accurate scheme -> small error at 32^2."""
    L = 2 * np.pi
    d = Deck(ng=(32, 32, 4), l=(L, L, L / 8), gtype=1, gr=0., visci=10., inivel="tgw", sgstype="none", cfl=0.5)
    s = Sim(d)
    u0 = s.U[0][1:-1, 1:-1, 1:-1].copy()
    for _ in range(20):
        s.step()
    u = s.U[0][1:-1, 1:-1, 1:-1]
    expect = u0 * np.exp(-2 * d.visc * s.time)
    assert np.abs(u - expect).max() < 2e-3 * np.abs(u0).max()


@pytest.mark.parametrize("deck", [deck_channel(ng=(16, 8, 12), sgstype="dsmag"), deck_channel(ng=(16, 8, 12), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.),
                                  deck_tgv(ng=(8, 8, 8)), deck_cavity(ng=(8, 8, 8)), deck_duct(ng=(8, 12, 12), wall_model=True)])
@pytest.mark.parametrize("dims", [(2, 1), (1, 2), (2, 2)])
def test_emulated_ranks_match_single_rank(deck, dims):
    """The decomposed run (halo exchange order, is_bound masks, plane averages) reproduces the single-rank run."""
    a = Sim(deck); b = Sim(deck.copy(dims=dims))
    for _ in range(2):
        a.step(); b.step()
    for nm in ("U", "V", "W", "VISCT"):
        x, y = a.world.gather(getattr(a, nm)), b.world.gather(getattr(b, nm))
        assert np.abs(x - y).max() <= 1e-12 * max(np.abs(x).max(), 1e-30), nm


def test_wallmodel_newton_converges_to_loglaw():
    from oracle import wmodel
    from oracle.param import b_log, kap_log
    uh = np.array([0.5, 1.0, 2.0]); vh = np.array([0.1, 0.0, -0.3])
    h, visc = 0.1, 1e-5
    t1, t2 = wmodel.wallmodel(1, uh, vh, h, 2.0, visc)
    utau = (t1 ** 2 + t2 ** 2) ** 0.25
    upar = np.hypot(uh, vh)
    assert np.allclose(upar / utau, np.log(h * utau / visc) / kap_log + b_log, rtol=1e-7)
    assert wmodel.last_iters.max() <= 8
