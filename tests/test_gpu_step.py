"""GPU parity tests of the assembled path: solver, boundary conditions, SGS models and whole RK3
steps, CUDA (through the C ABI, driven like main.f90) against the CPU oracle.

Tolerances (BASELINE.json north_star): Poisson solution and post-correction divergence 1e-12
relative; fields after N steps 1e-10 relative (L-inf, normalised by max|field|).  Pressure-like
fields are compared after removing the volume mean: for all-periodic/Neumann problems the (0,0)
mode is singular and regularised by `+eps` pivots (solver.f90:165-170), so its additive constant
is round-off noise in the reference itself."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
from conftest import need_gpu  # noqa: E402


@pytest.fixture(autouse=True)
def _gpu(arith):
    """every test of this module runs for both arithmetic variants of the library (conftest.py `arith`)"""
    need_gpu()


def relerr(a, b, demean=False, scale=None):
    a = a[1:-1, 1:-1, 1:-1]; b = b[1:-1, 1:-1, 1:-1]
    if demean:
        a = a - a.mean(); b = b - b.mean()
    return float(np.abs(a - b).max() / max(np.abs(b).max() if scale is None else scale, 1e-300))


def vel_scale(ref):
    """velocity components are normalised by the largest of the three (a component that is identically zero in the
    reference, e.g. v of the laminar duct, carries O(1e-19) contraction residue x*y - x*y in the fma build)"""
    return max(float(np.abs(ref[k][1:-1, 1:-1, 1:-1]).max()) for k in ("u", "v", "w"))


def field_scale(nm, ref, vs):
    """normalisation of the error of field nm: velocities by the largest velocity component, the pressure by
    max(|p - mean|, U^2) (the pressure of the laminar duct is identically zero: its scale is the dynamic pressure)"""
    if nm in ("u", "v", "w"):
        return vs
    if nm == "p":
        r = ref[1:-1, 1:-1, 1:-1]
        return max(float(np.abs(r - r.mean()).max()), vs * vs)
    return None


def make_pair(name, ng=None, **kw):
    """The same deck for the oracle and for the product."""
    import oracle.param as op
    import cales_b200.deck as pd
    from oracle.main import Sim
    from cales_b200.driver import Simulation
    args = dict(kw)
    if ng is not None:
        args["ng"] = ng
    od = getattr(op, name)(**args)
    dd = getattr(pd, name)(**args)
    return od, dd, Sim, Simulation


from parity_mgpu import CASES  # noqa: E402  (one table for the single- and multi-GPU parity tests)


def run_pair(case, nsteps, impdiff=None):
    name, kw = CASES[case]
    od, dd, Sim, Simulation = make_pair(name, **kw)
    if impdiff:
        for d in (od, dd):
            d.impdiff = True
            d.impdiff_1d = impdiff == "1d"
    o = Sim(od)
    g = Simulation(dd)
    g.init_flow()
    g.start()
    assert abs(g.dt - o.dt) <= 1e-12 * o.dt      # (the device-generated initial field differs from numpy by round-off)
    out = []
    for _ in range(nsteps):
        ro = o.step(icheck=1)
        rg = g.step(icheck=1)
        out.append((ro, rg))
    return o, g, out


def compare(o, g, tol):
    errs = {}
    vs = vel_scale({"u": o.U[0], "v": o.V[0], "w": o.W[0]})
    for nm, on in (("u", "U"), ("v", "V"), ("w", "W"), ("p", "P"), ("visct", "VISCT")):
        errs[nm] = relerr(g.get(nm), getattr(o, on)[0], demean=(nm == "p"), scale=field_scale(nm, getattr(o, on)[0], vs))
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, errs
    return errs


@pytest.mark.parametrize("case", list(CASES))
def test_ten_steps(case):
    o, g, out = run_pair(case, 10)
    compare(o, g, 1e-10)
    for ro, rg in out:
        # post-correction divergence: both at round-off level and equal to 1e-12 of the velocity gradient scale
        assert rg[1] < 1e-9 and abs(rg[1] - ro[1]) <= 1e-12 * max(1.0, np.abs(g.get("u")).max() * max(g.deck.dli))
    assert abs(g.dt - o.dt) <= 1e-10 * o.dt
    g.close()


@pytest.mark.parametrize("case", ["channel_dsmag", "tgv_smag", "channel_wm_smag"])
def test_hundred_steps(case):
    """north_star: velocity/pressure fields after 100 steps agree to relative 1e-10."""
    o, g, out = run_pair(case, 100)
    compare(o, g, 1e-10)
    g.close()


@pytest.mark.parametrize("case", ["channel_dsmag", "tgv_smag"])
def test_restart_from_checkpoint(case, tmp_path):
    """`fld.bin` (load.f90:20-187): 2 steps + save + load in a fresh context + 2 steps == 4 steps to round-off (the first
    RK substep has rkpar(2) = 0, so no history crosses the restart; main.f90:524 suggests icheck=1 for this check)."""
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    name, kw = CASES[case]
    a = Simulation(getattr(pd, name)(**kw))
    a.init_flow(); a.start()
    for _ in range(2):
        a.step(icheck=1)
    fn = str(tmp_path / "fld.bin")
    a.save(fn)
    for _ in range(2):
        a.step(icheck=1)
    b = Simulation(getattr(pd, name)(**kw))
    b.load(fn)
    assert b.istep == 2 and b.time > 0.
    b.start()
    for _ in range(2):
        b.step(icheck=1)
    assert b.istep == a.istep == 4 and abs(b.time - a.time) <= 1e-14 * a.time and abs(b.dt - a.dt) <= 1e-14 * a.dt
    for nm in ("u", "v", "w", "p", "visct"):
        fa, fb = a.get(nm)[1:-1, 1:-1, 1:-1], b.get(nm)[1:-1, 1:-1, 1:-1]
        assert np.abs(fa - fb).max() <= 1e-13 * max(np.abs(fa).max(), 1e-300), nm
    a.close(); b.close()


@pytest.mark.parametrize("case,mode", [("channel_smag", "3d"), ("channel_smag", "1d"), ("tgv_smag", "1d"), ("channel_wm_smag", "1d"),
                                       ("tgv_smag", "3d"), ("duct_smag", "3d"), ("cavity_smag", "3d"), ("duct_smag", "1d")])
def test_implicit_diffusion(case, mode):
    """Crank-Nicolson paths (_IMPDIFF / _IMPDIFF_1D, main.f90:423-491).  The 3-D Helmholtz solves of the duct and the
    cavity run the face-centred transforms (RODFT00 for the wall-normal component, fft.f90:221-244) that the
    reference's own GPU path lacks (fft.f90:567-569)."""
    o, g, out = run_pair(case, 5, impdiff=mode)
    compare(o, g, 1e-10)
    g.close()


@pytest.mark.parametrize("bcs", ["PPPPNN", "PPPPPP", "NNNNNN", "PPNNNN", "DDPPDD", "PPDDNN"])
@pytest.mark.parametrize("ng", [(32, 24, 16), (64, 64, 64), (96, 30, 40)])
def test_poisson_solver(bcs, ng):
    """solver(): CUDA vs oracle to 1e-12 (de-meaned) and the discrete-Laplacian residual of both."""
    import ctypes as C
    import oracle.param as op
    import cales_b200.deck as pd
    from oracle.main import Sim
    from cales_b200.driver import Simulation
    def mk(mod):
        d = mod.Deck(ng=ng, l=(3.0, 2.0, 1.5), gtype=1, gr=2.0, inivel="zer", sgstype="none")
        for idir in range(3):
            for ib in range(2):
                c = bcs[2 * idir + ib]
                d.cbcpre[ib, idir] = c
                for ivel in range(3):
                    d.cbcvel[ib, idir, ivel] = "P" if c == "P" else "D"
                d.cbcsgs[ib, idir] = "P" if c == "P" else "D"
        return d
    o = Sim(mk(op)); g = Simulation(mk(pd))
    rng = np.random.default_rng(11)
    rhs = np.asfortranarray(rng.standard_normal((ng[0] + 2, ng[1] + 2, ng[2] + 2)))
    singular = all(c in "PN" for c in bcs)
    if singular:
        # compatibility: the volume-weighted mean of the rhs must vanish
        wgt = o.st[0].dzf[1:-1][None, None, :]
        rhs[1:-1, 1:-1, 1:-1] -= (rhs[1:-1, 1:-1, 1:-1] * wgt).sum() / (wgt.sum() * ng[0] * ng[1])
    # exact lambdaxy/a/b/c parity of initsolver (index/eigenvalue maps must be bit-exact)
    assert np.array_equal(g.poi["lam_h"], o.lambdaxyp[0])
    assert np.array_equal(g.poi["a_h"], o.ap) and np.array_equal(g.poi["b_h"], o.bp) and np.array_equal(g.poi["c_h"], o.cp)
    assert g.poi["normfft"] == o.plan_p.normfft
    po = [rhs.copy(order="F")]
    o.solve_poisson(po)
    g.set_fields(pp=rhs)
    g.solver(g.poi, "pp")
    pg = g.get("pp")
    # noise floor of the reference algorithm itself: the same oracle solve with the rhs perturbed by one
    # ulp-sized relative noise.  For the singular (0,0) mode with periodic z, gaussel_periodic divides an
    # O(eps) residual by an O(eps) pivot (solver.f90:142-143) and adds the resulting huge constant to every
    # point, so the de-meaned field carries its rounding ("conditioning caps parity").
    pert = [np.asfortranarray(rhs * (1. + 1.1e-16 * np.sign(rng.standard_normal(rhs.shape))))]
    o.solve_poisson(pert)
    floor = relerr(pert[0], po[0], demean=singular)
    e = relerr(pg, po[0], demean=singular)
    # strict build: dgtsv_homebrewed's own operation order, within 20 floors.  Product build: the two-way (twisted)
    # factorisation is a different, equally stable elimination -- its rounding is independent of the reference's, so the
    # distance is that of two independent realisations of the floor (measured: up to 21 floors on the stretched
    # 96x30x40 all-Neumann grid, whose floor is 1.7e-13); 50 floors bound it.
    from cales_b200 import lib as L_
    assert e <= max(1e-12, (20. if L_.DEFAULT_ARITH == "strict" else 50.) * floor), (e, floor)
    # residual of the discrete problem: boundp + Laplacian applied to the GPU result and, as the floor, to the
    # oracle's own result (the additive constant of the singular mode is removed first: it would drown the
    # check in rounding; with periodic z its profile noise is the reference algorithm's own)
    from oracle import bound as ob
    s = o.st[0]
    I = (slice(1, -1),) * 3
    k = np.arange(1, ng[2] + 1)

    def residual(field):
        q = [field - (field[1:-1, 1:-1, 1:-1].mean() if singular else 0.)]
        ob.boundp(o.world, o.deck.cbcpre, o.st, "bcp", q)
        p = q[0]
        lap = (p[2:, 1:-1, 1:-1] - 2 * p[I] + p[:-2, 1:-1, 1:-1]) * s.dli[0] ** 2 + \
              (p[1:-1, 2:, 1:-1] - 2 * p[I] + p[1:-1, :-2, 1:-1]) * s.dli[1] ** 2 + \
              ((p[1:-1, 1:-1, 2:] - p[I]) * s.dzci[k] - (p[I] - p[1:-1, 1:-1, :-2]) * s.dzci[k - 1]) * s.dzfi[k]
        return np.abs(lap - rhs[I]).max() / np.abs(rhs[I]).max()
    res, res_o = residual(pg), residual(po[0])
    assert res <= max(1e-10, 20. * res_o), (res, res_o)
    g.close()


def test_full_size_properties():
    """BASELINE config 2 size (256^3 tri-periodic, smag): size-independent properties of the GPU path --
    divergence after projection at round-off, Poisson round trip, kinetic energy decays."""
    from cales_b200.deck import deck_tgv
    from cales_b200.driver import Simulation
    g = Simulation(deck_tgv(ng=(256, 256, 256)))
    g.init_flow(); g.start()
    def ke():
        return float(sum((g.fields[c] ** 2).sum().item() for c in "uvw"))
    e0 = ke()
    for _ in range(3):
        tot, mx = g.step(icheck=1)
        assert mx < 1e-11, mx
    assert ke() < e0
    g.close()


@pytest.mark.parametrize("case", ["channel_dsmag", "tgv_smag", "channel_wm_smag", "duct_wm_smag", "cavity_smag"])
def test_fused_step_identical_to_per_procedure_sequence(case, arith):
    """the fused entries (cales_substep / cales_step: update fused into the momentum kernel, out-of-place correc +
    updatep, CUDA-graph replay) must give the results of the per-procedure sequence main.f90:418-506 -- identical bits in the
    strict build, 1e-12 in the contraction build (the compiler contracts the two code shapes differently); graph replay vs the
    eager fused sequence: identical bits in both.  dt is re-evaluated every second step, so graphs are re-captured on the way."""
    import cales_b200.deck as pd
    from cales_b200.driver import Simulation
    name, kw = CASES[case]
    sims = [Simulation(getattr(pd, name)(**kw), fused=f, graph=g) for f, g in ((False, False), (True, False), (True, True))]
    assert sims[2].graph and sims[1].fused and not sims[0].fused
    for s in sims:
        s.init_flow(); s.start()
        for _ in range(7):
            s.step(icheck=2)
    ref = {nm: sims[0].get(nm) for nm in ("u", "v", "w", "p", "visct")}
    vs = vel_scale(ref)
    for nm in ref:
        b, c = sims[1].get(nm), sims[2].get(nm)
        assert np.array_equal(b, c), (nm, "graph replay differs from the eager fused sequence")
        if arith == "strict":
            assert np.array_equal(ref[nm], b), nm
        else:
            assert relerr(b, ref[nm], demean=(nm == "p"), scale=field_scale(nm, ref[nm], vs)) <= 1e-12, nm
    assert sims[0].dt == sims[1].dt or arith == "fma"
    for s in sims:
        s.close()


@pytest.mark.parametrize("case,ave,f2d", [("tgv_smag", "dit", False), ("duct_smag", "duct", False), ("cavity_smag", "cavity", False),
                                          ("channel_dsmag", "channel", True), ("duct_wm_smag", "duct", True)])
def test_dsmag_averaging_and_filter_variants(case, ave, f2d):
    """The build-time variants of the dynamic model: averaging over the whole volume (_DIT, ave0d_dit
    sgs.f90:388-431), along the streamwise direction (_DUCT, ave2d_duct 540-614), none (_CAVITY), and the 2-D test filter
    (-D_FILTER_2D, filter2d 824-848), each against the oracle on the matching flow, 5 steps, 1e-10."""
    import oracle.param as op
    import cales_b200.deck as pd
    from oracle.main import Sim
    from cales_b200.driver import Simulation
    name, kw = CASES[case]
    kw = dict(kw); kw["sgstype"] = "dsmag"
    od, dd = getattr(op, name)(**kw), getattr(pd, name)(**kw)
    if name == "deck_duct":                 # the laminar duct profile is x-uniform (nu_t = 0): start from an unforced TGV field instead
        for d in (od, dd):
            d.inivel = "tgv"; d.is_forced = (False, False, False)
    o = Sim(od, ave=ave, filter_2d=f2d)
    g = Simulation(dd, ave=ave, filter_2d=f2d)
    g.init_flow(); g.start()
    for _ in range(5):
        o.step(icheck=1); g.step(icheck=1)
    compare(o, g, 1e-10)
    assert float(np.abs(o.VISCT[0]).max()) > 0.
    g.close()
