"""GPU parity tests, kernel by kernel: every call goes through the C ABI and is compared with the CPU oracle on
the same seeded inputs, for BOTH arithmetic variants of the library (conftest.py `arith`):
  strict  libcales_b200_strict.so (-fmad=false): pure stencil arithmetic must agree BIT FOR BIT with numpy;
  fma     libcales_b200.so (the product build, fp64 contraction on as in the reference's own GPU build): the same
          kernels to FMA_TOL = 1e-13 relative to max|reference| (a contraction changes one rounding per a*b+c).
Transcendental functions (exp/log/pow) and re-ordered sums get a stated round-off tolerance in both."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from conftest import need_gpu  # noqa: E402

FMA_TOL = 1e-13
_ARITH = ["strict"]


@pytest.fixture
def env(arith):
    need_gpu()
    from cales_b200 import lib as L
    lib = L.load(arith)
    _ARITH[0] = arith
    return L, lib


def same(got, ref):
    """strict build: identical bits; fma build: FMA_TOL relative to max|ref|."""
    if _ARITH[0] == "strict":
        return np.array_equal(got, ref)
    return got.shape == ref.shape and float(np.abs(got - ref).max()) <= FMA_TOL * max(float(np.abs(ref).max()), 1e-300)


class Ctx:
    def __init__(self, L, lib, ng, cbcpre="PPPPPP", diffusion=0):
        self.L, self.lib = L, lib
        self.ctx = C.c_void_p()
        # the library works on torch's current stream, like the uploads of the test
        L.check(None, lib.cales_init(C.byref(self.ctx), L._ia(ng), L._ia([1, 1]), 1, cbcpre.encode(), 0, 1, None, 0,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream), diffusion), lib)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.lib.cales_finalize(self.ctx)

    def chk(self, rc):
        self.L.check(self.ctx, rc, self.lib)


_KEEP = []


def dev(a):
    """Upload; the tensor is kept alive until the next test starts (kernels are asynchronous and the caching
    allocator would otherwise hand the storage of a temporary to the next upload)."""
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))).cuda()
    _KEEP.append(t)
    return t


@pytest.fixture(autouse=True)
def _clear_keep():
    _KEEP.clear()
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def host(t, shape):
    return t.cpu().numpy().reshape(shape, order="F")


def rnd_fields(n, seed, count):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray(rng.standard_normal((n[0] + 2, n[1] + 2, n[2] + 2))) for _ in range(count)]


def grid(n3, seed=7):
    rng = np.random.default_rng(seed)
    dzf = 0.5 + rng.random(n3 + 2)
    dzc = np.zeros(n3 + 2)
    dzc[:-1] = .5 * (dzf[:-1] + dzf[1:]); dzc[-1] = dzc[-2]
    return dzc, dzf, 1. / dzc, 1. / dzf


SIZES = [(16, 12, 10), (67, 9, 33), (64, 64, 64)]


@pytest.mark.parametrize("n", SIZES)
def test_fillps_correc_updatep_bitexact(env, n):
    from oracle import ops
    L, lib = env
    u, v, w, p, pp = rnd_fields(n, 1, 5)
    dzc, dzf, dzci, dzfi = grid(n[2])
    dli = np.array([3.1, 2.7, 1.9]); dt = 0.0123
    with Ctx(L, lib, n) as c:
        du, dv, dw, dp, dpp = map(dev, (u, v, w, p, pp))
        ddzci, ddzfi = dev(dzci), dev(dzfi)
        c.chk(lib.cales_fillps(c.ctx, L._ia(n), L._da(dli), ddzfi.data_ptr(), 1. / dt, du.data_ptr(), dv.data_ptr(), dw.data_ptr(), dp.data_ptr()))
        p_ref = p.copy(order="F"); ops.fillps(n, dli, dzfi, 1. / dt, u, v, w, p_ref)
        assert same(host(dp, p.shape), p_ref)
        c.chk(lib.cales_correc(c.ctx, L._ia(n), L._da(dli), ddzci.data_ptr(), dt, dpp.data_ptr(), du.data_ptr(), dv.data_ptr(), dw.data_ptr()))
        ur, vr, wr = u.copy(order="F"), v.copy(order="F"), w.copy(order="F")
        ops.correc(n, dli, dzci, dt, pp, ur, vr, wr)
        assert same(host(du, u.shape), ur) and same(host(dv, u.shape), vr) and same(host(dw, u.shape), wr)
        dp2 = dev(p)
        c.chk(lib.cales_updatep(c.ctx, L._ia(n), L._da(dli), ddzci.data_ptr(), ddzfi.data_ptr(), 0.0, dpp.data_ptr(), dp2.data_ptr()))
        p_ref = p.copy(order="F"); ops.updatep(n, dli, dzci, dzfi, 0.0, pp, p_ref)
        assert same(host(dp2, p.shape), p_ref)
    for mode, (imp, imp1) in ((1, (True, False)), (2, (True, True))):
        with Ctx(L, lib, n, diffusion=mode) as c:
            dp2, dpp = dev(p), dev(pp)
            c.chk(lib.cales_updatep(c.ctx, L._ia(n), L._da(dli), dev(dzci).data_ptr(), dev(dzfi).data_ptr(), -0.37, dpp.data_ptr(), dp2.data_ptr()))
            p_ref = p.copy(order="F"); ops.updatep(n, dli, dzci, dzfi, -0.37, pp, p_ref, imp, imp1)
            assert same(host(dp2, p.shape), p_ref)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_mom_xyz_ad_bitexact(env, n, mode):
    from oracle import mom
    L, lib = env
    u, v, w, s = rnd_fields(n, 2, 4)
    s = np.abs(s)
    dzc, dzf, dzci, dzfi = grid(n[2])
    dxi, dyi, visc = 3.3, 2.1, 1e-2
    (ru, rv, rw), imp = mom.mom_xyz_ad(n, dxi, dyi, dzci, dzfi, visc, u, v, w, s, mode > 0, mode == 2)
    with Ctx(L, lib, n, diffusion=mode) as c:
        out = [torch.zeros(n[0] * n[1] * n[2], dtype=torch.float64, device="cuda") for _ in range(6)]
        c.chk(lib.cales_mom_xyz_ad(c.ctx, L._ia(n), dxi, dyi, dev(dzci).data_ptr(), dev(dzfi).data_ptr(), visc, dev(u).data_ptr(),
                                   dev(v).data_ptr(), dev(w).data_ptr(), dev(s).data_ptr(), *[o.data_ptr() for o in out]))
        for o, r in zip(out[:3], (ru, rv, rw)):
            assert same(host(o, tuple(n)), r)
        if mode:
            for o, r in zip(out[3:], imp):
                assert same(host(o, tuple(n)), r)


def test_mom_polynomial_check(env):
    """The reference's own analytical check (src/mom.f90:20-21): with visct = x+y+z and u=v=w=x*y*z the
    SGS cross terms are polynomial; compare oracle and device, and both against the closed form of the
    u-equation diffusion part for a uniform grid."""
    from oracle import mom
    L, lib = env
    n = (24, 20, 16)
    dl = np.array([0.1, 0.2, 0.05])
    i = np.arange(0, n[0] + 2)[:, None, None]; j = np.arange(0, n[1] + 2)[None, :, None]; k = np.arange(0, n[2] + 2)[None, None, :]
    xc, yc, zc = (i - .5) * dl[0], (j - .5) * dl[1], (k - .5) * dl[2]
    xf, yf, zf = i * dl[0], j * dl[1], k * dl[2]
    u = np.asfortranarray(xf * yc * zc); v = np.asfortranarray(xc * yf * zc); w = np.asfortranarray(xc * yc * zf)
    s = np.asfortranarray(xc + yc + zc + 0 * u)
    dzci = np.full(n[2] + 2, 1. / dl[2]); dzfi = dzci.copy()
    (ru, rv, rw), _ = mom.mom_xyz_ad(n, 1 / dl[0], 1 / dl[1], dzci, dzfi, 0.0, u, v, w, s)
    with Ctx(L, lib, n) as c:
        out = [torch.zeros(n[0] * n[1] * n[2], dtype=torch.float64, device="cuda") for _ in range(3)]
        c.chk(lib.cales_mom_xyz_ad(c.ctx, L._ia(n), 1 / dl[0], 1 / dl[1], dev(dzci).data_ptr(), dev(dzfi).data_ptr(), 0.0, dev(u).data_ptr(),
                                   dev(v).data_ptr(), dev(w).data_ptr(), dev(s).data_ptr(), *[o.data_ptr() for o in out], None, None, None))
        assert same(host(out[0], tuple(n)), ru)
    # closed form: d/dx[2 nu_t u_x] + d/dy[nu_t(u_y+v_x)] + d/dz[nu_t(u_z+w_x)] - div(u u) for the u-equation;
    # with u = x y z etc. second differences are exact for these polynomials
    I = (slice(1, n[0] + 1), slice(1, n[1] + 1), slice(1, n[2] + 1))
    X, Y, Z = (xf + 0 * u)[I], (yc + 0 * u)[I], (zc + 0 * u)[I]
    sgs = 2 * Y * Z + (X * Z + Y * Z) + (X * Y + Y * Z)      # nu_t derivative terms: d(nu_t)/dx_j = 1
    # + nu_t * (2 u_xx + u_yy + v_xy + u_zz + w_xz) = nu_t * (0 + 0 + Z + 0 + Y)
    sgs = sgs + (X + Y + Z) * (Z + Y)
    adv = -(2 * X * Y * Z * Y * Z + 2 * X * X * Y * Z * Z + 2 * X * X * Y * Y * Z)   # -(d(uu)/dx + d(vu)/dy + d(wu)/dz)
    # advective products are O(h^2)-accurate only; check the (exact) viscous part by subtracting the advective part of the oracle
    (au, _, _), _ = mom.mom_xyz_ad(n, 1 / dl[0], 1 / dl[1], dzci, dzfi, 0.0, u, v, w, 0 * s)
    assert np.allclose(ru - au, sgs, rtol=1e-10, atol=1e-10)
    assert np.allclose(au, adv, rtol=0, atol=0.05 * np.abs(adv).max())


@pytest.mark.parametrize("n", SIZES)
def test_strain_filter_bitexact(env, n):
    from oracle import sgs
    L, lib = env
    u, v, w = rnd_fields(n, 3, 3)
    dzc, dzf, dzci, dzfi = grid(n[2])
    dli = np.array([3.1, 2.7, 1.9])
    s0 = np.zeros_like(u); sij = [np.zeros_like(u) for _ in range(6)]
    sgs.strain_rate(n, dli, dzci, dzfi, u, v, w, s0, sij)
    pf = np.zeros_like(u); sgs.filter3d(n, u, pf)
    with Ctx(L, lib, n) as c:
        ds0 = torch.zeros(u.size, dtype=torch.float64, device="cuda")
        dsij = torch.zeros(6 * u.size, dtype=torch.float64, device="cuda")
        c.chk(lib.cales_strain_rate(c.ctx, L._ia(n), L._da(dli), dev(dzci).data_ptr(), dev(dzfi).data_ptr(), dev(u).data_ptr(),
                                    dev(v).data_ptr(), dev(w).data_ptr(), ds0.data_ptr(), dsij.data_ptr()))
        assert same(host(ds0, u.shape), s0)
        got = dsij.cpu().numpy().reshape((6,) + (u.size,))
        for m in range(6):
            assert same(got[m].reshape(u.shape, order="F"), sij[m])
        # s0 alone (sij = NULL): the two-rows-per-thread kernel strain2_k
        ds0.zero_()
        c.chk(lib.cales_strain_rate(c.ctx, L._ia(n), L._da(dli), dev(dzci).data_ptr(), dev(dzfi).data_ptr(), dev(u).data_ptr(),
                                    dev(v).data_ptr(), dev(w).data_ptr(), ds0.data_ptr(), None))
        assert same(host(ds0, u.shape), s0)
        dpf = torch.zeros(u.size, dtype=torch.float64, device="cuda")
        c.chk(lib.cales_filter3d(c.ctx, L._ia(n), dev(u).data_ptr(), dpf.data_ptr()))
        assert same(host(dpf, u.shape), pf)


@pytest.mark.parametrize("n", SIZES)
def test_reductions(env, n):
    """chkdiv / chkdt / bulk_mean: max is exact; sums differ from the sequential Fortran order only by
    re-association (tolerance 1e-13 relative to sum|x|)."""
    from oracle import ops, rk
    L, lib = env
    u, v, w, s = rnd_fields(n, 4, 4)
    s = np.abs(s) * 1e-3
    dzc, dzf, dzci, dzfi = grid(n[2])
    dl = np.array([0.31, 0.27, 0.19]); dli = 1. / dl
    lo = np.array([1, 1, 1], dtype=np.int32); hi = np.array(n, dtype=np.int32)
    with Ctx(L, lib, n) as c:
        tot, mx = C.c_double(), C.c_double()
        c.chk(lib.cales_chkdiv(c.ctx, L._ia(lo), L._ia(hi), L._da(dli), dev(dzfi).data_ptr(), dev(u).data_ptr(), dev(v).data_ptr(),
                               dev(w).data_ptr(), C.byref(tot), C.byref(mx)))
        rt, rm = ops.chkdiv_local(n, dli, dzfi, u, v, w)
        assert mx.value == rm if _ARITH[0] == "strict" else abs(mx.value - rm) <= FMA_TOL * rm
        scale = np.abs(u).sum() * dli.max() * 6
        assert abs(tot.value - rt) <= 1e-13 * scale
        out = C.c_double()
        c.chk(lib.cales_chkdt(c.ctx, L._ia(n), L._da(dl), dev(dzci).data_ptr(), dev(dzfi).data_ptr(), 1e-3, dev(s).data_ptr(),
                              dev(u).data_ptr(), dev(v).data_ptr(), dev(w).data_ptr(), C.byref(out)))
        rdt = ops.chkdt_local(n, dl, dzci, dzfi, 1e-3, s, u, v, w)
        assert out.value == rdt if _ARITH[0] == "strict" else abs(out.value - rdt) <= FMA_TOL * rdt
        gvr = dzf / dzf[1:-1].sum() / (n[0] * n[1])
        c.chk(lib.cales_bulk_mean(c.ctx, L._ia(n), dev(gvr).data_ptr(), dev(u).data_ptr(), C.byref(out)))
        ref = rk.bulk_mean_local(n, gvr, u)
        assert abs(out.value - ref) <= 1e-13 * np.abs(u[1:-1, 1:-1, 1:-1] * gvr[None, None, 1:-1]).sum()


# the ten BC/stagger combinations of find_fft (fft.f90:192-245): (bc, c_or_f, forward kind, backward kind)
KINDS = [("PP", "c", "R2HC", "HC2R"), ("NN", "c", "REDFT10", "REDFT01"), ("DD", "c", "RODFT10", "RODFT01"),
         ("ND", "c", "REDFT11", "REDFT11"), ("DN", "c", "RODFT11", "RODFT11"),
         ("PP", "f", "R2HC", "HC2R"), ("NN", "f", "REDFT00", "REDFT00"), ("DD", "f", "RODFT00", "RODFT00"),
         ("ND", "f", "REDFT10", "REDFT01"), ("DN", "f", "RODFT01", "RODFT10")]


@pytest.mark.parametrize("bc,cf,kf,kb", KINDS)
@pytest.mark.parametrize("n", [(16, 12, 3), (64, 48, 5), (96, 30, 4), (192, 256, 2), (512, 2, 2), (1024, 6, 2), (14, 22, 3), (2, 4, 2)])
def test_fft_lines_vs_fftw_definitions(env, bc, cf, kf, kb, n):
    """Each transform kind, both directions, forward and backward, against the FFTW r2r definitions
    (scipy/pocketfft).  Face-centred DD transforms n-1 points and leaves the last one alone (fft.f90:66-69).
    Tolerance: 2e-15 * log2(n) relative to max|result| (FFT round-off); the type-I kinds run a complex FFT of
    n-1 / n points through generic-radix butterflies (n-1 = 3.5.17, 7.73, 3.11.31 ...), whose O(R) sums are
    allowed 4x that."""
    from oracle import solver as osl
    L, lib = env
    rng = np.random.default_rng(5)
    a = np.asfortranarray(rng.standard_normal(n))
    ix = 1 if (bc == "DD" and cf == "f") else 0
    with Ctx(L, lib, n) as c:
        for dir_ in (0, 1):
            for backward, kind in ((0, kf), (1, kb)):
                ref = a.copy(order="F"); osl.fft(kind, n[dir_] - ix, ref, dir_)
                da = dev(a)
                c.chk(lib.cales_fft_lines(c.ctx, L._ia(n), dir_, bc.encode(), cf.encode(), backward, da.data_ptr()))
                got = host(da, a.shape)
                tol = (8e-15 if kind.endswith("00") else 2e-15) * max(1., np.log2(n[dir_])) * np.abs(ref).max()
                assert np.abs(got - ref).max() <= tol, (bc, cf, dir_, backward, np.abs(got - ref).max(), tol)
                if ix:
                    idx = [slice(None)] * 3; idx[dir_] = n[dir_] - 1
                    assert np.array_equal(got[tuple(idx)], a[tuple(idx)])


@pytest.mark.parametrize("bc,cf,kf,kb", KINDS)
def test_fft_round_trip_normfft(env, bc, cf, kf, kb):
    """bwd(fwd(x)) * normfft = x for every kind, with normfft as fftini computes it (fft.f90:99,136,142)."""
    from oracle import solver as osl
    L, lib = env
    n = (64, 32, 2)
    rng = np.random.default_rng(7)
    a = np.asfortranarray(rng.standard_normal(n))
    with Ctx(L, lib, n) as c:
        for dir_ in (0, 1):
            _, _, norm = osl.find_fft(bc, cf)
            ix = 1 if (bc == "DD" and cf == "f") else 0
            nf = 1. / (norm[0] * (n[dir_] + norm[1] - ix))
            da = dev(a)
            for backward in (0, 1):
                c.chk(lib.cales_fft_lines(c.ctx, L._ia(n), dir_, bc.encode(), cf.encode(), backward, da.data_ptr()))
            got = host(da, a.shape) * nf
            idx = [slice(None)] * 3; idx[dir_] = slice(0, n[dir_] - ix)
            assert np.abs(got[tuple(idx)] - a[tuple(idx)]).max() <= 1e-14 * np.abs(a).max()


@pytest.mark.parametrize("periodic", [0, 1])
# nxy even -> TMA kernel (partial column blocks, partial / whole level groups, periodic closure inside or outside the last box);
# nxy odd or very long columns -> general kernel
@pytest.mark.parametrize("n", [(16, 12, 10), (33, 7, 64), (64, 64, 256), (8, 8, 600), (10, 5, 37), (32, 2, 16), (32, 3, 17), (6, 7, 9),
                               (2, 1, 3), (48, 40, 129)])
def test_gaussel_bitexact(env, n, periodic):
    from oracle import solver as osl
    L, lib = env
    rng = np.random.default_rng(6)
    nx, ny, nz = n
    a = 0.5 + rng.random(nz); c_ = 0.5 + rng.random(nz); b = -(a + c_)
    lam = -np.asfortranarray(rng.random((nx, ny))) * 3
    p = np.asfortranarray(rng.standard_normal(n))
    ref = p.copy(order="F")
    (osl.gaussel_periodic if periodic else osl.gaussel)(nx, ny, nz, a, b, c_, ref, lam)
    with Ctx(L, lib, n) as c:
        dp = dev(p)
        c.chk(lib.cales_gaussel(c.ctx, nx, ny, nz, periodic, dev(a).data_ptr(), dev(b).data_ptr(), dev(c_).data_ptr(), dev(lam).data_ptr(), dp.data_ptr()))
        assert same(host(dp, p.shape), ref)
        # Thomas against a dense solve (non-periodic, well-conditioned): the reference's `+eps` pivots are O(eps)
        if not periodic and nz <= 64:
            i, j = min(1, nx - 1), min(2, ny - 1)
            A = np.diag(b + lam[i, j]) + np.diag(a[1:], -1) + np.diag(c_[:-1], 1)
            x = np.linalg.solve(A, p[i, j, :])
            assert np.allclose(ref[i, j, :], x, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("periodic", [0, 1])
@pytest.mark.parametrize("n,P", [((16, 12, 16), 2), ((33, 7, 64), 4), ((64, 64, 256), 8), ((10, 5, 37), 3), ((32, 2, 17), 2), ((6, 7, 19), 8),
                                 ((48, 40, 129), 4), ((32, 32, 600), 2)])
def test_zdist_emulated_ranks(env, n, P, periodic):
    """The distributed z solve (csrc/zdist.cu: z stays decomposed, two boundary planes per rank are exchanged instead of the
    y <-> z transposes of solver.f90:56-62) with P emulated ranks on one device against gaussel / gaussel_periodic of the
    oracle (solver.f90:82-151) on the same columns: round-off agreement (a different elimination order, both builds), for
    well-conditioned columns, nearly singular ones (|lambda| = 1e-6) and the singular mean mode (lambda = 0, compatible
    right-hand side, compared up to its additive constant as everywhere else)."""
    from oracle import solver as osl
    L, lib = env
    rng = np.random.default_rng(16)
    nx, ny, nz = n
    dz = 1. + 0.5 * rng.random(nz + 2)
    a = 1. / dz[1:nz + 1] / (0.5 * (dz[0:nz] + dz[1:nz + 1])); c_ = 1. / dz[1:nz + 1] / (0.5 * (dz[1:nz + 1] + dz[2:nz + 2])); b = -(a + c_)
    if periodic:
        a[0] = c_[-1] = 1. / dz[1] / (0.5 * (dz[1] + dz[nz])); b[0] = -(a[0] + c_[0]); b[-1] = -(a[-1] + c_[-1])   # a(1) couples to x(n): keep A.1 = 0
    else:
        b[0] += a[0]; b[-1] += c_[-1]                       # Neumann ends (initsolver.f90:156-160)
    lam = -np.asfortranarray(rng.random((nx, ny))) * 3
    lam[0, 0] = 0.; lam[min(1, nx - 1), 0] = -1e-6
    p = np.asfortranarray(rng.standard_normal(n))
    # compatible right-hand side for the singular column: orthogonal to the left null vector of A
    A0 = np.diag(b) + np.diag(a[1:], -1) + np.diag(c_[:-1], 1)
    if periodic:
        A0[0, nz - 1] += a[0]; A0[nz - 1, 0] += c_[nz - 1]
    wl = np.linalg.svd(A0)[0][:, -1]
    p[0, 0, :] -= wl * (wl @ p[0, 0, :])
    ref = p.copy(order="F")
    (osl.gaussel_periodic if periodic else osl.gaussel)(nx, ny, nz, a, b, c_, ref, lam)
    with Ctx(L, lib, n) as c:
        dp = dev(p)
        c.chk(lib.cales_zdist_emulate(c.ctx, nx, ny, nz, P, periodic, 1, dev(a).data_ptr(), dev(b).data_ptr(), dev(c_).data_ptr(), dev(lam).data_ptr(), dp.data_ptr()))
        got = host(dp, p.shape)
    # the mean mode: compare the residual-free part (constant removed); dense least squares as the reference for that column
    x0 = np.linalg.lstsq(A0, p[0, 0, :], rcond=None)[0]
    g0 = got[0, 0, :] - got[0, 0, :].mean(); x0 = x0 - x0.mean()
    assert np.abs(g0 - x0).max() <= 1e-9 * max(np.abs(x0).max(), 1e-300)
    got[0, 0, :] = ref[0, 0, :] = 0.
    # conditioning of a column ~ 1/|lambda| (in units of the coefficients): the round-off budget follows it
    cond = 1. + 1. / np.maximum(np.abs(lam), 1e-300); cond[0, 0] = 1.
    err = np.abs(got - ref).max(axis=2) / np.maximum(np.abs(ref).max(axis=2), 1e-300)
    assert (err <= 2e-14 * cond * nz).all(), float((err / cond).max())


@pytest.mark.parametrize("n", [(16, 12, 10), (67, 9, 33), (64, 64, 64)])
def test_rk_update_direct(env, n):
    """rk (src/rk.f90:17-121) kernel by kernel: two consecutive calls (the second reads the old right-hand side the first
    one left: rkpar(2) /= 0 and the new<->old swap, rk.f90:98-100), through cales_rk (mom_k + rk_update_k) and through
    cales_rk_fused (update inside the momentum kernel), each against the oracle's rk_update: identical bits (strict)."""
    from oracle import rk as ork
    from cales_b200.deck import rkcoeff
    L, lib = env
    u0, v0, w0, p, s = rnd_fields(n, 21, 5)
    s = np.abs(s)
    dzc, dzf, dzci, dzfi = grid(n[2])
    dli = np.array([3.3, 2.1, 1.7]); visc, dt = 1e-2, 3e-3
    bforce = (0.3, -0.2, 0.1)
    gvr = dzf / dzf[1:-1].sum() / (n[0] * n[1])
    noforce = L._ia(np.zeros(3, dtype=np.int32))
    # oracle: two substeps
    st = ork.RkState(n)
    ur, vr, wr = u0.copy(order="F"), v0.copy(order="F"), w0.copy(order="F")
    refs = []
    for irk in (0, 1):
        ork.rk_update(rkcoeff[irk], n, dli, dzci, dzfi, visc, dt, p, bforce, s, ur, vr, wr, st)
        refs.append((ur.copy(order="F"), vr.copy(order="F"), wr.copy(order="F")))
    I = (slice(1, -1),) * 3
    for fused in (False, True):
        with Ctx(L, lib, n) as c:
            du, dv, dw, dp, ds = map(dev, (u0, v0, w0, p, s))
            outs = [dev(np.zeros_like(u0)) for _ in range(3)]
            ddzci, ddzfi, dg = dev(dzci), dev(dzfi), dev(gvr)
            cur = [du, dv, dw]
            for irk in (0, 1):
                if fused:
                    c.chk(lib.cales_rk_fused(c.ctx, L._da(rkcoeff[irk]), L._ia(n), L._da(dli), ddzci.data_ptr(), ddzfi.data_ptr(), dg.data_ptr(), dg.data_ptr(),
                                             visc, dt, dp.data_ptr(), noforce, L._da(np.zeros(3)), L._da(bforce), ds.data_ptr(),
                                             cur[0].data_ptr(), cur[1].data_ptr(), cur[2].data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr()))
                    # the fused kernel writes the interior only: carry the (unchanged) ghost cells over for the next call
                    for a, b in zip(cur, outs):
                        full = host(a, u0.shape); new = host(b, u0.shape); full[I] = new[I]
                        b.copy_(torch.from_numpy(np.ascontiguousarray(full.ravel(order="F"))))
                    cur, outs = outs, cur
                else:
                    c.chk(lib.cales_rk(c.ctx, L._da(rkcoeff[irk]), L._ia(n), L._da(dli), ddzci.data_ptr(), ddzfi.data_ptr(), dg.data_ptr(), dg.data_ptr(),
                                       visc, dt, dp.data_ptr(), noforce, L._da(np.zeros(3)), L._da(bforce), ds.data_ptr(),
                                       cur[0].data_ptr(), cur[1].data_ptr(), cur[2].data_ptr(), None))
                for t, r in zip(cur, refs[irk]):
                    assert same(host(t, u0.shape)[I], r[I]), (fused, irk)


WM_CASES = {
    "channel_log": ("deck_channel", dict(ng=(24, 16, 20), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), 1),
    "channel_lam": ("deck_channel", dict(ng=(24, 16, 20), sgstype="smag", wall_model=True, gtype=6, gr=0., l=(12.8, 4.8, 2.), visci=43500.), -1),
    "duct_log": ("deck_duct", dict(ng=(16, 20, 24), wall_model=True), 1),
    "duct_lam": ("deck_duct", dict(ng=(16, 20, 24), wall_model=True), -1),
    "cavity": ("deck_cavity", dict(ng=(20, 16, 12)), 0),
    "tgv": ("deck_tgv", dict(ng=(16, 16, 16)), 0),
}


@pytest.mark.parametrize("case", list(WM_CASES))
@pytest.mark.parametrize("is_correc", [False, True])
def test_bounduvw_setbc_wallmodel_direct(arith, case, is_correc):
    """bounduvw (src/bound.f90:18-154) on its own: set_bc for every BC kind of the decks (P, D, N; cell- and face-centred), the
    wall-model update of the BC planes (log law WM_LOG = 1 and the parabolic WM_LAM = -1, src/wmodel.f90:288-335) and the
    `is_correc` variant that leaves the wall-normal component alone -- random interior fields, every cell of the haloed
    arrays and every wall-model BC plane compared with the oracle (the Newton iteration and pow/log: 1e-13)."""
    from conftest import need_gpu
    need_gpu()
    import oracle.param as op
    import cales_b200.deck as pd
    from oracle import bound as ob
    from oracle.main import Sim
    from cales_b200.driver import Simulation
    name, kw, mtype = WM_CASES[case]
    od, dd = getattr(op, name)(**kw), getattr(pd, name)(**kw)
    if mtype:
        od.lwm[od.lwm != 0] = mtype; dd.lwm[dd.lwm != 0] = mtype
    o = Sim(od)
    g = Simulation(dd, graph=False)
    rng = np.random.default_rng(31)
    shp = g.shape
    f = [np.asfortranarray(0.5 + 0.3 * rng.standard_normal(shp)) for _ in range(3)]
    g.set_fields(u=f[0], v=f[1], w=f[2])
    U, V, W = [f[0].copy(order="F")], [f[1].copy(order="F")], [f[2].copy(order="F")]
    g.bounduvw(True, is_correc)
    ob.bounduvw(o.world, o.cbcvel if hasattr(o, "cbcvel") else od.cbcvel, o.st, True, is_correc, U, V, W)
    tol = 0. if (arith == "strict" and not mtype) else 1e-13
    for nm, ref in (("u", U[0]), ("v", V[0]), ("w", W[0])):
        got = g.get(nm)
        assert np.abs(got - ref).max() <= tol * max(1., np.abs(ref).max()), (nm, np.abs(got - ref).max())
    if mtype:
        for bname in ("bcu", "bcv", "bcw"):
            hb = getattr(g, bname).host()
            rb = getattr(o.st[0], bname)
            for ax in "xyz":
                assert np.abs(hb[ax] - rb[ax]).max() <= 1e-12 * max(1., np.abs(rb[ax]).max()), (bname, ax)
    g.close()
