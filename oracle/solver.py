"""FFT-based direct Poisson/Helmholtz solver (CPU path of the reference).

Follows src/initsolver.f90:17-169 (eigenvalues, tridmatrix), src/fft.f90:23-143 + 192-245 (`fftini`,
`find_fft`, normfft), src/solver.f90:20-80 (`solver`), 82-179 (`gaussel`, `gaussel_periodic`,
`dgtsv_homebrewed`) and 182-233 (`solver_gaussel_z`).

Third-party arithmetic: the reference plans FFTW3 r2r transforms (unpinned `libfftw3-dev`, call
sites fft.f90:83-84,120-121 via fftw.f90:20-39).  FFTW is absent here; scipy.fft (pocketfft) is
used with the published FFTW definitions: R2HC = rfft in halfcomplex order (r0..r_{n/2},
i_{(n+1)/2-1}..i_1), HC2R = unnormalised inverse; REDFT00/10/01/11 = dct type 1/2/3/4 and
RODFT00/10/01/11 = dst type 1/2/3/4 with norm=None (FFTW manual section 4.8.3-4.8.4)."""
import numpy as np
import scipy.fft as sfft

from .param import eps, pi
from .mom import SLAB_MIN_CELLS, slab_threads


def _workers(x):
    """threads for the batched transforms / column solves of large arrays (the BASELINE-size parity cases); the lines and
    columns are independent, small arrays stay on one thread"""
    return slab_threads() if x.size >= SLAB_MIN_CELLS else 1


def _column_chunks(nx, nt):
    step = -(-nx // nt)
    return [(i0, min(nx, i0 + step)) for i0 in range(0, nx, step)]


def _dgtsv_threads(n, a, b, c, p):
    """dgtsv_homebrewed on chunks of independent columns (leading axis) in parallel"""
    nt = min(_workers(p), p.shape[0])
    if nt <= 1:
        return dgtsv_homebrewed(n, a, b, c, p)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda ch: dgtsv_homebrewed(n, a, b if b.ndim == 1 else b[ch[0]:ch[1]], c, p[ch[0]:ch[1]]),
                    _column_chunks(p.shape[0], nt)))


# ---- initsolver ------------------------------------------------------------------------------
def eigenvalues(n, cbc, c_or_f):
    """initsolver.f90:66-125 (CPU ordering: no iswap; the GPU build's re-ordering at 79-98 is a
    storage-format detail of cuFFT R2C)."""
    lam = np.zeros(n)
    bc = cbc[0] + cbc[1]
    l = np.arange(1, n + 1)
    if bc == "PP":
        lam = -2. * (1. - np.cos((2 * (l - 1)) * pi / (1. * n)))
    elif bc == "NN":
        if c_or_f == "c":
            lam = -2. * (1. - np.cos((l - 1) * pi / (1. * n)))
        else:
            lam = -2. * (1. - np.cos((l - 1) * pi / (1. * (n - 1 + 1))))
    elif bc == "DD":
        if c_or_f == "c":
            lam = -2. * (1. - np.cos(l * pi / (1. * n)))
        else:
            lam = -2. * (1. - np.cos(l * pi / (1. * (n + 1 - 1))))
            lam[n - 1] = 0.
    elif bc in ("ND", "DN"):
        lam = -2. * (1. - np.cos((2 * l - 1) * pi / (2. * n)))
    return lam


def tridmatrix(cbc, n, dzci, dzfi, c_or_f):
    """initsolver.f90:127-169.  dzci,dzfi global (0:n+1).  Returns a,b,c (n)."""
    a = np.zeros(n); c = np.zeros(n)
    k = np.arange(1, n + 1)
    if c_or_f == "c":
        a = dzfi[k] * dzci[k - 1]
        c = dzfi[k] * dzci[k]
    else:
        a = dzfi[k] * dzci[k]
        c = dzfi[k + 1] * dzci[k]
    b = -(a + c)
    factor = [{"P": 0., "D": -1., "N": 1.}[cbc[ib]] for ib in range(2)]
    if c_or_f == "c":
        b[0] = b[0] + factor[0] * a[0]
        b[n - 1] = b[n - 1] + factor[1] * c[n - 1]
    else:
        if cbc[0] == "N":
            b[0] = b[0] + factor[0] * a[0]
        if cbc[1] == "N":
            b[n - 1] = b[n - 1] + factor[1] * c[n - 1]
    return a, b, c


def find_fft(bc, c_or_f):
    """fft.f90:192-245.  Returns (kind_fwd, kind_bwd, norm) with kinds named after FFTW."""
    t = bc[0] + bc[1]
    if c_or_f == "c":
        return {"PP": ("R2HC", "HC2R", (1., 0.)), "NN": ("REDFT10", "REDFT01", (2., 0.)),
                "DD": ("RODFT10", "RODFT01", (2., 0.)), "ND": ("REDFT11", "REDFT11", (2., 0.)),
                "DN": ("RODFT11", "RODFT11", (2., 0.))}[t]
    return {"PP": ("R2HC", "HC2R", (1., 0.)), "NN": ("REDFT00", "REDFT00", (2., -1.)),
            "DD": ("RODFT00", "RODFT00", (2., 1.)), "ND": ("REDFT10", "REDFT01", (2., 0.)),
            "DN": ("RODFT01", "RODFT10", (2., 0.))}[t]


class Plan:
    """What fftini (fft.f90:23-143) returns: transform kinds, lengths and normfft."""

    def __init__(self, ng, bcxy, c_or_f):
        self.kinds = []
        normfft = 1.
        for d in range(2):
            kf, kb, norm = find_fft(bcxy[:, d], c_or_f[d])
            ix = 1 if (bcxy[0, d] + bcxy[1, d] == "DD" and c_or_f[d] == "f") else 0
            self.kinds.append((kf, kb, ng[d] - ix))
            normfft = normfft * norm[0] * (ng[d] + norm[1] - ix)
        self.normfft = normfft ** (-1)


def initsolver(ng, lo_z, hi_z, dli, dzci_g, dzfi_g, cbc, c_or_f):
    """initsolver.f90:17-64.  Returns lambdaxy (n_z(1),n_z(2)), a,b,c (ng3), Plan."""
    lambdax = eigenvalues(ng[0], cbc[:, 0], c_or_f[0]) * dli[0] ** 2
    lambday = eigenvalues(ng[1], cbc[:, 1], c_or_f[1]) * dli[1] ** 2
    lambdaxy = np.asfortranarray(lambdax[lo_z[0] - 1:hi_z[0], None] + lambday[None, lo_z[1] - 1:hi_z[1]])
    a, b, c = tridmatrix(cbc[:, 2], ng[2], dzci_g, dzfi_g, c_or_f[2])
    plan = Plan(ng, cbc[:, 0:2], c_or_f[0:2])
    return lambdaxy, a, b, c, plan


# ---- transforms (FFTW r2r kinds) ----------------------------------------------------------------
def _ax(ndim, axis, sl):
    idx = [slice(None)] * ndim
    idx[axis] = sl
    return tuple(idx)


def _r2hc(x, axis):
    n = x.shape[axis]
    X = sfft.rfft(x, axis=axis, workers=_workers(x))
    re = X.real
    im = np.flip(X.imag[_ax(X.ndim, axis, slice(1, (n + 1) // 2))], axis=axis)
    return np.concatenate([re, im], axis=axis)


def _hc2r(h, axis):
    n = h.shape[axis]
    nre = n // 2 + 1
    re = h[_ax(h.ndim, axis, slice(0, nre))]
    im_tail = np.flip(h[_ax(h.ndim, axis, slice(nre, n))], axis=axis)           # i_1 .. i_{(n+1)/2-1}
    X = np.zeros(re.shape, dtype=complex, order="F" if h.flags.f_contiguous else "C")
    X.real = re
    X.imag[_ax(h.ndim, axis, slice(1, 1 + im_tail.shape[axis]))] = im_tail
    return sfft.irfft(X, n=n, axis=axis, workers=_workers(h)) * n


_R2R = {"REDFT00": ("dct", 1), "REDFT10": ("dct", 2), "REDFT01": ("dct", 3), "REDFT11": ("dct", 4),
        "RODFT00": ("dst", 1), "RODFT10": ("dst", 2), "RODFT01": ("dst", 3), "RODFT11": ("dst", 4)}


def fft(kind, nlen, arr, axis):
    """Execute one FFTW r2r plan in place along `axis` on the first `nlen` points (fft.f90:176-190;
    the transform is one point shorter for face-centred DD, fft.f90:66-69)."""
    idx = [slice(None)] * arr.ndim
    idx[axis] = slice(0, nlen)
    x = arr[tuple(idx)]
    if kind == "R2HC":
        y = _r2hc(x, axis)
    elif kind == "HC2R":
        y = _hc2r(x, axis)
    else:
        f, t = _R2R[kind]
        y = getattr(sfft, f)(x, type=t, axis=axis, norm=None, workers=_workers(x))
    arr[tuple(idx)] = y


# ---- tridiagonal solves ---------------------------------------------------------------------------
def dgtsv_homebrewed(n, a, b, c, p):
    """solver.f90:153-179 vectorised over the leading (i,j) axes of p(..., 1:n); b may carry (i,j)
    dependence: shape (nx,ny,n) or (n,).  Operation order identical to the Fortran."""
    b = np.broadcast_to(b, p.shape[:-1] + (b.shape[-1],)) if b.ndim == 1 else b
    d = np.zeros(p.shape[:-1] + (n,))
    z = 1. / (b[..., 0] + eps)
    d[..., 0] = c[0] * z
    p[..., 0] = p[..., 0] * z
    for l in range(1, n):
        z = 1. / (b[..., l] - a[l] * d[..., l - 1] + eps)
        d[..., l] = c[l] * z
        p[..., l] = (p[..., l] - a[l] * p[..., l - 1]) * z
    for l in range(n - 2, -1, -1):
        p[..., l] = p[..., l] - d[..., l] * p[..., l + 1]


def gaussel(nx, ny, n, a, b, c, p, lambdaxy=None):
    """solver.f90:82-107.  p: (nx,ny,>=n), solved in place on 1:n."""
    if lambdaxy is not None:
        bb = b[None, None, 0:n] + lambdaxy[:, :, None]
    else:
        bb = b[0:n]
    pp = p[:, :, 0:n].copy()
    _dgtsv_threads(n, a, bb, c, pp)
    p[:, :, 0:n] = pp


def gaussel_periodic(nx, ny, n, a, b, c, p, lambdaxy=None):
    """solver.f90:109-151."""
    if lambdaxy is not None:
        bb = b[None, None, 0:n] + lambdaxy[:, :, None]
    else:
        bb = np.broadcast_to(b[None, None, 0:n], (nx, ny, n))
    p1 = p[:, :, 0:n - 1].copy()
    _dgtsv_threads(n - 1, a, bb[:, :, 0:n - 1], c, p1)
    p2 = np.zeros((nx, ny, n - 1))
    p2[:, :, 0] = -a[0]
    p2[:, :, n - 2] = -c[n - 2]
    _dgtsv_threads(n - 1, a, bb[:, :, 0:n - 1], c, p2)
    pn = (p[:, :, n - 1] - c[n - 1] * p1[:, :, 0] - a[n - 1] * p1[:, :, n - 2]) / \
         (bb[:, :, n - 1] + c[n - 1] * p2[:, :, 0] + a[n - 1] * p2[:, :, n - 2] + eps)
    p[:, :, n - 1] = pn
    p[:, :, 0:n - 1] = p1 + p2 * pn[:, :, None]


# ---- the solver proper -------------------------------------------------------------------------------
def solver(world, plan, lambdaxy_l, a, b, c, bc, c_or_f, P):
    """solver.f90:20-80 on all ranks.  P: per-rank haloed arrays; lambdaxy_l: per-rank (n_z(1),n_z(2)).
    The 2decomp transposes are emulated by re-slicing the assembled global array (decomp.World)."""
    ip = world.ipencil
    which = {1: "x", 2: "y", 3: "z"}[ip]
    loc = [np.asfortranarray(p[1:-1, 1:-1, 1:-1].copy()) for p in P]
    px = loc if ip == 1 else world.transpose(loc, which, "x")
    for a_ in px:
        fft(plan.kinds[0][0], plan.kinds[0][2], a_, 0)
    py = world.transpose(px, "x", "y")
    for a_ in py:
        fft(plan.kinds[1][0], plan.kinds[1][2], a_, 1)
    pz = world.transpose(py, "y", "z")
    q = 1 if (c_or_f[2] == "f" and bc[1, 2] == "D") else 0
    for r, a_ in zip(world.ranks, pz):
        nz = r.n_z
        if bc[0, 2] + bc[1, 2] == "PP":
            gaussel_periodic(nz[0], nz[1], nz[2] - q, a, b, c, a_, lambdaxy_l[r.id])
        else:
            gaussel(nz[0], nz[1], nz[2] - q, a, b, c, a_, lambdaxy_l[r.id])
    py = world.transpose(pz, "z", "y")
    for a_ in py:
        fft(plan.kinds[1][1], plan.kinds[1][2], a_, 1)
    px = world.transpose(py, "y", "x")
    for a_ in px:
        fft(plan.kinds[0][1], plan.kinds[0][2], a_, 0)
    out = px if ip == 1 else world.transpose(px, "x", which)
    for p, o in zip(P, out):
        p[1:-1, 1:-1, 1:-1] = o * plan.normfft


def solver_gaussel_z(world, a, b, c, bcz, c_or_f, P):
    """solver.f90:182-233: z-only implicit solve (no lambdaxy)."""
    ip = world.ipencil
    which = {1: "x", 2: "y", 3: "z"}[ip]
    loc = [np.asfortranarray(p[1:-1, 1:-1, 1:-1].copy()) for p in P]
    pz = loc if ip == 3 else world.transpose(loc, which, "z")
    q = 1 if (c_or_f[2] == "f" and bcz[1] == "D") else 0
    for r, a_ in zip(world.ranks, pz):
        nz = r.n_z
        if bcz[0] + bcz[1] == "PP":
            gaussel_periodic(nz[0], nz[1], nz[2] - q, a, b, c, a_)
        else:
            gaussel(nz[0], nz[1], nz[2] - q, a, b, c, a_)
    out = pz if ip == 3 else world.transpose(pz, "z", which)
    for p, o in zip(P, out):
        p[1:-1, 1:-1, 1:-1] = o
