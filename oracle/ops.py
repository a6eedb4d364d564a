"""Pressure-correction stencils and step checks.  Follows src/fillps.f90:14-48, src/correc.f90:14-68,
src/updatep.f90:14-49, src/chkdiv.f90:16-52, src/chkdt.f90:17-99."""
import numpy as np

from .param import eps
from .mom import SLAB_MIN_CELLS, run_slabs


def _slabs(n, nlev, fn):
    """fn(k0, nb) over the level offsets 0 .. nlev-1: in one piece on small grids, slab by slab on a thread pool on the
    BASELINE-size ones (per-cell work: the same bits either way, see oracle/mom.py)"""
    if n[0] * n[1] * n[2] < SLAB_MIN_CELLS:
        fn(0, nlev)
    else:
        run_slabs(nlev, n[0] * n[1], fn)


def fillps(n, dli, dzfi, dti, u, v, w, p):
    """fillps.f90:33-47."""
    n1, n2, n3 = n
    dtidxi = dti * dli[0]
    dtidyi = dti * dli[1]
    def job(k0, nb):
        K = slice(k0 + 1, k0 + nb + 1); Km = slice(k0, k0 + nb)
        I = (slice(1, n1 + 1), slice(1, n2 + 1), K)
        dzfi_k = dzfi[K][None, None, :]
        p[I] = ((w[I] - w[1:n1 + 1, 1:n2 + 1, Km]) * dti * dzfi_k +
                (v[I] - v[1:n1 + 1, 0:n2, K]) * dtidyi +
                (u[I] - u[0:n1, 1:n2 + 1, K]) * dtidxi)
    _slabs(n, n3, job)


def correc(n, dli, dzci, dt, p, u, v, w):
    """correc.f90:41-67: ghost rows included (i=0:n1, j=0:n2+1, k=0:n3+1 for u, ...)."""
    n1, n2, n3 = n
    factori = dt * dli[0]
    factorj = dt * dli[1]
    def job_uv(k0, nb):
        K = slice(k0, k0 + nb)
        u[0:n1 + 1, :, K] = u[0:n1 + 1, :, K] - factori * (p[1:n1 + 2, :, K] - p[0:n1 + 1, :, K])
        v[:, 0:n2 + 1, K] = v[:, 0:n2 + 1, K] - factorj * (p[:, 1:n2 + 2, K] - p[:, 0:n2 + 1, K])

    def job_w(k0, nb):
        K = slice(k0, k0 + nb); Kp = slice(k0 + 1, k0 + nb + 1)
        dzci_k = dzci[K][None, None, :]
        w[:, :, K] = w[:, :, K] - dt * dzci_k * (p[:, :, Kp] - p[:, :, K])
    _slabs(n, n3 + 2, job_uv)
    _slabs(n, n3 + 1, job_w)


def updatep(n, dli, dzci, dzfi, alpha, pp, p, impdiff=False, impdiff_1d=False):
    """updatep.f90:26-48."""
    n1, n2, n3 = n
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    if impdiff:
        dxi, dyi = dli[0], dli[1]
        c = pp[I]
        k = np.arange(1, n3 + 1)
        lap = ((pp[1:n1 + 1, 1:n2 + 1, 2:n3 + 2] - c) * dzci[k][None, None, :] -
               (c - pp[1:n1 + 1, 1:n2 + 1, 0:n3]) * dzci[k - 1][None, None, :]) * dzfi[k][None, None, :]
        if not impdiff_1d:
            lap = (pp[2:n1 + 2, 1:n2 + 1, 1:n3 + 1] - 2. * c + pp[0:n1, 1:n2 + 1, 1:n3 + 1]) * (dxi ** 2) + \
                  (pp[1:n1 + 1, 2:n2 + 2, 1:n3 + 1] - 2. * c + pp[1:n1 + 1, 0:n2, 1:n3 + 1]) * (dyi ** 2) + \
                  lap
        p[I] = p[I] + c + alpha * lap
    else:
        p[I] = p[I] + pp[I]


def chkdiv_local(n, dli, dzfi, u, v, w):
    """chkdiv.f90:27-47 (rank-local part): sequential sum, i fastest.  Returns (divtot, divmax)."""
    n1, n2, n3 = n
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    dzfi_k = dzfi[1:n3 + 1][None, None, :]
    div = (w[I] - w[1:n1 + 1, 1:n2 + 1, 0:n3]) * dzfi_k + \
          (v[I] - v[1:n1 + 1, 0:n2, 1:n3 + 1]) * dli[1] + \
          (u[I] - u[0:n1, 1:n2 + 1, 1:n3 + 1]) * dli[0]
    divtot = float(np.cumsum(div.ravel(order="F"))[-1])
    divmax = float(np.max(np.abs(div)))
    return divtot, divmax


def chkdt_local(n, dl, dzci, dzfi, visc, visct, u, v, w, impdiff=False, impdiff_1d=False):
    """chkdt.f90:40-97 (rank-local part).  Returns dtmax."""
    n1, n2, n3 = n
    dxi = 1.0 / dl[0]
    dyi = 1.0 / dl[1]
    dl2i = dxi * dxi + dyi * dyi

    res = []            # (dti, dtid) per slab of levels; the maximum of the slab maxima is the maximum (exact)

    def job(k0, nb):
        def S(a, di, dj, dk):
            return a[1 + di:n1 + 1 + di, 1 + dj:n2 + 1 + dj, k0 + 1 + dk:k0 + nb + 1 + dk]
        k = np.arange(k0 + 1, k0 + nb + 1)
        dzfi_k = dzfi[k][None, None, :]
        dzci_k = dzci[k][None, None, :]
        ux = np.abs(S(u, 0, 0, 0))
        vx = 0.25 * np.abs(S(v, 0, 0, 0) + S(v, 0, -1, 0) + S(v, 1, 0, 0) + S(v, 1, -1, 0))
        wx = 0.25 * np.abs(S(w, 0, 0, 0) + S(w, 0, 0, -1) + S(w, 1, 0, 0) + S(w, 1, 0, -1))
        uy = 0.25 * np.abs(S(u, 0, 0, 0) + S(u, 0, 1, 0) + S(u, -1, 1, 0) + S(u, -1, 0, 0))
        vy = np.abs(S(v, 0, 0, 0))
        wy = 0.25 * np.abs(S(w, 0, 0, 0) + S(w, 0, 1, 0) + S(w, 0, 1, -1) + S(w, 0, 0, -1))
        uz = 0.25 * np.abs(S(u, 0, 0, 0) + S(u, -1, 0, 0) + S(u, -1, 0, 1) + S(u, 0, 0, 1))
        vz = 0.25 * np.abs(S(v, 0, 0, 0) + S(v, 0, -1, 0) + S(v, 0, -1, 1) + S(v, 0, 0, 1))
        wz = np.abs(S(w, 0, 0, 0))
        dtix = ux * dxi + vx * dyi + wx * dzfi_k
        dtiy = uy * dxi + vy * dyi + wy * dzfi_k
        dtiz = uz * dxi + vz * dyi + wz * dzci_k
        dti = max(0.0, float(dtix.max()), float(dtiy.max()), float(dtiz.max()))
        viscx = 0.5 * (S(visct, 0, 0, 0) + S(visct, 1, 0, 0))
        viscy = 0.5 * (S(visct, 0, 0, 0) + S(visct, 0, 1, 0))
        viscz = 0.5 * (S(visct, 0, 0, 0) + S(visct, 0, 0, 1))
        dtidx = viscx * (dl2i + dzfi_k * dzfi_k)
        dtidy = viscy * (dl2i + dzfi_k * dzfi_k)
        dtidz = viscz * (dl2i + dzci_k * dzci_k)
        if impdiff and not impdiff_1d:
            pass
        else:
            dtidx = dtidx + visc * dl2i
            dtidy = dtidy + visc * dl2i
            dtidz = dtidz + visc * dl2i
            if not impdiff_1d:
                dtidx = dtidx + visc * (dzfi_k * dzfi_k)
                dtidy = dtidy + visc * (dzfi_k * dzfi_k)
                dtidz = dtidz + visc * (dzci_k * dzci_k)
        dtid = max(0.0, float(dtidx.max()), float(dtidy.max()), float(dtidz.max()))
        res.append((dti, dtid))
    _slabs(n, n3, job)
    dti = max(r[0] for r in res)
    dtid = max(r[1] for r in res)
    if dti == 0.0:
        dti = 1.0
    if dtid == 0.0:
        dtid = eps
    return min(0.4125 / dtid, 1.732 / dti)
