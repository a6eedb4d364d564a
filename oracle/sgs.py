"""Sub-grid-scale eddy viscosity.  Follows src/sgs.f90: cmpt_sgs 21-386 ('none', 'smag' 69-152,
'dsmag' 153-380), ave1d_channel 433-538 (+ave0d_dit 388-431, ave2d_duct 540-614), filter3d 616-680,
extrapolate 682-767, cmpt_alph2 769-822, filter2d 824-848, interpolate 850-870, strain_rate
1019-1110.  The dsmag plane averaging geometry is the reference's hard-wired `#define _CHANNEL`
(sgs.f90:8) unless `ave` says otherwise."""
import numpy as np

from . import bound as bnd
from .param import big, c_smag
from .mom import SLAB_MIN_CELLS, run_slabs


def _S(a, n):
    n1, n2, n3 = n
    return lambda di, dj, dk: a[1 + di:n1 + 1 + di, 1 + dj:n2 + 1 + dj, 1 + dk:n3 + 1 + dk]


def strain_rate(n, dli, dzci, dzfi, u, v, w, s0, sij=None):
    """sgs.f90:1019-1110.  Writes the interior of s0 (and sij[m], m=0..5 = s11,s22,s33,s12,s13,s23).
    Large grids: slab by slab in z on a thread pool (elementwise work: the same bits, see oracle/mom.py)."""
    n1, n2, n3 = n
    if n1 * n2 * n3 < SLAB_MIN_CELLS:
        return _strain_block(n, dli, dzci, dzfi, u, v, w, s0, sij)

    def job(k0, nb):
        z = slice(k0, k0 + nb + 2)
        _strain_block((n1, n2, nb), dli, dzci[z], dzfi[z], u[:, :, z], v[:, :, z], w[:, :, z], s0[:, :, z],
                      None if sij is None else [a[:, :, z] for a in sij])
    run_slabs(n3, n1 * n2, job)


def _strain_block(n, dli, dzci, dzfi, u, v, w, s0, sij=None):
    n1, n2, n3 = n
    dxi, dyi = dli[0], dli[1]
    U, V, W = _S(u, n), _S(v, n), _S(w, n)
    k = np.arange(1, n3 + 1)
    dzci_k = dzci[k][None, None, :]; dzci_km = dzci[k - 1][None, None, :]; dzfi_k = dzfi[k][None, None, :]
    u_mcm = U(-1, 0, -1); u_ccm = U(0, 0, -1); u_mmc = U(-1, -1, 0); u_cmc = U(0, -1, 0); u_mcc = U(-1, 0, 0)
    u_ccc = U(0, 0, 0); u_mpc = U(-1, 1, 0); u_cpc = U(0, 1, 0); u_mcp = U(-1, 0, 1); u_ccp = U(0, 0, 1)
    v_cmm = V(0, -1, -1); v_ccm = V(0, 0, -1); v_mmc = V(-1, -1, 0); v_cmc = V(0, -1, 0); v_pmc = V(1, -1, 0)
    v_mcc = V(-1, 0, 0); v_ccc = V(0, 0, 0); v_pcc = V(1, 0, 0); v_cmp = V(0, -1, 1); v_ccp = V(0, 0, 1)
    w_cmm = W(0, -1, -1); w_mcm = W(-1, 0, -1); w_ccm = W(0, 0, -1); w_pcm = W(1, 0, -1); w_cpm = W(0, 1, -1)
    w_cmc = W(0, -1, 0); w_mcc = W(-1, 0, 0); w_ccc = W(0, 0, 0); w_pcc = W(1, 0, 0); w_cpc = W(0, 1, 0)
    s11 = (u_ccc - u_mcc) * dxi
    s22 = (v_ccc - v_cmc) * dyi
    s33 = (w_ccc - w_ccm) * dzfi_k
    s12 = .125 * ((u_cpc - u_ccc) * dyi + (v_pcc - v_ccc) * dxi +
                  (u_ccc - u_cmc) * dyi + (v_pmc - v_cmc) * dxi +
                  (u_mpc - u_mcc) * dyi + (v_ccc - v_mcc) * dxi +
                  (u_mcc - u_mmc) * dyi + (v_cmc - v_mmc) * dxi)
    s13 = .125 * ((u_ccp - u_ccc) * dzci_k + (w_pcc - w_ccc) * dxi +
                  (u_ccc - u_ccm) * dzci_km + (w_pcm - w_ccm) * dxi +
                  (u_mcp - u_mcc) * dzci_k + (w_ccc - w_mcc) * dxi +
                  (u_mcc - u_mcm) * dzci_km + (w_ccm - w_mcm) * dxi)
    s23 = .125 * ((v_ccp - v_ccc) * dzci_k + (w_cpc - w_ccc) * dyi +
                  (v_ccc - v_ccm) * dzci_km + (w_cpm - w_ccm) * dyi +
                  (v_cmp - v_cmc) * dzci_k + (w_ccc - w_cmc) * dyi +
                  (v_cmc - v_cmm) * dzci_km + (w_ccm - w_cmm) * dyi)
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    s0[I] = np.sqrt(2. * (s11 ** 2 + s22 ** 2 + s33 ** 2 + 2. * (s12 ** 2 + s13 ** 2 + s23 ** 2)))
    if sij is not None:
        for m, s in enumerate((s11, s22, s33, s12, s13, s23)):
            sij[m][I] = s


def filter3d(n, p, pf):
    """sgs.f90:616-680: 27-point trapezoidal top-hat, weights 8/4/2/1 over 64, summed in the
    reference's order."""
    n1, n2, n3 = n
    P = _S(p, n)
    c = P(0, 0, 0)
    f = (P(-1, 0, 0) + P(0, -1, 0) + P(0, 0, -1) + P(1, 0, 0) + P(0, 1, 0) + P(0, 0, 1))
    e = (P(0, -1, -1) + P(-1, 0, -1) + P(-1, -1, 0) +
         P(0, 1, -1) + P(1, 0, -1) + P(1, -1, 0) +
         P(0, -1, 1) + P(-1, 0, 1) + P(-1, 1, 0) +
         P(0, 1, 1) + P(1, 0, 1) + P(1, 1, 0))
    v = (P(-1, -1, -1) + P(1, -1, -1) + P(-1, 1, -1) + P(1, 1, -1) +
         P(-1, -1, 1) + P(1, -1, 1) + P(-1, 1, 1) + P(1, 1, 1))
    pf[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = (8. * c + 4. * f + 2. * e + 1. * v) / 64.


def filter2d(n, p, pf):
    """sgs.f90:824-848."""
    n1, n2, n3 = n
    P = _S(p, n)
    pf[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = (4. * P(0, 0, 0) +
                                       2. * (P(-1, 0, 0) + P(0, -1, 0) + P(1, 0, 0) + P(0, 1, 0)) +
                                       1. * (P(-1, -1, 0) + P(1, -1, 0) + P(-1, 1, 0) + P(1, 1, 0))) / 16.


def interpolate(n, u, v, w, uc, vc, wc):
    """sgs.f90:850-870."""
    n1, n2, n3 = n
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    uc[I] = 0.5 * (u[I] + u[0:n1, 1:n2 + 1, 1:n3 + 1])
    vc[I] = 0.5 * (v[I] + v[1:n1 + 1, 0:n2, 1:n3 + 1])
    wc[I] = 0.5 * (w[I] + w[1:n1 + 1, 1:n2 + 1, 0:n3])


def extrapolate(n, is_bound, dzci, p, iface, cbc=None, lwm=None):
    """sgs.f90:682-767.  iface: 0 (cell centre) or 1,2,3 (the face direction of the component)."""
    n1, n2, n3 = n
    dzc = 1. / dzci
    is_done = np.zeros((2, 3), dtype=bool)
    if cbc is not None:
        factor0 = 1.
        factor1 = 1.
        for d in range(3):
            for ib in range(2):
                is_done[ib, d] = is_bound[ib, d] and cbc[ib, d, d] == "D" and iface != d + 1
    elif lwm is not None:
        factor0 = dzc[0] * dzci[1]
        factor1 = dzc[n3] * dzci[n3 - 1]
        for d in range(3):
            for ib in range(2):
                is_done[ib, d] = is_bound[ib, d] and lwm[ib, d] != 0 and iface != d + 1
    if is_done[0, 0]:
        p[0, :, :] = 2. * p[1, :, :] - p[2, :, :]
    if is_done[1, 0]:
        p[n1 + 1, :, :] = 2. * p[n1, :, :] - p[n1 - 1, :, :]
    if is_done[0, 1]:
        p[:, 0, :] = 2. * p[:, 1, :] - p[:, 2, :]
    if is_done[1, 1]:
        p[:, n2 + 1, :] = 2. * p[:, n2, :] - p[:, n2 - 1, :]
    if is_done[0, 2]:
        p[:, :, 0] = (1. + factor0) * p[:, :, 1] - factor0 * p[:, :, 2]
    if is_done[1, 2]:
        p[:, :, n3 + 1] = (1. + factor1) * p[:, :, n3] - factor1 * p[:, :, n3 - 1]


def cmpt_alph2(n, is_bound, cbc, filter_2d=False):
    """sgs.f90:769-822."""
    n1, n2, n3 = n
    alph2 = np.full((n1 + 2, n2 + 2, n3 + 2), 4.00, order="F")
    if filter_2d:
        alph2[:] = 2.52
        return alph2
    if is_bound[0, 0] and cbc[0, 0, 0] == "D": alph2[1, :, :] = 2.52
    if is_bound[1, 0] and cbc[1, 0, 0] == "D": alph2[n1, :, :] = 2.52
    if is_bound[0, 1] and cbc[0, 1, 1] == "D": alph2[:, 1, :] = 2.52
    if is_bound[1, 1] and cbc[1, 1, 1] == "D": alph2[:, n2, :] = 2.52
    if is_bound[0, 2] and cbc[0, 2, 2] == "D": alph2[:, :, 1] = 2.52
    if is_bound[1, 2] and cbc[1, 2, 2] == "D": alph2[:, :, n3] = 2.52
    return alph2


def _seqsum(a):
    """Sequential (loop-order, i fastest) sum, as the Fortran accumulation loops."""
    t = np.asarray(a).ravel(order="F")
    return float(np.cumsum(t)[-1]) if t.size else 0.0


def ave1d_channel(world, st, idir, P):
    """sgs.f90:433-538, idir=3 (the only call, sgs.f90:363-364); others kept for completeness."""
    ng = world.ng
    s0 = st[0]
    l, dl = s0.l, s0.dl
    p1d_all = []
    for r, s in zip(world.ranks, st):
        p = P[r.id]
        n1, n2, n3 = r.n
        p1d = np.zeros(ng[idir])
        if idir == 2:
            gar = dl[0] * dl[1] / (l[0] * l[1])
            for k in range(1, n3 + 1):
                p1d[r.lo[2] - 1 + k - 1] = _seqsum(p[1:n1 + 1, 1:n2 + 1, k]) * gar
        elif idir == 1:
            gar = dl[0] / (l[0] * l[2])
            for j in range(1, n2 + 1):
                # loop order k outer, i inner
                t = (p[1:n1 + 1, j, 1:n3 + 1] * s.dzf[None, 1:n3 + 1])
                p1d[r.lo[1] - 1 + j - 1] = _seqsum(t) * gar
        else:
            gar = dl[1] / (l[1] * l[2])
            for i in range(1, n1 + 1):
                t = (p[i, 1:n2 + 1, 1:n3 + 1] * s.dzf[None, 1:n3 + 1])
                p1d[r.lo[0] - 1 + i - 1] = _seqsum(t) * gar
        p1d_all.append(p1d)
    p1d = world.allreduce_sum(p1d_all)
    for r in world.ranks:
        p = P[r.id]
        n1, n2, n3 = r.n
        if idir == 2:
            for k in range(1, n3 + 1):
                p[:, :, k] = p1d[r.lo[2] - 1 + k - 1]
        elif idir == 1:
            for j in range(1, n2 + 1):
                p[:, j, :] = p1d[r.lo[1] - 1 + j - 1]
        else:
            for i in range(1, n1 + 1):
                p[i, :, :] = p1d[r.lo[0] - 1 + i - 1]


def ave0d_dit(world, st, P):
    """sgs.f90:388-431."""
    s0 = st[0]
    l, dl = s0.l, s0.dl
    gar = dl[0] * dl[1] / (l[0] * l[1])
    parts = []
    for r, s in zip(world.ranks, st):
        n1, n2, n3 = r.n
        t = P[r.id][1:n1 + 1, 1:n2 + 1, 1:n3 + 1] * gar * s.dzf[None, None, 1:n3 + 1] / l[2]
        parts.append(_seqsum(t))
    p0d = world.allreduce_sum(parts)
    for r in world.ranks:
        n1, n2, n3 = r.n
        P[r.id][1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = p0d


def ave2d_duct(world, st, idir, P):
    """sgs.f90:540-614 (idir = streamwise direction, 0 or 1 here)."""
    ng = world.ng
    s0 = st[0]
    l, dl = s0.l, s0.dl
    parts = []
    for r in world.ranks:
        n1, n2, n3 = r.n
        p = P[r.id]
        if idir == 0:
            p2d = np.zeros((ng[1], ng[2]))
            acc = np.cumsum(p[1:n1 + 1, 1:n2 + 1, 1:n3 + 1], axis=0)[-1]
            p2d[r.lo[1] - 1:r.hi[1], r.lo[2] - 1:r.hi[2]] = acc
        else:
            p2d = np.zeros((ng[0], ng[2]))
            acc = np.cumsum(p[1:n1 + 1, 1:n2 + 1, 1:n3 + 1], axis=1)[:, -1, :]
            p2d[r.lo[0] - 1:r.hi[0], r.lo[2] - 1:r.hi[2]] = acc
        parts.append(p2d)
    p2d = world.allreduce_sum(parts)
    gar = dl[idir] / l[idir]
    for r in world.ranks:
        n1, n2, n3 = r.n
        p = P[r.id]
        if idir == 0:
            p[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = (p2d[r.lo[1] - 1:r.hi[1], r.lo[2] - 1:r.hi[2]] * gar)[None, :, :]
        else:
            p[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = (p2d[r.lo[0] - 1:r.hi[0], r.lo[2] - 1:r.hi[2]] * gar)[:, None, :]


class SgsState:
    """The `save`d work arrays of cmpt_sgs (sgs.f90:51-57)."""

    def __init__(self):
        self.is_first = True


def _smag_local(r, s, deck, cbcvel, u, v, w, visct):
    """sgs.f90:69-152 on one rank."""
    n = r.n
    n1, n2, n3 = n
    g = s.sgs
    if g.is_first:
        g.is_first = False
        g.s0 = np.zeros((n1 + 2, n2 + 2, n3 + 2), order="F")
        g.is_wall = np.zeros(6)
        for d in range(3):
            for ib in range(2):
                if r.is_bound[ib, d] and cbcvel[ib, d, d] == "D":
                    g.is_wall[2 * d + ib] = 1.
    wk = [u.copy(order="F"), v.copy(order="F"), w.copy(order="F")]
    for c in range(3):
        extrapolate(n, r.is_bound, s.dzci, wk[c], iface=c + 1, lwm=s.lwm)
    strain_rate(n, s.dli, s.dzci, s.dzfi, wk[0], wk[1], wk[2], g.s0)
    dl, l, visc = s.dl, s.l, s.visc
    dxi, dyi = s.dli[0], s.dli[1]
    visci = 1. / visc
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    kk = np.arange(1, n3 + 1)
    dele = (dl[0] * dl[1] * s.dzf[kk]) ** (1. / 3.)
    if np.sum(g.is_wall) == 0:
        fd = np.ones((n1, n2, n3))
        visct[I] = (c_smag * dele[None, None, :] * fd) ** 2 * g.s0[I]
        return
    i = np.arange(1, n1 + 1); j = np.arange(1, n2 + 1)
    J = slice(1, n2 + 1); Jm = slice(0, n2)
    Ii = slice(1, n1 + 1); Im = slice(0, n1)

    def mag(t1, t2, f):
        return np.sqrt(t1 * t1 + t2 * t2) * f

    def vandriest(k0, nb):
        """levels k0+1 .. k0+nb (every operation is per cell: any k range gives the same bits)"""
        kb = np.arange(k0 + 1, k0 + nb + 1)
        K = slice(k0 + 1, k0 + nb + 1); Km = slice(k0, k0 + nb)
        shp = (n1, n2, nb)
        dw = np.empty((6,) + shp)
        dw[0] = (dl[0] * (i - 0.5))[:, None, None]
        dw[1] = (dl[0] * (n1 - i + 0.5))[:, None, None]
        dw[2] = (dl[1] * (j - 0.5))[None, :, None]
        dw[3] = (dl[1] * (n2 - j + 0.5))[None, :, None]
        dw[4] = s.zc[kb][None, None, :]
        dw[5] = (l[2] - s.zc[kb])[None, None, :]
        iw = g.is_wall[:, None, None, None]
        dw = dw * iw + big * (1. - iw)
        loc = np.argmin(dw, axis=0)                         # minloc: first minimum
        dw_min = np.take_along_axis(dw, loc[None], axis=0)[0]
        tw = np.zeros((6,) + shp)
        t1 = v[1, J, K] - v[0, J, K] + v[1, Jm, K] - v[0, Jm, K]
        t2 = w[1, J, K] - w[0, J, K] + w[1, J, Km] - w[0, J, Km]
        tw[0] = mag(t1, t2, dxi)[None, :, :]
        t1 = v[n1, J, K] - v[n1 + 1, J, K] + v[n1, Jm, K] - v[n1 + 1, Jm, K]
        t2 = w[n1, J, K] - w[n1 + 1, J, K] + w[n1, J, Km] - w[n1 + 1, J, Km]
        tw[1] = mag(t1, t2, dxi)[None, :, :]
        t1 = u[Ii, 1, K] - u[Ii, 0, K] + u[Im, 1, K] - u[Im, 0, K]
        t2 = w[Ii, 1, K] - w[Ii, 0, K] + w[Ii, 1, Km] - w[Ii, 0, Km]
        tw[2] = mag(t1, t2, dyi)[:, None, :]
        t1 = u[Ii, n2, K] - u[Ii, n2 + 1, K] + u[Im, n2, K] - u[Im, n2 + 1, K]
        t2 = w[Ii, n2, K] - w[Ii, n2 + 1, K] + w[Ii, n2, Km] - w[Ii, n2 + 1, Km]
        tw[3] = mag(t1, t2, dyi)[:, None, :]
        t1 = u[Ii, J, 1] - u[Ii, J, 0] + u[Im, J, 1] - u[Im, J, 0]
        t2 = v[Ii, J, 1] - v[Ii, J, 0] + v[Ii, Jm, 1] - v[Ii, Jm, 0]
        tw[4] = mag(t1, t2, s.dzci[0])[:, :, None]
        t1 = u[Ii, J, n3] - u[Ii, J, n3 + 1] + u[Im, J, n3] - u[Im, J, n3 + 1]
        t2 = v[Ii, J, n3] - v[Ii, J, n3 + 1] + v[Ii, Jm, n3] - v[Ii, Jm, n3 + 1]
        tw[5] = mag(t1, t2, s.dzci[n3])[:, :, None]
        tauw_s = np.take_along_axis(tw, loc[None], axis=0)[0]
        tauw_s = 0.5 * visc * tauw_s
        dw_plus = dw_min * np.sqrt(tauw_s) * visci
        fd = 1. - np.exp(-dw_plus / 25.)
        visct[Ii, J, K] = (c_smag * dele[None, None, k0:k0 + nb] * fd) ** 2 * g.s0[Ii, J, K]
    if n1 * n2 * n3 < SLAB_MIN_CELLS:
        vandriest(0, n3)
    else:
        run_slabs(n3, n1 * n2, vandriest)


def cmpt_sgs(world, st, deck, cbcvel, U, V, W, VISCT, ave="channel", filter_2d=False):
    """sgs.f90:21-386 on all ranks (cbcvel: the wall-model-adjusted one from initbc)."""
    sgstype = deck.sgstype.strip()
    if sgstype == "none":
        for r, s in zip(world.ranks, st):
            if s.sgs.is_first:
                s.sgs.is_first = False
                VISCT[r.id][:] = 0.
        return
    if sgstype == "smag":
        for r, s in zip(world.ranks, st):
            _smag_local(r, s, deck, cbcvel, U[r.id], V[r.id], W[r.id], VISCT[r.id])
        return
    if sgstype != "dsmag":
        raise ValueError("unknown SGS model " + sgstype)
    R = world.ranks
    cbcsgs = deck.cbcsgs
    for r, s in zip(R, st):
        g = s.sgs
        if g.is_first:
            g.is_first = False
            z = lambda: np.zeros((r.n[0] + 2, r.n[1] + 2, r.n[2] + 2), order="F")
            g.uc, g.vc, g.wc, g.uf, g.vf, g.wf, g.s0 = z(), z(), z(), z(), z(), z(), z()
            g.wk = [z() for _ in range(6)]; g.sij = [z() for _ in range(6)]; g.mij = [z() for _ in range(6)]
            g.alph2 = cmpt_alph2(r.n, r.is_bound, cbcvel, filter_2d)
    G = [s.sgs for s in st]
    for r, s, g in zip(R, st, G):                           # sgs.f90:173-185
        g.wk[0][:] = U[r.id]; g.wk[1][:] = V[r.id]; g.wk[2][:] = W[r.id]
        for c in range(3):
            extrapolate(r.n, r.is_bound, s.dzci, g.wk[c], iface=c + 1, lwm=s.lwm)
        strain_rate(r.n, s.dli, s.dzci, s.dzfi, g.wk[0], g.wk[1], g.wk[2], g.s0, g.sij)
        VISCT[r.id][:] = g.s0
    bnd.boundp(world, cbcsgs, st, "bcs", [g.s0 for g in G])   # sgs.f90:191-197
    for m in range(6):
        bnd.boundp(world, cbcsgs, st, "bcs", [g.sij[m] for g in G])
    filt = filter2d if filter_2d else filter3d
    for r, s, g in zip(R, st, G):                           # sgs.f90:198-235
        for m in range(6):
            g.wk[m][:] = g.s0 * g.sij[m]
        if not filter_2d:
            for m in range(6):
                extrapolate(r.n, r.is_bound, s.dzci, g.wk[m], iface=0, cbc=cbcvel)
        for m in range(6):
            filt(r.n, g.wk[m], g.mij[m])
        if not filter_2d:
            g.wk[0][:] = U[r.id]; g.wk[1][:] = V[r.id]; g.wk[2][:] = W[r.id]
            for c in range(3):
                extrapolate(r.n, r.is_bound, s.dzci, g.wk[c], iface=c + 1, cbc=cbcvel)
            filt(r.n, g.wk[0], g.uf); filt(r.n, g.wk[1], g.vf); filt(r.n, g.wk[2], g.wf)
        else:
            filt(r.n, U[r.id], g.uf); filt(r.n, V[r.id], g.vf); filt(r.n, W[r.id], g.wf)
    bnd.bounduvw(world, cbcvel, st, False, False, [g.uf for g in G], [g.vf for g in G], [g.wf for g in G],
                 bcu=[s.bcuf for s in st], bcv=[s.bcvf for s in st], bcw=[s.bcwf for s in st])   # sgs.f90:256-257
    for r, s, g in zip(R, st, G):                           # sgs.f90:258-272
        extrapolate(r.n, r.is_bound, s.dzci, g.uf, iface=1, lwm=s.lwm)
        extrapolate(r.n, r.is_bound, s.dzci, g.vf, iface=2, lwm=s.lwm)
        extrapolate(r.n, r.is_bound, s.dzci, g.wf, iface=3, lwm=s.lwm)
        strain_rate(r.n, s.dli, s.dzci, s.dzfi, g.uf, g.vf, g.wf, g.s0, g.sij)
        n1, n2, n3 = r.n
        I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
        for m in range(6):
            g.mij[m][I] = 2. * (g.mij[m][I] - g.alph2[I] * g.s0[I] * g.sij[m][I])
        interpolate(r.n, U[r.id], V[r.id], W[r.id], g.uc, g.vc, g.wc)   # sgs.f90:277
    for name in ("uc", "vc", "wc"):                         # sgs.f90:280-282
        bnd.boundp(world, cbcsgs, st, "bcs", [getattr(g, name) for g in G])
    for r, s, g in zip(R, st, G):                           # sgs.f90:283-358
        lij = g.sij
        g.wk[0][:] = g.uc * g.uc; g.wk[1][:] = g.vc * g.vc; g.wk[2][:] = g.wc * g.wc
        g.wk[3][:] = g.uc * g.vc; g.wk[4][:] = g.uc * g.wc; g.wk[5][:] = g.vc * g.wc
        if not filter_2d:
            for m in range(6):
                extrapolate(r.n, r.is_bound, s.dzci, g.wk[m], iface=0, cbc=cbcvel)
        for m in range(6):
            filt(r.n, g.wk[m], lij[m])
        if not filter_2d:
            for a in (g.uc, g.vc, g.wc):
                extrapolate(r.n, r.is_bound, s.dzci, a, iface=0, cbc=cbcvel)
        filt(r.n, g.uc, g.uf); filt(r.n, g.vc, g.vf); filt(r.n, g.wc, g.wf)
        n1, n2, n3 = r.n
        I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
        uf, vf, wf = g.uf[I], g.vf[I], g.wf[I]
        M = [g.mij[m][I] for m in range(6)]
        L = [lij[m][I] for m in range(6)]
        L[0] = L[0] - uf * uf; L[1] = L[1] - vf * vf; L[2] = L[2] - wf * wf
        L[3] = L[3] - uf * vf; L[4] = L[4] - uf * wf; L[5] = L[5] - vf * wf
        g.wk[0][I] = M[0] * L[0] + M[1] * L[1] + M[2] * L[2] + (M[3] * L[3] + M[4] * L[4] + M[5] * L[5]) * 2.
        g.wk[1][I] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + (M[3] * M[3] + M[4] * M[4] + M[5] * M[5]) * 2.
    for m in (0, 1):                                        # sgs.f90:359-370
        F = [g.wk[m] for g in G]
        if ave == "channel":
            ave1d_channel(world, st, 2, F)
        elif ave == "dit":
            ave0d_dit(world, st, F)
        elif ave == "duct":
            ave2d_duct(world, st, 0, F)
    for r, g in zip(R, G):                                  # sgs.f90:372-380
        n1, n2, n3 = r.n
        I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
        with np.errstate(divide="ignore", invalid="ignore"):
            vt = VISCT[r.id][I] * g.wk[0][I] / g.wk[1][I]
        # Fortran max(x,0.) with x=NaN: gfortran's MAX returns the non-NaN operand
        VISCT[r.id][I] = np.fmax(vt, 0.)
