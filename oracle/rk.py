"""Low-storage RK3 update.  Follows src/rk.f90:17-121 (`rk`), 197-222 (`cmpt_bulk_forcing`) and
src/utils.f90:16-47 (`bulk_mean`).  The `save`d arrays of the Fortran routine live in RkState."""
import numpy as np

from .mom import mom_xyz_ad, SLAB_MIN_CELLS, run_slabs


class RkState:
    """rk.f90:36-72: dudtrko etc. start at zero and are swapped with dudtrk after every call."""

    def __init__(self, n):
        self.o = [np.zeros(tuple(n), order="F") for _ in range(3)]


def bulk_mean_local(n, grid_vol_ratio, p):
    """utils.f90:33-44: sequential accumulation, i fastest (rank-local part)."""
    n1, n2, n3 = n
    t = (p[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] * grid_vol_ratio[None, None, 1:n3 + 1]).ravel(order="F")
    return float(np.cumsum(t)[-1]) if t.size else 0.0


def rk_update(rkpar, n, dli, dzci, dzfi, visc, dt, p, bforce, visct, u, v, w, state,
              impdiff=False, impdiff_1d=False):
    """rk.f90:45-100: momentum RHS + update + swap.  Returns the implicit RHS parts (or None)."""
    n1, n2, n3 = n
    factor1 = rkpar[0] * dt
    factor2 = rkpar[1] * dt
    factor12 = factor1 + factor2
    (dudtrk, dvdtrk, dwdtrk), imp = mom_xyz_ad(n, dli[0], dli[1], dzci, dzfi, visc, u, v, w, visct,
                                              impdiff, impdiff_1d)
    def update(k0, nb):
        """levels k0+1 .. k0+nb (per-cell work: any k range gives the same bits)"""
        K = slice(k0 + 1, k0 + nb + 1); Kp = slice(k0 + 2, k0 + nb + 2); Kh = slice(k0, k0 + nb)
        I = (slice(1, n1 + 1), slice(1, n2 + 1), K)
        pc = p[I]
        dzci_k = dzci[K][None, None, :]
        unew = u[I] + factor1 * dudtrk[:, :, Kh] + factor2 * state.o[0][:, :, Kh] + \
            factor12 * (bforce[0] - dli[0] * (p[2:n1 + 2, 1:n2 + 1, K] - pc))
        vnew = v[I] + factor1 * dvdtrk[:, :, Kh] + factor2 * state.o[1][:, :, Kh] + \
            factor12 * (bforce[1] - dli[1] * (p[1:n1 + 1, 2:n2 + 2, K] - pc))
        wnew = w[I] + factor1 * dwdtrk[:, :, Kh] + factor2 * state.o[2][:, :, Kh] + \
            factor12 * (bforce[2] - dzci_k * (p[1:n1 + 1, 1:n2 + 1, Kp] - pc))
        if impdiff:
            unew = unew + factor12 * imp[0][:, :, Kh]
            vnew = vnew + factor12 * imp[1][:, :, Kh]
            wnew = wnew + factor12 * imp[2][:, :, Kh]
        u[I] = unew; v[I] = vnew; w[I] = wnew
    if n1 * n2 * n3 < SLAB_MIN_CELLS:
        update(0, n3)
    else:
        run_slabs(n3, n1 * n2, update)
    state.o = [dudtrk, dvdtrk, dwdtrk]                      # swap: the new RHS becomes the old one
    return imp, factor12


def rk_impdiff_rhs(n, factor12, imp, u, v, w):
    """rk.f90:106-120: Helmholtz right-hand side."""
    n1, n2, n3 = n
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))
    u[I] = u[I] - .5 * factor12 * imp[0]
    v[I] = v[I] - .5 * factor12 * imp[1]
    w[I] = w[I] - .5 * factor12 * imp[2]


def rk(world, st, rkpar, dt, P, VISCT, U, V, W, deck):
    """World-level `rk`: update on every rank, then cmpt_bulk_forcing (rk.f90:197-222) with the
    MPI_ALLREDUCE emulated in rank order.  Returns f(3)."""
    imps = []
    for r, s in zip(world.ranks, st):
        imp, f12 = rk_update(rkpar, r.n, s.dli, s.dzci, s.dzfi, s.visc, dt, P[r.id], deck.bforce, VISCT[r.id],
                             U[r.id], V[r.id], W[r.id], s.rk, deck.impdiff, deck.impdiff_1d)
        imps.append((imp, f12))
    f = [0.0, 0.0, 0.0]
    for c, (F_, gname) in enumerate(((U, "grid_vol_ratio_f"), (V, "grid_vol_ratio_f"), (W, "grid_vol_ratio_c"))):
        if deck.is_forced[c]:
            mean = world.allreduce_sum([bulk_mean_local(r.n, getattr(s, gname), F_[r.id])
                                        for r, s in zip(world.ranks, st)])
            f[c] = deck.velf[c] - mean
    if deck.impdiff:
        for r, (imp, f12) in zip(world.ranks, imps):
            rk_impdiff_rhs(r.n, f12, imp, U[r.id], V[r.id], W[r.id])
    return f
