"""On-the-fly statistics of the channel -- TEST INFRASTRUCTURE (checker of cales_out1d_chan).
Restates the first block of out1d_single_point_chan, src/output.f90:509-691 (idir = 3): 27 plane-averaged single-point
profiles, accumulated per rank over its interior and summed over ranks (MPI_ALLREDUCE, output.f90:683).
PARITY UNPINNED: the reference ships no reference output for this routine."""
import numpy as np

NVARS = 27


def out1d_single_point_chan_local(ng, lo, hi, l, dl, dzc, dzf, u, v, w, p, visct):
    """One rank's contribution: buf(27, ng3) with the rows of its own k range filled (output.f90:541-680).
    dzc, dzf: the rank-local slices (0:n3+1) of dzc_g, dzf_g; fields haloed (0:n+1)."""
    n1, n2, n3 = (hi[q] - lo[q] + 1 for q in range(3))
    I = (slice(1, n1 + 1), slice(1, n2 + 1), slice(1, n3 + 1))

    def sh(a, di=0, dj=0, dk=0):
        return a[1 + di:n1 + 1 + di, 1 + dj:n2 + 1 + dj, 1 + dk:n3 + 1 + dk]
    dzck = dzc[1:n3 + 1][None, None, :]; dzfk = dzf[1:n3 + 1][None, None, :]; dzfkp = dzf[2:n3 + 2][None, None, :]
    uc, vc, wc, pc = u[I], v[I], w[I], p[I]
    q = []
    q += [uc, vc, wc, uc ** 2, vc ** 2, wc ** 2]
    q.append(0.25 * (sh(u, dk=1) + uc) * (wc + sh(w, di=1)))
    q += [uc ** 3, vc ** 3, wc ** 3, uc ** 4, vc ** 4, wc ** 4, pc, pc ** 2]
    tx = (sh(w, dj=1) - wc) / dl[1] - (sh(v, dk=1) - vc) / dzck
    ty = (sh(u, dk=1) - uc) / dzck - (sh(w, di=1) - wc) / dl[0]
    tz = (sh(v, di=1) - vc) / dl[0] - (sh(u, dj=1) - uc) / dl[1]
    q += [tx, ty, tz, tx ** 2, ty ** 2, tz ** 2]
    s_ccc, s_pcc, s_cpc, s_ccp, s_pcp = visct[I], sh(visct, di=1), sh(visct, dj=1), sh(visct, dk=1), sh(visct, di=1, dk=1)
    dudx_ip = (sh(u, di=1) - uc) / dl[0]; dudx_im = (uc - sh(u, di=-1)) / dl[0]
    dvdy_jp = (sh(v, dj=1) - vc) / dl[1]; dvdy_jm = (vc - sh(v, dj=-1)) / dl[1]
    dwdz_kp = (sh(w, dk=1) - wc) / dzfkp; dwdz_km = (wc - sh(w, dk=-1)) / dzfk
    dudz = (sh(u, dk=1) - uc) / dzck; dwdx = (sh(w, di=1) - wc) / dl[0]
    q.append(-0.5 * (s_pcc * (dudx_ip + dudx_ip) + s_ccc * (dudx_im + dudx_im)))
    q.append(-0.5 * (s_cpc * (dvdy_jp + dvdy_jp) + s_ccc * (dvdy_jm + dvdy_jm)))
    q.append(-0.5 * (s_ccp * (dwdz_kp + dwdz_kp) + s_ccc * (dwdz_km + dwdz_km)))
    q.append(-0.25 * (s_ccc + s_pcc + s_ccp + s_pcp) * (dudz + dwdx))
    q += [s_ccc, dudz + 0. * uc]
    gar = dl[0] * dl[1] / (l[0] * l[1])
    buf = np.zeros((NVARS, ng[2]), order="F")
    for m, a in enumerate(q):
        buf[m, lo[2] - 1:hi[2]] = a.sum(axis=(0, 1)) * gar
    return buf


def out1d_single_point_chan(world, st, deck, U, V, W, P, VISCT):
    """All ranks + the MPI_ALLREDUCE (rank order)."""
    tot = np.zeros((NVARS, deck.ng[2]), order="F")
    for r, s in zip(world.ranks, st):
        tot += out1d_single_point_chan_local(deck.ng, r.lo, r.hi, deck.l, deck.dl, s.dzc, s.dzf, U[r.id], V[r.id], W[r.id], P[r.id], VISCT[r.id])
    return tot
