"""Boundary conditions.  Follows src/bound.f90: set_bc 202-399, bounduvw 18-154, boundp 156-200,
cmpt_rhs_b/bc_rhs 447-560, updt_rhs_b 562-617, initbc 726-867.  A `bound` (typedef.f90:10-14)
is a dict {'x','y','z'} of F-ordered planes x(0:n2+1,0:n3+1,0:1), y(0:n1+1,0:n3+1,0:1),
z(0:n1+1,0:n2+1,0:1)."""
import numpy as np

from . import wmodel

_AX = "xyz"


def new_bound(n, value=0.0):
    return {"x": np.full((n[1] + 2, n[2] + 2, 2), value, order="F"),
            "y": np.full((n[0] + 2, n[2] + 2, 2), value, order="F"),
            "z": np.full((n[0] + 2, n[1] + 2, 2), value, order="F")}


def copy_bound(b):
    return {k: v.copy(order="F") for k, v in b.items()}


def _pl(idir, idx):
    s = [slice(None)] * 3
    s[idir] = idx
    return tuple(s)


def set_bc(ctype, ibound, idir, nh, centered, bc, dr, p):
    """bound.f90:202-399 with nh=1 (dh=0).  Whole-plane array syntax: ghost rows of the other two
    directions are included, which defines edge/corner ghosts by call order."""
    assert nh == 1
    n = p.shape[idir] - 2 * nh
    sgn = 1.0
    if ctype == "D" and centered:
        sgn = -1.0
    b = bc[:, :, ibound]
    P = lambda idx: p[_pl(idir, idx)]
    if ctype == "P":
        p[_pl(idir, 0)] = P(n)
        p[_pl(idir, n + 1)] = P(1)
    elif ctype == "D":
        if centered:
            if ibound == 0:
                p[_pl(idir, 0)] = 2.0 * b + sgn * P(1)
            else:
                p[_pl(idir, n + 1)] = 2.0 * b + sgn * P(n)
        else:
            if ibound == 0:
                p[_pl(idir, 0)] = b
            else:
                p[_pl(idir, n + 1)] = P(n - 1)      # unused
                p[_pl(idir, n)] = b
    elif ctype == "N":
        if centered:
            if ibound == 0:
                p[_pl(idir, 0)] = -dr * b + sgn * P(1)
            else:
                p[_pl(idir, n + 1)] = dr * b + sgn * P(n)
        else:
            if ibound == 0:
                p[_pl(idir, 0)] = -dr * b + P(1)
            else:
                p[_pl(idir, n + 1)] = P(n)          # unused
                p[_pl(idir, n)] = dr * b + P(n - 1)


def bounduvw_local(cbc, n, bcu, bcv, bcw, bcu_mag, bcv_mag, bcw_mag, is_bound, lwm, l, dl, zc, zf, dzc, dzf,
                   visc, h, index_wm, is_updt_wm, is_correc, u, v, w):
    """bound.f90:53-148: everything after the halo exchange (rank-local)."""
    nh = 1
    vel = (u, v, w)
    bcs = (bcu, bcv, bcw)
    drn = (dl[0], dl[0]), (dl[1], dl[1]), (dzf[0], dzf[n[2]])       # normal component spacing per face
    drt = (dl[0], dl[0]), (dl[1], dl[1]), (dzc[0], dzc[n[2]])       # tangential components
    for idir in range(3):
        impose_norm_bc = (not is_correc) or (cbc[0, idir, idir] + cbc[1, idir, idir] == "PP")
        others = [c for c in range(3) if c != idir]
        for ib in range(2):
            if is_bound[ib, idir]:
                if impose_norm_bc:
                    set_bc(cbc[ib, idir, idir], ib, idir, nh, False, bcs[idir][_AX[idir]], drn[idir][ib], vel[idir])
                if lwm[ib, idir] == 0:
                    for c in others:
                        set_bc(cbc[ib, idir, c], ib, idir, nh, True, bcs[c][_AX[idir]], drt[idir][ib], vel[c])
    if is_updt_wm:
        wmodel.updt_wallmodelbc(n, is_bound, lwm, l, dl, zc, zf, dzc, dzf, visc, h, index_wm, u, v, w,
                                bcu, bcv, bcw, bcu_mag, bcv_mag, bcw_mag)
    for idir in range(3):
        others = [c for c in range(3) if c != idir]
        for ib in range(2):
            if is_bound[ib, idir] and lwm[ib, idir] != 0:
                for c in others:
                    set_bc(cbc[ib, idir, c], ib, idir, nh, True, bcs[c][_AX[idir]], drt[idir][ib], vel[c])


def boundp_local(cbc, n, bcp, is_bound, dl, dzc, p):
    """bound.f90:181-199."""
    nh = 1
    dr = (dl[0], dl[0]), (dl[1], dl[1]), (dzc[0], dzc[n[2]])
    for idir in range(3):
        for ib in range(2):
            if is_bound[ib, idir]:
                set_bc(cbc[ib, idir], ib, idir, nh, True, bcp[_AX[idir]], dr[idir][ib], p)


# ---- world-level wrappers (halo exchange + local part) ---------------------------------------
def bounduvw(world, cbc, st, is_updt_wm, is_correc, U, V, W, bcu=None, bcv=None, bcw=None):
    """`st` is the list of per-rank state objects (see main.RankState)."""
    for idir in range(3):                                   # bound.f90:42-46
        world.updthalo(U, idir)
        world.updthalo(V, idir)
        world.updthalo(W, idir)
    for r, s in zip(world.ranks, st):
        bounduvw_local(cbc, r.n, (bcu or [x.bcu for x in st])[r.id], (bcv or [x.bcv for x in st])[r.id],
                       (bcw or [x.bcw for x in st])[r.id], s.bcu_mag, s.bcv_mag, s.bcw_mag, r.is_bound,
                       s.lwm, s.l, s.dl, s.zc, s.zf, s.dzc, s.dzf, s.visc, s.hwm, s.index_wm,
                       is_updt_wm, is_correc, U[r.id], V[r.id], W[r.id])


def boundp(world, cbc, st, bcname, P):
    for idir in range(3):                                   # bound.f90:175-177
        world.updthalo(P, idir)
    for r, s in zip(world.ranks, st):
        boundp_local(cbc, r.n, getattr(s, bcname), r.is_bound, s.dl, s.dzc, P[r.id])


# ---- rhs boundary contributions -----------------------------------------------------------------
def bc_rhs(cbc, bc, dlc, dlf, c_or_f):
    """bound.f90:497-560.  cbc: (2,) chars; bc: plane (0:n1+1,0:n2+1,0:1); returns rhs (n1,n2,0:1)."""
    n1 = bc.shape[0] - 2
    n2 = bc.shape[1] - 2
    rhs = np.zeros((n1, n2, 2), order="F")
    for ibound in range(2):
        b = bc[1:n1 + 1, 1:n2 + 1, ibound]
        c = cbc[ibound]
        if c_or_f == "c":
            if c == "P":
                rhs[:, :, ibound] = 0.0
            elif c == "D":
                rhs[:, :, ibound] = -2.0 * b / dlc[ibound] / dlf[ibound]
            elif c == "N":
                sgn = 1.0 if ibound == 0 else -1.0
                rhs[:, :, ibound] = sgn * b / dlf[ibound]
        else:
            if c == "P":
                rhs[:, :, ibound] = 0.0
            elif c == "D":
                rhs[:, :, ibound] = -b / dlc[ibound] / dlf[ibound]
            elif c == "N":
                sgn = 1.0 if ibound == 0 else -1.0
                rhs[:, :, ibound] = sgn * b / dlc[ibound]
    return rhs


def cmpt_rhs_b(ng, dl, dzc_g, dzf_g, cbc, bc, c_or_f):
    """bound.f90:447-495.  dzc_g/dzf_g are the GLOBAL z metrics (0:ng3+1).  Returns rhsbx,rhsby,rhsbz."""
    dxc01 = [dl[0], dl[0]]; dxf01 = [dl[0], dl[0]]
    dyc01 = [dl[1], dl[1]]; dyf01 = [dl[1], dl[1]]
    dzc01_c = [dzc_g[0], dzc_g[ng[2]]]
    dzf01_c = [dzf_g[1], dzf_g[ng[2]]]
    dzc01_f = [dzc_g[1], dzc_g[ng[2] - 1]]
    dzf01_f = [dzf_g[1], dzf_g[ng[2]]]
    rhsbx = bc_rhs(cbc[:, 0], bc["x"], dxc01, dxf01, c_or_f[0])
    rhsby = bc_rhs(cbc[:, 1], bc["y"], dyc01, dyf01, c_or_f[1])
    if c_or_f[2] == "c":
        rhsbz = bc_rhs(cbc[:, 2], bc["z"], dzc01_c, dzf01_c, c_or_f[2])
    else:
        rhsbz = bc_rhs(cbc[:, 2], bc["z"], dzc01_f, dzf01_f, c_or_f[2])
    return rhsbx, rhsby, rhsbz


def updt_rhs_b(c_or_f, cbc, n, is_bound, rhsbx, rhsby, rhsbz, p):
    """bound.f90:562-617 (rhsb? may be None = absent optional)."""
    q = [0, 0, 0]
    for idir in range(3):
        if c_or_f[idir] == "f" and cbc[1, idir] == "D":
            q[idir] = 1
    n1, n2, n3 = n
    if rhsbx is not None:
        if is_bound[0, 0]:
            p[1, 1:n2 + 1, 1:n3 + 1] = p[1, 1:n2 + 1, 1:n3 + 1] + rhsbx[:, :, 0]
        if is_bound[1, 0]:
            nn = n1 - q[0]
            p[nn, 1:n2 + 1, 1:n3 + 1] = p[nn, 1:n2 + 1, 1:n3 + 1] + rhsbx[:, :, 1]
    if rhsby is not None:
        if is_bound[0, 1]:
            p[1:n1 + 1, 1, 1:n3 + 1] = p[1:n1 + 1, 1, 1:n3 + 1] + rhsby[:, :, 0]
        if is_bound[1, 1]:
            nn = n2 - q[1]
            p[1:n1 + 1, nn, 1:n3 + 1] = p[1:n1 + 1, nn, 1:n3 + 1] + rhsby[:, :, 1]
    if rhsbz is not None:
        if is_bound[0, 2]:
            p[1:n1 + 1, 1:n2 + 1, 1] = p[1:n1 + 1, 1:n2 + 1, 1] + rhsbz[:, :, 0]
        if is_bound[1, 2]:
            nn = n3 - q[2]
            p[1:n1 + 1, 1:n2 + 1, nn] = p[1:n1 + 1, 1:n2 + 1, nn] + rhsbz[:, :, 1]


def initbc(deck, n, is_bound, zc, dzc):
    """bound.f90:726-867.  Returns (cbcvel, bcu,bcv,bcw,bcp,bcs,bcu_mag,bcv_mag,bcw_mag,bcuf,bcvf,bcwf,index_wm);
    cbcvel is the deck's with wall-model faces forced to D (normal) / N (tangential)."""
    cbcvel = deck.cbcvel.copy()
    lwm, l, dl, h = deck.lwm, deck.l, deck.dl, deck.hwm
    for idir in range(3):
        for i in range(2):
            if lwm[i, idir] != 0:
                for ivel in range(3):
                    cbcvel[i, idir, ivel] = "D" if ivel == idir else "N"
    bcu, bcv, bcw, bcp, bcs = (new_bound(n) for _ in range(5))
    for idir in range(3):
        for ib in range(2):
            bcu[_AX[idir]][:, :, ib] = deck.bcvel[ib, idir, 0]
            bcv[_AX[idir]][:, :, ib] = deck.bcvel[ib, idir, 1]
            bcw[_AX[idir]][:, :, ib] = deck.bcvel[ib, idir, 2]
            bcp[_AX[idir]][:, :, ib] = deck.bcpre[ib, idir]
            bcs[_AX[idir]][:, :, ib] = deck.bcsgs[ib, idir]
    bcu_mag, bcv_mag, bcw_mag = copy_bound(bcu), copy_bound(bcv), copy_bound(bcw)
    bcuf, bcvf, bcwf = copy_bound(bcu), copy_bound(bcv), copy_bound(bcw)
    index_wm = np.zeros((2, 3), dtype=np.int32)
    for idir in range(2):                                   # x and y: uniform spacing
        if is_bound[0, idir] and lwm[0, idir] != 0:
            i = 1
            while (i - 0.5) * dl[idir] < h:
                i = i + 1
            index_wm[0, idir] = i
        if is_bound[1, idir] and lwm[1, idir] != 0:
            i = n[idir]
            while (n[idir] - i + 0.5) * dl[idir] < h:
                i = i - 1
            index_wm[1, idir] = i
    if is_bound[0, 2] and lwm[0, 2] != 0:
        k = 1
        while zc[k] < h:
            k = k + 1
        index_wm[0, 2] = k
    if is_bound[1, 2] and lwm[1, 2] != 0:
        k = n[2]
        while l[2] - zc[k] < h:
            k = k - 1
        index_wm[1, 2] = k
    return cbcvel, bcu, bcv, bcw, bcp, bcs, bcu_mag, bcv_mag, bcw_mag, bcuf, bcvf, bcwf, index_wm
