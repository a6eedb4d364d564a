"""Initial fields.  Follows src/initflow.f90:17-283 (deterministic branches) and the helpers
340-434.  `add_noise` (initflow.f90:285-315) draws from the Fortran compiler's random_number
stream and is therefore not reproducible outside that compiler: branches that need it
('tbl','log','hcl') are rejected here."""
import numpy as np

from .param import pi


def poiseuille(n, zc, norm):                               # initflow.f90:349-363
    p = np.zeros(n)
    for k in range(1, n + 1):
        z = zc[k]
        p[k - 1] = 6. * z * (1. - z) * norm
    return p


def couette(n, zc, norm):                                  # initflow.f90:335-347
    p = np.zeros(n)
    for k in range(1, n + 1):
        z = zc[k]
        p[k - 1] = .5 * (1. - 2. * z) * norm
    return p


def _fz(zc):                                               # initflow.f90:410-414
    return (1. - zc ** 2) ** 2


def _dfz(zc):                                              # initflow.f90:416-420
    return -4. * zc * (1. - zc ** 2)


def _gxy(xc, yc):                                          # initflow.f90:422-426
    return yc * np.exp(-4. * (4. * xc ** 2 + yc ** 2))


def _dgxy(xc, yc):                                         # initflow.f90:428-432
    return np.exp(-4. * (4. * xc ** 2 + yc ** 2)) * (1. - 8. * yc ** 2)


def initflow(deck, lo, n, zc, zf, dzc, dzf, allreduce_sum=lambda x: x):
    """Returns u,v,w,p of shape (n+2) in F order for the rank whose global lower corner is `lo`
    (1-based, as in the reference).  `allreduce_sum` emulates MPI_ALLREDUCE in set_mean."""
    inivel = deck.inivel.strip()
    l, dl, visc = deck.l, deck.dl, deck.visc
    bcvel, is_forced, velf = deck.bcvel, deck.is_forced, deck.velf
    shp = (n[0] + 2, n[1] + 2, n[2] + 2)
    u = np.zeros(shp, order="F"); v = np.zeros(shp, order="F")
    w = np.zeros(shp, order="F"); p = np.zeros(shp, order="F")
    is_mean = False
    uref = 1.0
    ubulk = uref
    if is_forced[0]:
        ubulk = velf[0]
    u1d = np.zeros(n[2])
    i = np.arange(1, n[0] + 1)[:, None, None]
    j = np.arange(1, n[1] + 1)[None, :, None]
    kk = np.arange(1, n[2] + 1)
    I = (slice(1, n[0] + 1), slice(1, n[1] + 1), slice(1, n[2] + 1))
    if inivel == "cou":
        uref = bcvel[0, 2, 0] - bcvel[1, 2, 0]
        u1d = couette(n[2], zc / l[2], uref)
    elif inivel == "poi":
        u1d = poiseuille(n[2], zc / l[2], ubulk)
        is_mean = True
    elif inivel == "iop":
        ubulk = .5 * abs(bcvel[0, 2, 0] + bcvel[1, 2, 0])
        u1d = poiseuille(n[2], zc / l[2], ubulk)
        u1d = u1d - ubulk
        is_mean = True
    elif inivel == "zer":
        u1d[:] = 0.0
    elif inivel == "uni":
        u1d[:] = uref
    elif inivel == "tgv":                                  # initflow.f90:103-118
        zcc = (zc[kk] / l[2] * 2. * pi)[None, None, :]
        yc = (j + lo[1] - 1 - .5) * dl[1] / l[1] * 2. * pi
        yf = (j + lo[1] - 1 - .0) * dl[1] / l[1] * 2. * pi
        xc = (i + lo[0] - 1 - .5) * dl[0] / l[0] * 2. * pi
        xf = (i + lo[0] - 1 - .0) * dl[0] / l[0] * 2. * pi
        u[I] = np.sin(xf) * np.cos(yc) * np.cos(zcc) * uref
        v[I] = -np.cos(xc) * np.sin(yf) * np.cos(zcc) * uref
    elif inivel == "tgw":                                  # initflow.f90:119-133
        yc = (j + lo[1] - 1 - .5) * dl[1]
        yf = (j + lo[1] - 1 - .0) * dl[1]
        xc = (i + lo[0] - 1 - .5) * dl[0]
        xf = (i + lo[0] - 1 - .0) * dl[0]
        one = np.ones((1, 1, n[2]))
        u[I] = np.cos(xf) * np.sin(yc) * uref * one
        v[I] = -np.sin(xc) * np.cos(yf) * uref * one
        p[I] = -(np.cos(2. * xc) + np.cos(2. * yc)) / 4. * uref ** 2 * one
    elif inivel == "pdc":                                  # initflow.f90:155-180
        lref = l[2] / 2.
        if deck.is_wallturb:
            uref = (deck.bforce[0] * lref) ** (0.5)
            retau = uref * lref / visc
            reb = (retau / .09) ** (1. / .88)
            ubulk = reb * visc / (2 * lref)
        else:
            ubulk = (deck.bforce[0] * lref ** 2 / (3. * visc))
        u1d = poiseuille(n[2], zc / l[2], ubulk)
        is_mean = True
    elif inivel == "duc":                                  # initflow.f90:181-201
        ly = .5 * l[1]
        lz = .5 * l[2]
        for k in range(1, n[2] + 1):
            for jj in range(1, n[1] + 1):
                sum_term = 0.0
                xi = -1. + (jj + lo[1] - 1.5) * dl[1] / ly
                eta = -1. + zc[k] / lz
                for m in range(0, 101):
                    cosh_term = np.cosh((2 * m + 1) * pi * ly / (2 * lz) * xi) / np.cosh((2 * m + 1) * pi * ly / (2 * lz))
                    cos_term = np.cos((2 * m + 1) * pi / 2 * eta)
                    term = (-1.) ** m / (2 * m + 1) ** 3 * cosh_term * cos_term
                    sum_term = sum_term + term
                # u(:,j,k): the whole x extent including halo cells
                u[:, jj, k] = .5 * lz ** 2 * (1. - eta ** 2 - 4. * (2. / pi) ** 3 * sum_term)
        is_mean = True
    else:
        raise ValueError("oracle.initflow: inivel '%s' is not restated (needs add_noise or unknown)" % inivel)
    if inivel not in ("tgv", "tgw", "ant", "duc"):          # initflow.f90:211-222
        u[I] = u1d[None, None, :]
    if is_mean and inivel != "iop":                        # initflow.f90:228-232, set_mean 317-333
        gvr = dzf / l[2] * (dl[0] / l[0]) * (dl[1] / l[1])
        ui = u[I]
        meanold = 0.0
        # sequential accumulation in i-fastest order, as the Fortran loop nest
        meanold = float(np.cumsum((ui * gvr[None, None, 1:n[2] + 1]).ravel(order="F"))[-1])
        meanold = allreduce_sum(meanold)
        if meanold != 0.0:
            u[I] = ui / meanold * ubulk
    if deck.is_wallturb:                                   # initflow.f90:233-260 vortex pair
        zcc = (2. * zc[kk] / l[2] - 1.)[None, None, :]
        zff = (2. * (zc[kk] / l[2] + .5 * dzf[kk] / l[2]) - 1.)[None, None, :]
        yc = ((lo[1] - 1 + j - 0.5) * dl[1] - .5 * l[1]) * 2. / l[2]
        yf = ((lo[1] - 1 + j - 0.0) * dl[1] - .5 * l[1]) * 2. / l[2]
        xc = ((lo[0] - 1 + i - 0.5) * dl[0] - .5 * l[0]) * 2. / l[2]
        v[I] = -1. * _gxy(yf, xc) * _dfz(zcc) * ubulk * 1.5
        w[I] = 1. * _fz(zff) * _dgxy(yc, xc) * ubulk * 1.5
        p[I] = 0.0
    return u, v, w, p
