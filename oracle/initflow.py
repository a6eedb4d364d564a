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


def temporal_bl(n, zc, d, nu, norm):
    """initflow.f90:375-390."""
    theta = 54. * nu / norm
    p = np.zeros(n)
    for k in range(1, n + 1):
        p[k - 1] = (0.5 + (0.5) * np.tanh((d / (2. * theta)) * (1. - zc[k] / d))) * norm
    return p


def log_profile(n, zc, reb):
    """initflow.f90:392-407."""
    retau = 0.09 * reb ** (0.88)
    p = np.zeros(n)
    for k in range(1, n + 1):
        z = zc[k] * 2. * retau
        if z >= retau:
            z = 2. * retau - z
        p[k - 1] = 2.5 * np.log(z) + 5.5
        if z <= 11.6:
            p[k - 1] = z
    return p


def add_noise(ng, lo, n, iseed, norm, p):
    """initflow.f90:285-315 with numpy's PCG64(iseed) standing in for the Fortran compiler's random_number stream (which
    cannot be reproduced outside that compiler): one draw per global cell, i fastest, decomposition independent."""
    rng = np.random.Generator(np.random.PCG64(iseed))
    for k in range(1, ng[2] + 1):
        kk = k - (lo[2] - 1)
        rn = rng.random(ng[0] * ng[1])
        if 1 <= kk <= n[2]:
            for jj in range(1, n[1] + 1):
                jg = jj + lo[1] - 1
                for ii in range(1, n[0] + 1):
                    ig = ii + lo[0] - 1
                    p[ii, jj, kk] = p[ii, jj, kk] + 2. * (rn[(ig - 1) + ng[0] * (jg - 1)] - .5) * norm


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
    is_noise = False
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
    elif inivel == "tbl":                                  # initflow.f90:60-62
        u1d = temporal_bl(n[2], zc, 1.0, visc, uref)
        is_noise = True
    elif inivel == "log":                                  # initflow.f90:76-80
        reb = ubulk * l[2] / visc
        u1d = log_profile(n[2], zc / l[2], reb)
        is_noise = True
        is_mean = True
    elif inivel in ("hcl", "hcp"):                         # initflow.f90:81-102
        zc2 = np.zeros(2 * n[2] + 2)
        zc2[1:n[2] + 1] = zc[1:n[2] + 1]
        zc2[n[2] + 1:2 * n[2] + 1] = 2 * l[2] - zc[n[2]:0:-1]
        zc2[0] = -zc[0]
        zc2[2 * n[2] + 1] = 2 * l[2] + zc[0]
        if inivel == "hcl":
            reb = ubulk * (2 * l[2]) / visc
            u1d = log_profile(2 * n[2], zc2 / (2 * l[2]), reb)
            is_noise = True
        else:
            u1d = poiseuille(2 * n[2], zc2 / (2 * l[2]), ubulk)
        is_mean = True
    elif inivel == "ant":                                  # initflow.f90:134-156
        c0 = 4. * np.sqrt(2.) / 3. / np.sqrt(3.)
        for k in range(1, n[2] + 1):
            zcc = zc[k] / l[2] * 2. * pi + 0.5 * pi
            zff = zf[k] / l[2] * 2. * pi + 0.5 * pi
            for jj in range(1, n[1] + 1):
                yc = (jj + lo[1] - 1 - .5) * dl[1] / l[1] * 2. * pi + 0.5 * pi
                yf = (jj + lo[1] - 1 - .0) * dl[1] / l[1] * 2. * pi + 0.5 * pi
                ii = np.arange(1, n[0] + 1)
                xc = (ii + lo[0] - 1 - .5) * dl[0] / l[0] * 2. * pi + 0.5 * pi
                xf = (ii + lo[0] - 1 - .0) * dl[0] / l[0] * 2. * pi + 0.5 * pi
                u[1:n[0] + 1, jj, k] = c0 * (np.sin(xf - 5. * pi / 6.) * np.cos(yc - 1. * pi / 6.) * np.sin(zcc) -
                                             np.sin(xf - 1. * pi / 6.) * np.sin(yc) * np.cos(zcc - 5. * pi / 6.)) * uref
                v[1:n[0] + 1, jj, k] = c0 * (np.sin(xc) * np.sin(yf - 5. * pi / 6.) * np.sin(zcc - 1. * pi / 6.) -
                                             np.cos(xc - 5. * pi / 6.) * np.sin(yf - 1. * pi / 6.) * np.sin(zcc)) * uref
                w[1:n[0] + 1, jj, k] = c0 * (np.cos(xc - 1. * pi / 6.) * np.sin(yc) * np.sin(zff - 5. * pi / 6.) -
                                             np.sin(xc) * np.cos(yc - 5. * pi / 6.) * np.sin(zff - 1. * pi / 6.)) * uref
        p[I] = -(u[I] ** 2 + v[I] ** 2 + w[I] ** 2) / 2.
    elif inivel in ("pdc", "hdc"):                         # initflow.f90:157-180
        lref = l[2] / 2.
        if inivel != "pdc":
            lref = 2. * lref
        if deck.is_wallturb:
            uref = (deck.bforce[0] * lref) ** (0.5)
            retau = uref * lref / visc
            reb = (retau / .09) ** (1. / .88)
            ubulk = reb * visc / (2 * lref)
        else:
            ubulk = (deck.bforce[0] * lref ** 2 / (3. * visc))
        if inivel == "pdc":
            u1d = poiseuille(n[2], zc / l[2], ubulk)
        else:
            zc2 = np.zeros(2 * n[2] + 2)
            zc2[1:n[2] + 1] = zc[1:n[2] + 1]
            zc2[n[2] + 1:2 * n[2] + 1] = 2 * l[2] - zc[n[2]:0:-1]
            zc2[0] = -zc[0]
            zc2[2 * n[2] + 1] = 2 * l[2] + zc[0]
            u1d = poiseuille(2 * n[2], zc2 / (2 * l[2]), ubulk)
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
        raise ValueError("oracle.initflow: invalid name for initial velocity field '%s'" % inivel)
    if inivel not in ("tgv", "tgw", "ant", "duc"):          # initflow.f90:211-222
        u[I] = u1d[None, None, :n[2]]
    if is_noise:                                            # initflow.f90:223-227
        add_noise(deck.ng, lo, n, 123, .05, u)
        add_noise(deck.ng, lo, n, 456, .05, v)
        add_noise(deck.ng, lo, n, 789, .05, w)
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
