"""Host-side input synthesis: the caller side of the boundary.

Mirrors what the reference's Fortran host does before the time loop: `initgrid`
(src/initgrid.f90:15-196), `initflow` (src/initflow.f90:17-283, deterministic branches) and
`initbc` (src/bound.f90:726-867).  These produce the constant operands the device path consumes;
they are plain numpy on the host, as they are plain Fortran on the host in the reference."""
import numpy as np

pi = float(np.arccos(-1.0))
f32 = np.float32


def initgrid(gtype, n, gr, lz):
    """z grid (initgrid.f90:15-81).  The Fortran evaluates z0=(k-0.)/(1.*n) and dzc=0.1*32./nzg in default
    (single) precision; reproduced here.  Returns dzc,dzf,zc,zf with extent 0:n+1."""
    k = np.arange(1, n + 1)
    z0 = (k.astype(f32) / f32(n)).astype(np.float64)
    a = gr
    if gtype == 1 or gtype not in (2, 3, 4, 5, 6):
        z = 0.5 * (1. + np.tanh((z0 - 0.5) * a) / np.tanh(a / 2.)) if a != 0. else z0
    elif gtype == 2:
        z = 1.0 * (1. + np.tanh((z0 - 1.0) * a) / np.tanh(a / 1.)) if a != 0. else z0
    elif gtype == 3:
        z = 1. - 1.0 * (1. + np.tanh((1. - z0 - 1.0) * a) / np.tanh(a / 1.)) if a != 0. else z0
    elif gtype == 4:
        if a != 0.:
            z = np.where(z0 <= 0.5, 0.5 * (1. - 1. + np.tanh(2. * a * (z0 - 0.)) / np.tanh(a)),
                         0.5 * (1. + 1. + np.tanh(2. * a * (z0 - 1.)) / np.tanh(a)))
        else:
            z = z0
    elif gtype == 6:
        dzc_ = float(f32(f32(0.1) * f32(32.0)) / f32(n))
        z = z0 - (dzc_ * n / 2. - 1.) / (2. * pi) * np.sin(2. * pi * z0)
    else:  # gtype 5, Pirozzoli & Orlandi natural stretching (scalar arithmetic, as the Fortran routine)
        kb, alpha, c_eta, dyp = 32., pi / 1.5, 0.8, 0.05
        nn = n / 2.
        retau = 1. / (1. + (nn / kb) ** 2) * (dyp * nn + (3. / 4. * alpha * c_eta * nn) ** (4. / 3.) * (nn / kb) ** 2)
        z = np.zeros(n)
        for kg in range(1, n + 1):
            kk = 1. * min(kg, n - kg)
            zz = 1. / (1. + (kk / kb) ** 2) * (dyp * kk + (3. / 4. * alpha * c_eta * kk) ** (4. / 3.) * (kk / kb) ** 2) / (2. * retau)
            z[kg - 1] = 1. - zz if kg > n - kg else zz
    zf = np.zeros(n + 2); dzf = np.zeros(n + 2); dzc = np.zeros(n + 2); zc = np.zeros(n + 2)
    zf[1:n + 1] = z * lz
    dzf[1:n + 1] = zf[1:n + 1] - zf[0:n]
    dzf[0] = dzf[1]; dzf[n + 1] = dzf[n]
    dzc[0:n + 1] = .5 * (dzf[0:n + 1] + dzf[1:n + 2])
    dzc[n + 1] = dzc[n]
    zc[0] = -dzc[0] / 2.
    zf[0] = 0.
    for q in range(1, n + 2):        # running sums, as the Fortran loop
        zc[q] = zc[q - 1] + dzc[q - 1]
        zf[q] = zf[q - 1] + dzf[q]
    return dzc, dzf, zc, zf


def add_noise(ng, lo, n, iseed, norm, p):
    """`add_noise`, initflow.f90:285-315: one random number per GLOBAL cell in i-fastest order, added as 2(rn-.5)norm to the
    cells this rank owns, so the field does not depend on the decomposition.  The reference draws from the Fortran
    compiler's `random_number` stream seeded with `iseed`, which cannot be reproduced outside that compiler (SURVEY.md
    section 8d); here the stream is numpy's PCG64 seeded with `iseed` -- same distribution, same amplitude, same
    decomposition independence, different numbers."""
    rng = np.random.Generator(np.random.PCG64(iseed))
    plane = int(ng[0]) * int(ng[1])
    i0, j0 = lo[0] - 1, lo[1] - 1
    for k in range(1, int(ng[2]) + 1):
        kk = k - (lo[2] - 1)
        if kk < 1 or kk > n[2]:
            rng.bit_generator.advance(plane)                 # one 64-bit draw per double
            continue
        rn = rng.random(plane).reshape((int(ng[0]), int(ng[1])), order="F")
        p[1:n[0] + 1, 1:n[1] + 1, kk] += 2. * (rn[i0:i0 + n[0], j0:j0 + n[1]] - .5) * norm


def initflow(deck, lo, n, zc, zf, dzc, dzf, mean_allreduce=None):
    """Initial u,v,w,p (haloed, Fortran order) of the rank with 1-based lower corner `lo`
    (initflow.f90:17-283).  `mean_allreduce(partial)` supplies the global sum used by set_mean."""
    inivel = deck.inivel.strip()
    l, dl, visc = deck.l, deck.dl, deck.visc
    shp = (n[0] + 2, n[1] + 2, n[2] + 2)
    u, v, w, p = (np.zeros(shp, order="F") for _ in range(4))
    I = (slice(1, n[0] + 1), slice(1, n[1] + 1), slice(1, n[2] + 1))
    i = np.arange(1, n[0] + 1)[:, None, None]
    j = np.arange(1, n[1] + 1)[None, :, None]
    kk = np.arange(1, n[2] + 1)
    uref = 1.0
    ubulk = deck.velf[0] if deck.is_forced[0] else uref
    is_mean = False
    is_noise = False
    u1d = None
    zcn = zc[kk] / l[2]

    def mirrored():
        """zc2/(2 l3) of the half-channel cases (initflow.f90:84-88, 96-100); only levels 1..n3 are used afterwards."""
        return zc[kk] / (2. * l[2])
    if inivel == "poi":
        u1d = 6. * zcn * (1. - zcn) * ubulk; is_mean = True
    elif inivel == "cou":
        uref = deck.bcvel[0, 2, 0] - deck.bcvel[1, 2, 0]
        u1d = .5 * (1. - 2. * zcn) * uref
    elif inivel == "iop":
        ubulk = .5 * abs(deck.bcvel[0, 2, 0] + deck.bcvel[1, 2, 0])
        u1d = 6. * zcn * (1. - zcn) * ubulk - ubulk; is_mean = True
    elif inivel == "zer":
        u1d = np.zeros(n[2])
    elif inivel == "uni":
        u1d = np.full(n[2], uref)
    elif inivel == "tbl":                                     # initflow.f90:60-62, temporal_bl 375-390
        theta = 54. * visc / uref
        u1d = (0.5 + 0.5 * np.tanh((1. / (2. * theta)) * (1. - zc[kk] / 1.))) * uref
        is_noise = True
    elif inivel in ("log", "hcl"):                            # initflow.f90:76-92, log_profile 392-407
        half = inivel == "hcl"
        reb = ubulk * ((2. * l[2]) if half else l[2]) / visc
        retau = 0.09 * reb ** 0.88
        z = (mirrored() if half else zcn) * 2. * retau
        z = np.where(z >= retau, 2. * retau - z, z)
        u1d = np.where(z <= 11.6, z, 2.5 * np.log(z) + 5.5)
        is_noise = True; is_mean = True
    elif inivel == "hcp":                                     # initflow.f90:93-102
        zz = mirrored()
        u1d = 6. * zz * (1. - zz) * ubulk; is_mean = True
    elif inivel == "ant":                                     # initflow.f90:134-156 (Antuono, JFM 890, A23)
        a = 4. * np.sqrt(2.) / 3. / np.sqrt(3.)
        zcc = (zc[kk] / l[2] * 2. * pi + 0.5 * pi)[None, None, :]; zff = (zf[kk] / l[2] * 2. * pi + 0.5 * pi)[None, None, :]
        yc = (j + lo[1] - 1 - .5) * dl[1] / l[1] * 2. * pi + 0.5 * pi; yf = (j + lo[1] - 1 - .0) * dl[1] / l[1] * 2. * pi + 0.5 * pi
        xc = (i + lo[0] - 1 - .5) * dl[0] / l[0] * 2. * pi + 0.5 * pi; xf = (i + lo[0] - 1 - .0) * dl[0] / l[0] * 2. * pi + 0.5 * pi
        u[I] = a * (np.sin(xf - 5. * pi / 6.) * np.cos(yc - 1. * pi / 6.) * np.sin(zcc) -
                    np.sin(xf - 1. * pi / 6.) * np.sin(yc) * np.cos(zcc - 5. * pi / 6.)) * uref
        v[I] = a * (np.sin(xc) * np.sin(yf - 5. * pi / 6.) * np.sin(zcc - 1. * pi / 6.) -
                    np.cos(xc - 5. * pi / 6.) * np.sin(yf - 1. * pi / 6.) * np.sin(zcc)) * uref
        w[I] = a * (np.cos(xc - 1. * pi / 6.) * np.sin(yc) * np.sin(zff - 5. * pi / 6.) -
                    np.sin(xc) * np.cos(yc - 5. * pi / 6.) * np.sin(zff - 1. * pi / 6.)) * uref
        p[I] = -(u[I] ** 2 + v[I] ** 2 + w[I] ** 2) / 2.
    elif inivel in ("pdc", "hdc"):
        lref = l[2] / 2.
        if inivel != "pdc":
            lref = 2. * lref
        if deck.is_wallturb:
            uref = (deck.bforce[0] * lref) ** 0.5
            retau = uref * lref / visc
            reb = (retau / .09) ** (1. / .88)
            ubulk = reb * visc / (2 * lref)
        else:
            ubulk = deck.bforce[0] * lref ** 2 / (3. * visc)
        zz = zcn if inivel == "pdc" else mirrored()
        u1d = 6. * zz * (1. - zz) * ubulk; is_mean = True
    elif inivel == "tgv":
        zcc = (zc[kk] / l[2] * 2. * pi)[None, None, :]
        yc = (j + lo[1] - 1 - .5) * dl[1] / l[1] * 2. * pi; yf = (j + lo[1] - 1 - .0) * dl[1] / l[1] * 2. * pi
        xc = (i + lo[0] - 1 - .5) * dl[0] / l[0] * 2. * pi; xf = (i + lo[0] - 1 - .0) * dl[0] / l[0] * 2. * pi
        u[I] = np.sin(xf) * np.cos(yc) * np.cos(zcc) * uref
        v[I] = -np.cos(xc) * np.sin(yf) * np.cos(zcc) * uref
    elif inivel == "tgw":
        yc = (j + lo[1] - 1 - .5) * dl[1]; yf = (j + lo[1] - 1 - .0) * dl[1]
        xc = (i + lo[0] - 1 - .5) * dl[0]; xf = (i + lo[0] - 1 - .0) * dl[0]
        one = np.ones((1, 1, n[2]))
        u[I] = np.cos(xf) * np.sin(yc) * uref * one
        v[I] = -np.sin(xc) * np.cos(yf) * uref * one
        p[I] = -(np.cos(2. * xc) + np.cos(2. * yc)) / 4. * uref ** 2 * one
    elif inivel == "duc":
        ly, lz = .5 * l[1], .5 * l[2]
        xi = (-1. + (np.arange(1, n[1] + 1) + lo[1] - 1.5) * dl[1] / ly)[:, None]
        eta = (-1. + zc[kk] / lz)[None, :]
        sum_term = np.zeros((n[1], n[2]))
        for m in range(0, 101):
            cosh_term = np.cosh((2 * m + 1) * pi * ly / (2 * lz) * xi) / np.cosh((2 * m + 1) * pi * ly / (2 * lz))
            cos_term = np.cos((2 * m + 1) * pi / 2 * eta)
            sum_term = sum_term + (-1.) ** m / (2 * m + 1) ** 3 * cosh_term * cos_term
        u[:, 1:n[1] + 1, 1:n[2] + 1] = (.5 * lz ** 2 * (1. - eta ** 2 - 4. * (2. / pi) ** 3 * sum_term))[None, :, :]
        is_mean = True
    else:
        raise ValueError("invalid name for initial velocity field: '%s'" % inivel)      # initflow.f90:202-209
    if u1d is not None:
        u[I] = u1d[None, None, :]
    if is_noise:                                             # initflow.f90:223-227
        for fld, seed in ((u, 123), (v, 456), (w, 789)):
            add_noise(deck.ng, lo, n, seed, .05, fld)
    if is_mean and inivel != "iop":
        gvr = dzf / l[2] * (dl[0] / l[0]) * (dl[1] / l[1])
        part = float(np.cumsum((u[I] * gvr[None, None, 1:n[2] + 1]).ravel(order="F"))[-1])
        meanold = mean_allreduce(part) if mean_allreduce else part
        if meanold != 0.:
            u[I] = u[I] / meanold * ubulk
    if deck.is_wallturb:                                     # vortex pair, initflow.f90:233-260
        zcc = (2. * zc[kk] / l[2] - 1.)[None, None, :]
        zff = (2. * (zc[kk] / l[2] + .5 * dzf[kk] / l[2]) - 1.)[None, None, :]
        yc = ((lo[1] - 1 + j - 0.5) * dl[1] - .5 * l[1]) * 2. / l[2]
        yf = ((lo[1] - 1 + j - 0.0) * dl[1] - .5 * l[1]) * 2. / l[2]
        xc = ((lo[0] - 1 + i - 0.5) * dl[0] - .5 * l[0]) * 2. / l[2]
        gxy = xc * np.exp(-4. * (4. * yf ** 2 + xc ** 2))                    # gxy(yf,xc)
        dfz = -4. * zcc * (1. - zcc ** 2)
        fz = (1. - zff ** 2) ** 2
        dgxy = np.exp(-4. * (4. * yc ** 2 + xc ** 2)) * (1. - 8. * xc ** 2)  # dgxy(yc,xc)
        v[I] = -1. * gxy * dfz * ubulk * 1.5
        w[I] = 1. * fz * dgxy * ubulk * 1.5
        p[I] = 0.
    return u, v, w, p


def initbc(deck, n, is_bound, zc, dzc):
    """bound.f90:726-867: wall-model faces become D (normal) / N (tangential); constant BC planes;
    interpolation index for the wall-model height.  Planes are dicts {'x','y','z'} of F-ordered arrays."""
    cbcvel = deck.cbcvel.copy()
    lwm, l, dl, h = deck.lwm, deck.l, deck.dl, deck.hwm
    for idir in range(3):
        for ib in range(2):
            if lwm[ib, idir] != 0:
                for ivel in range(3):
                    cbcvel[ib, idir, ivel] = "D" if ivel == idir else "N"
    shapes = {"x": (n[1] + 2, n[2] + 2, 2), "y": (n[0] + 2, n[2] + 2, 2), "z": (n[0] + 2, n[1] + 2, 2)}

    def planes(vals):          # vals[ib, idir]
        out = {}
        for idir, ax in enumerate("xyz"):
            a = np.zeros(shapes[ax], order="F")
            a[:, :, 0] = vals[0, idir]; a[:, :, 1] = vals[1, idir]
            out[ax] = a
        return out
    bcu, bcv, bcw = (planes(deck.bcvel[:, :, c]) for c in range(3))
    bcp, bcs = planes(deck.bcpre), planes(deck.bcsgs)
    index_wm = np.zeros((2, 3), dtype=np.int32)
    for idir in range(2):
        if is_bound[0, idir] and lwm[0, idir] != 0:
            q = 1
            while (q - 0.5) * dl[idir] < h:
                q += 1
            index_wm[0, idir] = q
        if is_bound[1, idir] and lwm[1, idir] != 0:
            q = n[idir]
            while (n[idir] - q + 0.5) * dl[idir] < h:
                q -= 1
            index_wm[1, idir] = q
    if is_bound[0, 2] and lwm[0, 2] != 0:
        q = 1
        while zc[q] < h:
            q += 1
        index_wm[0, 2] = q
    if is_bound[1, 2] and lwm[1, 2] != 0:
        q = n[2]
        while l[2] - zc[q] < h:
            q -= 1
        index_wm[1, 2] = q
    return cbcvel, bcu, bcv, bcw, bcp, bcs, index_wm
