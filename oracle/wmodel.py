"""Log-law / laminar wall model.  Follows src/wmodel.f90: updt_wallmodelbc 19-63,
cmpt_wallmodelbc 65-273, vel_relative 275-286, wallmodel 288-335.  The Newton-Raphson loop is
vectorised over the wall plane with a per-point active mask so that every point performs exactly
the iterations the scalar `do while(conv>0.5e-4)` would."""
import numpy as np

from .param import b_log, eps, kap_log

WM_LAM, WM_LOG = -1, 1
last_iters = None        # iteration-count histogram input of the most recent wallmodel() call


def vel_relative(v1, v2, coef, bcv_mag):                    # wmodel.f90:275-286
    vr = (1.0 - coef) * v1 + coef * v2
    vr = vr - bcv_mag
    return vr


def wallmodel(mtype, uh, vh, h, l1d, visc):                 # wmodel.f90:288-335
    global last_iters
    uh = np.asarray(uh, dtype=np.float64)
    vh = np.asarray(vh, dtype=np.float64)
    if mtype == WM_LOG:
        upar = np.sqrt(uh * uh + vh * vh)
        utau = np.maximum(np.sqrt(upar / h * visc), visc / h * np.exp(-kap_log * b_log))
        conv = np.ones_like(upar)
        iters = np.zeros(upar.shape, dtype=np.int32)
        active = conv > 0.5e-4
        while active.any():
            utau_old = utau
            f = upar / utau - 1.0 / kap_log * np.log(h * utau / visc) - b_log
            fp = -1.0 / utau * (upar / utau + 1.0 / kap_log)
            utau_new = np.abs(utau - f / fp)
            conv_new = np.abs(utau_new / utau_old - 1.0)
            utau = np.where(active, utau_new, utau)
            conv = np.where(active, conv_new, conv)
            iters += active
            active = conv > 0.5e-4
        last_iters = iters
        tauw_tot = utau * utau
        return tauw_tot * uh / (upar + eps), tauw_tot * vh / (upar + eps)
    elif mtype == WM_LAM:
        upar = np.sqrt(uh * uh + vh * vh)
        dl_ = 0.5 * l1d
        umax = upar / (h / dl_ * (2.0 - h / dl_))
        tauw_tot = 2.0 / dl_ * umax * visc
        return tauw_tot * uh / (upar + eps), tauw_tot * vh / (upar + eps)
    raise ValueError(mtype)


def cmpt_wallmodelbc(n, ibound, idir, mtype, l, dl, zc, zf, dzc, dzf, visc, h, index, vel1, vel2,
                     bcvel1, bcvel2, bcvel1_mag, bcvel2_mag):
    """wmodel.f90:65-273.  (vel1,vel2) = (v,w) for idir=0, (u,w) for idir=1, (u,v) for idir=2."""
    visci = 1.0 / visc
    n1, n2, n3 = n
    if idir == 0:
        if ibound == 0:
            i2 = index; i1 = index - 1
            coef = (h - (i1 - 0.5) * dl[0]) / dl[0]; sgn = 1.0
        else:
            i2 = index; i1 = index + 1
            coef = (h - (n1 - i1 + 0.5) * dl[0]) / dl[0]; sgn = -1.0
        v, w, bcv, bcw, bcv_mag, bcw_mag = vel1, vel2, bcvel1, bcvel2, bcvel1_mag, bcvel2_mag
        J = slice(0, n2 + 1); Jp = slice(1, n2 + 2)
        K = slice(1, n3 + 1); Km = slice(0, n3)
        v1 = v[i1, J, K]; v2 = v[i2, J, K]
        w1 = 0.25 * (w[i1, J, K] + w[i1, Jp, K] + w[i1, J, Km] + w[i1, Jp, Km])
        w2 = 0.25 * (w[i2, J, K] + w[i2, Jp, K] + w[i2, J, Km] + w[i2, Jp, Km])
        v_mag = bcv_mag[J, K, ibound]
        w_mag = 0.25 * (bcw_mag[J, K, ibound] + bcw_mag[Jp, K, ibound] + bcw_mag[J, Km, ibound] + bcw_mag[Jp, Km, ibound])
        vh = vel_relative(v1, v2, coef, v_mag); wh = vel_relative(w1, w2, coef, w_mag)
        t1, _ = wallmodel(mtype, vh, wh, h, l[0], visc)
        bcv[J, K, ibound] = sgn * visci * t1
        K = slice(0, n3 + 1); Kp = slice(1, n3 + 2)
        J = slice(1, n2 + 1); Jm = slice(0, n2)
        wei = ((zf[K] - zc[K]) / dzc[K])[None, :]
        v1 = 0.5 * ((1.0 - wei) * (v[i1, Jm, K] + v[i1, J, K]) + wei * (v[i1, Jm, Kp] + v[i1, J, Kp]))
        v2 = 0.5 * ((1.0 - wei) * (v[i2, Jm, K] + v[i2, J, K]) + wei * (v[i2, Jm, Kp] + v[i2, J, Kp]))
        w1 = w[i1, J, K]; w2 = w[i2, J, K]
        v_mag = 0.5 * ((1.0 - wei) * (bcv_mag[Jm, K, ibound] + bcv_mag[J, K, ibound]) +
                       wei * (bcv_mag[Jm, Kp, ibound] + bcv_mag[J, Kp, ibound]))
        w_mag = bcw_mag[J, K, ibound]
        vh = vel_relative(v1, v2, coef, v_mag); wh = vel_relative(w1, w2, coef, w_mag)
        _, t2 = wallmodel(mtype, vh, wh, h, l[0], visc)
        bcw[J, K, ibound] = sgn * visci * t2
    elif idir == 1:
        if ibound == 0:
            j2 = index; j1 = index - 1
            coef = (h - (j1 - 0.5) * dl[1]) / dl[1]; sgn = 1.0
        else:
            j2 = index; j1 = index + 1
            coef = (h - (n2 - j1 + 0.5) * dl[1]) / dl[1]; sgn = -1.0
        u, w, bcu, bcw, bcu_mag, bcw_mag = vel1, vel2, bcvel1, bcvel2, bcvel1_mag, bcvel2_mag
        I = slice(0, n1 + 1); Ip = slice(1, n1 + 2)
        K = slice(1, n3 + 1); Km = slice(0, n3)
        u1 = u[I, j1, K]; u2 = u[I, j2, K]
        w1 = 0.25 * (w[I, j1, K] + w[Ip, j1, K] + w[I, j1, Km] + w[Ip, j1, Km])
        w2 = 0.25 * (w[I, j2, K] + w[Ip, j2, K] + w[I, j2, Km] + w[Ip, j2, Km])
        u_mag = bcu_mag[I, K, ibound]
        w_mag = 0.25 * (bcw_mag[I, K, ibound] + bcw_mag[Ip, K, ibound] + bcw_mag[I, Km, ibound] + bcw_mag[Ip, Km, ibound])
        uh = vel_relative(u1, u2, coef, u_mag); wh = vel_relative(w1, w2, coef, w_mag)
        t1, _ = wallmodel(mtype, uh, wh, h, l[1], visc)
        bcu[I, K, ibound] = sgn * visci * t1
        K = slice(0, n3 + 1); Kp = slice(1, n3 + 2)
        I = slice(1, n1 + 1); Im = slice(0, n1)
        wei = ((zf[K] - zc[K]) / dzc[K])[None, :]
        u1 = 0.5 * ((1.0 - wei) * (u[Im, j1, K] + u[I, j1, K]) + wei * (u[Im, j1, Kp] + u[I, j1, Kp]))
        u2 = 0.5 * ((1.0 - wei) * (u[Im, j2, K] + u[I, j2, K]) + wei * (u[Im, j2, Kp] + u[I, j2, Kp]))
        w1 = w[I, j1, K]; w2 = w[I, j2, K]
        u_mag = 0.5 * ((1.0 - wei) * (bcu_mag[Im, K, ibound] + bcu_mag[I, K, ibound]) +
                       wei * (bcu_mag[Im, Kp, ibound] + bcu_mag[I, Kp, ibound]))
        w_mag = bcw_mag[I, K, ibound]
        uh = vel_relative(u1, u2, coef, u_mag); wh = vel_relative(w1, w2, coef, w_mag)
        _, t2 = wallmodel(mtype, uh, wh, h, l[1], visc)
        bcw[I, K, ibound] = sgn * visci * t2
    else:
        if ibound == 0:
            k2 = index; k1 = index - 1
            coef = (h - zc[k1]) / dzc[k1]; sgn = 1.0
        else:
            k2 = index; k1 = index + 1
            coef = (h - (l[2] - zc[k1])) / (dzc[k2]); sgn = -1.0
        u, v, bcu, bcv, bcu_mag, bcv_mag = vel1, vel2, bcvel1, bcvel2, bcvel1_mag, bcvel2_mag
        I = slice(0, n1 + 1); Ip = slice(1, n1 + 2)
        J = slice(1, n2 + 1); Jm = slice(0, n2)
        u1 = u[I, J, k1]; u2 = u[I, J, k2]
        v1 = 0.25 * (v[I, J, k1] + v[Ip, J, k1] + v[I, Jm, k1] + v[Ip, Jm, k1])
        v2 = 0.25 * (v[I, J, k2] + v[Ip, J, k2] + v[I, Jm, k2] + v[Ip, Jm, k2])
        u_mag = bcu_mag[I, J, ibound]
        v_mag = 0.25 * (bcv_mag[I, J, ibound] + bcv_mag[Ip, J, ibound] + bcv_mag[I, Jm, ibound] + bcv_mag[Ip, Jm, ibound])
        uh = vel_relative(u1, u2, coef, u_mag); vh = vel_relative(v1, v2, coef, v_mag)
        t1, _ = wallmodel(mtype, uh, vh, h, l[2], visc)
        bcu[I, J, ibound] = sgn * visci * t1
        J = slice(0, n2 + 1); Jp = slice(1, n2 + 2)
        I = slice(1, n1 + 1); Im = slice(0, n1)
        u1 = 0.25 * (u[Im, J, k1] + u[I, J, k1] + u[Im, Jp, k1] + u[I, Jp, k1])
        u2 = 0.25 * (u[Im, J, k2] + u[I, J, k2] + u[Im, Jp, k2] + u[I, Jp, k2])
        v1 = v[I, J, k1]; v2 = v[I, J, k2]
        u_mag = 0.25 * (bcu_mag[Im, J, ibound] + bcu_mag[I, J, ibound] + bcu_mag[Im, Jp, ibound] + bcu_mag[I, Jp, ibound])
        v_mag = bcv_mag[I, J, ibound]
        uh = vel_relative(u1, u2, coef, u_mag); vh = vel_relative(v1, v2, coef, v_mag)
        _, t2 = wallmodel(mtype, uh, vh, h, l[2], visc)
        bcv[I, J, ibound] = sgn * visci * t2


def updt_wallmodelbc(n, is_bound, lwm, l, dl, zc, zf, dzc, dzf, visc, h, index_wm, u, v, w,
                     bcu, bcv, bcw, bcu_mag, bcv_mag, bcw_mag):
    """wmodel.f90:19-63."""
    vel = (u, v, w); bc = (bcu, bcv, bcw); mag = (bcu_mag, bcv_mag, bcw_mag)
    for idir in range(3):
        c1, c2 = [c for c in range(3) if c != idir]
        ax = "xyz"[idir]
        for ib in range(2):
            if is_bound[ib, idir] and lwm[ib, idir] != 0:
                cmpt_wallmodelbc(n, ib, idir, lwm[ib, idir], l, dl, zc, zf, dzc, dzf, visc, h, index_wm[ib, idir],
                                 vel[c1], vel[c2], bc[c1][ax], bc[c2][ax], mag[c1][ax], mag[c2][ax])
