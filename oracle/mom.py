"""Momentum right-hand side.  Follows src/mom.f90:17-309 (`mom_xyz_ad`) expression by expression
(same association order, so a non-contracting compiler gives bit-identical results) and
src/mom.f90:311-335 (`bulk_forcing`)."""
import numpy as np


def _views(a, n):
    n1, n2, n3 = n

    def s(di, dj, dk):
        return a[1 + di:n1 + 1 + di, 1 + dj:n2 + 1 + dj, 1 + dk:n3 + 1 + dk]
    return s


# Large grids (the BASELINE-size parity cases) are evaluated slab by slab in z on a thread pool: every operation below is
# elementwise on shifted views, so cutting the k range changes neither the operations nor their order -- the results are
# the same bits (tests/test_oracle_identities.py::test_mom_slabs_identical) -- but the temporaries stay in cache and
# numpy's inner loops (which release the GIL) run on all host cores.
SLAB_CELLS = 1 << 18          # cells per slab
SLAB_MIN_CELLS = 1 << 21      # grids below this size are evaluated in one piece


def slab_threads():
    import os
    try:
        return max(1, min(32, len(os.sched_getaffinity(0))))
    except AttributeError:
        return max(1, min(32, os.cpu_count() or 1))


def run_slabs(n3, nplane, fn):
    """fn(k0, nb) for consecutive k ranges [k0+1, k0+nb] (1-based interior levels) covering 1..n3."""
    nb = max(1, SLAB_CELLS // max(nplane, 1))
    jobs = [(k0, min(nb, n3 - k0)) for k0 in range(0, n3, nb)]
    nt = min(slab_threads(), len(jobs))
    if nt <= 1:
        for j in jobs:
            fn(*j)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda j: fn(*j), jobs))


def mom_xyz_ad(n, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, impdiff=False, impdiff_1d=False):
    """Returns (dudt,dvdt,dwdt) and, with impdiff, also (dudtd,dvdtd,dwdtd); arrays (n1,n2,n3)."""
    n1, n2, n3 = n
    if n1 * n2 * n3 < SLAB_MIN_CELLS:
        return _mom_block(n, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, impdiff, impdiff_1d)
    out = [np.empty((n1, n2, n3), order="F") for _ in range(3)]
    outd = [np.empty((n1, n2, n3), order="F") for _ in range(3)] if impdiff else None

    def job(k0, nb):
        z = slice(k0, k0 + nb + 2)
        r, rd = _mom_block((n1, n2, nb), dxi, dyi, dzci[z], dzfi[z], visc, u[:, :, z], v[:, :, z], w[:, :, z], visct[:, :, z],
                           impdiff, impdiff_1d)
        for c in range(3):
            out[c][:, :, k0:k0 + nb] = r[c]
            if impdiff:
                outd[c][:, :, k0:k0 + nb] = rd[c]
    run_slabs(n3, n1 * n2, job)
    return tuple(out), (tuple(outd) if impdiff else None)


def _mom_block(n, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, impdiff=False, impdiff_1d=False):
    n1, n2, n3 = n
    U, V, W, S = _views(u, n), _views(v, n), _views(w, n), _views(visct, n)
    k = np.arange(1, n3 + 1)
    dzci_k = dzci[k][None, None, :]; dzci_km = dzci[k - 1][None, None, :]
    dzfi_k = dzfi[k][None, None, :]; dzfi_kp = dzfi[k + 1][None, None, :]
    names = dict(ccm=(0, 0, -1), pcm=(1, 0, -1), cpm=(0, 1, -1), cmc=(0, -1, 0), pmc=(1, -1, 0), mcc=(-1, 0, 0),
                 ccc=(0, 0, 0), pcc=(1, 0, 0), mpc=(-1, 1, 0), cpc=(0, 1, 0), cmp=(0, -1, 1), mcp=(-1, 0, 1),
                 ccp=(0, 0, 1))
    u_ = {k_: U(*o) for k_, o in names.items()}
    v_ = {k_: V(*o) for k_, o in names.items()}
    w_ = {k_: W(*o) for k_, o in names.items()}
    s_ = {k_: S(*o) for k_, o in names.items()}
    s_["ppc"] = S(1, 1, 0); s_["pcp"] = S(1, 0, 1); s_["cpp"] = S(0, 1, 1)
    u_ccm, u_pcm, u_cpm, u_cmc, u_pmc, u_mcc, u_ccc, u_pcc, u_mpc, u_cpc, u_cmp, u_mcp, u_ccp = (u_[x] for x in names)
    v_ccm, v_pcm, v_cpm, v_cmc, v_pmc, v_mcc, v_ccc, v_pcc, v_mpc, v_cpc, v_cmp, v_mcp, v_ccp = (v_[x] for x in names)
    w_ccm, w_pcm, w_cpm, w_cmc, w_pmc, w_mcc, w_ccc, w_pcc, w_mpc, w_cpc, w_cmp, w_mcp, w_ccp = (w_[x] for x in names)
    s_ccm, s_pcm, s_cpm, s_cmc, s_pmc, s_mcc, s_ccc, s_pcc, s_mpc, s_cpc, s_cmp, s_mcp, s_ccp = (s_[x] for x in names)
    s_ppc, s_pcp, s_cpp = s_["ppc"], s_["pcp"], s_["cpp"]
    #
    # x momentum (mom.f90:142-186)
    #
    visc_ip = s_pcc
    visc_im = s_ccc
    visc_jp = 0.25 * (s_ccc + s_pcc + s_cpc + s_ppc)
    visc_jm = 0.25 * (s_ccc + s_pcc + s_cmc + s_pmc)
    visc_kp = 0.25 * (s_ccc + s_pcc + s_ccp + s_pcp)
    visc_km = 0.25 * (s_ccc + s_pcc + s_ccm + s_pcm)
    dudx_ip = (u_pcc - u_ccc) * dxi
    dudx_im = (u_ccc - u_mcc) * dxi
    dudy_jp = (u_cpc - u_ccc) * dyi
    dudy_jm = (u_ccc - u_cmc) * dyi
    dudz_kp = (u_ccp - u_ccc) * dzci_k
    dudz_km = (u_ccc - u_ccm) * dzci_km
    dvdx_jp = (v_pcc - v_ccc) * dxi
    dvdx_jm = (v_pmc - v_cmc) * dxi
    dwdx_kp = (w_pcc - w_ccc) * dxi
    dwdx_km = (w_pcm - w_ccm) * dxi
    uu_ip = 0.25 * (u_pcc + u_ccc) * (u_ccc + u_pcc)
    uu_im = 0.25 * (u_mcc + u_ccc) * (u_ccc + u_mcc)
    vu_jp = 0.25 * (v_pcc + v_ccc) * (u_ccc + u_cpc)
    vu_jm = 0.25 * (v_pmc + v_cmc) * (u_ccc + u_cmc)
    wu_kp = 0.25 * (w_pcc + w_ccc) * (u_ccc + u_ccp)
    wu_km = 0.25 * (w_pcm + w_ccm) * (u_ccc + u_ccm)
    dudtd_xy_s = visc * (dudx_ip - dudx_im) * dxi + \
                 visc * (dudy_jp - dudy_jm) * dyi
    dudtd_z_s = visc * (dudz_kp - dudz_km) * dzfi_k
    dudt_s = -(uu_ip - uu_im) * dxi - \
              (vu_jp - vu_jm) * dyi - \
              (wu_kp - wu_km) * dzfi_k \
             + (visc_ip * (dudx_ip + dudx_ip) - visc_im * (dudx_im + dudx_im)) * dxi + \
               (visc_jp * (dudy_jp + dvdx_jp) - visc_jm * (dudy_jm + dvdx_jm)) * dyi + \
               (visc_kp * (dudz_kp + dwdx_kp) - visc_km * (dudz_km + dwdx_km)) * dzfi_k
    #
    # y momentum (mom.f90:187-231)
    #
    visc_ip = 0.25 * (s_ccc + s_cpc + s_pcc + s_ppc)
    visc_im = 0.25 * (s_ccc + s_cpc + s_mcc + s_mpc)
    visc_jp = s_cpc
    visc_jm = s_ccc
    visc_kp = 0.25 * (s_ccc + s_cpc + s_ccp + s_cpp)
    visc_km = 0.25 * (s_ccc + s_cpc + s_ccm + s_cpm)
    dvdx_ip = (v_pcc - v_ccc) * dxi
    dvdx_im = (v_ccc - v_mcc) * dxi
    dvdy_jp = (v_cpc - v_ccc) * dyi
    dvdy_jm = (v_ccc - v_cmc) * dyi
    dvdz_kp = (v_ccp - v_ccc) * dzci_k
    dvdz_km = (v_ccc - v_ccm) * dzci_km
    dudy_ip = (u_cpc - u_ccc) * dyi
    dudy_im = (u_mpc - u_mcc) * dyi
    dwdy_kp = (w_cpc - w_ccc) * dyi
    dwdy_km = (w_cpm - w_ccm) * dyi
    uv_ip = 0.25 * (u_ccc + u_cpc) * (v_ccc + v_pcc)
    uv_im = 0.25 * (u_mcc + u_mpc) * (v_ccc + v_mcc)
    vv_jp = 0.25 * (v_ccc + v_cpc) * (v_ccc + v_cpc)
    vv_jm = 0.25 * (v_ccc + v_cmc) * (v_ccc + v_cmc)
    wv_kp = 0.25 * (w_ccc + w_cpc) * (v_ccc + v_ccp)
    wv_km = 0.25 * (w_ccm + w_cpm) * (v_ccc + v_ccm)
    dvdtd_xy_s = visc * (dvdx_ip - dvdx_im) * dxi + \
                 visc * (dvdy_jp - dvdy_jm) * dyi
    dvdtd_z_s = visc * (dvdz_kp - dvdz_km) * dzfi_k
    dvdt_s = -(uv_ip - uv_im) * dxi - \
              (vv_jp - vv_jm) * dyi - \
              (wv_kp - wv_km) * dzfi_k \
             + (visc_ip * (dvdx_ip + dudy_ip) - visc_im * (dvdx_im + dudy_im)) * dxi + \
               (visc_jp * (dvdy_jp + dvdy_jp) - visc_jm * (dvdy_jm + dvdy_jm)) * dyi + \
               (visc_kp * (dvdz_kp + dwdy_kp) - visc_km * (dvdz_km + dwdy_km)) * dzfi_k
    #
    # z momentum (mom.f90:232-276)
    #
    visc_ip = 0.25 * (s_ccc + s_ccp + s_pcc + s_pcp)
    visc_im = 0.25 * (s_ccc + s_ccp + s_mcc + s_mcp)
    visc_jp = 0.25 * (s_ccc + s_ccp + s_cpc + s_cpp)
    visc_jm = 0.25 * (s_ccc + s_ccp + s_cmc + s_cmp)
    visc_kp = s_ccp
    visc_km = s_ccc
    dwdx_ip = (w_pcc - w_ccc) * dxi
    dwdx_im = (w_ccc - w_mcc) * dxi
    dwdy_jp = (w_cpc - w_ccc) * dyi
    dwdy_jm = (w_ccc - w_cmc) * dyi
    dwdz_kp = (w_ccp - w_ccc) * dzfi_kp
    dwdz_km = (w_ccc - w_ccm) * dzfi_k
    dudz_ip = (u_ccp - u_ccc) * dzci_k
    dudz_im = (u_mcp - u_mcc) * dzci_k
    dvdz_jp = (v_ccp - v_ccc) * dzci_k
    dvdz_jm = (v_cmp - v_cmc) * dzci_k
    uw_ip = 0.25 * (u_ccc + u_ccp) * (w_ccc + w_pcc)
    uw_im = 0.25 * (u_mcc + u_mcp) * (w_ccc + w_mcc)
    vw_jp = 0.25 * (v_ccc + v_ccp) * (w_ccc + w_cpc)
    vw_jm = 0.25 * (v_cmc + v_cmp) * (w_ccc + w_cmc)
    ww_kp = 0.25 * (w_ccc + w_ccp) * (w_ccc + w_ccp)
    ww_km = 0.25 * (w_ccc + w_ccm) * (w_ccc + w_ccm)
    dwdtd_xy_s = visc * (dwdx_ip - dwdx_im) * dxi + \
                 visc * (dwdy_jp - dwdy_jm) * dyi
    dwdtd_z_s = visc * (dwdz_kp - dwdz_km) * dzci_k
    dwdt_s = -(uw_ip - uw_im) * dxi - \
              (vw_jp - vw_jm) * dyi - \
              (ww_kp - ww_km) * dzci_k \
             + (visc_ip * (dwdx_ip + dudz_ip) - visc_im * (dwdx_im + dudz_im)) * dxi + \
               (visc_jp * (dwdy_jp + dvdz_jp) - visc_jm * (dwdy_jm + dvdz_jm)) * dyi + \
               (visc_kp * (dwdz_kp + dwdz_kp) - visc_km * (dwdz_km + dwdz_km)) * dzci_k
    F = np.asfortranarray
    if impdiff:                                             # mom.f90:277-295
        if impdiff_1d:
            dudt_s = dudt_s + dudtd_xy_s
            dvdt_s = dvdt_s + dvdtd_xy_s
            dwdt_s = dwdt_s + dwdtd_xy_s
            dudtd_s, dvdtd_s, dwdtd_s = dudtd_z_s, dvdtd_z_s, dwdtd_z_s
        else:
            dudtd_s = dudtd_xy_s + dudtd_z_s
            dvdtd_s = dvdtd_xy_s + dvdtd_z_s
            dwdtd_s = dwdtd_xy_s + dwdtd_z_s
        return (F(dudt_s), F(dvdt_s), F(dwdt_s)), (F(dudtd_s * np.ones_like(dudt_s)),
                                                   F(dvdtd_s * np.ones_like(dudt_s)),
                                                   F(dwdtd_s * np.ones_like(dudt_s)))
    dudt_s = dudt_s + dudtd_xy_s + dudtd_z_s                # mom.f90:296-302
    dvdt_s = dvdt_s + dvdtd_xy_s + dvdtd_z_s
    dwdt_s = dwdt_s + dwdtd_xy_s + dwdtd_z_s
    return (F(dudt_s), F(dvdt_s), F(dwdt_s)), None


def bulk_forcing(n, is_forced, f, u, v, w):
    """mom.f90:311-335."""
    n1, n2, n3 = n
    for c, a in enumerate((u, v, w)):
        if is_forced[c]:
            a[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] = a[1:n1 + 1, 1:n2 + 1, 1:n3 + 1] + f[c]
