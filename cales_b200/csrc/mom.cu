// Momentum right-hand side and the low-storage RK3 update.
//   mom_xyz_ad  src/mom.f90:17-309   (one fused kernel: advection + diffusion + SGS stress divergence)
//   rk          src/rk.f90:17-121    cmpt_bulk_forcing src/rk.f90:197-222
// Layout: a CTA owns a TX x TY tile of (i,j) columns and marches in k through a z-chunk.  The four
// input fields are staged plane by plane in shared memory with a one-cell halo ring (three rotating
// planes k-1,k,k+1), so each field value is fetched from HBM/L2 once per CTA and the 13-16 point
// stencils are served from shared memory.  Arithmetic keeps the reference's association order
// (-fmad=false): bit-identical to a non-contracting CPU build.
#include "common.cuh"

#define TX 32
#define TY 8
#define PX (TX + 2)
#define PY (TY + 2)
#define PLANE (PX * PY)

// Per-thread staging descriptors: every thread moves the same (at most two) tile points of each plane, so
// the tile-local index and the global offset are computed once, not per plane.
struct Stage {
  long g0, g1;      // global offsets (without the k term) of my two tile points, -1 if outside the array
  int q0, q1;       // their positions in the PX x PY tile, -1 if none
};

__device__ __forceinline__ Stage make_stage(const Dims& d, int i0, int j0) {
  Stage st;
  const int t = threadIdx.x + TX * threadIdx.y;
  st.q0 = t;                                   // PLANE = 340 > 256 = TX*TY: first point always exists
  st.q1 = t + TX * TY < PLANE ? t + TX * TY : -1;
  {
    const int li = st.q0 % PX, lj = st.q0 / PX;
    const int i = i0 + li - 1, j = j0 + lj - 1;
    st.g0 = (i <= d.n1 + 1 && j <= d.n2 + 1) ? (long)i + d.s1 * j : -1;
  }
  if (st.q1 >= 0) {
    const int li = st.q1 % PX, lj = st.q1 / PX;
    const int i = i0 + li - 1, j = j0 + lj - 1;
    st.g1 = (i <= d.n1 + 1 && j <= d.n2 + 1) ? (long)i + d.s1 * j : -1;
  } else st.g1 = -1;
  return st;
}

template <int MODE>  // 0 explicit, 1 _IMPDIFF, 2 _IMPDIFF + _IMPDIFF_1D
__global__ void __launch_bounds__(TX* TY, 2) mom_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci, const double* __restrict__ dzfi,
                                                    double visc, const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, const double* __restrict__ s,
                                                    double* __restrict__ dudt, double* __restrict__ dvdt, double* __restrict__ dwdt,
                                                    double* __restrict__ dudtd, double* __restrict__ dvdtd, double* __restrict__ dwdtd, int kc) {
  extern __shared__ double smem[];   // [4 fields][3 planes][PLANE]
  const int i0 = blockIdx.x * TX + 1, j0 = blockIdx.y * TY + 1;
  const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  const double* fld[4] = {u, v, w, s};
  const Stage st = make_stage(d, i0, j0);
  double r0[4], r1[4];                // register stage of the plane in flight
#define FETCH(k_)                                                                      \
  {                                                                                    \
    const long ko = d.s2 * (long)(k_);                                                 \
    _Pragma("unroll") for (int f = 0; f < 4; ++f) {                                    \
      r0[f] = st.g0 >= 0 ? fld[f][st.g0 + ko] : 0.;                                    \
      r1[f] = st.g1 >= 0 ? fld[f][st.g1 + ko] : 0.;                                    \
    }                                                                                  \
  }
#define COMMIT(slot)                                                                   \
  {                                                                                    \
    _Pragma("unroll") for (int f = 0; f < 4; ++f) {                                    \
      smem[(f * 3 + (slot)) * PLANE + st.q0] = r0[f];                                  \
      if (st.q1 >= 0) smem[(f * 3 + (slot)) * PLANE + st.q1] = r1[f];                  \
    }                                                                                  \
  }
  FETCH(k0 - 1) COMMIT(0)
  FETCH(k0) COMMIT(1)
  FETCH(k0 + 1) COMMIT(2)
  __syncthreads();
  int pm = 0, pc = 1, pp = 2;
  const bool active = i <= d.n1 && j <= d.n2;
  const int c = (threadIdx.x + 1) + PX * (threadIdx.y + 1);
  const long n12 = (long)d.n1 * d.n2;
  for (int k = k0; k <= k1; ++k) {
    const bool more = k < k1;
    if (more) FETCH(k + 2)            // in flight while plane k is computed (k+2 <= n3+1)
    if (active) {
      const double* um = smem + (0 * 3 + pm) * PLANE; const double* uc = smem + (0 * 3 + pc) * PLANE; const double* up = smem + (0 * 3 + pp) * PLANE;
      const double* vm = smem + (1 * 3 + pm) * PLANE; const double* vc = smem + (1 * 3 + pc) * PLANE; const double* vp = smem + (1 * 3 + pp) * PLANE;
      const double* wm = smem + (2 * 3 + pm) * PLANE; const double* wc = smem + (2 * 3 + pc) * PLANE; const double* wp = smem + (2 * 3 + pp) * PLANE;
      const double* sm_ = smem + (3 * 3 + pm) * PLANE; const double* sc = smem + (3 * 3 + pc) * PLANE; const double* sp = smem + (3 * 3 + pp) * PLANE;
#define LD(prefix, M, C, P)                                                                                       \
  const double prefix##_ccm = M[c], prefix##_pcm = M[c + 1], prefix##_cpm = M[c + PX], prefix##_cmc = C[c - PX],     \
               prefix##_pmc = C[c + 1 - PX], prefix##_mcc = C[c - 1], prefix##_ccc = C[c], prefix##_pcc = C[c + 1],   \
               prefix##_mpc = C[c - 1 + PX], prefix##_cpc = C[c + PX], prefix##_cmp = P[c - PX], prefix##_mcp = P[c - 1], \
               prefix##_ccp = P[c];
      LD(u, um, uc, up)
      LD(v, vm, vc, vp)
      LD(w, wm, wc, wp)
      LD(s, sm_, sc, sp)
#undef LD
      const double s_ppc = sc[c + 1 + PX], s_pcp = sp[c + 1], s_cpp = sp[c + PX];
      const double dzci_k = dzci[k], dzci_km = dzci[k - 1], dzfi_k = dzfi[k], dzfi_kp = dzfi[k + 1];
      (void)u_pmc; (void)u_cpm; (void)v_pcm; (void)v_mpc; (void)w_mcp; (void)w_mpc; (void)w_cmp; (void)w_pmc;
      (void)u_pcm; (void)v_mcp; (void)s_mcp; (void)s_cmp; (void)s_mpc;
      double visc_ip, visc_im, visc_jp, visc_jm, visc_kp, visc_km;
      // ---- x momentum (mom.f90:142-186)
      visc_ip = s_pcc;
      visc_im = s_ccc;
      visc_jp = 0.25 * (s_ccc + s_pcc + s_cpc + s_ppc);
      visc_jm = 0.25 * (s_ccc + s_pcc + s_cmc + s_pmc);
      visc_kp = 0.25 * (s_ccc + s_pcc + s_ccp + s_pcp);
      visc_km = 0.25 * (s_ccc + s_pcc + s_ccm + s_pcm);
      const double dudx_ip = (u_pcc - u_ccc) * dxi;
      const double dudx_im = (u_ccc - u_mcc) * dxi;
      const double dudy_jp = (u_cpc - u_ccc) * dyi;
      const double dudy_jm = (u_ccc - u_cmc) * dyi;
      const double dudz_kp = (u_ccp - u_ccc) * dzci_k;
      const double dudz_km = (u_ccc - u_ccm) * dzci_km;
      const double dvdx_jp = (v_pcc - v_ccc) * dxi;
      const double dvdx_jm = (v_pmc - v_cmc) * dxi;
      const double dwdx_kp = (w_pcc - w_ccc) * dxi;
      const double dwdx_km = (w_pcm - w_ccm) * dxi;
      const double uu_ip = 0.25 * (u_pcc + u_ccc) * (u_ccc + u_pcc);
      const double uu_im = 0.25 * (u_mcc + u_ccc) * (u_ccc + u_mcc);
      const double vu_jp = 0.25 * (v_pcc + v_ccc) * (u_ccc + u_cpc);
      const double vu_jm = 0.25 * (v_pmc + v_cmc) * (u_ccc + u_cmc);
      const double wu_kp = 0.25 * (w_pcc + w_ccc) * (u_ccc + u_ccp);
      const double wu_km = 0.25 * (w_pcm + w_ccm) * (u_ccc + u_ccm);
      const double dudtd_xy_s = visc * (dudx_ip - dudx_im) * dxi + visc * (dudy_jp - dudy_jm) * dyi;
      const double dudtd_z_s = visc * (dudz_kp - dudz_km) * dzfi_k;
      double dudt_s = -(uu_ip - uu_im) * dxi - (vu_jp - vu_jm) * dyi - (wu_kp - wu_km) * dzfi_k +
                      (visc_ip * (dudx_ip + dudx_ip) - visc_im * (dudx_im + dudx_im)) * dxi +
                      (visc_jp * (dudy_jp + dvdx_jp) - visc_jm * (dudy_jm + dvdx_jm)) * dyi +
                      (visc_kp * (dudz_kp + dwdx_kp) - visc_km * (dudz_km + dwdx_km)) * dzfi_k;
      // ---- y momentum (mom.f90:187-231)
      visc_ip = 0.25 * (s_ccc + s_cpc + s_pcc + s_ppc);
      visc_im = 0.25 * (s_ccc + s_cpc + s_mcc + s_mpc);
      visc_jp = s_cpc;
      visc_jm = s_ccc;
      visc_kp = 0.25 * (s_ccc + s_cpc + s_ccp + s_cpp);
      visc_km = 0.25 * (s_ccc + s_cpc + s_ccm + s_cpm);
      const double dvdx_ip = (v_pcc - v_ccc) * dxi;
      const double dvdx_im = (v_ccc - v_mcc) * dxi;
      const double dvdy_jp = (v_cpc - v_ccc) * dyi;
      const double dvdy_jm = (v_ccc - v_cmc) * dyi;
      const double dvdz_kp = (v_ccp - v_ccc) * dzci_k;
      const double dvdz_km = (v_ccc - v_ccm) * dzci_km;
      const double dudy_ip = (u_cpc - u_ccc) * dyi;
      const double dudy_im = (u_mpc - u_mcc) * dyi;
      const double dwdy_kp = (w_cpc - w_ccc) * dyi;
      const double dwdy_km = (w_cpm - w_ccm) * dyi;
      const double uv_ip = 0.25 * (u_ccc + u_cpc) * (v_ccc + v_pcc);
      const double uv_im = 0.25 * (u_mcc + u_mpc) * (v_ccc + v_mcc);
      const double vv_jp = 0.25 * (v_ccc + v_cpc) * (v_ccc + v_cpc);
      const double vv_jm = 0.25 * (v_ccc + v_cmc) * (v_ccc + v_cmc);
      const double wv_kp = 0.25 * (w_ccc + w_cpc) * (v_ccc + v_ccp);
      const double wv_km = 0.25 * (w_ccm + w_cpm) * (v_ccc + v_ccm);
      const double dvdtd_xy_s = visc * (dvdx_ip - dvdx_im) * dxi + visc * (dvdy_jp - dvdy_jm) * dyi;
      const double dvdtd_z_s = visc * (dvdz_kp - dvdz_km) * dzfi_k;
      double dvdt_s = -(uv_ip - uv_im) * dxi - (vv_jp - vv_jm) * dyi - (wv_kp - wv_km) * dzfi_k +
                      (visc_ip * (dvdx_ip + dudy_ip) - visc_im * (dvdx_im + dudy_im)) * dxi +
                      (visc_jp * (dvdy_jp + dvdy_jp) - visc_jm * (dvdy_jm + dvdy_jm)) * dyi +
                      (visc_kp * (dvdz_kp + dwdy_kp) - visc_km * (dvdz_km + dwdy_km)) * dzfi_k;
      // ---- z momentum (mom.f90:232-276)
      visc_ip = 0.25 * (s_ccc + s_ccp + s_pcc + s_pcp);
      visc_im = 0.25 * (s_ccc + s_ccp + s_mcc + s_mcp);
      visc_jp = 0.25 * (s_ccc + s_ccp + s_cpc + s_cpp);
      visc_jm = 0.25 * (s_ccc + s_ccp + s_cmc + s_cmp);
      visc_kp = s_ccp;
      visc_km = s_ccc;
      const double dwdx_ip = (w_pcc - w_ccc) * dxi;
      const double dwdx_im = (w_ccc - w_mcc) * dxi;
      const double dwdy_jp = (w_cpc - w_ccc) * dyi;
      const double dwdy_jm = (w_ccc - w_cmc) * dyi;
      const double dwdz_kp = (w_ccp - w_ccc) * dzfi_kp;
      const double dwdz_km = (w_ccc - w_ccm) * dzfi_k;
      const double dudz_ip = (u_ccp - u_ccc) * dzci_k;
      const double dudz_im = (u_mcp - u_mcc) * dzci_k;
      const double dvdz_jp = (v_ccp - v_ccc) * dzci_k;
      const double dvdz_jm = (v_cmp - v_cmc) * dzci_k;
      const double uw_ip = 0.25 * (u_ccc + u_ccp) * (w_ccc + w_pcc);
      const double uw_im = 0.25 * (u_mcc + u_mcp) * (w_ccc + w_mcc);
      const double vw_jp = 0.25 * (v_ccc + v_ccp) * (w_ccc + w_cpc);
      const double vw_jm = 0.25 * (v_cmc + v_cmp) * (w_ccc + w_cmc);
      const double ww_kp = 0.25 * (w_ccc + w_ccp) * (w_ccc + w_ccp);
      const double ww_km = 0.25 * (w_ccc + w_ccm) * (w_ccc + w_ccm);
      const double dwdtd_xy_s = visc * (dwdx_ip - dwdx_im) * dxi + visc * (dwdy_jp - dwdy_jm) * dyi;
      const double dwdtd_z_s = visc * (dwdz_kp - dwdz_km) * dzci_k;
      double dwdt_s = -(uw_ip - uw_im) * dxi - (vw_jp - vw_jm) * dyi - (ww_kp - ww_km) * dzci_k +
                      (visc_ip * (dwdx_ip + dudz_ip) - visc_im * (dwdx_im + dudz_im)) * dxi +
                      (visc_jp * (dwdy_jp + dvdz_jp) - visc_jm * (dwdy_jm + dvdz_jm)) * dyi +
                      (visc_kp * (dwdz_kp + dwdz_kp) - visc_km * (dwdz_km + dwdz_km)) * dzci_k;
      const long o = (i - 1) + (long)d.n1 * (j - 1) + n12 * (k - 1);   // dudt(n1,n2,n3): no halo
      if (MODE == 0) {                                                    // mom.f90:296-302
        dudt[o] = dudt_s + dudtd_xy_s + dudtd_z_s;
        dvdt[o] = dvdt_s + dvdtd_xy_s + dvdtd_z_s;
        dwdt[o] = dwdt_s + dwdtd_xy_s + dwdtd_z_s;
      } else if (MODE == 1) {                                             // mom.f90:286-295
        dudt[o] = dudt_s; dvdt[o] = dvdt_s; dwdt[o] = dwdt_s;
        dudtd[o] = dudtd_xy_s + dudtd_z_s;
        dvdtd[o] = dvdtd_xy_s + dvdtd_z_s;
        dwdtd[o] = dwdtd_xy_s + dwdtd_z_s;
      } else {                                                            // mom.f90:278-284
        dudt[o] = dudt_s + dudtd_xy_s; dvdt[o] = dvdt_s + dvdtd_xy_s; dwdt[o] = dwdt_s + dwdtd_xy_s;
        dudtd[o] = dudtd_z_s; dvdtd[o] = dvdtd_z_s; dwdtd[o] = dwdtd_z_s;
      }
    }
    __syncthreads();                 // everyone is done reading plane k-1 (slot pm)
    if (more) COMMIT(pm)
    __syncthreads();
    const int tmp = pm; pm = pc; pc = pp; pp = tmp;
  }
#undef FETCH
#undef COMMIT
}

static int mom_launch(cales_ctx* ctx, const int n[3], double dxi, double dyi, const double* dzci, const double* dzfi, double visc,
                      const double* u, const double* v, const double* w, const double* visct, double* dudt, double* dvdt,
                      double* dwdt, double* dudtd, double* dvdtd, double* dwdtd) {
  Dims d(n);
  long cols = (long)cdiv(n[0], TX) * cdiv(n[1], TY);
  const int kc = pick_chunk(cols, n[2], 148 * 2, 12, 2);
  dim3 g(cdiv(n[0], TX), cdiv(n[1], TY), cdiv(n[2], kc)), b(TX, TY);
  const size_t sh = 12 * PLANE * sizeof(double);
  if (ctx->diffusion == CALES_DIFF_EXPLICIT)
    mom_k<0><<<g, b, sh, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, dudt, dvdt, dwdt, dudtd, dvdtd, dwdtd, kc);
  else if (ctx->diffusion == CALES_DIFF_IMPLICIT_3D)
    mom_k<1><<<g, b, sh, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, dudt, dvdt, dwdt, dudtd, dvdtd, dwdtd, kc);
  else
    mom_k<2><<<g, b, sh, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, dudt, dvdt, dwdt, dudtd, dvdtd, dwdtd, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

extern "C" int cales_mom_xyz_ad(cales_ctx* ctx, const int n[3], double dxi, double dyi, const double* dzci, const double* dzfi,
                                double visc, const double* u, const double* v, const double* w, const double* visct,
                                double* dudt, double* dvdt, double* dwdt, double* dudtd, double* dvdtd, double* dwdtd) {
  CHECK_CTX(ctx);
  if (ctx->diffusion != CALES_DIFF_EXPLICIT && (!dudtd || !dvdtd || !dwdtd))
    return cales_fail(ctx, CALES_ERR_INVALID, "mom_xyz_ad: dudtd/dvdtd/dwdtd required with implicit diffusion");
  return mom_launch(ctx, n, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, dudt, dvdt, dwdt, dudtd, dvdtd, dwdtd);
}

// ---- RK update (rk.f90:77-94) ---------------------------------------------------------------------------
#define BX 64
#define BY 4
// F2: factor2 != 0 (the first RK substep has rkpar(2) = 0: the old right-hand side is not read at all; adding
// 0*dudtrko changes no bit of a finite sum).  Loads of RU consecutive levels are issued together before any store so
// that each thread keeps ~10*RU independent requests in flight.
#define RU 2
template <int IMP, int F2>
__global__ void __launch_bounds__(BX* BY) rk_update_k(Dims d, double f1, double f2, double f12, double dxi, double dyi,
                                                       const double* __restrict__ dzci, double bfx, double bfy, double bfz,
                                                       const double* __restrict__ p, const double* __restrict__ du,
                                                       const double* __restrict__ dv, const double* __restrict__ dw,
                                                       const double* __restrict__ duo, const double* __restrict__ dvo,
                                                       const double* __restrict__ dwo, const double* __restrict__ dud,
                                                       const double* __restrict__ dvd, const double* __restrict__ dwd,
                                                       double* __restrict__ u, double* __restrict__ v, double* __restrict__ w, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  const long n12 = (long)d.n1 * d.n2, s1 = d.s1, s2 = d.s2;
  long o = (i - 1) + (long)d.n1 * (j - 1) + n12 * (k0 - 1);
  double pc = p[c];
  int k = k0;
  for (; k + RU - 1 <= k1; k += RU, c += RU * s2, o += RU * n12) {
    double uu[RU], vv[RU], ww[RU], pk[RU], pi[RU], pj[RU], a[RU], b[RU], e[RU], ao[RU], bo[RU], eo[RU], ad[RU], bd[RU], ed[RU], dz[RU];
#pragma unroll
    for (int q = 0; q < RU; ++q) {
      const long cq = c + q * s2, oq = o + q * n12;
      uu[q] = u[cq]; vv[q] = v[cq]; ww[q] = w[cq];
      pk[q] = p[cq + s2]; pi[q] = p[cq + 1]; pj[q] = p[cq + s1];
      a[q] = du[oq]; b[q] = dv[oq]; e[q] = dw[oq];
      if (F2) { ao[q] = duo[oq]; bo[q] = dvo[oq]; eo[q] = dwo[oq]; }
      if (IMP) { ad[q] = dud[oq]; bd[q] = dvd[oq]; ed[q] = dwd[oq]; }
      dz[q] = dzci[k + q];
    }
#pragma unroll
    for (int q = 0; q < RU; ++q) {
      const long cq = c + q * s2;
      double un, vn, wn;
      if (F2) {
        un = uu[q] + f1 * a[q] + f2 * ao[q] + f12 * (bfx - dxi * (pi[q] - pc));
        vn = vv[q] + f1 * b[q] + f2 * bo[q] + f12 * (bfy - dyi * (pj[q] - pc));
        wn = ww[q] + f1 * e[q] + f2 * eo[q] + f12 * (bfz - dz[q] * (pk[q] - pc));
      } else {
        un = uu[q] + f1 * a[q] + f12 * (bfx - dxi * (pi[q] - pc));
        vn = vv[q] + f1 * b[q] + f12 * (bfy - dyi * (pj[q] - pc));
        wn = ww[q] + f1 * e[q] + f12 * (bfz - dz[q] * (pk[q] - pc));
      }
      if (IMP) { un = un + f12 * ad[q]; vn = vn + f12 * bd[q]; wn = wn + f12 * ed[q]; }
      u[cq] = un; v[cq] = vn; w[cq] = wn;
      pc = pk[q];
    }
  }
  for (; k <= k1; ++k, c += s2, o += n12) {
    const double pk = p[c + s2];
    double un = u[c] + f1 * du[o], vn = v[c] + f1 * dv[o], wn = w[c] + f1 * dw[o];
    if (F2) { un = un + f2 * duo[o]; vn = vn + f2 * dvo[o]; wn = wn + f2 * dwo[o]; }
    un = un + f12 * (bfx - dxi * (p[c + 1] - pc));
    vn = vn + f12 * (bfy - dyi * (p[c + s1] - pc));
    wn = wn + f12 * (bfz - dzci[k] * (pk - pc));
    if (IMP) { un = un + f12 * dud[o]; vn = vn + f12 * dvd[o]; wn = wn + f12 * dwd[o]; }
    u[c] = un; v[c] = vn; w[c] = wn;
    pc = pk;
  }
}

// rk.f90:110-119
__global__ void __launch_bounds__(BX* BY) rk_imprhs_k(Dims d, double hf12, const double* __restrict__ dud, const double* __restrict__ dvd,
                                                       const double* __restrict__ dwd, double* __restrict__ u, double* __restrict__ v,
                                                       double* __restrict__ w, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  const long n12 = (long)d.n1 * d.n2;
  long o = (i - 1) + (long)d.n1 * (j - 1) + n12 * (k0 - 1);
  for (int k = k0; k <= k1; ++k, c += d.s2, o += n12) {
    u[c] = u[c] - hf12 * dud[o]; v[c] = v[c] - hf12 * dvd[o]; w[c] = w[c] - hf12 * dwd[o];
  }
}

__global__ void bulkf_k(const double* __restrict__ mean, int is_u, int is_v, int is_w, double vu, double vv, double vw, double* __restrict__ f) {
  f[0] = is_u ? vu - mean[0] : 0.;
  f[1] = is_v ? vv - mean[1] : 0.;
  f[2] = is_w ? vw - mean[2] : 0.;
}

int k_bulk_mean_dev(cales_ctx* ctx, const int n[3], const double* gvr, const double* p, double* out);

// device-resident rk: leaves f(3) in ctx->fdev, no host synchronisation
int k_rk_dev(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci, const double* dzfi,
             const double* gvr_c, const double* gvr_f, double visc, double dt, const double* p, const int is_forced[3],
             const double velf[3], const double bforce[3], const double* visct, double* u, double* v, double* w) {
  const double factor1 = rkpar[0] * dt, factor2 = rkpar[1] * dt, factor12 = factor1 + factor2;
  const size_t nb = (size_t)n[0] * n[1] * n[2] * sizeof(double);
  const bool imp = ctx->diffusion != CALES_DIFF_EXPLICIT;
  double* r[9];
  static const char* names[9] = {"rk_du0", "rk_dv0", "rk_dw0", "rk_du1", "rk_dv1", "rk_dw1", "rk_dud", "rk_dvd", "rk_dwd"};
  for (int q = 0; q < (imp ? 9 : 6); ++q) {
    r[q] = (double*)cales_scratch(ctx, names[q], nb);
    if (!r[q]) return CALES_ERR_NOMEM;
  }
  if (!imp) r[6] = r[7] = r[8] = nullptr;
  if (ctx->rk_first) {                               // rk.f90:45-72: dudtrko = 0 on first call
    ctx->rk_first = false;
    ctx->rk_swap = 0;
    for (int q = 3; q < 6; ++q) CUDA_TRY(ctx, cudaMemsetAsync(r[q], 0, nb, ctx->stream));
  }
  double** nw = ctx->rk_swap ? r + 3 : r;            // dudtrk
  double** ol = ctx->rk_swap ? r : r + 3;            // dudtrko
  int rc = mom_launch(ctx, n, dli[0], dli[1], dzci, dzfi, visc, u, v, w, visct, nw[0], nw[1], nw[2], r[6], r[7], r[8]);
  if (rc) return rc;
  Dims d(n);
  long cols = (long)cdiv(n[0], BX) * cdiv(n[1], BY);
  const int kc = pick_chunk(cols, n[2], 148 * 4, 8, 1);
  dim3 g(cdiv(n[0], BX), cdiv(n[1], BY), cdiv(n[2], kc)), b(BX, BY);
#define RKU(IMP_, F2_) rk_update_k<IMP_, F2_><<<g, b, 0, ctx->stream>>>(d, factor1, factor2, factor12, dli[0], dli[1], dzci, bforce[0], bforce[1], \
                                                                         bforce[2], p, nw[0], nw[1], nw[2], ol[0], ol[1], ol[2], r[6], r[7], r[8], u, v, w, kc)
  if (imp) { if (factor2 != 0.) RKU(1, 1); else RKU(1, 0); }
  else { if (factor2 != 0.) RKU(0, 1); else RKU(0, 0); }
#undef RKU
  KERNEL_CHECK(ctx);
  ctx->rk_swap ^= 1;                                 // rk.f90:98-100
  // cmpt_bulk_forcing (rk.f90:197-222)
  double* mean = ctx->red + 8;
  CUDA_TRY(ctx, cudaMemsetAsync(mean, 0, 3 * sizeof(double), ctx->stream));
  if (is_forced[0] && (rc = k_bulk_mean_dev(ctx, n, gvr_f, u, mean + 0))) return rc;
  if (is_forced[1] && (rc = k_bulk_mean_dev(ctx, n, gvr_f, v, mean + 1))) return rc;
  if (is_forced[2] && (rc = k_bulk_mean_dev(ctx, n, gvr_c, w, mean + 2))) return rc;
  bulkf_k<<<1, 1, 0, ctx->stream>>>(mean, is_forced[0], is_forced[1], is_forced[2], velf[0], velf[1], velf[2], ctx->fdev);
  KERNEL_CHECK(ctx);
  if (imp) {
    rk_imprhs_k<<<g, b, 0, ctx->stream>>>(d, .5 * factor12, r[6], r[7], r[8], u, v, w, kc);
    KERNEL_CHECK(ctx);
  }
  return CALES_OK;
}

extern "C" int cales_rk(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci,
                        const double* dzfi, const double* grid_vol_ratio_c, const double* grid_vol_ratio_f, double visc,
                        double dt, const double* p, const int is_forced[3], const double velf[3], const double bforce[3],
                        const double* visct, double* u, double* v, double* w, double f[3]) {
  CHECK_CTX(ctx);
  int rc = k_rk_dev(ctx, rkpar, n, dli, dzci, dzfi, grid_vol_ratio_c, grid_vol_ratio_f, visc, dt, p, is_forced, velf, bforce, visct, u, v, w);
  if (rc) return rc;
  if (!f) return CALES_OK;                             // f(3) stays on the device: no synchronisation (see cales_bulk_forcing)
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red_host + 8, ctx->fdev, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < 3; ++c) f[c] = ctx->red_host[8 + c];
  return CALES_OK;
}
