// Momentum right-hand side and the low-storage RK3 update.
//   mom_xyz_ad  src/mom.f90:17-309   (one fused kernel: advection + diffusion + SGS stress divergence)
//   rk          src/rk.f90:17-121    cmpt_bulk_forcing src/rk.f90:197-222
// Layout: a CTA owns a TX x TY tile of (i,j) columns and marches in k through a z-chunk.  The four
// input fields are staged plane by plane in shared memory with a one-cell halo ring by cp.async (three
// rotating slots: planes k and k+1 in use, k+2 in flight; one barrier per plane), so each field value is
// fetched from HBM/L2 once per CTA and the stencils are served from shared memory.  Arithmetic keeps the
// reference's association order (-fmad=false): bit-identical to a non-contracting CPU build.
#include "common.cuh"

#include "tile.cuh"

// Shared sub-expressions.  Every face/edge quantity of the staggered stencil is needed by two cells and (for the
// cross terms) by two momentum components; mom.f90 writes it out at each use (e.g. wu_km of level k is wu_kp of level
// k-1, uv_ip is vu_jp with the factors swapped).  Floating-point + and * are commutative, so those duplicates are the
// same bits, and this kernel computes each of them once:
//   * the k-1/2 quantities of level k are the k+1/2 quantities of level k-1 -> carried in registers along the z march
//     (9 values + 3 eddy viscosities), so plane k-1 is not read at all;
//   * the 0.25 of the advective averages and of the four-point eddy-viscosity averages is a power of two: scaling by
//     it commutes with rounding, so it is folded into the metric factor of the difference, 0.25*dxi etc. (exact unless a
//     flux underflows, |flux| < 2^-1020).
// The association order of everything else is the reference's; the library is built with -fmad=false.
// Fused RK update (RK > 0, explicit diffusion): the low-storage update of rk.f90:77-94 is applied to the right-hand side
// while it is still in registers -- the pressure is staged as a fifth tile field, the old right-hand side (RK == 2:
// rkpar(2) /= 0) is read once, and the new velocity goes to a SECOND set of arrays (un,vn,wn: the stencil of the
// neighbouring tiles still reads the old one), so that rk = mom + update moves 112 instead of 160 B/cell.  Same
// operations in the same order as mom_k + rk_update_k: identical bits.
struct RkFuse {
  double f1, f2, f12, bfx, bfy, bfz;
  const double* p;                       // pressure (haloed)
  const double *duo, *dvo, *dwo;         // old right-hand side (halo-free)
  double *un, *vn, *wn;                  // updated velocity (haloed; interior written)
};

template <int MODE, bool V16, int RK>  // MODE 0 explicit, 1 _IMPDIFF, 2 _IMPDIFF + _IMPDIFF_1D; V16: 16-byte plane staging (tile.cuh)
__global__ void __launch_bounds__(TX* TY, 2) mom_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci, const double* __restrict__ dzfi,
                                                    double visc, const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, const double* __restrict__ s,
                                                    double* __restrict__ dudt, double* __restrict__ dvdt, double* __restrict__ dwdt,
                                                    double* __restrict__ dudtd, double* __restrict__ dvdtd, double* __restrict__ dwdtd, int kc,
                                                    const __grid_constant__ RkFuse R) {
  constexpr int NF = RK ? 5 : 4;
  extern __shared__ __align__(16) double smem[];   // [4 slots][NF fields][PLANE]: planes k, k+1 in use, k+2 and k+3 in flight
  const int i0 = blockIdx.x * TX + 1, j0 = blockIdx.y * TY + 1;
  const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  Stager<NF, V16> st(d, i0, j0, u, v, w, s, smem, k0 - 1, k1 + 1, R.p);
  const bool active = i <= d.n1 && j <= d.n2;
  const long n12 = (long)d.n1 * d.n2;
  // RK == 2: the old right-hand side of my cell rides the same ring (thread-private 8-byte copies issued with the plane
  // of its level, two levels ahead of its use), so its HBM latency is hidden like that of the tiles
  double* const fl = smem + TSLOTS * NF * PLANE + (threadIdx.x + TX * threadIdx.y);      // [slot][3][TX*TY], my element
  long ofl = (i - 1) + (long)d.n1 * (j - 1) + n12 * (k0 - 2);                            // my cell at plane st.knext (= k0-1 now)
  auto flat = [&](auto sl_) {
    constexpr int SLT = decltype(sl_)::v;
    if (RK == 2) {
      const int kp = st.knext - 1;                                                       // the plane just issued
      if (active && kp >= k0 && kp <= k1) {
        tile_cp8(fl + (SLT * 3 + 0) * (TX * TY), R.duo + ofl);
        tile_cp8(fl + (SLT * 3 + 1) * (TX * TY), R.dvo + ofl);
        tile_cp8(fl + (SLT * 3 + 2) * (TX * TY), R.dwo + ofl);
      }
      ofl += n12;
    }
    tile_commit();
  };
  st.template issue<0, false>(); flat(Slot<0>{});
  st.template issue<1, false>(); flat(Slot<1>{});
  st.template issue<2, false>(); flat(Slot<2>{});
  st.template issue<3, false>(); flat(Slot<3>{});
  tile_wait_1();
  __syncthreads();
  const double* const sm = smem + (threadIdx.x + 1) + PX * (threadIdx.y + 1);     // my cell in field 0 of slot 0
  constexpr int SL = NF * PLANE;
  const double qdxi = 0.25 * dxi, qdyi = 0.25 * dyi;
  // k-1/2 quantities of level k0 = k+1/2 quantities of level k0-1 (planes k0-1, k0 = slots 0, 1)
  double wu_km, dudz_km, sxz_km, wv_km, dvdz_km, syz_km, ww_km, dwdz_km, fzz_km, s_ccm, s_pcm, s_cpm;
  {
    const double* uc = sm; const double* vc = sm + PLANE; const double* wc = sm + 2 * PLANE; const double* sc = sm + 3 * PLANE;
    const double* up = uc + SL; const double* vp = vc + SL; const double* wp = wc + SL; const double* sp = sc + SL;
    const double dzci_m = dzci[k0 - 1], dzfi_k = dzfi[k0];
    const double u_ccc = uc[0], u_ccp = up[0], v_ccc = vc[0], v_ccp = vp[0], w_ccc = wc[0], w_pcc = wc[1], w_cpc = wc[PX], w_ccp = wp[0];
    wu_km = (w_pcc + w_ccc) * (u_ccc + u_ccp);
    dudz_km = (u_ccp - u_ccc) * dzci_m;
    sxz_km = dudz_km + (w_pcc - w_ccc) * dxi;
    wv_km = (w_ccc + w_cpc) * (v_ccc + v_ccp);
    dvdz_km = (v_ccp - v_ccc) * dzci_m;
    syz_km = dvdz_km + (w_cpc - w_ccc) * dyi;
    ww_km = (w_ccc + w_ccp) * (w_ccc + w_ccp);
    dwdz_km = (w_ccp - w_ccc) * dzfi_k;
    fzz_km = sp[0] * (dwdz_km + dwdz_km);
    s_ccm = sc[0]; s_pcm = sc[1]; s_cpm = sc[PX];
  }
  __syncthreads();                   // slot 0 (plane k0-1) may now be overwritten
  long o = (i - 1) + (long)d.n1 * (j - 1) + n12 * (k0 - 1);   // dudt(n1,n2,n3): no halo
  long ch = RK ? d.idx(i, j, k0) : 0;                          // the same cell in a haloed array
  int k = k0;
  // one level: planes k, k+1 in slots SC, SP; plane k+3 goes into slot SN (which held plane k-1).  Threads outside the
  // array compute on whatever their tile cells hold and store nothing.
  auto step = [&](auto sc_, auto sp_, auto sn_) {
    constexpr int SC = decltype(sc_)::v, SP = decltype(sp_)::v, SN = decltype(sn_)::v;
    st.template issue<SN, false>(); flat(Slot<SN>{});
    {
      const double* uc = sm + SC * SL; const double* up = sm + SP * SL;
      const double* vc = uc + PLANE; const double* vp = up + PLANE;
      const double* wc = uc + 2 * PLANE; const double* wp = up + 2 * PLANE;
      const double* sc = uc + 3 * PLANE; const double* sp = up + 3 * PLANE;
      const double u_cmc = uc[-PX], u_mcc = uc[-1], u_ccc = uc[0], u_pcc = uc[1], u_mpc = uc[PX - 1], u_cpc = uc[PX];
      const double u_mcp = up[-1], u_ccp = up[0];
      const double v_cmc = vc[-PX], v_pmc = vc[1 - PX], v_mcc = vc[-1], v_ccc = vc[0], v_pcc = vc[1], v_cpc = vc[PX];
      const double v_cmp = vp[-PX], v_ccp = vp[0];
      const double w_cmc = wc[-PX], w_mcc = wc[-1], w_ccc = wc[0], w_pcc = wc[1], w_cpc = wc[PX], w_ccp = wp[0];
      const double s_cmc = sc[-PX], s_pmc = sc[1 - PX], s_mcc = sc[-1], s_ccc = sc[0], s_pcc = sc[1], s_mpc = sc[PX - 1],
                   s_cpc = sc[PX], s_ppc = sc[1 + PX];
      const double s_cmp = sp[-PX], s_mcp = sp[-1], s_ccp = sp[0], s_pcp = sp[1], s_cpp = sp[PX];
      const double dzci_k = dzci[k], dzfi_k = dzfi[k], dzfi_kp = dzfi[k + 1];
      const double qdzfi_k = 0.25 * dzfi_k, qdzci_k = 0.25 * dzci_k;
      // ---- face/edge quantities on the + side of this cell (shared with the neighbour at k+1 through the carry,
      //      and between the two momentum components that meet at the edge)
      const double dudx_ip = (u_pcc - u_ccc) * dxi;
      const double dudy_jp = (u_cpc - u_ccc) * dyi;              // = dudy_ip of the y equation
      const double dudz_kp = (u_ccp - u_ccc) * dzci_k;           // = dudz_ip of the z equation
      const double dvdx_jp = (v_pcc - v_ccc) * dxi;              // = dvdx_ip
      const double dvdy_jp = (v_cpc - v_ccc) * dyi;
      const double dvdz_kp = (v_ccp - v_ccc) * dzci_k;           // = dvdz_jp of the z equation
      const double dwdx_kp = (w_pcc - w_ccc) * dxi;              // = dwdx_ip
      const double dwdy_kp = (w_cpc - w_ccc) * dyi;              // = dwdy_jp
      const double dwdz_kp = (w_ccp - w_ccc) * dzfi_kp;
      const double sxy_p = dudy_jp + dvdx_jp;                    // (dudy_jp+dvdx_jp) = (dvdx_ip+dudy_ip)
      const double sxz_p = dudz_kp + dwdx_kp;                    // (dudz_kp+dwdx_kp) = (dwdx_ip+dudz_ip)
      const double syz_p = dvdz_kp + dwdy_kp;                    // (dvdz_kp+dwdy_kp) = (dwdy_jp+dvdz_jp)
      const double uv_p = (v_pcc + v_ccc) * (u_ccc + u_cpc);     // 4 vu_jp = 4 uv_ip
      const double uw_p = (w_pcc + w_ccc) * (u_ccc + u_ccp);     // 4 wu_kp = 4 uw_ip
      const double vw_p = (w_ccc + w_cpc) * (v_ccc + v_ccp);     // 4 wv_kp = 4 vw_jp
      // ---- x momentum (mom.f90:142-186)
      const double sx = s_ccc + s_pcc;
      const double dudx_im = (u_ccc - u_mcc) * dxi;
      const double dudy_jm = (u_ccc - u_cmc) * dyi;
      const double dvdx_jm = (v_pmc - v_cmc) * dxi;
      const double uu_ip = (u_pcc + u_ccc) * (u_ccc + u_pcc);
      const double uu_im = (u_mcc + u_ccc) * (u_ccc + u_mcc);
      const double vu_jm = (v_pmc + v_cmc) * (u_ccc + u_cmc);
      const double dudtd_xy_s = visc * (dudx_ip - dudx_im) * dxi + visc * (dudy_jp - dudy_jm) * dyi;
      const double dudtd_z_s = visc * (dudz_kp - dudz_km) * dzfi_k;
      const double dudt_s = -(uu_ip - uu_im) * qdxi - (uv_p - vu_jm) * qdyi - (uw_p - wu_km) * qdzfi_k +
                            (s_pcc * (dudx_ip + dudx_ip) - s_ccc * (dudx_im + dudx_im)) * dxi +
                            ((sx + s_cpc + s_ppc) * sxy_p - (sx + s_cmc + s_pmc) * (dudy_jm + dvdx_jm)) * qdyi +
                            ((sx + s_ccp + s_pcp) * sxz_p - (sx + s_ccm + s_pcm) * sxz_km) * qdzfi_k;
      // ---- y momentum (mom.f90:187-231)
      const double sy = s_ccc + s_cpc;
      const double dvdx_im = (v_ccc - v_mcc) * dxi;
      const double dvdy_jm = (v_ccc - v_cmc) * dyi;
      const double dudy_im = (u_mpc - u_mcc) * dyi;
      const double uv_im = (u_mcc + u_mpc) * (v_ccc + v_mcc);
      const double vv_jp = (v_ccc + v_cpc) * (v_ccc + v_cpc);
      const double vv_jm = (v_ccc + v_cmc) * (v_ccc + v_cmc);
      const double dvdtd_xy_s = visc * (dvdx_jp - dvdx_im) * dxi + visc * (dvdy_jp - dvdy_jm) * dyi;
      const double dvdtd_z_s = visc * (dvdz_kp - dvdz_km) * dzfi_k;
      const double dvdt_s = -(uv_p - uv_im) * qdxi - (vv_jp - vv_jm) * qdyi - (vw_p - wv_km) * qdzfi_k +
                            ((sy + s_pcc + s_ppc) * sxy_p - (sy + s_mcc + s_mpc) * (dvdx_im + dudy_im)) * qdxi +
                            (s_cpc * (dvdy_jp + dvdy_jp) - s_ccc * (dvdy_jm + dvdy_jm)) * dyi +
                            ((sy + s_ccp + s_cpp) * syz_p - (sy + s_ccm + s_cpm) * syz_km) * qdzfi_k;
      // ---- z momentum (mom.f90:232-276)
      const double sz = s_ccc + s_ccp;
      const double dwdx_im = (w_ccc - w_mcc) * dxi;
      const double dwdy_jm = (w_ccc - w_cmc) * dyi;
      const double dudz_im = (u_mcp - u_mcc) * dzci_k;
      const double dvdz_jm = (v_cmp - v_cmc) * dzci_k;
      const double uw_im = (u_mcc + u_mcp) * (w_ccc + w_mcc);
      const double vw_jm = (v_cmc + v_cmp) * (w_ccc + w_cmc);
      const double ww_kp = (w_ccc + w_ccp) * (w_ccc + w_ccp);
      const double fzz_kp = s_ccp * (dwdz_kp + dwdz_kp);
      const double dwdtd_xy_s = visc * (dwdx_kp - dwdx_im) * dxi + visc * (dwdy_kp - dwdy_jm) * dyi;
      const double dwdtd_z_s = visc * (dwdz_kp - dwdz_km) * dzci_k;
      const double dwdt_s = -(uw_p - uw_im) * qdxi - (vw_p - vw_jm) * qdyi - (ww_kp - ww_km) * qdzci_k +
                            ((sz + s_pcc + s_pcp) * sxz_p - (sz + s_mcc + s_mcp) * (dwdx_im + dudz_im)) * qdxi +
                            ((sz + s_cpc + s_cpp) * syz_p - (sz + s_cmc + s_cmp) * (dwdy_jm + dvdz_jm)) * qdyi +
                            (fzz_kp - fzz_km) * dzci_k;
      if (active) {
        if (MODE == 0) {                                                    // mom.f90:296-302
          const double du_ = dudt_s + dudtd_xy_s + dudtd_z_s;
          const double dv_ = dvdt_s + dvdtd_xy_s + dvdtd_z_s;
          const double dw_ = dwdt_s + dwdtd_xy_s + dwdtd_z_s;
          dudt[o] = du_; dvdt[o] = dv_; dwdt[o] = dw_;
          if (RK) {                                                         // rk.f90:77-86
            const double* pc_ = uc + 4 * PLANE;
            const double p_ccc = pc_[0];
            double un_, vn_, wn_;
            if (RK == 2) {
              const double* fo = fl + SC * 3 * (TX * TY);
              un_ = u_ccc + R.f1 * du_ + R.f2 * fo[0] + R.f12 * (R.bfx - dxi * (pc_[1] - p_ccc));
              vn_ = v_ccc + R.f1 * dv_ + R.f2 * fo[TX * TY] + R.f12 * (R.bfy - dyi * (pc_[PX] - p_ccc));
              wn_ = w_ccc + R.f1 * dw_ + R.f2 * fo[2 * TX * TY] + R.f12 * (R.bfz - dzci_k * (up[4 * PLANE] - p_ccc));
            } else {
              un_ = u_ccc + R.f1 * du_ + R.f12 * (R.bfx - dxi * (pc_[1] - p_ccc));
              vn_ = v_ccc + R.f1 * dv_ + R.f12 * (R.bfy - dyi * (pc_[PX] - p_ccc));
              wn_ = w_ccc + R.f1 * dw_ + R.f12 * (R.bfz - dzci_k * (up[4 * PLANE] - p_ccc));
            }
            R.un[ch] = un_; R.vn[ch] = vn_; R.wn[ch] = wn_;
          }
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
      wu_km = uw_p; dudz_km = dudz_kp; sxz_km = sxz_p;
      wv_km = vw_p; dvdz_km = dvdz_kp; syz_km = syz_p;
      ww_km = ww_kp; dwdz_km = dwdz_kp; fzz_km = fzz_kp;
      s_ccm = s_ccc; s_pcm = s_pcc; s_cpm = s_cpc;
    }
    tile_wait_1();                   // plane k+2 has landed (k+3 may still be in flight)
    __syncthreads();                 // ... for everyone, and everyone is done reading plane k
    o += n12;
    if (RK) ch += d.s2;
    return ++k <= k1;
  };
  while (step(Slot<1>{}, Slot<2>{}, Slot<0>{}) && step(Slot<2>{}, Slot<3>{}, Slot<1>{}) && step(Slot<3>{}, Slot<0>{}, Slot<2>{}) &&
         step(Slot<0>{}, Slot<1>{}, Slot<3>{})) {}
}

static int mom_launch(cales_ctx* ctx, const int n[3], double dxi, double dyi, const double* dzci, const double* dzfi, double visc,
                      const double* u, const double* v, const double* w, const double* visct, double* dudt, double* dvdt,
                      double* dwdt, double* dudtd, double* dvdtd, double* dwdtd, const RkFuse* rk = nullptr) {
  Dims d(n);
  long cols = (long)cdiv(n[0], TX) * cdiv(n[1], TY);
  const int kc = pick_chunk(cols, n[2], 148 * 2, 12, 2);
  dim3 g(cdiv(n[0], TX), cdiv(n[1], TY), cdiv(n[2], kc)), b(TX, TY);
  const size_t sh = (TSLOTS * (rk ? 5 : 4) * PLANE + (rk && rk->f2 != 0. ? TSLOTS * 3 * TX * TY : 0)) * sizeof(double);
  const bool v16 = tile_v16(n[0], u, v, w, visct) && (!rk || ((uintptr_t)rk->p & 15) == 0);
  RkFuse R;
  memset(&R, 0, sizeof R);
  if (rk) R = *rk;
#define MOM_GO(M_, V_, R_)                                                                                           \
  {                                                                                                                  \
    static bool attr = false;                                                                                        \
    if (!attr) { attr = true; cudaFuncSetAttribute(mom_k<M_, V_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((TSLOTS * 5 * PLANE + TSLOTS * 3 * TX * TY) * sizeof(double))); } \
    mom_k<M_, V_, R_><<<g, b, sh, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, u, v, w, visct, dudt, dvdt, dwdt, dudtd, dvdtd, dwdtd, kc, R); \
  }
  if (rk) {
    if (ctx->diffusion != CALES_DIFF_EXPLICIT) return cales_fail(ctx, CALES_ERR_INVALID, "fused mom+rk update needs explicit diffusion");
    if (rk->f2 != 0.) { if (v16) MOM_GO(0, true, 2) else MOM_GO(0, false, 2) }
    else { if (v16) MOM_GO(0, true, 1) else MOM_GO(0, false, 1) }
  } else if (ctx->diffusion == CALES_DIFF_EXPLICIT) { if (v16) MOM_GO(0, true, 0) else MOM_GO(0, false, 0) }
  else if (ctx->diffusion == CALES_DIFF_IMPLICIT_3D) { if (v16) MOM_GO(1, true, 0) else MOM_GO(1, false, 0) }
  else { if (v16) MOM_GO(2, true, 0) else MOM_GO(2, false, 0) }
#undef MOM_GO
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

// device-resident rk: leaves f(3) in ctx->fdev, no host synchronisation.  With un/vn/wn (explicit diffusion only) the
// update is fused into the momentum kernel and the new velocity is written THERE instead of in place (see RkFuse).
int k_rk_dev(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci, const double* dzfi,
             const double* gvr_c, const double* gvr_f, double visc, double dt, const double* p, const int is_forced[3],
             const double velf[3], const double bforce[3], const double* visct, double* u, double* v, double* w,
             double* un = nullptr, double* vn = nullptr, double* wn = nullptr) {
  const double factor1 = rkpar[0] * dt, factor2 = rkpar[1] * dt, factor12 = factor1 + factor2;
  const size_t nb = (size_t)n[0] * n[1] * n[2] * sizeof(double);
  const bool imp = ctx->diffusion != CALES_DIFF_EXPLICIT;
  const bool fused = un != nullptr;
  if (fused && imp) return cales_fail(ctx, CALES_ERR_INVALID, "fused rk needs explicit diffusion");
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
  int rc;
  Dims d(n);
  long cols = (long)cdiv(n[0], BX) * cdiv(n[1], BY);
  const int kc = pick_chunk(cols, n[2], 148 * 4, 8, 1);
  dim3 g(cdiv(n[0], BX), cdiv(n[1], BY), cdiv(n[2], kc)), b(BX, BY);
  if (fused) {
    RkFuse R;
    R.f1 = factor1; R.f2 = factor2; R.f12 = factor12; R.bfx = bforce[0]; R.bfy = bforce[1]; R.bfz = bforce[2];
    R.p = p; R.duo = ol[0]; R.dvo = ol[1]; R.dwo = ol[2]; R.un = un; R.vn = vn; R.wn = wn;
    if ((rc = mom_launch(ctx, n, dli[0], dli[1], dzci, dzfi, visc, u, v, w, visct, nw[0], nw[1], nw[2], nullptr, nullptr, nullptr, &R))) return rc;
  } else {
    if ((rc = mom_launch(ctx, n, dli[0], dli[1], dzci, dzfi, visc, u, v, w, visct, nw[0], nw[1], nw[2], r[6], r[7], r[8]))) return rc;
#define RKU(IMP_, F2_) rk_update_k<IMP_, F2_><<<g, b, 0, ctx->stream>>>(d, factor1, factor2, factor12, dli[0], dli[1], dzci, bforce[0], bforce[1], \
                                                                         bforce[2], p, nw[0], nw[1], nw[2], ol[0], ol[1], ol[2], r[6], r[7], r[8], u, v, w, kc)
    if (imp) { if (factor2 != 0.) RKU(1, 1); else RKU(1, 0); }
    else { if (factor2 != 0.) RKU(0, 1); else RKU(0, 0); }
#undef RKU
    KERNEL_CHECK(ctx);
  }
  ctx->rk_swap ^= 1;                                 // rk.f90:98-100
  // cmpt_bulk_forcing (rk.f90:197-222)
  double* uo = fused ? un : u; double* vo = fused ? vn : v; double* wo = fused ? wn : w;
  double* mean = ctx->red + 8;
  CUDA_TRY(ctx, cudaMemsetAsync(mean, 0, 3 * sizeof(double), ctx->stream));
  if (is_forced[0] && (rc = k_bulk_mean_dev(ctx, n, gvr_f, uo, mean + 0))) return rc;
  if (is_forced[1] && (rc = k_bulk_mean_dev(ctx, n, gvr_f, vo, mean + 1))) return rc;
  if (is_forced[2] && (rc = k_bulk_mean_dev(ctx, n, gvr_c, wo, mean + 2))) return rc;
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

// rk with the update fused into the momentum kernel (explicit diffusion): u,v,w are read only, the updated velocity goes to
// un,vn,wn (haloed arrays; interior written).  f(3) stays on the device (cales_bulk_forcing(f = NULL) consumes it).
extern "C" int cales_rk_fused(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci,
                              const double* dzfi, const double* grid_vol_ratio_c, const double* grid_vol_ratio_f, double visc,
                              double dt, const double* p, const int is_forced[3], const double velf[3], const double bforce[3],
                              const double* visct, const double* u, const double* v, const double* w, double* un, double* vn, double* wn) {
  CHECK_CTX(ctx);
  if (!un || !vn || !wn || un == u || vn == v || wn == w) return cales_fail(ctx, CALES_ERR_INVALID, "rk_fused: un,vn,wn must be arrays other than u,v,w");
  return k_rk_dev(ctx, rkpar, n, dli, dzci, dzfi, grid_vol_ratio_c, grid_vol_ratio_f, visc, dt, p, is_forced, velf, bforce, visct,
                  (double*)u, (double*)v, (double*)w, un, vn, wn);
}
