// Sub-grid-scale eddy viscosity.
//   cmpt_sgs     src/sgs.f90:21-386   ('none', 'smag' 69-152 with van Driest damping, 'dsmag' 153-380)
//   strain_rate  src/sgs.f90:1019-1110   filter3d 616-680   extrapolate 682-767   cmpt_alph2 769-822
//   interpolate  850-870                 ave1d_channel 433-538 (the hard-wired `_CHANNEL` average, sgs.f90:8)
// The dynamic procedure keeps the reference's operation order (Appendix B of SURVEY.md) but works on
// batches: one halo exchange + ghost fill for the 7 (then 3) cell-centred arrays, one launch for the six
// products, one for the six extrapolations and one for the six 27-point filters, the Germano contraction
// and the x-y plane sums fused in one kernel, and the plane average consumed on the device.
#include <cstdlib>

#include "common.cuh"
#include "reduce.cuh"
#include "tile.cuh"

#define BX 64
#define BY 4
#define CSMAG 0.11
#define BIG 1.7976931348623157e308

struct Ptr6 { double* p[6]; };
struct CPtr6 { const double* p[6]; };

static inline int pick_kc(int ni, int nj, int nk) {
  return pick_chunk((long)cdiv(ni, BX) * cdiv(nj, BY), nk, 148 * 5, 8, 2);
}

// ---- Smagorinsky + van Driest (sgs.f90:98-152), fused into the strain-rate kernel ---------------------------------
struct SmagArgs {
  double is_wall[6];
  double dl0, dl1, l2, dxi, dyi, visc;
  int any_wall;
  const double* zc; const double* dzci0;      // zc(0:n3+1), dzci(0:n3+1)
  const double* delk;                         // (dl(1)*dl(2)*dzf(k))**(1/3), precomputed per k
  const double *u, *v, *w;                    // the UN-extrapolated velocity (wall shear, sgs.f90:117-143)
  // in-place extrapolation (wall-model faces in z only): the original ghost planes k = 0 / n3+1 of u and v, saved aside
  // (nullptr: read the arrays)
  const double *ug0, *vg0, *ug1, *vg1;
};

// del(k) = (dl(1)*dl(2)*dzf(k))**(1/3)   (sgs.f90:149)
__global__ void smag_del_k(int n3, double dl0, double dl1, const double* __restrict__ dzf, double* __restrict__ delk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= n3 + 1) delk[k] = pow(dl0 * dl1 * dzf[k], 1. / 3.);
}

// sqrt(tau_w) of the z walls (sgs.f90:135-142) depends on (i,j) only: a thread evaluates it once for its column and
// reuses it at every level (same operations, same bits as evaluating it per cell)
__device__ __forceinline__ double wall_sqrt_tau_z(const Dims& d, const SmagArgs& A, int i, int j, int top) {
  const double* __restrict__ u = A.u; const double* __restrict__ v = A.v;
  const int ka = top ? d.n3 : 1, kb = top ? d.n3 + 1 : 0;
  const double* ug = top ? A.ug1 : A.ug0; const double* vg = top ? A.vg1 : A.vg0;     // ghost plane kb: saved copy or the array itself
  const long o = i + d.s1 * (long)j;
  const double* ub = ug ? ug + o : u + d.idx(i, j, kb); const double* vb = vg ? vg + o : v + d.idx(i, j, kb);
  const double t1 = u[d.idx(i, j, ka)] - ub[0] + u[d.idx(i - 1, j, ka)] - ub[-1];
  const double t2 = v[d.idx(i, j, ka)] - vb[0] + v[d.idx(i, j - 1, ka)] - vb[-d.s1];
  const double tauw_s = sqrt(t1 * t1 + t2 * t2) * A.dzci0[top ? d.n3 : 0];
  return sqrt(0.5 * A.visc * tauw_s);
}

__device__ __forceinline__ double van_driest(const Dims& d, const SmagArgs& A, int i, int j, int k, double sq_bot, double sq_top) {
  const double* __restrict__ u = A.u; const double* __restrict__ v = A.v; const double* __restrict__ w = A.w;
  const int n1 = d.n1, n2 = d.n2;
  double dw[6];
  dw[0] = A.dl0 * (i - 0.5); dw[1] = A.dl0 * (n1 - i + 0.5);
  dw[2] = A.dl1 * (j - 0.5); dw[3] = A.dl1 * (n2 - j + 0.5);
  dw[4] = A.zc[k]; dw[5] = A.l2 - A.zc[k];
  int loc = 0;
  double dw_min = BIG;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    dw[q] = dw[q] * A.is_wall[q] + BIG * (1. - A.is_wall[q]);
    if (q == 0 || dw[q] < dw_min) { dw_min = dw[q]; loc = q; }     // minloc: first minimum
  }
  double sq;
  if (loc == 4) sq = sq_bot;
  else if (loc == 5) sq = sq_top;
  else {
    double t1, t2, tauw_s;
#define U(ii, jj, kk) u[d.idx(ii, jj, kk)]
#define V(ii, jj, kk) v[d.idx(ii, jj, kk)]
#define W(ii, jj, kk) w[d.idx(ii, jj, kk)]
    if (loc == 0) {
      t1 = V(1, j, k) - V(0, j, k) + V(1, j - 1, k) - V(0, j - 1, k);
      t2 = W(1, j, k) - W(0, j, k) + W(1, j, k - 1) - W(0, j, k - 1);
      tauw_s = sqrt(t1 * t1 + t2 * t2) * A.dxi;
    } else if (loc == 1) {
      t1 = V(n1, j, k) - V(n1 + 1, j, k) + V(n1, j - 1, k) - V(n1 + 1, j - 1, k);
      t2 = W(n1, j, k) - W(n1 + 1, j, k) + W(n1, j, k - 1) - W(n1 + 1, j, k - 1);
      tauw_s = sqrt(t1 * t1 + t2 * t2) * A.dxi;
    } else if (loc == 2) {
      t1 = U(i, 1, k) - U(i, 0, k) + U(i - 1, 1, k) - U(i - 1, 0, k);
      t2 = W(i, 1, k) - W(i, 0, k) + W(i, 1, k - 1) - W(i, 0, k - 1);
      tauw_s = sqrt(t1 * t1 + t2 * t2) * A.dyi;
    } else {
      t1 = U(i, n2, k) - U(i, n2 + 1, k) + U(i - 1, n2, k) - U(i - 1, n2 + 1, k);
      t2 = W(i, n2, k) - W(i, n2 + 1, k) + W(i, n2, k - 1) - W(i, n2 + 1, k - 1);
      tauw_s = sqrt(t1 * t1 + t2 * t2) * A.dyi;
    }
#undef U
#undef V
#undef W
    sq = sqrt(0.5 * A.visc * tauw_s);
  }
  const double dw_plus = dw_min * sq * (1. / A.visc);
  return 1. - exp(-dw_plus / 25.);
}

// ---- strain rate (sgs.f90:1019-1110) -----------------------------------------------------------------------------
// Same skeleton as mom_k: a TX x TY tile marches in k, the u,v,w planes k and k+1 sit in shared memory (cp.async ring,
// one barrier per plane).  The four k-1/2 terms of s13 and of s23 at level k are the k+1/2 terms of level k-1 (same
// expression, same bits): they are carried in registers, so plane k-1 is never read.  The eight-term sums keep the
// reference's order.
template <int SIJ, int SMAG, bool V16>
__global__ void __launch_bounds__(TX* TY, 3) strain_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci,
                                                    const double* __restrict__ dzfi, const double* __restrict__ u,
                                                    const double* __restrict__ v, const double* __restrict__ w,
                                                    double* __restrict__ s0, Ptr6 sij, double* __restrict__ s0copy, int kc, SmagArgs A,
                                                    double* __restrict__ visct) {
  extern __shared__ __align__(16) double smem[];   // [4 slots][3 fields][PLANE]
  const int i0 = blockIdx.x * TX + 1, j0 = blockIdx.y * TY + 1;
  const int i = i0 + threadIdx.x, j = j0 + threadIdx.y;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  Stager<3, V16> st(d, i0, j0, u, v, w, nullptr, smem, k0 - 1, k1 + 1);
  st.template issue<0>();
  st.template issue<1>();
  st.template issue<2>();
  st.template issue<3>();
  tile_wait_1();
  __syncthreads();
  const bool active = i <= d.n1 && j <= d.n2;
  const double* const sm = smem + (threadIdx.x + 1) + PX * (threadIdx.y + 1);     // my cell in field 0 of slot 0
  constexpr int SL = 3 * PLANE;
  // k+1/2 terms of level k0-1 (planes k0-1, k0 = slots 0, 1)
  double a_uz = 0., a_wx = 0., a_uzm = 0., a_wxm = 0., b_vz = 0., b_wy = 0., b_vzm = 0., b_wym = 0., w_ccm = 0.;
  if (active) {
    const double* uc = sm; const double* vc = sm + PLANE; const double* wc = sm + 2 * PLANE;
    const double* up = uc + SL; const double* vp = vc + SL;
    const double dz = dzci[k0 - 1];
    const double w_ccc = wc[0];
    a_uz = (up[0] - uc[0]) * dz;          a_wx = (wc[1] - w_ccc) * dxi;
    a_uzm = (up[-1] - uc[-1]) * dz;       a_wxm = (w_ccc - wc[-1]) * dxi;
    b_vz = (vp[0] - vc[0]) * dz;          b_wy = (wc[PX] - w_ccc) * dyi;
    b_vzm = (vp[-PX] - vc[-PX]) * dz;     b_wym = (w_ccc - wc[-PX]) * dyi;
    w_ccm = w_ccc;
  }
  __syncthreads();
  long o = d.idx(i, j, k0);
  int k = k0;
  double sq_bot = 0., sq_top = 0.;
  if (SMAG && A.any_wall && active) {
    if (A.is_wall[4] != 0.) sq_bot = wall_sqrt_tau_z(d, A, i, j, 0);
    if (A.is_wall[5] != 0.) sq_top = wall_sqrt_tau_z(d, A, i, j, 1);
  }
  // one level: planes k, k+1 in slots SC, SP; plane k+3 goes into slot SN (which held plane k-1)
  auto step = [&](auto sc_, auto sp_, auto sn_) {
    constexpr int SC = decltype(sc_)::v, SP = decltype(sp_)::v, SN = decltype(sn_)::v;
    st.template issue<SN>();
    {   // threads outside the array compute on whatever their tile cells hold and store nothing
      const double* uc = sm + SC * SL; const double* up = sm + SP * SL;
      const double* vc = uc + PLANE; const double* vp = up + PLANE;
      const double* wc = uc + 2 * PLANE;
      const double u_mmc = uc[-1 - PX], u_cmc = uc[-PX], u_mcc = uc[-1], u_ccc = uc[0], u_mpc = uc[-1 + PX], u_cpc = uc[PX];
      const double u_mcp = up[-1], u_ccp = up[0];
      const double v_mmc = vc[-1 - PX], v_cmc = vc[-PX], v_pmc = vc[1 - PX], v_mcc = vc[-1], v_ccc = vc[0], v_pcc = vc[1];
      const double v_cmp = vp[-PX], v_ccp = vp[0];
      const double w_cmc = wc[-PX], w_mcc = wc[-1], w_ccc = wc[0], w_pcc = wc[1], w_cpc = wc[PX];
      const double dzci_k = dzci[k];
      const double s11 = (u_ccc - u_mcc) * dxi;
      const double s22 = (v_ccc - v_cmc) * dyi;
      const double s33 = (w_ccc - w_ccm) * dzfi[k];
      const double n_uz = (u_ccp - u_ccc) * dzci_k, n_wx = (w_pcc - w_ccc) * dxi, n_uzm = (u_mcp - u_mcc) * dzci_k, n_wxm = (w_ccc - w_mcc) * dxi;
      const double n_vz = (v_ccp - v_ccc) * dzci_k, n_wy = (w_cpc - w_ccc) * dyi, n_vzm = (v_cmp - v_cmc) * dzci_k, n_wym = (w_ccc - w_cmc) * dyi;
      const double s12 = .125 * ((u_cpc - u_ccc) * dyi + (v_pcc - v_ccc) * dxi + (u_ccc - u_cmc) * dyi + (v_pmc - v_cmc) * dxi +
                                 (u_mpc - u_mcc) * dyi + (v_ccc - v_mcc) * dxi + (u_mcc - u_mmc) * dyi + (v_cmc - v_mmc) * dxi);
      const double s13 = .125 * (n_uz + n_wx + a_uz + a_wx + n_uzm + n_wxm + a_uzm + a_wxm);
      const double s23 = .125 * (n_vz + n_wy + b_vz + b_wy + n_vzm + n_wym + b_vzm + b_wym);
      const double s = sqrt(2. * (s11 * s11 + s22 * s22 + s33 * s33 + 2. * (s12 * s12 + s13 * s13 + s23 * s23)));
      if (SMAG) {                                          // visct = (c_smag*del*fd)**2*s0   (sgs.f90:150)
        const double fd = (A.any_wall && active) ? van_driest(d, A, i, j, k, sq_bot, sq_top) : 1.;
        const double t = CSMAG * A.delk[k] * fd;
        if (active) visct[o] = t * t * s;
      } else if (active) s0[o] = s;
      if (SIJ && active) {
        sij.p[0][o] = s11; sij.p[1][o] = s22; sij.p[2][o] = s33; sij.p[3][o] = s12; sij.p[4][o] = s13; sij.p[5][o] = s23;
        if (s0copy) s0copy[o] = s;
      }
      a_uz = n_uz; a_wx = n_wx; a_uzm = n_uzm; a_wxm = n_wxm;
      b_vz = n_vz; b_wy = n_wy; b_vzm = n_vzm; b_wym = n_wym;
      w_ccm = w_ccc;
    }
    tile_wait_1();                   // plane k+2 has landed (k+3 may still be in flight)
    __syncthreads();                 // ... for everyone, and everyone is done reading plane k
    o += d.s2;
    return ++k <= k1;
  };
  while (step(Slot<1>{}, Slot<2>{}, Slot<0>{}) && step(Slot<2>{}, Slot<3>{}, Slot<1>{}) && step(Slot<3>{}, Slot<0>{}, Slot<2>{}) &&
         step(Slot<0>{}, Slot<1>{}, Slot<3>{})) {}
}


// Two rows per thread (the kernel of cales_cmpt_sgs('smag') and of the s0-only strain rate): a CTA of 32 x 4 threads owns the
// same 32 x 8 tile, thread (tx, ty) the cells (i, j) and (i, j+1), j = j0 + 2 ty.  Every y-difference and x-difference that
// the two cells share (half of the s12 terms, half of the s23 terms, their k-1/2 copies) is computed once and the tile is
// read with 16 instead of 24 shared-memory loads per cell -- same expressions, same summation order, hence the same bits as
// strain_k; two independent cells per thread double the instruction-level parallelism (ncu r2o: strain_k is bound by
// dependent-issue waits at 6 warps per scheduler, neither the fp64 pipe (45 %) nor shared memory (73 %) is saturated).
#ifndef STRAIN2_MINB
#define STRAIN2_MINB 4
#endif
template <int SMAG, bool V16>
__global__ void __launch_bounds__(TX* TY / 2, STRAIN2_MINB) strain2_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci,
                                                                     const double* __restrict__ dzfi, const double* __restrict__ u,
                                                                     const double* __restrict__ v, const double* __restrict__ w,
                                                                     double* __restrict__ s0, int kc, SmagArgs A, double* __restrict__ visct) {
  extern __shared__ __align__(16) double smem[];   // [4 slots][3 fields][PLANE]
  const int i0 = blockIdx.x * TX + 1, j0 = blockIdx.y * TY + 1;
  const int i = i0 + threadIdx.x, j = j0 + 2 * threadIdx.y;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  Stager<3, V16, TX * TY / 2> st(d, i0, j0, u, v, w, nullptr, smem, k0 - 1, k1 + 1);
  st.template issue<0>();
  st.template issue<1>();
  st.template issue<2>();
  st.template issue<3>();
  tile_wait_1();
  __syncthreads();
  const bool actA = i <= d.n1 && j <= d.n2, actB = i <= d.n1 && j + 1 <= d.n2;
  const double* const sm = smem + (threadIdx.x + 1) + PX * (2 * threadIdx.y + 1);     // cell A in field 0 of slot 0; cell B = +PX
  constexpr int SL = 3 * PLANE;
  // k+1/2 terms of level k0-1 (planes k0-1, k0 = slots 0, 1).  Row-indexed: [0] = row j-1, [1] = row j, [2] = row j+1
  double a_uz[2], a_wx[2], a_uzm[2], a_wxm[2], w_ccm[2], b_vz[3], b_wy[3];
  {
    const double* uc = sm; const double* vc = sm + PLANE; const double* wc = sm + 2 * PLANE;
    const double* up = uc + SL; const double* vp = vc + SL;
    const double dz = dzci[k0 - 1];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double w_ccc = wc[r * PX];
      a_uz[r] = (up[r * PX] - uc[r * PX]) * dz;           a_wx[r] = (wc[r * PX + 1] - w_ccc) * dxi;
      a_uzm[r] = (up[r * PX - 1] - uc[r * PX - 1]) * dz;  a_wxm[r] = (w_ccc - wc[r * PX - 1]) * dxi;
      w_ccm[r] = w_ccc;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      b_vz[r] = (vp[(r - 1) * PX] - vc[(r - 1) * PX]) * dz;
      b_wy[r] = (wc[r * PX] - wc[(r - 1) * PX]) * dyi;
    }
  }
  __syncthreads();
  long o = d.idx(i, j, k0);
  int k = k0;
  double sq_bot[2] = {0., 0.}, sq_top[2] = {0., 0.};
  constexpr bool WALLS = false;      // strain2_k is launched for wall-free Smagorinsky only (see cales_cmpt_sgs): no van Driest code
  if (SMAG && WALLS && A.any_wall) {
    if (actA) { if (A.is_wall[4] != 0.) sq_bot[0] = wall_sqrt_tau_z(d, A, i, j, 0); if (A.is_wall[5] != 0.) sq_top[0] = wall_sqrt_tau_z(d, A, i, j, 1); }
    if (actB) { if (A.is_wall[4] != 0.) sq_bot[1] = wall_sqrt_tau_z(d, A, i, j + 1, 0); if (A.is_wall[5] != 0.) sq_top[1] = wall_sqrt_tau_z(d, A, i, j + 1, 1); }
  }
  auto step = [&](auto sc_, auto sp_, auto sn_) {
    constexpr int SC = decltype(sc_)::v, SP = decltype(sp_)::v, SN = decltype(sn_)::v;
    st.template issue<SN>();
    {   // threads outside the array compute on whatever their tile cells hold and store nothing
      const double* uc = sm + SC * SL; const double* up = sm + SP * SL;
      const double* vc = uc + PLANE; const double* vp = up + PLANE;
      const double* wc = uc + 2 * PLANE;
      // u: columns i-1 (m), i (c); rows j-1 .. j+2
      double um[4], ucn[4], vm[3], vcn[3], vpn[3], wcn[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { um[r] = uc[(r - 1) * PX - 1]; ucn[r] = uc[(r - 1) * PX]; wcn[r] = wc[(r - 1) * PX]; }
#pragma unroll
      for (int r = 0; r < 3; ++r) { vm[r] = vc[(r - 1) * PX - 1]; vcn[r] = vc[(r - 1) * PX]; vpn[r] = vc[(r - 1) * PX + 1]; }
      const double u_mcp[2] = {up[-1], up[PX - 1]}, u_ccp[2] = {up[0], up[PX]};
      const double v_p[3] = {vp[-PX], vp[0], vp[PX]};
      const double w_m[2] = {wc[-1], wc[PX - 1]}, w_p[2] = {wc[1], wc[PX + 1]};
      const double dzci_k = dzci[k], dzfi_k = dzfi[k];
      // shared differences: UY[col][r] = (u(col, row r+1) - u(col, row r)) dyi, VX[col][r] = (v(col+1, r) - v(col, r)) dxi, rows j-1, j, j+1
      double uyc[3], uym[3], vxc[3], vxm[3], n_vz[3], n_wy[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        uyc[r] = (ucn[r + 1] - ucn[r]) * dyi; uym[r] = (um[r + 1] - um[r]) * dyi;
        vxc[r] = (vpn[r] - vcn[r]) * dxi;     vxm[r] = (vcn[r] - vm[r]) * dxi;
        n_vz[r] = (v_p[r] - vcn[r]) * dzci_k; n_wy[r] = (wcn[r + 1] - wcn[r]) * dyi;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {                       // cell A (c = 0): row index 1; cell B: row index 2
        const int r = c + 1;
        const double u_ccc = ucn[r], u_mcc = um[r], v_ccc = vcn[r], v_cmc = vcn[r - 1], w_ccc = wcn[r];
        const double s11 = (u_ccc - u_mcc) * dxi;
        const double s22 = (v_ccc - v_cmc) * dyi;
        const double s33 = (w_ccc - w_ccm[c]) * dzfi_k;
        const double n_uz = (u_ccp[c] - u_ccc) * dzci_k, n_wx = (w_p[c] - w_ccc) * dxi, n_uzm = (u_mcp[c] - u_mcc) * dzci_k, n_wxm = (w_ccc - w_m[c]) * dxi;
        const double s12 = .125 * (uyc[r] + vxc[r] + uyc[r - 1] + vxc[r - 1] + uym[r] + vxm[r] + uym[r - 1] + vxm[r - 1]);
        const double s13 = .125 * (n_uz + n_wx + a_uz[c] + a_wx[c] + n_uzm + n_wxm + a_uzm[c] + a_wxm[c]);
        const double s23 = .125 * (n_vz[r] + n_wy[r] + b_vz[r] + b_wy[r] + n_vz[r - 1] + n_wy[r - 1] + b_vz[r - 1] + b_wy[r - 1]);
        const double s = sqrt(2. * (s11 * s11 + s22 * s22 + s33 * s33 + 2. * (s12 * s12 + s13 * s13 + s23 * s23)));
        const bool act = c ? actB : actA;
        const long oc = o + c * d.s1;
        if (SMAG) {                                        // visct = (c_smag*del*fd)**2*s0   (sgs.f90:150)
          const double fd = (WALLS && A.any_wall && act) ? van_driest(d, A, i, j + c, k, sq_bot[c], sq_top[c]) : 1.;
          const double t = CSMAG * A.delk[k] * fd;
          if (act) visct[oc] = t * t * s;
        } else if (act) s0[oc] = s;
        a_uz[c] = n_uz; a_wx[c] = n_wx; a_uzm[c] = n_uzm; a_wxm[c] = n_wxm;
        w_ccm[c] = w_ccc;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) { b_vz[r] = n_vz[r]; b_wy[r] = n_wy[r]; }
    }
    tile_wait_1();                   // plane k+2 has landed (k+3 may still be in flight)
    __syncthreads();                 // ... for everyone, and everyone is done reading plane k
    o += d.s2;
    return ++k <= k1;
  };
  while (step(Slot<1>{}, Slot<2>{}, Slot<0>{}) && step(Slot<2>{}, Slot<3>{}, Slot<1>{}) && step(Slot<3>{}, Slot<0>{}, Slot<2>{}) &&
         step(Slot<0>{}, Slot<1>{}, Slot<3>{})) {}
}

#define STRAIN_SMEM (TSLOTS * 3 * PLANE * sizeof(double))
// two-rows-per-thread kernel (strain2_k) for the launches without the six s_ij outputs; CALES_STRAIN2=0 keeps strain_k
static inline bool strain2_on() { static const int v = getenv("CALES_STRAIN2") ? atoi(getenv("CALES_STRAIN2")) : 1; return v != 0; }

static inline dim3 strain_grid(const int n[3], int& kc, int resident_per_sm = 3) {
  kc = pick_chunk((long)cdiv(n[0], TX) * cdiv(n[1], TY), n[2], 148 * resident_per_sm, 12, 2);
  return dim3(cdiv(n[0], TX), cdiv(n[1], TY), cdiv(n[2], kc));
}

static int strain_launch(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, const double* dzfi, const double* u,
                         const double* v, const double* w, double* s0, double* const* sij, double* s0copy) {
  Dims d(n);
  int kc;
  const dim3 g = strain_grid(n, kc), b(TX, TY);
  Ptr6 P;
  for (int m = 0; m < 6; ++m) P.p[m] = sij ? sij[m] : nullptr;
  SmagArgs none{};
  const bool v16 = tile_v16(n[0], u, v, w);
#define STRAIN_GO(SIJ_, V_, S0C_) strain_k<SIJ_, 0, V_><<<g, b, STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, u, v, w, s0, P, S0C_, kc, none, nullptr)
  if (sij) { if (v16) STRAIN_GO(1, true, s0copy); else STRAIN_GO(1, false, s0copy); }
  else if (strain2_on()) {
    int kc2;
    const dim3 g2 = strain_grid(n, kc2, STRAIN2_MINB);
    if (v16) strain2_k<0, true><<<g2, dim3(TX, TY / 2), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, u, v, w, s0, kc2, none, nullptr);
    else strain2_k<0, false><<<g2, dim3(TX, TY / 2), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, u, v, w, s0, kc2, none, nullptr);
  }
  else { if (v16) STRAIN_GO(0, true, nullptr); else STRAIN_GO(0, false, nullptr); }
#undef STRAIN_GO
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

extern "C" int cales_strain_rate(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, const double* dzfi,
                                 const double* u, const double* v, const double* w, double* s0, double* sij) {
  CHECK_CTX(ctx);
  Dims d(n);
  double* ps[6];
  for (int m = 0; m < 6; ++m) ps[m] = sij ? sij + m * d.size() : nullptr;
  return strain_launch(ctx, n, dli, dzci, dzfi, u, v, w, s0, sij ? ps : nullptr, nullptr);
}

// ---- filter3d (sgs.f90:616-680), NF arrays per launch (blockIdx.z / nkb selects the array) ---------------------------
// z-march with the 3 x 3 x 3 neighbourhood held in registers: 9 loads per cell instead of 27; the sum keeps the reference's
// association order (same bits).  F2D: filter2d (sgs.f90:824-848, the -D_FILTER_2D variant): 9 points of the plane, /16.
template <bool F2D>
__global__ void __launch_bounds__(BX* BY) filter3d_k(Dims d, CPtr6 in, Ptr6 out, int kc, int nkb) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int f = blockIdx.z / nkb, kb = blockIdx.z - f * nkb;
  const int k0 = kb * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  const double* __restrict__ p = in.p[f];
  double* __restrict__ pf = out.p[f];
  const long s1 = d.s1, s2 = d.s2;
  long c = d.idx(i, j, k0);
  if (F2D) {
    for (int k = k0; k <= k1; ++k, c += s2) {
#define Q(di, dj) p[c + (di) + (dj) * s1]
      pf[c] = (4. * Q(0, 0) + 2. * (Q(-1, 0) + Q(0, -1) + Q(1, 0) + Q(0, 1)) + 1. * (Q(-1, -1) + Q(1, -1) + Q(-1, 1) + Q(1, 1))) / 16.;
#undef Q
    }
    return;
  }
  double m[3][3], q0[3][3], pl[3][3];      // planes k-1, k, k+1: [dj+1][di+1]
#pragma unroll
  for (int dj = 0; dj < 3; ++dj)
#pragma unroll
    for (int di = 0; di < 3; ++di) { m[dj][di] = p[c - s2 + (di - 1) + (dj - 1) * s1]; q0[dj][di] = p[c + (di - 1) + (dj - 1) * s1]; }
  for (int k = k0; k <= k1; ++k, c += s2) {
#pragma unroll
    for (int dj = 0; dj < 3; ++dj)
#pragma unroll
      for (int di = 0; di < 3; ++di) pl[dj][di] = p[c + s2 + (di - 1) + (dj - 1) * s1];
#define Q(di, dj, dk) ((dk) < 0 ? m[(dj) + 1][(di) + 1] : (dk) == 0 ? q0[(dj) + 1][(di) + 1] : pl[(dj) + 1][(di) + 1])
    pf[c] = (8. * Q(0, 0, 0) +
             4. * (Q(-1, 0, 0) + Q(0, -1, 0) + Q(0, 0, -1) + Q(1, 0, 0) + Q(0, 1, 0) + Q(0, 0, 1)) +
             2. * (Q(0, -1, -1) + Q(-1, 0, -1) + Q(-1, -1, 0) + Q(0, 1, -1) + Q(1, 0, -1) + Q(1, -1, 0) +
                   Q(0, -1, 1) + Q(-1, 0, 1) + Q(-1, 1, 0) + Q(0, 1, 1) + Q(1, 0, 1) + Q(1, 1, 0)) +
             1. * (Q(-1, -1, -1) + Q(1, -1, -1) + Q(-1, 1, -1) + Q(1, 1, -1) + Q(-1, -1, 1) + Q(1, -1, 1) + Q(-1, 1, 1) + Q(1, 1, 1))) / 64.;
#undef Q
#pragma unroll
    for (int dj = 0; dj < 3; ++dj)
#pragma unroll
      for (int di = 0; di < 3; ++di) { m[dj][di] = q0[dj][di]; q0[dj][di] = pl[dj][di]; }
  }
}

static int filter_launch(cales_ctx* ctx, const int n[3], double* const* in, double* const* out, int nf, bool f2d = false) {
  Dims d(n);
  const int kc = pick_kc(n[0], n[1], n[2]);
  const int nkb = cdiv(n[2], kc);
  CPtr6 I; Ptr6 O;
  for (int m = 0; m < nf; ++m) { I.p[m] = in[m]; O.p[m] = out[m]; }
  if (f2d) filter3d_k<true><<<dim3(cdiv(n[0], BX), cdiv(n[1], BY), nkb * nf), dim3(BX, BY), 0, ctx->stream>>>(d, I, O, kc, nkb);
  else filter3d_k<false><<<dim3(cdiv(n[0], BX), cdiv(n[1], BY), nkb * nf), dim3(BX, BY), 0, ctx->stream>>>(d, I, O, kc, nkb);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

extern "C" int cales_filter3d(cales_ctx* ctx, const int n[3], const double* p, double* pf) {
  CHECK_CTX(ctx);
  double* in[1] = {(double*)p};
  double* out[1] = {pf};
  return filter_launch(ctx, n, in, out, 1);
}

// ---- extrapolate (sgs.f90:682-767): one launch per direction for up to 6 arrays ---------------------------------------
struct ExTask { double* p; int lo, hi; };   // lo/hi: extrapolate the lower/upper ghost plane of this direction
struct ExBatch { ExTask t[6]; int nt, idir; const double* dzci; int use_factor; };

__global__ void __launch_bounds__(256) extrap_k(Dims d, ExBatch b) {
  const ExTask& t = b.t[blockIdx.z];
  const int idir = b.idir;
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const long sn = idir == 0 ? 1 : idir == 1 ? d.s1 : d.s2;
  const long base = idir == 0 ? d.idx(0, a, c) : idir == 1 ? d.idx(a, 0, c) : d.idx(a, c, 0);
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  double* p = t.p;
  double f0 = 1., f1 = 1.;
  if (idir == 2 && b.use_factor) {                       // lwm variant: factor0 = dzc(0)*dzci(1), factor1 = dzc(n3)*dzci(n3-1)
    f0 = (1. / b.dzci[0]) * b.dzci[1];
    f1 = (1. / b.dzci[n]) * b.dzci[n - 1];
  }
#define P(q) p[base + sn * (long)(q)]
  if (idir < 2) {
    if (t.lo) P(0) = 2. * P(1) - P(2);
    if (t.hi) P(n + 1) = 2. * P(n) - P(n - 1);
  } else {
    if (t.lo) P(0) = (1. + f0) * P(1) - f0 * P(2);
    if (t.hi) P(n + 1) = (1. + f1) * P(n) - f1 * P(n - 1);
  }
#undef P
}

// is_done(ib,idir) per array: mode 0 = cbc variant (D walls), mode 1 = lwm variant (wall-model faces)
static int extrap_launch(cales_ctx* ctx, const int n[3], const int is_bound[6], const double* dzci, double* const* ps, const int* ifaces,
                         int np, int mode, const char* cbc, const int* lwm) {
  Dims d(n);
  for (int idir = 0; idir < 3; ++idir) {
    ExBatch b; b.nt = 0; b.idir = idir; b.dzci = dzci; b.use_factor = mode == 1;
    for (int f = 0; f < np; ++f) {
      int done[2];
      for (int ib = 0; ib < 2; ++ib) {
        const bool wall = mode == 0 ? cbc[ib + 2 * idir + 6 * idir] == 'D' : lwm[tb(ib, idir)] != 0;
        done[ib] = is_bound[tb(ib, idir)] && wall && ifaces[f] != idir + 1;
      }
      if (done[0] || done[1]) { b.t[b.nt].p = ps[f]; b.t[b.nt].lo = done[0]; b.t[b.nt].hi = done[1]; b.nt++; }
    }
    if (b.nt == 0) continue;
    const int m1 = idir == 0 ? n[1] + 2 : n[0] + 2, m2 = idir == 2 ? n[1] + 2 : n[2] + 2;
    extrap_k<<<dim3(cdiv(m1, 64), cdiv(m2, 4), b.nt), dim3(64, 4), 0, ctx->stream>>>(d, b);
    KERNEL_CHECK(ctx);
  }
  return CALES_OK;
}

// ---- dsmag building blocks ------------------------------------------------------------------------------------------------
__global__ void copy3_k(long n, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                        double* __restrict__ da, double* __restrict__ db, double* __restrict__ dc) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { da[i] = a[i]; db[i] = b[i]; dc[i] = c[i]; }
}

// wk(m) = s0*sij(m) over the full index range (sgs.f90:198-210)
__global__ void prod_s_k(long n, const double* __restrict__ s0, CPtr6 sij, Ptr6 wk) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double s = s0[i];
#pragma unroll
    for (int m = 0; m < 6; ++m) wk.p[m][i] = s * sij.p[m][i];
  }
}

// wk = (uc*uc, vc*vc, wc*wc, uc*vc, uc*wc, vc*wc) over the full range (sgs.f90:283-295)
__global__ void prod_u_k(long n, const double* __restrict__ uc, const double* __restrict__ vc, const double* __restrict__ wc, Ptr6 wk) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double a = uc[i], b = vc[i], c = wc[i];
    wk.p[0][i] = a * a; wk.p[1][i] = b * b; wk.p[2][i] = c * c; wk.p[3][i] = a * b; wk.p[4][i] = a * c; wk.p[5][i] = b * c;
  }
}

struct Alph { int wall[6]; int all; };   // is_bound && cbc(ib,idir,idir)=='D' (cmpt_alph2, sgs.f90:783-816); all: _FILTER_2D (817-821)

// mij = 2*(mij - alph2*s0*sij) interior (sgs.f90:262-272) fused with interpolate (sgs.f90:850-870)
__global__ void __launch_bounds__(BX* BY) mij_interp_k(Dims d, Alph al, const double* __restrict__ s0, CPtr6 sij, Ptr6 mij,
                                                        const double* __restrict__ u, const double* __restrict__ v,
                                                        const double* __restrict__ w, double* __restrict__ uc,
                                                        double* __restrict__ vc, double* __restrict__ wc, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  for (int k = k0; k <= k1; ++k, c += d.s2) {
    double alph2 = 4.00;
    if (al.all || (al.wall[0] && i == 1) || (al.wall[1] && i == d.n1) || (al.wall[2] && j == 1) || (al.wall[3] && j == d.n2) ||
        (al.wall[4] && k == 1) || (al.wall[5] && k == d.n3)) alph2 = 2.52;
    const double s = s0[c];
#pragma unroll
    for (int m = 0; m < 6; ++m) mij.p[m][c] = 2. * (mij.p[m][c] - alph2 * s * sij.p[m][c]);
    uc[c] = 0.5 * (u[c] + u[c - 1]);
    vc[c] = 0.5 * (v[c] + v[c - d.s1]);
    wc[c] = 0.5 * (w[c] + w[c - d.s2]);
  }
}

// Germano contraction (sgs.f90:328-358) fused with the x-y plane sums of ave1d_channel (sgs.f90:455-474):
// one CTA per (k, x-y tile); per-plane partials are folded in a fixed order afterwards.
// cml/cmm (non-null): the per-cell contractions are stored as well (the _DUCT and _CAVITY averaging variants need them).
__global__ void __launch_bounds__(BX* BY) contract_k(Dims d, CPtr6 mij, CPtr6 lij, const double* __restrict__ uf,
                                                      const double* __restrict__ vf, const double* __restrict__ wf,
                                                      double* __restrict__ part, int ntile, double* __restrict__ cml, double* __restrict__ cmm) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1, k = blockIdx.z + 1;
  double ml = 0., mm = 0.;
  if (i <= d.n1 && j <= d.n2) {
    const long c = d.idx(i, j, k);
    double M[6], L[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) { M[m] = mij.p[m][c]; L[m] = lij.p[m][c]; }
    const double a = uf[c], b = vf[c], e = wf[c];
    L[0] = L[0] - a * a; L[1] = L[1] - b * b; L[2] = L[2] - e * e;
    L[3] = L[3] - a * b; L[4] = L[4] - a * e; L[5] = L[5] - b * e;
    ml = M[0] * L[0] + M[1] * L[1] + M[2] * L[2] + (M[3] * L[3] + M[4] * L[4] + M[5] * L[5]) * 2.;
    mm = M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + (M[3] * M[3] + M[4] * M[4] + M[5] * M[5]) * 2.;
    if (cml) { cml[c] = ml; cmm[c] = mm; }
  }
  ml = block_sum<BX * BY>(ml);
  mm = block_sum<BX * BY>(mm);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int tile = blockIdx.x + gridDim.x * blockIdx.y;
    part[(2 * (k - 1) + 0) * (long)ntile + tile] = ml;
    part[(2 * (k - 1) + 1) * (long)ntile + tile] = mm;
  }
}

// fold tile partials of plane k into p1d[kg] (kg = global k index), scaled by grid_area_ratio
__global__ void plane_fold_k(const double* __restrict__ part, int ntile, int n3, int lo3, int ng3, double gar, double* __restrict__ p1d) {
  const int q = blockIdx.x;          // q = 2*(k-1)+which
  double v = 0.;
  for (int t = threadIdx.x; t < ntile; t += 256) v = v + part[q * (long)ntile + t];
  v = block_sum<256>(v);
  if (threadIdx.x == 0) {
    const int k = q / 2, which = q % 2;
    p1d[which * ng3 + (lo3 - 1 + k)] = v * gar;
  }
}

// ave0d_dit (sgs.f90:388-431): volume average = sum_k plane average(k) dzf(k)/l(3) over the local levels (then all-reduced)
__global__ void dit_fold_k(const double* __restrict__ p1d, int ng3, int lo3, int n3, const double* __restrict__ dzf, double l3, double* __restrict__ p0d) {
  const int which = blockIdx.x;
  double v = 0.;
  for (int k = threadIdx.x + 1; k <= n3; k += 256) v = v + p1d[which * ng3 + lo3 - 1 + k - 1] * dzf[k] / l3;
  v = block_sum<256>(v);
  if (threadIdx.x == 0) p0d[which] = v;
}

// ave2d_duct (sgs.f90:540-614, streamwise direction x): one warp per (j,k) row sums the per-cell contractions along i
// (x is never decomposed in X-pencils) and scales by dl(1)/l(1)
__global__ void __launch_bounds__(256) duct_rows_k(Dims d, const double* __restrict__ cml, const double* __restrict__ cmm, double gar,
                                                    double* __restrict__ ml2d, double* __restrict__ mm2d) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= d.n2 * d.n3) return;
  const int j = row % d.n2 + 1, k = row / d.n2 + 1;
  const long c0 = d.idx(0, j, k);
  double a = 0., b = 0.;
  for (int i = lane + 1; i <= d.n1; i += 32) { a = a + cml[c0 + i]; b = b + cmm[c0 + i]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a = a + __shfl_down_sync(0xffffffffu, a, o); b = b + __shfl_down_sync(0xffffffffu, b, o); }
  if (lane == 0) { ml2d[row] = a * gar; mm2d[row] = b * gar; }
}

// MODE 0 _DIT (two scalars), 2 _DUCT (per (j,k) row), 3 _CAVITY (no averaging: the cell's own contraction)
template <int MODE>
__global__ void __launch_bounds__(BX* BY) dsmag_final_var_k(Dims d, const double* __restrict__ a0, const double* __restrict__ a1,
                                                            double* __restrict__ visct, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  for (int k = k0; k <= k1; ++k, c += d.s2) {
    double ml, mm;
    if (MODE == 0) { ml = a0[0]; mm = a0[1]; }
    else if (MODE == 2) { const long r = (j - 1) + (long)d.n2 * (k - 1); ml = a0[r]; mm = a1[r]; }
    else { ml = a0[c]; mm = a1[c]; }
    visct[c] = fmax(visct[c] * ml / mm, 0.);
  }
}

// visct = max(visct*<ML>/<MM>, 0) (sgs.f90:372-380)
__global__ void __launch_bounds__(BX* BY) dsmag_final_k(Dims d, const double* __restrict__ p1d, int ng3, int lo3, double* __restrict__ visct, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  for (int k = k0; k <= k1; ++k, c += d.s2) {
    const double vt = visct[c] * p1d[lo3 - 1 + k - 1] / p1d[ng3 + lo3 - 1 + k - 1];
    visct[c] = fmax(vt, 0.);
  }
}

// ---- cmpt_sgs ---------------------------------------------------------------------------------------------------------------
extern "C" int cales_cmpt_sgs(cales_ctx* ctx, const char* sgstype, const int n[3], const int ng[3], const int lo[3],
                              const int hi[3], const char cbcvel[18], const char cbcsgs[6], const cales_bound* bcs,
                              const int nb[6], const int is_bound[6], const int lwm[6], const double l[3], const double dl[3],
                              const double dli[3], const double* zc, const double* zf, const double* dzc, const double* dzf,
                              const double* dzci, const double* dzfi, double visc, double h, const int index_wm[6],
                              const double* u, const double* v, const double* w, const cales_bound* bcuf,
                              const cales_bound* bcvf, const cales_bound* bcwf, const cales_bound* bcu_mag,
                              const cales_bound* bcv_mag, const cales_bound* bcw_mag, double* visct) {
  CHECK_CTX(ctx);
  (void)hi;
  Dims d(n);
  const size_t fb = (size_t)d.size() * sizeof(double);
  const int kc = pick_kc(n[0], n[1], n[2]);
  dim3 g(cdiv(n[0], BX), cdiv(n[1], BY), cdiv(n[2], kc)), b(BX, BY);
  int rc;
  if (!strcmp(sgstype, "none")) {
    if (ctx->sgs_first) { ctx->sgs_first = false; CUDA_TRY(ctx, cudaMemsetAsync(visct, 0, fb, ctx->stream)); }
    return CALES_OK;
  }
  bool any_wm = false;
  for (int q = 0; q < 6; ++q) any_wm |= (is_bound[q] && lwm[q] != 0);
  const int ifaces3[3] = {1, 2, 3}, ifaces0[6] = {0, 0, 0, 0, 0, 0};
  const long nflat = d.size();
  const int gflat = 148 * 8;
  if (!strcmp(sgstype, "smag")) {
    const double *us = u, *vs = v, *ws = w;
    // Wall-model faces in z only and no x/y walls (the channel): instead of copying u,v,w (sgs.f90:84-90, 48 B/cell) the
    // extrapolation (which touches the ghost planes k = 0 / n3+1 of u and v only) is done IN PLACE on the caller's arrays
    // and undone afterwards; the four original planes are saved aside -- the van Driest wall shear reads them there.
    bool inplace = any_wm;
    for (int q6 = 0; q6 < 4; ++q6) inplace = inplace && !(is_bound[q6] && lwm[q6] != 0);
    for (int idir = 0; idir < 2 && inplace; ++idir)
      for (int ib = 0; ib < 2; ++ib) if (is_bound[tb(ib, idir)] && cbcvel[ib + 2 * idir + 6 * idir] == 'D') inplace = false;
    static const bool no_inplace = getenv("CALES_SGS_COPY") != nullptr;
    double* gsave = nullptr;
    const size_t pb_ = (size_t)d.s2 * sizeof(double);
    if (inplace && !no_inplace) {
      gsave = (double*)cales_scratch(ctx, "sgs_gsave", 4 * pb_);
      if (!gsave) return CALES_ERR_NOMEM;
      const double* srcs[2] = {u, v};
      for (int f = 0; f < 2; ++f)
        for (int ib = 0; ib < 2; ++ib)
          if (is_bound[tb(ib, 2)] && lwm[tb(ib, 2)] != 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(gsave + (size_t)(2 * f + ib) * d.s2, srcs[f] + (size_t)d.s2 * (ib ? n[2] + 1 : 0), pb_, cudaMemcpyDeviceToDevice, ctx->stream));
      double* wkp[3] = {(double*)u, (double*)v, (double*)w};
      if ((rc = extrap_launch(ctx, n, is_bound, dzci, wkp, ifaces3, 3, 1, nullptr, lwm))) return rc;
    } else if (any_wm) {                                     // sgs.f90:84-90: copies only matter where extrapolate acts
      double* wk[3];
      static const char* nm[3] = {"sgs_wk0", "sgs_wk1", "sgs_wk2"};
      for (int c = 0; c < 3; ++c) if (!(wk[c] = (double*)cales_scratch(ctx, nm[c], fb))) return CALES_ERR_NOMEM;
      copy3_k<<<gflat, 256, 0, ctx->stream>>>(nflat, u, v, w, wk[0], wk[1], wk[2]);
      KERNEL_CHECK(ctx);
      if ((rc = extrap_launch(ctx, n, is_bound, dzci, wk, ifaces3, 3, 1, nullptr, lwm))) return rc;
      us = wk[0]; vs = wk[1]; ws = wk[2];
    }
    double* delk = (double*)cales_scratch(ctx, "sgs_delk", (size_t)(n[2] + 2) * sizeof(double));
    if (!delk) return CALES_ERR_NOMEM;
    smag_del_k<<<cdiv(n[2] + 2, 128), 128, 0, ctx->stream>>>(n[2], dl[0], dl[1], dzf, delk);
    KERNEL_CHECK(ctx);
    SmagArgs A;
    A.any_wall = 0;
    for (int idir = 0; idir < 3; ++idir)
      for (int ib = 0; ib < 2; ++ib) {
        const bool wall = is_bound[tb(ib, idir)] && cbcvel[ib + 2 * idir + 6 * idir] == 'D';
        A.is_wall[2 * idir + ib] = wall ? 1. : 0.;
        A.any_wall |= wall;
      }
    A.dl0 = dl[0]; A.dl1 = dl[1]; A.l2 = l[2]; A.dxi = dli[0]; A.dyi = dli[1]; A.visc = visc;
    A.zc = zc; A.dzci0 = dzci; A.delk = delk; A.u = u; A.v = v; A.w = w;
    A.ug0 = A.vg0 = A.ug1 = A.vg1 = nullptr;
    if (gsave) {
      if (is_bound[tb(0, 2)] && lwm[tb(0, 2)] != 0) { A.ug0 = gsave; A.vg0 = gsave + 2 * (size_t)d.s2; }
      if (is_bound[tb(1, 2)] && lwm[tb(1, 2)] != 0) { A.ug1 = gsave + (size_t)d.s2; A.vg1 = gsave + 3 * (size_t)d.s2; }
    }
    Ptr6 P{};
    int kcs;
    // two rows per thread where no wall damping is evaluated (TGV / wall-free directions): 0.164 -> 0.147 ms at 256^3; with
    // van Driest damping the second inlined copy costs more registers than the shared loads save (512x256x192 wall-modelled
    // channel: 0.43 -> 0.49 ms), so walls keep strain_k
    const bool two = strain2_on() && !A.any_wall;
    const dim3 gs = strain_grid(n, kcs, two ? STRAIN2_MINB : 3);
    if (two) {
      if (tile_v16(n[0], us, vs, ws)) strain2_k<1, true><<<gs, dim3(TX, TY / 2), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, us, vs, ws, nullptr, kcs, A, visct);
      else strain2_k<1, false><<<gs, dim3(TX, TY / 2), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, us, vs, ws, nullptr, kcs, A, visct);
    } else if (tile_v16(n[0], us, vs, ws))
      strain_k<0, 1, true><<<gs, dim3(TX, TY), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, us, vs, ws, nullptr, P, nullptr, kcs, A, visct);
    else
      strain_k<0, 1, false><<<gs, dim3(TX, TY), STRAIN_SMEM, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, us, vs, ws, nullptr, P, nullptr, kcs, A, visct);
    KERNEL_CHECK(ctx);
    if (gsave) {                                             // undo the in-place extrapolation: the caller's ghost planes return
      double* dsts[2] = {(double*)u, (double*)v};
      for (int f = 0; f < 2; ++f)
        for (int ib = 0; ib < 2; ++ib)
          if (is_bound[tb(ib, 2)] && lwm[tb(ib, 2)] != 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(dsts[f] + (size_t)d.s2 * (ib ? n[2] + 1 : 0), gsave + (size_t)(2 * f + ib) * d.s2, pb_, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return CALES_OK;
  }
  if (strcmp(sgstype, "dsmag")) return cales_fail(ctx, CALES_ERR_INVALID, "unknown SGS model '%s'", sgstype);
  // ---- dynamic Smagorinsky (sgs.f90:153-380) ----
  static const char* nm1[7] = {"sgs_s0", "sgs_uc", "sgs_vc", "sgs_wc", "sgs_uf", "sgs_vf", "sgs_wf"};
  double* a1[7];
  for (int q = 0; q < 7; ++q) if (!(a1[q] = (double*)cales_scratch(ctx, nm1[q], fb, true))) return CALES_ERR_NOMEM;
  double *s0 = a1[0], *uc = a1[1], *vc = a1[2], *wc = a1[3], *uf = a1[4], *vf = a1[5], *wf = a1[6];
  double *wk[6], *sij[6], *mij[6];
  static const char* nwk[6] = {"sgs_wk0", "sgs_wk1", "sgs_wk2", "sgs_wk3", "sgs_wk4", "sgs_wk5"};
  static const char* nsij[6] = {"sgs_sij0", "sgs_sij1", "sgs_sij2", "sgs_sij3", "sgs_sij4", "sgs_sij5"};
  static const char* nmij[6] = {"sgs_mij0", "sgs_mij1", "sgs_mij2", "sgs_mij3", "sgs_mij4", "sgs_mij5"};
  for (int m = 0; m < 6; ++m) {
    wk[m] = (double*)cales_scratch(ctx, nwk[m], fb, true);
    sij[m] = (double*)cales_scratch(ctx, nsij[m], fb, true);
    mij[m] = (double*)cales_scratch(ctx, nmij[m], fb, true);
    if (!wk[m] || !sij[m] || !mij[m]) return CALES_ERR_NOMEM;
  }
  CPtr6 Csij, Cmij; Ptr6 Pwk, Pmij;
  for (int m = 0; m < 6; ++m) { Csij.p[m] = sij[m]; Cmij.p[m] = mij[m]; Pwk.p[m] = wk[m]; Pmij.p[m] = mij[m]; }
  // (2)-(3) strain rate of the (wall-model extrapolated) velocity; visct <- s0      sgs.f90:173-185
  const double *us = u, *vs = v, *ws = w;
  if (any_wm) {
    copy3_k<<<gflat, 256, 0, ctx->stream>>>(nflat, u, v, w, wk[0], wk[1], wk[2]);
    KERNEL_CHECK(ctx);
    if ((rc = extrap_launch(ctx, n, is_bound, dzci, wk, ifaces3, 3, 1, nullptr, lwm))) return rc;
    us = wk[0]; vs = wk[1]; ws = wk[2];
  }
  if ((rc = strain_launch(ctx, n, dli, dzci, dzfi, us, vs, ws, s0, sij, visct))) return rc;
  // (4) ghost cells of s0 and sij                                                        sgs.f90:191-197
  {
    double* ps[7] = {s0, sij[0], sij[1], sij[2], sij[3], sij[4], sij[5]};
    if ((rc = k_boundp_multi(ctx, cbcsgs, n, bcs, nb, is_bound, dl, dzc, ps, 7))) return rc;
  }
  // (5) Mij first part: filter(s0*sij)                                                   sgs.f90:198-223
  const bool f2d = ctx->sgs_filter2d != 0;                   // -D_FILTER_2D: filter2d, no extrapolation (sgs.f90:236-247, 316-327)
  prod_s_k<<<gflat, 256, 0, ctx->stream>>>(nflat, s0, Csij, Pwk);
  KERNEL_CHECK(ctx);
  if (!f2d && (rc = extrap_launch(ctx, n, is_bound, dzci, wk, ifaces0, 6, 0, cbcvel, nullptr))) return rc;
  if ((rc = filter_launch(ctx, n, wk, mij, 6, f2d))) return rc;
  // (6) filtered velocity                                                                sgs.f90:225-235
  if (!f2d) {
    copy3_k<<<gflat, 256, 0, ctx->stream>>>(nflat, u, v, w, wk[0], wk[1], wk[2]);
    KERNEL_CHECK(ctx);
    if ((rc = extrap_launch(ctx, n, is_bound, dzci, wk, ifaces3, 3, 0, cbcvel, nullptr))) return rc;
    double* out[3] = {uf, vf, wf};
    if ((rc = filter_launch(ctx, n, wk, out, 3))) return rc;
  } else {
    double* in3[3] = {(double*)u, (double*)v, (double*)w};
    double* out[3] = {uf, vf, wf};
    if ((rc = filter_launch(ctx, n, in3, out, 3, true))) return rc;
  }
  // (7) BCs on the filtered velocity, strain rate of it                                   sgs.f90:256-261
  if ((rc = cales_bounduvw(ctx, cbcvel, n, bcuf, bcvf, bcwf, bcu_mag, bcv_mag, bcw_mag, nb, is_bound, lwm, l, dl, zc, zf, dzc, dzf,
                           visc, h, index_wm, 0, 0, uf, vf, wf))) return rc;
  if (any_wm) {
    double* ps[3] = {uf, vf, wf};
    if ((rc = extrap_launch(ctx, n, is_bound, dzci, ps, ifaces3, 3, 1, nullptr, lwm))) return rc;
  }
  if ((rc = strain_launch(ctx, n, dli, dzci, dzfi, uf, vf, wf, s0, sij, nullptr))) return rc;
  // (8)-(9a) Mij second part and cell-centred velocity                                     sgs.f90:262-277
  Alph al;
  for (int idir = 0; idir < 3; ++idir)
    for (int ib = 0; ib < 2; ++ib) al.wall[2 * idir + ib] = is_bound[tb(ib, idir)] && cbcvel[ib + 2 * idir + 6 * idir] == 'D';
  al.all = f2d ? 1 : 0;
  mij_interp_k<<<g, b, 0, ctx->stream>>>(d, al, s0, Csij, Pmij, u, v, w, uc, vc, wc, kc);
  KERNEL_CHECK(ctx);
  {
    double* ps[3] = {uc, vc, wc};
    if ((rc = k_boundp_multi(ctx, cbcsgs, n, bcs, nb, is_bound, dl, dzc, ps, 3))) return rc;   // sgs.f90:280-282
  }
  // (9b) Lij (stored in sij)                                                               sgs.f90:283-315
  prod_u_k<<<gflat, 256, 0, ctx->stream>>>(nflat, uc, vc, wc, Pwk);
  KERNEL_CHECK(ctx);
  if (!f2d && (rc = extrap_launch(ctx, n, is_bound, dzci, wk, ifaces0, 6, 0, cbcvel, nullptr))) return rc;
  if ((rc = filter_launch(ctx, n, wk, sij, 6, f2d))) return rc;
  {
    double* ps[3] = {uc, vc, wc};
    double* out[3] = {uf, vf, wf};
    if (!f2d && (rc = extrap_launch(ctx, n, is_bound, dzci, ps, ifaces0, 3, 0, cbcvel, nullptr))) return rc;
    if ((rc = filter_launch(ctx, n, ps, out, 3, f2d))) return rc;
  }
  // (10)-(11) contraction + averaging over the homogeneous directions                       sgs.f90:328-370
  // ctx->sgs_ave: 1 _CHANNEL (x-y planes; the reference's hard-wired choice, sgs.f90:8), 0 _DIT (whole volume), 2 _DUCT
  // (streamwise x), 3 _CAVITY (none)
  const int ave = ctx->sgs_ave;
  const int ntile = cdiv(n[0], BX) * cdiv(n[1], BY);
  double* part = (double*)cales_scratch(ctx, "sgs_part", (size_t)2 * n[2] * ntile * sizeof(double));
  double* p1d = (double*)cales_scratch(ctx, "sgs_p1d", (size_t)(2 * ng[2] + 2) * sizeof(double));
  if (!part || !p1d) return CALES_ERR_NOMEM;
  CUDA_TRY(ctx, cudaMemsetAsync(p1d, 0, (size_t)(2 * ng[2] + 2) * sizeof(double), ctx->stream));
  double *cml = nullptr, *cmm = nullptr;
  if (ave >= 2) { cml = wk[0]; cmm = wk[1]; }               // the reference keeps them in wk(:,:,:,1:2) too (sgs.f90:345-357)
  contract_k<<<dim3(cdiv(n[0], BX), cdiv(n[1], BY), n[2]), b, 0, ctx->stream>>>(d, Cmij, Csij, uf, vf, wf, part, ntile, cml, cmm);
  KERNEL_CHECK(ctx);
  if (ave == 2 || ave == 3) {
    if (ave == 2) {
      if (ctx->ipencil != 1) return cales_fail(ctx, CALES_ERR_INVALID, "_DUCT averaging needs X-aligned pencils");
      double* r2 = (double*)cales_scratch(ctx, "sgs_p2d", (size_t)2 * n[1] * n[2] * sizeof(double));
      if (!r2) return CALES_ERR_NOMEM;
      duct_rows_k<<<cdiv((long)n[1] * n[2], 8), 256, 0, ctx->stream>>>(d, cml, cmm, dl[0] / l[0], r2, r2 + (size_t)n[1] * n[2]);
      KERNEL_CHECK(ctx);
      dsmag_final_var_k<2><<<g, b, 0, ctx->stream>>>(d, r2, r2 + (size_t)n[1] * n[2], visct, kc);
    } else dsmag_final_var_k<3><<<g, b, 0, ctx->stream>>>(d, cml, cmm, visct, kc);
    KERNEL_CHECK(ctx);
    return CALES_OK;
  }
  const double gar = dl[0] * dl[1] / (l[0] * l[1]);
  plane_fold_k<<<2 * n[2], 256, 0, ctx->stream>>>(part, ntile, n[2], lo[2], ng[2], gar, p1d);
  KERNEL_CHECK(ctx);
  if (ave == 0) {
    double* p0d = p1d + 2 * ng[2];
    dit_fold_k<<<2, 256, 0, ctx->stream>>>(p1d, ng[2], lo[2], n[2], dzf, l[2], p0d);
    KERNEL_CHECK(ctx);
    if ((rc = k_allreduce_sum(ctx, p0d, 2))) return rc;                                      // sgs.f90:418
    dsmag_final_var_k<0><<<g, b, 0, ctx->stream>>>(d, p0d, nullptr, visct, kc);
    KERNEL_CHECK(ctx);
    return CALES_OK;
  }
  if ((rc = k_allreduce_sum(ctx, p1d, 2 * ng[2]))) return rc;                                // sgs.f90:475
  // (12)                                                                                   sgs.f90:372-380
  dsmag_final_k<<<g, b, 0, ctx->stream>>>(d, p1d, ng[2], lo[2], visct, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// the cpp switches of src/sgs.f90 that pick the dsmag averaging geometry and the test filter, selected at run time:
// ave 0 _DIT (ave0d_dit, sgs.f90:388-431), 1 _CHANNEL (ave1d_channel 433-538; the reference's hard-wired default, sgs.f90:8),
// 2 _DUCT (ave2d_duct 540-614, streamwise x), 3 _CAVITY (no averaging); filter_2d != 0: -D_FILTER_2D (filter2d 824-848)
extern "C" int cales_set_sgs_options(cales_ctx* ctx, int ave, int filter_2d) {
  CHECK_CTX(ctx);
  if (ave < 0 || ave > 3) return cales_fail(ctx, CALES_ERR_INVALID, "sgs averaging mode %d: 0 _DIT, 1 _CHANNEL, 2 _DUCT, 3 _CAVITY", ave);
  ctx->sgs_ave = ave; ctx->sgs_filter2d = filter_2d ? 1 : 0;
  return CALES_OK;
}
