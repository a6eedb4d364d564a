// Streaming stencils of the pressure-correction chain and the step checks.
//   fillps  src/fillps.f90:14-48      correc  src/correc.f90:14-68     updatep src/updatep.f90:14-49
//   chkdiv  src/chkdiv.f90:16-52      chkdt   src/chkdt.f90:17-99      bulk_mean src/utils.f90:16-47
//   bulk_forcing src/mom.f90:311-335
// Layout: thread (x,y) owns column (i,j) of a z-chunk and marches in k, so every array is read
// once per cell from HBM (k-1 values are carried in registers, i-1/j-1 values come from L1).
// Expressions keep the reference's association order; the library is built with -fmad=false so
// results are bit-identical to a non-contracting CPU build.
#include "common.cuh"
#include "reduce.cuh"

#define BX 64
#define BY 4

static inline dim3 grid_for(int ni, int nj, int nk, int kchunk) { return dim3(cdiv(ni, BX), cdiv(nj, BY), cdiv(nk, kchunk)); }

// choose a z-chunk so that the grid has a few waves of CTAs on 148 SMs
static inline int pick_kchunk(int ni, int nj, int nk) {
  return pick_chunk((long)cdiv(ni, BX) * cdiv(nj, BY), nk, 148 * 8, 8, 1);
}

__global__ void __launch_bounds__(BX* BY) fillps_k(Dims d, double dxi, double dyi, const double* __restrict__ dzfi, double dti,
                                                    const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, double* __restrict__ p, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  const double dtidxi = dti * dxi, dtidyi = dti * dyi;
  long c = d.idx(i, j, k0);
  double wm = w[c - d.s2];
#pragma unroll 4
  for (int k = k0; k <= k1; ++k, c += d.s2) {
    const double wc = w[c];
    p[c] = ((wc - wm) * dti * dzfi[k] + (v[c] - v[c - d.s1]) * dtidyi + (u[c] - u[c - 1]) * dtidxi);
    wm = wc;
  }
}

extern "C" int cales_fillps(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzfi, double dti,
                            const double* u, const double* v, const double* w, double* p) {
  CHECK_CTX(ctx);
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  fillps_k<<<grid_for(n[0], n[1], n[2], kc), dim3(BX, BY), 0, ctx->stream>>>(d, dli[0], dli[1], dzfi, dti, u, v, w, p, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// correc: u over i=0:n1, j=0:n2+1, k=0:n3+1; v over i=0:n1+1, j=0:n2, k=0:n3+1; w over k=0:n3 (ghost rows included).
// The update covers whole haloed planes, so threads are laid over the linear plane index q = i + (n1+2) j (no partial
// tiles, perfectly coalesced); CU levels are loaded together before any store.
#define CU 2
// FUSED (the fused substep, substep.cu): the corrected velocity is written to OTHER arrays than the ones read (uo,vo,wo:
// the caller's fields; u,v,w: the library's intermediate velocity), rows the correction leaves alone are copied, and the
// explicit pressure update p = p + pp (updatep.f90:45-47) of the interior rides along -- one pass instead of two.
template <bool FUSED>
__global__ void __launch_bounds__(256) correc_k(Dims d, double factori, double factorj, double dt, const double* __restrict__ dzci,
                                                const double* __restrict__ p, const double* u, const double* v, const double* w,
                                                double* uo, double* vo, double* wo, double* __restrict__ pres, int kc) {
  const long q = blockIdx.x * 256L + threadIdx.x;
  if (q >= d.s2) return;
  const int j = (int)(q / d.s1), i = (int)(q - j * d.s1);
  const bool du = i <= d.n1, dv = j <= d.n2;
  const bool inner = FUSED && i >= 1 && i <= d.n1 && j >= 1 && j <= d.n2;
  const int k0 = blockIdx.y * kc, k1 = min(k0 + kc - 1, d.n3 + 1);
  const long s1 = d.s1, s2 = d.s2;
  long c = q + s2 * k0;
  // p(i+1), p(j+1) of the last ghost row/plane are never used; keep the addresses inside the array
  const long oi = du ? 1 : 0, oj = dv ? s1 : 0;
  double pc = p[c];
  int k = k0;
  for (; k + CU - 1 <= k1 && k + CU - 1 <= d.n3; k += CU, c += CU * s2) {
    double pk[CU], pi[CU], pj[CU], uu[CU], vv[CU], ww[CU], dz[CU], pr[CU];
#pragma unroll
    for (int r = 0; r < CU; ++r) {
      const long cr = c + r * s2;
      pk[r] = p[cr + s2]; pi[r] = p[cr + oi]; pj[r] = p[cr + oj];
      uu[r] = u[cr]; vv[r] = v[cr]; ww[r] = w[cr]; dz[r] = dzci[k + r];
      if (FUSED) pr[r] = (inner && k + r >= 1) ? pres[cr] : 0.;
    }
#pragma unroll
    for (int r = 0; r < CU; ++r) {
      const long cr = c + r * s2;
      if (du) uo[cr] = uu[r] - factori * (pi[r] - pc); else if (FUSED) uo[cr] = uu[r];
      if (dv) vo[cr] = vv[r] - factorj * (pj[r] - pc); else if (FUSED) vo[cr] = vv[r];
      wo[cr] = ww[r] - dt * dz[r] * (pk[r] - pc);
      if (FUSED && inner && k + r >= 1) pres[cr] = pr[r] + pc;
      pc = pk[r];
    }
  }
  for (; k <= k1; ++k, c += s2) {
    const double pk = k <= d.n3 ? p[c + s2] : 0.0;
    if (du) uo[c] = u[c] - factori * (p[c + oi] - pc); else if (FUSED) uo[c] = u[c];
    if (dv) vo[c] = v[c] - factorj * (p[c + oj] - pc); else if (FUSED) vo[c] = v[c];
    if (k <= d.n3) wo[c] = w[c] - dt * dzci[k] * (pk - pc); else if (FUSED) wo[c] = w[c];
    if (FUSED && inner && k >= 1 && k <= d.n3) pres[c] = pres[c] + pc;
    pc = pk;
  }
}

extern "C" int cales_correc(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, double dt,
                            const double* p, double* u, double* v, double* w) {
  CHECK_CTX(ctx);
  Dims d(n);
  const int cols = cdiv(d.s2, 256);
  const int kc = pick_chunk(cols, n[2] + 2, 148 * 8, 8, 1);
  correc_k<false><<<dim3(cols, cdiv(n[2] + 2, kc)), 256, 0, ctx->stream>>>(d, dt * dli[0], dt * dli[1], dt, dzci, p, u, v, w, u, v, w, nullptr, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// correc from (us,vs,ws) into (u,v,w) + explicit updatep, one pass (fused substep)
int k_correc_updatep(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, double dt, const double* pp,
                     const double* us, const double* vs, const double* ws, double* u, double* v, double* w, double* p) {
  Dims d(n);
  const int cols = cdiv(d.s2, 256);
  const int kc = pick_chunk(cols, n[2] + 2, 148 * 8, 8, 1);
  correc_k<true><<<dim3(cols, cdiv(n[2] + 2, kc)), 256, 0, ctx->stream>>>(d, dt * dli[0], dt * dli[1], dt, dzci, pp, us, vs, ws, u, v, w, p, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

template <int MODE>  // 0 explicit, 1 implicit 3-D, 2 implicit z only
__global__ void __launch_bounds__(BX* BY) updatep_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci,
                                                     const double* __restrict__ dzfi, double alpha,
                                                     const double* __restrict__ pp, double* __restrict__ p, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  long c = d.idx(i, j, k0);
  if (MODE == 0) {
#pragma unroll 4
    for (int k = k0; k <= k1; ++k, c += d.s2) p[c] = p[c] + pp[c];
  } else {
    double pm = pp[c - d.s2], pc = pp[c];
    for (int k = k0; k <= k1; ++k, c += d.s2) {
      const double pk = pp[c + d.s2];
      double lap = ((pk - pc) * dzci[k] - (pc - pm) * dzci[k - 1]) * dzfi[k];
      if (MODE == 1)
        lap = (pp[c + 1] - 2. * pc + pp[c - 1]) * (dxi * dxi) + (pp[c + d.s1] - 2. * pc + pp[c - d.s1]) * (dyi * dyi) + lap;
      p[c] = p[c] + pc + alpha * lap;
      pm = pc; pc = pk;
    }
  }
}

extern "C" int cales_updatep(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, const double* dzfi,
                             double alpha, const double* pp, double* p) {
  CHECK_CTX(ctx);
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  dim3 g = grid_for(n[0], n[1], n[2], kc), b(BX, BY);
  if (ctx->diffusion == CALES_DIFF_EXPLICIT) updatep_k<0><<<g, b, 0, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, alpha, pp, p, kc);
  else if (ctx->diffusion == CALES_DIFF_IMPLICIT_3D) updatep_k<1><<<g, b, 0, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, alpha, pp, p, kc);
  else updatep_k<2><<<g, b, 0, ctx->stream>>>(d, dli[0], dli[1], dzci, dzfi, alpha, pp, p, kc);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// ---- bulk forcing -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX* BY) addconst_k(Dims d, const double* __restrict__ fdev, int comp, double* __restrict__ u, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  if (i > d.n1 || j > d.n2) return;
  const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
  const double ff = fdev[comp];
  long c = d.idx(i, j, k0);
  for (int k = k0; k <= k1; ++k, c += d.s2) u[c] = u[c] + ff;
}

int k_bulk_forcing_dev(cales_ctx* ctx, const int n[3], const int is_forced[3], const double* fdev, double* u, double* v, double* w) {
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  double* f[3] = {u, v, w};
  for (int c = 0; c < 3; ++c)
    if (is_forced[c]) {
      addconst_k<<<grid_for(n[0], n[1], n[2], kc), dim3(BX, BY), 0, ctx->stream>>>(d, fdev, c, f[c], kc);
      KERNEL_CHECK(ctx);
    }
  return CALES_OK;
}

extern "C" int cales_bulk_forcing(cales_ctx* ctx, const int n[3], const int is_forced[3], const double f[3], double* u,
                                  double* v, double* w) {
  CHECK_CTX(ctx);
  // f == NULL: use the device-resident f(3) left by the last cales_rk (no host round trip)
  if (!f) return k_bulk_forcing_dev(ctx, n, is_forced, ctx->fdev, u, v, w);
  double* tmp = (double*)cales_scratch(ctx, "bulkf_arg", 3 * sizeof(double));
  if (!tmp) return CALES_ERR_NOMEM;
  CUDA_TRY(ctx, cudaMemcpyAsync(tmp, f, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return k_bulk_forcing_dev(ctx, n, is_forced, tmp, u, v, w);
}

// ---- reductions ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BX* BY) bulk_mean_k(Dims d, const double* __restrict__ gvr, const double* __restrict__ p,
                                                       double* __restrict__ part, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  double s = 0.;
  if (i <= d.n1 && j <= d.n2) {
    const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
    long c = d.idx(i, j, k0);
    for (int k = k0; k <= k1; ++k, c += d.s2) s = s + p[c] * gvr[k];
  }
  s = block_sum<BX * BY>(s);
  if (threadIdx.x == 0 && threadIdx.y == 0) part[blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)] = s;
}

// device-resident result in out[0]; all-reduced over ranks
int k_bulk_mean_dev(cales_ctx* ctx, const int n[3], const double* gvr, const double* p, double* out) {
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  dim3 g = grid_for(n[0], n[1], n[2], kc);
  const int nb = g.x * g.y * g.z;
  double* part = (double*)cales_scratch(ctx, "red_part", (size_t)nb * 4 * sizeof(double));
  if (!part) return CALES_ERR_NOMEM;
  bulk_mean_k<<<g, dim3(BX, BY), 0, ctx->stream>>>(d, gvr, p, part, kc);
  KERNEL_CHECK(ctx);
  final_reduce_k<0><<<1, 256, 0, ctx->stream>>>(part, nb, out);
  KERNEL_CHECK(ctx);
  return k_allreduce_sum(ctx, out, 1);
}

extern "C" int cales_bulk_mean(cales_ctx* ctx, const int n[3], const double* grid_vol_ratio, const double* p, double* mean) {
  CHECK_CTX(ctx);
  int rc = k_bulk_mean_dev(ctx, n, grid_vol_ratio, p, ctx->red);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red_host, ctx->red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *mean = ctx->red_host[0];
  return CALES_OK;
}

__global__ void __launch_bounds__(BX* BY) chkdiv_k(Dims d, double dxi, double dyi, const double* __restrict__ dzfi,
                                                    const double* __restrict__ u, const double* __restrict__ v,
                                                    const double* __restrict__ w, double* __restrict__ part, int nblk, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  double s = 0., m = 0.;
  if (i <= d.n1 && j <= d.n2) {
    const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
    long c = d.idx(i, j, k0);
    double wm = w[c - d.s2];
    for (int k = k0; k <= k1; ++k, c += d.s2) {
      const double wc = w[c];
      const double div = (wc - wm) * dzfi[k] + (v[c] - v[c - d.s1]) * dyi + (u[c] - u[c - 1]) * dxi;
      m = fmax(m, fabs(div));
      s = s + div;
      wm = wc;
    }
  }
  s = block_sum<BX * BY>(s);
  m = block_max<BX * BY>(m);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    part[b] = s;
    part[nblk + b] = m;
  }
}

extern "C" int cales_chkdiv(cales_ctx* ctx, const int lo[3], const int hi[3], const double dli[3], const double* dzfi,
                            const double* u, const double* v, const double* w, double* divtot, double* divmax) {
  CHECK_CTX(ctx);
  const int n[3] = {hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  dim3 g = grid_for(n[0], n[1], n[2], kc);
  const int nb = g.x * g.y * g.z;
  double* part = (double*)cales_scratch(ctx, "red_part", (size_t)nb * 4 * sizeof(double));
  if (!part) return CALES_ERR_NOMEM;
  chkdiv_k<<<g, dim3(BX, BY), 0, ctx->stream>>>(d, dli[0], dli[1], dzfi, u, v, w, part, nb, kc);
  KERNEL_CHECK(ctx);
  final_reduce_k<0><<<1, 256, 0, ctx->stream>>>(part, nb, ctx->red);
  KERNEL_CHECK(ctx);
  final_reduce_k<1><<<1, 256, 0, ctx->stream>>>(part + nb, nb, ctx->red + 1);
  KERNEL_CHECK(ctx);
  int rc = k_allreduce_sum(ctx, ctx->red, 1);
  if (rc) return rc;
  rc = k_allreduce_minmax(ctx, ctx->red + 1, 1, 1);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red_host, ctx->red, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *divtot = ctx->red_host[0];
  *divmax = ctx->red_host[1];
  return CALES_OK;
}

template <int MODE>
__global__ void __launch_bounds__(BX* BY) chkdt_k(Dims d, double dxi, double dyi, const double* __restrict__ dzci,
                                                   const double* __restrict__ dzfi, double visc,
                                                   const double* __restrict__ visct, const double* __restrict__ u,
                                                   const double* __restrict__ v, const double* __restrict__ w,
                                                   double* __restrict__ part, int nblk, int kc) {
  const int i = blockIdx.x * BX + threadIdx.x + 1, j = blockIdx.y * BY + threadIdx.y + 1;
  double dti = 0., dtid = 0.;
  if (i <= d.n1 && j <= d.n2) {
    const double dl2i = dxi * dxi + dyi * dyi;
    const int k0 = blockIdx.z * kc + 1, k1 = min(k0 + kc - 1, d.n3);
    long c = d.idx(i, j, k0);
    const long s1 = d.s1, s2 = d.s2;
    for (int k = k0; k <= k1; ++k, c += s2) {
      const double ux = fabs(u[c]);
      const double vx = 0.25 * fabs(v[c] + v[c - s1] + v[c + 1] + v[c + 1 - s1]);
      const double wx = 0.25 * fabs(w[c] + w[c - s2] + w[c + 1] + w[c + 1 - s2]);
      const double uy = 0.25 * fabs(u[c] + u[c + s1] + u[c - 1 + s1] + u[c - 1]);
      const double vy = fabs(v[c]);
      const double wy = 0.25 * fabs(w[c] + w[c + s1] + w[c + s1 - s2] + w[c - s2]);
      const double uz = 0.25 * fabs(u[c] + u[c - 1] + u[c - 1 + s2] + u[c + s2]);
      const double vz = 0.25 * fabs(v[c] + v[c - s1] + v[c - s1 + s2] + v[c + s2]);
      const double wz = fabs(w[c]);
      const double dtix = ux * dxi + vx * dyi + wx * dzfi[k];
      const double dtiy = uy * dxi + vy * dyi + wy * dzfi[k];
      const double dtiz = uz * dxi + vz * dyi + wz * dzci[k];
      dti = fmax(fmax(fmax(dti, dtix), dtiy), dtiz);
      const double viscx = 0.5 * (visct[c] + visct[c + 1]);
      const double viscy = 0.5 * (visct[c] + visct[c + s1]);
      const double viscz = 0.5 * (visct[c] + visct[c + s2]);
      double dtidx = viscx * (dl2i + dzfi[k] * dzfi[k]);
      double dtidy = viscy * (dl2i + dzfi[k] * dzfi[k]);
      double dtidz = viscz * (dl2i + dzci[k] * dzci[k]);
      if (MODE != 1) {
        dtidx = dtidx + visc * dl2i; dtidy = dtidy + visc * dl2i; dtidz = dtidz + visc * dl2i;
        if (MODE == 0) {
          dtidx = dtidx + visc * (dzfi[k] * dzfi[k]);
          dtidy = dtidy + visc * (dzfi[k] * dzfi[k]);
          dtidz = dtidz + visc * (dzci[k] * dzci[k]);
        }
      }
      dtid = fmax(fmax(fmax(dtid, dtidx), dtidy), dtidz);
    }
  }
  dti = block_max<BX * BY>(dti);
  dtid = block_max<BX * BY>(dtid);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    part[b] = dti;
    part[nblk + b] = dtid;
  }
}

__global__ void chkdt_final_k(double* r, double eps) {
  double dti = r[0], dtid = r[1];
  if (dti == 0.) dti = 1.;
  if (dtid == 0.) dtid = eps;
  r[2] = fmin(0.4125 / dtid, 1.732 / dti);
}

extern "C" int cales_chkdt(cales_ctx* ctx, const int n[3], const double dl[3], const double* dzci, const double* dzfi,
                           double visc, const double* visct, const double* u, const double* v, const double* w, double* dtmax) {
  CHECK_CTX(ctx);
  Dims d(n);
  const int kc = pick_kchunk(n[0], n[1], n[2]);
  dim3 g = grid_for(n[0], n[1], n[2], kc), b(BX, BY);
  const int nb = g.x * g.y * g.z;
  double* part = (double*)cales_scratch(ctx, "red_part", (size_t)nb * 4 * sizeof(double));
  if (!part) return CALES_ERR_NOMEM;
  const double dxi = 1. / dl[0], dyi = 1. / dl[1];
  if (ctx->diffusion == CALES_DIFF_EXPLICIT) chkdt_k<0><<<g, b, 0, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, visct, u, v, w, part, nb, kc);
  else if (ctx->diffusion == CALES_DIFF_IMPLICIT_3D) chkdt_k<1><<<g, b, 0, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, visct, u, v, w, part, nb, kc);
  else chkdt_k<2><<<g, b, 0, ctx->stream>>>(d, dxi, dyi, dzci, dzfi, visc, visct, u, v, w, part, nb, kc);
  KERNEL_CHECK(ctx);
  final_reduce_k<1><<<1, 256, 0, ctx->stream>>>(part, nb, ctx->red);
  KERNEL_CHECK(ctx);
  final_reduce_k<1><<<1, 256, 0, ctx->stream>>>(part + nb, nb, ctx->red + 1);
  KERNEL_CHECK(ctx);
  chkdt_final_k<<<1, 1, 0, ctx->stream>>>(ctx->red, 2.220446049250313e-16);
  KERNEL_CHECK(ctx);
  int rc = k_allreduce_minmax(ctx, ctx->red + 2, 1, 0);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->red_host, ctx->red + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *dtmax = ctx->red_host[0];
  return CALES_OK;
}

// ---- small vector helpers of the implicit-diffusion sequence (src/main.f90:426-441) --------------------
__global__ void scale_k(long n, double alpha, const double* __restrict__ s, double* __restrict__ d) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) d[i] = s[i] * alpha;
}

extern "C" int cales_scale(cales_ctx* ctx, long count, double alpha, const double* src, double* dst) {
  CHECK_CTX(ctx);
  if (count <= 0) return CALES_OK;
  scale_k<<<cdiv(count, 256), 256, 0, ctx->stream>>>(count, alpha, src, dst);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

__global__ void helm_k(int n3, double alpha, const double* __restrict__ a, const double* __restrict__ b,
                       const double* __restrict__ c, double* __restrict__ aa, double* __restrict__ bb, double* __restrict__ cc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) { aa[i] = a[i] * alpha; bb[i] = b[i] * alpha + 1.; cc[i] = c[i] * alpha; }
}

extern "C" int cales_helmholtz_coeffs(cales_ctx* ctx, int n3, long nxy, double alpha, const double* a, const double* b,
                                      const double* c, const double* lambdaxy_in, double* aa, double* bb, double* cc,
                                      double* lambdaxy_out) {
  CHECK_CTX(ctx);
  helm_k<<<cdiv(n3, 128), 128, 0, ctx->stream>>>(n3, alpha, a, b, c, aa, bb, cc);
  KERNEL_CHECK(ctx);
  if (lambdaxy_in && lambdaxy_out) return cales_scale(ctx, nxy, alpha, lambdaxy_in, lambdaxy_out);
  return CALES_OK;
}
