// FFT-based direct Poisson/Helmholtz solver.
//   initsolver/eigenvalues/tridmatrix  src/initsolver.f90:17-169       fftini/find_fft src/fft.f90:23-143,192-245
//   solver        src/solver.f90:20-80  (GPU twin src/solver_gpu.f90:32-164)
//   gaussel / gaussel_periodic / dgtsv_homebrewed   src/solver.f90:82-179
//   solver_gaussel_z  src/solver.f90:182-233
// Data flow on one rank (X-pencils): the forward x pass reads the haloed field directly and writes the
// halo-free work array, y passes work in place on [x-chunk]x[all y] tiles, the z pass is one thread per
// (i,j) column (coalesced in i) running the reference's Thomas recurrence in its exact operation order
// (pivot regularisation `+eps` included), and the backward x pass writes the haloed field scaled by
// normfft: five sweeps, 16 B/cell each, no transposes and no copy-in/out passes.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

int k_fft_pass(cales_ctx* ctx, int dir, const char bc[2], char c_or_f, int backward, int n1, int n2, int n3,
               const double* in, long ip1, long ip2, double* out, long op1, long op2, double scale, const DivSrc* ds = nullptr);
int k_transpose(cales_ctx* ctx, int which, const double* src, double* dst);
int k_transpose_p2p(cales_ctx* ctx, int which, const double* src, PeerBuf* dst);
bool k_fft_peer_capable(const char bc[2], char c_or_f, int n);
bool k_gauss_tma_fits(int nxy, int n, int periodic);
int k_gaussel_tab(cales_ctx* ctx, int nx, int ny, int n, long sz, int periodic, const double* a, const double* b, const double* c,
                  const double* lambdaxy, double* p);
int k_zdist_prepare(cales_ctx* ctx, int plan, const Plan& pl, const double* lambdaxy, const double* a, const double* b, const double* c, bool zper);
int k_zdist_solve(cales_ctx* ctx, int plan, const double* lambdaxy, const double* a, const double* b, const double* c, double* w);

#define EPS 2.220446049250313e-16

// ---- tridiagonal solves along z -----------------------------------------------------------------------------------
// One thread per (i,j) column, coalesced in i.  The Thomas recurrences of solver.f90:165-178 are run in the
// reference's exact operation order, but in register tiles of TZ levels: every tile first issues its TZ
// independent loads, then walks the dependent chain, so the sweep is bandwidth- rather than latency-bound.
// The pivot recurrence d(l) depends only on (a,b,c,lambda): instead of storing it (the reference keeps a
// 3-D scratch array, solver_gpu.f90:166-231) the backward sweep recomputes it per tile from per-tile
// checkpoints held in shared memory -- identical arithmetic, hence bit-identical results, and the solve
// moves 32 B/cell (48 periodic) instead of 48 (88).
#define TZ 8
#define GT 64

// The recurrences are written branch-free so that every level costs one reciprocal and a handful of DFMA-pipe
// operations: with d(0)=p(0)=0 the first level needs no special case ((b+lam) - a*0 + eps is exact), with
// p(n+1):=0 neither does the last level of the back substitution, and the right-hand side of the second
// periodic system is a (mostly zero) table.  Coefficients a,b,c live in shared memory.
#define LOAD_TILE(dst, t, nlev, FULL)                                                        \
  {                                                                                          \
    const double* pt_ = pp + (long)(t) * TZ * sz;                                            \
    _Pragma("unroll") for (int q = 0; q < TZ; ++q) if (FULL || (t) * TZ + q < (nlev)) dst[q] = pt_[q * sz]; \
  }
#define STORE_TILE(src, t, nlev, FULL)                                                       \
  {                                                                                          \
    double* pt_ = pp + (long)(t) * TZ * sz;                                                  \
    _Pragma("unroll") for (int q = 0; q < TZ; ++q) if (FULL || (t) * TZ + q < (nlev)) pt_[q * sz] = src[q]; \
  }
#define COPY_TILE(dst, src) { _Pragma("unroll") for (int q = 0; q < TZ; ++q) dst[q] = src[q]; }
#define PIVOT(l) const double al = sa[l]; const double z = __drcp_rn((sb[l] + lam) - al * dl + EPS); dl = sc[l] * z;

template <int LAM>
__global__ void __launch_bounds__(GT) gaussel_k(int nxy, int n, long sz, const double* __restrict__ a, const double* __restrict__ b,
                                                 const double* __restrict__ c, const double* __restrict__ lambdaxy, double* __restrict__ p) {
  extern __shared__ double sh[];
  const int ntile = (n + TZ - 1) / TZ, nfull = n / TZ;
  double* sa = sh; double* sb = sa + n; double* sc = sb + n;
  double* ck = sc + n;                                 // checkpoints [ntile][GT]: d at the level before the tile
  for (int l = threadIdx.x; l < n; l += GT) { sa[l] = a[l]; sb[l] = b[l]; sc[l] = c[l]; }
  __syncthreads();
  const int col = blockIdx.x * GT + threadIdx.x;
  if (col >= nxy) return;
  const double lam = LAM ? lambdaxy[col] : 0.;
  double* pp = p + col;
  double dl = 0., pl = 0.;
  double r[TZ], rn[TZ];
  // ---- forward elimination (solver.f90:165-173)
  if (nfull > 0) LOAD_TILE(r, 0, n, true) else LOAD_TILE(r, 0, n, false)
  for (int t = 0; t < nfull; ++t) {
    if (t + 1 < nfull) LOAD_TILE(rn, t + 1, n, true) else if (t + 1 < ntile) LOAD_TILE(rn, t + 1, n, false)
    ck[t * GT + threadIdx.x] = dl;
#pragma unroll
    for (int q = 0; q < TZ; ++q) {
      const int l = t * TZ + q;
      const double dprev = dl;
      PIVOT(l)
      (void)dprev;
      pl = (r[q] - al * pl) * z;
      r[q] = pl;
    }
    STORE_TILE(r, t, n, true)
    COPY_TILE(r, rn)
  }
  if (nfull < ntile) {
    const int t = nfull;
    ck[t * GT + threadIdx.x] = dl;
#pragma unroll
    for (int q = 0; q < TZ; ++q) {
      const int l = t * TZ + q;
      if (l < n) { PIVOT(l) pl = (r[q] - al * pl) * z; r[q] = pl; }
    }
    STORE_TILE(r, t, n, false)
  }
  // ---- backward substitution (solver.f90:176-178): p(l) = p(l) - d(l)*p(l+1) with p(n+1) := 0
  pl = 0.;
  if (nfull < ntile) {
    const int t = nfull;
    double d[TZ];
    LOAD_TILE(r, t, n, false)
    dl = ck[t * GT + threadIdx.x];
#pragma unroll
    for (int q = 0; q < TZ; ++q) { const int l = t * TZ + q; if (l < n) { PIVOT(l) (void)z; d[q] = dl; } }
#pragma unroll
    for (int q = TZ - 1; q >= 0; --q) { const int l = t * TZ + q; if (l < n) { pl = r[q] - d[q] * pl; r[q] = pl; } }
    STORE_TILE(r, t, n, false)
  }
  if (nfull > 0) LOAD_TILE(r, nfull - 1, n, true)
  for (int t = nfull - 1; t >= 0; --t) {
    double d[TZ];
    if (t > 0) LOAD_TILE(rn, t - 1, n, true)
    dl = ck[t * GT + threadIdx.x];
#pragma unroll
    for (int q = 0; q < TZ; ++q) { const int l = t * TZ + q; PIVOT(l) (void)z; d[q] = dl; }
#pragma unroll
    for (int q = TZ - 1; q >= 0; --q) { pl = r[q] - d[q] * pl; r[q] = pl; }
    STORE_TILE(r, t, n, true)
    COPY_TILE(r, rn)
  }
}

// periodic: solver.f90:109-151.  Two systems of size n-1 share the matrix: rhs p(1:n-1) and
// [-a(1),0,...,0,-c(n-1)] (p2).  p2 needs no loads, so it is recomputed where needed; three sweeps.
template <int LAM>
__global__ void __launch_bounds__(GT) gaussel_periodic_k(int nxy, int n, long sz, const double* __restrict__ a,
                                                          const double* __restrict__ b, const double* __restrict__ c,
                                                          const double* __restrict__ lambdaxy, double* __restrict__ p) {
  extern __shared__ double sh[];
  const int nm = n - 1;
  const int ntile = (nm + TZ - 1) / TZ, nfull = nm / TZ;
  double* sa = sh; double* sb = sa + n; double* sc = sb + n; double* s2 = sc + n;
  double* ckd = s2 + n;                                // [2][ntile][GT]: d and forward-p2 before each tile
  double* ck2 = ckd + (size_t)ntile * GT;
  for (int l = threadIdx.x; l < n; l += GT) { sa[l] = a[l]; sb[l] = b[l]; sc[l] = c[l]; s2[l] = 0.; }
  __syncthreads();
  if (threadIdx.x == 0) { s2[0] = -a[0]; s2[nm - 1] = -c[nm - 1]; }   // p2(1) = -a(1); p2(n-1) = -c(n-1) (in that order)
  __syncthreads();
  const int col = blockIdx.x * GT + threadIdx.x;
  if (col >= nxy) return;
  const double lam = LAM ? lambdaxy[col] : 0.;
  double* pp = p + col;
  double dl = 0., p1l = 0., p2l = 0.;
  double r[TZ], rn[TZ];
  // ---- forward elimination of both systems
  if (nfull > 0) LOAD_TILE(r, 0, nm, true) else LOAD_TILE(r, 0, nm, false)
  for (int t = 0; t < ntile; ++t) {
    const bool full = t < nfull;
    if (t + 1 < nfull) LOAD_TILE(rn, t + 1, nm, true) else if (t + 1 < ntile) LOAD_TILE(rn, t + 1, nm, false)
    ckd[t * GT + threadIdx.x] = dl;
    ck2[t * GT + threadIdx.x] = p2l;
#pragma unroll
    for (int q = 0; q < TZ; ++q) {
      const int l = t * TZ + q;
      if (full || l < nm) {
        PIVOT(l)
        p1l = (r[q] - al * p1l) * z;
        p2l = (s2[l] - al * p2l) * z;
        r[q] = p1l;
      }
    }
    if (full) STORE_TILE(r, t, nm, true) else STORE_TILE(r, t, nm, false)
    COPY_TILE(r, rn)
  }
  const double p1n = p1l, p2n = p2l;                   // p1(n-1), p2(n-1): unchanged by the back substitution
  const double plast = pp[(long)(n - 1) * sz];
  // ---- backward substitution of both systems (p(n) := 0 makes the last level uniform)
  p1l = 0.; p2l = 0.;
  if (ntile - 1 < nfull) LOAD_TILE(r, ntile - 1, nm, true) else LOAD_TILE(r, ntile - 1, nm, false)
  for (int t = ntile - 1; t >= 0; --t) {
    const bool full = t < nfull;
    double d[TZ], f2[TZ];
    if (t > 0) LOAD_TILE(rn, t - 1, nm, true)
    dl = ckd[t * GT + threadIdx.x];
    double g2 = ck2[t * GT + threadIdx.x];
#pragma unroll
    for (int q = 0; q < TZ; ++q) {
      const int l = t * TZ + q;
      if (full || l < nm) { PIVOT(l) g2 = (s2[l] - al * g2) * z; d[q] = dl; f2[q] = g2; }
    }
#pragma unroll
    for (int q = TZ - 1; q >= 0; --q) {
      const int l = t * TZ + q;
      if (full || l < nm) { p1l = r[q] - d[q] * p1l; p2l = f2[q] - d[q] * p2l; r[q] = p1l; }
    }
    if (full) STORE_TILE(r, t, nm, true) else STORE_TILE(r, t, nm, false)
    COPY_TILE(r, rn)
  }
  // p1l = p1(1), p2l = p2(1)                                                     solver.f90:142-144
  const double pn = (plast - sc[n - 1] * p1l - sa[n - 1] * p1n) / ((sb[n - 1] + lam) + sc[n - 1] * p2l + sa[n - 1] * p2n + EPS);
  pp[(long)(n - 1) * sz] = pn;
  // ---- p(1:n-1) = p1 + p2*p(n): p2 recomputed once more
  p2l = 0.;
  if (ntile - 1 < nfull) LOAD_TILE(r, ntile - 1, nm, true) else LOAD_TILE(r, ntile - 1, nm, false)
  for (int t = ntile - 1; t >= 0; --t) {
    const bool full = t < nfull;
    double d[TZ], f2[TZ];
    if (t > 0) LOAD_TILE(rn, t - 1, nm, true)
    dl = ckd[t * GT + threadIdx.x];
    double g2 = ck2[t * GT + threadIdx.x];
#pragma unroll
    for (int q = 0; q < TZ; ++q) {
      const int l = t * TZ + q;
      if (full || l < nm) { PIVOT(l) g2 = (s2[l] - al * g2) * z; d[q] = dl; f2[q] = g2; }
    }
#pragma unroll
    for (int q = TZ - 1; q >= 0; --q) {
      const int l = t * TZ + q;
      if (full || l < nm) { p2l = f2[q] - d[q] * p2l; r[q] = r[q] + p2l * pn; }
    }
    if (full) STORE_TILE(r, t, nm, true) else STORE_TILE(r, t, nm, false)
    COPY_TILE(r, rn)
  }
}

// p: halo-free (nx,ny,>=n) array, plane stride sz
int k_gaussel(cales_ctx* ctx, int nx, int ny, int n, long sz, int periodic, const double* a, const double* b, const double* c,
              const double* lambdaxy, double* p) {
  const int nxy = nx * ny;
  if (periodic && n < 3) return cales_fail(ctx, CALES_ERR_INVALID, "periodic tridiagonal solve needs n >= 3");
  {
    // fast path: cached pivots (gaussel_tab.cu); CALES_GAUSSEL_DIRECT=1 forces the recompute-everything kernels below
    static const bool direct = getenv("CALES_GAUSSEL_DIRECT") != nullptr;
    if (!direct) {
      const int rc = k_gaussel_tab(ctx, nx, ny, n, sz, periodic, a, b, c, lambdaxy, p);
      if (rc < 0) return -rc;
      if (rc == 1) return CALES_OK;
    }
  }
  if (ctx->gauss_peer_out) return cales_fail(ctx, CALES_ERR_INVALID, "peer-fused z solve requested but the cached-pivot kernel is unavailable");
  const int ntile = ((periodic ? n - 1 : n) + TZ - 1) / TZ;
  const size_t sh = ((size_t)ntile * GT * (periodic ? 2 : 1) + (size_t)n * (periodic ? 4 : 3)) * sizeof(double);
  if (sh > 200 * 1024) return cales_fail(ctx, CALES_ERR_INVALID, "tridiagonal system of %d points exceeds the checkpoint buffer", n);
  dim3 g(cdiv(nxy, GT));
  static bool attr = false;
  if (!attr) {
    attr = true;
    cudaFuncSetAttribute(gaussel_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gaussel_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gaussel_periodic_k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gaussel_periodic_k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  if (!periodic) {
    if (lambdaxy) gaussel_k<1><<<g, GT, sh, ctx->stream>>>(nxy, n, sz, a, b, c, lambdaxy, p);
    else gaussel_k<0><<<g, GT, sh, ctx->stream>>>(nxy, n, sz, a, b, c, lambdaxy, p);
  } else {
    if (lambdaxy) gaussel_periodic_k<1><<<g, GT, sh, ctx->stream>>>(nxy, n, sz, a, b, c, lambdaxy, p);
    else gaussel_periodic_k<0><<<g, GT, sh, ctx->stream>>>(nxy, n, sz, a, b, c, lambdaxy, p);
  }
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

extern "C" int cales_gaussel(cales_ctx* ctx, int nx, int ny, int n, int periodic, const double* a, const double* b,
                             const double* c, const double* lambdaxy, double* p) {
  CHECK_CTX(ctx);
  return k_gaussel(ctx, nx, ny, n, (long)nx * ny, periodic, a, b, c, lambdaxy, p);
}

// ---- initsolver --------------------------------------------------------------------------------------------------------
static void eigenvalues(int n, const char cbc[2], char c_or_f, std::vector<double>& lam) {   // initsolver.f90:66-125
  const double pi = acos(-1.0);
  lam.assign(n, 0.);
  const bool PP = cbc[0] == 'P' && cbc[1] == 'P', NN = cbc[0] == 'N' && cbc[1] == 'N', DD = cbc[0] == 'D' && cbc[1] == 'D';
  for (int l = 1; l <= n; ++l) {
    double v = 0.;
    if (PP) v = -2. * (1. - cos((2 * (l - 1)) * pi / (1. * n)));
    else if (NN) v = c_or_f == 'c' ? -2. * (1. - cos((l - 1) * pi / (1. * n))) : -2. * (1. - cos((l - 1) * pi / (1. * (n - 1 + 1))));
    else if (DD) {
      if (c_or_f == 'c') v = -2. * (1. - cos(l * pi / (1. * n)));
      else v = l < n ? -2. * (1. - cos(l * pi / (1. * (n + 1 - 1)))) : 0.;
    } else v = -2. * (1. - cos((2 * l - 1) * pi / (2. * n)));
    lam[l - 1] = v;
  }
}

extern "C" int cales_initsolver(cales_ctx* ctx, const int ng[3], const int n_x_fft[3], const int n_y_fft[3], const int lo_z[3],
                                const int hi_z[3], const double dli[3], const double* dzci_g, const double* dzfi_g,
                                const char cbc[6], const char c_or_f[3], double* lambdaxy, double* a, double* b, double* c,
                                int* plan, double* normfft) {
  CHECK_CTX(ctx);
  (void)n_x_fft; (void)n_y_fft;
  std::vector<double> lx, ly;
  const char bcx[2] = {cbc[tb(0, 0)], cbc[tb(1, 0)]}, bcy[2] = {cbc[tb(0, 1)], cbc[tb(1, 1)]}, bcz[2] = {cbc[tb(0, 2)], cbc[tb(1, 2)]};
  eigenvalues(ng[0], bcx, c_or_f[0], lx);
  eigenvalues(ng[1], bcy, c_or_f[1], ly);
  for (auto& v : lx) v = v * (dli[0] * dli[0]);
  for (auto& v : ly) v = v * (dli[1] * dli[1]);
  const int nzx = hi_z[0] - lo_z[0] + 1;
  for (int j = lo_z[1]; j <= hi_z[1]; ++j)                  // initsolver.f90:51-55
    for (int i = lo_z[0]; i <= hi_z[0]; ++i) lambdaxy[(i - lo_z[0]) + (long)nzx * (j - lo_z[1])] = lx[i - 1] + ly[j - 1];
  const int n = ng[2];                                      // tridmatrix, initsolver.f90:127-169
  for (int k = 1; k <= n; ++k) {
    if (c_or_f[2] == 'c') { a[k - 1] = dzfi_g[k] * dzci_g[k - 1]; c[k - 1] = dzfi_g[k] * dzci_g[k]; }
    else { a[k - 1] = dzfi_g[k] * dzci_g[k]; c[k - 1] = dzfi_g[k + 1] * dzci_g[k]; }
    b[k - 1] = -(a[k - 1] + c[k - 1]);
  }
  double factor[2];
  for (int ib = 0; ib < 2; ++ib) factor[ib] = bcz[ib] == 'P' ? 0. : bcz[ib] == 'D' ? -1. : 1.;
  if (c_or_f[2] == 'c') {
    b[0] = b[0] + factor[0] * a[0];
    b[n - 1] = b[n - 1] + factor[1] * c[n - 1];
  } else {
    if (bcz[0] == 'N') b[0] = b[0] + factor[0] * a[0];
    if (bcz[1] == 'N') b[n - 1] = b[n - 1] + factor[1] * c[n - 1];
  }
  // fftini: normfft (fft.f90:66-69,99,106,136,142) and the plan record
  Plan pl;
  pl.used = true;
  double nf = 1.;
  for (int d = 0; d < 2; ++d) {
    const char* bc = d == 0 ? bcx : bcy;
    double norm[2];
    const bool PP = bc[0] == 'P' && bc[1] == 'P', NN = bc[0] == 'N' && bc[1] == 'N', DD = bc[0] == 'D' && bc[1] == 'D';
    if (PP) { norm[0] = 1.; norm[1] = 0.; }
    else if (c_or_f[d] == 'f' && NN) { norm[0] = 2.; norm[1] = -1.; }
    else if (c_or_f[d] == 'f' && DD) { norm[0] = 2.; norm[1] = 1.; }
    else { norm[0] = 2.; norm[1] = 0.; }
    const int ix = (DD && c_or_f[d] == 'f') ? 1 : 0;
    nf = nf * norm[0] * (ng[d] + norm[1] - ix);
    pl.bc[d][0] = bc[0]; pl.bc[d][1] = bc[1]; pl.c_or_f[d] = c_or_f[d];
  }
  pl.normfft = 1. / nf;
  memcpy(pl.ng, ng, sizeof pl.ng);
  pl.lx = lx; pl.ly = ly; pl.a.assign(a, a + n); pl.b.assign(b, b + n); pl.c.assign(c, c + n);
  pl.bcz[0] = bcz[0]; pl.bcz[1] = bcz[1];
  *normfft = pl.normfft;
  int h = -1;
  for (size_t q = 0; q < ctx->plans.size(); ++q) if (!ctx->plans[q].used) { h = (int)q; break; }
  if (h < 0) { ctx->plans.push_back(pl); h = (int)ctx->plans.size() - 1; } else ctx->plans[h] = pl;
  *plan = h;
  return CALES_OK;
}

extern "C" const char* cales_solver_exchange(const cales_ctx* ctx) {
  static const char* nm[5] = {"none (z on one rank)", "distributed z solve (zdist.cu): 2 boundary planes per rank, no y<->z transposes",
                              "copy-engine pipelined y<->z transposes", "kernel-fused y<->z transposes", "separate transpose kernels / NCCL"};
  return ctx && ctx->solver_path >= 0 && ctx->solver_path < 5 ? nm[ctx->solver_path] : "?";
}

extern "C" int cales_fftend(cales_ctx* ctx, int plan) {
  CHECK_CTX(ctx);
  if (plan < 0 || plan >= (int)ctx->plans.size()) return cales_fail(ctx, CALES_ERR_INVALID, "fftend: bad plan handle %d", plan);
  ctx->plans[plan].used = false;
  return CALES_OK;
}

// ---- solver --------------------------------------------------------------------------------------------------------------
extern "C" int cales_solver(cales_ctx* ctx, const int n[3], const int ng[3], int plan, double normfft, const double* lambdaxy,
                            const double* a, const double* b, const double* c, const char bc[6], const char c_or_f[3], double* p) {
  CHECK_CTX(ctx);
  if (plan < 0 || plan >= (int)ctx->plans.size() || !ctx->plans[plan].used) return cales_fail(ctx, CALES_ERR_INVALID, "solver: bad plan handle %d", plan);
  if (ctx->ipencil != 1) return cales_fail(ctx, CALES_ERR_INVALID, "solver: only X-aligned pencils (_DECOMP_X, the reference default) are implemented");
  const Plan& pl = ctx->plans[plan];
  Dims d(n);
  const char bcz[2] = {bc[tb(0, 2)], bc[tb(1, 2)]};
  const int q = (c_or_f[2] == 'f' && bcz[1] == 'D') ? 1 : 0;
  const bool zper = bcz[0] == 'P' && bcz[1] == 'P';
  int rc;
  // Right-hand side still to be formed (cales_substep hands fillps + updt_rhs_b over): the forward x pass computes it on the
  // fly where the register-blocked transform runs on whole planes (one rank; distributed z solve) -- 32 instead of 48 B/cell
  // for fillps + pass, one launch less; everywhere else the two kernels run here first.  CALES_FUSE_FILLPS=0 always splits.
  const DivSrc* ds_in = ctx->div_src;
  ctx->div_src = nullptr;
  DivSrc dsv;
  const DivSrc* ds = nullptr;
  if (ds_in) {
    static const int fuse_env = getenv("CALES_FUSE_FILLPS") ? atoi(getenv("CALES_FUSE_FILLPS")) : 1;
    dsv = *ds_in;
    const long o111 = d.idx(1, 1, 1);
    dsv.u += o111; dsv.v += o111; dsv.w += o111;
    // measured at 256^3 on one GPU: step 3.635 -> 3.619 ms (the forward x pass is bound by its load/shared-memory pipe, so
    // most of the 16 B/cell saved comes back as extra load instructions).  On by default on one rank, where the fused
    // step is held to the per-procedure sequence by the tests; across ranks only with CALES_FUSE_FILLPS=2.
    if (fuse_env && (ctx->nranks == 1 || fuse_env == 2) && k_fft_peer_capable(pl.bc[0], pl.c_or_f[0], n[0])) ds = &dsv;
  }
  auto form_rhs = [&]() -> int {                            // the unfused pair, for the paths that do not take `ds`
    if (!ds_in) return CALES_OK;
    const double dli_[3] = {ds_in->dxi, ds_in->dyi, 0.};
    int r_;
    if ((r_ = cales_fillps(ctx, n, dli_, ds_in->dzfi, ds_in->dti, ds_in->u, ds_in->v, ds_in->w, p))) return r_;
    return cales_updt_rhs_b(ctx, "ccc", bc, n, ds_in->bnd, ds_in->rbx, ds_in->rby, ds_in->rbz, p);
  };
  if (ctx->nranks == 1) {
    if (!ds && (rc = form_rhs())) return rc;
    const long p1 = n[0], p2 = (long)n[0] * n[1];
    double* wk = (double*)cales_scratch(ctx, "solver_wk", (size_t)p2 * n[2] * sizeof(double));
    if (!wk) return CALES_ERR_NOMEM;
    // fwd x: haloed p -> wk ; fwd y in place ; z solve ; bwd y ; bwd x: wk -> haloed p * normfft
    if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, n[0], n[1], n[2], p + d.idx(1, 1, 1), d.s1, d.s2, wk, p1, p2, 1.0, ds))) return rc;
    if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, n[0], n[1], n[2], wk, p1, p2, wk, p1, p2, 1.0))) return rc;
    if ((rc = k_gaussel(ctx, n[0], n[1], n[2] - q, p2, zper, a, b, c, lambdaxy, wk))) return rc;
    if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 1, n[0], n[1], n[2], wk, p1, p2, wk, p1, p2, 1.0))) return rc;
    if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 1, n[0], n[1], n[2], wk, p1, p2, p + d.idx(1, 1, 1), d.s1, d.s2, normfft))) return rc;
    return CALES_OK;
  }
  // distributed: x-pencil (ng1, n2, n3) -> y-pencil -> z-pencil and back (solver.f90:48-69)
  const int* xs = ctx->xsz; const int* ys = ctx->ysz; const int* zs = ctx->zsz;
  const size_t bx = (size_t)xs[0] * xs[1] * xs[2], by = (size_t)ys[0] * ys[1] * ys[2], bz = (size_t)zs[0] * zs[1] * zs[2];
  size_t bmax = bx > by ? bx : by; bmax = bmax > bz ? bmax : bz;
  // the largest pencil over ALL ranks bounds the peer buffers (uneven splits)
  {
    size_t m = 0;
    for (int r = 0; r < ctx->nranks; ++r)
      for (int ax = 1; ax <= 3; ++ax) {
        int lo[3], hi[3], sz[3];
        cales_pencil(ctx->ng, ctx->dims, r, ax, lo, hi, sz);
        const size_t v = (size_t)sz[0] * sz[1] * sz[2];
        if (v > m) m = v;
      }
    bmax = m;
  }
  PeerBuf* pb0 = k_peer_buffer(ctx, "solver_wk", bmax * sizeof(double));
  PeerBuf* pb1 = pb0 ? k_peer_buffer(ctx, "solver_wk1", bmax * sizeof(double)) : nullptr;
  const bool p2p = pb0 && pb1;
  double* w0 = p2p ? (double*)pb0->local : (double*)cales_scratch(ctx, "solver_wk", bmax * sizeof(double));
  double* w1 = p2p ? (double*)pb1->local : (double*)cales_scratch(ctx, "solver_wk1", bmax * sizeof(double));
  if (!w0 || !w1) return CALES_ERR_NOMEM;
  // ---- distributed z solve (zdist.cu; default in the product build with peer memory): z stays decomposed, the y <-> z
  // transposes disappear -- two boundary planes per rank cross NVLink instead of the whole spectrum twice.  Pressure plans
  // only (static coefficients; q == 0); CALES_ZDIST=0/1 forces it off/on (on in the strict build = tolerance parity there).
  {
#ifdef CALES_FMA
    const int zd_default = 1;
#else
    const int zd_default = 0;
#endif
    const int zd_env = getenv("CALES_ZDIST") ? atoi(getenv("CALES_ZDIST")) : -1;   // read per call: bench.py switches it off after a failed parity check
    if (p2p && (zd_env >= 0 ? zd_env != 0 : zd_default != 0) && ctx->dims[1] > 1 && ctx->dims[1] <= 8 && lambdaxy && q == 0 && n[2] >= 2) {
      const int zr = k_zdist_prepare(ctx, plan, pl, lambdaxy, a, b, c, zper);
      if (zr < 0) return -zr;
      if (zr == 1) {
        ctx->solver_path = 1;
        double *cur = w0, *oth = w1;
        PeerBuf *pcur = pb0, *poth = pb1;
        const bool xy = ctx->dims[0] > 1;
        const long ypl = (long)ys[0] * ys[1];
        if (!ds && (rc = form_rhs())) return rc;
        if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, xs[0], xs[1], xs[2], p + d.idx(1, 1, 1), d.s1, d.s2, cur, xs[0], (long)xs[0] * xs[1], 1.0, ds))) return rc;
        if (xy) { if ((rc = k_transpose_p2p(ctx, 0, cur, poth))) return rc; std::swap(cur, oth); std::swap(pcur, poth); }
        if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, ys[0], ys[1], ys[2], cur, ys[0], ypl, cur, ys[0], ypl, 1.0))) return rc;
        if ((rc = k_zdist_solve(ctx, plan, lambdaxy, a, b, c, cur))) return rc;
        if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 1, ys[0], ys[1], ys[2], cur, ys[0], ypl, cur, ys[0], ypl, 1.0))) return rc;
        if (xy) { if ((rc = k_transpose_p2p(ctx, 3, cur, poth))) return rc; std::swap(cur, oth); std::swap(pcur, poth); }
        return k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 1, xs[0], xs[1], xs[2], cur, xs[0], (long)xs[0] * xs[1], p + d.idx(1, 1, 1), d.s1, d.s2, normfft);
      }
    }
  }
  if ((rc = form_rhs())) return rc;                         // the transposing paths below read their right-hand side from p
  // ---- pipelined exchange (default with peer memory): the y <-> z transposes are done by the COPY ENGINES, chunk by chunk,
  // while the SMs work on the next chunk.  y -> z: the local z range is cut into chunks; as soon as the x and y transforms of
  // a chunk are done, one strided 2-D DMA per peer pushes its [all x, y of that peer, chunk] box straight into the peer's
  // Z-pencil (contiguous there) over NVLink.  z -> y: the columns of my Z-pencil are cut into y chunks; each chunk is
  // solved and then pushed into the peers' Y-pencils the same way.  One flag barrier per direction.  Peer writes run at the
  // link rate whatever the producing kernel looks like (tools/p2pbench: 650-700 GB/s per direction), every transform kind
  // and process grid takes this path, and no kernel stores across NVLink.  CALES_SOLVER_PIPE=0 restores the older paths.
  // Selection (measured, profiles/r2h_*): where the kernel-fused exchange below exists (1 x P grids, transform with a peer
  // post-stage, column short enough for the TMA z solve) it wins at 256^3 per GPU (0.70 vs 0.78 ms at 2 GPUs: the chunks
  // cost wave quantisation); everywhere else the pipeline replaces the unfused chain.  CALES_SOLVER_PIPE=0/1 forces a path.
  static const int pipe_env = getenv("CALES_SOLVER_PIPE") ? atoi(getenv("CALES_SOLVER_PIPE")) : -1;
  // chunks: 1 below ~24 M cells per rank (at 256^3 per GPU chunked kernels lose more to wave quantisation than the overlap
  // returns: 6.49 ms/step with 1 chunk, 9.33 with 4 at 8 GPUs), 2 above
  static const int nchunk_cfg = getenv("CALES_SOLVER_CHUNKS") ? atoi(getenv("CALES_SOLVER_CHUNKS")) : 0;
  const int nchunk_env = nchunk_cfg > 0 ? nchunk_cfg : ((long)xs[0] * xs[1] * xs[2] < 24000000L ? 1 : 2);
  // (the kernel-fused periodic z solve is verified on two ranks only: with more, the pipeline takes over)
  const bool fused_ok = ctx->dims[0] == 1 && ctx->dims[1] <= 8 && lambdaxy && k_fft_peer_capable(pl.bc[1], pl.c_or_f[1], ys[1]) && q == 0 &&
                        k_gauss_tma_fits(zs[0] * zs[1], zs[2], zper) && (!zper || ctx->dims[1] == 2);
  const bool pipe = pipe_env >= 0 ? pipe_env != 0 : !fused_ok;
  if (p2p && pipe && ctx->dims[1] > 1 && ctx->dims[1] <= 16 && lambdaxy) {
    const int P = ctx->dims[1], me = ctx->coord[1];
    ctx->solver_path = 2;
    if (!ctx->side[0]) {
      for (auto& st_ : ctx->side) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
      for (auto& sev : ctx->side_ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&sev, cudaEventDisableTiming));
    }
    std::vector<int> yst(P), yen(P), yszq(P), zst(P), zen(P), zszq(P);
    cales_distribute(ctx->ng[1], P, yst.data(), yen.data(), yszq.data());
    cales_distribute(ctx->ng[2], P, zst.data(), zen.data(), zszq.data());
    const long nxl = ys[0], ny = ys[1], nzl = ys[2], nyl = zs[1];
    const size_t E = sizeof(double);
    auto rankof = [&](int q) { return ctx->coord[0] * ctx->dims[1] + q; };
    double *cur = w0, *oth = w1;
    PeerBuf *pcur = pb0, *poth = pb1;
    const bool xy = ctx->dims[0] > 1;
    if (xy) {                                               // x transform + x -> y transpose first (not pipelined)
      if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, xs[0], xs[1], xs[2], p + d.idx(1, 1, 1), d.s1, d.s2, cur, xs[0], (long)xs[0] * xs[1], 1.0))) return rc;
      if ((rc = k_transpose_p2p(ctx, 0, cur, poth))) return rc;
      std::swap(cur, oth); std::swap(pcur, poth);
    }
    // ---- forward: [x transform,] y transform and push, chunk by chunk over my z planes.  Streams side[0..2]: NVLink pushes
    // (chunks round-robin), side[3]: the local boxes (another copy engine).
    auto join_side = [&]() -> int {                          // the compute stream waits for every copy issued so far
      for (int q4 = 0; q4 < 4; ++q4) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->side_ev[4 + q4], ctx->side[q4]));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->side_ev[4 + q4], 0));
      }
      return CALES_OK;
    };
    const int nck = std::max(1, std::min(std::min(nchunk_env, 4), (int)nzl));
    for (int ch = 0; ch < nck; ++ch) {
      const int k0 = (int)((long)nzl * ch / nck), k1 = (int)((long)nzl * (ch + 1) / nck);
      if (k1 <= k0) continue;
      if (!xy && (rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, xs[0], xs[1], k1 - k0, p + d.idx(1, 1, 1 + k0), d.s1, d.s2,
                                  cur + (long)xs[0] * xs[1] * k0, xs[0], (long)xs[0] * xs[1], 1.0))) return rc;
      double* yc = cur + nxl * ny * k0;
      if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, ys[0], ys[1], k1 - k0, yc, ys[0], nxl * ny, yc, ys[0], nxl * ny, 1.0))) return rc;
      CUDA_TRY(ctx, cudaEventRecord(ctx->side_ev[ch % 4], ctx->stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side[ch % 3], ctx->side_ev[ch % 4], 0));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side[3], ctx->side_ev[ch % 4], 0));
      for (int dq = 0; dq < P; ++dq) {
        const int pq = (me + 1 + dq) % P;
        const double* src = cur + nxl * (yst[pq] - 1) + nxl * ny * k0;
        double* dst = (double*)poth->ptr[rankof(pq)] + nxl * yszq[pq] * (long)(zst[me] - 1 + k0);
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)nxl * yszq[pq] * E, src, (size_t)nxl * ny * E, (size_t)nxl * yszq[pq] * E, (size_t)(k1 - k0),
                                        cudaMemcpyDeviceToDevice, pq == me ? ctx->side[3] : ctx->side[ch % 3]));
      }
      ctx->launches += P;
    }
    if ((rc = join_side())) return rc;
    if ((rc = k_barrier(ctx))) return rc;
    // ---- z solve and push, chunk by chunk over my y columns
    const int ncj = std::max(1, std::min(std::min(nchunk_env, 4), (int)nyl / 2));
    for (int ch = 0; ch < ncj; ++ch) {
      int j0 = (int)((long)nyl * ch / ncj), j1 = (int)((long)nyl * (ch + 1) / ncj);
      j0 -= j0 & 1; if (ch + 1 < ncj) j1 -= j1 & 1;          // even boundaries keep the 16-byte alignment of the TMA kernel
      if (j1 <= j0) continue;
      if ((rc = k_gaussel(ctx, zs[0], j1 - j0, zs[2] - q, nxl * nyl, zper, a, b, c, lambdaxy + nxl * j0, oth + nxl * j0))) return rc;
      CUDA_TRY(ctx, cudaEventRecord(ctx->side_ev[ch % 4], ctx->stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side[ch % 3], ctx->side_ev[ch % 4], 0));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side[3], ctx->side_ev[ch % 4], 0));
      for (int dq = 0; dq < P; ++dq) {
        const int qq = (me + 1 + dq) % P;
        const double* src = oth + nxl * j0 + nxl * nyl * (long)(zst[qq] - 1);
        double* dst = (double*)pcur->ptr[rankof(qq)] + nxl * (long)(yst[me] - 1 + j0);
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)nxl * ny * E, src, (size_t)nxl * nyl * E, (size_t)nxl * (j1 - j0) * E, (size_t)zszq[qq],
                                        cudaMemcpyDeviceToDevice, qq == me ? ctx->side[3] : ctx->side[ch % 3]));
      }
      ctx->launches += P;
    }
    if ((rc = join_side())) return rc;
    if ((rc = k_barrier(ctx))) return rc;
    // ---- backward: y transform, [y -> x transpose,] x transform
    if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 1, ys[0], ys[1], ys[2], cur, ys[0], nxl * ny, cur, ys[0], nxl * ny, 1.0))) return rc;
    if (xy) {
      if ((rc = k_transpose_p2p(ctx, 3, cur, poth))) return rc;
      std::swap(cur, oth); std::swap(pcur, poth);
    }
    return k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 1, xs[0], xs[1], xs[2], cur, xs[0], (long)xs[0] * xs[1], p + d.idx(1, 1, 1), d.s1, d.s2, normfft);
  }
  // ---- peer-fused path (process grid 1 x P: x->y is the identity): the forward y pass scatters its spectrum straight
  // into the Z-pencils of the owning ranks and the z solve scatters its solution straight into their Y-pencils, so the
  // two transposes cost no pass over memory at all -- only the NVLink stores inside the producing kernels and one
  // stream-ordered barrier each.
  static const bool nofuse = getenv("CALES_NO_FUSED_TRANSPOSE") != nullptr;
  if (p2p && !nofuse && ctx->dims[0] == 1 && ctx->dims[1] <= 8 && lambdaxy && k_fft_peer_capable(pl.bc[1], pl.c_or_f[1], ys[1]) &&
      (pipe_env == 0 || fused_ok)) {
    const int P = ctx->dims[1], me = ctx->coord[1];
    std::vector<int> yst(P), yen(P), ysz(P), zst(P), zen(P), zsz(P);
    cales_distribute(ctx->ng[1], P, yst.data(), yen.data(), ysz.data());
    cales_distribute(ctx->ng[2], P, zst.data(), zen.data(), zsz.data());
    FftPeerOut FP; GPeer GP;
    ctx->solver_path = 3;
    FP.np = GP.np = P; FP.nx = ys[0]; FP.zoff = zst[me] - 1;
    GP.plane = (long)ys[0] * ys[1]; GP.coff = (long)ys[0] * (yst[me] - 1);
    for (int pr = 0; pr < P; ++pr) {
      const int r = ctx->coord[0] * ctx->dims[1] + pr;
      FP.pbase[pr] = (double*)pb1->ptr[r]; FP.pys[pr] = yst[pr] - 1; FP.pny[pr] = ysz[pr];
      GP.pbase[pr] = (double*)pb0->ptr[r]; GP.pzs[pr] = zst[pr] - 1;
    }
    FP.pys[P] = ctx->ng[1]; GP.pzs[P] = ctx->ng[2];
    if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, xs[0], xs[1], xs[2], p + d.idx(1, 1, 1), d.s1, d.s2, w0, xs[0], (long)xs[0] * xs[1], 1.0))) return rc;
    // bit 0: the forward y pass pushes its spectrum, bit 1: the z solve pushes its solution (asynchronous bulk tensor stores
    // of gauss_tma_k; only when every level is solved, i.e. not for the shortened face-centred Dirichlet system)
    static const int fmask = getenv("CALES_FUSE_MASK") ? atoi(getenv("CALES_FUSE_MASK")) : 3;
    if (fmask & 1) {
      ctx->fft_peer_out = &FP;
      rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, ys[0], ys[1], ys[2], w0, ys[0], (long)ys[0] * ys[1], w1, ys[0], (long)ys[0] * ys[1], 1.0);
      ctx->fft_peer_out = nullptr;
      if (rc) return rc;
      if ((rc = k_barrier(ctx))) return rc;
    } else {
      if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, ys[0], ys[1], ys[2], w0, ys[0], (long)ys[0] * ys[1], w0, ys[0], (long)ys[0] * ys[1], 1.0))) return rc;
      if ((rc = k_transpose_p2p(ctx, 1, w0, pb1))) return rc;
    }
    if ((fmask & 2) && q == 0 && k_gauss_tma_fits(zs[0] * zs[1], zs[2], zper)) {
      ctx->gauss_peer_out = &GP;
      rc = k_gaussel(ctx, zs[0], zs[1], zs[2] - q, (long)zs[0] * zs[1], zper, a, b, c, lambdaxy, w1);
      ctx->gauss_peer_out = nullptr;
      if (rc) return rc;
      if ((rc = k_barrier(ctx))) return rc;
    } else {
      if ((rc = k_gaussel(ctx, zs[0], zs[1], zs[2] - q, (long)zs[0] * zs[1], zper, a, b, c, lambdaxy, w1))) return rc;
      if ((rc = k_transpose_p2p(ctx, 2, w1, pb0))) return rc;
    }
    if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 1, ys[0], ys[1], ys[2], w0, ys[0], (long)ys[0] * ys[1], w0, ys[0], (long)ys[0] * ys[1], 1.0))) return rc;
    if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 1, xs[0], xs[1], xs[2], w0, xs[0], (long)xs[0] * xs[1], p + d.idx(1, 1, 1), d.s1, d.s2, normfft))) return rc;
    return CALES_OK;
  }
  double *cur = w0, *oth = w1, *t_;
  PeerBuf *pcur = pb0, *poth = pb1, *pt_;
  ctx->solver_path = 4;
#define TRANSPOSE(which, P)                                                                       \
  if ((P) > 1) {                                                                                  \
    if ((rc = p2p ? k_transpose_p2p(ctx, which, cur, poth) : k_transpose(ctx, which, cur, oth))) return rc; \
    t_ = cur; cur = oth; oth = t_;                                                                \
    pt_ = pcur; pcur = poth; poth = pt_;                                                          \
  }
  if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 0, xs[0], xs[1], xs[2], p + d.idx(1, 1, 1), d.s1, d.s2, cur, xs[0], (long)xs[0] * xs[1], 1.0))) return rc;
  TRANSPOSE(0, ctx->dims[0])                                                            // x -> y
  if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 0, ys[0], ys[1], ys[2], cur, ys[0], (long)ys[0] * ys[1], cur, ys[0], (long)ys[0] * ys[1], 1.0))) return rc;
  TRANSPOSE(1, ctx->dims[1])                                                            // y -> z
  if ((rc = k_gaussel(ctx, zs[0], zs[1], zs[2] - q, (long)zs[0] * zs[1], zper, a, b, c, lambdaxy, cur))) return rc;
  TRANSPOSE(2, ctx->dims[1])                                                            // z -> y
  if ((rc = k_fft_pass(ctx, 1, pl.bc[1], pl.c_or_f[1], 1, ys[0], ys[1], ys[2], cur, ys[0], (long)ys[0] * ys[1], cur, ys[0], (long)ys[0] * ys[1], 1.0))) return rc;
  TRANSPOSE(3, ctx->dims[0])                                                            // y -> x
  if ((rc = k_fft_pass(ctx, 0, pl.bc[0], pl.c_or_f[0], 1, xs[0], xs[1], xs[2], cur, xs[0], (long)xs[0] * xs[1], p + d.idx(1, 1, 1), d.s1, d.s2, normfft))) return rc;
#undef TRANSPOSE
  (void)ng;
  return CALES_OK;
}

// copy interior <-> halo-free work array
__global__ void strip_k(Dims d, const double* __restrict__ p, double* __restrict__ w, int to_work) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
  if (i >= d.n1) return;
  const long c = d.idx(i + 1, j + 1, k + 1), o = i + (long)d.n1 * (j + (long)d.n2 * k);
  if (to_work) w[o] = p[c]; else ((double*)p)[c] = w[o];
}

extern "C" int cales_solver_gaussel_z(cales_ctx* ctx, const int n[3], const double* a, const double* b, const double* c,
                                      const char bcz[2], const char c_or_f[3], double* p) {
  CHECK_CTX(ctx);
  Dims d(n);
  const int q = (c_or_f[2] == 'f' && bcz[1] == 'D') ? 1 : 0;
  const bool zper = bcz[0] == 'P' && bcz[1] == 'P';
  int rc;
  if (ctx->nranks == 1 || ctx->dims[1] == 1) {
    // z is rank-local: solve directly on the haloed array (plane stride s2, columns = the full (n1+2)x(n2+2) plane
    // would touch ghosts; restrict to interior rows by solving row by row)
    const long p2 = (long)n[0] * n[1];
    double* wk = (double*)cales_scratch(ctx, "solver_wk", (size_t)p2 * n[2] * sizeof(double));
    if (!wk) return CALES_ERR_NOMEM;
    strip_k<<<dim3(cdiv(n[0], 128), n[1], n[2]), 128, 0, ctx->stream>>>(d, p, wk, 1);
    KERNEL_CHECK(ctx);
    if ((rc = k_gaussel(ctx, n[0], n[1], n[2] - q, p2, zper, a, b, c, nullptr, wk))) return rc;
    strip_k<<<dim3(cdiv(n[0], 128), n[1], n[2]), 128, 0, ctx->stream>>>(d, p, wk, 0);
    KERNEL_CHECK(ctx);
    return CALES_OK;
  }
  // z is decomposed: transpose to Z-pencils, solve, transpose back (src/solver.f90:199-231)
  const int* xs = ctx->xsz; const int* zs = ctx->zsz;
  size_t bmax = 0;
  for (int r = 0; r < ctx->nranks; ++r)
    for (int ax = 1; ax <= 3; ++ax) {
      int lo[3], hi[3], sz[3];
      cales_pencil(ctx->ng, ctx->dims, r, ax, lo, hi, sz);
      bmax = std::max(bmax, (size_t)sz[0] * sz[1] * sz[2]);
    }
  PeerBuf* pb0 = k_peer_buffer(ctx, "solver_wk", bmax * sizeof(double));
  PeerBuf* pb1 = pb0 ? k_peer_buffer(ctx, "solver_wk1", bmax * sizeof(double)) : nullptr;
  const bool p2p = pb0 && pb1;
  double* cur = p2p ? (double*)pb0->local : (double*)cales_scratch(ctx, "solver_wk", bmax * sizeof(double));
  double* oth = p2p ? (double*)pb1->local : (double*)cales_scratch(ctx, "solver_wk1", bmax * sizeof(double));
  if (!cur || !oth) return CALES_ERR_NOMEM;
  PeerBuf *pcur = pb0, *poth = pb1;
  if (xs[0] != n[0] || xs[1] != n[1] || xs[2] != n[2]) return cales_fail(ctx, CALES_ERR_INVALID, "solver_gaussel_z: n does not match the X-pencil of this rank");
  strip_k<<<dim3(cdiv(n[0], 128), n[1], n[2]), 128, 0, ctx->stream>>>(d, p, cur, 1);
  KERNEL_CHECK(ctx);
#define TRANSPOSE(which, P)                                                                       \
  if ((P) > 1) {                                                                                  \
    if ((rc = p2p ? k_transpose_p2p(ctx, which, cur, poth) : k_transpose(ctx, which, cur, oth))) return rc; \
    std::swap(cur, oth); std::swap(pcur, poth);                                                   \
  }
  TRANSPOSE(0, ctx->dims[0])
  TRANSPOSE(1, ctx->dims[1])
  if ((rc = k_gaussel(ctx, zs[0], zs[1], zs[2] - q, (long)zs[0] * zs[1], zper, a, b, c, nullptr, cur))) return rc;
  TRANSPOSE(2, ctx->dims[1])
  TRANSPOSE(3, ctx->dims[0])
#undef TRANSPOSE
  strip_k<<<dim3(cdiv(n[0], 128), n[1], n[2]), 128, 0, ctx->stream>>>(d, p, cur, 0);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}
