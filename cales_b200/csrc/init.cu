// Device-side input synthesis: the deterministic initial conditions of
//   initflow  src/initflow.f90:17-283  (profiles 340-434, set_mean 317-338, vortex pair 233-260)
// generated directly in the caller's device arrays, so that the 1024 x 512 x 512 fields never exist on the host.
// The noisy cases ('log', 'hcl', 'tbl': add_noise draws from the Fortran compiler's random_number stream,
// initflow.f90:285-315) stay with the host (cales_b200/hostinit.py).
#include <cmath>

#include "common.cuh"

int k_bulk_mean_dev(cales_ctx* ctx, const int n[3], const double* gvr, const double* p, double* out);

#define IBX 64
#define IBY 4
#define PI 3.14159265358979323846      // acos(-1._rp), param.f90:18

enum { P_ZER = 0, P_UNI, P_COU, P_POI, P_IOP, P_HALF };   // 1-D profiles u1d(k)

// u1d(k): initflow.f90:340-373 (couette, poiseuille) and the inline cases 60-102
__global__ void u1d_k(int kind, int n3, const double* __restrict__ zc, double l3, double uref, double ubulk, double* __restrict__ u1d) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (k > n3) return;
  const double z = zc[k] / l3, zh = zc[k] / (2. * l3);
  double v = 0.;
  if (kind == P_UNI) v = uref;
  else if (kind == P_COU) v = .5 * (1. - 2. * z) * uref;
  else if (kind == P_POI) v = 6. * z * (1. - z) * ubulk;
  else if (kind == P_IOP) v = 6. * z * (1. - z) * ubulk - ubulk;
  else if (kind == P_HALF) v = 6. * zh * (1. - zh) * ubulk;           // 'hcp', 'hdc': the lower half of a mirrored channel
  u1d[k] = v;
}

struct InitArgs {
  int kind;                 // 0: u = u1d(k), v = w = p = 0;  1 'tgv', 2 'tgw', 3 'ant', 4 'duc'
  int lo1, lo2;
  double l1, l2, l3, dl1, dl2, uref;
  const double *zc, *zf, *u1d;
  double *u, *v, *w, *p;
};

__global__ void __launch_bounds__(IBX* IBY) initflow_k(Dims d, InitArgs A) {
  const int i = blockIdx.x * IBX + threadIdx.x + 1, j = blockIdx.y * IBY + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > d.n1 || j > d.n2) return;
  const long c = d.idx(i, j, k);
  double u = 0., v = 0., w = 0., p = 0.;
  if (A.kind == 0) u = A.u1d[k];
  else if (A.kind == 1) {                                             // initflow.f90:103-118
    const double zcc = A.zc[k] / A.l3 * 2. * PI;
    const double yc = (j + A.lo2 - 1 - .5) * A.dl2 / A.l2 * 2. * PI, yf = (j + A.lo2 - 1 - .0) * A.dl2 / A.l2 * 2. * PI;
    const double xc = (i + A.lo1 - 1 - .5) * A.dl1 / A.l1 * 2. * PI, xf = (i + A.lo1 - 1 - .0) * A.dl1 / A.l1 * 2. * PI;
    u = sin(xf) * cos(yc) * cos(zcc) * A.uref;
    v = -cos(xc) * sin(yf) * cos(zcc) * A.uref;
  } else if (A.kind == 2) {                                           // initflow.f90:119-133
    const double yc = (j + A.lo2 - 1 - .5) * A.dl2, yf = (j + A.lo2 - 1 - .0) * A.dl2;
    const double xc = (i + A.lo1 - 1 - .5) * A.dl1, xf = (i + A.lo1 - 1 - .0) * A.dl1;
    u = cos(xf) * sin(yc) * A.uref;
    v = -sin(xc) * cos(yf) * A.uref;
    p = -(cos(2. * xc) + cos(2. * yc)) / 4. * (A.uref * A.uref);
  } else if (A.kind == 3) {                                           // initflow.f90:134-156 (Antuono, JFM 890, A23)
    const double a = 4. * sqrt(2.) / 3. / sqrt(3.);
    const double zcc = A.zc[k] / A.l3 * 2. * PI + 0.5 * PI, zff = A.zf[k] / A.l3 * 2. * PI + 0.5 * PI;
    const double yc = (j + A.lo2 - 1 - .5) * A.dl2 / A.l2 * 2. * PI + 0.5 * PI, yf = (j + A.lo2 - 1 - .0) * A.dl2 / A.l2 * 2. * PI + 0.5 * PI;
    const double xc = (i + A.lo1 - 1 - .5) * A.dl1 / A.l1 * 2. * PI + 0.5 * PI, xf = (i + A.lo1 - 1 - .0) * A.dl1 / A.l1 * 2. * PI + 0.5 * PI;
    u = a * (sin(xf - 5. * PI / 6.) * cos(yc - 1. * PI / 6.) * sin(zcc) - sin(xf - 1. * PI / 6.) * sin(yc) * cos(zcc - 5. * PI / 6.)) * A.uref;
    v = a * (sin(xc) * sin(yf - 5. * PI / 6.) * sin(zcc - 1. * PI / 6.) - cos(xc - 5. * PI / 6.) * sin(yf - 1. * PI / 6.) * sin(zcc)) * A.uref;
    w = a * (cos(xc - 1. * PI / 6.) * sin(yc) * sin(zff - 5. * PI / 6.) - sin(xc) * cos(yc - 5. * PI / 6.) * sin(zff - 1. * PI / 6.)) * A.uref;
    p = -(u * u + v * v + w * w) / 2.;
  } else {                                                            // 'duc', initflow.f90:181-201: 101 terms of the series
    const double ly = .5 * A.l2, lz = .5 * A.l3;
    const double xi = -1. + (j + A.lo2 - 1.5) * A.dl2 / ly, eta = -1. + A.zc[k] / lz;
    double sum = 0.;
    for (int m = 0; m <= 100; ++m) {
      const double cosh_term = cosh((2 * m + 1) * PI * ly / (2 * lz) * xi) / cosh((2 * m + 1) * PI * ly / (2 * lz));
      const double cos_term = cos((2 * m + 1) * PI / 2 * eta);
      const double odd = (double)(2 * m + 1);
      sum = sum + ((m & 1) ? -1. : 1.) / (odd * odd * odd) * cosh_term * cos_term;
    }
    u = .5 * (lz * lz) * (1. - eta * eta - 4. * ((2. / PI) * (2. / PI) * (2. / PI)) * sum);
  }
  A.u[c] = u; A.v[c] = v; A.w[c] = w; A.p[c] = p;
}

// set_mean (initflow.f90:317-338): u = u / meanold * mean when meanold /= 0
__global__ void __launch_bounds__(IBX* IBY) set_mean_k(Dims d, const double* __restrict__ meanold, double mean, double* __restrict__ u) {
  const int i = blockIdx.x * IBX + threadIdx.x + 1, j = blockIdx.y * IBY + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > d.n1 || j > d.n2) return;
  const double m = meanold[0];
  if (m != 0.) { const long c = d.idx(i, j, k); u[c] = u[c] / m * mean; }
}

__global__ void gvr_k(int n3, const double* __restrict__ dzf, double l3, double f, double* __restrict__ gvr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= n3 + 1) gvr[k] = dzf[k] / l3 * f;
}

// the streamwise vortex pair of Henningson & Kim (initflow.f90:233-260): v, w, p of every cell; u untouched
__global__ void __launch_bounds__(IBX* IBY) vortex_pair_k(Dims d, int lo1, int lo2, double l1, double l2, double l3, double dl1, double dl2,
                                                           const double* __restrict__ zc, const double* __restrict__ dzf, double ubulk,
                                                           double* __restrict__ v, double* __restrict__ w, double* __restrict__ p) {
  const int i = blockIdx.x * IBX + threadIdx.x + 1, j = blockIdx.y * IBY + threadIdx.y + 1, k = blockIdx.z + 1;
  if (i > d.n1 || j > d.n2) return;
  const double zcc = 2. * zc[k] / l3 - 1.;
  const double zff = 2. * (zc[k] / l3 + .5 * dzf[k] / l3) - 1.;
  const double yc = ((lo2 - 1 + j - 0.5) * dl2 - .5 * l2) * 2. / l3, yf = ((lo2 - 1 + j - 0.0) * dl2 - .5 * l2) * 2. / l3;
  const double xc = ((lo1 - 1 + i - 0.5) * dl1 - .5 * l1) * 2. / l3;
  const double gxy = xc * exp(-4. * (4. * (yf * yf) + xc * xc));                 // gxy(yf,xc), initflow.f90:424-428
  const double dfz = -4. * zcc * (1. - zcc * zcc);                               // dfz, 414-418
  const double fz = (1. - zff * zff) * (1. - zff * zff);                         // fz, 409-413
  const double dgxy = exp(-4. * (4. * (yc * yc) + xc * xc)) * (1. - 8. * (xc * xc));   // dgxy, 429-433
  const long c = d.idx(i, j, k);
  v[c] = -1. * gxy * dfz * ubulk * 1.5;
  w[c] = 1. * fz * dgxy * ubulk * 1.5;
  p[c] = 0.;
}

extern "C" int cales_initflow(cales_ctx* ctx, const char* inivel, const double bcvel[18], const int ng[3], const int lo[3], const int n[3],
                              const double l[3], const double dl[3], const double* zc, const double* zf, const double* dzc,
                              const double* dzf, double visc, const int is_forced[3], const double velf[3], const double bforce[3],
                              int is_wallturb, double* u, double* v, double* w, double* p) {
  CHECK_CTX(ctx);
  (void)ng; (void)dzc;
  Dims d(n);
  const size_t fb = (size_t)d.size() * sizeof(double);
  double uref = 1., ubulk = is_forced[0] ? velf[0] : 1.;
  bool is_mean = false;
  int kind3d = 0, prof = P_ZER;
  const std::string s(inivel);
  // bcvel(0:1,3,3) in Fortran order: element (ib,idir,ivel) at ib + 2*(idir-1) + 6*(ivel-1)
  const double ubot = bcvel[0 + 2 * 2 + 6 * 0], utop = bcvel[1 + 2 * 2 + 6 * 0];
  if (s == "zer") prof = P_ZER;
  else if (s == "uni") prof = P_UNI;
  else if (s == "cou") { uref = ubot - utop; prof = P_COU; }
  else if (s == "poi") { prof = P_POI; is_mean = true; }
  else if (s == "iop") { ubulk = .5 * fabs(ubot + utop); prof = P_IOP; }          // ('iop' skips set_mean, initflow.f90:229)
  else if (s == "hcp") { prof = P_HALF; is_mean = true; }
  else if (s == "pdc" || s == "hdc") {                                            // initflow.f90:157-180
    double lref = l[2] / 2.;
    if (s != "pdc") lref = 2. * lref;
    if (is_wallturb) {
      uref = sqrt(bforce[0] * lref);
      const double retau = uref * lref / visc, reb = pow(retau / .09, 1. / .88);
      ubulk = reb * visc / (2 * lref);
    } else ubulk = bforce[0] * (lref * lref) / (3. * visc);
    prof = s == "pdc" ? P_POI : P_HALF; is_mean = true;
  } else if (s == "tgv") kind3d = 1;
  else if (s == "tgw") kind3d = 2;
  else if (s == "ant") kind3d = 3;
  else if (s == "duc") { kind3d = 4; is_mean = true; }
  else if (s == "log" || s == "hcl" || s == "tbl")
    return cales_fail(ctx, CALES_ERR_INVALID, "initflow: '%s' adds noise from the host's random stream (add_noise, initflow.f90:285-315): use the host path", inivel);
  else return cales_fail(ctx, CALES_ERR_INVALID, "invalid name for initial velocity field: '%s'", inivel);   // initflow.f90:202-209
  double* scr = (double*)cales_scratch(ctx, "init_1d", (size_t)2 * (n[2] + 2) * sizeof(double));
  if (!scr) return CALES_ERR_NOMEM;
  double* u1d = scr; double* gvr = scr + n[2] + 2;
  // ghost cells are not part of the initial condition (the caller fills them with bounduvw/boundp, main.f90:370-372)
  CUDA_TRY(ctx, cudaMemsetAsync(u, 0, fb, ctx->stream)); CUDA_TRY(ctx, cudaMemsetAsync(v, 0, fb, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(w, 0, fb, ctx->stream)); CUDA_TRY(ctx, cudaMemsetAsync(p, 0, fb, ctx->stream));
  if (!kind3d) {
    u1d_k<<<cdiv(n[2], 128), 128, 0, ctx->stream>>>(prof, n[2], zc, l[2], uref, ubulk, u1d);
    KERNEL_CHECK(ctx);
  }
  InitArgs A;
  A.kind = kind3d; A.lo1 = lo[0]; A.lo2 = lo[1]; A.l1 = l[0]; A.l2 = l[1]; A.l3 = l[2]; A.dl1 = dl[0]; A.dl2 = dl[1]; A.uref = uref;
  A.zc = zc; A.zf = zf; A.u1d = u1d; A.u = u; A.v = v; A.w = w; A.p = p;
  const dim3 g(cdiv(n[0], IBX), cdiv(n[1], IBY), n[2]), b(IBX, IBY);
  initflow_k<<<g, b, 0, ctx->stream>>>(d, A);
  KERNEL_CHECK(ctx);
  if (is_mean) {                                                       // set_mean, initflow.f90:228-232, 317-338
    gvr_k<<<cdiv(n[2] + 2, 128), 128, 0, ctx->stream>>>(n[2], dzf, l[2], (dl[0] / l[0]) * (dl[1] / l[1]), gvr);
    KERNEL_CHECK(ctx);
    double* mean = ctx->red + 16;
    int rc = k_bulk_mean_dev(ctx, n, gvr, u, mean);
    if (rc) return rc;
    set_mean_k<<<g, b, 0, ctx->stream>>>(d, mean, ubulk, u);
    KERNEL_CHECK(ctx);
  }
  if (is_wallturb) {
    vortex_pair_k<<<g, b, 0, ctx->stream>>>(d, lo[0], lo[1], l[0], l[1], l[2], dl[0], dl[1], zc, dzf, ubulk, v, w, p);
    KERNEL_CHECK(ctx);
  }
  return CALES_OK;
}
