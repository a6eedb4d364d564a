/* oracle/c/cales_cpu.c -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never on the product path).
 *
 * Plain C / OpenMP restatement of the reference's per-RK3-substep path, explicit diffusion, static or dynamic Smagorinsky, one
 * rank, for (i) the tri-periodic configuration (BASELINE config 2, the bench.py workload) and (ii) the plane channel:
 * periodic x and y, no-slip walls in z (velocity D, pressure N, eddy viscosity D), bulk-velocity forcing, stretched
 * grids, van Driest damping, optionally the log-law wall model on the z walls, and the dynamic Smagorinsky model
 * (sgs.f90:153-380 with filter3d 616-680, extrapolate 682-767, cmpt_alph2 769-822, interpolate 850-870, ave1d_channel
 * 433-538) for both (BASELINE configs 1, 3 and 5); (iii) static Smagorinsky with walls in x and / or y as well -- square
 * duct, lid-driven cavity (BASELINE config 4): the general set_bc sequence of bounduvw / boundp, van Driest distance over all
 * walls, the wall model on y and z walls (cmpt_wallmodelbc case(2) and case(3)), and the cell-centred Neumann-Neumann
 * transforms REDFT10 / REDFT01 (fft.f90:192-245) through the same complex FFT:
 *   src/bound.f90:18-154 (bounduvw), 156-200 (boundp), 202-399 (set_bc: P, D and N, centred and face),
 *   src/wmodel.f90:19-335 (updt_wallmodelbc / cmpt_wallmodelbc case(3) / vel_relative / wallmodel),
 *   src/sgs.f90:69-152 with extrapolate 682-767, src/rk.f90:197-222 + src/utils.f90:16-47 + src/mom.f90:311-335
 *   (bulk forcing), src/initsolver.f90:127-169 (tridmatrix with the N fold-in), src/solver.f90:82-107 (gaussel).
 * The loops of
 *   src/mom.f90:17-309 (mom_xyz_ad), src/rk.f90:17-121 (rk), src/fillps.f90:14-48, src/solver.f90:20-80
 *   (solver: x transform, y transform, gaussel_periodic 109-151 with dgtsv_homebrewed 153-179, backward),
 *   src/correc.f90:14-68, src/updatep.f90:14-49, src/sgs.f90:69-152 ('smag') + strain_rate 1019-1110,
 *   src/bound.f90:18-200 (periodic ghost fill only), src/chkdt.f90:17-99, src/chkdiv.f90:16-52,
 * sequenced as src/main.f90:417-507.  It exists (i) as a second, independently written restatement that
 * tests/test_oracle_c.py holds against the numpy oracle, and (ii) as the multi-threaded CPU arm of
 * bench.py (`cpu_baseline`, `--impl reference`): the Fortran/MPI/FFTW reference cannot be built here.
 *
 * PARITY UNPINNED (see oracle/__init__.py): the reference ships no golden vectors for this path.
 *
 * Third-party arithmetic: the reference plans FFTW3 R2HC/HC2R (unpinned libfftw3-dev; call sites
 * src/fft.f90:83-84,120-121).  FFTW is absent; the transforms below are an own Stockham mixed-radix
 * complex FFT, two real lines per complex transform, producing FFTW's halfcomplex layout
 * (r0..r_{n/2}, i_{(n+1)/2-1}..i_1), unnormalised, forward sign -1.
 *
 * Arrays are Fortran ordered with one ghost cell: (i,j,k) -> i + (n1+2)*(j + (n2+2)*k).
 * Build with -ffp-contract=off: the stencil loops then agree bit for bit with the numpy oracle.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define C_SMAG 0.11 /* param.f90:33 */
static const double RKCOEFF[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}}; /* param.f90:27-29 */

typedef struct { double re, im; } cpx;

typedef struct {
  int n, nfac, fac[32];
  cpx *tw; /* tw[k] = exp(-2 pi i k / n) */
  cpx *tq; /* tq[k] = exp(-i pi k / (2 n)): the quarter-wave factors of the cosine transforms */
} fftplan;

typedef struct {
  int n1, n2, n3;
  long sj, sk, ntot, nint;
  double l[3], dl[3], dli[3], visc;
  double *dzc, *dzf, *dzci, *dzfi;     /* 0:n3+1 */
  double *u, *v, *w, *p, *pp, *visct, *s0;
  double *rhs[3], *rhso[3];            /* dudtrk.., dudtrko.. (n1,n2,n3), rk.f90:36-72 */
  double *a, *b, *c, *lambdaxy, normfft;
  double *wk;                          /* solver work array (n1,n2,n3) */
  fftplan px, py;
  double eps;
  /* channel family (zwall = 1): z walls, forcing, wall model */
  int zwall;                           /* 0: z periodic; 1: walls at both z faces (cbcvel D, cbcpre N, cbcsgs D) */
  int lwm[2], index_wm[2];             /* wall model on the lower / upper z wall (lwm(0:1,3)), interpolation level k2 */
  double hwm;
  int is_forced[3];
  double velf[3], bforce[3], f[3];
  double *zc, *zf, *gvr_c, *gvr_f;     /* 0:n3+1 */
  double *bcu_z, *bcv_z;               /* wall-model Neumann planes bcu%z, bcv%z: (0:n1+1, 0:n2+1, 0:1) */
  double *wku, *wkv;                   /* copies of u, v for the extrapolation of cmpt_sgs (sgs.f90:84-90) */
  int dsmag, sgs_done;                 /* 0: 'smag', 1: 'dsmag', 2: 'none' */
  char kx, ky;                         /* transform kind of x and y: 'P' (R2HC/HC2R) or 'N' (REDFT10/REDFT01) */
  int gen;                             /* 1: walls in x and / or y as well (duct, cavity): the general ghost fills below */
  char cbcvel[2][3][3], cbcpre[2][3], cbcsgs[2][3];   /* [ib][idir][ivel] */
  double bcvel[2][3][3];
  int lwm_y[2], index_wm_y[2];         /* wall model on the y walls (lwm(0:1,2)): general mode only */
  double *bcu_y, *bcw_y, *wkw;         /* bcu%y, bcw%y: (0:n1+1, 0:n3+1, 0:1); copy of w for the extrapolation of cmpt_sgs */
  double *dyn[24];                     /* work arrays of the dynamic model (sgs.f90:156-166): uc,vc,wc, uf,vf,wf, wk(6), sij(6), mij(6) */
} cpu_t;

#define IDX(s, i, j, k) ((long)(i) + (s)->sj * ((long)(j) + (long)((s)->n2 + 2) * (long)(k)))

/* ------------------------------------------------------------------------------------------------ FFT */
static void plan_init(fftplan *pl, int n) {
  pl->n = n; pl->nfac = 0;
  int m = n;
  while (m % 4 == 0) { pl->fac[pl->nfac++] = 4; m /= 4; }
  for (int r = 2; m > 1;) { if (m % r == 0) { pl->fac[pl->nfac++] = r; m /= r; } else r++; }
  pl->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  const double pi = acos(-1.0);
  for (int k = 0; k < n; k++) { pl->tw[k].re = cos(2.0 * pi * k / n); pl->tw[k].im = -sin(2.0 * pi * k / n); }
  pl->tq = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  for (int k = 0; k < n; k++) { pl->tq[k].re = cos(pi * k / (2.0 * n)); pl->tq[k].im = -sin(pi * k / (2.0 * n)); }
}

/* Stockham autosort, decimation in frequency; sign=-1 forward, +1 backward (conjugated twiddles).
 * Returns the buffer (x or y) that holds the result. */
static cpx *fft_exec(const fftplan *pl, cpx *x, cpx *y, int sign) {
  int n = pl->n, s = 1;
  for (int f = 0; f < pl->nfac; f++) {
    int r = pl->fac[f], m = n / r, tstep = pl->n / n; /* w_n^p = tw[p*tstep] */
    if (r == 2) {
      for (int p = 0; p < m; p++) {
        cpx w = pl->tw[p * tstep]; if (sign > 0) w.im = -w.im;
        for (int q = 0; q < s; q++) {
          cpx a = x[q + s * p], b = x[q + s * (p + m)];
          cpx d = {a.re - b.re, a.im - b.im};
          y[q + s * (2 * p)].re = a.re + b.re; y[q + s * (2 * p)].im = a.im + b.im;
          y[q + s * (2 * p + 1)].re = d.re * w.re - d.im * w.im; y[q + s * (2 * p + 1)].im = d.re * w.im + d.im * w.re;
        }
      }
    } else if (r == 4) {
      for (int p = 0; p < m; p++) {
        cpx w1 = pl->tw[p * tstep], w2 = pl->tw[2 * p * tstep], w3 = pl->tw[3 * p * tstep];
        if (sign > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
        for (int q = 0; q < s; q++) {
          cpx a = x[q + s * p], b = x[q + s * (p + m)], c = x[q + s * (p + 2 * m)], d = x[q + s * (p + 3 * m)];
          cpx apc = {a.re + c.re, a.im + c.im}, amc = {a.re - c.re, a.im - c.im};
          cpx bpd = {b.re + d.re, b.im + d.im}, bmd = {b.re - d.re, b.im - d.im};
          /* multiply (b-d) by -i (forward) or +i (backward) */
          cpx jb = sign < 0 ? (cpx){bmd.im, -bmd.re} : (cpx){-bmd.im, bmd.re};
          cpx t0 = {apc.re + bpd.re, apc.im + bpd.im}, t2 = {apc.re - bpd.re, apc.im - bpd.im};
          cpx t1 = {amc.re + jb.re, amc.im + jb.im}, t3 = {amc.re - jb.re, amc.im - jb.im};
          cpx *o = &y[q + s * (4 * p)];
          o[0] = t0;
          o[s].re = t1.re * w1.re - t1.im * w1.im; o[s].im = t1.re * w1.im + t1.im * w1.re;
          o[2 * s].re = t2.re * w2.re - t2.im * w2.im; o[2 * s].im = t2.re * w2.im + t2.im * w2.re;
          o[3 * s].re = t3.re * w3.re - t3.im * w3.im; o[3 * s].im = t3.re * w3.im + t3.im * w3.re;
        }
      }
    } else { /* generic radix */
      int rstep = pl->n / r; /* W_r^t = tw[t*rstep] */
      for (int p = 0; p < m; p++)
        for (int q = 0; q < s; q++)
          for (int uo = 0; uo < r; uo++) {
            double sr = 0., si = 0.;
            for (int t = 0; t < r; t++) {
              cpx a = x[q + s * (p + t * m)], w = pl->tw[((long)t * uo % r) * rstep];
              if (sign > 0) w.im = -w.im;
              sr += a.re * w.re - a.im * w.im; si += a.re * w.im + a.im * w.re;
            }
            cpx w = pl->tw[((long)p * uo * tstep) % pl->n]; if (sign > 0) w.im = -w.im;
            y[q + s * (r * p + uo)].re = sr * w.re - si * w.im; y[q + s * (r * p + uo)].im = sr * w.im + si * w.re;
          }
    }
    cpx *t = x; x = y; y = t;
    n = m; s *= r;
  }
  return x;
}

/* Two real lines (a,b; element stride `st`) -> halfcomplex in place (FFTW R2HC). b may be NULL. */
static void r2hc_pair(const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  int n = pl->n;
  for (int j = 0; j < n; j++) { x[j].re = a[j * st]; x[j].im = b ? b[j * st] : 0.0; }
  cpx *z = fft_exec(pl, x, y, -1);
  a[0] = z[0].re; if (b) b[0] = z[0].im;
  for (int k = 1; k <= n / 2; k++) {
    cpx zk = z[k], zc = z[n - k];
    double are = 0.5 * (zk.re + zc.re), aim = 0.5 * (zk.im - zc.im);
    double bre = 0.5 * (zk.im + zc.im), bim = -0.5 * (zk.re - zc.re);
    a[k * st] = are; if (2 * k < n) a[(n - k) * st] = aim;
    if (b) { b[k * st] = bre; if (2 * k < n) b[(n - k) * st] = bim; }
  }
}

/* Halfcomplex -> real, unnormalised (FFTW HC2R), two lines at once. */
static void hc2r_pair(const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  int n = pl->n;
  x[0].re = a[0]; x[0].im = b ? b[0] : 0.0;
  for (int k = 1; k <= n / 2; k++) {
    double are = a[k * st], aim = (2 * k < n) ? a[(n - k) * st] : 0.0;
    double bre = b ? b[k * st] : 0.0, bim = (b && 2 * k < n) ? b[(n - k) * st] : 0.0;
    /* Z[k] = A[k] + i B[k];  Z[n-k] = conj(A[k]) + i conj(B[k]) */
    x[k].re = are - bim; x[k].im = aim + bre;
    x[n - k].re = are + bim; x[n - k].im = -aim + bre;
  }
  cpx *z = fft_exec(pl, x, y, +1);
  for (int j = 0; j < n; j++) { a[j * st] = z[j].re; if (b) b[j * st] = z[j].im; }
}

/* FFTW REDFT10 (DCT-II, Y_k = 2 sum_j X_j cos(pi (j + 1/2) k / n)) of two real lines in place, through ONE complex transform
 * of length n (Makhoul's even/odd re-ordering; the pair is separated as in r2hc_pair).  fft.f90:221-244 plans these kinds
 * for Neumann-Neumann, cell-centred directions. */
static void redft10_pair(const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  const int n = pl->n, h = (n + 1) / 2;
  for (int j = 0; j < h; j++) { x[j].re = a[(long)(2 * j) * st]; x[j].im = b ? b[(long)(2 * j) * st] : 0.0; }
  for (int j = 0; j < n / 2; j++) { x[n - 1 - j].re = a[(long)(2 * j + 1) * st]; x[n - 1 - j].im = b ? b[(long)(2 * j + 1) * st] : 0.0; }
  cpx *z = fft_exec(pl, x, y, -1);
  for (int k = 0; k < n; k++) {
    const cpx zk = z[k], zc = z[(n - k) % n], t = pl->tq[k];
    const double are = 0.5 * (zk.re + zc.re), aim = 0.5 * (zk.im - zc.im);
    const double bre = 0.5 * (zk.im + zc.im), bim = -0.5 * (zk.re - zc.re);
    a[(long)k * st] = 2. * (are * t.re - aim * t.im);
    if (b) b[(long)k * st] = 2. * (bre * t.re - bim * t.im);
  }
}

/* FFTW REDFT01 (DCT-III, Y_j = X_0 + 2 sum_{k>=1} X_k cos(pi k (j + 1/2) / n)), the unnormalised inverse of REDFT10 (x 2n):
 * W_k = e^{i pi k / 2n} (X_k - i X_{n-k}), X_n = 0, is Hermitian, so its backward transform is real and two lines ride
 * one complex transform. */
static void redft01_pair(const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  const int n = pl->n, h = (n + 1) / 2;
  for (int k = 0; k < n; k++) {
    const double xa = a[(long)k * st], xan = k ? a[(long)(n - k) * st] : 0.0;
    const double xb = b ? b[(long)k * st] : 0.0, xbn = (b && k) ? b[(long)(n - k) * st] : 0.0;
    const double cr = pl->tq[k].re, ci = -pl->tq[k].im;
    const double war = cr * xa + ci * xan, wai = ci * xa - cr * xan;
    const double wbr = cr * xb + ci * xbn, wbi = ci * xb - cr * xbn;
    x[k].re = war - wbi; x[k].im = wai + wbr;
  }
  cpx *z = fft_exec(pl, x, y, +1);
  for (int j = 0; j < h; j++) { a[(long)(2 * j) * st] = z[j].re; if (b) b[(long)(2 * j) * st] = z[j].im; }
  for (int j = 0; j < n / 2; j++) { a[(long)(2 * j + 1) * st] = z[n - 1 - j].re; if (b) b[(long)(2 * j + 1) * st] = z[n - 1 - j].im; }
}

/* forward / backward transform of a pair of lines by the BC pair of the direction: 'P' R2HC / HC2R, 'N' REDFT10 / REDFT01 */
static void fwd_pair(char kind, const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  if (kind == 'N') redft10_pair(pl, a, b, st, x, y); else r2hc_pair(pl, a, b, st, x, y);
}
static void bwd_pair(char kind, const fftplan *pl, double *a, double *b, long st, cpx *x, cpx *y) {
  if (kind == 'N') redft01_pair(pl, a, b, st, x, y); else hc2r_pair(pl, a, b, st, x, y);
}

/* ------------------------------------------------------------------------------------------ ghost fill */
/* bound.f90:175-199 / 42-46 with all-periodic BCs on one rank: direction by direction, full extent. */
static void bound_periodic(const cpu_t *s, double *p) {
  int n1 = s->n1, n2 = s->n2, n3 = s->n3;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int j = 0; j <= n2 + 1; j++) { p[IDX(s, 0, j, k)] = p[IDX(s, n1, j, k)]; p[IDX(s, n1 + 1, j, k)] = p[IDX(s, 1, j, k)]; }
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int i = 0; i <= n1 + 1; i++) { p[IDX(s, i, 0, k)] = p[IDX(s, i, n2, k)]; p[IDX(s, i, n2 + 1, k)] = p[IDX(s, i, 1, k)]; }
#pragma omp parallel for schedule(static)
  for (int j = 0; j <= n2 + 1; j++)
    for (int i = 0; i <= n1 + 1; i++) { p[IDX(s, i, j, 0)] = p[IDX(s, i, j, n3)]; p[IDX(s, i, j, n3 + 1)] = p[IDX(s, i, j, 1)]; }
}

/* x and y periodic fills, full extent (x: set_bc('P'), bound.f90:232-248; y: the self halo exchange of one rank) */
static void bound_xy_periodic(const cpu_t *s, double *p) {
  int n1 = s->n1, n2 = s->n2, n3 = s->n3;
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int i = 0; i <= n1 + 1; i++) { p[IDX(s, i, 0, k)] = p[IDX(s, i, n2, k)]; p[IDX(s, i, n2 + 1, k)] = p[IDX(s, i, 1, k)]; }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int j = 0; j <= n2 + 1; j++) { p[IDX(s, 0, j, k)] = p[IDX(s, n1, j, k)]; p[IDX(s, n1 + 1, j, k)] = p[IDX(s, 1, j, k)]; }
}

/* set_bc in z on whole planes (bound.f90:202-399): ctype 'D' / 'N', centred or face, boundary value planes bc0 / bc1
 * (NULL = the constant 0 of the channel decks), dr = dzc(0) / dzc(n3) (dzf for the face-centred w, unused by D). */
static void set_bc_z(const cpu_t *s, char ctype, int centered, const double *bc0, const double *bc1, double dr0, double dr1, double *p,
                     int do0, int do1) {
  const int n1 = s->n1, n2 = s->n2, n = s->n3;
  const long pl = (long)(n1 + 2) * (n2 + 2);
  const double sgn = (ctype == 'D' && centered) ? -1. : 1.;
#pragma omp parallel for schedule(static)
  for (int j = 0; j <= n2 + 1; j++)
    for (int i = 0; i <= n1 + 1; i++) {
      const long q = (long)i + (long)(n1 + 2) * j;
      const double b0 = bc0 ? bc0[q] : 0., b1 = bc1 ? bc1[q] : 0.;
      if (ctype == 'D' && centered) {                                            /* 250-282 */
        if (do0) p[q] = 2. * b0 + sgn * p[q + pl];
        if (do1) p[q + pl * (n + 1)] = 2. * b1 + sgn * p[q + pl * n];
      } else if (ctype == 'D') {                                                 /* 283-319 */
        if (do0) p[q] = b0;
        if (do1) { p[q + pl * (n + 1)] = p[q + pl * (n - 1)]; p[q + pl * n] = b1; }
      } else if (centered) {                                                     /* 'N', 320-353 */
        if (do0) p[q] = -dr0 * b0 + sgn * p[q + pl];
        if (do1) p[q + pl * (n + 1)] = dr1 * b1 + sgn * p[q + pl * n];
      }
    }
}

/* set_bc (bound.f90:202-399) for a face of any direction with a CONSTANT boundary value (all the decks covered here), on
 * the whole plane, ghost rows of the other directions included */
static void set_bc_c(const cpu_t *s, char ctype, int ibound, int idir, int centered, double bc, double dr, double *p) {
  const int nn[3] = {s->n1, s->n2, s->n3};
  const long st[3] = {1, s->sj, s->sk};
  const int n = nn[idir], d1 = (idir + 1) % 3, d2 = (idir + 2) % 3;
  const long sd = st[idir];
  const double sgn = (ctype == 'D' && centered) ? -1. : 1.;
#pragma omp parallel for schedule(static)
  for (int b = 0; b <= nn[d2] + 1; b++)
    for (int a = 0; a <= nn[d1] + 1; a++) {
      double *q = p + (long)a * st[d1] + (long)b * st[d2];                       /* q[m * sd] = p(.., m, ..) */
      if (ctype == 'P') {                                                        /* 232-248 */
        q[0] = q[(long)n * sd]; q[(long)(n + 1) * sd] = q[sd];
      } else if (ctype == 'D' && centered) {                                     /* 250-282 */
        if (ibound == 0) q[0] = 2. * bc + sgn * q[sd];
        else q[(long)(n + 1) * sd] = 2. * bc + sgn * q[(long)n * sd];
      } else if (ctype == 'D') {                                                 /* 283-319 */
        if (ibound == 0) q[0] = bc;
        else { q[(long)(n + 1) * sd] = q[(long)(n - 1) * sd]; q[(long)n * sd] = bc; }
      } else if (ctype == 'N' && centered) {                                     /* 320-353 */
        if (ibound == 0) q[0] = -dr * bc + sgn * q[sd];
        else q[(long)(n + 1) * sd] = dr * bc + sgn * q[(long)n * sd];
      } else if (ctype == 'N') {                                                 /* 354-396 */
        if (ibound == 0) q[0] = -dr * bc + q[sd];
        else { q[(long)(n + 1) * sd] = q[(long)n * sd]; q[(long)n * sd] = dr * bc + q[(long)(n - 1) * sd]; }
      }
    }
}

/* halo exchange of one rank (bound.f90:619-696): a decomposed direction (y, z) that is periodic is its own neighbour */
static void halo_self(const cpu_t *s, int idir, double *p) { set_bc_c(s, 'P', 0, idir, 1, 0., 0., p); }

static void cmpt_wallmodelbc_y(cpu_t *s, int ibound);
static void cmpt_wallmodelbc_z(cpu_t *s, int ibound);
static void set_bc_y_plane(const cpu_t *s, int ibound, const double *bc, double dr, double *p);

/* bounduvw (bound.f90:18-154) in general: walls in any direction; wall model on y and / or z walls */
static void bounduvw_gen(cpu_t *s, double *u, double *v, double *w, int is_correc) {
  double *f[3] = {u, v, w};
  for (int idir = 1; idir < 3; idir++)                                           /* updthalo, 42-46: x is the pencil direction */
    if (s->cbcpre[0][idir] == 'P') for (int m = 0; m < 3; m++) halo_self(s, idir, f[m]);
  const int n3 = s->n3;
  for (int idir = 0; idir < 3; idir++) {
    const int is_bound = idir == 0 || s->cbcpre[0][idir] != 'P';                 /* initmpi.f90:169-176 */
    if (!is_bound) continue;
    const int pp = s->cbcvel[0][idir][idir] == 'P' && s->cbcvel[1][idir][idir] == 'P';
    const int impose_norm_bc = !is_correc || pp;
    for (int ib = 0; ib < 2; ib++) {
      const double drf = idir == 0 ? s->dl[0] : idir == 1 ? s->dl[1] : (ib == 0 ? s->dzf[0] : s->dzf[n3]);
      const double drc = idir == 0 ? s->dl[0] : idir == 1 ? s->dl[1] : (ib == 0 ? s->dzc[0] : s->dzc[n3]);
      if (impose_norm_bc) set_bc_c(s, s->cbcvel[ib][idir][idir], ib, idir, 0, s->bcvel[ib][idir][idir], drf, f[idir]);
      const int wm = idir == 1 ? s->lwm_y[ib] : idir == 2 ? s->lwm[ib] : 0;
      if (wm != 0) continue;                                                     /* lwm /= 0: after the wall model, below */
      for (int m = 0; m < 3; m++)                                                /* the two wall-parallel components, in index order */
        if (m != idir) set_bc_c(s, s->cbcvel[ib][idir][m], ib, idir, 1, s->bcvel[ib][idir][m], drc, f[m]);
    }
  }
  if (!(s->lwm_y[0] || s->lwm_y[1] || s->lwm[0] || s->lwm[1])) return;
  /* updt_wallmodelbc (wmodel.f90:19-62; is_updt_wm = .true. on this path), then the Neumann ghosts of the wall-model faces */
  for (int ib = 0; ib < 2; ib++) if (s->lwm_y[ib]) cmpt_wallmodelbc_y(s, ib);
  for (int ib = 0; ib < 2; ib++) if (s->lwm[ib]) cmpt_wallmodelbc_z(s, ib);
  const long ply = (long)(s->n1 + 2) * (s->n3 + 2), plz = (long)(s->n1 + 2) * (s->n2 + 2);
  for (int ib = 0; ib < 2; ib++)
    if (s->lwm_y[ib]) { set_bc_y_plane(s, ib, s->bcu_y + ply * ib, s->dl[1], u); set_bc_y_plane(s, ib, s->bcw_y + ply * ib, s->dl[1], w); }
  for (int ib = 0; ib < 2; ib++)
    if (s->lwm[ib]) {
      set_bc_z(s, 'N', 1, s->bcu_z, s->bcu_z + plz, s->dzc[0], s->dzc[n3], u, ib == 0, ib == 1);
      set_bc_z(s, 'N', 1, s->bcv_z, s->bcv_z + plz, s->dzc[0], s->dzc[n3], v, ib == 0, ib == 1);
    }
}

/* boundp (bound.f90:156-200) in general, boundary values 0 */
static void boundp_gen(const cpu_t *s, const char cbc[2][3], double *p) {
  for (int idir = 1; idir < 3; idir++) if (cbc[0][idir] == 'P') halo_self(s, idir, p);
  const int n3 = s->n3;
  for (int idir = 0; idir < 3; idir++) {
    const int is_bound = idir == 0 || s->cbcpre[0][idir] != 'P';
    if (!is_bound) continue;
    for (int ib = 0; ib < 2; ib++)
      set_bc_c(s, cbc[ib][idir], ib, idir, 1, 0., idir == 0 ? s->dl[0] : idir == 1 ? s->dl[1] : (ib == 0 ? s->dzc[0] : s->dzc[n3]), p);
  }
}

/* wmodel.f90:288-335, WM_LOG */
static void wallmodel_log(double uh, double vh, double h, double visc, double eps, double tauw[2]) {
  const double kap_log = 0.41, b_log = 5.20;                                     /* param.f90:31-32 */
  double conv = 1.;
  const double upar = sqrt(uh * uh + vh * vh);
  double utau = fmax(sqrt(upar / h * visc), visc / h * exp(-kap_log * b_log));
  while (conv > 0.5e-4) {
    const double utau_old = utau;
    const double f = upar / utau - 1. / kap_log * log(h * utau / visc) - b_log;
    const double fp = -1. / utau * (upar / utau + 1. / kap_log);
    utau = fabs(utau - f / fp);
    conv = fabs(utau / utau_old - 1.);
  }
  const double tauw_tot = utau * utau;
  tauw[0] = tauw_tot * uh / (upar + eps);
  tauw[1] = tauw_tot * vh / (upar + eps);
}

/* cmpt_wallmodelbc, case(3) (wmodel.f90:215-271), wall at rest (bcu_mag = bcv_mag = 0: the channel decks) */
static void cmpt_wallmodelbc_z(cpu_t *s, int ibound) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj;
  const double h = s->hwm, visc = s->visc, visci = 1. / visc;
  const double *u = s->u, *v = s->v;
  int k1, k2; double coef, sgn;
  if (ibound == 0) { k2 = s->index_wm[0]; k1 = k2 - 1; coef = (h - s->zc[k1]) / s->dzc[k1]; sgn = 1.; }
  else { k2 = s->index_wm[1]; k1 = k2 + 1; coef = (h - (s->l[2] - s->zc[k1])) / (s->dzc[k2]); sgn = -1.; }
  double *bcu = s->bcu_z + (long)ibound * (n1 + 2) * (n2 + 2), *bcv = s->bcv_z + (long)ibound * (n1 + 2) * (n2 + 2);
  (void)n3;
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= n2; j++)
    for (int i = 0; i <= n1; i++) {
      const long c1 = IDX(s, i, j, k1), c2 = IDX(s, i, j, k2);
      const double u1 = u[c1], u2 = u[c2];
      const double v1 = 0.25 * (v[c1] + v[c1 + 1] + v[c1 - sj] + v[c1 + 1 - sj]);
      const double v2 = 0.25 * (v[c2] + v[c2 + 1] + v[c2 - sj] + v[c2 + 1 - sj]);
      const double u_mag = 0., v_mag = 0.25 * (0. + 0. + 0. + 0.);
      double uh = (1. - coef) * u1 + coef * u2; uh = uh - u_mag;                  /* vel_relative, 273-286 */
      double vh = (1. - coef) * v1 + coef * v2; vh = vh - v_mag;
      double tauw[2];
      wallmodel_log(uh, vh, h, visc, s->eps, tauw);
      bcu[(long)i + (long)(n1 + 2) * j] = sgn * visci * tauw[0];
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j <= n2; j++)
    for (int i = 1; i <= n1; i++) {
      const long c1 = IDX(s, i, j, k1), c2 = IDX(s, i, j, k2);
      const double u1 = 0.25 * (u[c1 - 1] + u[c1] + u[c1 - 1 + sj] + u[c1 + sj]);
      const double u2 = 0.25 * (u[c2 - 1] + u[c2] + u[c2 - 1 + sj] + u[c2 + sj]);
      const double v1 = v[c1], v2 = v[c2];
      const double u_mag = 0.25 * (0. + 0. + 0. + 0.), v_mag = 0.;
      double uh = (1. - coef) * u1 + coef * u2; uh = uh - u_mag;
      double vh = (1. - coef) * v1 + coef * v2; vh = vh - v_mag;
      double tauw[2];
      wallmodel_log(uh, vh, h, visc, s->eps, tauw);
      bcv[(long)i + (long)(n1 + 2) * j] = sgn * visci * tauw[1];
    }
}

/* cmpt_wallmodelbc, case(2) (wmodel.f90:171-214), walls at rest */
static void cmpt_wallmodelbc_y(cpu_t *s, int ibound) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sk = s->sk;
  const double h = s->hwm, visc = s->visc, visci = 1. / visc;
  const double *u = s->u, *w = s->w;
  int j1, j2; double coef, sgn;
  if (ibound == 0) { j2 = s->index_wm_y[0]; j1 = j2 - 1; coef = (h - (j1 - 0.5) * s->dl[1]) / s->dl[1]; sgn = 1.; }
  else { j2 = s->index_wm_y[1]; j1 = j2 + 1; coef = (h - (n2 - j1 + 0.5) * s->dl[1]) / s->dl[1]; sgn = -1.; }
  const long pl = (long)(n1 + 2) * (n3 + 2);
  double *bcu = s->bcu_y + pl * ibound, *bcw = s->bcw_y + pl * ibound;
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int i = 0; i <= n1; i++) {
      const long c1 = IDX(s, i, j1, k), c2 = IDX(s, i, j2, k);
      const double u1 = u[c1], u2 = u[c2];
      const double w1 = 0.25 * (w[c1] + w[c1 + 1] + w[c1 - sk] + w[c1 + 1 - sk]);
      const double w2 = 0.25 * (w[c2] + w[c2 + 1] + w[c2 - sk] + w[c2 + 1 - sk]);
      const double u_mag = 0., w_mag = 0.25 * (0. + 0. + 0. + 0.);
      double uh = (1. - coef) * u1 + coef * u2; uh = uh - u_mag;
      double wh = (1. - coef) * w1 + coef * w2; wh = wh - w_mag;
      double tauw[2];
      wallmodel_log(uh, wh, h, visc, s->eps, tauw);
      bcu[(long)i + (long)(n1 + 2) * k] = sgn * visci * tauw[0];
    }
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= n3; k++)
    for (int i = 1; i <= n1; i++) {
      const long c1 = IDX(s, i, j1, k), c2 = IDX(s, i, j2, k);
      const double wei = (s->zf[k] - s->zc[k]) / s->dzc[k];
      const double u1 = 0.5 * ((1. - wei) * (u[c1 - 1] + u[c1]) + wei * (u[c1 - 1 + sk] + u[c1 + sk]));
      const double u2 = 0.5 * ((1. - wei) * (u[c2 - 1] + u[c2]) + wei * (u[c2 - 1 + sk] + u[c2 + sk]));
      const double w1 = w[c1], w2 = w[c2];
      const double u_mag = 0.5 * ((1. - wei) * (0. + 0.) + wei * (0. + 0.)), w_mag = 0.;
      double uh = (1. - coef) * u1 + coef * u2; uh = uh - u_mag;
      double wh = (1. - coef) * w1 + coef * w2; wh = wh - w_mag;
      double tauw[2];
      wallmodel_log(uh, wh, h, visc, s->eps, tauw);
      bcw[(long)i + (long)(n1 + 2) * k] = sgn * visci * tauw[1];
    }
}

/* set_bc('N', centred) on a y face with a plane of boundary values (0:n1+1, 0:n3+1) (bound.f90:320-353) */
static void set_bc_y_plane(const cpu_t *s, int ibound, const double *bc, double dr, double *p) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int i = 0; i <= n1 + 1; i++) {
      const double b = bc[(long)i + (long)(n1 + 2) * k];
      if (ibound == 0) p[IDX(s, i, 0, k)] = -dr * b + p[IDX(s, i, 1, k)];
      else p[IDX(s, i, n2 + 1, k)] = dr * b + p[IDX(s, i, n2, k)];
    }
}

/* bounduvw (bound.f90:18-154) for the channel on one rank.  is_updt_wm = 1: the fields are s->u, v, w and the wall-model
 * planes bcu%z, bcv%z are recomputed first (125-128); is_updt_wm = 0 (the filtered velocities of the dynamic model,
 * sgs.f90:256-257, with bcuf = bcvf = the initial planes = 0): the Neumann planes are zero. */
static void bounduvw_ch(cpu_t *s, double *u, double *v, double *w, int is_correc, int is_updt_wm) {
  const int n3 = s->n3;
  const long pl = (long)(s->n1 + 2) * (s->n2 + 2);
  bound_xy_periodic(s, u); bound_xy_periodic(s, v); bound_xy_periodic(s, w);
  const int impose_norm_bc = !is_correc;                                         /* cbc(:,3,3) = 'D','D': never 'PP' */
  if (impose_norm_bc) set_bc_z(s, 'D', 0, NULL, NULL, s->dzf[0], s->dzf[n3], w, 1, 1);
  for (int ib = 0; ib < 2; ib++)
    if (s->lwm[ib] == 0) {
      set_bc_z(s, 'D', 1, NULL, NULL, s->dzc[0], s->dzc[n3], u, ib == 0, ib == 1);
      set_bc_z(s, 'D', 1, NULL, NULL, s->dzc[0], s->dzc[n3], v, ib == 0, ib == 1);
    }
  if (is_updt_wm)
    for (int ib = 0; ib < 2; ib++) if (s->lwm[ib] != 0) cmpt_wallmodelbc_z(s, ib); /* updt_wallmodelbc, 125-128 */
  for (int ib = 0; ib < 2; ib++)
    if (s->lwm[ib] != 0) {                                                       /* cbcvel(:,3,1:2) = 'N' (initbc, bound.f90:746-758) */
      const double *b0u = is_updt_wm ? s->bcu_z : NULL, *b0v = is_updt_wm ? s->bcv_z : NULL;
      set_bc_z(s, 'N', 1, b0u, b0u ? b0u + pl : NULL, s->dzc[0], s->dzc[n3], u, ib == 0, ib == 1);
      set_bc_z(s, 'N', 1, b0v, b0v ? b0v + pl : NULL, s->dzc[0], s->dzc[n3], v, ib == 0, ib == 1);
    }
}
static void bounduvw_channel(cpu_t *s, int is_correc) { bounduvw_ch(s, s->u, s->v, s->w, is_correc, 1); }

/* boundp (bound.f90:156-200) for the channel: ctype 'N' (pressure) or 'D' (eddy viscosity), boundary value 0 */
static void boundp_channel(const cpu_t *s, char ctype, double *p) {
  bound_xy_periodic(s, p);
  set_bc_z(s, ctype, 1, NULL, NULL, s->dzc[0], s->dzc[s->n3], p, 1, 1);
}

static void fill_uvw(cpu_t *s, int is_correc) {
  if (s->gen) bounduvw_gen(s, s->u, s->v, s->w, is_correc);
  else if (s->zwall) bounduvw_channel(s, is_correc);
  else { bound_periodic(s, s->u); bound_periodic(s, s->v); bound_periodic(s, s->w); }
}
static void fill_p(cpu_t *s, double *p, char ctype) {
  if (s->gen) boundp_gen(s, ctype == 'N' ? s->cbcpre : s->cbcsgs, p);
  else if (s->zwall) boundp_channel(s, ctype, p); else bound_periodic(s, p);
}

/* ------------------------------------------------------------------------------------------------ SGS */
/* sgs.f90:1019-1110 (s0 only) fused with the 'smag' model of sgs.f90:69-152; no walls -> fd = 1. */
static void cmpt_sgs_smag(cpu_t *s) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
  const double dxi = s->dli[0], dyi = s->dli[1];
  const double *u = s->u, *v = s->v, *w = s->w;
  const double *uo = s->u, *vo = s->v;                                          /* the un-extrapolated fields: van Driest wall shear */
  if (!s->gen && s->zwall && (s->lwm[0] || s->lwm[1])) {
    /* sgs.f90:84-90: copies, then extrapolate(iface = 1, 2, lwm) (682-767): wall-parallel ghosts on the wall-model
     * faces by linear extrapolation; w (iface = 3) is not extrapolated on the z walls */
    memcpy(s->wku, s->u, sizeof(double) * (size_t)s->ntot); memcpy(s->wkv, s->v, sizeof(double) * (size_t)s->ntot);
    const double factor0 = s->dzc[0] * s->dzci[1], factor1 = s->dzc[n3] * s->dzci[n3 - 1];
    double *wk[2] = {s->wku, s->wkv};
    for (int m = 0; m < 2; m++) {
      double *p = wk[m];
#pragma omp parallel for schedule(static)
      for (int j = 0; j <= n2 + 1; j++)
        for (int i = 0; i <= n1 + 1; i++) {
          if (s->lwm[0]) p[IDX(s, i, j, 0)] = (1. + factor0) * p[IDX(s, i, j, 1)] - factor0 * p[IDX(s, i, j, 2)];
          if (s->lwm[1]) p[IDX(s, i, j, n3 + 1)] = (1. + factor1) * p[IDX(s, i, j, n3)] - factor1 * p[IDX(s, i, j, n3 - 1)];
        }
    }
    u = s->wku; v = s->wkv;
  }
  if (s->gen && (s->lwm_y[0] || s->lwm_y[1] || s->lwm[0] || s->lwm[1])) {
    /* the same in general (sgs.f90:84-90, extrapolate 682-767 with lwm): per array the y faces first, then the z faces;
     * u (iface 1): y and z, v (iface 2): z only, w (iface 3): y only */
    const size_t nb = sizeof(double) * (size_t)s->ntot;
    memcpy(s->wku, s->u, nb); memcpy(s->wkv, s->v, nb); memcpy(s->wkw, s->w, nb);
    const double factor0 = s->dzc[0] * s->dzci[1], factor1 = s->dzc[n3] * s->dzci[n3 - 1];
    double *wk3[3] = {s->wku, s->wkv, s->wkw};
    for (int m = 0; m < 3; m++) {
      double *p = wk3[m];
      if (m != 1)
        for (int k = 0; k <= n3 + 1; k++)
          for (int i = 0; i <= n1 + 1; i++) {
            if (s->lwm_y[0]) p[IDX(s, i, 0, k)] = 2. * p[IDX(s, i, 1, k)] - p[IDX(s, i, 2, k)];
            if (s->lwm_y[1]) p[IDX(s, i, n2 + 1, k)] = 2. * p[IDX(s, i, n2, k)] - p[IDX(s, i, n2 - 1, k)];
          }
      if (m != 2)
        for (int j = 0; j <= n2 + 1; j++)
          for (int i = 0; i <= n1 + 1; i++) {
            if (s->lwm[0]) p[IDX(s, i, j, 0)] = (1. + factor0) * p[IDX(s, i, j, 1)] - factor0 * p[IDX(s, i, j, 2)];
            if (s->lwm[1]) p[IDX(s, i, j, n3 + 1)] = (1. + factor1) * p[IDX(s, i, j, n3)] - factor1 * p[IDX(s, i, j, n3 - 1)];
          }
    }
    u = s->wku; v = s->wkv; w = s->wkw;
  }
  const double visc = s->visc, visci = 1. / visc;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double dzci_k = s->dzci[k], dzci_km = s->dzci[k - 1], dzfi_k = s->dzfi[k];
      const double dele = pow(s->dl[0] * s->dl[1] * s->dzf[k], 1. / 3.);
      for (int i = 1; i <= n1; i++) {
        double fd = 1.0;
        if (s->gen) {                                                            /* van Driest, sgs.f90:106-147, any set of walls */
          const double big = 1.7976931348623157e308;                             /* huge(1._rp), param.f90:25 */
          double dw[6] = {s->dl[0] * (i - 0.5), s->dl[0] * (n1 - i + 0.5), s->dl[1] * (j - 0.5), s->dl[1] * (n2 - j + 0.5), s->zc[k], s->l[2] - s->zc[k]};
          int loc = 0;
          for (int q = 0; q < 6; q++) {
            const double isw = (s->cbcvel[q & 1][q >> 1][q >> 1] == 'D' && ((q >> 1) == 0 || s->cbcpre[0][q >> 1] != 'P')) ? 1. : 0.;
            dw[q] = dw[q] * isw + big * (1. - isw);
          }
          for (int q = 1; q < 6; q++) if (dw[q] < dw[loc]) loc = q;              /* minloc: the first minimum */
          const double dw_min = dw[loc];
          double t1, t2, tauw_s;
          const double *ww = s->w;
#define F(f, a, b, c_) f[IDX(s, a, b, c_)]
          if (loc == 0) { t1 = F(vo, 1, j, k) - F(vo, 0, j, k) + F(vo, 1, j - 1, k) - F(vo, 0, j - 1, k); t2 = F(ww, 1, j, k) - F(ww, 0, j, k) + F(ww, 1, j, k - 1) - F(ww, 0, j, k - 1); tauw_s = sqrt(t1 * t1 + t2 * t2) * dxi; }
          else if (loc == 1) { t1 = F(vo, n1, j, k) - F(vo, n1 + 1, j, k) + F(vo, n1, j - 1, k) - F(vo, n1 + 1, j - 1, k); t2 = F(ww, n1, j, k) - F(ww, n1 + 1, j, k) + F(ww, n1, j, k - 1) - F(ww, n1 + 1, j, k - 1); tauw_s = sqrt(t1 * t1 + t2 * t2) * dxi; }
          else if (loc == 2) { t1 = F(uo, i, 1, k) - F(uo, i, 0, k) + F(uo, i - 1, 1, k) - F(uo, i - 1, 0, k); t2 = F(ww, i, 1, k) - F(ww, i, 0, k) + F(ww, i, 1, k - 1) - F(ww, i, 0, k - 1); tauw_s = sqrt(t1 * t1 + t2 * t2) * dyi; }
          else if (loc == 3) { t1 = F(uo, i, n2, k) - F(uo, i, n2 + 1, k) + F(uo, i - 1, n2, k) - F(uo, i - 1, n2 + 1, k); t2 = F(ww, i, n2, k) - F(ww, i, n2 + 1, k) + F(ww, i, n2, k - 1) - F(ww, i, n2 + 1, k - 1); tauw_s = sqrt(t1 * t1 + t2 * t2) * dyi; }
          else if (loc == 4) { t1 = F(uo, i, j, 1) - F(uo, i, j, 0) + F(uo, i - 1, j, 1) - F(uo, i - 1, j, 0); t2 = F(vo, i, j, 1) - F(vo, i, j, 0) + F(vo, i, j - 1, 1) - F(vo, i, j - 1, 0); tauw_s = sqrt(t1 * t1 + t2 * t2) * s->dzci[0]; }
          else { t1 = F(uo, i, j, n3) - F(uo, i, j, n3 + 1) + F(uo, i - 1, j, n3) - F(uo, i - 1, j, n3 + 1); t2 = F(vo, i, j, n3) - F(vo, i, j, n3 + 1) + F(vo, i, j - 1, n3) - F(vo, i, j - 1, n3 + 1); tauw_s = sqrt(t1 * t1 + t2 * t2) * s->dzci[n3]; }
#undef F
          tauw_s = 0.5 * visc * tauw_s;
          const double dw_plus = dw_min * sqrt(tauw_s) * visci;
          fd = 1. - exp(-dw_plus / 25.);
        } else if (s->zwall) {                                                   /* the same for the channel: walls 5 and 6 only */
          const double dw5 = s->zc[k], dw6 = s->l[2] - s->zc[k];
          double dw_min, tauw_s;
          if (dw5 <= dw6) {                                                      /* minloc: the first minimum */
            const long c1 = IDX(s, i, j, 1), c0 = IDX(s, i, j, 0);
            const double t1 = uo[c1] - uo[c0] + uo[c1 - 1] - uo[c0 - 1], t2 = vo[c1] - vo[c0] + vo[c1 - sj] - vo[c0 - sj];
            dw_min = dw5; tauw_s = sqrt(t1 * t1 + t2 * t2) * s->dzci[0];
          } else {
            const long c1 = IDX(s, i, j, n3), c0 = IDX(s, i, j, n3 + 1);
            const double t1 = uo[c1] - uo[c0] + uo[c1 - 1] - uo[c0 - 1], t2 = vo[c1] - vo[c0] + vo[c1 - sj] - vo[c0 - sj];
            dw_min = dw6; tauw_s = sqrt(t1 * t1 + t2 * t2) * s->dzci[n3];
          }
          tauw_s = 0.5 * visc * tauw_s;
          const double dw_plus = dw_min * sqrt(tauw_s) * visci;
          fd = 1. - exp(-dw_plus / 25.);
        }
        const double cs = C_SMAG * dele * fd;
        const long c = IDX(s, i, j, k);
        double s11 = (u[c] - u[c - 1]) * dxi;
        double s22 = (v[c] - v[c - sj]) * dyi;
        double s33 = (w[c] - w[c - sk]) * dzfi_k;
        double s12 = .125 * ((u[c + sj] - u[c]) * dyi + (v[c + 1] - v[c]) * dxi +
                             (u[c] - u[c - sj]) * dyi + (v[c + 1 - sj] - v[c - sj]) * dxi +
                             (u[c - 1 + sj] - u[c - 1]) * dyi + (v[c] - v[c - 1]) * dxi +
                             (u[c - 1] - u[c - 1 - sj]) * dyi + (v[c - sj] - v[c - 1 - sj]) * dxi);
        double s13 = .125 * ((u[c + sk] - u[c]) * dzci_k + (w[c + 1] - w[c]) * dxi +
                             (u[c] - u[c - sk]) * dzci_km + (w[c + 1 - sk] - w[c - sk]) * dxi +
                             (u[c - 1 + sk] - u[c - 1]) * dzci_k + (w[c] - w[c - 1]) * dxi +
                             (u[c - 1] - u[c - 1 - sk]) * dzci_km + (w[c - sk] - w[c - 1 - sk]) * dxi);
        double s23 = .125 * ((v[c + sk] - v[c]) * dzci_k + (w[c + sj] - w[c]) * dyi +
                             (v[c] - v[c - sk]) * dzci_km + (w[c + sj - sk] - w[c - sk]) * dyi +
                             (v[c - sj + sk] - v[c - sj]) * dzci_k + (w[c] - w[c - sj]) * dyi +
                             (v[c - sj] - v[c - sj - sk]) * dzci_km + (w[c - sk] - w[c - sj - sk]) * dyi);
        double s0 = sqrt(2. * (s11 * s11 + s22 * s22 + s33 * s33 + 2. * (s12 * s12 + s13 * s13 + s23 * s23)));
        s->s0[c] = s0;
        s->visct[c] = (cs * cs) * s0;
      }
    }
}

/* ---------------------------------------------------------------------------------------- dynamic Smagorinsky */
/* strain_rate with sij (sgs.f90:1019-1110): interior of s0 and of sij(1:6) = s11, s22, s33, s12, s13, s23 */
static void strain_rate_sij(const cpu_t *s, const double *u, const double *v, const double *w, double *s0a, double *const sij[6]) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
  const double dxi = s->dli[0], dyi = s->dli[1];
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double dzci_k = s->dzci[k], dzci_km = s->dzci[k - 1], dzfi_k = s->dzfi[k];
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
        double s11 = (u[c] - u[c - 1]) * dxi;
        double s22 = (v[c] - v[c - sj]) * dyi;
        double s33 = (w[c] - w[c - sk]) * dzfi_k;
        double s12 = .125 * ((u[c + sj] - u[c]) * dyi + (v[c + 1] - v[c]) * dxi +
                             (u[c] - u[c - sj]) * dyi + (v[c + 1 - sj] - v[c - sj]) * dxi +
                             (u[c - 1 + sj] - u[c - 1]) * dyi + (v[c] - v[c - 1]) * dxi +
                             (u[c - 1] - u[c - 1 - sj]) * dyi + (v[c - sj] - v[c - 1 - sj]) * dxi);
        double s13 = .125 * ((u[c + sk] - u[c]) * dzci_k + (w[c + 1] - w[c]) * dxi +
                             (u[c] - u[c - sk]) * dzci_km + (w[c + 1 - sk] - w[c - sk]) * dxi +
                             (u[c - 1 + sk] - u[c - 1]) * dzci_k + (w[c] - w[c - 1]) * dxi +
                             (u[c - 1] - u[c - 1 - sk]) * dzci_km + (w[c - sk] - w[c - 1 - sk]) * dxi);
        double s23 = .125 * ((v[c + sk] - v[c]) * dzci_k + (w[c + sj] - w[c]) * dyi +
                             (v[c] - v[c - sk]) * dzci_km + (w[c + sj - sk] - w[c - sk]) * dyi +
                             (v[c - sj + sk] - v[c - sj]) * dzci_k + (w[c] - w[c - sj]) * dyi +
                             (v[c - sj] - v[c - sj - sk]) * dzci_km + (w[c - sk] - w[c - sj - sk]) * dyi);
        s0a[c] = sqrt(2. * (s11 * s11 + s22 * s22 + s33 * s33 + 2. * (s12 * s12 + s13 * s13 + s23 * s23)));
        sij[0][c] = s11; sij[1][c] = s22; sij[2][c] = s33; sij[3][c] = s12; sij[4][c] = s13; sij[5][c] = s23;
      }
    }
}

/* filter3d, sgs.f90:616-680: 27-point top hat, trapezoidal weights 8/4/2/1 over 64, the reference's summation order */
static void filter3d(const cpu_t *s, const double *p, double *pf) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
#define P(di, dj, dk) p[c + (di) + (dj) * sj + (dk) * sk]
        pf[c] = (8. * (P(0, 0, 0)) +
                 4. * (P(-1, 0, 0) + P(0, -1, 0) + P(0, 0, -1) + P(1, 0, 0) + P(0, 1, 0) + P(0, 0, 1)) +
                 2. * (P(0, -1, -1) + P(-1, 0, -1) + P(-1, -1, 0) + P(0, 1, -1) + P(1, 0, -1) + P(1, -1, 0) +
                       P(0, -1, 1) + P(-1, 0, 1) + P(-1, 1, 0) + P(0, 1, 1) + P(1, 0, 1) + P(1, 1, 0)) +
                 1. * (P(-1, -1, -1) + P(1, -1, -1) + P(-1, 1, -1) + P(1, 1, -1) + P(-1, -1, 1) + P(1, -1, 1) + P(-1, 1, 1) + P(1, 1, 1))) / 64.;
#undef P
      }
}

/* extrapolate (sgs.f90:682-767) on the z faces, whole planes: with `cbc` (factor 1, faces with cbcvel(:,3,3) = 'D': both walls
 * of the channel) or with `lwm` (grid-ratio factors, wall-model faces only) */
static void extrapolate_z(const cpu_t *s, double *p, int use_lwm) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  if (!s->zwall) return;
  const double f0 = use_lwm ? s->dzc[0] * s->dzci[1] : 1., f1 = use_lwm ? s->dzc[n3] * s->dzci[n3 - 1] : 1.;
  const int do0 = use_lwm ? s->lwm[0] != 0 : 1, do1 = use_lwm ? s->lwm[1] != 0 : 1;
#pragma omp parallel for schedule(static)
  for (int j = 0; j <= n2 + 1; j++)
    for (int i = 0; i <= n1 + 1; i++) {
      if (do0) p[IDX(s, i, j, 0)] = (1. + f0) * p[IDX(s, i, j, 1)] - f0 * p[IDX(s, i, j, 2)];
      if (do1) p[IDX(s, i, j, n3 + 1)] = (1. + f1) * p[IDX(s, i, j, n3)] - f1 * p[IDX(s, i, j, n3 - 1)];
    }
}

/* ave1d_channel(idir = 3), sgs.f90:455-482: per-level sequential sum (i fastest), times the area ratio, written back to the
 * whole plane */
static void ave1d_channel_z(const cpu_t *s, double *p) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const double grid_area_ratio = s->dl[0] * s->dl[1] / (s->l[0] * s->l[1]);
#pragma omp parallel for schedule(static)
  for (int k = 1; k <= n3; k++) {
    double p1d_s = 0.;
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) p1d_s = p1d_s + p[IDX(s, i, j, k)];
    const double p1d = p1d_s * grid_area_ratio;
    for (int j = 0; j <= n2 + 1; j++)
      for (int i = 0; i <= n1 + 1; i++) p[IDX(s, i, j, k)] = p1d;
  }
}

/* cmpt_sgs('dsmag'), sgs.f90:153-380, 3-D test filter, plane averaging of the hard-wired _CHANNEL build (sgs.f90:8) */
static void cmpt_sgs_dsmag(cpu_t *s) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const size_t nb = sizeof(double) * (size_t)s->ntot;
  if (!s->dyn[0]) for (int m = 0; m < 24; m++) s->dyn[m] = (double *)calloc((size_t)s->ntot, 8);   /* 154-166 */
  double *uc = s->dyn[0], *vc = s->dyn[1], *wc = s->dyn[2], *uf = s->dyn[3], *vf = s->dyn[4], *wf = s->dyn[5];
  double *wk[6], *sij[6], *mij[6];
  for (int m = 0; m < 6; m++) { wk[m] = s->dyn[6 + m]; sij[m] = s->dyn[12 + m]; mij[m] = s->dyn[18 + m]; }
  double *s0 = s->s0;
  const char sg = 'D';                                                           /* cbcsgs(:,3) of the channel; periodic decks ignore it */
  /* 173-185 */
  memcpy(wk[0], s->u, nb); memcpy(wk[1], s->v, nb); memcpy(wk[2], s->w, nb);
  extrapolate_z(s, wk[0], 1); extrapolate_z(s, wk[1], 1);                        /* iface = 3 (w): not on the z faces */
  strain_rate_sij(s, wk[0], wk[1], wk[2], s0, sij);
  memcpy(s->visct, s0, nb);
  /* Mij, 191-223 */
  fill_p(s, s0, sg);
  for (int m = 0; m < 6; m++) fill_p(s, sij[m], sg);
  for (int m = 0; m < 6; m++) {
    double *a = wk[m]; const double *b = sij[m];
#pragma omp parallel for schedule(static)
    for (long q = 0; q < s->ntot; q++) a[q] = s0[q] * b[q];
  }
  for (int m = 0; m < 6; m++) extrapolate_z(s, wk[m], 0);
  for (int m = 0; m < 6; m++) filter3d(s, wk[m], mij[m]);
  /* 225-235 */
  memcpy(wk[0], s->u, nb); memcpy(wk[1], s->v, nb); memcpy(wk[2], s->w, nb);
  extrapolate_z(s, wk[0], 0); extrapolate_z(s, wk[1], 0);                        /* iface = 3: no */
  filter3d(s, wk[0], uf); filter3d(s, wk[1], vf); filter3d(s, wk[2], wf);
  /* 256-272 */
  if (s->zwall) bounduvw_ch(s, uf, vf, wf, 0, 0);
  else { bound_periodic(s, uf); bound_periodic(s, vf); bound_periodic(s, wf); }
  extrapolate_z(s, uf, 1); extrapolate_z(s, vf, 1);
  strain_rate_sij(s, uf, vf, wf, s0, sij);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double alph2 = (s->zwall && (k == 1 || k == n3)) ? 2.52 : 4.00;      /* cmpt_alph2, 769-822 */
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
        for (int m = 0; m < 6; m++) mij[m][c] = 2. * (mij[m][c] - alph2 * s0[c] * sij[m][c]);
      }
    }
  /* Lij (stored in sij), 277-315 */
  double **lij = sij;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) {                                            /* interpolate, 860-869 */
        const long c = IDX(s, i, j, k);
        uc[c] = 0.5 * (s->u[c] + s->u[c - 1]);
        vc[c] = 0.5 * (s->v[c] + s->v[c - s->sj]);
        wc[c] = 0.5 * (s->w[c] + s->w[c - s->sk]);
      }
  fill_p(s, uc, sg); fill_p(s, vc, sg); fill_p(s, wc, sg);
#pragma omp parallel for schedule(static)
  for (long q = 0; q < s->ntot; q++) {
    wk[0][q] = uc[q] * uc[q]; wk[1][q] = vc[q] * vc[q]; wk[2][q] = wc[q] * wc[q];
    wk[3][q] = uc[q] * vc[q]; wk[4][q] = uc[q] * wc[q]; wk[5][q] = vc[q] * wc[q];
  }
  for (int m = 0; m < 6; m++) extrapolate_z(s, wk[m], 0);
  for (int m = 0; m < 6; m++) filter3d(s, wk[m], lij[m]);
  extrapolate_z(s, uc, 0); extrapolate_z(s, vc, 0); extrapolate_z(s, wc, 0);
  filter3d(s, uc, uf); filter3d(s, vc, vf); filter3d(s, wc, wf);
  /* 328-358 */
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
        double mij_s[6], lij_s[6];
        for (int m = 0; m < 6; m++) { mij_s[m] = mij[m][c]; lij_s[m] = lij[m][c]; }
        lij_s[0] = lij_s[0] - uf[c] * uf[c];
        lij_s[1] = lij_s[1] - vf[c] * vf[c];
        lij_s[2] = lij_s[2] - wf[c] * wf[c];
        lij_s[3] = lij_s[3] - uf[c] * vf[c];
        lij_s[4] = lij_s[4] - uf[c] * wf[c];
        lij_s[5] = lij_s[5] - vf[c] * wf[c];
        wk[0][c] = mij_s[0] * lij_s[0] + mij_s[1] * lij_s[1] + mij_s[2] * lij_s[2] +
                   (mij_s[3] * lij_s[3] + mij_s[4] * lij_s[4] + mij_s[5] * lij_s[5]) * 2.;
        wk[1][c] = mij_s[0] * mij_s[0] + mij_s[1] * mij_s[1] + mij_s[2] * mij_s[2] +
                   (mij_s[3] * mij_s[3] + mij_s[4] * mij_s[4] + mij_s[5] * mij_s[5]) * 2.;
      }
  ave1d_channel_z(s, wk[0]); ave1d_channel_z(s, wk[1]);                          /* 363-364 */
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)                                                  /* 372-380 */
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
        double t = s->visct[c] * wk[0][c] / wk[1][c];
        s->visct[c] = fmax(t, 0.);
      }
}

/* cmpt_sgs (sgs.f90:21-386): 'none' (58-68: the eddy viscosity is zeroed on the first call and never touched again), 'smag', 'dsmag' */
static void cmpt_sgs(cpu_t *s) {
  if (s->dsmag == 2) { if (!s->sgs_done) { s->sgs_done = 1; memset(s->visct, 0, sizeof(double) * (size_t)s->ntot); } return; }
  if (s->dsmag) cmpt_sgs_dsmag(s); else cmpt_sgs_smag(s);
}

/* ---------------------------------------------------------------------------------------------- mom + rk */
/* mom.f90:142-302 (explicit branch) and rk.f90:45-100, then cmpt_bulk_forcing (rk.f90:197-222). */
static void rk_substep(cpu_t *s, const double rkpar[2], double dt) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
  const double dxi = s->dli[0], dyi = s->dli[1], visc = s->visc;
  const double *u = s->u, *v = s->v, *w = s->w, *t = s->visct;
  double *du = s->rhs[0], *dv = s->rhs[1], *dw = s->rhs[2];
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double dzci_k = s->dzci[k], dzci_km = s->dzci[k - 1], dzfi_k = s->dzfi[k], dzfi_kp = s->dzfi[k + 1];
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k), o = (long)(i - 1) + (long)n1 * ((long)(j - 1) + (long)n2 * (long)(k - 1));
#define A(f, di, dj, dk) f[c + (di) + (dj) * sj + (dk) * sk]
        const double u_ccm = A(u, 0, 0, -1), u_pcm = A(u, 1, 0, -1), u_cpm = A(u, 0, 1, -1), u_cmc = A(u, 0, -1, 0), u_pmc = A(u, 1, -1, 0),
                     u_mcc = A(u, -1, 0, 0), u_ccc = A(u, 0, 0, 0), u_pcc = A(u, 1, 0, 0), u_mpc = A(u, -1, 1, 0), u_cpc = A(u, 0, 1, 0),
                     u_cmp = A(u, 0, -1, 1), u_mcp = A(u, -1, 0, 1), u_ccp = A(u, 0, 0, 1);
        const double v_ccm = A(v, 0, 0, -1), v_pcm = A(v, 1, 0, -1), v_cpm = A(v, 0, 1, -1), v_cmc = A(v, 0, -1, 0), v_pmc = A(v, 1, -1, 0),
                     v_mcc = A(v, -1, 0, 0), v_ccc = A(v, 0, 0, 0), v_pcc = A(v, 1, 0, 0), v_mpc = A(v, -1, 1, 0), v_cpc = A(v, 0, 1, 0),
                     v_cmp = A(v, 0, -1, 1), v_mcp = A(v, -1, 0, 1), v_ccp = A(v, 0, 0, 1);
        const double w_ccm = A(w, 0, 0, -1), w_pcm = A(w, 1, 0, -1), w_cpm = A(w, 0, 1, -1), w_cmc = A(w, 0, -1, 0), w_pmc = A(w, 1, -1, 0),
                     w_mcc = A(w, -1, 0, 0), w_ccc = A(w, 0, 0, 0), w_pcc = A(w, 1, 0, 0), w_mpc = A(w, -1, 1, 0), w_cpc = A(w, 0, 1, 0),
                     w_cmp = A(w, 0, -1, 1), w_mcp = A(w, -1, 0, 1), w_ccp = A(w, 0, 0, 1);
        const double s_ccm = A(t, 0, 0, -1), s_pcm = A(t, 1, 0, -1), s_cpm = A(t, 0, 1, -1), s_cmc = A(t, 0, -1, 0), s_pmc = A(t, 1, -1, 0),
                     s_mcc = A(t, -1, 0, 0), s_ccc = A(t, 0, 0, 0), s_pcc = A(t, 1, 0, 0), s_mpc = A(t, -1, 1, 0), s_cpc = A(t, 0, 1, 0),
                     s_cmp = A(t, 0, -1, 1), s_mcp = A(t, -1, 0, 1), s_ccp = A(t, 0, 0, 1), s_ppc = A(t, 1, 1, 0), s_pcp = A(t, 1, 0, 1),
                     s_cpp = A(t, 0, 1, 1);
#undef A
        (void)u_pcm; (void)u_cpm; (void)u_pmc; (void)u_cmp; (void)v_pcm; (void)v_cpm; (void)v_mpc; (void)v_mcp; (void)w_pmc; (void)w_mpc;
        (void)w_cmp; (void)w_mcp;
        double visc_ip, visc_im, visc_jp, visc_jm, visc_kp, visc_km;
        /* x momentum, mom.f90:142-186 */
        visc_ip = s_pcc; visc_im = s_ccc;
        visc_jp = 0.25 * (s_ccc + s_pcc + s_cpc + s_ppc);
        visc_jm = 0.25 * (s_ccc + s_pcc + s_cmc + s_pmc);
        visc_kp = 0.25 * (s_ccc + s_pcc + s_ccp + s_pcp);
        visc_km = 0.25 * (s_ccc + s_pcc + s_ccm + s_pcm);
        {
          double dudx_ip = (u_pcc - u_ccc) * dxi, dudx_im = (u_ccc - u_mcc) * dxi;
          double dudy_jp = (u_cpc - u_ccc) * dyi, dudy_jm = (u_ccc - u_cmc) * dyi;
          double dudz_kp = (u_ccp - u_ccc) * dzci_k, dudz_km = (u_ccc - u_ccm) * dzci_km;
          double dvdx_jp = (v_pcc - v_ccc) * dxi, dvdx_jm = (v_pmc - v_cmc) * dxi;
          double dwdx_kp = (w_pcc - w_ccc) * dxi, dwdx_km = (w_pcm - w_ccm) * dxi;
          double uu_ip = 0.25 * (u_pcc + u_ccc) * (u_ccc + u_pcc), uu_im = 0.25 * (u_mcc + u_ccc) * (u_ccc + u_mcc);
          double vu_jp = 0.25 * (v_pcc + v_ccc) * (u_ccc + u_cpc), vu_jm = 0.25 * (v_pmc + v_cmc) * (u_ccc + u_cmc);
          double wu_kp = 0.25 * (w_pcc + w_ccc) * (u_ccc + u_ccp), wu_km = 0.25 * (w_pcm + w_ccm) * (u_ccc + u_ccm);
          double d_xy = visc * (dudx_ip - dudx_im) * dxi + visc * (dudy_jp - dudy_jm) * dyi;
          double d_z = visc * (dudz_kp - dudz_km) * dzfi_k;
          double r = -(uu_ip - uu_im) * dxi - (vu_jp - vu_jm) * dyi - (wu_kp - wu_km) * dzfi_k +
                     (visc_ip * (dudx_ip + dudx_ip) - visc_im * (dudx_im + dudx_im)) * dxi +
                     (visc_jp * (dudy_jp + dvdx_jp) - visc_jm * (dudy_jm + dvdx_jm)) * dyi +
                     (visc_kp * (dudz_kp + dwdx_kp) - visc_km * (dudz_km + dwdx_km)) * dzfi_k;
          du[o] = r + d_xy + d_z;
        }
        /* y momentum, mom.f90:187-231 */
        visc_ip = 0.25 * (s_ccc + s_cpc + s_pcc + s_ppc);
        visc_im = 0.25 * (s_ccc + s_cpc + s_mcc + s_mpc);
        visc_jp = s_cpc; visc_jm = s_ccc;
        visc_kp = 0.25 * (s_ccc + s_cpc + s_ccp + s_cpp);
        visc_km = 0.25 * (s_ccc + s_cpc + s_ccm + s_cpm);
        {
          double dvdx_ip = (v_pcc - v_ccc) * dxi, dvdx_im = (v_ccc - v_mcc) * dxi;
          double dvdy_jp = (v_cpc - v_ccc) * dyi, dvdy_jm = (v_ccc - v_cmc) * dyi;
          double dvdz_kp = (v_ccp - v_ccc) * dzci_k, dvdz_km = (v_ccc - v_ccm) * dzci_km;
          double dudy_ip = (u_cpc - u_ccc) * dyi, dudy_im = (u_mpc - u_mcc) * dyi;
          double dwdy_kp = (w_cpc - w_ccc) * dyi, dwdy_km = (w_cpm - w_ccm) * dyi;
          double uv_ip = 0.25 * (u_ccc + u_cpc) * (v_ccc + v_pcc), uv_im = 0.25 * (u_mcc + u_mpc) * (v_ccc + v_mcc);
          double vv_jp = 0.25 * (v_ccc + v_cpc) * (v_ccc + v_cpc), vv_jm = 0.25 * (v_ccc + v_cmc) * (v_ccc + v_cmc);
          double wv_kp = 0.25 * (w_ccc + w_cpc) * (v_ccc + v_ccp), wv_km = 0.25 * (w_ccm + w_cpm) * (v_ccc + v_ccm);
          double d_xy = visc * (dvdx_ip - dvdx_im) * dxi + visc * (dvdy_jp - dvdy_jm) * dyi;
          double d_z = visc * (dvdz_kp - dvdz_km) * dzfi_k;
          double r = -(uv_ip - uv_im) * dxi - (vv_jp - vv_jm) * dyi - (wv_kp - wv_km) * dzfi_k +
                     (visc_ip * (dvdx_ip + dudy_ip) - visc_im * (dvdx_im + dudy_im)) * dxi +
                     (visc_jp * (dvdy_jp + dvdy_jp) - visc_jm * (dvdy_jm + dvdy_jm)) * dyi +
                     (visc_kp * (dvdz_kp + dwdy_kp) - visc_km * (dvdz_km + dwdy_km)) * dzfi_k;
          dv[o] = r + d_xy + d_z;
        }
        /* z momentum, mom.f90:232-276 */
        visc_ip = 0.25 * (s_ccc + s_ccp + s_pcc + s_pcp);
        visc_im = 0.25 * (s_ccc + s_ccp + s_mcc + s_mcp);
        visc_jp = 0.25 * (s_ccc + s_ccp + s_cpc + s_cpp);
        visc_jm = 0.25 * (s_ccc + s_ccp + s_cmc + s_cmp);
        visc_kp = s_ccp; visc_km = s_ccc;
        {
          double dwdx_ip = (w_pcc - w_ccc) * dxi, dwdx_im = (w_ccc - w_mcc) * dxi;
          double dwdy_jp = (w_cpc - w_ccc) * dyi, dwdy_jm = (w_ccc - w_cmc) * dyi;
          double dwdz_kp = (w_ccp - w_ccc) * dzfi_kp, dwdz_km = (w_ccc - w_ccm) * dzfi_k;
          double dudz_ip = (u_ccp - u_ccc) * dzci_k, dudz_im = (u_mcp - u_mcc) * dzci_k;
          double dvdz_jp = (v_ccp - v_ccc) * dzci_k, dvdz_jm = (v_cmp - v_cmc) * dzci_k;
          double uw_ip = 0.25 * (u_ccc + u_ccp) * (w_ccc + w_pcc), uw_im = 0.25 * (u_mcc + u_mcp) * (w_ccc + w_mcc);
          double vw_jp = 0.25 * (v_ccc + v_ccp) * (w_ccc + w_cpc), vw_jm = 0.25 * (v_cmc + v_cmp) * (w_ccc + w_cmc);
          double ww_kp = 0.25 * (w_ccc + w_ccp) * (w_ccc + w_ccp), ww_km = 0.25 * (w_ccc + w_ccm) * (w_ccc + w_ccm);
          double d_xy = visc * (dwdx_ip - dwdx_im) * dxi + visc * (dwdy_jp - dwdy_jm) * dyi;
          double d_z = visc * (dwdz_kp - dwdz_km) * dzci_k;
          double r = -(uw_ip - uw_im) * dxi - (vw_jp - vw_jm) * dyi - (ww_kp - ww_km) * dzci_k +
                     (visc_ip * (dwdx_ip + dudz_ip) - visc_im * (dwdx_im + dudz_im)) * dxi +
                     (visc_jp * (dwdy_jp + dvdz_jp) - visc_jm * (dwdy_jm + dvdz_jm)) * dyi +
                     (visc_kp * (dwdz_kp + dwdz_kp) - visc_km * (dwdz_km + dwdz_km)) * dzci_k;
          dw[o] = r + d_xy + d_z;
        }
      }
    }
  /* rk.f90:76-100: update (needs the old u,v,w everywhere above, hence a second sweep) */
  const double factor1 = rkpar[0] * dt, factor2 = rkpar[1] * dt, factor12 = factor1 + factor2;
  double *uu = s->u, *vv = s->v, *ww = s->w;
  const double *p = s->p;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double dzci_k = s->dzci[k];
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k), o = (long)(i - 1) + (long)n1 * ((long)(j - 1) + (long)n2 * (long)(k - 1));
        uu[c] = uu[c] + factor1 * du[o] + factor2 * s->rhso[0][o] + factor12 * (s->bforce[0] - dxi * (p[c + 1] - p[c]));
        vv[c] = vv[c] + factor1 * dv[o] + factor2 * s->rhso[1][o] + factor12 * (s->bforce[1] - dyi * (p[c + sj] - p[c]));
        ww[c] = ww[c] + factor1 * dw[o] + factor2 * s->rhso[2][o] + factor12 * (s->bforce[2] - dzci_k * (p[c + sk] - p[c]));
      }
    }
  for (int m = 0; m < 3; m++) { double *tmp = s->rhs[m]; s->rhs[m] = s->rhso[m]; s->rhso[m] = tmp; } /* swap, rk.f90:92-100 */
  /* cmpt_bulk_forcing (rk.f90:197-222) with bulk_mean (utils.f90:33-44): ONE sequential accumulation, i fastest -- the
   * summation order is part of the result (the reference compiles that routine with -O0 for this reason) */
  double *fld[3] = {s->u, s->v, s->w};
  const double *gvr[3] = {s->gvr_f, s->gvr_f, s->gvr_c};
  for (int m = 0; m < 3; m++) {
    s->f[m] = 0.;
    if (!s->is_forced[m]) continue;
    double mean = 0.;
    for (int k = 1; k <= n3; k++)
      for (int j = 1; j <= n2; j++)
        for (int i = 1; i <= n1; i++) mean = mean + fld[m][IDX(s, i, j, k)] * gvr[m][k];
    s->f[m] = s->velf[m] - mean;
  }
}

/* bulk_forcing, mom.f90:311-335 */
static void bulk_forcing(cpu_t *s) {
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  double *fld[3] = {s->u, s->v, s->w};
  for (int m = 0; m < 3; m++) {
    if (!s->is_forced[m]) continue;
    const double ff = s->f[m];
    double *a = fld[m];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= n3; k++)
      for (int j = 1; j <= n2; j++)
        for (int i = 1; i <= n1; i++) { const long c = IDX(s, i, j, k); a[c] = a[c] + ff; }
  }
}

/* -------------------------------------------------------------------------------- pressure correction */
static void fillps(cpu_t *s, double dti) { /* fillps.f90:33-47 */
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
  const double dtidxi = dti * s->dli[0], dtidyi = dti * s->dli[1];
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++) {
      const double dzfi_k = s->dzfi[k];
      for (int i = 1; i <= n1; i++) {
        const long c = IDX(s, i, j, k);
        s->pp[c] = (s->w[c] - s->w[c - sk]) * dti * dzfi_k + (s->v[c] - s->v[c - sj]) * dtidyi + (s->u[c] - s->u[c - 1]) * dtidxi;
      }
    }
}

/* solver.f90:153-179 for the columns of one j-row, vectorised over i (same operation order per column). */
static void dgtsv_row(int n1, int n, const double *a, const double *bb, const double *c, double *p, double *d, double eps) {
  /* bb, p, d: [l][i] with row length n1 */
  for (int i = 0; i < n1; i++) { double z = 1. / (bb[i] + eps); d[i] = c[0] * z; p[i] = p[i] * z; }
  for (int l = 1; l < n; l++)
    for (int i = 0; i < n1; i++) {
      double z = 1. / (bb[(long)l * n1 + i] - a[l] * d[(long)(l - 1) * n1 + i] + eps);
      d[(long)l * n1 + i] = c[l] * z;
      p[(long)l * n1 + i] = (p[(long)l * n1 + i] - a[l] * p[(long)(l - 1) * n1 + i]) * z;
    }
  for (int l = n - 2; l >= 0; l--)
    for (int i = 0; i < n1; i++) p[(long)l * n1 + i] = p[(long)l * n1 + i] - d[(long)l * n1 + i] * p[(long)(l + 1) * n1 + i];
}

static void solver(cpu_t *s) { /* solver.f90:20-80, one rank: no transposes */
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  double *wk = s->wk;
  const int nmax = n1 > n2 ? n1 : n2;
#pragma omp parallel
  {
    cpx *x = (cpx *)malloc(sizeof(cpx) * (size_t)nmax), *y = (cpx *)malloc(sizeof(cpx) * (size_t)nmax);
#pragma omp for schedule(static)
    for (int k = 0; k < n3; k++) {
      double *pl = wk + (long)n1 * n2 * k;
      for (int j = 0; j < n2; j++) memcpy(pl + (long)n1 * j, s->pp + IDX(s, 1, j + 1, k + 1), sizeof(double) * (size_t)n1);
      for (int j = 0; j < n2; j += 2) fwd_pair(s->kx, &s->px, pl + (long)n1 * j, j + 1 < n2 ? pl + (long)n1 * (j + 1) : NULL, 1, x, y);
      for (int i = 0; i < n1; i += 2) fwd_pair(s->ky, &s->py, pl + i, i + 1 < n1 ? pl + i + 1 : NULL, n1, x, y);
    }
    free(x); free(y);
  }
  if (s->zwall) {                                                                /* gaussel, solver.f90:82-107 */
#pragma omp parallel
    {
      const int n = n3;
      size_t rowsz = (size_t)n1 * (size_t)n;
      double *bb = (double *)malloc(sizeof(double) * rowsz), *p1 = (double *)malloc(sizeof(double) * rowsz), *d = (double *)malloc(sizeof(double) * rowsz);
#pragma omp for schedule(static)
      for (int j = 0; j < n2; j++) {
        for (int l = 0; l < n; l++)
          for (int i = 0; i < n1; i++) {
            bb[(long)l * n1 + i] = s->b[l] + s->lambdaxy[i + (long)n1 * j];
            p1[(long)l * n1 + i] = wk[i + (long)n1 * (j + (long)n2 * l)];
          }
        dgtsv_row(n1, n, s->a, bb, s->c, p1, d, s->eps);
        for (int l = 0; l < n; l++)
          for (int i = 0; i < n1; i++) wk[i + (long)n1 * (j + (long)n2 * l)] = p1[(long)l * n1 + i];
      }
      free(bb); free(p1); free(d);
    }
  } else
  /* gaussel_periodic, solver.f90:109-151 */
#pragma omp parallel
  {
    const int n = n3;
    size_t rowsz = (size_t)n1 * (size_t)n;
    double *bb = (double *)malloc(sizeof(double) * rowsz), *p1 = (double *)malloc(sizeof(double) * rowsz),
           *p2 = (double *)malloc(sizeof(double) * rowsz), *d = (double *)malloc(sizeof(double) * rowsz);
#pragma omp for schedule(static)
    for (int j = 0; j < n2; j++) {
      for (int l = 0; l < n; l++)
        for (int i = 0; i < n1; i++) {
          bb[(long)l * n1 + i] = s->b[l] + s->lambdaxy[i + (long)n1 * j];
          p1[(long)l * n1 + i] = wk[i + (long)n1 * (j + (long)n2 * l)];
          p2[(long)l * n1 + i] = 0.0;
        }
      for (int i = 0; i < n1; i++) { p2[i] = -s->a[0]; p2[(long)(n - 2) * n1 + i] = -s->c[n - 2]; }
      dgtsv_row(n1, n - 1, s->a, bb, s->c, p1, d, s->eps);
      dgtsv_row(n1, n - 1, s->a, bb, s->c, p2, d, s->eps);
      for (int i = 0; i < n1; i++) {
        double pn = (p1[(long)(n - 1) * n1 + i] - s->c[n - 1] * p1[i] - s->a[n - 1] * p1[(long)(n - 2) * n1 + i]) /
                    (bb[(long)(n - 1) * n1 + i] + s->c[n - 1] * p2[i] + s->a[n - 1] * p2[(long)(n - 2) * n1 + i] + s->eps);
        wk[i + (long)n1 * (j + (long)n2 * (n - 1))] = pn;
        for (int l = 0; l < n - 1; l++) wk[i + (long)n1 * (j + (long)n2 * l)] = p1[(long)l * n1 + i] + p2[(long)l * n1 + i] * pn;
      }
    }
    free(bb); free(p1); free(p2); free(d);
  }
#pragma omp parallel
  {
    cpx *x = (cpx *)malloc(sizeof(cpx) * (size_t)nmax), *y = (cpx *)malloc(sizeof(cpx) * (size_t)nmax);
#pragma omp for schedule(static)
    for (int k = 0; k < n3; k++) {
      double *pl = wk + (long)n1 * n2 * k;
      for (int i = 0; i < n1; i += 2) bwd_pair(s->ky, &s->py, pl + i, i + 1 < n1 ? pl + i + 1 : NULL, n1, x, y);
      for (int j = 0; j < n2; j += 2) bwd_pair(s->kx, &s->px, pl + (long)n1 * j, j + 1 < n2 ? pl + (long)n1 * (j + 1) : NULL, 1, x, y);
      for (int j = 0; j < n2; j++) {
        double *dst = s->pp + IDX(s, 1, j + 1, k + 1);
        for (int i = 0; i < n1; i++) dst[i] = pl[(long)n1 * j + i] * s->normfft;
      }
    }
    free(x); free(y);
  }
}

static void correc(cpu_t *s, double dt) { /* correc.f90:41-67 (ghost rows included) */
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  const long sj = s->sj, sk = s->sk;
  const double factori = dt * s->dli[0], factorj = dt * s->dli[1];
  const double *p = s->pp;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= n3 + 1; k++)
    for (int j = 0; j <= n2 + 1; j++) {
      for (int i = 0; i <= n1 + 1; i++) {
        const long c = IDX(s, i, j, k);
        if (i <= n1) s->u[c] = s->u[c] - factori * (p[c + 1] - p[c]);
        if (j <= n2) s->v[c] = s->v[c] - factorj * (p[c + sj] - p[c]);
        if (k <= n3) s->w[c] = s->w[c] - dt * s->dzci[k] * (p[c + sk] - p[c]);
      }
    }
}

static void updatep(cpu_t *s) { /* updatep.f90:44-48 (explicit diffusion) */
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; k++)
    for (int j = 1; j <= n2; j++)
      for (int i = 1; i <= n1; i++) { const long c = IDX(s, i, j, k); s->p[c] = s->p[c] + s->pp[c]; }
}

/* ------------------------------------------------------------------------------------------------ API */
void *cales_cpu_new(int n1, int n2, int n3, const double *l, double visc, const double *dzc, const double *dzf) {
  cpu_t *s = (cpu_t *)calloc(1, sizeof(cpu_t));
  s->n1 = n1; s->n2 = n2; s->n3 = n3; s->sj = n1 + 2; s->sk = (long)(n1 + 2) * (n2 + 2);
  s->ntot = s->sk * (n3 + 2); s->nint = (long)n1 * n2 * n3;
  int ng[3] = {n1, n2, n3};
  for (int d = 0; d < 3; d++) { s->l[d] = l[d]; s->dl[d] = l[d] / (1. * ng[d]); s->dli[d] = 1. / s->dl[d]; } /* param.f90:152-153 */
  s->visc = visc; s->eps = 2.220446049250313e-16; /* epsilon(1._rp), param.f90:20 */
  size_t nz = (size_t)(n3 + 2);
  s->dzc = (double *)malloc(8 * nz); s->dzf = (double *)malloc(8 * nz); s->dzci = (double *)malloc(8 * nz); s->dzfi = (double *)malloc(8 * nz);
  for (int k = 0; k < n3 + 2; k++) { s->dzc[k] = dzc[k]; s->dzf[k] = dzf[k]; s->dzci[k] = 1. / dzc[k]; s->dzfi[k] = 1. / dzf[k]; }
  double **f[] = {&s->u, &s->v, &s->w, &s->p, &s->pp, &s->visct, &s->s0};
  for (int m = 0; m < 7; m++) *f[m] = (double *)calloc((size_t)s->ntot, 8);
  for (int m = 0; m < 3; m++) { s->rhs[m] = (double *)calloc((size_t)s->nint, 8); s->rhso[m] = (double *)calloc((size_t)s->nint, 8); }
  s->wk = (double *)calloc((size_t)s->nint, 8);
  plan_init(&s->px, n1); plan_init(&s->py, n2);
  s->kx = 'P'; s->ky = 'P';
  /* initsolver.f90:17-64 with cbcpre = P/P/P, c_or_f = c,c,c: eigenvalues 66-78, tridmatrix 127-169, normfft fft.f90:99,136 */
  const double pi = acos(-1.0);
  double *lx = (double *)malloc(8 * (size_t)n1), *ly = (double *)malloc(8 * (size_t)n2);
  for (int i = 0; i < n1; i++) lx[i] = -2. * (1. - cos((2 * i) * pi / (1. * n1))) * (s->dli[0] * s->dli[0]);
  for (int j = 0; j < n2; j++) ly[j] = -2. * (1. - cos((2 * j) * pi / (1. * n2))) * (s->dli[1] * s->dli[1]);
  s->lambdaxy = (double *)malloc(8 * (size_t)n1 * n2);
  for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) s->lambdaxy[i + (long)n1 * j] = lx[i] + ly[j];
  free(lx); free(ly);
  s->a = (double *)malloc(8 * (size_t)n3); s->b = (double *)malloc(8 * (size_t)n3); s->c = (double *)malloc(8 * (size_t)n3);
  for (int k = 1; k <= n3; k++) { s->a[k - 1] = s->dzfi[k] * s->dzci[k - 1]; s->c[k - 1] = s->dzfi[k] * s->dzci[k]; s->b[k - 1] = -(s->a[k - 1] + s->c[k - 1]); }
  s->normfft = 1. / ((1. * (n1 + 0.)) * (1. * (n2 + 0.)));
  return s;
}

/* Channel family: walls at both z faces (cbcvel(:,3,:) = 'D', cbcpre(:,3) = 'N', cbcsgs(:,3) = 'D', boundary values 0),
 * lwm[2] = lwm(0:1,3), hwm, forcing.  Call once, right after cales_cpu_new and before the fields are filled.
 * initbc (bound.f90:846-863): index_wm; initsolver: tridmatrix fold-in b(1) += a(1), b(n) += c(n) for 'N'
 * (initsolver.f90:149-164); grid_vol_ratio (main.f90:270-283). */
int cales_cpu_set_channel(void *h, const double *zc, const double *zf, const int *lwm, double hwm, const int *is_forced,
                          const double *velf, const double *bforce) {
  cpu_t *s = (cpu_t *)h;
  const int n3 = s->n3;
  s->zwall = 1;
  size_t nz = (size_t)(n3 + 2);
  s->zc = (double *)malloc(8 * nz); s->zf = (double *)malloc(8 * nz); s->gvr_c = (double *)malloc(8 * nz); s->gvr_f = (double *)malloc(8 * nz);
  for (int k = 0; k < n3 + 2; k++) {
    s->zc[k] = zc[k]; s->zf[k] = zf[k];
    s->gvr_c[k] = s->dl[0] * s->dl[1] * s->dzc[k] / (s->l[0] * s->l[1] * s->l[2]);
    s->gvr_f[k] = s->dl[0] * s->dl[1] * s->dzf[k] / (s->l[0] * s->l[1] * s->l[2]);
  }
  for (int m = 0; m < 3; m++) { s->is_forced[m] = is_forced[m]; s->velf[m] = velf[m]; s->bforce[m] = bforce[m]; }
  s->lwm[0] = lwm[0]; s->lwm[1] = lwm[1]; s->hwm = hwm;
  if (lwm[0] || lwm[1]) {
    if (lwm[0] != 1 && lwm[0] != 0) return 1;                                    /* WM_LOG only */
    if (lwm[1] != 1 && lwm[1] != 0) return 1;
    const size_t pl = (size_t)(s->n1 + 2) * (size_t)(s->n2 + 2);
    s->bcu_z = (double *)calloc(2 * pl, 8); s->bcv_z = (double *)calloc(2 * pl, 8);
    s->wku = (double *)calloc((size_t)s->ntot, 8); s->wkv = (double *)calloc((size_t)s->ntot, 8);
    if (lwm[0]) { int k = 1; while (s->zc[k] < hwm) k = k + 1; s->index_wm[0] = k; }
    if (lwm[1]) { int k = n3; while (s->l[2] - s->zc[k] < hwm) k = k - 1; s->index_wm[1] = k; }
  }
  s->b[0] = s->b[0] + 1. * s->a[0];
  s->b[n3 - 1] = s->b[n3 - 1] + 1. * s->c[n3 - 1];
  return 0;
}

/* Walls in x and / or y as well (square duct, lid-driven cavity; no wall model, static Smagorinsky): cbcvel(0:1,3,3),
 * bcvel(0:1,3,3), cbcpre(0:1,3), cbcsgs(0:1,3) in Fortran order; pressure and eddy-viscosity boundary values are 0.  Call
 * right after cales_cpu_new.  Transform kinds by find_fft (fft.f90:192-245): P,P -> R2HC/HC2R, N,N cell-centred ->
 * REDFT10/REDFT01; eigenvalues initsolver.f90:66-103; normfft fft.f90:99,136,142; tridmatrix fold-in initsolver.f90:149-164. */
int cales_cpu_set_bc(void *h, const char *cbcvel, const double *bcvel, const char *cbcpre, const char *cbcsgs, const double *zc, const double *zf,
                     const int *is_forced, const double *velf, const double *bforce) {
  cpu_t *s = (cpu_t *)h;
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  for (int ivel = 0; ivel < 3; ivel++)
    for (int idir = 0; idir < 3; idir++)
      for (int ib = 0; ib < 2; ib++) {
        s->cbcvel[ib][idir][ivel] = cbcvel[ib + 2 * (idir + 3 * ivel)];
        s->bcvel[ib][idir][ivel] = bcvel[ib + 2 * (idir + 3 * ivel)];
      }
  for (int idir = 0; idir < 3; idir++)
    for (int ib = 0; ib < 2; ib++) { s->cbcpre[ib][idir] = cbcpre[ib + 2 * idir]; s->cbcsgs[ib][idir] = cbcsgs[ib + 2 * idir]; }
  for (int idir = 0; idir < 3; idir++) {
    const char a = s->cbcpre[0][idir], b = s->cbcpre[1][idir];
    if (!((a == 'P' && b == 'P') || (a == 'N' && b == 'N'))) return 1;            /* the kinds restated here */
  }
  s->gen = 1;
  s->zwall = s->cbcpre[0][2] != 'P';
  size_t nz = (size_t)(n3 + 2);
  s->zc = (double *)malloc(8 * nz); s->zf = (double *)malloc(8 * nz); s->gvr_c = (double *)malloc(8 * nz); s->gvr_f = (double *)malloc(8 * nz);
  for (int k = 0; k < n3 + 2; k++) {
    s->zc[k] = zc[k]; s->zf[k] = zf[k];
    s->gvr_c[k] = s->dl[0] * s->dl[1] * s->dzc[k] / (s->l[0] * s->l[1] * s->l[2]);
    s->gvr_f[k] = s->dl[0] * s->dl[1] * s->dzf[k] / (s->l[0] * s->l[1] * s->l[2]);
  }
  for (int m = 0; m < 3; m++) { s->is_forced[m] = is_forced[m]; s->velf[m] = velf[m]; s->bforce[m] = bforce[m]; }
  s->kx = s->cbcpre[0][0]; s->ky = s->cbcpre[0][1];
  const double pi = acos(-1.0);
  double *lx = (double *)malloc(8 * (size_t)n1), *ly = (double *)malloc(8 * (size_t)n2);
  for (int i = 0; i < n1; i++) lx[i] = (s->kx == 'P' ? -2. * (1. - cos((2 * i) * pi / (1. * n1))) : -2. * (1. - cos((i) * pi / (1. * n1)))) * (s->dli[0] * s->dli[0]);
  for (int j = 0; j < n2; j++) ly[j] = (s->ky == 'P' ? -2. * (1. - cos((2 * j) * pi / (1. * n2))) : -2. * (1. - cos((j) * pi / (1. * n2)))) * (s->dli[1] * s->dli[1]);
  for (int j = 0; j < n2; j++) for (int i = 0; i < n1; i++) s->lambdaxy[i + (long)n1 * j] = lx[i] + ly[j];
  free(lx); free(ly);
  s->normfft = 1. / (((s->kx == 'P' ? 1. : 2.) * (n1 + 0.)) * ((s->ky == 'P' ? 1. : 2.) * (n2 + 0.)));
  if (s->zwall) { s->b[0] = s->b[0] + 1. * s->a[0]; s->b[n3 - 1] = s->b[n3 - 1] + 1. * s->c[n3 - 1]; }
  return 0;
}

/* Wall model (WM_LOG) on y and / or z walls in the general mode: lwm(0:1,3) in Fortran order, after cales_cpu_set_bc.
 * initbc (bound.f90:746-758): the wall-parallel components of a wall-model face become 'N', the normal one 'D';
 * index_wm 830-863. */
int cales_cpu_set_wm(void *h, const int *lwm, double hwm) {
  cpu_t *s = (cpu_t *)h;
  const int n1 = s->n1, n2 = s->n2, n3 = s->n3;
  if (!s->gen || lwm[0] || lwm[1]) return 1;                                     /* x walls: not restated */
  for (int q = 2; q < 6; q++) if (lwm[q] != 0 && lwm[q] != 1) return 1;
  s->hwm = hwm;
  for (int ib = 0; ib < 2; ib++) { s->lwm_y[ib] = lwm[ib + 2]; s->lwm[ib] = lwm[ib + 4]; }
  for (int idir = 1; idir < 3; idir++)
    for (int ib = 0; ib < 2; ib++)
      if (lwm[ib + 2 * idir]) for (int ivel = 0; ivel < 3; ivel++) s->cbcvel[ib][idir][ivel] = ivel == idir ? 'D' : 'N';
  s->bcu_y = (double *)calloc(2 * (size_t)(n1 + 2) * (size_t)(n3 + 2), 8); s->bcw_y = (double *)calloc(2 * (size_t)(n1 + 2) * (size_t)(n3 + 2), 8);
  s->bcu_z = (double *)calloc(2 * (size_t)(n1 + 2) * (size_t)(n2 + 2), 8); s->bcv_z = (double *)calloc(2 * (size_t)(n1 + 2) * (size_t)(n2 + 2), 8);
  s->wku = (double *)calloc((size_t)s->ntot, 8); s->wkv = (double *)calloc((size_t)s->ntot, 8); s->wkw = (double *)calloc((size_t)s->ntot, 8);
  if (s->lwm_y[0]) { int j = 1; while ((j - 0.5) * s->dl[1] < hwm) j = j + 1; s->index_wm_y[0] = j; }
  if (s->lwm_y[1]) { int j = n2; while ((n2 - j + 0.5) * s->dl[1] < hwm) j = j - 1; s->index_wm_y[1] = j; }
  if (s->lwm[0]) { int k = 1; while (s->zc[k] < hwm) k = k + 1; s->index_wm[0] = k; }
  if (s->lwm[1]) { int k = n3; while (s->l[2] - s->zc[k] < hwm) k = k - 1; s->index_wm[1] = k; }
  return 0;
}

double cales_cpu_forcing(void *h, int m) { return ((cpu_t *)h)->f[m]; }

double *cales_cpu_field(void *h, int which) { /* 0 u, 1 v, 2 w, 3 p, 4 pp, 5 visct, 6 s0 */
  cpu_t *s = (cpu_t *)h;
  double *f[] = {s->u, s->v, s->w, s->p, s->pp, s->visct, s->s0};
  return (which >= 0 && which < 7) ? f[which] : NULL;
}

double *cales_cpu_lambdaxy(void *h) { return ((cpu_t *)h)->lambdaxy; }

/* main.f90:370-375: ghost fill of the initial fields and the first eddy viscosity */
void cales_cpu_start(void *h) {
  cpu_t *s = (cpu_t *)h;
  fill_uvw(s, 0); fill_p(s, s->p, 'N');
  cmpt_sgs(s); fill_p(s, s->visct, 'D');
}

void cales_cpu_cmpt_sgs(void *h) { cmpt_sgs((cpu_t *)h); }
void cales_cpu_set_sgs(void *h, int dsmag) { ((cpu_t *)h)->dsmag = dsmag; }
void cales_cpu_boundp(void *h, int which) { fill_p((cpu_t *)h, cales_cpu_field(h, which), which == 5 ? 'D' : 'N'); }
void cales_cpu_solver(void *h) { solver((cpu_t *)h); }
void cales_cpu_fillps(void *h, double dti) { fillps((cpu_t *)h, dti); }
void cales_cpu_correc(void *h, double dt) { correc((cpu_t *)h, dt); }
void cales_cpu_rk(void *h, int irk, double dt) { rk_substep((cpu_t *)h, RKCOEFF[irk], dt); }

/* one time step = 3 substeps, main.f90:417-507 */
void cales_cpu_step(void *h, double dt) {
  cpu_t *s = (cpu_t *)h;
  for (int irk = 0; irk < 3; irk++) {
    const double dtrk = (RKCOEFF[irk][0] + RKCOEFF[irk][1]) * dt, dtrki = 1. / dtrk;
    rk_substep(s, RKCOEFF[irk], dt);                                             /* main.f90:420 */
    bulk_forcing(s);                                                             /* 422 */
    fill_uvw(s, 0);                                                              /* 493 */
    fillps(s, dtrki);                                                            /* 495 (updt_rhs_b 496: the rhsb planes of P and of N with value 0 are zero) */
    solver(s);                                                                   /* 497 */
    fill_p(s, s->pp, 'N');                                                       /* 498 */
    correc(s, dtrk);                                                             /* 499 */
    fill_uvw(s, 1);                                                              /* 500 */
    updatep(s);                                                                  /* 502 */
    fill_p(s, s->p, 'N');                                                        /* 503 */
    cmpt_sgs(s);                                                                 /* 504 */
    fill_p(s, s->visct, 'D');                                                    /* 506 */
  }
}

/* chkdiv.f90:27-47 (max part; the sum is order dependent and left to the numpy oracle) */
double cales_cpu_divmax(void *h) {
  cpu_t *s = (cpu_t *)h;
  double mx = 0.;
#pragma omp parallel for collapse(2) reduction(max : mx) schedule(static)
  for (int k = 1; k <= s->n3; k++)
    for (int j = 1; j <= s->n2; j++)
      for (int i = 1; i <= s->n1; i++) {
        const long c = IDX(s, i, j, k);
        double div = (s->w[c] - s->w[c - s->sk]) * s->dzfi[k] + (s->v[c] - s->v[c - s->sj]) * s->dli[1] + (s->u[c] - s->u[c - 1]) * s->dli[0];
        if (fabs(div) > mx) mx = fabs(div);
      }
  return mx;
}

/* chkdt.f90:40-97, explicit diffusion */
double cales_cpu_chkdt(void *h) {
  cpu_t *s = (cpu_t *)h;
  const long sj = s->sj, sk = s->sk;
  const double dxi = 1. / s->dl[0], dyi = 1. / s->dl[1], dl2i = dxi * dxi + dyi * dyi, visc = s->visc;
  const double *u = s->u, *v = s->v, *w = s->w, *t = s->visct;
  double dti = 0., dtid = 0.;
#pragma omp parallel for collapse(2) reduction(max : dti, dtid) schedule(static)
  for (int k = 1; k <= s->n3; k++)
    for (int j = 1; j <= s->n2; j++)
      for (int i = 1; i <= s->n1; i++) {
        const long c = IDX(s, i, j, k);
        const double dzfi_k = s->dzfi[k], dzci_k = s->dzci[k];
        double ux = fabs(u[c]), vx = 0.25 * fabs(v[c] + v[c - sj] + v[c + 1] + v[c + 1 - sj]), wx = 0.25 * fabs(w[c] + w[c - sk] + w[c + 1] + w[c + 1 - sk]);
        double uy = 0.25 * fabs(u[c] + u[c + sj] + u[c - 1 + sj] + u[c - 1]), vy = fabs(v[c]), wy = 0.25 * fabs(w[c] + w[c + sj] + w[c + sj - sk] + w[c - sk]);
        double uz = 0.25 * fabs(u[c] + u[c - 1] + u[c - 1 + sk] + u[c + sk]), vz = 0.25 * fabs(v[c] + v[c - sj] + v[c - sj + sk] + v[c + sk]), wz = fabs(w[c]);
        double dtix = ux * dxi + vx * dyi + wx * dzfi_k, dtiy = uy * dxi + vy * dyi + wy * dzfi_k, dtiz = uz * dxi + vz * dyi + wz * dzci_k;
        double m = fmax(dtix, fmax(dtiy, dtiz)); if (m > dti) dti = m;
        double dtidx = 0.5 * (t[c] + t[c + 1]) * (dl2i + dzfi_k * dzfi_k) + visc * dl2i + visc * (dzfi_k * dzfi_k);
        double dtidy = 0.5 * (t[c] + t[c + sj]) * (dl2i + dzfi_k * dzfi_k) + visc * dl2i + visc * (dzfi_k * dzfi_k);
        double dtidz = 0.5 * (t[c] + t[c + sk]) * (dl2i + dzci_k * dzci_k) + visc * dl2i + visc * (dzci_k * dzci_k);
        m = fmax(dtidx, fmax(dtidy, dtidz)); if (m > dtid) dtid = m;
      }
  if (dti == 0.) dti = 1.;
  if (dtid == 0.) dtid = s->eps;
  return fmin(0.4125 / dtid, 1.732 / dti);
}

int cales_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py sets the thread count explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers) */
void cales_cpu_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void cales_cpu_free(void *h) {
  cpu_t *s = (cpu_t *)h;
  if (!s) return;
  double *f[] = {s->u, s->v, s->w, s->p, s->pp, s->visct, s->s0, s->wk, s->dzc, s->dzf, s->dzci, s->dzfi, s->a, s->b, s->c, s->lambdaxy,
                 s->zc, s->zf, s->gvr_c, s->gvr_f, s->bcu_z, s->bcv_z, s->wku, s->wkv, s->bcu_y, s->bcw_y, s->wkw};
  for (size_t m = 0; m < sizeof(f) / sizeof(f[0]); m++) free(f[m]);
  for (int m = 0; m < 3; m++) { free(s->rhs[m]); free(s->rhso[m]); }
  for (int m = 0; m < 24; m++) free(s->dyn[m]);
  free(s->px.tw); free(s->py.tw); free(s->px.tq); free(s->py.tq); free(s);
}
