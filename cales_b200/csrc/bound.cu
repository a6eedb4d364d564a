// Boundary conditions, ghost fills and the wall model.
//   set_bc     src/bound.f90:202-399      bounduvw src/bound.f90:18-154    boundp src/bound.f90:156-200
//   cmpt_rhs_b src/bound.f90:447-560      updt_rhs_b src/bound.f90:562-617
//   updt_wallmodelbc / cmpt_wallmodelbc / vel_relative / wallmodel   src/wmodel.f90:19-335
// Ghost-fill ORDER is semantics (whole-plane array syntax incl. ghost rows, x then y then z, after
// all halo exchanges): one launch per direction fills both faces of every field of the call, so a
// bounduvw is 3 halo rounds + 3 fills (+ wall model) instead of ~20 tiny launches.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

#define MAXT 12
struct BcTask {
  double* p;           // field
  const double* bc;    // bc plane pair (0:m1+1, 0:m2+1, 0:1)
  const double* drp;   // device grid vector for dr (z faces) or nullptr
  double drv;          // host-known dr (x,y faces)
  int dri;             // index into drp
  char ctype;          // 'P','D','N'
  int ibound, centered;
};
struct BcBatch {
  int idir, nt;
  BcTask t[MAXT];
};

__global__ void __launch_bounds__(256) setbc_k(Dims d, BcBatch b) {
  const BcTask& t = b.t[blockIdx.z];
  const int idir = b.idir;
  // plane coordinates (a,b) over the two other directions, full extent 0..n+1
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2;
  const int m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const long sn = idir == 0 ? 1 : idir == 1 ? d.s1 : d.s2;       // stride along the normal
  const long base = idir == 0 ? d.idx(0, a, c) : idir == 1 ? d.idx(a, 0, c) : d.idx(a, c, 0);
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  double* p = t.p;
#define P(q) p[base + sn * (long)(q)]
  if (t.ctype == 'P') {                                          // bound.f90:232-248
    if (t.ibound == 0) { P(0) = P(n); P(n + 1) = P(1); }
    return;
  }
  const double bcv = t.bc[a + (long)m1 * (c + (long)m2 * t.ibound)];
  const double dr = t.drp ? t.drp[t.dri] : t.drv;
  double sgn = 1.;
  if (t.ctype == 'D' && t.centered) sgn = -1.;
  if (t.ctype == 'D') {                                          // bound.f90:250-319
    if (t.centered) {
      if (t.ibound == 0) P(0) = 2. * bcv + sgn * P(1);
      else P(n + 1) = 2. * bcv + sgn * P(n);
    } else {
      if (t.ibound == 0) P(0) = bcv;
      else { P(n + 1) = P(n - 1); P(n) = bcv; }
    }
  } else if (t.ctype == 'N') {                                   // bound.f90:320-396
    if (t.centered) {
      if (t.ibound == 0) P(0) = -dr * bcv + sgn * P(1);
      else P(n + 1) = dr * bcv + sgn * P(n);
    } else {
      if (t.ibound == 0) P(0) = -dr * bcv + P(1);
      else { P(n + 1) = P(n); P(n) = dr * bcv + P(n - 1); }
    }
  }
#undef P
}

static int run_batch(cales_ctx* ctx, const Dims& d, BcBatch& b) {
  if (b.nt == 0) return CALES_OK;
  const int m1 = b.idir == 0 ? d.n2 + 2 : d.n1 + 2;
  const int m2 = b.idir == 2 ? d.n2 + 2 : d.n3 + 2;
  setbc_k<<<dim3(cdiv(m1, 64), cdiv(m2, 4), b.nt), dim3(64, 4), 0, ctx->stream>>>(d, b);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// a periodic face pair is one task (ibound 0 does both planes); add it once
static void add_task(BcBatch& b, double* p, const double* bc, char ctype, int ibound, int centered, double drv, const double* drp, int dri) {
  if (ctype == 'P') {
    for (int q = 0; q < b.nt; ++q)
      if (b.t[q].p == p && b.t[q].ctype == 'P') return;
    ibound = 0;
  }
  BcTask& t = b.t[b.nt++];
  t.p = p; t.bc = bc; t.ctype = ctype; t.ibound = ibound; t.centered = centered; t.drv = drv; t.drp = drp; t.dri = dri;
}

static const double* plane_of(const cales_bound* b, int idir) { return idir == 0 ? b->x : idir == 1 ? b->y : b->z; }

// ---- fused ghost fill ------------------------------------------------------------------------------------------
// The reference fills ghosts by a fixed SEQUENCE of whole-plane operations (periodic self-copies of the halo phase in
// y, z, then set_bc in x, y, z; bound.f90:42-100,175-199), each of which reads cells the previous ones wrote -- that is
// how edges and corners get their values.  Every operation maps a ghost index of ONE direction to an interior index of
// that direction (copy / 2 bc - p / -+dr bc + p / constant bc), so the final value of any shell cell is the original
// value of one interior cell pushed through at most three such maps, applied in sequence order.  One launch evaluates
// that chain for every shell cell of every field: same arithmetic, same order, bit-identical -- instead of one launch
// per direction and phase.  (Not expressible this way: the face-centred Neumann top face, which reads a cell its own
// operation overwrites; the host keeps the per-direction launches for it.)
#define FG_MAXF 8
enum { FG_NONE = 0, FG_P = 1, FG_DC = 2, FG_DF = 3, FG_NC = 4, FG_NF = 5 };
struct FgRule { signed char kind[2]; int dri[2]; const double* bc; const double* drp; double drv; };
struct FgStep { int dir; FgRule r[FG_MAXF]; };
struct FgArgs { int ns, nf; double* p[FG_MAXF]; FgStep s[6]; };

__device__ __forceinline__ bool fg_eval(const FgArgs& A, const Dims& d, int f, int i, int j, int k, double& out) {
  int idx[3] = {i, j, k};
  const int nn[3] = {d.n1, d.n2, d.n3};
  int nops = 0, opk[3];
  double opc[3];
  bool konst = false;
  double cval = 0.;
  for (int s = A.ns - 1; s >= 0; --s) {
    const int dir = A.s[s].dir;
    const FgRule& R = A.s[s].r[f];
    const int q = idx[dir], n = nn[dir];
    int ib;
    if (q == 0) ib = 0; else if (q == n + 1 || q == n) ib = 1; else continue;
    const int kind = R.kind[ib];
    if (kind == FG_NONE || (q == n && kind != FG_DF)) continue;
    if (kind == FG_P) { idx[dir] = q == 0 ? n : 1; continue; }              // plain copy: no arithmetic
    if (kind == FG_DF && q == n + 1) { idx[dir] = n - 1; continue; }        // P(n+1) = P(n-1)
    const int a = dir == 0 ? idx[1] : idx[0], c = dir == 2 ? idx[1] : idx[2];
    const int m1 = dir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = dir == 2 ? d.n2 + 2 : d.n3 + 2;
    const double bcv = R.bc[a + (long)m1 * (c + (long)m2 * ib)];
    if (kind == FG_DF) { konst = true; cval = bcv; break; }                 // P(0) = bc, P(n) = bc
    const double dr = R.drp ? R.drp[R.dri[ib]] : R.drv;
    opk[nops] = kind;
    opc[nops] = kind == FG_DC ? 2. * bcv : (ib == 0 ? -dr * bcv : dr * bcv);
    ++nops;
    idx[dir] = ib == 0 ? 1 : n;
  }
  if (!konst && nops == 0 && idx[0] == i && idx[1] == j && idx[2] == k) return false;   // nothing writes this cell
  double v = konst ? cval : A.p[f][d.idx(idx[0], idx[1], idx[2])];
  for (int o = nops - 1; o >= 0; --o) v = opk[o] == FG_DC ? opc[o] + (-1.) * v : opc[o] + v;
  out = v;
  return true;
}

// slab 2: k in {0, n3, n3+1}, all (i,j); slab 1: j in {0, n2, n2+1}, k in 1..n3-1, all i; slab 0: i in {0, n1, n1+1},
// j in 1..n2-1, k in 1..n3-1: every shell cell exactly once.
__global__ void __launch_bounds__(256) fused_fill_k(Dims d, FgArgs A) {
  const long t = blockIdx.x * 256L + threadIdx.x;
  const int slab = blockIdx.y;
  int i, j, k;
  if (slab == 2) {
    if (t >= 3 * d.s2) return;
    const int kq = (int)(t / d.s2); const long r = t - kq * d.s2;
    j = (int)(r / d.s1); i = (int)(r - j * d.s1);
    k = kq == 0 ? 0 : d.n3 - 1 + kq;
  } else if (slab == 1) {
    const long per = (long)d.s1 * 3;
    if (t >= per * (d.n3 - 1)) return;
    const int kk = (int)(t / per); const long r = t - kk * per;
    const int jq = (int)(r / d.s1); i = (int)(r - jq * d.s1);
    j = jq == 0 ? 0 : d.n2 - 1 + jq; k = kk + 1;
  } else {
    const long per = 3L * (d.n2 - 1);
    if (t >= per * (d.n3 - 1)) return;
    const int kk = (int)(t / per); const long r = t - kk * per;
    const int jj = (int)(r / 3); const int iq = (int)(r - jj * 3);
    i = iq == 0 ? 0 : d.n1 - 1 + iq; j = jj + 1; k = kk + 1;
  }
  const long c = d.idx(i, j, k);
  for (int f = 0; f < A.nf; ++f) {
    double v;
    if (fg_eval(A, d, f, i, j, k, v)) A.p[f][c] = v;
  }
}

// Ghost fill of `nf` fields: halo phase (exchange with real neighbours, periodic self-copies) followed by the set_bc
// batches of the three directions.  Everything after the last direction that needs communication is one launch.
static int run_batch(cales_ctx* ctx, const Dims& d, BcBatch& b);
static int ghost_fill(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nf, BcBatch* batches /*[3]*/) {
  Dims d(n);
  int rc;
  static const bool nofuse = getenv("CALES_NO_FUSED_FILL") != nullptr;
  int self_mask = 0, comm_mask = 0, last_comm = -1;
  for (int idir = 0; idir < 3; ++idir) {
    if (idir + 1 == ctx->ipencil) continue;
    const int nb0 = nb[tb(0, idir)], nb1 = nb[tb(1, idir)];
    if (nb0 < 0 && nb1 < 0) continue;
    if (nb0 == ctx->rank && nb1 == ctx->rank) self_mask |= 1 << idir;
    else { comm_mask |= 1 << idir; last_comm = idir; }
  }
  bool ok = !nofuse && nf <= FG_MAXF && n[0] >= 2 && n[1] >= 2 && n[2] >= 2;
  for (int b = 0; b < 3 && ok; ++b)
    for (int q = 0; q < batches[b].nt; ++q) {
      const BcTask& t = batches[b].t[q];
      if (t.ctype == 'N' && !t.centered && t.ibound == 1) ok = false;
      bool known = false;
      for (int f = 0; f < nf; ++f) known |= fields[f] == t.p;
      if (!known) ok = false;
    }
  if (!ok) {
    if ((rc = k_halo_exchange(ctx, n, nb, fields, nf))) return rc;
    for (int b = 0; b < 3; ++b) if ((rc = run_batch(ctx, d, batches[b]))) return rc;
    return CALES_OK;
  }
  // halo phase up to and including the last communicating direction: as before
  int early = 0;
  for (int idir = 0; idir <= last_comm; ++idir) early |= 1 << idir;
  if (early && (rc = k_halo_exchange_dirs(ctx, n, nb, fields, nf, early))) return rc;
  FgArgs A;
  memset(&A, 0, sizeof A);
  A.nf = nf;
  for (int f = 0; f < nf; ++f) A.p[f] = fields[f];
  for (int idir = last_comm + 1; idir < 3; ++idir)
    if (self_mask & (1 << idir)) {
      FgStep& S = A.s[A.ns++];
      S.dir = idir;
      for (int f = 0; f < nf; ++f) S.r[f].kind[0] = S.r[f].kind[1] = FG_P;
    }
  for (int b = 0; b < 3; ++b) {
    if (batches[b].nt == 0) continue;
    FgStep& S = A.s[A.ns++];
    S.dir = batches[b].idir;
    for (int q = 0; q < batches[b].nt; ++q) {
      const BcTask& t = batches[b].t[q];
      int f = 0;
      while (fields[f] != t.p) ++f;
      FgRule& R = S.r[f];
      R.bc = t.bc; R.drp = t.drp; R.drv = t.drv;
      if (t.ctype == 'P') { R.kind[0] = R.kind[1] = FG_P; continue; }
      R.kind[t.ibound] = t.ctype == 'D' ? (t.centered ? FG_DC : FG_DF) : (t.centered ? FG_NC : FG_NF);
      R.dri[t.ibound] = t.dri;
    }
  }
  if (A.ns == 0) return CALES_OK;
  const long cnt = std::max(std::max(3 * d.s2, 3L * d.s1 * (n[2] - 1)), 3L * (n[1] - 1) * (n[2] - 1));
  fused_fill_k<<<dim3(cdiv(cnt, 256), 3), 256, 0, ctx->stream>>>(d, A);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// ---- wall model (wmodel.f90) --------------------------------------------------------------------------------
struct WmFace {
  int idir, ibound, mtype, index;
  double h, visc, l1d, dl_n;                // dl_n: dl(idir) for x,y
  const double *zc, *zf, *dzc;
  double lz;
  const double *vel1, *vel2;                // (v,w) | (u,w) | (u,v)
  double *bc1, *bc2;                        // bc planes of the two tangential components
  const double *mag1, *mag2;
};

__device__ __forceinline__ double vel_relative(double v1, double v2, double coef, double mag) {   // wmodel.f90:275-286
  double r = (1. - coef) * v1 + coef * v2;
  return r - mag;
}

__device__ void wallmodel(int mtype, double uh, double vh, double h, double l1d, double visc, double& t1, double& t2) {  // wmodel.f90:288-335
  const double kap_log = 0.41, b_log = 5.20, eps = 2.220446049250313e-16;
  const double upar = sqrt(uh * uh + vh * vh);
  double tauw_tot;
  if (mtype == 1) {
    double conv = 1.;
    double utau = fmax(sqrt(upar / h * visc), visc / h * exp(-kap_log * b_log));
    while (conv > 0.5e-4) {
      const double utau_old = utau;
      const double f = upar / utau - 1. / kap_log * log(h * utau / visc) - b_log;
      const double fp = -1. / utau * (upar / utau + 1. / kap_log);
      utau = fabs(utau - f / fp);
      conv = fabs(utau / utau_old - 1.);
    }
    tauw_tot = utau * utau;
  } else {
    const double del = 0.5 * l1d;
    const double umax = upar / (h / del * (2. - h / del));
    tauw_tot = 2. / del * umax * visc;
  }
  t1 = tauw_tot * uh / (upar + eps);
  t2 = tauw_tot * vh / (upar + eps);
}

// one thread per plane point and per component (blockIdx.z = 0: first tangential component, 1: second)
__global__ void __launch_bounds__(256) wm_k(Dims d, WmFace f) {
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  const int comp = blockIdx.z;
  const double visci = 1. / f.visc;
  const long s1 = d.s1, s2 = d.s2;
  const double* A = f.vel1; const double* B = f.vel2;
  double coef, sgn;
  int q1, q2;   // near-wall and far layers
  if (f.idir == 0) {                                        // wmodel.f90:118-166
    if (f.ibound == 0) { q2 = f.index; q1 = f.index - 1; coef = (f.h - (q1 - 0.5) * f.dl_n) / f.dl_n; sgn = 1.; }
    else { q2 = f.index; q1 = f.index + 1; coef = (f.h - (d.n1 - q1 + 0.5) * f.dl_n) / f.dl_n; sgn = -1.; }
    const int m1 = d.n2 + 2, m2 = d.n3 + 2;
    const int j = a, k = c;
    if (comp == 0) {
      if (j > d.n2 || k < 1 || k > d.n3) return;
      const double v1 = A[d.idx(q1, j, k)], v2 = A[d.idx(q2, j, k)];
      const double w1 = 0.25 * (B[d.idx(q1, j, k)] + B[d.idx(q1, j + 1, k)] + B[d.idx(q1, j, k - 1)] + B[d.idx(q1, j + 1, k - 1)]);
      const double w2 = 0.25 * (B[d.idx(q2, j, k)] + B[d.idx(q2, j + 1, k)] + B[d.idx(q2, j, k - 1)] + B[d.idx(q2, j + 1, k - 1)]);
      const long o = (long)m1 * m2 * f.ibound;
      const double v_mag = f.mag1[j + m1 * k + o];
      const double w_mag = 0.25 * (f.mag2[j + m1 * k + o] + f.mag2[j + 1 + m1 * k + o] + f.mag2[j + m1 * (k - 1) + o] + f.mag2[j + 1 + m1 * (k - 1) + o]);
      double t1, t2;
      wallmodel(f.mtype, vel_relative(v1, v2, coef, v_mag), vel_relative(w1, w2, coef, w_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc1[j + m1 * k + o] = sgn * visci * t1;
    } else {
      if (j < 1 || j > d.n2 || k > d.n3) return;
      const double wei = (f.zf[k] - f.zc[k]) / f.dzc[k];
      const double v1 = 0.5 * ((1. - wei) * (A[d.idx(q1, j - 1, k)] + A[d.idx(q1, j, k)]) + wei * (A[d.idx(q1, j - 1, k + 1)] + A[d.idx(q1, j, k + 1)]));
      const double v2 = 0.5 * ((1. - wei) * (A[d.idx(q2, j - 1, k)] + A[d.idx(q2, j, k)]) + wei * (A[d.idx(q2, j - 1, k + 1)] + A[d.idx(q2, j, k + 1)]));
      const double w1 = B[d.idx(q1, j, k)], w2 = B[d.idx(q2, j, k)];
      const long o = (long)m1 * m2 * f.ibound;
      const double v_mag = 0.5 * ((1. - wei) * (f.mag1[j - 1 + m1 * k + o] + f.mag1[j + m1 * k + o]) +
                                  wei * (f.mag1[j - 1 + m1 * (k + 1) + o] + f.mag1[j + m1 * (k + 1) + o]));
      const double w_mag = f.mag2[j + m1 * k + o];
      double t1, t2;
      wallmodel(f.mtype, vel_relative(v1, v2, coef, v_mag), vel_relative(w1, w2, coef, w_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc2[j + m1 * k + o] = sgn * visci * t2;
    }
  } else if (f.idir == 1) {                                 // wmodel.f90:167-217
    if (f.ibound == 0) { q2 = f.index; q1 = f.index - 1; coef = (f.h - (q1 - 0.5) * f.dl_n) / f.dl_n; sgn = 1.; }
    else { q2 = f.index; q1 = f.index + 1; coef = (f.h - (d.n2 - q1 + 0.5) * f.dl_n) / f.dl_n; sgn = -1.; }
    const int m1 = d.n1 + 2, m2 = d.n3 + 2;
    const int i = a, k = c;
    if (comp == 0) {
      if (i > d.n1 || k < 1 || k > d.n3) return;
      const double u1 = A[d.idx(i, q1, k)], u2 = A[d.idx(i, q2, k)];
      const double w1 = 0.25 * (B[d.idx(i, q1, k)] + B[d.idx(i + 1, q1, k)] + B[d.idx(i, q1, k - 1)] + B[d.idx(i + 1, q1, k - 1)]);
      const double w2 = 0.25 * (B[d.idx(i, q2, k)] + B[d.idx(i + 1, q2, k)] + B[d.idx(i, q2, k - 1)] + B[d.idx(i + 1, q2, k - 1)]);
      const long o = (long)m1 * m2 * f.ibound;
      const double u_mag = f.mag1[i + m1 * k + o];
      const double w_mag = 0.25 * (f.mag2[i + m1 * k + o] + f.mag2[i + 1 + m1 * k + o] + f.mag2[i + m1 * (k - 1) + o] + f.mag2[i + 1 + m1 * (k - 1) + o]);
      double t1, t2;
      wallmodel(f.mtype, vel_relative(u1, u2, coef, u_mag), vel_relative(w1, w2, coef, w_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc1[i + m1 * k + o] = sgn * visci * t1;
    } else {
      if (i < 1 || i > d.n1 || k > d.n3) return;
      const double wei = (f.zf[k] - f.zc[k]) / f.dzc[k];
      const double u1 = 0.5 * ((1. - wei) * (A[d.idx(i - 1, q1, k)] + A[d.idx(i, q1, k)]) + wei * (A[d.idx(i - 1, q1, k + 1)] + A[d.idx(i, q1, k + 1)]));
      const double u2 = 0.5 * ((1. - wei) * (A[d.idx(i - 1, q2, k)] + A[d.idx(i, q2, k)]) + wei * (A[d.idx(i - 1, q2, k + 1)] + A[d.idx(i, q2, k + 1)]));
      const double w1 = B[d.idx(i, q1, k)], w2 = B[d.idx(i, q2, k)];
      const long o = (long)m1 * m2 * f.ibound;
      const double u_mag = 0.5 * ((1. - wei) * (f.mag1[i - 1 + m1 * k + o] + f.mag1[i + m1 * k + o]) +
                                  wei * (f.mag1[i - 1 + m1 * (k + 1) + o] + f.mag1[i + m1 * (k + 1) + o]));
      const double w_mag = f.mag2[i + m1 * k + o];
      double t1, t2;
      wallmodel(f.mtype, vel_relative(u1, u2, coef, u_mag), vel_relative(w1, w2, coef, w_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc2[i + m1 * k + o] = sgn * visci * t2;
    }
  } else {                                                  // wmodel.f90:218-272
    if (f.ibound == 0) { q2 = f.index; q1 = f.index - 1; coef = (f.h - f.zc[q1]) / f.dzc[q1]; sgn = 1.; }
    else { q2 = f.index; q1 = f.index + 1; coef = (f.h - (f.lz - f.zc[q1])) / (f.dzc[q2]); sgn = -1.; }
    const int m1 = d.n1 + 2, m2 = d.n2 + 2;
    const int i = a, j = c;
    const long o = (long)m1 * m2 * f.ibound;
    if (comp == 0) {
      if (i > d.n1 || j < 1 || j > d.n2) return;
      const double u1 = A[d.idx(i, j, q1)], u2 = A[d.idx(i, j, q2)];
      const double v1 = 0.25 * (B[d.idx(i, j, q1)] + B[d.idx(i + 1, j, q1)] + B[d.idx(i, j - 1, q1)] + B[d.idx(i + 1, j - 1, q1)]);
      const double v2 = 0.25 * (B[d.idx(i, j, q2)] + B[d.idx(i + 1, j, q2)] + B[d.idx(i, j - 1, q2)] + B[d.idx(i + 1, j - 1, q2)]);
      const double u_mag = f.mag1[i + m1 * j + o];
      const double v_mag = 0.25 * (f.mag2[i + m1 * j + o] + f.mag2[i + 1 + m1 * j + o] + f.mag2[i + m1 * (j - 1) + o] + f.mag2[i + 1 + m1 * (j - 1) + o]);
      double t1, t2;
      wallmodel(f.mtype, vel_relative(u1, u2, coef, u_mag), vel_relative(v1, v2, coef, v_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc1[i + m1 * j + o] = sgn * visci * t1;
    } else {
      if (i < 1 || i > d.n1 || j > d.n2) return;
      const double u1 = 0.25 * (A[d.idx(i - 1, j, q1)] + A[d.idx(i, j, q1)] + A[d.idx(i - 1, j + 1, q1)] + A[d.idx(i, j + 1, q1)]);
      const double u2 = 0.25 * (A[d.idx(i - 1, j, q2)] + A[d.idx(i, j, q2)] + A[d.idx(i - 1, j + 1, q2)] + A[d.idx(i, j + 1, q2)]);
      const double v1 = B[d.idx(i, j, q1)], v2 = B[d.idx(i, j, q2)];
      const double u_mag = 0.25 * (f.mag1[i - 1 + m1 * j + o] + f.mag1[i + m1 * j + o] + f.mag1[i - 1 + m1 * (j + 1) + o] + f.mag1[i + m1 * (j + 1) + o]);
      const double v_mag = f.mag2[i + m1 * j + o];
      double t1, t2;
      wallmodel(f.mtype, vel_relative(u1, u2, coef, u_mag), vel_relative(v1, v2, coef, v_mag), f.h, f.l1d, f.visc, t1, t2);
      f.bc2[i + m1 * j + o] = sgn * visci * t2;
    }
  }
}

// ---- bounduvw / boundp ------------------------------------------------------------------------------------------
extern "C" int cales_bounduvw(cales_ctx* ctx, const char cbc[18], const int n[3], const cales_bound* bcu, const cales_bound* bcv,
                              const cales_bound* bcw, const cales_bound* bcu_mag, const cales_bound* bcv_mag,
                              const cales_bound* bcw_mag, const int nb[6], const int is_bound[6], const int lwm[6],
                              const double l[3], const double dl[3], const double* zc, const double* zf, const double* dzc,
                              const double* dzf, double visc, double h, const int index_wm[6], int is_updt_wm, int is_correc,
                              double* u, double* v, double* w) {
  CHECK_CTX(ctx);
  Dims d(n);
  double* vel[3] = {u, v, w};
  const cales_bound* bcs[3] = {bcu, bcv, bcw};
  const cales_bound* mags[3] = {bcu_mag, bcv_mag, bcw_mag};
  int rc;
  BcBatch bb[3];
#define CBC(ib, idir, ivel) cbc[(ib) + 2 * (idir) + 6 * (ivel)]
  for (int idir = 0; idir < 3; ++idir) {                    // bound.f90:56-100
    BcBatch& b = bb[idir]; b.idir = idir; b.nt = 0;
    const bool impose_norm_bc = (!is_correc) || (CBC(0, idir, idir) == 'P' && CBC(1, idir, idir) == 'P');
    for (int ib = 0; ib < 2; ++ib) {
      if (!is_bound[tb(ib, idir)]) continue;
      const int kk = ib == 0 ? 0 : n[2];
      if (impose_norm_bc)
        add_task(b, vel[idir], plane_of(bcs[idir], idir), CBC(ib, idir, idir), ib, 0, dl[idir < 2 ? idir : 0], idir == 2 ? dzf : nullptr, kk);
      if (lwm[tb(ib, idir)] == 0)
        for (int c = 0; c < 3; ++c)
          if (c != idir) add_task(b, vel[c], plane_of(bcs[c], idir), CBC(ib, idir, c), ib, 1, dl[idir < 2 ? idir : 0], idir == 2 ? dzc : nullptr, kk);
    }
  }
  if ((rc = ghost_fill(ctx, n, nb, vel, 3, bb))) return rc;   // halo exchange (bound.f90:42-52) + the three set_bc rounds
  bool any_wm = false;
  for (int q = 0; q < 6; ++q) any_wm |= (is_bound[q] && lwm[q] != 0);
  if (!any_wm) return CALES_OK;
  if (is_updt_wm) {                                         // bound.f90:120-123 -> wmodel.f90:19-63
    for (int idir = 0; idir < 3; ++idir)
      for (int ib = 0; ib < 2; ++ib) {
        if (!(is_bound[tb(ib, idir)] && lwm[tb(ib, idir)] != 0)) continue;
        int c1 = -1, c2 = -1;
        for (int c = 0; c < 3; ++c) if (c != idir) { if (c1 < 0) c1 = c; else c2 = c; }
        WmFace f;
        f.idir = idir; f.ibound = ib; f.mtype = lwm[tb(ib, idir)]; f.index = index_wm[tb(ib, idir)];
        f.h = h; f.visc = visc; f.l1d = l[idir]; f.dl_n = dl[idir < 2 ? idir : 0];
        f.zc = zc; f.zf = zf; f.dzc = dzc; f.lz = l[2];
        f.vel1 = vel[c1]; f.vel2 = vel[c2];
        f.bc1 = (double*)plane_of(bcs[c1], idir); f.bc2 = (double*)plane_of(bcs[c2], idir);
        f.mag1 = plane_of(mags[c1], idir); f.mag2 = plane_of(mags[c2], idir);
        const int m1 = idir == 0 ? n[1] + 2 : n[0] + 2;
        const int m2 = idir == 2 ? n[1] + 2 : n[2] + 2;
        wm_k<<<dim3(cdiv(m1, 64), cdiv(m2, 4), 2), dim3(64, 4), 0, ctx->stream>>>(d, f);
        KERNEL_CHECK(ctx);
      }
  }
  for (int idir = 0; idir < 3; ++idir) {                    // bound.f90:125-148
    BcBatch b; b.idir = idir; b.nt = 0;
    for (int ib = 0; ib < 2; ++ib) {
      if (!(is_bound[tb(ib, idir)] && lwm[tb(ib, idir)] != 0)) continue;
      const int kk = ib == 0 ? 0 : n[2];
      for (int c = 0; c < 3; ++c)
        if (c != idir) add_task(b, vel[c], plane_of(bcs[c], idir), CBC(ib, idir, c), ib, 1, dl[idir < 2 ? idir : 0], idir == 2 ? dzc : nullptr, kk);
    }
    if ((rc = run_batch(ctx, d, b))) return rc;
  }
#undef CBC
  return CALES_OK;
}

int k_boundp_multi(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
                   const int is_bound[6], const double dl[3], const double* dzc, double* const* ps, int np) {
  Dims d(n);
  int rc;
  if (np <= MAXT / 2) {
    BcBatch bb[3];
    for (int idir = 0; idir < 3; ++idir) {                  // bound.f90:181-199
      BcBatch& b = bb[idir]; b.idir = idir; b.nt = 0;
      for (int f = 0; f < np; ++f)
        for (int ib = 0; ib < 2; ++ib)
          if (is_bound[tb(ib, idir)])
            add_task(b, ps[f], plane_of(bcp, idir), cbc[tb(ib, idir)], ib, 1, dl[idir < 2 ? idir : 0], idir == 2 ? dzc : nullptr, ib == 0 ? 0 : n[2]);
    }
    return ghost_fill(ctx, n, nb, ps, np, bb);
  }
  if ((rc = k_halo_exchange(ctx, n, nb, ps, np))) return rc;             // bound.f90:175-180
  for (int idir = 0; idir < 3; ++idir) {                    // bound.f90:181-199
    for (int f0 = 0; f0 < np; f0 += MAXT / 2) {
      BcBatch b; b.idir = idir; b.nt = 0;
      for (int f = f0; f < np && f < f0 + MAXT / 2; ++f)
        for (int ib = 0; ib < 2; ++ib)
          if (is_bound[tb(ib, idir)])
            add_task(b, ps[f], plane_of(bcp, idir), cbc[tb(ib, idir)], ib, 1, dl[idir < 2 ? idir : 0], idir == 2 ? dzc : nullptr, ib == 0 ? 0 : n[2]);
      if ((rc = run_batch(ctx, d, b))) return rc;
    }
  }
  return CALES_OK;
}

int k_boundp(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
             const int is_bound[6], const double dl[3], const double* dzc, double* p) {
  double* ps[1] = {p};
  return k_boundp_multi(ctx, cbc, n, bcp, nb, is_bound, dl, dzc, ps, 1);
}

extern "C" int cales_boundp(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
                            const int is_bound[6], const double dl[3], const double* dzc, double* p) {
  CHECK_CTX(ctx);
  return k_boundp(ctx, cbc, n, bcp, nb, is_bound, dl, dzc, p);
}

// ---- rhs boundary planes -------------------------------------------------------------------------------------------
// bc_rhs (bound.f90:497-560): rhs(m1,m2,0:1) from bc(0:m1+1,0:m2+1,0:1)
__global__ void bc_rhs_k(int m1, int m2, const double* __restrict__ bc, double* __restrict__ rhs, char c0, char c1, char cf,
                         double dlc0, double dlc1, double dlf0, double dlf1) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int ib = blockIdx.z;
  if (a >= m1) return;
  const char c = ib == 0 ? c0 : c1;
  const double dlc = ib == 0 ? dlc0 : dlc1, dlf = ib == 0 ? dlf0 : dlf1;
  const double bv = bc[(a + 1) + (long)(m1 + 2) * ((b + 1) + (long)(m2 + 2) * ib)];
  double r = 0.;
  if (cf == 'c') {
    if (c == 'D') r = -2. * bv / dlc / dlf;
    else if (c == 'N') r = (ib == 0 ? 1. : -1.) * bv / dlf;
  } else {
    if (c == 'D') r = -bv / dlc / dlf;
    else if (c == 'N') r = (ib == 0 ? 1. : -1.) * bv / dlc;
  }
  rhs[a + (long)m1 * (b + (long)m2 * ib)] = r;
}

extern "C" int cales_cmpt_rhs_b(cales_ctx* ctx, const int ng[3], const int n[3], const double dl[3], const double* dzc_g,
                                const double* dzf_g, const char cbc[6], const cales_bound* bc, const char c_or_f[3],
                                double* rhsbx, double* rhsby, double* rhsbz) {
  CHECK_CTX(ctx);
  // bound.f90:466-477 (the z metrics are the GLOBAL ones)
  const double dzc01_c[2] = {dzc_g[0], dzc_g[ng[2]]}, dzf01_c[2] = {dzf_g[1], dzf_g[ng[2]]};
  const double dzc01_f[2] = {dzc_g[1], dzc_g[ng[2] - 1]}, dzf01_f[2] = {dzf_g[1], dzf_g[ng[2]]};
  if (rhsbx) {
    bc_rhs_k<<<dim3(cdiv(n[1], 128), n[2], 2), 128, 0, ctx->stream>>>(n[1], n[2], bc->x, rhsbx, cbc[tb(0, 0)], cbc[tb(1, 0)], c_or_f[0], dl[0], dl[0], dl[0], dl[0]);
    KERNEL_CHECK(ctx);
  }
  if (rhsby) {
    bc_rhs_k<<<dim3(cdiv(n[0], 128), n[2], 2), 128, 0, ctx->stream>>>(n[0], n[2], bc->y, rhsby, cbc[tb(0, 1)], cbc[tb(1, 1)], c_or_f[1], dl[1], dl[1], dl[1], dl[1]);
    KERNEL_CHECK(ctx);
  }
  if (rhsbz) {
    const double* zc01 = c_or_f[2] == 'c' ? dzc01_c : dzc01_f;
    const double* zf01 = c_or_f[2] == 'c' ? dzf01_c : dzf01_f;
    bc_rhs_k<<<dim3(cdiv(n[0], 128), n[1], 2), 128, 0, ctx->stream>>>(n[0], n[1], bc->z, rhsbz, cbc[tb(0, 2)], cbc[tb(1, 2)], c_or_f[2], zc01[0], zc01[1], zf01[0], zf01[1]);
    KERNEL_CHECK(ctx);
  }
  return CALES_OK;
}

// updt_rhs_b (bound.f90:562-617): p(face plane) += rhsb(:,:,ib)
__global__ void updt_rhs_k(Dims d, int idir, int pos, int ib, const double* __restrict__ rhs, double* __restrict__ p) {
  const int m1 = idir == 0 ? d.n2 : d.n1, m2 = idir == 2 ? d.n2 : d.n3;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= m1) return;
  const long c = idir == 0 ? d.idx(pos, a + 1, b + 1) : idir == 1 ? d.idx(a + 1, pos, b + 1) : d.idx(a + 1, b + 1, pos);
  p[c] = p[c] + rhs[a + (long)m1 * (b + (long)m2 * ib)];
}

extern "C" int cales_updt_rhs_b(cales_ctx* ctx, const char c_or_f[3], const char cbc[6], const int n[3], const int is_bound[6],
                                const double* rhsbx, const double* rhsby, const double* rhsbz, double* p) {
  CHECK_CTX(ctx);
  Dims d(n);
  const double* rhs[3] = {rhsbx, rhsby, rhsbz};
  for (int idir = 0; idir < 3; ++idir) {
    if (!rhs[idir]) continue;
    const int q = (c_or_f[idir] == 'f' && cbc[tb(1, idir)] == 'D') ? 1 : 0;
    const int m1 = idir == 0 ? n[1] : n[0], m2 = idir == 2 ? n[1] : n[2];
    for (int ib = 0; ib < 2; ++ib) {
      if (!is_bound[tb(ib, idir)]) continue;
      const int pos = ib == 0 ? 1 : n[idir] - q;
      updt_rhs_k<<<dim3(cdiv(m1, 128), m2), 128, 0, ctx->stream>>>(d, idir, pos, ib, rhs[idir], p);
      KERNEL_CHECK(ctx);
    }
  }
  return CALES_OK;
}
