// Register-blocked batched real transforms (power-of-two lengths): the fast path of the solver's x/y passes.
// Same transforms and conventions as fft.cu (R2HC/HC2R, REDFT10/01, RODFT10/01 of src/fft.f90:192-245, FFTW
// definitions, unnormalised), different machine mapping:
//   * lanes run ACROSS lines (NL = 16 neighbouring lines per CTA), threadIdx.y = t walks along the transform, so every
//     shared-memory access of the butterflies is a contiguous 16 x 16 B row (conflict-free by construction) and every
//     index, twiddle address and branch is warp-uniform;
//   * the transform length, the radix schedule, the direction and the transform family are template parameters, so
//     all Stockham positions and twiddle indices fold to shifts and immediates;
//   * each thread keeps E = 8 or 16 complex points in registers and performs radix-16/8/4/2 butterflies on them;
//     stages exchange data through ONE in-place shared buffer (read -> sync -> butterfly -> write -> sync); stage
//     twiddles come from a shared-memory copy of the table (warp-uniform broadcast reads);
//   * y-lines (stride = row pitch): the NL lines of a CTA are NL consecutive x, so the first stage loads straight
//     from global memory (128 B rows) and the even/odd split / Makhoul post-stage stores straight to global memory:
//     one read + one write of the array, 4 shared-memory passes;
//   * x-lines (contiguous): the phases that touch global memory use a second thread mapping (lanes ALONG the line,
//     coalesced) and the shared buffer is XOR-swizzled so both mappings are conflict-free: same traffic as y-lines.
#pragma once
#include <cstdlib>
#include "common.cuh"

enum { KB_PP = 0, KB_NN = 1, KB_DD = 2 };

struct FftBArgs {
  const double* in; double* out;
  long ies, il1, il2, oes, ol1, ol2;
  int nl1, nl2, dd;        // lines along l1 / l2; dd: DST flavour of the Makhoul family
  double scale;
  const double2 *wm, *wn, *h4;
  // peer-fused y -> z transpose (forward y-lines only): the spectral line is scattered straight into the Z-pencils of
  // the ranks owning each slab of ky (local or over NVLink); element (x, ky, z) of rank r lives at
  // pbase[r] + x + nx (ky - pys[r]) + nx pny[r] (zoff + z_local)
  int np, zoff, nx;
  double* pbase[8];
  int pys[9], pny[8];
  // fillps fused into the forward x pass (su != nullptr): the line element is not read from `in` but computed from the
  // haloed velocity arrays with the strides il1, il2 of `in` -- fillps (src/fillps.f90:14-48) followed by updt_rhs_b
  // (src/bound.f90:562-617; rb*: the rhsb planes, bnd: is_bound) -- so the right-hand side never makes a round trip to memory
  const double *su, *sv, *sw, *sdzfi, *rbx, *rby, *rbz;
  double dti, dtidxi, dtidyi;
  int bnd[6], sn1, sn2, sn3;
};

// element i0 (0-based) of line (j, k) (1-based) of the pressure right-hand side; off = offset of the line's first element
__device__ __forceinline__ double fftb_div_src(const FftBArgs& A, long off, int i0, int j, int k, double dzfi_k) {
  const long c = off + i0;
  double val = ((A.sw[c] - A.sw[c - A.il2]) * A.dti * dzfi_k + (A.sv[c] - A.sv[c - A.il1]) * A.dtidyi + (A.su[c] - A.su[c - 1]) * A.dtidxi);
  if (A.bnd[0] && i0 == 0) val = val + A.rbx[(j - 1) + (long)A.sn2 * (k - 1)];
  if (A.bnd[1] && i0 == A.sn1 - 1) val = val + A.rbx[(j - 1) + (long)A.sn2 * ((k - 1) + (long)A.sn3)];
  if (A.bnd[2] && j == 1) val = val + A.rby[i0 + (long)A.sn1 * (k - 1)];
  if (A.bnd[3] && j == A.sn2) val = val + A.rby[i0 + (long)A.sn1 * ((k - 1) + (long)A.sn3)];
  if (A.bnd[4] && k == 1) val = val + A.rbz[i0 + (long)A.sn1 * (j - 1)];
  if (A.bnd[5] && k == A.sn3) val = val + A.rbz[i0 + (long)A.sn1 * ((j - 1) + (long)A.sn2)];
  return val;
}

#define NLB 16

namespace fb {
__device__ __forceinline__ double2 mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 conj(double2 a) { return make_double2(a.x, -a.y); }
template <bool INV> __device__ __forceinline__ double2 muli(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

__device__ __forceinline__ void dft2(double2& a, double2& b) { const double2 t = a; a = add(t, b); b = sub(t, b); }
template <bool INV> __device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3) {
  const double2 t0 = add(a0, a2), t1 = sub(a0, a2), t2 = add(a1, a3), t3 = muli<INV>(sub(a1, a3));
  a0 = add(t0, t2); a1 = add(t1, t3); a2 = sub(t0, t2); a3 = sub(t1, t3);
}
#define FB_RH 0.70710678118654752440
#define FB_C1 0.92387953251128675613
#define FB_S1 0.38268343236508977173
// o * exp(-+ i pi K/8)
template <bool INV, int K> __device__ __forceinline__ double2 mw16(double2 o) {
  if (K == 0) return o;
  if (K == 4) return muli<INV>(o);
  if (K == 2) return INV ? make_double2(FB_RH * (o.x - o.y), FB_RH * (o.x + o.y)) : make_double2(FB_RH * (o.x + o.y), FB_RH * (o.y - o.x));
  if (K == 6) return INV ? make_double2(FB_RH * (-o.x - o.y), FB_RH * (o.x - o.y)) : make_double2(FB_RH * (o.y - o.x), FB_RH * (-o.x - o.y));
  constexpr double c = K == 1 ? FB_C1 : K == 3 ? FB_S1 : -FB_C1;                 // K = 1, 3, 9
  constexpr double s0 = K == 1 ? FB_S1 : K == 3 ? FB_C1 : -FB_S1;
  constexpr double s = INV ? -s0 : s0;                                           // forward: c - i s
  return make_double2(o.x * c + o.y * s, o.y * c - o.x * s);
}
template <bool INV> __device__ __forceinline__ void dft8(double2 (&v)[8]) {
  dft4<INV>(v[0], v[2], v[4], v[6]);
  dft4<INV>(v[1], v[3], v[5], v[7]);
  const double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  const double2 o0 = v[1], o1 = mw16<INV, 2>(v[3]), o2 = muli<INV>(v[5]), o3 = mw16<INV, 6>(v[7]);
  v[0] = add(e0, o0); v[4] = sub(e0, o0);
  v[1] = add(e1, o1); v[5] = sub(e1, o1);
  v[2] = add(e2, o2); v[6] = sub(e2, o2);
  v[3] = add(e3, o3); v[7] = sub(e3, o3);
}
template <bool INV> __device__ __forceinline__ void dft16(double2 (&v)[16]) {
  // a = 4 a1 + a0: 4-point DFTs over a1 for each a0, twiddle w16^(a0 p), 4-point DFTs over a0 for each p
#pragma unroll
  for (int a0 = 0; a0 < 4; ++a0) dft4<INV>(v[a0], v[4 + a0], v[8 + a0], v[12 + a0]);     // Y_{a0}(p) in v[4p + a0]
  v[5] = mw16<INV, 1>(v[5]); v[6] = mw16<INV, 2>(v[6]); v[7] = mw16<INV, 3>(v[7]);
  v[9] = mw16<INV, 2>(v[9]); v[10] = mw16<INV, 4>(v[10]); v[11] = mw16<INV, 6>(v[11]);
  v[13] = mw16<INV, 3>(v[13]); v[14] = mw16<INV, 6>(v[14]); v[15] = mw16<INV, 9>(v[15]);
#pragma unroll
  for (int p = 0; p < 4; ++p) dft4<INV>(v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);  // X_{p + 4r} in v[4p + r]
  double2 o[16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int r = 0; r < 4; ++r) o[p + 4 * r] = v[4 * p + r];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = o[q];
}
template <int R, bool INV> __device__ __forceinline__ void dftR(double2 (&v)[R]) {
  if constexpr (R == 2) dft2(v[0], v[1]);
  else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
  else if constexpr (R == 8) dft8<INV>(v);
  else dft16<INV>(v);
}

// One Stockham stage of radix R (Ns = product of the previous radices) on the E register points of thread t (point e
// sits at complex index t + T e): butterfly b handles j = t + b T with inputs j + a M/R = x[b + a E/R]; output q goes
// to position (j - k) R + k + q Ns, k = j mod Ns.  sw: stage twiddles exp(-2 pi i t / M) in shared memory.
template <int M, int E, int R, int Ns, bool INV>
__device__ __forceinline__ void stage(double2 (&x)[E], int (&pos)[E], int t, const double2* __restrict__ sw) {
  constexpr int NB = E / R, T = M / E, tstep = M / (Ns * R);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = t + b * T;
    const int k = j & (Ns - 1);
    double2 v[R];
#pragma unroll
    for (int a = 0; a < R; ++a) v[a] = x[b + a * NB];
    if (Ns > 1) {
      const int ts = k * tstep;
#pragma unroll
      for (int a = 1; a < R; ++a) {
        double2 w = sw[a * ts];
        if (INV) w.y = -w.y;
        v[a] = mul(v[a], w);
      }
    }
    dftR<R, INV>(v);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) { x[b + q * NB] = v[q]; pos[b + q * NB] = base + q * Ns; }
  }
}

// Makhoul: line element that holds sample vi of the permuted sequence v (v_j = x_2j, v_{n-1-j} = x_{2j+1})
template <int MK> __device__ __forceinline__ int src_of(int vi, int n) {
  if (!MK) return vi;
  return vi < (n >> 1) ? 2 * vi : 2 * (n - 1 - vi) + 1;
}
template <int MK> __device__ __forceinline__ int slot_of(int e, int n) {       // inverse of src_of
  if (!MK) return e;
  return (e & 1) ? n - 1 - (e >> 1) : (e >> 1);
}
}  // namespace fb

// M complex points per line (n = 2M reals), E points per thread, XD: x-lines / y-lines, INV: backward transform,
// MK: Makhoul family (DCT/DST) instead of the periodic real FFT.
// Two thread mappings are used.  "Standard" (ls, ts): lanes across the NL lines, ts along the transform -- the
// butterfly stages in the middle.  "Global" (lx, tx): the mapping of every phase that touches global memory (first
// forward stage, forward post-stage, backward pre-stage, last backward stage).  For y-lines the two coincide (the NL
// lines are NL consecutive x, so lanes across lines ARE coalesced); for x-lines lanes run along the line (tx fastest)
// so that global accesses are coalesced, and the shared buffer is XOR-swizzled so that both mappings are conflict-free.
template <int M, int E, int XD, bool INV, int MK, bool PEER = false, bool DIVSRC = false>
__global__ void __launch_bounds__(NLB*(M / E)) fftb_k(FftBArgs A) {
  using namespace fb;
  extern __shared__ double2 S[];
  constexpr int n = 2 * M, T = M / E, NT = NLB * T;
  double2* sw = S + M * NLB;                  // stage twiddles, M entries
  const int ls = threadIdx.x, ts = threadIdx.y;
  const int tid = ls + NLB * ts;
  const int lx = XD ? tid / T : ls, tx = XD ? tid % T : ts;
  const int L0 = blockIdx.x * NLB;
  const int nl = min(NLB, A.nl1 - L0);
  const bool on = lx < nl;
  const bool dd = MK && A.dd;
  const double* gl = A.in + (long)blockIdx.y * A.il2 + (long)(L0 + lx) * A.il1;
  double* go = A.out + (long)blockIdx.y * A.ol2 + (long)(L0 + lx) * A.ol1;
  const long ies = XD ? 1 : A.ies, oes = XD ? 1 : A.oes;
#define SI(p_, l_) (XD ? (p_) * NLB + ((l_) ^ (((p_) ^ ((p_) >> 3) ^ ((p_) >> 4)) & (NLB - 1))) : (p_) * NLB + (l_))
  double2 x[E];
  int pos[E];

  double2* swn = sw + M;                      // split twiddles wn[0..M/2] of the pre-/post-stage
  double2* sh4 = swn + (M / 2 + 1);           // Makhoul twiddles h4[0..M] (MK only)
  for (int q = tid; q < M; q += NT) sw[q] = __ldg(A.wm + q);
  for (int q = tid; q <= M / 2; q += NT) swn[q] = __ldg(A.wn + q);
  if (MK) for (int q = tid; q <= M; q += NT) sh4[q] = __ldg(A.h4 + q);
  double** sbase = (double**)(sh4 + (M + 1)); // PEER: per-rank row bases for this CTA's (x-tile, z)
  if (PEER && tid < A.np)
    sbase[tid] = A.pbase[tid] + L0 + (long)A.nx * A.pny[tid] * (A.zoff + (int)blockIdx.y) - (long)A.nx * A.pys[tid];

  if (!INV) {
    // ---- forward: first-stage operands straight from global memory
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = tx + T * e;
      int e0, e1;
      if (!MK) { e0 = 2 * c; e1 = 2 * c + 1; }
      else if (e < E / 2) { e0 = 4 * c; e1 = 4 * c + 2; }               // c < M/2  <=>  e < E/2
      else { e0 = 2 * n - 1 - 4 * c; e1 = 2 * n - 3 - 4 * c; }
      double a = 0., b = 0.;
      if (DIVSRC) {
        if (on) {
          const long off = (long)blockIdx.y * A.il2 + (long)(L0 + lx) * A.il1;
          const int jj = L0 + lx + 1, kk = (int)blockIdx.y + 1;
          const double dzk = __ldg(A.sdzfi + kk);
          a = fftb_div_src(A, off, e0, jj, kk, dzk); b = fftb_div_src(A, off, e1, jj, kk, dzk);
        }
      } else if (on) { a = gl[(long)e0 * ies]; b = gl[(long)e1 * ies]; }
      if (MK && e >= E / 2 && dd) { a = -a; b = -b; }                  // odd line elements change sign for the DST
      x[e] = make_double2(a, b);
    }
  } else {
    // ---- backward pre-stage: half spectrum (global) -> packed complex Z (shared).  All 4 (E/2 + 1) operands of the thread are
    // loaded before the first one is used (register-blocked, like the forward pass): the loads overlap instead of
    // alternating with the combine arithmetic and the shared-memory stores (ncu r2o: long-scoreboard stall 9.9 per issue)
    double rk[E / 2 + 1], rnk[E / 2 + 1], rmk[E / 2 + 1], rnmk[E / 2 + 1];
#pragma unroll
    for (int b = 0; b <= E / 2; ++b) {
      const int k = tx + T * b, mk = M - k;
      rk[b] = rnk[b] = rmk[b] = rnmk[b] = 0.;
      if (b == E / 2 && tx > 0) continue;          // k = M/2 belongs to tx = 0
      const int snk = k > 0 ? n - k : 0, snmk = n - mk;
      if (on) {
#define GI(s_) gl[(long)(dd ? n - 1 - (s_) : (s_)) * ies]
        rk[b] = GI(k); rnk[b] = GI(snk); rmk[b] = GI(mk); rnmk[b] = GI(snmk);
#undef GI
      }
    }
    __syncthreads();                               // twiddle tables staged (their loads overlapped the operand loads above)
#pragma unroll
    for (int b = 0; b <= E / 2; ++b) {
      const int k = tx + T * b, mk = M - k;
      if (b == E / 2 && tx > 0) continue;
      double2 Xk, Xmk;
      if (!MK) {
        Xk = make_double2(rk[b], (k > 0 && k < M) ? rnk[b] : 0.);
        Xmk = make_double2(rmk[b], (mk > 0 && mk < M) ? rnmk[b] : 0.);
      } else {
        const double2 hk = conj(sh4[k]), hmk = conj(sh4[mk]);
        Xk = mul(make_double2(rk[b], k > 0 ? -rnk[b] : 0.), hk);
        Xmk = mul(make_double2(rmk[b], -rnmk[b]), hmk);
      }
      const double2 Aa = add(Xk, conj(Xmk));
      const double2 Bb = mul(sub(Xk, conj(Xmk)), conj(swn[k]));
      S[SI(k, lx)] = make_double2(Aa.x - Bb.y, Aa.y + Bb.x);
      if (k > 0 && mk != k) S[SI(mk, lx)] = make_double2(Aa.x + Bb.y, -Aa.y + Bb.x);
    }
    __syncthreads();
  }

  // ---- complex FFT of length M: Stockham stages on registers, in-place exchange through S.  Radix schedule: E, E, ...,
  // then the remaining power of two.  Forward: stage 0 runs in the global mapping; backward: the last stage does.
  constexpr int R0 = M >= E ? E : M;
  constexpr int M1 = M / R0, R1 = M1 >= E ? E : M1;
  constexpr int M2 = M1 / R1, R2 = M2 >= E ? E : M2;
  constexpr int M3 = M2 / R2;
  static_assert(M3 == 1 && M1 > 1, "two or three stages");
  constexpr int NST = M2 > 1 ? 3 : 2;
#define PUT(l_) { _Pragma("unroll") for (int e = 0; e < E; ++e) S[SI(pos[e], l_)] = x[e]; __syncthreads(); }
#define GET(l_, t_) { _Pragma("unroll") for (int e = 0; e < E; ++e) x[e] = S[SI((t_) + T * e, l_)]; }
  if (!INV) {
    stage<M, E, R0, 1, INV>(x, pos, tx, sw);
    PUT(lx)
    GET(ls, ts) __syncthreads();
    stage<M, E, R1, R0, INV>(x, pos, ts, sw);
    PUT(ls)
    if constexpr (NST == 3) {
      GET(ls, ts) __syncthreads();
      stage<M, E, R2, R0 * R1, INV>(x, pos, ts, sw);
      PUT(ls)
    }
  } else {
    GET(ls, ts) __syncthreads();
    stage<M, E, R0, 1, INV>(x, pos, ts, sw);
    PUT(ls)
    if constexpr (NST == 3) {
      GET(ls, ts) __syncthreads();
      stage<M, E, R1, R0, INV>(x, pos, ts, sw);
      PUT(ls)
      GET(lx, tx)
      stage<M, E, R2, R0 * R1, INV>(x, pos, tx, sw);
    } else {
      GET(lx, tx)
      stage<M, E, R1, R0, INV>(x, pos, tx, sw);
    }
  }
#undef PUT
#undef GET

  if (!INV) {
    // ---- forward post-stage: even/odd split (+ Makhoul twiddles), output ordering, straight to global memory
#pragma unroll
    for (int b = 0; b <= E / 2; ++b) {
      const int k = tx + T * b, mk = M - k;
      if (b == E / 2 && tx > 0) continue;
      const double2 Zk = S[SI(k, lx)], Zmk = conj(S[SI(k == 0 ? 0 : mk, lx)]);
      const double2 Ev = make_double2(0.5 * (Zk.x + Zmk.x), 0.5 * (Zk.y + Zmk.y));
      const double2 D = sub(Zk, Zmk);
      const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);
      const double2 Tw = mul(swn[k], O);
      const double2 Xk = add(Ev, Tw), Xmk = conj(sub(Ev, Tw));
      int i0, i1, i2, i3;            // output slots of the four reals (-1: none)
      double r0, r1, r2, r3;
      if (!MK) {
        i0 = k; r0 = Xk.x;
        i1 = (k > 0 && k < M) ? n - k : -1; r1 = Xk.y;
        i2 = mk; r2 = Xmk.x;
        i3 = (mk > 0 && mk < M) ? n - mk : -1; r3 = Xmk.y;
      } else {
        const double2 Yk = mul(sh4[k], Xk), Ymk = mul(sh4[mk], Xmk);
        r0 = 2. * Yk.x; r1 = -2. * Yk.y; r2 = 2. * Ymk.x; r3 = -2. * Ymk.y;
        if (!dd) { i0 = k; i1 = k > 0 ? n - k : -1; i2 = mk; i3 = n - mk; }
        else { i0 = n - 1 - k; i1 = k > 0 ? k - 1 : -1; i2 = n - 1 - mk; i3 = mk - 1; }
      }
      if (on) {
        if (PEER) {
#define FB_PUT(iq_, rq_)                                                          \
  {                                                                               \
    int rr = 0;                                                                   \
    for (int q = 1; q < A.np; ++q) rr += (iq_) >= A.pys[q];                        \
    sbase[rr][lx + (long)A.nx * (iq_)] = (rq_) * A.scale;                          \
  }
          FB_PUT(i0, r0) if (i1 >= 0) FB_PUT(i1, r1)
          FB_PUT(i2, r2) if (i3 >= 0) FB_PUT(i3, r3)
#undef FB_PUT
        } else {
          go[(long)i0 * oes] = r0 * A.scale; if (i1 >= 0) go[(long)i1 * oes] = r1 * A.scale;
          go[(long)i2 * oes] = r2 * A.scale; if (i3 >= 0) go[(long)i3 * oes] = r3 * A.scale;
        }
      }
    }
  } else if (on) {
    // ---- backward store: z_c = (v_2c, v_2c+1), x = inverse Makhoul permutation of v
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = pos[e];
      const int e0 = src_of<MK>(2 * c, n), e1 = src_of<MK>(2 * c + 1, n);
      double a = x[e].x, b = x[e].y;
      if (dd) { if (e0 & 1) a = -a; if (e1 & 1) b = -b; }
      go[(long)e0 * oes] = a * A.scale;
      go[(long)e1 * oes] = b * A.scale;
    }
  }
#undef SI
}

template <int M, int E, int XD>
static inline int fftb_launch(cales_ctx* ctx, const FftBArgs& A, int kind, int backward) {
  const size_t sh = ((size_t)M * NLB + M + (M / 2 + 1) + (M + 1)) * sizeof(double2) + 8 * sizeof(double*);
  dim3 g(cdiv(A.nl1, NLB), A.nl2), b(NLB, M / E);
#define FB_GO(INV_, MK_)                                                                                           \
  {                                                                                                                \
    static bool attr = false;                                                                                      \
    if (!attr) { attr = true; cudaFuncSetAttribute(fftb_k<M, E, XD, INV_, MK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); } \
    fftb_k<M, E, XD, INV_, MK_><<<g, b, sh, ctx->stream>>>(A);                                                      \
  }
  const int mk = kind != KB_PP;
  if (XD == 0 && A.np > 0 && !backward) {
#define FB_GOP(MK_)                                                                                                \
  {                                                                                                                \
    static bool attr = false;                                                                                      \
    if (!attr) { attr = true; cudaFuncSetAttribute(fftb_k<M, E, 0, false, MK_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); } \
    fftb_k<M, E, 0, false, MK_, true><<<g, b, sh, ctx->stream>>>(A);                                                \
  }
    if (mk) FB_GOP(1) else FB_GOP(0)
#undef FB_GOP
  } else if (XD == 1 && A.su && !backward) {
    if constexpr (XD == 1) {
#define FB_GOD(MK_)                                                                                                \
  {                                                                                                                \
    static bool attr = false;                                                                                      \
    if (!attr) { attr = true; cudaFuncSetAttribute(fftb_k<M, E, 1, false, MK_, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); } \
    fftb_k<M, E, 1, false, MK_, false, true><<<g, b, sh, ctx->stream>>>(A);                                         \
  }
      if (mk) FB_GOD(1) else FB_GOD(0)
#undef FB_GOD
    }
  } else if (!backward) { if (mk) FB_GO(false, 1) else FB_GO(false, 0) }
  else { if (mk) FB_GO(true, 1) else FB_GO(true, 0) }
#undef FB_GO
  ctx->launches++;
  if (cudaGetLastError() != cudaSuccess) return -cales_fail(ctx, CALES_ERR_CUDA, "fftb_k launch failed");
  return 1;
}

// points per thread for n = 128 / 256: 8 (three / three stages) or 16 (two stages: one shared-memory exchange less); CALES_FFT_E16
static inline bool fftb_e16() { static const int v = getenv("CALES_FFT_E16") ? atoi(getenv("CALES_FFT_E16")) : 0; return v != 0; }

template <int XD>
static inline int fftb_dispatch(cales_ctx* ctx, int n, const FftBArgs& A, int kind, int backward) {
  switch (n) {
    case 32: return fftb_launch<16, 8, XD>(ctx, A, kind, backward);
    case 64: return fftb_launch<32, 8, XD>(ctx, A, kind, backward);
    case 128: return fftb_e16() ? fftb_launch<64, 16, XD>(ctx, A, kind, backward) : fftb_launch<64, 8, XD>(ctx, A, kind, backward);
    case 256: return fftb_e16() ? fftb_launch<128, 16, XD>(ctx, A, kind, backward) : fftb_launch<128, 8, XD>(ctx, A, kind, backward);
    case 512: return fftb_launch<256, 16, XD>(ctx, A, kind, backward);
    case 1024: return fftb_launch<512, 16, XD>(ctx, A, kind, backward);
    default: return 0;
  }
}
