// Batched 1-D real transforms of the Poisson/Helmholtz solver, hand-written (no cuFFT).
// Replaces the FFTW r2r plans of the reference CPU path (src/fft.f90:23-143,176-190) and the cuFFT
// + `signal_processing` re-ordering passes of its GPU path (src/fft.f90:247-661):
//   PP  R2HC / HC2R             (halfcomplex order r0..r_{n/2}, i_{(n+1)/2-1}..i_1, as FFTW)
//   NN  REDFT10 / REDFT01       (DCT-II / DCT-III, Makhoul's N-point algorithm)
//   DD  RODFT10 / RODFT01       (DST-II / DST-III via sign flip + index reversal of the DCT)
//   NN-f REDFT00, DD-f RODFT00  (DCT-I / DST-I: packed real FFT of the even / odd extension, length 2(n-1) / 2n)
//   ND-c REDFT11, DN-c RODFT11  (DCT-IV / DST-IV: n/2-point complex FFT with pre- and post-twiddles)
//   ND-f REDFT10 / REDFT01, DN-f RODFT01 / RODFT10 (the DCT/DST-II/III kernels, directions swapped for DN)
// -- the ten BC/stagger combinations of find_fft (src/fft.f90:192-245; the reference's own GPU path covers only the
// first three, src/fft.f90:527-569) -- all unnormalised with FFTW's factor-2 convention.
// One CTA stages NL lines in shared memory, packs each real line of n points into n/2 complex
// points, runs a Stockham autosort FFT (radix 4/2/3/5/generic) between two shared buffers and
// applies the even/odd split, the DCT twiddles, the index permutations and the final scale while
// loading/storing, so a transform pass reads and writes every element exactly once (16 B/cell).
// x-lines are contiguous; y-lines are handled as [NL consecutive x] x [all y] tiles so global access
// stays coalesced without any transpose pass.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

enum { K_PP = 0, K_NN = 1, K_DD = 2, K_C1 = 3, K_S1 = 4, K_C4 = 5, K_S4 = 6 };

struct FftArgs {
  const double* in; double* out;
  long ies, il1, il2;      // input strides: element, line index 1, line index 2
  long oes, ol1, ol2;      // output strides
  int n, nl1, nl2;         // transform length, number of lines along l1 / l2
  int kind, backward;
  double scale;
  const double2* wm;       // exp(-2 pi i t/m), t<m   (m = n/2)
  const double2* wn;       // exp(-2 pi i t/n), t<n
  const double2* h4;       // exp(-pi i t/(2n)), t<=n
  int nfac; int fac[24];   // radices of m
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i (forward) or +i (inverse)
__device__ __forceinline__ double2 cmuli(double2 a, bool inv) { return inv ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }
__device__ __forceinline__ double2 twid(const double2* __restrict__ w, int t, bool inv) {
  double2 v = __ldg(w + t);
  if (inv) v.y = -v.y;
  return v;
}

// Thread mapping: a CTA of FT threads owns NL lines; TPL = FT/NL consecutive threads work on one line, so no
// integer division appears in any loop (NL, TPL are compile-time powers of two).
#define FT 128

// one Stockham stage of radix R for my line: src -> dst (complex, m points); lane = my index within the line team
template <int R, int TPL>
__device__ __forceinline__ void stage(const double2* __restrict__ s, double2* __restrict__ dbase, int lane, int m, int Ns,
                                      const double2* __restrict__ wm, bool inv) {
  const int nb = m / R;
  const int tstep = m / (Ns * R);
  const bool p2 = (Ns & (Ns - 1)) == 0;
  for (int j = lane; j < nb; j += TPL) {
    const int k = p2 ? (j & (Ns - 1)) : (j % Ns);
    double2* d = dbase + (j - k) * R + k;
    if (R == 2) {
      double2 a = s[j], b = s[j + nb];
      if (Ns > 1) b = cmul(b, twid(wm, k * tstep, inv));
      d[0] = cadd(a, b); d[Ns] = csub(a, b);
    } else if (R == 4) {
      double2 a = s[j], b = s[j + nb], c = s[j + 2 * nb], e = s[j + 3 * nb];
      if (Ns > 1) {
        b = cmul(b, twid(wm, k * tstep, inv));
        c = cmul(c, twid(wm, 2 * k * tstep, inv));
        e = cmul(e, twid(wm, 3 * k * tstep, inv));
      }
      const double2 t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, e), t3 = cmuli(csub(b, e), inv);
      d[0] = cadd(t0, t2); d[Ns] = cadd(t1, t3); d[2 * Ns] = csub(t0, t2); d[3 * Ns] = csub(t1, t3);
    } else {
      // generic radix R (3, 5, any prime; R == 0 means "runtime radix in Rrt"): O(R^2) DFT with table twiddles
      const int rstep = m / R;
      for (int q = 0; q < R; ++q) {
        double2 acc = s[j];
        for (int a = 1; a < R; ++a) {
          double2 v = s[j + a * nb];
          if (Ns > 1) v = cmul(v, twid(wm, a * k * tstep, inv));
          acc = cadd(acc, cmul(v, twid(wm, ((a * q) % R) * rstep, inv)));
        }
        d[q * Ns] = acc;
      }
    }
  }
}

template <int TPL>
__device__ __forceinline__ void stage_any(int R, const double2* __restrict__ s, double2* __restrict__ dbase, int lane, int m, int Ns,
                                          const double2* __restrict__ wm, bool inv) {
  const int nb = m / R, tstep = m / (Ns * R), rstep = m / R;
  for (int j = lane; j < nb; j += TPL) {
    const int k = j % Ns;
    double2* d = dbase + (j - k) * R + k;
    for (int q = 0; q < R; ++q) {
      double2 acc = s[j];
      for (int a = 1; a < R; ++a) {
        double2 v = s[j + a * nb];
        if (Ns > 1) v = cmul(v, twid(wm, (int)(((long)a * k * tstep) % m), inv));
        acc = cadd(acc, cmul(v, twid(wm, ((a * q) % R) * rstep, inv)));
      }
      d[q * Ns] = acc;
    }
  }
}

// position in the packed complex array (as a double index) of real sample e of a forward input line
__device__ __forceinline__ int fwd_slot(int e, int n, int kind) {
  if (kind == K_PP) return e;
  return (e & 1) ? n - 1 - (e >> 1) : (e >> 1);            // Makhoul: v_j = x_2j, v_{n-1-j} = x_{2j+1}
}

template <int DIR, int NL>
__global__ void __launch_bounds__(FT) fft_lines_k(FftArgs A, int LS) {
  constexpr int TPL = FT / NL;
  extern __shared__ double2 sm[];
  double2* b0 = sm;
  double2* b1 = sm + (size_t)NL * LS;
  const int n = A.n, m = n >> 1;
  const int l1_0 = blockIdx.x * NL, l2 = blockIdx.y;
  const int nl = min(NL, A.nl1 - l1_0);
  const bool inv = A.backward != 0;
  const int kind = A.kind;
  const double* gin = A.in + (long)l2 * A.il2 + (long)l1_0 * A.il1;
  double* gout = A.out + (long)l2 * A.ol2 + (long)l1_0 * A.ol1;
  const int LSD = 2 * LS;
  // global <-> shared mapping: x-lines: TPL consecutive threads sweep one line; y-lines: NL consecutive threads
  // take the same element of NL neighbouring lines (contiguous in memory)
  const int gl = DIR == 0 ? threadIdx.x / TPL : threadIdx.x % NL;       // line of my global accesses
  const int ge0 = DIR == 0 ? threadIdx.x % TPL : threadIdx.x / NL;      // first element
  const int gstep = TPL;
  // compute mapping: my line and my lane within its team
  const int cl = threadIdx.x / TPL, lane = threadIdx.x % TPL;
  // ---- load ---------------------------------------------------------------------------------------------
  if (gl < nl) {
    double* dst = (double*)(inv ? b1 : b0) + gl * LSD;  // forward: packed complex in b0; backward: raw reals in b1
    const double* g = gin + (long)gl * A.il1;
    for (int e = ge0; e < n; e += gstep) {
      double v = g[(long)e * A.ies];
      int slot;
      if (!inv) {
        slot = fwd_slot(e, n, kind);
        if (kind == K_DD && (e & 1)) v = -v;           // DST-II(x)_k = DCT-II((-1)^j x_j)_{n-1-k}
      } else {
        slot = kind == K_DD ? n - 1 - e : e;           // DST-III(a)_k = (-1)^k DCT-III(reversed a)_k
      }
      dst[slot] = v;
    }
  }
  __syncthreads();
  const bool on = cl < nl;
  // ---- backward pre-stage: half spectrum -> packed complex Z (b1 reals -> b0 complex) --------------------------------
  if (inv) {
    if (on) {
      const int np = m / 2 + 1;
      const double* R = (const double*)(b1 + (size_t)cl * LS);
      double2* Z = b0 + (size_t)cl * LS;
      for (int k = lane; k < np; k += TPL) {
        const int mk = m - k;
        double2 Xk, Xmk;
        if (kind == K_PP) {
          Xk = make_double2(R[k], (k > 0 && k < m) ? R[n - k] : 0.);
          Xmk = make_double2(R[mk], (mk > 0 && mk < m) ? R[n - mk] : 0.);
        } else {                                       // V_k = (a_k - i a_{n-k}) e^{+i pi k/(2n)}, a_n := 0
          const double2 hk = cconj(__ldg(A.h4 + k)), hmk = cconj(__ldg(A.h4 + mk));
          Xk = cmul(make_double2(R[k], k > 0 ? -R[n - k] : 0.), hk);
          Xmk = cmul(make_double2(R[mk], -R[n - mk]), hmk);
        }
        const double2 Aa = cadd(Xk, cconj(Xmk));
        const double2 Bb = cmul(csub(Xk, cconj(Xmk)), cconj(__ldg(A.wn + k)));
        Z[k] = make_double2(Aa.x - Bb.y, Aa.y + Bb.x);
        if (k > 0 && mk != k) Z[mk] = make_double2(Aa.x + Bb.y, -Aa.y + Bb.x);
      }
    }
    __syncthreads();
  }
  // ---- complex FFT of length m (Stockham autosort, ping-pong b0 <-> b1) ------------------------------------------------
  double2* src = b0;
  double2* dst = b1;
  int Ns = 1;
  for (int st = 0; st < A.nfac; ++st) {
    const int R = A.fac[st];
    if (on) {
      const double2* sl = src + (size_t)cl * LS;
      double2* dl = dst + (size_t)cl * LS;
      if (R == 4) stage<4, TPL>(sl, dl, lane, m, Ns, A.wm, inv);
      else if (R == 2) stage<2, TPL>(sl, dl, lane, m, Ns, A.wm, inv);
      else if (R == 3) stage<3, TPL>(sl, dl, lane, m, Ns, A.wm, inv);
      else if (R == 5) stage<5, TPL>(sl, dl, lane, m, Ns, A.wm, inv);
      else stage_any<TPL>(R, sl, dl, lane, m, Ns, A.wm, inv);
    }
    Ns *= R;
    __syncthreads();
    double2* t = src; src = dst; dst = t;
  }
  // result is in `src`; `dst` is free
  if (!inv) {
    // ---- forward post-stage: Z -> X (even/odd split) -> output ordering, written as reals into dst ------------------------
    if (on) {
      const int np = m / 2 + 1;
      const double2* Z = src + (size_t)cl * LS;
      double* R = (double*)(dst + (size_t)cl * LS);
      for (int k = lane; k < np; k += TPL) {
        const int mk = m - k;
        const double2 Zk = Z[k], Zmk = cconj(Z[k == 0 ? 0 : mk]);
        const double2 E = make_double2(0.5 * (Zk.x + Zmk.x), 0.5 * (Zk.y + Zmk.y));
        const double2 D = csub(Zk, Zmk);
        const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);          // (Zk - conj Zmk)/(2i)
        const double2 T = cmul(__ldg(A.wn + k), O);
        const double2 Xk = cadd(E, T), Xmk = cconj(csub(E, T));        // X[k], X[m-k]
        if (kind == K_PP) {
          R[k] = Xk.x;
          if (k > 0 && k < m) R[n - k] = Xk.y;
          R[mk] = Xmk.x;
          if (mk > 0 && mk < m) R[n - mk] = Xmk.y;
        } else {
          const double2 Yk = cmul(__ldg(A.h4 + k), Xk), Ymk = cmul(__ldg(A.h4 + mk), Xmk);
          if (kind == K_NN) {
            R[k] = 2. * Yk.x;
            if (k > 0) R[n - k] = -2. * Yk.y;
            R[mk] = 2. * Ymk.x;
            if (mk < n && mk > 0) R[n - mk] = -2. * Ymk.y;
          } else {                                                     // reversed order for the DST
            R[n - 1 - k] = 2. * Yk.x;
            if (k > 0) R[k - 1] = -2. * Yk.y;
            R[n - 1 - mk] = 2. * Ymk.x;
            if (mk > 0) R[mk - 1] = -2. * Ymk.y;
          }
        }
      }
    }
    __syncthreads();
    if (gl < nl) {
      const double* Rb = (const double*)dst + gl * LSD;
      double* g = gout + (long)gl * A.ol1;
      for (int e = ge0; e < n; e += gstep) g[(long)e * A.oes] = Rb[e] * A.scale;
    }
  } else {
    // ---- backward store: z_j -> x_2j, x_2j+1 with the inverse Makhoul permutation ------------------------------------------
    if (gl < nl) {
      const double* Zb = (const double*)src + gl * LSD;
      double* g = gout + (long)gl * A.ol1;
      for (int e = ge0; e < n; e += gstep) {
        const int q = fwd_slot(e, n, kind);
        double v = Zb[q];
        if (kind == K_DD && (e & 1)) v = -v;
        g[(long)e * A.oes] = v * A.scale;
      }
    }
  }
}


// ---- the remaining FFTW r2r kinds (generic path only; none of them is on a BASELINE configuration's pressure solve) ------
//   K_C1  REDFT00  Y_k = X_0 + (-1)^k X_{N-1} + 2 sum_{0<j<N-1} X_j cos(pi j k/(N-1)): real DFT of the even extension
//         (length L = 2(N-1)), computed as a packed complex FFT of m = N-1 points; Y_k = Re X~_k, k = 0..m
//   K_S1  RODFT00  Y_k = 2 sum_j X_j sin(pi (j+1)(k+1)/(N+1)): odd extension of length L = 2(N+1), m = N+1;
//         Y_{k-1} = -Im X~_k, k = 1..N
//   K_C4  REDFT11  Y_k = 2 sum_j X_j cos(pi (j+1/2)(k+1/2)/N), N even: c_j = (X_2j + i X_{N-1-2j}) e^{-i pi j/N},
//         A = FFT_{N/2}(c) . e^{-i pi (4k+1)/(4N)};  Y_2k = 2 Re A_k,  Y_{N-1-2k} = -2 Im A_k
//   K_S4  RODFT11  = (-1)^k DCT-IV(reversed X)_k
// All four are their own inverses up to the normalisation that normfft carries.  A line may be longer than the
// transform (face-centred DD: n-1 points, fft.f90:66-69); the extra points pass through (scaled like the rest).
struct R2RArgs {
  const double* in; double* out;
  long ies, il1, il2, oes, ol1, ol2;
  int nline, nt, m, nl1, nl2, kind;   // points per line, transform length, complex FFT length
  double scale;
  const double2 *wm, *wl, *h4, *g4;   // exp(-2 pi i t/m); exp(-2 pi i t/(2m)); exp(-pi i t/(2N)); exp(-pi i (4t+1)/(4N))
  int nfac; int fac[24];
};

template <int DIR, int NL>
__global__ void __launch_bounds__(FT) r2r_lines_k(R2RArgs A, int LS) {
  constexpr int TPL = FT / NL;
  extern __shared__ double2 sm[];
  double2* b0 = sm;
  double2* b1 = sm + (size_t)NL * LS;
  const int nt = A.nt, m = A.m, kind = A.kind;
  const int l1_0 = blockIdx.x * NL, l2 = blockIdx.y;
  const int nl = min(NL, A.nl1 - l1_0);
  const double* gin = A.in + (long)l2 * A.il2 + (long)l1_0 * A.il1;
  double* gout = A.out + (long)l2 * A.ol2 + (long)l1_0 * A.ol1;
  const int LSD = 2 * LS;
  const int gl = DIR == 0 ? threadIdx.x / TPL : threadIdx.x % NL;
  const int ge0 = DIR == 0 ? threadIdx.x % TPL : threadIdx.x / NL;
  const int gstep = TPL;
  const int cl = threadIdx.x / TPL, lane = threadIdx.x % TPL;
  // ---- load: build the packed complex sequence in b0
  if (gl < nl) {
    double* dst = (double*)b0 + gl * LSD;
    const double* g = gin + (long)gl * A.il1;
    if (kind == K_S1 && ge0 == 0) { dst[0] = 0.; dst[m] = 0.; }
    for (int e = ge0; e < nt; e += gstep) {
      const double v = g[(long)e * A.ies];
      if (kind == K_C1) { dst[e] = v; if (e > 0 && e < m) dst[2 * m - e] = v; }
      else if (kind == K_S1) { dst[e + 1] = v; dst[2 * m - 1 - e] = -v; }
      else {
        const int r = kind == K_S4 ? nt - 1 - e : e;
        dst[(r & 1) ? nt - r : r] = v;
      }
    }
  }
  __syncthreads();
  const bool on = cl < nl;
  if (kind >= K_C4) {
    if (on) {
      double2* Z = b0 + (size_t)cl * LS;
      for (int j = lane; j < m; j += TPL) Z[j] = cmul(Z[j], __ldg(A.h4 + 2 * j));
    }
    __syncthreads();
  }
  double2* src = b0;
  double2* dst = b1;
  int Ns = 1;
  for (int st = 0; st < A.nfac; ++st) {
    const int R = A.fac[st];
    if (on) {
      const double2* sl = src + (size_t)cl * LS;
      double2* dl = dst + (size_t)cl * LS;
      if (R == 4) stage<4, TPL>(sl, dl, lane, m, Ns, A.wm, false);
      else if (R == 2) stage<2, TPL>(sl, dl, lane, m, Ns, A.wm, false);
      else if (R == 3) stage<3, TPL>(sl, dl, lane, m, Ns, A.wm, false);
      else if (R == 5) stage<5, TPL>(sl, dl, lane, m, Ns, A.wm, false);
      else stage_any<TPL>(R, sl, dl, lane, m, Ns, A.wm, false);
    }
    Ns *= R;
    __syncthreads();
    double2* t = src; src = dst; dst = t;
  }
  // ---- post-stage: results as reals into dst
  if (on) {
    const double2* Z = src + (size_t)cl * LS;
    double* R = (double*)(dst + (size_t)cl * LS);
    if (kind <= K_S1) {
      const int np = m / 2 + 1;
      for (int k = lane; k < np; k += TPL) {
        const int mk = m - k;
        const double2 Zk = Z[k == m ? 0 : k], Zmk = cconj(Z[k == 0 ? 0 : mk]);
        const double2 E = make_double2(0.5 * (Zk.x + Zmk.x), 0.5 * (Zk.y + Zmk.y));
        const double2 D = csub(Zk, Zmk);
        const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);
        const double2 T = cmul(__ldg(A.wl + k), O);
        const double2 Xk = cadd(E, T), Xmk = cconj(csub(E, T));        // X~[k], X~[m-k]
        if (kind == K_C1) { R[k] = Xk.x; R[mk] = Xmk.x; }
        else { if (k > 0) { R[k - 1] = -Xk.y; R[mk - 1] = -Xmk.y; } }
      }
    } else {
      for (int k = lane; k < m; k += TPL) {
        const double2 a = cmul(Z[k], __ldg(A.g4 + k));
        R[2 * k] = 2. * a.x;
        R[nt - 1 - 2 * k] = kind == K_C4 ? -2. * a.y : 2. * a.y;
      }
    }
  }
  __syncthreads();
  if (gl < nl) {
    const double* Rb = (const double*)dst + gl * LSD;
    const double* gi = gin + (long)gl * A.il1;
    double* g = gout + (long)gl * A.ol1;
    for (int e = ge0; e < A.nline; e += gstep) g[(long)e * A.oes] = (e < nt ? Rb[e] : gi[(long)e * A.ies]) * A.scale;
  }
}

int k_fftb_pass(cales_ctx* ctx, int dir, int kind, int backward, int n, int nl1, int nl2, const double* in, long ies, long il1, long il2,
                double* out, long oes, long ol1, long ol2, double scale, const FftTables* T, const DivSrc* ds);
bool k_fftb_supported(int n);
int k_fft_pass(cales_ctx* ctx, int dir, const char bc[2], char c_or_f, int backward, int n1, int n2, int n3,
               const double* in, long ip1, long ip2, double* out, long op1, long op2, double scale, const DivSrc* ds = nullptr);

// ---- host side -------------------------------------------------------------------------------------------------------
FftTables* k_tables(cales_ctx* ctx, int n) {
  auto it = ctx->tables.find(n);
  if (it != ctx->tables.end()) return &it->second;
  // w: [0,m) exp(-2 pi i t/m) ; [m, m+n) exp(-2 pi i t/n) ; h: [0,n] exp(-pi i t/(2n)); long double on the host
  const int m = n / 2;
  std::vector<double2> w(m + n), h(n + 1 + m);   // h[n+1 ..): exp(-pi i (4t+1)/(4n)), t < m (DCT/DST-IV post-twiddles)
  const long double pi = acosl(-1.0L);
  for (int t = 0; t < m; ++t) w[t] = make_double2((double)cosl(2 * pi * t / m), (double)-sinl(2 * pi * t / m));
  for (int t = 0; t < n; ++t) w[m + t] = make_double2((double)cosl(2 * pi * t / n), (double)-sinl(2 * pi * t / n));
  for (int t = 0; t <= n; ++t) h[t] = make_double2((double)cosl(pi * t / (2 * n)), (double)-sinl(pi * t / (2 * n)));
  for (int t = 0; t < m; ++t) h[n + 1 + t] = make_double2((double)cosl(pi * (4 * t + 1) / (4 * n)), (double)-sinl(pi * (4 * t + 1) / (4 * n)));
  FftTables tb_;
  if (cudaMalloc(&tb_.w, w.size() * sizeof(double2)) != cudaSuccess || cudaMalloc(&tb_.h, h.size() * sizeof(double2)) != cudaSuccess) {
    cales_fail(ctx, CALES_ERR_NOMEM, "twiddle table allocation failed");
    return nullptr;
  }
  cudaMemcpyAsync(tb_.w, w.data(), w.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(tb_.h, h.data(), h.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);   // host vectors go out of scope
  ctx->tables[n] = tb_;
  return &ctx->tables[n];
}

// find_fft (src/fft.f90:192-245): kernel family for a BC pair and stagger; *swap = forward/backward kernels exchanged
static int kind_of(const char bc[2], char c_or_f, bool* swap) {
  *swap = false;
  const bool c = c_or_f == 'c', f = c_or_f == 'f';
  if (bc[0] == 'P' && bc[1] == 'P') return K_PP;
  if (bc[0] == 'N' && bc[1] == 'N') return c ? K_NN : f ? K_C1 : -1;       // REDFT10/01 | REDFT00
  if (bc[0] == 'D' && bc[1] == 'D') return c ? K_DD : f ? K_S1 : -1;       // RODFT10/01 | RODFT00
  if (bc[0] == 'N' && bc[1] == 'D') return c ? K_C4 : f ? K_NN : -1;       // REDFT11    | REDFT10/01
  if (bc[0] == 'D' && bc[1] == 'N') { *swap = f; return c ? K_S4 : f ? K_DD : -1; }   // RODFT11 | RODFT01/10
  return -1;
}

static int r2r_pass(cales_ctx* ctx, int dir, int kind, int n, int n1, int n2, int n3, const double* in, long ip1, long ip2, double* out,
                    long op1, long op2, double scale);

// dir 0: lines along x of an array with row pitch ps1 and plane pitch ps2 (elements); dir 1: along y.
int k_fft_pass(cales_ctx* ctx, int dir, const char bc[2], char c_or_f, int backward, int n1, int n2, int n3,
               const double* in, long ip1, long ip2, double* out, long op1, long op2, double scale, const DivSrc* ds) {
  bool swap;
  const int kind = kind_of(bc, c_or_f, &swap);
  if (kind < 0)
    return cales_fail(ctx, CALES_ERR_INVALID, "no transform for BC '%c%c' (%c-centred): find_fft knows P/P, N/N, D/D, N/D, D/N", bc[0], bc[1], c_or_f);
  if (swap) backward = !backward;
  const int n = dir == 0 ? n1 : n2;
  if (ds && (swap || kind >= K_C1 || !k_fftb_supported(n) || getenv("CALES_FFT_GENERIC")))
    return cales_fail(ctx, CALES_ERR_INVALID, "fused fillps needs the register-blocked transform (caller must check k_fft_peer_capable)");
  if (kind >= K_C1) return r2r_pass(ctx, dir, kind, n, n1, n2, n3, in, ip1, ip2, out, op1, op2, scale);
  if (n < 2 || (n & 1)) return cales_fail(ctx, CALES_ERR_INVALID, "transform length %d: only even lengths are implemented", n);
  FftTables* T = k_tables(ctx, n);
  if (!T) return CALES_ERR_NOMEM;
  {
    // fast path: register-blocked batched kernel (fftb.cu) for power-of-two lengths
    static const bool old_only = getenv("CALES_FFT_GENERIC") != nullptr;
    if (!old_only) {
      int rc;
      if (dir == 0) rc = k_fftb_pass(ctx, 0, kind, backward, n, n2, n3, in, 1, ip1, ip2, out, 1, op1, op2, scale, T, ds);
      else rc = k_fftb_pass(ctx, 1, kind, backward, n, n1, n3, in, ip1, 1, ip2, out, op1, 1, op2, scale, T, nullptr);
      if (rc < 0) return -rc;
      if (rc == 1) return CALES_OK;
    }
  }
  FftArgs A;
  A.in = in; A.out = out; A.n = n; A.kind = kind; A.backward = backward; A.scale = scale;
  const int m = n / 2;
  A.wm = T->w; A.wn = T->w + m; A.h4 = T->h;
  if (dir == 0) { A.ies = 1; A.il1 = ip1; A.il2 = ip2; A.oes = 1; A.ol1 = op1; A.ol2 = op2; A.nl1 = n2; A.nl2 = n3; }
  else { A.ies = ip1; A.il1 = 1; A.il2 = ip2; A.oes = op1; A.ol1 = 1; A.ol2 = op2; A.nl1 = n1; A.nl2 = n3; }
  // factorise m: radix 4 first, then 2, 3, 5, then any remaining primes
  int r = m; A.nfac = 0;
  while (r % 4 == 0) { A.fac[A.nfac++] = 4; r /= 4; }
  while (r % 2 == 0) { A.fac[A.nfac++] = 2; r /= 2; }
  for (int p = 3; r > 1; p += 2)
    while (r % p == 0) { A.fac[A.nfac++] = p; r /= p; if (A.nfac >= 23) break; }
  if (m == 1) A.nfac = 0;
  const int LS = m + 1;
  const size_t per_line = 2 * (size_t)LS * sizeof(double2);
  int NL = 8;
  while (NL > 2 && NL * per_line > 72 * 1024) NL >>= 1;
  if (NL * per_line > 200 * 1024) return cales_fail(ctx, CALES_ERR_INVALID, "transform length %d exceeds the shared-memory line buffer", n);
  const size_t sh = NL * per_line;
  static bool attr_set = false;
  if (!attr_set) {
    attr_set = true;
    cudaFuncSetAttribute(fft_lines_k<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fft_lines_k<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fft_lines_k<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fft_lines_k<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fft_lines_k<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(fft_lines_k<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 g(cdiv(A.nl1, NL), A.nl2);
#define GO(D_, N_) fft_lines_k<D_, N_><<<g, FT, sh, ctx->stream>>>(A, LS)
  if (dir == 0) { if (NL == 8) GO(0, 8); else if (NL == 4) GO(0, 4); else GO(0, 2); }
  else { if (NL == 8) GO(1, 8); else if (NL == 4) GO(1, 4); else GO(1, 2); }
#undef GO
  KERNEL_CHECK(ctx);
  return CALES_OK;
}


static int r2r_pass(cales_ctx* ctx, int dir, int kind, int n, int n1, int n2, int n3, const double* in, long ip1, long ip2, double* out,
                    long op1, long op2, double scale) {
  R2RArgs A;
  A.in = in; A.out = out; A.kind = kind; A.scale = scale; A.nline = n;
  if (kind == K_C1) { A.nt = n; A.m = n - 1; }
  else if (kind == K_S1) { A.nt = n - 1; A.m = n; }
  else { A.nt = n; A.m = n / 2; }
  if (A.nt < 1 || A.m < 1) return cales_fail(ctx, CALES_ERR_INVALID, "transform length %d is too short for this boundary condition", n);
  if (kind >= K_C4 && (n & 1)) return cales_fail(ctx, CALES_ERR_INVALID, "transform length %d: type-IV transforms need an even length", n);
  FftTables* T = k_tables(ctx, kind >= K_C4 ? n : 2 * A.m);
  if (!T) return CALES_ERR_NOMEM;
  A.wm = T->w; A.wl = T->w + A.m; A.h4 = T->h; A.g4 = T->h + n + 1;     // (g4 only meaningful for the type-IV tables)
  if (dir == 0) { A.ies = 1; A.il1 = ip1; A.il2 = ip2; A.oes = 1; A.ol1 = op1; A.ol2 = op2; A.nl1 = n2; A.nl2 = n3; }
  else { A.ies = ip1; A.il1 = 1; A.il2 = ip2; A.oes = op1; A.ol1 = 1; A.ol2 = op2; A.nl1 = n1; A.nl2 = n3; }
  int r = A.m; A.nfac = 0;
  while (r % 4 == 0) { A.fac[A.nfac++] = 4; r /= 4; }
  while (r % 2 == 0) { A.fac[A.nfac++] = 2; r /= 2; }
  for (int p = 3; r > 1; p += 2)
    while (r % p == 0) { A.fac[A.nfac++] = p; r /= p; if (A.nfac >= 23) break; }
  const int LS = A.m + 1;
  const size_t per_line = 2 * (size_t)LS * sizeof(double2);
  int NL = 8;
  while (NL > 2 && NL * per_line > 72 * 1024) NL >>= 1;
  if (NL * per_line > 200 * 1024) return cales_fail(ctx, CALES_ERR_INVALID, "transform length %d exceeds the shared-memory line buffer", n);
  const size_t sh = NL * per_line;
  static bool attr_set = false;
  if (!attr_set) {
    attr_set = true;
    cudaFuncSetAttribute(r2r_lines_k<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(r2r_lines_k<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(r2r_lines_k<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(r2r_lines_k<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(r2r_lines_k<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(r2r_lines_k<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 g(cdiv(A.nl1, NL), A.nl2);
#define GO(D_, N_) r2r_lines_k<D_, N_><<<g, FT, sh, ctx->stream>>>(A, LS)
  if (dir == 0) { if (NL == 8) GO(0, 8); else if (NL == 4) GO(0, 4); else GO(0, 2); }
  else { if (NL == 8) GO(1, 8); else if (NL == 4) GO(1, 4); else GO(1, 2); }
#undef GO
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// can the forward pass along a y-line of this BC/stagger/length run in the register-blocked kernel (fftb.cuh), which is
// the one that can scatter its spectrum straight into peer Z-pencils?
bool k_fftb_supported(int n);
bool k_fft_peer_capable(const char bc[2], char c_or_f, int n) {
  bool swap;
  const int kind = kind_of(bc, c_or_f, &swap);
  return kind >= 0 && kind <= K_DD && !swap && k_fftb_supported(n);
}

extern "C" int cales_fft_lines(cales_ctx* ctx, const int n[3], int dir, const char bc[2], char c_or_f, int backward, double* a) {
  CHECK_CTX(ctx);
  return k_fft_pass(ctx, dir, bc, c_or_f, backward, n[0], n[1], n[2], a, n[0], (long)n[0] * n[1], a, n[0], (long)n[0] * n[1], 1.0);
}
