// Shared internals of libcales_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/cales_b200.h"

// ---- indexing of haloed Fortran arrays (0:n1+1,0:n2+1,0:n3+1) -----------------------------------
struct Dims {
  int n1, n2, n3;  // interior extents
  long s1, s2;     // strides of j and k in elements: s1 = n1+2, s2 = (n1+2)*(n2+2)
  __host__ __device__ Dims() {}
  __host__ __device__ Dims(const int n[3]) : n1(n[0]), n2(n[1]), n3(n[2]), s1(n[0] + 2), s2((long)(n[0] + 2) * (n[1] + 2)) {}
  __host__ __device__ long idx(int i, int j, int k) const { return i + s1 * j + s2 * (long)k; }
  __host__ long size() const { return s2 * (n3 + 2); }
};

struct Plan {  // what fftini returns (src/fft.f90:23-143)
  bool used = false;
  int ng[3];
  char bc[2][2];     // [dir][ib]
  char c_or_f[2];
  double normfft;
  // what initsolver returned to the caller, kept on the host: eigenvalues of x and y (scaled by dli^2), tridiagonal
  // coefficients, z boundary types -- the distributed z solve (zdist.cu) builds its tables for the Y-pencil from them
  std::vector<double> lx, ly, a, b, c;
  char bcz[2] = {'P', 'P'};
};

struct FftTables {   // twiddle tables per transform length, device resident
  double2* w = nullptr;   // exp(-2 pi i k / n),      k = 0..n-1
  double2* h = nullptr;   // exp(-  pi i k / (2 n)),  k = 0..n-1   (Makhoul DCT twiddles)
};

// fillps + updt_rhs_b handed to the pressure solve instead of being run first (substep.cu -> solver.cu): haloed velocity arrays,
// dzfi(0:n3+1), the rhsb planes and is_bound; the forward x pass of the solver computes its input from them where it can
struct DivSrc { const double *u, *v, *w, *dzfi, *rbx, *rby, *rbz; double dti, dxi, dyi; int bnd[6]; };

// destinations of the kernel-fused transposes of the distributed solver (solver.cu sets them on the context around the
// producing kernel; fftb_y.cu / gaussel_tab.cu consume them)
struct FftPeerOut { int np, zoff, nx; double* pbase[8]; int pys[9], pny[8]; };   // forward y pass -> peers' Z-pencils
struct GPeer { int np; long plane, coff; double* pbase[8]; int pzs[9]; };        // z solve -> peers' Y-pencils: pbase[r] + (col + coff) + plane (l - pzs[r])

#define CALES_MAX_RANKS 64
struct PeerBuf {              // a buffer every rank allocated and mapped into every other rank (CUDA IPC over NVLink)
  void* local = nullptr;
  void* ptr[CALES_MAX_RANKS] = {nullptr};   // indexed by global rank; ptr[rank] == local
  size_t bytes = 0;
};

struct cales_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int diffusion = CALES_DIFF_EXPLICIT;
  // decomposition (initmpi)
  int ng[3], dims[2], ipencil, rank, nranks, coord[2];
  char cbcpre[6];
  int lo[3], hi[3], n[3];
  int xst[3], xen[3], xsz[3], yst[3], yen[3], ysz[3], zst[3], zen[3], zsz[3];
  int nb[6], is_bound[6];
  // comms (comm.cu)
  void* nccl = nullptr;                 // ncclComm_t
  cudaStream_t comm_stream = nullptr;
  std::map<std::string, PeerBuf> peerbufs;
  int p2p = -1;                         // -1 untested, 0 unavailable (NCCL send/recv transposes), 1 peer memory mapped
  double* bar = nullptr;                // barrier token (NCCL all-reduce barrier)
  unsigned long long bar_seq = 0;       // sequence number of the peer-memory flag barrier
  unsigned halo_seq = 0;                // parity of the double-buffered peer halo buffers
  unsigned long long halo_dir_seq[3] = {0, 0, 0};   // exchange count per direction (arrival flags of halo_fused_k)
  unsigned* halo_counter = nullptr;     // per direction: CTAs of halo_fused_k done pushing
  // scratch owned by the callee (the reference's `save`d allocatables and module buffers)
  std::map<std::string, std::pair<void*, size_t>> scratch;
  double* red = nullptr;                // device reduction slots
  double* red_host = nullptr;           // pinned mirror
  double* fdev = nullptr;               // bulk forcing f(3), device resident
  int rk_swap = 0;                      // which of the two RHS sets is "old" (rk.f90:98-100)
  bool rk_first = true;
  bool sgs_first = true;
  int sgs_ave = 1, sgs_filter2d = 0;    // dsmag averaging geometry / test filter (cales_set_sgs_options)
  cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // copy streams of the pipelined solver exchange (solver.cu)
  cudaEvent_t side_ev[8] = {nullptr};   // [0..3] chunk ready (main -> side), [4..7] chunk pushed (side -> main)
  const DivSrc* div_src = nullptr;      // set by cales_substep around cales_solver: right-hand side still to be formed (fused fillps)
  const FftPeerOut* fft_peer_out = nullptr;   // per-context (two contexts / host threads in one process do not interfere)
  const GPeer* gauss_peer_out = nullptr;
  void* gauss_state = nullptr;          // pivot tables of the tridiagonal solves (gaussel_tab.cu)
  void* zdist = nullptr;                // tables of the distributed z solve (zdist.cu)
  void* step_graphs = nullptr;          // captured time steps of cales_step (substep.cu)
  long step_calls = 0;                  // cales_step calls so far (the first ones run eagerly: lazy allocations)
  std::vector<Plan> plans;
  std::map<int, FftTables> tables;
  int solver_path = 0;                  // exchange of the last cales_solver call: 0 none (one rank), 1 distributed z solve, 2 copy-engine pipeline, 3 kernel-fused, 4 separate transposes
  long launches = 0;                    // kernels launched by this library (bench.py gpu_launches)
  char err[512] = {0};
};

extern char g_cales_err[512];

int cales_fail(cales_ctx* ctx, int code, const char* fmt, ...);
void* cales_scratch(cales_ctx* ctx, const char* name, size_t bytes, bool zero_on_create = false);

#define CUDA_TRY(ctx, call)                                                                        \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return cales_fail(ctx, CALES_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                             cudaGetErrorString(e_));                               \
  } while (0)

#define KERNEL_CHECK(ctx) do { (ctx)->launches++; CUDA_TRY(ctx, cudaGetLastError()); } while (0)
#define CHECK_CTX(ctx) do { if (!(ctx)) return CALES_ERR_INVALID; } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// z-chunk of a marching kernel: `cols` CTAs per chunk, `resident` CTAs fit on the GPU at once (148 SMs x CTAs/SM),
// every chunk re-reads `extra` planes (carried values / staging prologue).  Picks the chunk count that maximises
// (fill of the last wave) x (useful planes / planes read), with at least `min_kc` levels per chunk.
static inline int pick_chunk(long cols, int nk, int resident, int min_kc, int extra) {
  int best = nk; double beste = -1.;
  for (int chunks = 1; chunks <= nk; ++chunks) {
    const int kc = (nk + chunks - 1) / chunks;
    if (kc < min_kc && chunks > 1) break;
    const long ctas = cols * ((nk + kc - 1) / kc);
    const double waves = (double)ctas / resident;
    const double fill = waves / (double)((ctas + resident - 1) / resident);
    const double e = fill * kc / (double)(kc + extra) * (waves >= 3. ? 1. : 0.6 + 0.4 * waves / 3.);   // a few waves hide ramp-up/tail
    if (e > beste + 1e-9) { beste = e; best = kc; }
  }
  return best;
}

// tables indexed ib + 2*idir
__host__ __device__ inline int tb(int ib, int idir) { return ib + 2 * idir; }

// internal device-level entry points shared between translation units (no host sync)
int k_halo_exchange(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields);
int k_halo_exchange_dirs(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields, int dirmask);
int k_boundp(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
             const int is_bound[6], const double dl[3], const double* dzc, double* p);
int k_boundp_multi(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
                   const int is_bound[6], const double dl[3], const double* dzc, double* const* ps, int np);
int k_allreduce_sum(cales_ctx* ctx, double* dev, int count);
PeerBuf* k_peer_buffer(cales_ctx* ctx, const char* name, size_t bytes);   // collective; nullptr if peer mapping is unavailable
int k_barrier(cales_ctx* ctx);                                              // stream-ordered barrier over all ranks
int k_allreduce_minmax(cales_ctx* ctx, double* dev, int count, int is_max);
FftTables* k_tables(cales_ctx* ctx, int n);
