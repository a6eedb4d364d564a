// Communication: NCCL bootstrap, halo exchange, scalar/plane reductions.
//   updthalo      src/bound.f90:619-696  (MPI_SENDRECV of full-extent faces, dimension by dimension)
//   updthalo_gpu  src/bound.f90:698-723  (cuDecomp halos; periodic single-rank self-copy halo.h:143-171)
//   NCCL bootstrap as dependencies/cuDecomp/src/cudecomp.cc:66-80 (unique id broadcast by the host)
// One process per GPU; every collective is enqueued on the context stream, so the caller's program
// order is kept without host synchronisation.  All fields of one call share a single packed
// message per neighbour (the reference sends one message per field and direction).
#include <nccl.h>

#include <cstdlib>

#include <algorithm>
#include <cstdint>

#include "common.cuh"

// NCCL is bound at run time (dlopen) and only when nranks > 1: a host process that already carries an NCCL
// (PyTorch bundles its own libnccl.so.2) keeps using that one, and single-GPU use needs no NCCL at all.
#include <dlfcn.h>
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(cales_ctx* ctx) {
  if (g_nccl.h) return CALES_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already in the process (e.g. torch's)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return cales_fail(ctx, CALES_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(name) *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name); if (!g_nccl.name) return cales_fail(ctx, CALES_ERR_NCCL, "libnccl lacks nccl" #name)
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllReduce); SYM(AllGather); SYM(Send); SYM(Recv); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
  g_nccl.h = h;
  return CALES_OK;
}

#define NCCL_TRY(ctx, call)                                                                              \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) return cales_fail(ctx, CALES_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)

extern "C" int cales_get_unique_id(char uid[CALES_UNIQUE_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= CALES_UNIQUE_ID_BYTES, "unique id size");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return cales_fail(nullptr, CALES_ERR_NCCL, "ncclGetUniqueId failed");
  memset(uid, 0, CALES_UNIQUE_ID_BYTES);
  memcpy(uid, &id, sizeof id);
  return CALES_OK;
}

int comm_init(cales_ctx* ctx, const char* uid) {
  if (!uid) return cales_fail(ctx, CALES_ERR_INVALID, "nranks>1 requires an NCCL unique id (cales_get_unique_id on rank 0, broadcast by the host)");
  int rc = nccl_load(ctx);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, uid, sizeof id);
  ncclComm_t comm;
  NCCL_TRY(ctx, g_nccl.CommInitRank(&comm, ctx->nranks, id, ctx->rank));
  ctx->nccl = comm;
  return CALES_OK;
}

void comm_finalize(cales_ctx* ctx) {
  if (ctx->nccl && !ctx->peerbufs.empty()) { k_barrier(ctx); cudaStreamSynchronize(ctx->stream); }   // nobody still stores into my buffers
  for (auto& kv : ctx->peerbufs) {
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && kv.second.ptr[r]) cudaIpcCloseMemHandle(kv.second.ptr[r]);
    cudaFree(kv.second.local);
  }
  ctx->peerbufs.clear();
  if (ctx->nccl) { g_nccl.CommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
}

int k_allreduce_sum(cales_ctx* ctx, double* dev, int count) {
  if (ctx->nranks == 1) return CALES_OK;
  NCCL_TRY(ctx, g_nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return CALES_OK;
}

int k_allreduce_minmax(cales_ctx* ctx, double* dev, int count, int is_max) {
  if (ctx->nranks == 1) return CALES_OK;
  NCCL_TRY(ctx, g_nccl.AllReduce(dev, dev, count, ncclDouble, is_max ? ncclMax : ncclMin, (ncclComm_t)ctx->nccl, ctx->stream));
  return CALES_OK;
}

// ---- peer memory (CUDA IPC over NVLink / NVSwitch) --------------------------------------------------------------
// Every rank allocates the buffer, publishes its IPC handle with one ncclAllGather and maps the handles of all other
// ranks, so kernels can store straight into a peer's pencil (the transposes below, and the fused transform/solve
// kernels of the distributed solver).  CALES_NO_P2P=1 keeps the NCCL send/recv path.
PeerBuf* k_peer_buffer(cales_ctx* ctx, const char* name, size_t bytes) {
  if (ctx->nranks == 1 || !ctx->nccl || ctx->nranks > CALES_MAX_RANKS) return nullptr;
  if (ctx->p2p == -1) ctx->p2p = getenv("CALES_NO_P2P") ? 0 : 1;
  if (ctx->p2p == 0) return nullptr;
  auto it = ctx->peerbufs.find(name);
  if (it != ctx->peerbufs.end() && it->second.bytes >= bytes) return &it->second;
  if (it != ctx->peerbufs.end()) {                      // grow: unmap, free, re-create (collective: all ranks take this branch together)
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && it->second.ptr[r]) cudaIpcCloseMemHandle(it->second.ptr[r]);
    k_barrier(ctx); cudaStreamSynchronize(ctx->stream);
    cudaFree(it->second.local);
    ctx->peerbufs.erase(it);
  }
  PeerBuf pb;
  pb.bytes = bytes;
  if (cudaMalloc(&pb.local, bytes) != cudaSuccess) { cales_fail(ctx, CALES_ERR_NOMEM, "cudaMalloc(%zu) for peer buffer '%s' failed", bytes, name); return nullptr; }
  cudaMemsetAsync(pb.local, 0, bytes, ctx->stream);     // before the handle exchange below: no peer can touch it earlier
  cudaIpcMemHandle_t mine;
  const int n = ctx->nranks;
  bool ok = cudaIpcGetMemHandle(&mine, pb.local) == cudaSuccess;
  // exchange {ok flag, handle} records
  const size_t rec = 128;
  static_assert(sizeof(cudaIpcMemHandle_t) + 8 <= 128, "ipc record");
  std::vector<char> h(rec * n, 0);
  char* dsend = nullptr; char* drecv = nullptr;
  cudaMalloc(&dsend, rec); cudaMalloc(&drecv, rec * n);
  char mrec[128] = {0};
  mrec[0] = ok ? 1 : 0;
  memcpy(mrec + 8, &mine, sizeof mine);
  cudaMemcpyAsync(dsend, mrec, rec, cudaMemcpyHostToDevice, ctx->stream);
  const bool sent = g_nccl.AllGather(dsend, drecv, rec, ncclChar, (ncclComm_t)ctx->nccl, ctx->stream) == ncclSuccess;
  cudaMemcpyAsync(h.data(), drecv, rec * n, cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(dsend); cudaFree(drecv);
  bool all = sent;
  for (int r = 0; r < n; ++r) all = all && h[rec * r] == 1;
  if (all) {
    for (int r = 0; r < n && all; ++r) {
      if (r == ctx->rank) { pb.ptr[r] = pb.local; continue; }
      cudaIpcMemHandle_t hd;
      memcpy(&hd, h.data() + rec * r + 8, sizeof hd);
      if (cudaIpcOpenMemHandle(&pb.ptr[r], hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) all = false;
    }
  }
  // every rank must reach the same verdict: sum of failures
  double fail = all ? 0. : 1.;
  if (!ctx->bar) cudaMalloc(&ctx->bar, sizeof(double));
  cudaMemcpyAsync(ctx->bar, &fail, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  g_nccl.AllReduce(ctx->bar, ctx->bar, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream);
  cudaMemcpyAsync(&fail, ctx->bar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaGetLastError();
  if (fail != 0.) {
    for (int r = 0; r < n; ++r) if (r != ctx->rank && pb.ptr[r]) cudaIpcCloseMemHandle(pb.ptr[r]);
    cudaFree(pb.local);
    ctx->p2p = 0;
    fprintf(stderr, "cales_b200: CUDA IPC peer mapping unavailable on rank %d; using NCCL send/recv transposes\n", ctx->rank);
    return nullptr;
  }
  ctx->peerbufs[name] = pb;
  return &ctx->peerbufs[name];
}

// Stream-ordered barrier over all ranks.  With peer memory: one tiny kernel -- thread r stores the barrier's sequence
// number into rank r's flag word for me (st.release.sys over NVLink) and spins on my flag word for rank r
// (ld.acquire.sys), so everything the ranks stored into each other's memory before the barrier is visible after it.
// A few microseconds, against ~15 for the one-element NCCL all-reduce that remains the fallback.
struct BarArgs { unsigned long long* flags[CALES_MAX_RANKS]; int n, me; unsigned long long seq; };

__global__ void p2p_barrier_k(BarArgs A) {
  const int r = threadIdx.x;
  if (r >= A.n || r == A.me) return;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.flags[r] + A.me), "l"(A.seq) : "memory");
  const unsigned long long* mine = A.flags[A.me] + r;
  unsigned long long v;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
  } while (v < A.seq && clock64() - t0 < 60000000000LL);      // ~30 s: a peer died; fail loudly instead of hanging
  if (v < A.seq) __trap();
}

int k_barrier(cales_ctx* ctx) {
  if (ctx->nranks == 1) return CALES_OK;
  static const bool nccl_only = getenv("CALES_NCCL_BARRIER") != nullptr;
  PeerBuf* fb = nccl_only ? nullptr : k_peer_buffer(ctx, "barrier_flags", CALES_MAX_RANKS * sizeof(unsigned long long));
  if (fb) {
    BarArgs A;
    for (int r = 0; r < ctx->nranks; ++r) A.flags[r] = (unsigned long long*)fb->ptr[r];
    A.n = ctx->nranks; A.me = ctx->rank; A.seq = ++ctx->bar_seq;
    p2p_barrier_k<<<1, CALES_MAX_RANKS, 0, ctx->stream>>>(A);
    KERNEL_CHECK(ctx);
    return CALES_OK;
  }
  if (!ctx->bar) { CUDA_TRY(ctx, cudaMalloc(&ctx->bar, sizeof(double))); CUDA_TRY(ctx, cudaMemsetAsync(ctx->bar, 0, sizeof(double), ctx->stream)); }
  NCCL_TRY(ctx, g_nccl.AllReduce(ctx->bar, ctx->bar, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return CALES_OK;
}

// ---- halo exchange ------------------------------------------------------------------------------------
struct FieldList { double* p[12]; int nf; };

// face geometry for direction idir: m1 x m2 points, element (a,c) of the plane at normal index q
__device__ __forceinline__ long face_idx(const Dims& d, int idir, int q, int a, int c) {
  return idir == 0 ? d.idx(q, a, c) : idir == 1 ? d.idx(a, q, c) : d.idx(a, c, q);
}

// periodic self-neighbour: p(0)=p(n), p(n+1)=p(1) (cuDecomp halo.h:143-171; MPI_SENDRECV to self)
__global__ void __launch_bounds__(256) halo_self_k(Dims d, int idir, FieldList fl) {
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  double* p = fl.p[blockIdx.z];
  p[face_idx(d, idir, 0, a, c)] = p[face_idx(d, idir, n, a, c)];
  p[face_idx(d, idir, n + 1, a, c)] = p[face_idx(d, idir, 1, a, c)];
}

// pack planes 1 (-> buf[0..]) and n (-> buf[nf*m..]) of every field; unpack ghost planes 0 and n+1
__global__ void __launch_bounds__(256) halo_pack_k(Dims d, int idir, FieldList fl, double* __restrict__ buf) {
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  const long m = (long)m1 * m2;
  const int f = blockIdx.z;
  const double* p = fl.p[f];
  buf[f * m + a + (long)m1 * c] = p[face_idx(d, idir, 1, a, c)];
  buf[(fl.nf + f) * m + a + (long)m1 * c] = p[face_idx(d, idir, n, a, c)];
}

// peer-memory variant of the pack: plane 1 goes straight into the upper-ghost region of neighbour nb(0)'s receive buffer
// (`to_lo`), plane n into the lower-ghost region of nb(1)'s (`to_hi`): NVLink stores, no send/recv
__global__ void __launch_bounds__(256) halo_push_k(Dims d, int idir, FieldList fl, double* __restrict__ to_lo, double* __restrict__ to_hi) {
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  const long m = (long)m1 * m2;
  const int f = blockIdx.z;
  const double* p = fl.p[f];
  if (to_lo) to_lo[f * m + a + (long)m1 * c] = p[face_idx(d, idir, 1, a, c)];
  if (to_hi) to_hi[f * m + a + (long)m1 * c] = p[face_idx(d, idir, n, a, c)];
}

__global__ void __launch_bounds__(256) halo_unpack_k(Dims d, int idir, FieldList fl, const double* __restrict__ buf, int has_lo, int has_hi) {
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int a = blockIdx.x * 64 + threadIdx.x, c = blockIdx.y * 4 + threadIdx.y;
  if (a >= m1 || c >= m2) return;
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  const long m = (long)m1 * m2;
  const int f = blockIdx.z;
  double* p = fl.p[f];
  if (has_lo) p[face_idx(d, idir, 0, a, c)] = buf[f * m + a + (long)m1 * c];
  if (has_hi) p[face_idx(d, idir, n + 1, a, c)] = buf[(fl.nf + f) * m + a + (long)m1 * c];
}

// One-launch peer exchange of one direction: a persistent grid (every CTA resident) pushes its tiles of the boundary planes
// into the neighbours' receive buffers, the last CTA to finish raises the neighbours' arrival flags (st.release.sys after
// every CTA's system-scope fence), then every CTA waits for MY arrival flags (ld.acquire.sys) and unpacks its tiles of the
// ghost planes: what used to be push kernel + barrier kernel + unpack kernel, with only the two neighbours involved.
struct HaloX {
  double *to_lo, *to_hi;                 // receive regions in nb(0) / nb(1) (nullptr: no neighbour)
  const double* mine;                    // my receive buffer (lower ghosts first, then upper)
  unsigned long long *sig_lo, *sig_hi;   // flag words in nb(0) / nb(1) to raise
  const unsigned long long *wait_lo, *wait_hi;   // my flag words raised by nb(0) / nb(1)
  unsigned* counter;                     // CTAs done pushing (reset by the last one)
  unsigned long long seq;
};

__global__ void __launch_bounds__(256) halo_fused_k(Dims d, int idir, FieldList fl, HaloX X, int gx, int gy) {
  const int m1 = idir == 0 ? d.n2 + 2 : d.n1 + 2, m2 = idir == 2 ? d.n2 + 2 : d.n3 + 2;
  const int n = idir == 0 ? d.n1 : idir == 1 ? d.n2 : d.n3;
  const long m = (long)m1 * m2;
  const int ntile = gx * gy * fl.nf;
  for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int f = t / (gx * gy), r = t - f * gx * gy;
    const int a = (r % gx) * 64 + threadIdx.x, c = (r / gx) * 4 + threadIdx.y;
    if (a >= m1 || c >= m2) continue;
    const double* p = fl.p[f];
    if (X.to_lo) X.to_lo[f * m + a + (long)m1 * c] = p[face_idx(d, idir, 1, a, c)];
    if (X.to_hi) X.to_hi[f * m + a + (long)m1 * c] = p[face_idx(d, idir, n, a, c)];
  }
  __threadfence_system();
  __syncthreads();
  const bool lead = threadIdx.x == 0 && threadIdx.y == 0;
  if (lead) {
    const unsigned old = atomicAdd(X.counter, 1u);
    if (old == gridDim.x - 1) {
      *X.counter = 0;
      __threadfence_system();
      if (X.sig_lo) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(X.sig_lo), "l"(X.seq) : "memory");
      if (X.sig_hi) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(X.sig_hi), "l"(X.seq) : "memory");
    }
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
      const unsigned long long* w = side == 0 ? X.wait_lo : X.wait_hi;
      if (!w) continue;
      unsigned long long v;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
      } while (v < X.seq && clock64() - t0 < 60000000000LL);
      if (v < X.seq) __trap();              // a neighbour died: fail loudly instead of hanging
    }
  }
  __syncthreads();
  for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int f = t / (gx * gy), r = t - f * gx * gy;
    const int a = (r % gx) * 64 + threadIdx.x, c = (r / gx) * 4 + threadIdx.y;
    if (a >= m1 || c >= m2) continue;
    double* p = fl.p[f];
    if (X.wait_lo) p[face_idx(d, idir, 0, a, c)] = __ldcv(X.mine + f * m + a + (long)m1 * c);
    if (X.wait_hi) p[face_idx(d, idir, n + 1, a, c)] = __ldcv(X.mine + (fl.nf + f) * m + a + (long)m1 * c);
  }
}

int k_halo_exchange_dirs(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields, int dirmask);

int k_halo_exchange(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields) {
  return k_halo_exchange_dirs(ctx, n, nb, fields, nfields, 7);
}

int k_halo_exchange_dirs(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields, int dirmask) {
  Dims d(n);
  for (int f0 = 0; f0 < nfields; f0 += 12) {
    FieldList fl;
    fl.nf = nfields - f0 < 12 ? nfields - f0 : 12;
    for (int f = 0; f < fl.nf; ++f) fl.p[f] = fields[f0 + f];
    for (int idir = 0; idir < 3; ++idir) {
      if (idir + 1 == ctx->ipencil) continue;                 // bound.f90:634
      if (!(dirmask & (1 << idir))) continue;
      const int nb0 = nb[tb(0, idir)], nb1 = nb[tb(1, idir)];
      if (nb0 < 0 && nb1 < 0) continue;
      const int m1 = idir == 0 ? n[1] + 2 : n[0] + 2, m2 = idir == 2 ? n[1] + 2 : n[2] + 2;
      dim3 g(cdiv(m1, 64), cdiv(m2, 4), fl.nf), b(64, 4);
      if (nb0 == ctx->rank && nb1 == ctx->rank) {
        halo_self_k<<<g, b, 0, ctx->stream>>>(d, idir, fl);
        KERNEL_CHECK(ctx);
        continue;
      }
      if (!ctx->nccl) return cales_fail(ctx, CALES_ERR_INVALID, "halo exchange with rank %d/%d requested on a single-rank context", nb0, nb1);
      const long m = (long)m1 * m2, cnt = m * fl.nf;
      const long fxy = (long)(n[0] + 2) * (n[1] + 2), fxz = (long)(n[0] + 2) * (n[2] + 2), fyz = (long)(n[1] + 2) * (n[2] + 2);
      const long fmax_ = fxy > fxz ? (fxy > fyz ? fxy : fyz) : (fxz > fyz ? fxz : fyz);
      // peer memory: push my boundary planes into the neighbours' receive buffers, barrier, unpack mine.  The receive
      // buffer is double-buffered: a neighbour may still be unpacking exchange s while I push exchange s+1, but it has
      // passed the barrier of s+1 -- hence finished unpacking s -- before I push s+2.
      static const bool nccl_halo = getenv("CALES_NCCL_HALO") != nullptr;
      size_t fglob = 0;            // the largest face of any rank's pencil: both sides of a face must agree on the buffer layout
      for (int r = 0; r < ctx->nranks; ++r) {
        int lo_[3], hi_[3], sz_[3];
        cales_pencil(ctx->ng, ctx->dims, r, ctx->ipencil, lo_, hi_, sz_);
        const size_t a_ = sz_[0] + 2, b_ = sz_[1] + 2, c_ = sz_[2] + 2;
        fglob = std::max(fglob, std::max(a_ * b_, std::max(a_ * c_, b_ * c_)));
      }
      const size_t half = (size_t)2 * 12 * fglob;
      PeerBuf* hb = nccl_halo || (size_t)fmax_ > fglob ? nullptr : k_peer_buffer(ctx, "halo_peer", 2 * half * sizeof(double));
      static const bool fused_halo = !(getenv("CALES_HALO_FUSED") && atoi(getenv("CALES_HALO_FUSED")) == 0);
      PeerBuf* hf = hb && fused_halo ? k_peer_buffer(ctx, "halo_flags", 64 * sizeof(unsigned long long)) : nullptr;
      if (hb && hf) {
        if (!ctx->halo_counter) { CUDA_TRY(ctx, cudaMalloc(&ctx->halo_counter, 4 * sizeof(unsigned))); CUDA_TRY(ctx, cudaMemsetAsync(ctx->halo_counter, 0, 4 * sizeof(unsigned), ctx->stream)); }
        const size_t off = (ctx->halo_seq++ & 1u) * half;
        HaloX X;
        X.to_lo = nb0 >= 0 ? (double*)hb->ptr[nb0] + off + cnt : nullptr;      // my plane 1 = nb(0)'s upper ghost
        X.to_hi = nb1 >= 0 ? (double*)hb->ptr[nb1] + off : nullptr;            // my plane n = nb(1)'s lower ghost
        X.mine = (const double*)hb->local + off;
        X.sig_lo = nb0 >= 0 ? (unsigned long long*)hf->ptr[nb0] + 2 * idir + 1 : nullptr;
        X.sig_hi = nb1 >= 0 ? (unsigned long long*)hf->ptr[nb1] + 2 * idir + 0 : nullptr;
        X.wait_lo = nb0 >= 0 ? (const unsigned long long*)hf->local + 2 * idir + 0 : nullptr;
        X.wait_hi = nb1 >= 0 ? (const unsigned long long*)hf->local + 2 * idir + 1 : nullptr;
        X.counter = ctx->halo_counter + idir;
        X.seq = ++ctx->halo_dir_seq[idir];
        const int gx = cdiv(m1, 64), gy = cdiv(m2, 4);
        const int ntile = gx * gy * fl.nf;
        halo_fused_k<<<std::min(ntile, 148 * 4), b, 0, ctx->stream>>>(d, idir, fl, X, gx, gy);
        KERNEL_CHECK(ctx);
        continue;
      }
      if (hb) {
        const size_t off = (ctx->halo_seq++ & 1u) * half;
        double* to_lo = nb0 >= 0 ? (double*)hb->ptr[nb0] + off + cnt : nullptr;
        double* to_hi = nb1 >= 0 ? (double*)hb->ptr[nb1] + off : nullptr;
        halo_push_k<<<g, b, 0, ctx->stream>>>(d, idir, fl, to_lo, to_hi);
        KERNEL_CHECK(ctx);
        int rcb;
        if ((rcb = k_barrier(ctx))) return rcb;
        halo_unpack_k<<<g, b, 0, ctx->stream>>>(d, idir, fl, (double*)hb->local + off, nb0 >= 0, nb1 >= 0);
        KERNEL_CHECK(ctx);
        continue;
      }
      double* sbuf = (double*)cales_scratch(ctx, "halo_send", (size_t)2 * 12 * sizeof(double) * (size_t)fmax_);
      double* rbuf = (double*)cales_scratch(ctx, "halo_recv", ctx->scratch["halo_send"].second);
      if (!sbuf || !rbuf) return CALES_ERR_NOMEM;
      halo_pack_k<<<g, b, 0, ctx->stream>>>(d, idir, fl, sbuf);
      KERNEL_CHECK(ctx);
      ncclComm_t comm = (ncclComm_t)ctx->nccl;
      NCCL_TRY(ctx, g_nccl.GroupStart());
      // sends: my plane 1 to nb(0), my plane n to nb(1); receives in the matching order for the
      // two-rank periodic case (the peer's first message is its plane 1 = my upper ghost)
      if (nb0 >= 0) NCCL_TRY(ctx, g_nccl.Send(sbuf, cnt, ncclDouble, nb0, comm, ctx->stream));
      if (nb1 >= 0) NCCL_TRY(ctx, g_nccl.Send(sbuf + cnt, cnt, ncclDouble, nb1, comm, ctx->stream));
      if (nb1 >= 0) NCCL_TRY(ctx, g_nccl.Recv(rbuf + cnt, cnt, ncclDouble, nb1, comm, ctx->stream));
      if (nb0 >= 0) NCCL_TRY(ctx, g_nccl.Recv(rbuf, cnt, ncclDouble, nb0, comm, ctx->stream));
      NCCL_TRY(ctx, g_nccl.GroupEnd());
      halo_unpack_k<<<g, b, 0, ctx->stream>>>(d, idir, fl, rbuf, nb0 >= 0, nb1 >= 0);
      KERNEL_CHECK(ctx);
    }
  }
  return CALES_OK;
}

extern "C" int cales_updthalo(cales_ctx* ctx, const int n[3], const int nb[6], double* p) {
  CHECK_CTX(ctx);
  double* ps[1] = {p};
  return k_halo_exchange(ctx, n, nb, ps, 1);
}

// ---- pencil transposes ------------------------------------------------------------------------------------------
// 2decomp transpose_x_to_y / y_to_z / z_to_y / y_to_x (dependencies/2decomp-fft/src/transpose_x_to_y.f90:25-115)
// == cudecompTransposeXToY ... (dependencies/cuDecomp/include/internal/transpose.h:160-729).
// Pencil A is complete along axis `al` and split along `be` over the P ranks of a row/column of the process
// grid; pencil B is complete along `be` and split along `al`.  Rank q receives my sub-box al in range_al(q);
// I receive from q the sub-box be in range_be(q).  One pack launch builds all peer slabs, one NCCL group moves
// them over NVLink (the self slab never leaves the device), one unpack launch scatters them.
struct Seg { long soff, doff; int b0, b1, b2; long ss1, ss2, ds1, ds2; double* dptr; };
struct SegList { Seg s[16]; int n; };

__global__ void __launch_bounds__(256) boxcopy_k(const double* __restrict__ src, double* __restrict__ dst, SegList L) {
  const Seg& g = L.s[blockIdx.z];
  const int i = blockIdx.x * 64 + threadIdx.x;
  if (i >= g.b0) return;
  const long rows = (long)g.b1 * g.b2;
  for (long r = blockIdx.y * 4 + threadIdx.y; r < rows; r += (long)gridDim.y * 4) {
    const int j = (int)(r % g.b1), k = (int)(r / g.b1);
    (g.dptr ? g.dptr : dst)[g.doff + i + j * g.ds1 + k * g.ds2] = src[g.soff + i + j * g.ss1 + k * g.ss2];
  }
}

// 16-byte variant: every extent, offset and stride of every segment is even and all pointers are 16-byte aligned
__global__ void __launch_bounds__(256) boxcopy2_k(const double* __restrict__ src, double* __restrict__ dst, SegList L) {
  const Seg& g = L.s[blockIdx.z];
  const int i = 2 * (blockIdx.x * 64 + threadIdx.x);
  if (i >= g.b0) return;
  const long rows = (long)g.b1 * g.b2;
  double* out = g.dptr ? g.dptr : dst;
  for (long r = blockIdx.y * 4 + threadIdx.y; r < rows; r += (long)gridDim.y * 4) {
    const int j = (int)(r % g.b1), k = (int)(r / g.b1);
    *reinterpret_cast<double2*>(out + g.doff + i + j * g.ds1 + k * g.ds2) = *reinterpret_cast<const double2*>(src + g.soff + i + j * g.ss1 + k * g.ss2);
  }
}

// Copy-engine variant for boxes that are complete along i in both pencils (the y <-> z transposes): per peer ONE strided
// 2-D DMA (rows of b0*b1 contiguous values, b2 rows), each on its own stream so that the peers' links run concurrently;
// NVLink at copy-engine efficiency and no SM time.  Returns 0 if some segment does not collapse to 2-D.
static cudaStream_t g_ce_stream[8];
static cudaEvent_t g_ce_ev[17];
static bool g_ce_init = false;
static int boxcopy_ce(cales_ctx* ctx, const double* src, double* dst, const SegList& L, int first) {
  for (int q = 0; q < L.n; ++q) if (L.s[q].b0 != L.s[q].ss1 || L.s[q].b0 != L.s[q].ds1) return 0;
  if (!g_ce_init) {
    g_ce_init = true;
    for (auto& st : g_ce_stream) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto& e : g_ce_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  }
  cudaEventRecord(g_ce_ev[16], ctx->stream);
  for (int d = 0; d < L.n; ++d) {
    const int q = (first + d) % L.n;
    const Seg& g = L.s[q];
    cudaStream_t st = g_ce_stream[d % 8];
    cudaStreamWaitEvent(st, g_ce_ev[16], 0);
    if (cudaMemcpy2DAsync((g.dptr ? g.dptr : dst) + g.doff, (size_t)g.ds2 * sizeof(double), src + g.soff, (size_t)g.ss2 * sizeof(double),
                          (size_t)g.b0 * g.b1 * sizeof(double), (size_t)g.b2, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return -cales_fail(ctx, CALES_ERR_CUDA, "peer cudaMemcpy2DAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaEventRecord(g_ce_ev[d % 16], st);
    cudaStreamWaitEvent(ctx->stream, g_ce_ev[d % 16], 0);
  }
  ctx->launches += L.n;
  return 1;
}

static int boxcopy(cales_ctx* ctx, const double* src, double* dst, const SegList& L) {
  static const int mode = getenv("CALES_TRANSPOSE_MODE") ? atoi(getenv("CALES_TRANSPOSE_MODE")) : 1;   // 0: 8-byte kernel, 1: 16-byte kernel, 2: copy engines
  if (mode == 2 && L.n <= 16) {
    const int rc = boxcopy_ce(ctx, src, dst, L, (ctx->coord[1] + 1) % L.n);
    if (rc < 0) return -rc;
    if (rc == 1) return CALES_OK;
  }
  bool even = mode >= 1 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  for (int q = 0; q < L.n && even; ++q) {
    const Seg& g = L.s[q];
    even = !((g.b0 | g.soff | g.doff | g.ss1 | g.ss2 | g.ds1 | g.ds2) & 1) && ((uintptr_t)g.dptr & 15) == 0;
  }
  if (even) {
    int b0 = 2; long rows = 1;
    for (int q = 0; q < L.n; ++q) { if (L.s[q].b0 > b0) b0 = L.s[q].b0; if ((long)L.s[q].b1 * L.s[q].b2 > rows) rows = (long)L.s[q].b1 * L.s[q].b2; }
    long gy = (rows + 3) / 4; if (gy > 16384) gy = 16384;
    boxcopy2_k<<<dim3(cdiv(b0 / 2, 64), (unsigned)gy, L.n), dim3(64, 4), 0, ctx->stream>>>(src, dst, L);
    KERNEL_CHECK(ctx);
    return CALES_OK;
  }
  int b0 = 1; long rows = 1;
  for (int q = 0; q < L.n; ++q) { if (L.s[q].b0 > b0) b0 = L.s[q].b0; if ((long)L.s[q].b1 * L.s[q].b2 > rows) rows = (long)L.s[q].b1 * L.s[q].b2; }
  long gy = (rows + 3) / 4; if (gy > 16384) gy = 16384;
  boxcopy_k<<<dim3(cdiv(b0, 64), (unsigned)gy, L.n), dim3(64, 4), 0, ctx->stream>>>(src, dst, L);
  KERNEL_CHECK(ctx);
  return CALES_OK;
}

// Pure-integer exchange plan of one transpose (usable without a device; the world_size-2 gloo test drives it).
// For every peer q of my row/column: global rank, the sub-box of pencil A sent to q and the sub-box of pencil B
// received from q, each as {offset(3), extent(3)} in local 0-based indices.  Returns the pencil shapes too.
extern "C" int cales_transpose_plan(const int ng[3], const int dims[2], int rank, int which, int* npeers, int* peers,
                                    int* sendbox, int* recvbox, int shapeA[3], int shapeB[3]) {
  if (which < 0 || which > 3 || rank < 0 || rank >= dims[0] * dims[1]) return CALES_ERR_INVALID;
  int lo[3], hi[3], A[3], B[3];
  const int axA = which == 0 ? 1 : (which == 1 || which == 3) ? 2 : 3;
  const int axB = which == 0 ? 2 : which == 1 ? 3 : which == 2 ? 2 : 1;
  cales_pencil(ng, dims, rank, axA, lo, hi, A);
  cales_pencil(ng, dims, rank, axB, lo, hi, B);
  const int al = which == 0 ? 0 : which == 1 ? 1 : which == 2 ? 2 : 1;
  const int be = which == 0 ? 1 : which == 1 ? 2 : which == 2 ? 1 : 0;
  const bool colcomm = (which == 0 || which == 3);
  const int P = colcomm ? dims[0] : dims[1];
  const int coord[2] = {rank / dims[1], rank % dims[1]};
  std::vector<int> ast(P), aen(P), asz(P), bst(P), ben(P), bsz(P);
  cales_distribute(ng[al], P, ast.data(), aen.data(), asz.data());
  cales_distribute(ng[be], P, bst.data(), ben.data(), bsz.data());
  *npeers = P;
  for (int q = 0; q < P; ++q) {
    peers[q] = colcomm ? q * dims[1] + coord[1] : coord[0] * dims[1] + q;
    int* sb = sendbox + 6 * q; int* rb = recvbox + 6 * q;
    for (int d = 0; d < 3; ++d) { sb[d] = 0; sb[3 + d] = A[d]; rb[d] = 0; rb[3 + d] = B[d]; }
    sb[al] = ast[q] - 1; sb[3 + al] = asz[q];
    rb[be] = bst[q] - 1; rb[3 + be] = bsz[q];
  }
  for (int d = 0; d < 3; ++d) { shapeA[d] = A[d]; shapeB[d] = B[d]; }
  return CALES_OK;
}

int k_transpose(cales_ctx* ctx, int which, const double* src, double* dst) {
  // which: 0 x->y, 1 y->z, 2 z->y, 3 y->x
  int P, peers[16], sendbox[96], recvbox[96], A[3], B[3];
  const bool colcomm = (which == 0 || which == 3);
  if ((colcomm ? ctx->dims[0] : ctx->dims[1]) > 16) return cales_fail(ctx, CALES_ERR_INVALID, "transpose: at most 16 ranks per row/column supported");
  if (cales_transpose_plan(ctx->ng, ctx->dims, ctx->rank, which, &P, peers, sendbox, recvbox, A, B)) return cales_fail(ctx, CALES_ERR_INVALID, "transpose plan failed");
  const int me = colcomm ? ctx->coord[0] : ctx->coord[1];
  const size_t na = (size_t)A[0] * A[1] * A[2], nb_ = (size_t)B[0] * B[1] * B[2];
  if (P == 1) {                                                              // same data, same layout
    if (src != dst) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, na * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return CALES_OK;
  }
  double* sbuf = (double*)cales_scratch(ctx, "tr_send", (na > nb_ ? na : nb_) * sizeof(double));
  double* rbuf = (double*)cales_scratch(ctx, "tr_recv", (na > nb_ ? na : nb_) * sizeof(double));
  if (!sbuf || !rbuf) return CALES_ERR_NOMEM;
  SegList pk, up; pk.n = up.n = P;
  long soff[16], roff[16], scnt[16], rcnt[16];
  long so = 0, ro = 0;
  const long As1 = A[0], As2 = (long)A[0] * A[1], Bs1 = B[0], Bs2 = (long)B[0] * B[1];
  for (int q = 0; q < P; ++q) {
    const int* sb = sendbox + 6 * q; const int* rb = recvbox + 6 * q;
    Seg& g = pk.s[q];
    g.dptr = nullptr; up.s[q].dptr = nullptr;
    g.b0 = sb[3]; g.b1 = sb[4]; g.b2 = sb[5];
    g.soff = sb[0] + sb[1] * As1 + sb[2] * As2; g.ss1 = As1; g.ss2 = As2;
    g.doff = so; g.ds1 = sb[3]; g.ds2 = (long)sb[3] * sb[4];
    soff[q] = so; scnt[q] = (long)sb[3] * sb[4] * sb[5]; so += scnt[q];
    Seg& h = up.s[q];
    h.b0 = rb[3]; h.b1 = rb[4]; h.b2 = rb[5];
    h.soff = ro; h.ss1 = rb[3]; h.ss2 = (long)rb[3] * rb[4];
    h.doff = rb[0] + rb[1] * Bs1 + rb[2] * Bs2; h.ds1 = Bs1; h.ds2 = Bs2;
    roff[q] = ro; rcnt[q] = (long)rb[3] * rb[4] * rb[5]; ro += rcnt[q];
  }
  int rc;
  if ((rc = boxcopy(ctx, src, sbuf, pk))) return rc;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  NCCL_TRY(ctx, g_nccl.GroupStart());
  for (int d = 1; d < P; ++d) {
    const int qs = (me + d) % P, qr = (me - d + P) % P;
    NCCL_TRY(ctx, g_nccl.Send(sbuf + soff[qs], scnt[qs], ncclDouble, peers[qs], comm, ctx->stream));
    NCCL_TRY(ctx, g_nccl.Recv(rbuf + roff[qr], rcnt[qr], ncclDouble, peers[qr], comm, ctx->stream));
  }
  NCCL_TRY(ctx, g_nccl.GroupEnd());
  // the self slab never leaves the device
  CUDA_TRY(ctx, cudaMemcpyAsync(rbuf + roff[me], sbuf + soff[me], scnt[me] * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return boxcopy(ctx, rbuf, dst, up);
}

// Peer-memory transpose: ONE kernel reads my pencil and stores every sub-box straight into the destination pencil of
// its owner (local or over NVLink) in its final layout -- no pack, no unpack, no staging buffers -- followed by a
// stream-ordered barrier.  dst must be a peer buffer (k_peer_buffer); the caller guarantees by the barrier discipline
// of the solver that no rank still reads it.
int k_transpose_p2p(cales_ctx* ctx, int which, const double* src, PeerBuf* dst) {
  int P, peers[16], sendbox[96], recvbox[96], A[3], B[3];
  const bool colcomm = (which == 0 || which == 3);
  if ((colcomm ? ctx->dims[0] : ctx->dims[1]) > 16) return cales_fail(ctx, CALES_ERR_INVALID, "transpose: at most 16 ranks per row/column supported");
  if (cales_transpose_plan(ctx->ng, ctx->dims, ctx->rank, which, &P, peers, sendbox, recvbox, A, B)) return cales_fail(ctx, CALES_ERR_INVALID, "transpose plan failed");
  const int me = colcomm ? ctx->coord[0] : ctx->coord[1];
  const int axB = which == 0 ? 2 : which == 1 ? 3 : which == 2 ? 2 : 1;
  const int be = which == 0 ? 1 : which == 1 ? 2 : which == 2 ? 1 : 0;
  std::vector<int> bst(P), ben(P), bsz(P);
  cales_distribute(ctx->ng[be], P, bst.data(), ben.data(), bsz.data());
  SegList L; L.n = P;
  const long As1 = A[0], As2 = (long)A[0] * A[1];
  for (int q = 0; q < P; ++q) {
    const int* sb = sendbox + 6 * q;
    int lo[3], hi[3], Bq[3];
    cales_pencil(ctx->ng, ctx->dims, peers[q], axB, lo, hi, Bq);          // shape of the destination pencil on rank q
    Seg& g = L.s[q];
    g.b0 = sb[3]; g.b1 = sb[4]; g.b2 = sb[5];
    g.soff = sb[0] + sb[1] * As1 + sb[2] * As2; g.ss1 = As1; g.ss2 = As2;
    int off[3] = {0, 0, 0};
    off[be] = bst[me] - 1;                                               // my slab of the re-assembled direction
    g.ds1 = Bq[0]; g.ds2 = (long)Bq[0] * Bq[1];
    g.doff = off[0] + off[1] * g.ds1 + off[2] * g.ds2;
    g.dptr = (double*)dst->ptr[peers[q]];
  }
  int rc;
  if ((rc = boxcopy(ctx, src, nullptr, L))) return rc;
  return k_barrier(ctx);
}

extern "C" int cales_transpose(cales_ctx* ctx, int which, const double* src, double* dst) {
  CHECK_CTX(ctx);
  // a destination obtained from cales_peer_alloc takes the solver's own exchange (NVLink stores into the owners' pencils)
  for (auto& kv : ctx->peerbufs)
    if (kv.second.local == (void*)dst && ctx->nranks > 1) {
      const bool colcomm = (which == 0 || which == 3);
      if ((colcomm ? ctx->dims[0] : ctx->dims[1]) == 1) break;
      int rc = k_barrier(ctx);                       // every rank is done reading the destination of the previous call
      return rc ? rc : k_transpose_p2p(ctx, which, src, &kv.second);
    }
  return k_transpose(ctx, which, src, dst);
}

extern "C" int cales_peer_alloc(cales_ctx* ctx, const char* name, long bytes, void** ptr) {
  CHECK_CTX(ctx);
  if (!name || !ptr || bytes <= 0) return cales_fail(ctx, CALES_ERR_INVALID, "peer_alloc: bad arguments");
  *ptr = nullptr;
  PeerBuf* pb = k_peer_buffer(ctx, (std::string("user_") + name).c_str(), (size_t)bytes);
  if (pb) { *ptr = pb->local; return CALES_OK; }
  if (ctx->nranks > 1 && ctx->p2p != 0) return ctx->err[0] ? CALES_ERR_NOMEM : cales_fail(ctx, CALES_ERR_CUDA, "peer_alloc failed");
  *ptr = cales_scratch(ctx, (std::string("user_") + name).c_str(), (size_t)bytes, true);    // single rank / no peer access: plain device memory
  return *ptr ? CALES_OK : CALES_ERR_NOMEM;
}
