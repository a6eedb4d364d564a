// Communication: NCCL bootstrap, halo exchange, scalar/plane reductions.
//   updthalo      src/bound.f90:619-696  (MPI_SENDRECV of full-extent faces, dimension by dimension)
//   updthalo_gpu  src/bound.f90:698-723  (cuDecomp halos; periodic single-rank self-copy halo.h:143-171)
//   NCCL bootstrap as dependencies/cuDecomp/src/cudecomp.cc:66-80 (unique id broadcast by the host)
// One process per GPU; every collective is enqueued on the context stream, so the caller's program
// order is kept without host synchronisation.  All fields of one call share a single packed
// message per neighbour (the reference sends one message per field and direction).
#include <nccl.h>

#include "common.cuh"

#define NCCL_TRY(ctx, call)                                                                              \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) return cales_fail(ctx, CALES_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
  } while (0)

extern "C" int cales_get_unique_id(char uid[CALES_UNIQUE_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= CALES_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return cales_fail(nullptr, CALES_ERR_NCCL, "ncclGetUniqueId failed");
  memset(uid, 0, CALES_UNIQUE_ID_BYTES);
  memcpy(uid, &id, sizeof id);
  return CALES_OK;
}

int comm_init(cales_ctx* ctx, const char* uid) {
  if (!uid) return cales_fail(ctx, CALES_ERR_INVALID, "nranks>1 requires an NCCL unique id (cales_get_unique_id on rank 0, broadcast by the host)");
  ncclUniqueId id;
  memcpy(&id, uid, sizeof id);
  ncclComm_t comm;
  NCCL_TRY(ctx, ncclCommInitRank(&comm, ctx->nranks, id, ctx->rank));
  ctx->nccl = comm;
  return CALES_OK;
}

void comm_finalize(cales_ctx* ctx) {
  if (ctx->nccl) { ncclCommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
}

int k_allreduce_sum(cales_ctx* ctx, double* dev, int count) {
  if (ctx->nranks == 1) return CALES_OK;
  NCCL_TRY(ctx, ncclAllReduce(dev, dev, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return CALES_OK;
}

int k_allreduce_minmax(cales_ctx* ctx, double* dev, int count, int is_max) {
  if (ctx->nranks == 1) return CALES_OK;
  NCCL_TRY(ctx, ncclAllReduce(dev, dev, count, ncclDouble, is_max ? ncclMax : ncclMin, (ncclComm_t)ctx->nccl, ctx->stream));
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

int k_halo_exchange(cales_ctx* ctx, const int n[3], const int nb[6], double* const* fields, int nfields) {
  Dims d(n);
  for (int f0 = 0; f0 < nfields; f0 += 12) {
    FieldList fl;
    fl.nf = nfields - f0 < 12 ? nfields - f0 : 12;
    for (int f = 0; f < fl.nf; ++f) fl.p[f] = fields[f0 + f];
    for (int idir = 0; idir < 3; ++idir) {
      if (idir + 1 == ctx->ipencil) continue;                 // bound.f90:634
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
      double* sbuf = (double*)cales_scratch(ctx, "halo_send", (size_t)2 * 12 * sizeof(double) * (size_t)fmax_);
      double* rbuf = (double*)cales_scratch(ctx, "halo_recv", ctx->scratch["halo_send"].second);
      if (!sbuf || !rbuf) return CALES_ERR_NOMEM;
      halo_pack_k<<<g, b, 0, ctx->stream>>>(d, idir, fl, sbuf);
      KERNEL_CHECK(ctx);
      ncclComm_t comm = (ncclComm_t)ctx->nccl;
      NCCL_TRY(ctx, ncclGroupStart());
      // sends: my plane 1 to nb(0), my plane n to nb(1); receives in the matching order for the
      // two-rank periodic case (the peer's first message is its plane 1 = my upper ghost)
      if (nb0 >= 0) NCCL_TRY(ctx, ncclSend(sbuf, cnt, ncclDouble, nb0, comm, ctx->stream));
      if (nb1 >= 0) NCCL_TRY(ctx, ncclSend(sbuf + cnt, cnt, ncclDouble, nb1, comm, ctx->stream));
      if (nb1 >= 0) NCCL_TRY(ctx, ncclRecv(rbuf + cnt, cnt, ncclDouble, nb1, comm, ctx->stream));
      if (nb0 >= 0) NCCL_TRY(ctx, ncclRecv(rbuf, cnt, ncclDouble, nb0, comm, ctx->stream));
      NCCL_TRY(ctx, ncclGroupEnd());
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

// ---- pencil transposes (placeholder until the NCCL all-to-all path below is wired) -----------------------
int k_transpose(cales_ctx* ctx, int which, const double* src, double* dst) {
  (void)which; (void)src; (void)dst;
  return cales_fail(ctx, CALES_ERR_INVALID, "distributed transposes are not implemented yet");
}

extern "C" int cales_transpose(cales_ctx* ctx, int which, const double* src, double* dst) {
  CHECK_CTX(ctx);
  return k_transpose(ctx, which, src, dst);
}
