// Context, error handling, scratch and the pure-integer decomposition maps.
// Decomposition follows dependencies/2decomp-fft/src/decomp_2d.f90:1046-1149 (partition/distribute),
// decomp_2d_init_fin.f90:120-142 (rank = coord(1)*p_col + coord(2)) and src/initmpi.f90:141-204.
#include "common.cuh"

char g_cales_err[512] = {0};

int cales_fail(cales_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) strncpy(ctx->err, buf, sizeof ctx->err - 1);
  strncpy(g_cales_err, buf, sizeof g_cales_err - 1);
  return code;
}

void* cales_scratch(cales_ctx* ctx, const char* name, size_t bytes, bool zero_on_create) {
  auto it = ctx->scratch.find(name);
  if (it != ctx->scratch.end()) {
    if (it->second.second >= bytes) return it->second.first;
    cudaFree(it->second.first);
    ctx->scratch.erase(it);
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) {
    cales_fail(ctx, CALES_ERR_NOMEM, "cudaMalloc(%zu) for scratch '%s' failed", bytes, name);
    return nullptr;
  }
  if (zero_on_create) cudaMemsetAsync(p, 0, bytes, ctx->stream);
  ctx->scratch[name] = {p, bytes};
  return p;
}

extern "C" const char* cales_version(void) { return "cales_b200 0.1 (sm_100a)"; }

extern "C" const char* cales_last_error(const cales_ctx* ctx) { return ctx ? ctx->err : g_cales_err; }

// ---- integer maps ---------------------------------------------------------------------------------
extern "C" int cales_distribute(int data1, int proc, int* st, int* en, int* sz) {
  if (proc < 1 || data1 < 0) return CALES_ERR_INVALID;
  // decomp_2d.f90:1132-1145 (NEW_DISTRIBUTION): the first mod(data1,proc) ranks get one extra point
  const int q = data1 / proc, r = data1 % proc;
  for (int i = 0; i < proc; ++i) {
    sz[i] = q + (i < r ? 1 : 0);
    st[i] = 1 + i * q + (i < r ? i : r);
    en[i] = st[i] + sz[i] - 1;
  }
  return CALES_OK;
}

extern "C" int cales_pencil(const int ng[3], const int dims[2], int rank, int axis, int lo[3], int hi[3], int sz[3]) {
  if (axis < 1 || axis > 3 || dims[0] < 1 || dims[1] < 1 || rank < 0 || rank >= dims[0] * dims[1]) return CALES_ERR_INVALID;
  static const int pdim[3][3] = {{1, 2, 3}, {2, 1, 3}, {2, 3, 1}};  // x-, y-, z-pencil (decomp_2d.f90 get_decomp_info)
  const int coord[2] = {rank / dims[1], rank % dims[1]};
  for (int i = 0; i < 3; ++i) {
    const int pd = pdim[axis - 1][i];
    if (pd == 1) {
      lo[i] = 1; hi[i] = ng[i]; sz[i] = ng[i];
    } else {
      const int a = pd - 2;
      std::vector<int> st(dims[a]), en(dims[a]), s(dims[a]);
      cales_distribute(ng[i], dims[a], st.data(), en.data(), s.data());
      lo[i] = st[coord[a]]; hi[i] = en[coord[a]]; sz[i] = s[coord[a]];
    }
  }
  return CALES_OK;
}

extern "C" int cales_neighbours(const int dims[2], int ipencil, const char cbcpre[6], int rank, int nb[6], int is_bound[6]) {
  if (ipencil < 1 || ipencil > 3) return CALES_ERR_INVALID;
  int ipt[2], c = 0;
  for (int d = 1; d <= 3; ++d) if (d != ipencil) ipt[c++] = d;   // initmpi.f90:63
  const int coord[2] = {rank / dims[1], rank % dims[1]};
  for (int i = 0; i < 6; ++i) nb[i] = -1;                        // nb(:,ipencil) = MPI_PROC_NULL
  for (int a = 0; a < 2; ++a) {                                  // MPI_CART_SHIFT(comm_cart,a,1,...)
    const int idir = ipt[a] - 1;
    const bool periodic = cbcpre[tb(0, idir)] == 'P' && cbcpre[tb(1, idir)] == 'P';
    for (int ib = 0; ib < 2; ++ib) {
      int cc[2] = {coord[0], coord[1]};
      cc[a] += ib == 0 ? -1 : 1;
      if (cc[a] < 0 || cc[a] >= dims[a]) {
        if (!periodic) continue;
        cc[a] = (cc[a] + dims[a]) % dims[a];
      }
      nb[tb(ib, idir)] = cc[0] * dims[1] + cc[1];
    }
  }
  for (int i = 0; i < 6; ++i) is_bound[i] = nb[i] == -1;         // initmpi.f90:204
  return CALES_OK;
}

int comm_init(cales_ctx* ctx, const char* uid);   // comm.cu
void comm_finalize(cales_ctx* ctx);

extern "C" int cales_init(cales_ctx** out, const int ng[3], const int dims[2], int ipencil, const char cbcpre[6],
                          int rank, int nranks, const char* nccl_uid, int device, void* stream, int diffusion) {
  if (!out) return CALES_ERR_INVALID;
  *out = nullptr;
  if (dims[0] * dims[1] != nranks) return cales_fail(nullptr, CALES_ERR_INVALID, "dims(1)*dims(2)=%d != nranks=%d", dims[0] * dims[1], nranks);
  if (ipencil < 1 || ipencil > 3) return cales_fail(nullptr, CALES_ERR_INVALID, "ipencil must be 1,2 or 3");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return cales_fail(nullptr, CALES_ERR_CUDA, "no CUDA device available (%s): libcales_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0) device = rank % ndev;                          // initmpi.f90:85-87
  if ((e = cudaSetDevice(device)) != cudaSuccess) return cales_fail(nullptr, CALES_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  cales_ctx* ctx = new cales_ctx();
  ctx->device = device;
  ctx->stream = (cudaStream_t)stream;
  ctx->diffusion = diffusion;
  ctx->ipencil = ipencil; ctx->rank = rank; ctx->nranks = nranks;
  ctx->coord[0] = rank / dims[1]; ctx->coord[1] = rank % dims[1];
  memcpy(ctx->ng, ng, sizeof ctx->ng); memcpy(ctx->dims, dims, sizeof ctx->dims); memcpy(ctx->cbcpre, cbcpre, 6);
  cales_pencil(ng, dims, rank, 1, ctx->xst, ctx->xen, ctx->xsz);
  cales_pencil(ng, dims, rank, 2, ctx->yst, ctx->yen, ctx->ysz);
  cales_pencil(ng, dims, rank, 3, ctx->zst, ctx->zen, ctx->zsz);
  const int* st = ipencil == 1 ? ctx->xst : ipencil == 2 ? ctx->yst : ctx->zst;
  const int* en = ipencil == 1 ? ctx->xen : ipencil == 2 ? ctx->yen : ctx->zen;
  for (int i = 0; i < 3; ++i) { ctx->lo[i] = st[i]; ctx->hi[i] = en[i]; ctx->n[i] = en[i] - st[i] + 1; }
  cales_neighbours(dims, ipencil, cbcpre, rank, ctx->nb, ctx->is_bound);
  if (cudaMalloc(&ctx->red, 4096 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&ctx->red_host, 4096 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&ctx->fdev, 8 * sizeof(double)) != cudaSuccess) {
    int rc = cales_fail(nullptr, CALES_ERR_NOMEM, "context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return rc;
  }
  cudaMemsetAsync(ctx->fdev, 0, 8 * sizeof(double), ctx->stream);
  if (nranks > 1) {
    int rc = comm_init(ctx, nccl_uid);
    if (rc) { strncpy(g_cales_err, ctx->err, sizeof g_cales_err - 1); delete ctx; return rc; }
  }
  *out = ctx;
  return CALES_OK;
}

void k_gaussel_tab_free(cales_ctx* ctx);
void k_zdist_free(cales_ctx* ctx);
void k_step_graphs_free(cales_ctx* ctx);

extern "C" int cales_finalize(cales_ctx* ctx) {
  CHECK_CTX(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_finalize(ctx);
  for (auto& st_ : ctx->side) if (st_) cudaStreamDestroy(st_);
  for (auto& e_ : ctx->side_ev) if (e_) cudaEventDestroy(e_);
  k_step_graphs_free(ctx);
  k_gaussel_tab_free(ctx);
  k_zdist_free(ctx);
  for (auto& kv : ctx->scratch) cudaFree(kv.second.first);
  for (auto& kv : ctx->tables) { cudaFree(kv.second.w); cudaFree(kv.second.h); }
  cudaFree(ctx->red); cudaFreeHost(ctx->red_host); cudaFree(ctx->fdev); cudaFree(ctx->bar); cudaFree(ctx->halo_counter);
  delete ctx;
  return CALES_OK;
}

extern "C" int cales_get_decomp(const cales_ctx* ctx, int lo[3], int hi[3], int n[3], int n_x_fft[3], int n_y_fft[3],
                                int lo_z[3], int hi_z[3], int n_z[3], int nb[6], int is_bound[6]) {
  CHECK_CTX(ctx);
  for (int i = 0; i < 3; ++i) {
    lo[i] = ctx->lo[i]; hi[i] = ctx->hi[i]; n[i] = ctx->n[i];
    n_x_fft[i] = ctx->xsz[i]; n_y_fft[i] = ctx->ysz[i];
    lo_z[i] = ctx->zst[i]; hi_z[i] = ctx->zen[i]; n_z[i] = ctx->zsz[i];
  }
  for (int i = 0; i < 6; ++i) { nb[i] = ctx->nb[i]; is_bound[i] = ctx->is_bound[i]; }
  return CALES_OK;
}

extern "C" int cales_stream_synchronize(cales_ctx* ctx) {
  CHECK_CTX(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return CALES_OK;
}

extern "C" long cales_launch_count(const cales_ctx* ctx) { return ctx ? ctx->launches : 0; }
