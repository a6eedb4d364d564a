// Peer-write micro-benchmark (design input for the solver exchange): how fast can one GPU push a contiguous buffer into
// a peer's memory over NVLink with (a) plain 16-byte SM stores, (b) 4x16-byte unrolled SM stores, (c) TMA bulk copies
// staged through shared memory, (d) the copy engine -- one direction and both directions at once.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/p2pbench tools/p2pbench.cu && tools/p2pbench [MB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_plain(const double2* __restrict__ s, double2* __restrict__ d, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) d[i] = s[i];
}
template <int U>
__global__ void k_unroll(const double2* __restrict__ s, double2* __restrict__ d, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    double2 v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) v[q] = s[i + q * stride];
#pragma unroll
    for (int q = 0; q < U; ++q) d[i + q * stride] = v[q];
  }
  for (; i < n; i += stride) d[i] = s[i];
}
// the store pattern of a fused transform kernel: every warp instruction writes SEG-byte contiguous segments (W bytes per
// thread) that are `rowstride` bytes apart in the destination (rows of a pencil), reading the source linearly
template <typename T, int SEGT>   // SEGT threads per segment
__global__ void k_rows(const T* __restrict__ s, T* __restrict__ d, long n, long rowstride_elems, int rows) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const long seg = i / SEGT, lane = i - seg * SEGT;
    const long row = seg % rows, col = seg / rows;           // consecutive segments go to consecutive rows
    const long o = row * rowstride_elems + col * SEGT + lane;
    if (o < n) d[o] = s[i];
  }
}

// TMA: one elected thread per CTA moves CH-byte chunks global -> shared -> peer global, NST stages
template <int CH, int NST>
__global__ void __launch_bounds__(32) k_tma(const char* __restrict__ s, char* __restrict__ d, long bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar[NST];
  if (threadIdx.x != 0) return;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  for (int q = 0; q < NST; ++q) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[q]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const long nch = bytes / CH;
  long c = blockIdx.x;
  unsigned par[NST];
  for (int q = 0; q < NST; ++q) par[q] = 0;
  // prologue
  long issued = c; int st = 0;
  for (int q = 0; q < NST && issued < nch; ++q, issued += gridDim.x) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[q]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CH) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + q * CH), "l"(s + issued * CH), "r"(CH), "r"(b) : "memory");
  }
  for (; c < nch; c += gridDim.x) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[st]);
    unsigned ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(par[st]) : "memory");
    par[st] ^= 1u;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + c * CH), "r"(sbase + st * CH), "r"(CH) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (issued < nch) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stage's smem has been read by the store
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CH) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sbase + st * CH), "l"(s + issued * CH), "r"(CH), "r"(b) : "memory");
      issued += gridDim.x;
    }
    st = (st + 1) % NST;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const long mb = argc > 1 ? atol(argv[1]) : 64;
  const long bytes = mb << 20;
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
  char *src[2], *dst[2]; cudaStream_t st[2]; cudaEvent_t e0[2], e1[2];
  for (int g = 0; g < 2; ++g) {
    CK(cudaSetDevice(g));
    CK(cudaDeviceEnablePeerAccess(1 - g, 0));
    CK(cudaMalloc(&src[g], bytes)); CK(cudaMalloc(&dst[g], bytes));
    CK(cudaMemset(src[g], g + 1, bytes)); CK(cudaMemset(dst[g], 0, bytes));
    CK(cudaStreamCreate(&st[g])); CK(cudaEventCreate(&e0[g])); CK(cudaEventCreate(&e1[g]));
  }
  const long n2 = bytes / 16;
  for (int g = 0; g < 2; ++g) {
    CK(cudaSetDevice(g));
    CK(cudaFuncSetAttribute(k_tma<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
    CK(cudaFuncSetAttribute(k_tma<32768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768));
  }
  auto run = [&](const char* name, int ndir, auto launch) {
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
      for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
      for (int g = 0; g < ndir; ++g) { CK(cudaSetDevice(g)); CK(cudaEventRecord(e0[g], st[g])); launch(g); CK(cudaEventRecord(e1[g], st[g])); }
      float worst = 0;
      for (int g = 0; g < ndir; ++g) { CK(cudaSetDevice(g)); CK(cudaEventSynchronize(e1[g])); float ms; CK(cudaEventElapsedTime(&ms, e0[g], e1[g])); if (ms > worst) worst = ms; }
      if (it > 0 && worst < best) best = worst;
    }
    printf("%-34s %s  %8.3f ms  %8.1f GB/s per direction\n", name, ndir == 1 ? "one way " : "both ways", best, bytes / best / 1e6);
    fflush(stdout);
  };
  for (int ndir = 1; ndir <= 2; ++ndir) {
    run("copy engine (cudaMemcpyPeerAsync)", ndir, [&](int g) { CK(cudaMemcpyPeerAsync(dst[1 - g], 1 - g, src[g], g, bytes, st[g])); });
    for (int grid : {148, 592})
      for (int bs : {512}) {
        char nm[64]; snprintf(nm, sizeof nm, "plain 16B  grid %d x %d", grid, bs);
        run(nm, ndir, [&](int g) { k_plain<<<grid, bs, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2); });
      }
    {
      // rows of 2 KB (nx = 256 doubles): segment sizes 64 B ... 512 B, 8 or 16 bytes per thread
      const int rows = 4096; const long n8 = bytes / 8; const long rs8 = n8 / rows, rs16 = n2 / rows;
      run("rows: 8B/thread, 64B segments", ndir, [&](int g) { k_rows<double, 8><<<592, 512, 0, st[g]>>>((const double*)src[g], (double*)dst[1 - g], n8, rs8, rows); });
      run("rows: 8B/thread, 128B segments", ndir, [&](int g) { k_rows<double, 16><<<592, 512, 0, st[g]>>>((const double*)src[g], (double*)dst[1 - g], n8, rs8, rows); });
      run("rows: 8B/thread, 256B segments", ndir, [&](int g) { k_rows<double, 32><<<592, 512, 0, st[g]>>>((const double*)src[g], (double*)dst[1 - g], n8, rs8, rows); });
      run("rows: 16B/thread, 128B segments", ndir, [&](int g) { k_rows<double2, 8><<<592, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2, rs16, rows); });
      run("rows: 16B/thread, 256B segments", ndir, [&](int g) { k_rows<double2, 16><<<592, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2, rs16, rows); });
      run("rows: 16B/thread, 512B segments", ndir, [&](int g) { k_rows<double2, 32><<<592, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2, rs16, rows); });
      run("rows: 8B/thread, 128B seg, 148x256", ndir, [&](int g) { k_rows<double, 16><<<148, 256, 0, st[g]>>>((const double*)src[g], (double*)dst[1 - g], n8, rs8, rows); });
      run("rows: 8B/thread, 128B seg, 1184x256", ndir, [&](int g) { k_rows<double, 16><<<1184, 256, 0, st[g]>>>((const double*)src[g], (double*)dst[1 - g], n8, rs8, rows); });
    }
    for (int grid : {592}) {
      char nm[64];
      snprintf(nm, sizeof nm, "unroll4 16B grid %d x 512", grid);
      run(nm, ndir, [&](int g) { k_unroll<4><<<grid, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2); });
      snprintf(nm, sizeof nm, "unroll8 16B grid %d x 512", grid);
      run(nm, ndir, [&](int g) { k_unroll<8><<<grid, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[1 - g], n2); });
    }
    for (int grid : {296}) {
      char nm[64];
      snprintf(nm, sizeof nm, "TMA 16KB x 4 stages grid %d", grid);
      run(nm, ndir, [&](int g) { k_tma<16384, 4><<<grid, 32, 4 * 16384, st[g]>>>(src[g], dst[1 - g], bytes); });
      snprintf(nm, sizeof nm, "TMA 32KB x 2 stages grid %d", grid);
      run(nm, ndir, [&](int g) { k_tma<32768, 2><<<grid, 32, 2 * 32768, st[g]>>>(src[g], dst[1 - g], bytes); });
      snprintf(nm, sizeof nm, "TMA 8KB x 4 stages grid %d", grid);
      run(nm, ndir, [&](int g) { k_tma<8192, 4><<<grid, 32, 4 * 8192, st[g]>>>(src[g], dst[1 - g], bytes); });
    }
  }
  // local copy for scale
  run("local copy unroll4 (HBM)", 1, [&](int g) { k_unroll<4><<<592, 512, 0, st[g]>>>((const double2*)src[g], (double2*)dst[g], n2); });
  return 0;
}
