// On-the-fly statistics of the channel: the 27 plane-averaged single-point profiles of
//   out1d_single_point_chan  src/output.f90:509-691  (idir = 3; called every iout1d steps through out1d.h90:35-36)
// as device reductions: one pass over u,v,w,p,visct (40 B/cell), per-tile partial sums folded in a fixed order
// (deterministic), rank sums by one ncclAllReduce of 27 x ng(3) doubles, one small copy to the host.  The second block of
// that routine (38 budget terms, output.f90:692-920) is not covered.
#include "common.cuh"
#include "reduce.cuh"

#define SBX 64
#define SBY 4
#define NV 27

__global__ void __launch_bounds__(SBX* SBY) out1d_part_k(Dims d, double dl1, double dl2, const double* __restrict__ dzc, const double* __restrict__ dzf,
                                                          const double* __restrict__ u, const double* __restrict__ v, const double* __restrict__ w,
                                                          const double* __restrict__ p, const double* __restrict__ s, double* __restrict__ part) {
  const int i = blockIdx.x * SBX + threadIdx.x + 1, j = blockIdx.y * SBY + threadIdx.y + 1, k = blockIdx.z + 1;
  double b[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) b[q] = 0.;
  if (i <= d.n1 && j <= d.n2) {
    const long c = d.idx(i, j, k), s1 = d.s1, s2 = d.s2;
    const double uc = u[c], vc = v[c], wc = w[c], pc = p[c];
    const double u_kp = u[c + s2], u_ip = u[c + 1], u_im = u[c - 1], u_jp = u[c + s1];
    const double v_kp = v[c + s2], v_ip = v[c + 1], v_jp = v[c + s1], v_jm = v[c - s1];
    const double w_ip = w[c + 1], w_jp = w[c + s1], w_kp = w[c + s2], w_km = w[c - s2];
    const double dzck = dzc[k], dzfk = dzf[k], dzfkp = dzf[k + 1];
    b[0] = uc; b[1] = vc; b[2] = wc;
    b[3] = uc * uc; b[4] = vc * vc; b[5] = wc * wc;
    b[6] = 0.25 * (u_kp + uc) * (wc + w_ip);                                         // cell edge
    b[7] = uc * uc * uc; b[8] = vc * vc * vc; b[9] = wc * wc * wc;
    b[10] = (uc * uc) * (uc * uc); b[11] = (vc * vc) * (vc * vc); b[12] = (wc * wc) * (wc * wc);
    b[13] = pc; b[14] = pc * pc;
    const double tx = (w_jp - wc) / dl2 - (v_kp - vc) / dzck;                        // vorticity
    const double ty = (u_kp - uc) / dzck - (w_ip - wc) / dl1;
    const double tz = (v_ip - vc) / dl1 - (u_jp - uc) / dl2;
    b[15] = tx; b[16] = ty; b[17] = tz; b[18] = tx * tx; b[19] = ty * ty; b[20] = tz * tz;
    const double s_ccc = s[c], s_pcc = s[c + 1], s_cpc = s[c + s1], s_ccp = s[c + s2], s_pcp = s[c + 1 + s2];
    const double dudx_ip = (u_ip - uc) / dl1, dudx_im = (uc - u_im) / dl1;
    const double dvdy_jp = (v_jp - vc) / dl2, dvdy_jm = (vc - v_jm) / dl2;
    const double dwdz_kp = (w_kp - wc) / dzfkp, dwdz_km = (wc - w_km) / dzfk;
    const double dudz = (u_kp - uc) / dzck, dwdx = (w_ip - wc) / dl1;
    b[21] = -0.5 * (s_pcc * (dudx_ip + dudx_ip) + s_ccc * (dudx_im + dudx_im));      // modelled stresses
    b[22] = -0.5 * (s_cpc * (dvdy_jp + dvdy_jp) + s_ccc * (dvdy_jm + dvdy_jm));
    b[23] = -0.5 * (s_ccp * (dwdz_kp + dwdz_kp) + s_ccc * (dwdz_km + dwdz_km));
    b[24] = -0.25 * (s_ccc + s_pcc + s_ccp + s_pcp) * (dudz + dwdx);
    b[25] = s_ccc;
    b[26] = dudz;
  }
  const int tile = blockIdx.x + gridDim.x * blockIdx.y, ntile = gridDim.x * gridDim.y;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const double t = block_sum<SBX * SBY>(b[q]);
    if (threadIdx.x == 0 && threadIdx.y == 0) part[((long)blockIdx.z * NV + q) * ntile + tile] = t;
  }
}

// buf(q, kglobal) = grid_area_ratio * sum over tiles (fixed order); one warp per (k, q)
__global__ void out1d_fold_k(int n3, int ntile, int koff, double gar, const double* __restrict__ part, double* __restrict__ buf) {
  const int k = blockIdx.x, q = threadIdx.y;
  double t = 0.;
  for (int m = threadIdx.x; m < ntile; m += 32) t = t + part[((long)k * NV + q) * ntile + m];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = t + __shfl_down_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) buf[q + NV * (long)(koff + k)] = t * gar;
  (void)n3;
}

extern "C" int cales_out1d_chan(cales_ctx* ctx, const int ng[3], const int lo[3], const int hi[3], const double l[3], const double dl[3],
                                const double* dzc, const double* dzf, const double* u, const double* v, const double* w, const double* p,
                                const double* visct, double* buf) {
  CHECK_CTX(ctx);
  const int n[3] = {hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
  Dims d(n);
  dim3 g(cdiv(n[0], SBX), cdiv(n[1], SBY), n[2]);
  const int ntile = g.x * g.y;
  double* part = (double*)cales_scratch(ctx, "out1d_part", (size_t)n[2] * NV * ntile * sizeof(double));
  double* dbuf = (double*)cales_scratch(ctx, "out1d_buf", (size_t)NV * ng[2] * sizeof(double));
  if (!part || !dbuf) return CALES_ERR_NOMEM;
  CUDA_TRY(ctx, cudaMemsetAsync(dbuf, 0, (size_t)NV * ng[2] * sizeof(double), ctx->stream));
  out1d_part_k<<<g, dim3(SBX, SBY), 0, ctx->stream>>>(d, dl[0], dl[1], dzc, dzf, u, v, w, p, visct, part);
  KERNEL_CHECK(ctx);
  out1d_fold_k<<<n[2], dim3(32, NV), 0, ctx->stream>>>(n[2], ntile, lo[2] - 1, dl[0] * dl[1] / (l[0] * l[1]), part, dbuf);
  KERNEL_CHECK(ctx);
  int rc = k_allreduce_sum(ctx, dbuf, NV * ng[2]);                              // output.f90:683
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(buf, dbuf, (size_t)NV * ng[2] * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));                            // `!$acc wait(1)`, output.f90:682
  return CALES_OK;
}
