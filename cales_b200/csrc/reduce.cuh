// Deterministic block reductions (fixed tree => results independent of scheduling).
#pragma once

template <int NT>
__device__ inline double block_sum(double v) {
  __shared__ double sh_[NT / 32];
  const int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((t & 31) == 0) sh_[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    v = t < NT / 32 ? sh_[t] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

template <int NT>
__device__ inline double block_max(double v) {
  __shared__ double shm_[NT / 32];
  const int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((t & 31) == 0) shm_[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    v = t < NT / 32 ? shm_[t] : -1.7976931348623157e308;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;
}

// OP 0: sum, 1: max, 2: min.  One block of 256 threads folds `n` partials into out[0].
template <int OP>
__global__ void final_reduce_k(const double* __restrict__ part, int n, double* __restrict__ out) {
  double v = OP == 0 ? 0. : (OP == 1 ? -1.7976931348623157e308 : 1.7976931348623157e308);
  for (int i = threadIdx.x; i < n; i += 256) {
    const double x = part[i];
    v = OP == 0 ? v + x : (OP == 1 ? fmax(v, x) : fmin(v, x));
  }
  if (OP == 0) v = block_sum<256>(v);
  else if (OP == 1) v = block_max<256>(v);
  else v = -block_max<256>(-v);
  if (threadIdx.x == 0) out[0] = v;
}
