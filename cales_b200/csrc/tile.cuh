// Plane staging for the z-marching stencil kernels (mom_k, strain_k): a CTA owns a TX x TY tile of (i,j)
// columns; whole (TX+2) x (TY+2) planes of NF haloed fields are copied global -> shared with cp.async
// (no register staging), into a ring of three slots: planes k and k+1 in use, plane k+2 in flight.
#pragma once
#include "common.cuh"

#define TX 32
#define TY 8
#define PX (TX + 2)
#define PY (TY + 2)
#define PLANE (PX * PY)

// Every thread moves the same (at most two) tile points of each plane, so the tile-local index and the
// global offset are computed once, not per plane.
struct Stage {
  long g0, g1;      // global offsets (without the k term) of my two tile points, -1 if outside the array
  int q0, q1;       // their positions in the PX x PY tile
};

__device__ __forceinline__ Stage make_stage(const Dims& d, int i0, int j0) {
  Stage st;
  const int t = threadIdx.x + TX * threadIdx.y;
  st.q0 = t;                                   // PLANE = 340 > 256 = TX*TY: first point always exists
  st.q1 = t + TX * TY;
  {
    const int li = st.q0 % PX, lj = st.q0 / PX;
    const int i = i0 + li - 1, j = j0 + lj - 1;
    st.g0 = (i <= d.n1 + 1 && j <= d.n2 + 1) ? (long)i + d.s1 * j : -1;
  }
  st.g1 = -1;
  if (st.q1 < PLANE) {
    const int li = st.q1 % PX, lj = st.q1 / PX;
    const int i = i0 + li - 1, j = j0 + lj - 1;
    st.g1 = (i <= d.n1 + 1 && j <= d.n2 + 1) ? (long)i + d.s1 * j : -1;
  }
  return st;
}

__device__ __forceinline__ void tile_cp8(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}

// stage plane k of NF fields into slot `slot` of smem laid out [3 slots][NF fields][PLANE]; one commit group
template <int NF>
__device__ __forceinline__ void tile_issue(const Stage& st, const Dims& d, const double* const (&fld)[NF], double* smem, int k, int slot) {
  const long ko = d.s2 * (long)k;
  double* dst = smem + slot * (NF * PLANE);
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    if (st.g0 >= 0) tile_cp8(dst + f * PLANE + st.q0, fld[f] + st.g0 + ko);
    if (st.g1 >= 0) tile_cp8(dst + f * PLANE + st.q1, fld[f] + st.g1 + ko);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void tile_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
