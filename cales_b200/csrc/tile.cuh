// Plane staging for the z-marching stencil kernels (mom_k, strain_k): a CTA owns a TX x TY tile of (i,j)
// columns; whole (TX+2) x (TY+2) planes of NF haloed fields are copied global -> shared with cp.async
// (no register staging) into a ring of FOUR slots: planes k and k+1 in use, planes k+2 and k+3 in flight, so a
// plane has two compute iterations to arrive.  Slots are compile-time constants (the k loop is unrolled by four):
// every shared-memory access of the stencil is [one base register + immediate].
//
// V16: 16-byte copies.  The tile starts at i = 32*bx (even) and a row of the array holds n1+2 values, so with n1
// even and 16-byte aligned base pointers every pair (i, i+1) of a tile row is one aligned 16-byte chunk that lies
// entirely inside or outside the array.  Otherwise (V16 = false) the same code moves 8-byte chunks.
#pragma once
#include <cstdint>
#include "common.cuh"

#define TX 32
#define TY 8
#define PX (TX + 2)
#define PY (TY + 2)
#define PLANE (PX * PY)
#define TSLOTS 4

template <int S> struct Slot { static constexpr int v = S; };

template <int NF, bool V16, int NTH = TX * TY>     // NTH: threads of the CTA (blockDim = (TX, NTH / TX))
struct Stager {
  static constexpr int CE = V16 ? 2 : 1;            // elements per chunk
  static constexpr int CPR = PX / CE;               // chunks per tile row
  static constexpr int CPP = CPR * PY;              // chunks per field plane
  static constexpr int NCH = NF * CPP;              // chunks per slot
  static constexpr int NR = (NCH + NTH - 1) / NTH;
  static constexpr int SLOT_BYTES = NF * PLANE * 8;
  const double* src[NR];    // advancing source pointers (plane `knext`); nullptr = this thread has no chunk in round r
  unsigned dst[NR];         // shared-memory byte address of the chunk in slot 0
  long s2;
  int knext, klast;         // next plane to issue, last plane this CTA needs

  __device__ __forceinline__ Stager(const Dims& d, int i0, int j0, const double* f0, const double* f1, const double* f2,
                                    const double* f3, const double* smem, int kfirst, int klast_, const double* f4 = nullptr) {
    const int t = threadIdx.x + TX * threadIdx.y;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    s2 = d.s2; knext = kfirst; klast = klast_;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int q = t + r * NTH;
      const int f = q / CPP, rem = q - f * CPP;
      const int lj = rem / CPR, li = (rem - lj * CPR) * CE;
      const int i = i0 - 1 + li, j = j0 - 1 + lj;
      const double* fb = f == 0 ? f0 : f == 1 ? f1 : f == 2 ? f2 : f == 3 ? f3 : f4;
      const bool ok = q < NCH && i + CE - 1 <= d.n1 + 1 && j <= d.n2 + 1;
      src[r] = ok ? fb + i + d.s1 * j + d.s2 * (long)kfirst : nullptr;
      dst[r] = sbase + 8u * (unsigned)(f * PLANE + lj * PX + li);
    }
  }

  // stage plane `knext` into slot S (nothing if the CTA does not need it); always one commit group (COMMIT = false: the
  // caller adds copies of its own to the group and commits it with tile_commit())
  template <int S, bool COMMIT = true>
  __device__ __forceinline__ void issue() {
    if (knext <= klast) {
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (src[r]) {
          if (V16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst[r] + S * SLOT_BYTES), "l"(src[r]) : "memory");
          else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst[r] + S * SLOT_BYTES), "l"(src[r]) : "memory");
          src[r] += s2;
        }
    }
    ++knext;
    if (COMMIT) asm volatile("cp.async.commit_group;" ::: "memory");
  }
};

__device__ __forceinline__ void tile_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tile_cp8(const double* smem_dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}

// all but the most recent commit group have landed
__device__ __forceinline__ void tile_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

static inline bool tile_v16(int n1, const void* a, const void* b, const void* c, const void* e = nullptr) {
  return n1 % 2 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)e) & 15) == 0;
}
