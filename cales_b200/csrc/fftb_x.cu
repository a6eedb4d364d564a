// x-lines (contiguous) instantiations of the register-blocked batched transforms, see fftb.cuh
#include "fftb.cuh"

int k_fftb_x(cales_ctx* ctx, int n, const FftBArgs& A, int kind, int backward) { return fftb_dispatch<1>(ctx, n, A, kind, backward); }
