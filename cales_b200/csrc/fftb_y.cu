// y-lines (strided; lanes across neighbouring x) instantiations of the register-blocked batched transforms, see fftb.cuh
#include <cstdlib>

#include "fftb.cuh"

int k_fftb_x(cales_ctx* ctx, int n, const FftBArgs& A, int kind, int backward);
int k_fftb_y(cales_ctx* ctx, int n, const FftBArgs& A, int kind, int backward) { return fftb_dispatch<0>(ctx, n, A, kind, backward); }

bool k_fftb_supported(int n) { return n >= 32 && n <= 1024 && !(n & (n - 1)) && getenv("CALES_FFT_GENERIC") == nullptr; }

// returns 1 if handled, 0 if this length is not covered by the fast path (caller falls back), <0 on error

// ds (forward x pass only): the line elements are computed from the velocity arrays (fused fillps, see FftBArgs); `in` then
// only defines the strides and ds->u/v/w point at element (1,1,1) like `in` would
int k_fftb_pass(cales_ctx* ctx, int dir, int kind, int backward, int n, int nl1, int nl2, const double* in, long ies, long il1, long il2,
                double* out, long oes, long ol1, long ol2, double scale, const FftTables* T, const DivSrc* ds) {
  if (n < 32 || n > 1024 || (n & (n - 1))) return 0;
  if (dir == 0 && (ies != 1 || oes != 1)) return 0;
  const int m = n / 2;
  FftBArgs A;
  A.in = in; A.out = out; A.ies = ies; A.il1 = il1; A.il2 = il2; A.oes = oes; A.ol1 = ol1; A.ol2 = ol2;
  A.nl1 = nl1; A.nl2 = nl2; A.dd = kind == KB_DD; A.scale = scale;
  A.wm = T->w; A.wn = T->w + m; A.h4 = T->h;
  A.np = 0; A.zoff = 0; A.nx = 0;
  A.su = A.sv = A.sw = A.sdzfi = A.rbx = A.rby = A.rbz = nullptr;
  if (ds) {
    if (dir != 0 || backward) return -cales_fail(ctx, CALES_ERR_INVALID, "fused fillps: forward x pass only");
    A.su = ds->u; A.sv = ds->v; A.sw = ds->w; A.sdzfi = ds->dzfi; A.rbx = ds->rbx; A.rby = ds->rby; A.rbz = ds->rbz;
    A.dti = ds->dti; A.dtidxi = ds->dti * ds->dxi; A.dtidyi = ds->dti * ds->dyi;      // as fillps_k forms them
    for (int q = 0; q < 6; ++q) A.bnd[q] = ds->bnd[q] && (q < 2 ? ds->rbx : q < 4 ? ds->rby : ds->rbz) != nullptr;
    A.sn1 = n; A.sn2 = nl1; A.sn3 = nl2;
  }
  if (ctx->fft_peer_out && dir == 1 && !backward) {
    const FftPeerOut& P = *ctx->fft_peer_out;
    A.np = P.np; A.zoff = P.zoff; A.nx = P.nx;
    for (int q = 0; q < P.np; ++q) { A.pbase[q] = P.pbase[q]; A.pys[q] = P.pys[q]; A.pny[q] = P.pny[q]; }
    A.pys[P.np] = P.pys[P.np];
  }
  return dir == 0 ? k_fftb_x(ctx, n, A, kind, backward) : k_fftb_y(ctx, n, A, kind, backward);
}
