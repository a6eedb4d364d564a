// Fused entry points of the time loop (last row): one RK3 substep = src/main.f90:418-506 in ONE call, and
// one time step = three of them, optionally replayed from a CUDA graph.
//
// What the fusion buys (explicit diffusion; results identical to the per-procedure sequence, bit for bit in the strict build):
//   * rk: the low-storage update is applied inside the momentum kernel (mom_k<.., RK>), the new velocity going to a set of
//     intermediate arrays (us,vs,ws) owned by the library -- 112 instead of 160 B/cell;
//   * bulk forcing, ghost fill, wall model and fillps work on the intermediate velocity;
//   * correc reads the intermediate velocity and writes the CALLER's u,v,w (out of place: no extra traffic), with the
//     explicit pressure update p += pp of the interior in the same pass -- 72 instead of 80 B/cell, one launch less;
//   * cales_step: on one rank the ~200 launches of a step are captured once per (dt, RK history parity) and replayed as one
//     graph launch -- what matters for the launch-bound small grids (BASELINE config 1, 64^3).
#include <cstddef>

#include "common.cuh"

int k_rk_dev(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci, const double* dzfi,
             const double* gvr_c, const double* gvr_f, double visc, double dt, const double* p, const int is_forced[3],
             const double velf[3], const double bforce[3], const double* visct, double* u, double* v, double* w,
             double* un, double* vn, double* wn);
int k_bulk_forcing_dev(cales_ctx* ctx, const int n[3], const int is_forced[3], const double* fdev, double* u, double* v, double* w);
int k_correc_updatep(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, double dt, const double* pp,
                     const double* us, const double* vs, const double* ws, double* u, double* v, double* w, double* p);

static const double RKCOEFF[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};   // param.f90:27-29

extern "C" int cales_substep(cales_ctx* ctx, const cales_step_args* a, int irk, double dt) {
  CHECK_CTX(ctx);
  if (!a || irk < 1 || irk > 3) return cales_fail(ctx, CALES_ERR_INVALID, "substep: irk must be 1, 2 or 3");
  if (ctx->diffusion != CALES_DIFF_EXPLICIT)
    return cales_fail(ctx, CALES_ERR_INVALID, "substep: the fused entry covers explicit diffusion; drive _IMPDIFF runs through the per-procedure entries");
  const int* n = a->n;
  const size_t fb = (size_t)(n[0] + 2) * (n[1] + 2) * (n[2] + 2) * sizeof(double);
  double* us = (double*)cales_scratch(ctx, "step_us", fb, true);
  double* vs = (double*)cales_scratch(ctx, "step_vs", fb, true);
  double* ws = (double*)cales_scratch(ctx, "step_ws", fb, true);
  if (!us || !vs || !ws) return CALES_ERR_NOMEM;
  const double* rk = RKCOEFF[irk - 1];
  const double dtrk = (rk[0] + rk[1]) * dt, dtrki = 1. / dtrk;
  int rc;
  // rk + bulk forcing (main.f90:420-422)
  if ((rc = k_rk_dev(ctx, rk, n, a->dli, a->dzci, a->dzfi, a->grid_vol_ratio_c, a->grid_vol_ratio_f, a->visc, dt, a->p, a->is_forced, a->velf,
                     a->bforce, a->visct, a->u, a->v, a->w, us, vs, ws))) return rc;
  if ((rc = k_bulk_forcing_dev(ctx, n, a->is_forced, ctx->fdev, us, vs, ws))) return rc;
  // main.f90:493-497
  if ((rc = cales_bounduvw(ctx, a->cbcvel, n, &a->bcu, &a->bcv, &a->bcw, &a->bcu_mag, &a->bcv_mag, &a->bcw_mag, a->nb, a->is_bound, a->lwm, a->l, a->dl,
                           a->zc, a->zf, a->dzc, a->dzf, a->visc, a->hwm, a->index_wm, 1, 0, us, vs, ws))) return rc;
  // fillps + updt_rhs_b + solver (main.f90:495-497): the first two are handed to the solver, whose forward x pass forms the
  // right-hand side on the fly where it can (solver.cu) and runs the two kernels itself where it cannot
  {
    DivSrc S;
    S.u = us; S.v = vs; S.w = ws; S.dzfi = a->dzfi; S.rbx = a->rhsbx; S.rby = a->rhsby; S.rbz = a->rhsbz;
    S.dti = dtrki; S.dxi = a->dli[0]; S.dyi = a->dli[1];
    for (int q6 = 0; q6 < 6; ++q6) S.bnd[q6] = a->is_bound[q6];
    ctx->div_src = &S;
    rc = cales_solver(ctx, n, a->ng, a->plan, a->normfft, a->lambdaxy, a->a, a->b, a->c, a->cbcpre, "ccc", a->pp);
    ctx->div_src = nullptr;
    if (rc) return rc;
  }
  // main.f90:498-503
  if ((rc = cales_boundp(ctx, a->cbcpre, n, &a->bcp, a->nb, a->is_bound, a->dl, a->dzc, a->pp))) return rc;
  if ((rc = k_correc_updatep(ctx, n, a->dli, a->dzci, dtrk, a->pp, us, vs, ws, a->u, a->v, a->w, a->p))) return rc;
  if ((rc = cales_bounduvw(ctx, a->cbcvel, n, &a->bcu, &a->bcv, &a->bcw, &a->bcu_mag, &a->bcv_mag, &a->bcw_mag, a->nb, a->is_bound, a->lwm, a->l, a->dl,
                           a->zc, a->zf, a->dzc, a->dzf, a->visc, a->hwm, a->index_wm, 1, 1, a->u, a->v, a->w))) return rc;
  if ((rc = cales_boundp(ctx, a->cbcpre, n, &a->bcp, a->nb, a->is_bound, a->dl, a->dzc, a->p))) return rc;
  // main.f90:504-506
  if ((rc = cales_cmpt_sgs(ctx, a->sgstype, n, a->ng, a->lo, a->hi, a->cbcvel, a->cbcsgs, &a->bcs, a->nb, a->is_bound, a->lwm, a->l, a->dl, a->dli, a->zc,
                           a->zf, a->dzc, a->dzf, a->dzci, a->dzfi, a->visc, a->hwm, a->index_wm, a->u, a->v, a->w, &a->bcuf, &a->bcvf, &a->bcwf,
                           &a->bcu_mag, &a->bcv_mag, &a->bcw_mag, a->visct))) return rc;
  return cales_boundp(ctx, a->cbcsgs, n, &a->bcs, a->nb, a->is_bound, a->dl, a->dzc, a->visct);
}

// layout of cales_step_args as this library was compiled (a foreign-language binding checks its own against it)
extern "C" int cales_step_args_layout(long out[4]) {
  out[0] = (long)sizeof(cales_step_args);
  out[1] = (long)offsetof(cales_step_args, cbcvel);
  out[2] = (long)offsetof(cales_step_args, zc);
  out[3] = (long)offsetof(cales_step_args, visct);
  return CALES_OK;
}

// ---- one time step, optionally as a CUDA graph ---------------------------------------------------------------------------
// The captured work depends on dt (kernel arguments), on which RK history set is "old" at the start of the step (it flips
// every substep, hence every step) and on the argument block; graphs are cached under that key.  The first two steps of a
// context run eagerly (lazy scratch allocations are not capturable); afterwards a new key is captured and launched at once.  Single rank only: the peer-memory barrier and halo kernels carry sequence numbers as arguments.
struct StepGraph { double dt; int swap; unsigned long long hash; long nlaunch; cudaGraphExec_t exec; };
// the cache lives on the context (cales_ctx::step_graphs): contexts of different host threads share nothing
static std::vector<StepGraph>& step_graphs(cales_ctx* ctx) {
  if (!ctx->step_graphs) ctx->step_graphs = new std::vector<StepGraph>();
  return *(std::vector<StepGraph>*)ctx->step_graphs;
}

static unsigned long long hash_args(const cales_step_args* a) {
  const unsigned char* p = (const unsigned char*)a;
  unsigned long long h = 1469598103934665603ull;
  for (size_t q = 0; q < sizeof(*a); ++q) { h ^= p[q]; h *= 1099511628211ull; }
  return h;
}

void k_step_graphs_free(cales_ctx* ctx) {
  std::vector<StepGraph>* v = (std::vector<StepGraph>*)ctx->step_graphs;
  if (!v) return;
  for (auto& g : *v) if (g.exec) cudaGraphExecDestroy(g.exec);
  delete v;
  ctx->step_graphs = nullptr;
}

extern "C" int cales_step(cales_ctx* ctx, const cales_step_args* a, double dt, int use_graph) {
  CHECK_CTX(ctx);
  int rc;
  if (!use_graph || ctx->nranks > 1 || !ctx->stream) {                  // (the legacy default stream cannot be captured)
    for (int irk = 1; irk <= 3; ++irk) if ((rc = cales_substep(ctx, a, irk, dt))) return rc;
    return CALES_OK;
  }
  std::vector<StepGraph>& v = step_graphs(ctx);
  const unsigned long long h = hash_args(a);
  StepGraph* g = nullptr;
  for (auto& e : v) if (e.dt == dt && e.swap == ctx->rk_swap && e.hash == h) { g = &e; break; }
  if (!g) {
    if (v.size() >= 16) { for (auto& e : v) if (e.exec) cudaGraphExecDestroy(e.exec); v.clear(); }   // dt changes every icheck steps: bounded cache
    v.push_back(StepGraph{dt, ctx->rk_swap, h, 0, nullptr});
    g = &v.back();
  }
  if (g->exec) {
    CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->rk_swap ^= 1;                                                // three substeps flip the history parity once
    ctx->launches += g->nlaunch;                                      // the kernel nodes of the replayed graph
    return CALES_OK;
  }
  if (ctx->step_calls++ < 2 || ctx->rk_first) {      // the first two steps (one per history parity) run eagerly
    for (int irk = 1; irk <= 3; ++irk) if ((rc = cales_substep(ctx, a, irk, dt))) return rc;
    return CALES_OK;
  }
  const long l0 = ctx->launches;
  CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  rc = CALES_OK;
  for (int irk = 1; irk <= 3 && !rc; ++irk) rc = cales_substep(ctx, a, irk, dt);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  g->nlaunch = ctx->launches - l0;
  ctx->launches = l0;
  if (rc || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc ? rc : cales_fail(ctx, CALES_ERR_CUDA, "step: graph capture failed: %s", cudaGetErrorString(e));
  }
  e = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { g->exec = nullptr; return cales_fail(ctx, CALES_ERR_CUDA, "step: cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
  // the capture advanced the host-side state (rk_swap flipped three times = once) without running anything: launch it now
  CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
  ctx->launches += g->nlaunch;
  return CALES_OK;
}
