/*
 * cales_b200.h -- C ABI of libcales_b200.so: the B200 (sm_100a) implementation of the CaLES
 * per-RK3-substep hot path.
 *
 * The reference has no FFI: its boundary is the set of Fortran module procedures called from
 * src/main.f90:144-507.  Every export below names the procedure it replaces (file:line in the
 * reference tree) and keeps that procedure's argument list and meaning, so that a bind(C)
 * interface module lets main.f90 call it unchanged.
 *
 * Conventions
 *  - all functions return 0 on success, a CALES_ERR_* code otherwise; nothing throws across the
 *    boundary; cales_last_error() returns a sticky message.
 *  - fp64 throughout (rp = dp, src/precision.f90:14-20).
 *  - 3-D fields are DEVICE pointers to Fortran-ordered arrays (0:n1+1,0:n2+1,0:n3+1) (one halo
 *    cell), i fastest: element (i,j,k) is at i + (n1+2)*(j + (n2+2)*k).
 *  - 1-D grid vectors dzc,dzf,zc,zf,dzci,dzfi,grid_vol_ratio_* are DEVICE pointers of extent
 *    0:n3+1 unless marked host.
 *  - small descriptor vectors (n, ng, lo, hi, dl, dli, l, nb, is_bound, lwm, index_wm, BC
 *    characters) are HOST pointers.  Fortran logicals are int (0/1).  (0:1,3) tables are stored
 *    in Fortran order: element (ib,idir) at ib + 2*(idir-1); cbcvel(0:1,3,3) element
 *    (ib,idir,ivel) at ib + 2*(idir-1) + 6*(ivel-1).
 *  - a `bound` (src/typedef.f90:10-14) is three DEVICE planes x(0:n2+1,0:n3+1,0:1),
 *    y(0:n1+1,0:n3+1,0:1), z(0:n1+1,0:n2+1,0:1).
 *  - every device operation is enqueued on the stream given to cales_init (the reference's
 *    OpenACC queue 1, src/workspaces.f90:69); functions that return a scalar to the host
 *    synchronise that stream, exactly where the reference does `!$acc wait(1)`.
 *  - there is no CPU fallback: a missing/failed device makes every call return an error.
 */
#ifndef CALES_B200_H
#define CALES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cales_ctx cales_ctx;

typedef struct cales_bound {
  double* x;
  double* y;
  double* z;
} cales_bound;

enum {
  CALES_OK = 0,
  CALES_ERR_INVALID = 1,   /* invalid argument / unsupported configuration */
  CALES_ERR_CUDA = 2,      /* CUDA runtime error (no device, launch failure, ...) */
  CALES_ERR_NCCL = 3,      /* NCCL error */
  CALES_ERR_NOMEM = 4
};

enum { CALES_DIFF_EXPLICIT = 0, CALES_DIFF_IMPLICIT_3D = 1, CALES_DIFF_IMPLICIT_1D = 2 };

#define CALES_UNIQUE_ID_BYTES 128

/* ---- library / context -------------------------------------------------------------------- */
const char* cales_version(void);
const char* cales_last_error(const cales_ctx* ctx); /* ctx may be NULL: last error of a failed init */

/* NCCL bootstrap: rank 0 fills `uid` and the host broadcasts it (MPI_Bcast / torch.distributed),
 * as cuDecomp does in dependencies/cuDecomp/src/cudecomp.cc:66-80. */
int cales_get_unique_id(char uid[CALES_UNIQUE_ID_BYTES]);

/* replaces initmpi (src/initmpi.f90:34-206): builds the 2-D pencil decomposition of `ng` over
 * dims(1) x dims(2) ranks with pencils along `ipencil` (1,2,3 <- _DECOMP_X/_Y/_Z), selects the
 * device and, for nranks>1, creates the NCCL communicator.  `stream` is a cudaStream_t (NULL =
 * the legacy default stream).  `diffusion` selects what the reference selects with
 * -D_IMPDIFF / -D_IMPDIFF_1D. */
int cales_init(cales_ctx** ctx, const int ng[3], const int dims[2], int ipencil, const char cbcpre[6],
               int rank, int nranks, const char* nccl_uid, int device, void* stream, int diffusion);
int cales_finalize(cales_ctx* ctx);

/* the outputs of initmpi (src/initmpi.f90:42-44); 1-based lo/hi as in Fortran; nb = -1 where the
 * reference has MPI_PROC_NULL / CUDECOMP_RANK_NULL. */
int cales_get_decomp(const cales_ctx* ctx, int lo[3], int hi[3], int n[3], int n_x_fft[3], int n_y_fft[3],
                     int lo_z[3], int hi_z[3], int n_z[3], int nb[6], int is_bound[6]);

/* pure-integer decomposition maps, usable without a device (bit-exact contract):
 * 2decomp `distribute` (dependencies/2decomp-fft/src/decomp_2d.f90:1096-1147) == cuDecomp
 * `getSplits`/`cudecompGetPencilInfo` (dependencies/cuDecomp/src/cudecomp.cc:776-836). */
int cales_distribute(int data1, int proc, int* st, int* en, int* sz);
int cales_pencil(const int ng[3], const int dims[2], int rank, int axis /*1,2,3*/, int lo[3], int hi[3], int sz[3]);
int cales_neighbours(const int dims[2], int ipencil, const char cbcpre[6], int rank, int nb[6], int is_bound[6]);
/* exchange plan of one pencil transpose (which = 0 x->y, 1 y->z, 2 z->y, 3 y->x), the send/recv counts and offsets of
 * cuDecomp transpose.h:310-327 / 2decomp transpose_x_to_y.f90:87-96: per peer q of my row/column its global rank, the
 * sub-box {offset(3),extent(3)} of my source pencil sent to q and of my destination pencil received from q. */
int cales_transpose_plan(const int ng[3], const int dims[2], int rank, int which, int* npeers, int* peers, int* sendbox,
                         int* recvbox, int shapeA[3], int shapeB[3]);

int cales_stream_synchronize(cales_ctx* ctx);
/* number of kernels this library has launched on the context so far (bench.py `gpu_launches`) */
long cales_launch_count(const cales_ctx* ctx);

/* ---- solver set-up ------------------------------------------------------------------------------
 * replaces initsolver (src/initsolver.f90:17-64) + fftini (src/fft.f90:23-143).
 * dzci_g, dzfi_g: HOST vectors 0:ng3+1.  Outputs (HOST): lambdaxy(n_z(1),n_z(2)), a,b,c(ng3), normfft.
 * The spectral ordering is the reference CPU build's (FFTW halfcomplex, natural for DCT/DST), so
 * lambdaxy equals the reference's array element for element.  *plan receives an integer handle
 * (the role of arrplan(2,2)); cales_fftend releases it (src/fft.f90:145). */
int cales_initsolver(cales_ctx* ctx, const int ng[3], const int n_x_fft[3], const int n_y_fft[3], const int lo_z[3],
                     const int hi_z[3], const double dli[3], const double* dzci_g, const double* dzfi_g,
                     const char cbc[6], const char c_or_f[3], double* lambdaxy, double* a, double* b, double* c,
                     int* plan, double* normfft);
int cales_fftend(cales_ctx* ctx, int plan);

/* replaces solver / solver_gpu (src/solver.f90:20-80, src/solver_gpu.f90:32-164).
 * lambdaxy, a, b, c: DEVICE.  p: haloed field, solved in place on its interior. */
int cales_solver(cales_ctx* ctx, const int n[3], const int ng[3], int plan, double normfft, const double* lambdaxy,
                 const double* a, const double* b, const double* c, const char bc[6], const char c_or_f[3], double* p);
/* which exchange the last cales_solver call of this context used across ranks (diagnostic string; bench.py reports it) */
const char* cales_solver_exchange(const cales_ctx* ctx);
/* replaces solver_gaussel_z (src/solver.f90:182-233, src/solver_gpu.f90:374-477) */
int cales_solver_gaussel_z(cales_ctx* ctx, const int n[3], const double* a, const double* b, const double* c,
                           const char bcz[2], const char c_or_f[3], double* p);

/* ---- momentum / RK ---------------------------------------------------------------------------------
 * replaces rk (src/rk.f90:17-121) incl. mom_xyz_ad (src/mom.f90:17-309) and cmpt_bulk_forcing
 * (src/rk.f90:197-222).  The callee owns the RK history (the `save`d arrays, rk.f90:36-72).
 * f(3) is returned on the host (stream synchronised, as bulk_mean does, src/utils.f90:34-46); pass f = NULL to keep
 * it on the device only (no synchronisation) and hand NULL to cales_bulk_forcing, which then reads it there. */
int cales_rk(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci,
             const double* dzfi, const double* grid_vol_ratio_c, const double* grid_vol_ratio_f, double visc,
             double dt, const double* p, const int is_forced[3], const double velf[3], const double bforce[3],
             const double* visct, double* u, double* v, double* w, double f[3]);
/* rk with the update applied inside the momentum kernel (explicit diffusion; 112 instead of 160 B/cell): u,v,w are only
 * read, the updated velocity is written to the interior of un,vn,wn (other haloed arrays); same results as cales_rk,
 * bit for bit in the strict build; f(3) stays on the device.  cales_substep is built on it. */
int cales_rk_fused(cales_ctx* ctx, const double rkpar[2], const int n[3], const double dli[3], const double* dzci,
                   const double* dzfi, const double* grid_vol_ratio_c, const double* grid_vol_ratio_f, double visc,
                   double dt, const double* p, const int is_forced[3], const double velf[3], const double bforce[3],
                   const double* visct, const double* u, const double* v, const double* w, double* un, double* vn, double* wn);
/* mom_xyz_ad alone (src/mom.f90:17-309); dudtd.. may be NULL unless diffusion is implicit */
int cales_mom_xyz_ad(cales_ctx* ctx, const int n[3], double dxi, double dyi, const double* dzci, const double* dzfi,
                     double visc, const double* u, const double* v, const double* w, const double* visct,
                     double* dudt, double* dvdt, double* dwdt, double* dudtd, double* dvdtd, double* dwdtd);
/* replaces bulk_forcing (src/mom.f90:311-335) */
int cales_bulk_forcing(cales_ctx* ctx, const int n[3], const int is_forced[3], const double f[3], double* u,
                       double* v, double* w);
/* replaces bulk_mean (src/utils.f90:16-47) */
int cales_bulk_mean(cales_ctx* ctx, const int n[3], const double* grid_vol_ratio, const double* p, double* mean);

/* ---- boundary conditions ------------------------------------------------------------------------------
 * replace bounduvw (src/bound.f90:18-154, incl. updt_wallmodelbc src/wmodel.f90:19-335) and boundp
 * (src/bound.f90:156-200), halo exchange included (src/bound.f90:619-723). */
int cales_bounduvw(cales_ctx* ctx, const char cbc[18], const int n[3], const cales_bound* bcu, const cales_bound* bcv,
                   const cales_bound* bcw, const cales_bound* bcu_mag, const cales_bound* bcv_mag,
                   const cales_bound* bcw_mag, const int nb[6], const int is_bound[6], const int lwm[6],
                   const double l[3], const double dl[3], const double* zc, const double* zf, const double* dzc,
                   const double* dzf, double visc, double h, const int index_wm[6], int is_updt_wm, int is_correc,
                   double* u, double* v, double* w);
int cales_boundp(cales_ctx* ctx, const char cbc[6], const int n[3], const cales_bound* bcp, const int nb[6],
                 const int is_bound[6], const double dl[3], const double* dzc, double* p);
/* replace cmpt_rhs_b (src/bound.f90:447-495; dzc_g,dzf_g HOST 0:ng3+1; rhsb? DEVICE, may be NULL) and
 * updt_rhs_b (src/bound.f90:562-617) */
int cales_cmpt_rhs_b(cales_ctx* ctx, const int ng[3], const int n[3], const double dl[3], const double* dzc_g,
                     const double* dzf_g, const char cbc[6], const cales_bound* bc, const char c_or_f[3],
                     double* rhsbx, double* rhsby, double* rhsbz);
int cales_updt_rhs_b(cales_ctx* ctx, const char c_or_f[3], const char cbc[6], const int n[3], const int is_bound[6],
                     const double* rhsbx, const double* rhsby, const double* rhsbz, double* p);
/* scale rhsb planes: the `!$acc kernels` blocks of src/main.f90:426-432 (dst = src*alpha) */
int cales_scale(cales_ctx* ctx, long count, double alpha, const double* src, double* dst);
/* aa=a*alpha, bb=b*alpha+1, cc=c*alpha, lambdaxy=lambdaxyu*alpha (src/main.f90:434-441) */
int cales_helmholtz_coeffs(cales_ctx* ctx, int n3, long nxy, double alpha, const double* a, const double* b,
                           const double* c, const double* lambdaxy_in, double* aa, double* bb, double* cc,
                           double* lambdaxy_out);

/* ---- pressure correction ---------------------------------------------------------------------------------
 * replace fillps (src/fillps.f90:14-48), correc (src/correc.f90:14-68), updatep (src/updatep.f90:14-49) */
int cales_fillps(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzfi, double dti,
                 const double* u, const double* v, const double* w, double* p);
int cales_correc(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, double dt,
                 const double* p, double* u, double* v, double* w);
int cales_updatep(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, const double* dzfi,
                  double alpha, const double* pp, double* p);

/* ---- SGS model -----------------------------------------------------------------------------------------
 * replaces cmpt_sgs (src/sgs.f90:21-386): sgstype "none" | "smag" | "dsmag". */
int cales_cmpt_sgs(cales_ctx* ctx, const char* sgstype, const int n[3], const int ng[3], const int lo[3],
                   const int hi[3], const char cbcvel[18], const char cbcsgs[6], const cales_bound* bcs,
                   const int nb[6], const int is_bound[6], const int lwm[6], const double l[3], const double dl[3],
                   const double dli[3], const double* zc, const double* zf, const double* dzc, const double* dzf,
                   const double* dzci, const double* dzfi, double visc, double h, const int index_wm[6],
                   const double* u, const double* v, const double* w, const cales_bound* bcuf,
                   const cales_bound* bcvf, const cales_bound* bcwf, const cales_bound* bcu_mag,
                   const cales_bound* bcv_mag, const cales_bound* bcw_mag, double* visct);
/* the build-time switches of src/sgs.f90 for 'dsmag', selected at run time (before the first cales_cmpt_sgs):
 * ave: 0 _DIT (ave0d_dit, sgs.f90:388-431), 1 _CHANNEL (ave1d_channel, 433-538: the reference's hard-wired `#define`, sgs.f90:8;
 * the default), 2 _DUCT (ave2d_duct, 540-614, streamwise x), 3 _CAVITY (no averaging); filter_2d != 0: -D_FILTER_2D
 * (filter2d, sgs.f90:824-848, alph2 = 2.52 everywhere, no extrapolation) */
int cales_set_sgs_options(cales_ctx* ctx, int ave, int filter_2d);
/* building blocks exported for parity tests: strain_rate (src/sgs.f90:1019-1110; sij may be NULL,
 * else 6 haloed arrays back to back) and filter3d (src/sgs.f90:616-680) */
int cales_strain_rate(cales_ctx* ctx, const int n[3], const double dli[3], const double* dzci, const double* dzfi,
                      const double* u, const double* v, const double* w, double* s0, double* sij);
int cales_filter3d(cales_ctx* ctx, const int n[3], const double* p, double* pf);

/* ---- checks ----------------------------------------------------------------------------------------------
 * replace chkdt (src/chkdt.f90:17-99) and chkdiv (src/chkdiv.f90:16-52); host outputs. */
int cales_chkdt(cales_ctx* ctx, const int n[3], const double dl[3], const double* dzci, const double* dzfi,
                double visc, const double* visct, const double* u, const double* v, const double* w, double* dtmax);
int cales_chkdiv(cales_ctx* ctx, const int lo[3], const int hi[3], const double dli[3], const double* dzfi,
                 const double* u, const double* v, const double* w, double* divtot, double* divmax);

/* ---- input synthesis ---------------------------------------------------------------------------------------------
 * replaces initflow (src/initflow.f90:17-283) for the deterministic initial conditions, generated directly in the caller's
 * device arrays: inivel 'zer' 'uni' 'cou' 'poi' 'iop' 'hcp' 'pdc' 'hdc' 'tgv' 'tgw' 'ant' 'duc', set_mean (317-338) through
 * a device reduction (+ all-reduce), and the vortex pair of is_wallturb (233-260).  'log' 'hcl' 'tbl' add noise from the
 * Fortran compiler's random_number stream (285-315) and are refused (CALES_ERR_INVALID): they stay with the host.
 * zc,zf,dzc,dzf: DEVICE, rank-local (0:n3+1); bcvel(0:1,3,3) HOST in Fortran order; ghost cells are left zero. */
int cales_initflow(cales_ctx* ctx, const char* inivel, const double bcvel[18], const int ng[3], const int lo[3], const int n[3],
                   const double l[3], const double dl[3], const double* zc, const double* zf, const double* dzc,
                   const double* dzf, double visc, const int is_forced[3], const double velf[3], const double bforce[3],
                   int is_wallturb, double* u, double* v, double* w, double* p);

/* ---- on-the-fly statistics -----------------------------------------------------------------------------------
 * replaces the reductions of out1d_single_point_chan (src/output.f90:509-691, idir = 3; called every iout1d steps through
 * out1d.h90:35-36): the 27 plane-averaged single-point profiles (mean and moments of u,v,w, <uw>, p, p^2, vorticity,
 * modelled stresses, <nu_t>, du/dz), summed over ranks.  dzc, dzf: DEVICE, the rank-local slices (0:n3+1) of dzc_g, dzf_g;
 * u..visct: the rank-local haloed fields; buf: HOST, (27, ng3) in Fortran order.  Writing fname.out / fname.bin stays with
 * the host (cales_b200/stats.py).  The budget block of the same routine (output.f90:692-920) is not covered. */
int cales_out1d_chan(cales_ctx* ctx, const int ng[3], const int lo[3], const int hi[3], const double l[3], const double dl[3],
                     const double* dzc, const double* dzf, const double* u, const double* v, const double* w, const double* p,
                     const double* visct, double* buf);

/* ---- fused entries of the time loop (optional): identical results to the per-procedure sequence ------------
 * Everything the callees of one RK3 substep take (src/main.f90:418-506), gathered once.  HOST struct; its pointer members are
 * DEVICE arrays exactly as in the per-procedure exports above, the small descriptor vectors are held by value. */
typedef struct cales_step_args {
  int n[3], ng[3], lo[3], hi[3], nb[6], is_bound[6], lwm[6], index_wm[6], is_forced[3];
  int plan;                                     /* Poisson plan handle of cales_initsolver */
  double dl[3], dli[3], l[3], velf[3], bforce[3];
  double visc, hwm, normfft;
  char cbcvel[18], cbcpre[6], cbcsgs[6], sgstype[8];   /* sgstype NUL-terminated: "none" | "smag" | "dsmag" */
  const double *zc, *zf, *dzc, *dzf, *dzci, *dzfi, *grid_vol_ratio_c, *grid_vol_ratio_f;
  const double *lambdaxy, *a, *b, *c;           /* Poisson coefficients (cales_initsolver, uploaded) */
  const double *rhsbx, *rhsby, *rhsbz;          /* rhsbp planes of cales_cmpt_rhs_b (src/main.f90:317) */
  cales_bound bcu, bcv, bcw, bcp, bcs, bcu_mag, bcv_mag, bcw_mag, bcuf, bcvf, bcwf;
  double *u, *v, *w, *p, *pp, *visct;           /* the caller's fields */
} cales_step_args;
/* one RK3 substep irk = 1,2,3 of src/main.f90:418-506 (explicit diffusion): rk (update fused into the momentum kernel) ->
 * bulk_forcing -> bounduvw -> fillps -> updt_rhs_b -> solver -> boundp -> correc (+ updatep) -> bounduvw -> boundp ->
 * cmpt_sgs -> boundp; no host synchronisation; f(3) stays on the device */
int cales_substep(cales_ctx* ctx, const cales_step_args* a, int irk, double dt);
/* the three substeps of one time step; use_graph != 0 replays them from a CUDA graph captured per (dt, history parity)
 * on single-rank contexts with a non-default stream (falls back to the eager sequence otherwise) */
int cales_step(cales_ctx* ctx, const cales_step_args* a, double dt, int use_graph);
/* out = {sizeof(cales_step_args), offsetof cbcvel, offsetof zc, offsetof visct}: lets a binding verify its struct layout */
int cales_step_args_layout(long out[4]);

/* ---- building blocks exported for parity tests and benchmarks ---------------------------------------------- */
/* one batched 1-D transform pass of the solver on a halo-free array a(n1,n2,n3) in place:
 * dir 0 = along x, 1 = along y; bc = "PP","NN","DD",...; c_or_f 'c'|'f'; backward!=0 = inverse kind */
int cales_fft_lines(cales_ctx* ctx, const int n[3], int dir, const char bc[2], char c_or_f, int backward, double* a);
/* batched tridiagonal solve along z on a halo-free array (solver.f90:82-179); lambdaxy may be NULL */
int cales_gaussel(cales_ctx* ctx, int nx, int ny, int n, int periodic, const double* a, const double* b,
                  const double* c, const double* lambdaxy, double* p);
/* The distributed z solve of cales_solver (product build, z decomposed over 2..8 ranks with peer memory; csrc/zdist.cu) run
 * for P emulated ranks on ONE device: the P blocks of levels are solved by the kernels of the multi-GPU path (local Thomas
 * solve, boundary planes, tabulated inverse of the 2P x 2P interface system, correction pass), the exchange being a local
 * copy.  Same problem as cales_gaussel (solver.f90:82-179; a,b,c: DEVICE (n), lambdaxy: DEVICE (nx*ny), required), same
 * solution up to round-off.  pin != 0: the lambda = 0 column is singular (periodic / Neumann-Neumann z) and is regularised
 * by pinning one interface unknown, so that column is defined up to its additive constant. */
int cales_zdist_emulate(cales_ctx* ctx, int nx, int ny, int n, int P, int periodic, int pin, const double* a, const double* b,
                        const double* c, const double* lambdaxy, double* p);
/* pencil transposes (2decomp transpose_x_to_y etc. / cudecompTranspose*): which = 0 x->y, 1 y->z, 2 z->y, 3 y->x */
int cales_transpose(cales_ctx* ctx, int which, const double* src, double* dst);
/* work array the library owns, zero-filled, mapped into every rank of the node (CUDA IPC over NVLink; the role of cuDecomp's
 * cudecompMalloc, dependencies/cuDecomp/src/cudecomp.cc:913): a cales_transpose whose dst came from here stores each
 * sub-box straight into its owner's pencil over NVLink (the solver's own exchange) instead of pack + NCCL send/recv + unpack.
 * Collective over all ranks; plain device memory when peer access is unavailable.  Freed by cales_finalize. */
int cales_peer_alloc(cales_ctx* ctx, const char* name, long bytes, void** ptr);
/* halo exchange alone (src/bound.f90:619-723) */
int cales_updthalo(cales_ctx* ctx, const int n[3], const int nb[6], double* p);

#ifdef __cplusplus
}
#endif
#endif /* CALES_B200_H */
