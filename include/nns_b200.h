/*
 * nns_b200.h -- C ABI of the B200-native Navier-Stokes step library (libnns_b200.so).
 *
 * The reference (mhw32/neural-navier-stokes) is pure Python and has no FFI layer; this ABI
 * is the boundary a maintainer would bind UNDER the reference's solver classes.  Each entry
 * point names the reference code it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes stub that goes into the reference.
 *
 * Conventions
 *   - All fields are C-contiguous float64 [batch][nx][ny], index [i][j], j fastest; side ->
 *     index map as in src/boundary.py:39-46 (left=A[0,:], right=A[-1,:], bottom=A[:,0],
 *     top=A[:,-1]).
 *   - `*_run`, `*_step` and stage entry points take DEVICE pointers (caller-owned), are
 *     asynchronous and ordered on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream).  `*_host` entry points take HOST pointers and are synchronous; the
 *     host<->device copies happen inside the call.
 *   - Every function returns 0 on success or a negative nns_status; the message is
 *     available from nns_last_error() (thread-local).  No C++ exception crosses the ABI.
 *   - There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef NNS_B200_H
#define NNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNS_ABI_VERSION 2
#define NNS_MAX_BC 8 /* per field */

typedef enum nns_status {
    NNS_OK = 0,
    NNS_ERR_INVALID = -1,     /* bad argument (mirrors the reference's AssertionError / Exception) */
    NNS_ERR_CUDA = -2,        /* CUDA runtime failure, message carries cudaGetErrorString */
    NNS_ERR_UNSUPPORTED = -3, /* e.g. spectral Neumann BC (src/chorin_spectral/simulate.py:221) */
    NNS_ERR_NONFINITE = -4,   /* overflow/NaN seen; the reference raises (warnings are errors, chorin_fd:3) */
    NNS_ERR_NOMEM = -5
} nns_status;

enum { NNS_SOLVER_CHORIN_FD = 0, NNS_SOLVER_DIRECT_FD = 1, NNS_SOLVER_CHORIN_SPECTRAL = 2 };
enum { NNS_METHOD_EXPLICIT = 0, NNS_METHOD_SEMI_IMPLICIT = 1 };
enum { NNS_FIELD_U = 0, NNS_FIELD_V = 1, NNS_FIELD_P = 2 };
enum { NNS_SIDE_LEFT = 0, NNS_SIDE_RIGHT = 1, NNS_SIDE_BOTTOM = 2, NNS_SIDE_TOP = 3 };
enum { NNS_BC_DIRICHLET = 0, NNS_BC_NEUMANN = 1 };

/* One boundary condition; replaces a src/boundary.py BoundaryCondition object
 * (boundary.py:14-23).  The ORDER of the entries of one field is the order of the
 * reference's u_bc / v_bc / p_bc lists and is preserved (later entries overwrite corners). */
typedef struct nns_bc {
    int32_t field; /* NNS_FIELD_* */
    int32_t side;  /* NNS_SIDE_*  */
    int32_t type;  /* NNS_BC_*    */
    int32_t reserved;
    double value;
} nns_bc;

/* Constructor arguments of the reference's NavierStokesSystem
 * (src/chorin_fd/simulate.py:51-61, src/direct_fd/simulate.py:46-54,
 *  src/chorin_spectral/simulate.py:41-52) plus the ensemble batch. */
typedef struct nns_params {
    int32_t solver; /* NNS_SOLVER_* */
    int32_t method; /* NNS_METHOD_* (chorin_fd only) */
    int32_t nx, ny;
    int32_t batch;  /* independent simulations advanced together (1 = the reference's case) */
    int32_t nit;
    double dt, rho, nu, beta;
    double tol;     /* SOR exit tolerance; <= 0 selects the reference's 5e-6 (chorin_fd:183) */
    int32_t device; /* CUDA ordinal, -1 = current device */
    int32_t flags;  /* NNS_FLAG_* */
    double force_x; /* direct_fd with NNS_FLAG_PERIODIC_X: constant source added to u (dt * force_x per step), else ignored */
} nns_params;

#define NNS_FLAG_CHECK_FINITE 1 /* count non-finite outputs; *_host calls then return NNS_ERR_NONFINITE */
/* direct_fd only, EXTENSION (not in the reference, which has Dirichlet / Neumann conditions only, src/boundary.py:29-86;
 * BASELINE.json config 2 asks for it): the differenced axis 1 (dx, columns) is periodic -- every column is an interior
 * column whose neighbours wrap around -- and force_x drives the flow; the BC lists may then only name the sides
 * 'left' / 'right' (rows 0 and nx-1: the channel walls). */
#define NNS_FLAG_PERIODIC_X 4

typedef struct nns_handle nns_handle;

int32_t nns_abi_version(void);
const char *nns_last_error(void);

/* Create / destroy a solver instance (device workspace is allocated here, once).
 * nu_per_member: host array [batch] or NULL (use params->nu).
 * bc_value_per_member: host array [batch][n_bcs] overriding bcs[k].value per member, or NULL. */
int32_t nns_create(const nns_params *params, const nns_bc *bcs, int32_t n_bcs, const double *nu_per_member,
                   const double *bc_value_per_member, nns_handle **out);
int32_t nns_destroy(nns_handle *h);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t nns_launch_count(const nns_handle *h);
/* Non-finite values seen since creation (needs NNS_FLAG_CHECK_FINITE); synchronises the stream. */
int32_t nns_nonfinite_count(nns_handle *h, int64_t *count);

/* Sequential BC application on a device field [batch][nx][ny], in list order
 * (src/boundary.py:34-86; NavierStokesSystem._init_variables, chorin_fd/simulate.py:236-249).
 * Uses the handle's BC list of `field`. */
int32_t nns_apply_bc(nns_handle *h, int32_t field, double *a, void *stream);

/* ---- chorin_fd (src/chorin_fd/simulate.py) --------------------------------------- */

/* One time step, replaces NavierStokesSystem.step (chorin_fd/simulate.py:212-234):
 * predictor (:63-91 or :93-167) + u/v BCs (:221-225) + SOR pressure (:169-202, exact
 * lexicographic order, at most nit-1 sweeps, early exit at max|dp| <= tol) + p BCs
 * (:230-231) + projection (:204-210).  Reads u,v,u1,v1; writes u_out,v_out (must not
 * alias the inputs); p is updated in place.  sweeps_out: device int32 [batch] or NULL. */
int32_t nns_chorin_fd_step(nns_handle *h, const double *u, const double *v, const double *u1,
                           const double *v1, double *p, double *u_out, double *v_out, int32_t *sweeps_out,
                           void *stream);

/* The same step with HOST buffers (synchronous): the call the reference's step() would make.
 * 5 fields in (u, v, u1, v1, p), 3 out (u_out, v_out, p in place); sweeps_out host int32 [batch]
 * or NULL.  Member chunks are pipelined over internal streams; use pinned memory for speed. */
int32_t nns_chorin_fd_step_host(nns_handle *h, const double *u, const double *v, const double *u1,
                                const double *v1, double *p, double *u_out, double *v_out,
                                int32_t *sweeps_out);

/* nsteps time steps, replaces the loop of NavierStokesSystem.simulate (:258-265).  On
 * return u,v hold step n, u1,v1 step n-1, p step n.  traj_*: device [batch][nsteps][nx][ny]
 * or NULL (the reference keeps every step, :263-265).  sweeps_out: device int32
 * [nsteps][batch] or NULL. */
int32_t nns_chorin_fd_run(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p,
                          int32_t nsteps, double *traj_u, double *traj_v, double *traj_p,
                          int32_t *sweeps_out, void *stream);

/* Same with HOST buffers (synchronous; copies inside).  This is the call the reference's
 * simulate() would make.  traj_* / sweeps_out may be NULL. */
int32_t nns_chorin_fd_run_host(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p,
                               int32_t nsteps, double *traj_u, double *traj_v, double *traj_p,
                               int32_t *sweeps_out);

/* Stage entry points for unit parity (device pointers):
 *   predictor: _explicit/_semi_implicit_predictor_step + u_bc/v_bc   -> ui, vi
 *   pressure : _get_pressure (no p BCs), p in place                  -> p, sweeps_out[batch]
 *   correct  : p_bc + _correction_step, p in place                   -> u_out, v_out        */
int32_t nns_chorin_fd_predictor(nns_handle *h, const double *u, const double *v, const double *u1,
                                const double *v1, double *ui, double *vi, void *stream);
int32_t nns_chorin_fd_pressure(nns_handle *h, const double *ui, const double *vi, double *p,
                               int32_t *sweeps_out, void *stream);
int32_t nns_chorin_fd_correct(nns_handle *h, const double *ui, const double *vi, double *p, double *u_out,
                              double *v_out, void *stream);

/* ---- direct_fd (src/direct_fd/simulate.py) ---------------------------------------- */

/* nsteps of NavierStokesSystem.step (direct_fd/simulate.py:90-127): RHS b (:56-66), exactly
 * nit Jacobi sweeps with p BCs after every sweep (:68-88), upwind/central update (:98-118),
 * u/v BCs (:121-125).  u, v, p are advanced IN PLACE (as the reference does, :132).
 * traj_*: device [batch][nsteps][nx][ny] or NULL. */
int32_t nns_direct_fd_run(nns_handle *h, double *u, double *v, double *p, int32_t nsteps, double *traj_u,
                          double *traj_v, double *traj_p, void *stream);
int32_t nns_direct_fd_run_host(nns_handle *h, double *u, double *v, double *p, int32_t nsteps,
                               double *traj_u, double *traj_v, double *traj_p);

/* ---- chorin_spectral (src/chorin_spectral/simulate.py) ----------------------------- */

/* Number of operator arrays nns_spectral_set_operators expects, and the order:
 *   0 Dx  1 Dy  2 Dx_sqr  3 Dy_sqr                 interiors [1:-1,1:-1] of the matrices of :84-90
 *   4 uPinv 5 uQinv 6 uP 7 uQ  8 vPinv 9 vQinv 10 vP 11 vQ      Helmholtz eigenbases :174-183
 *   12 u_lambda_x 13 u_lambda_y 14 v_lambda_x 15 v_lambda_y
 *   16 pPinv 17 pQinv 18 pP 19 pQ 20 p_lambda_x 21 p_lambda_y   Uzawa operator :196-199
 *   22 DxDPx 23 DyDPy (:193-194)   24 S (n x m, :353-361)
 *   25 bvec_u 26 bvec_v = [b0_x(n) | bN_x(n) | b0_y(m) | bN_y(m)] (:102-118)
 *   27 sc_u 28 sc_v = [1/e_x, x0 constant, 1/e_y, y0 constant]    (:322-334)
 * with n = nx-2, m = ny-2; x-operators are n x n, y-operators m x m, all row-major float64. */
#define NNS_SPECTRAL_N_OPERATORS 29

/* Upload the operators of _pseudospectral_setup (chorin_spectral/simulate.py:59-199).  They are
 * built on the HOST exactly as the reference builds them (numpy + LAPACK eig/inv, one-time)
 * so that only the GEMM summation order differs from the reference.  ops: host pointers. */
int32_t nns_spectral_set_operators(nns_handle *h, const double *const *ops, int32_t n_ops);

/* Stage entry points (device pointers, [batch][nx][ny]):
 *   predictor: _predictor_step :232-337   un, vn, un1, vn1 -> ui, vi
 *   correct  : _correction_step :339-383  ui, vi, p -> u_out, v_out, p_out (p_out may alias p);
 *              q_out: optional [batch][nx-2][ny-2] copy of the pressure Q (:370-375) or NULL */
int32_t nns_spectral_predictor(nns_handle *h, const double *un, const double *vn, const double *un1,
                               const double *vn1, double *ui, double *vi, void *stream);
int32_t nns_spectral_correct(nns_handle *h, const double *ui, const double *vi, const double *p, double *u_out,
                             double *v_out, double *p_out, double *q_out, void *stream);

/* nsteps of step() with the rotation of simulate() (:547-570); same conventions as
 * nns_chorin_fd_run / nns_chorin_fd_run_host. */
int32_t nns_spectral_run(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p, int32_t nsteps,
                         double *traj_u, double *traj_v, double *traj_p, void *stream);
int32_t nns_spectral_run_host(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p,
                              int32_t nsteps, double *traj_u, double *traj_v, double *traj_p);

/* ---- chorin_fd on row slabs: ONE grid split over the GPUs of a box (new capability; the reference runs one
 * process) ---------------------------------------------------------------------------------------
 * Rank g owns the contiguous global rows [row0, row0 + nrows) (whole SOR tile rows; rank 0 also owns row 0,
 * the last rank row nx-1) and stores every field as [nrows + 2][ny]: one halo row above and below its rows.
 * nns_create is called with the GLOBAL nx, ny and batch = 1.  The lexicographic SOR of
 * src/chorin_fd/simulate.py:183-200 runs as a hyperplane of tiles across all slabs, neighbouring ranks swap
 * single rows of p over NCCL after every tick; results equal the single-GPU / reference results. */

/* Partition (pure host logic): rows owned by `rank`.  tile_rows <= 0 selects the library default. */
int32_t nns_slab_partition(int32_t nx, int32_t nranks, int32_t rank, int32_t tile_rows, int32_t *row0,
                           int32_t *nrows);
/* Tick plan (pure host logic): out8 = {tile rows TR, tile columns TC, nI, nJ, I0, I1, Ilo, Ihi}: the tile grid, the
 * tile rows [I0, I1) of `rank`, and the tile rows [Ilo, Ihi] it sweeps at (tick, sweep) -- empty if Ihi < Ilo.
 * Tile (I, J) = rows [1 + I*TR, 1 + (I+1)*TR) x columns [J*TC, (J+1)*TC) of the interior performs sweep s at tick
 * I + J + 2s; J = tick - 2*sweep - I. */
int32_t nns_slab_plan(int32_t nx, int32_t ny, int32_t nranks, int32_t rank, int32_t tile_rows, int32_t tick,
                      int32_t sweep, int32_t *out8);
/* The BC list of `field` in list order on a local slab (src/boundary.py:34-86; _init_variables :236-249). */
int32_t nns_slab_apply_bc(nns_handle *h, int32_t field, double *a, void *stream);
/* 128-byte NCCL id: rank 0 creates it and hands it to the other ranks (e.g. torch.distributed broadcast). */
int32_t nns_nccl_unique_id(uint8_t *id128);
/* Attach the handle to its slab; creates the NCCL communicator when nranks > 1 (collective call). */
int32_t nns_slab_attach(nns_handle *h, int32_t rank, int32_t nranks, const uint8_t *id128);
/* Optional peer-memory exchange for the SOR tick loop: every rank exports the 64-byte CUDA IPC handle of its
 * mailbox and maps the mailboxes of the ranks above / below (NULL at the ends).  Boundary rows of p then travel
 * as NVLink stores + flags issued from the tick loop's own stream (no NCCL launch per tick), and the tile rows
 * next to a neighbour are swept first so that the transfer hides behind the interior tile rows. */
int32_t nns_slab_ipc_export(nns_handle *h, uint8_t *handle64);
int32_t nns_slab_ipc_connect(nns_handle *h, const uint8_t *above64, const uint8_t *below64);
/* Swap the boundary rows of one local field with the neighbouring ranks (fills the halo rows). */
int32_t nns_slab_exchange(nns_handle *h, double *field, void *stream);
/* One time step (step(), chorin_fd/simulate.py:212-234) on the local slabs.  Halo rows of u, v, u1, v1, p must be
 * valid on entry (nns_slab_exchange once after initialisation); they are valid on return for u_out, v_out, p.
 * sweeps_out_host: host int32 or NULL.  The call synchronises the stream once (exit-test decision). */
int32_t nns_chorin_fd_slab_step(nns_handle *h, const double *u, const double *v, const double *u1, const double *v1,
                                double *p, double *u_out, double *v_out, int32_t *sweeps_out_host, void *stream);
/* Device time (CUDA events) of the SOR tick loop of the last step and its number of ticks (= sweep-kernel launches). */
int32_t nns_slab_last_timing(nns_handle *h, float *sor_ms, int32_t *ticks);

/* direct_fd on row slabs (src/direct_fd/simulate.py:56-127; one grid over the GPUs of a box): nsteps of step() on the
 * local slabs u, v, p ([nrows + 2][ny], halo rows valid on entry and on return; nns_slab_attach / nns_slab_apply_bc /
 * nns_slab_exchange as for chorin_fd).  Jacobi has no ordering dependency: every sweep is one kernel on the owned rows, the
 * p BCs and one halo-row exchange with the neighbouring ranks. */
int32_t nns_direct_fd_slab_run(nns_handle *h, double *u, double *v, double *p, int32_t nsteps, void *stream);

/* ---- trajectory sink (the interface between the time-step path and the rest of the reference) ---- */

/* Block means of device trajectories, replaces utils.spatial_coarsen (src/utils.py:13-60) for the u, v, p
 * sequences: inputs [frames][nx][ny] float64 (frames = members * nt), outputs [frames][nx/agg_x][ny/agg_y]
 * float64.  Every block mean is summed in NumPy's pairwise order and divided once, i.e. bit-identical to
 * np.mean over the flattened block; like the reference loop (utils.py:49) only the first ny // agg_x output
 * columns are written, the others are zero.  nx % agg_x, ny % agg_y != 0 -> NNS_ERR_INVALID (AssertionError
 * in the reference), ny // agg_x > ny // agg_y -> NNS_ERR_INVALID (IndexError there). */
int32_t nns_traj_coarsen(const double *u, const double *v, const double *p, int64_t frames, int32_t nx, int32_t ny,
                         int32_t agg_x, int32_t agg_y, double *u_out, double *v_out, double *p_out, void *stream);

/* The float32 observation tensor the neural scripts build from a data file
 * (src/neural_spectral/rnn.py:77-82, spectral_ode.py:158-163: stack([u, v, p]).permute(1, 0, 2, 3) of the
 * .float() fields): obs_out [frames][3][nx/agg_x][ny/agg_y] float32, optionally of the coarsened fields
 * (agg_x = agg_y = 1: plain conversion). */
int32_t nns_traj_observations(const double *u, const double *v, const double *p, int64_t frames, int32_t nx, int32_t ny,
                              int32_t agg_x, int32_t agg_y, float *obs_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NNS_B200_H */
