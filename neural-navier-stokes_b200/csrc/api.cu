// api.cu -- extern "C" entry points of libnns_b200 (see include/nns_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>

#include <new>
#include <vector>

#include "nns_common.cuh"

namespace nns {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// chorin_fd_chip.cu
struct ChipArgs;
int chorin_chip_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                    int step0, int phases, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps,
                    cudaStream_t st, int m0, int count);
bool chorin_chip_fits(const nns_handle *h);
void chorin_chip_free_plan(nns_handle *h);
// chorin_fd_stream.cu
bool chorin_stream_eligible(const nns_handle *h, int phases, int nsteps);
void chorin_stream_free(nns_handle *h);
void chorin_stream_prof_dump(nns_handle *h);
int chorin_stream_step(nns_handle *h, const double *uc, const double *vc, const double *up, const double *vp,
                       double *un, double *vn, double *p, double *tu, double *tv, double *tp,
                       size_t traj_member_stride, size_t traj_off, int32_t *sweeps, cudaStream_t st, int m0,
                       int count);
// chorin_fd_slab.cu
int chorin_tiled_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                     int step0, int phases, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps,
                     cudaStream_t st, int m0, int count);
int slab_partition(int nx, int nranks, int rank, int tile_rows, int *row0, int *nrows, int *I0, int *I1, int *nI);
void slab_free(nns_handle *h);
int slab_plan(int nx, int ny, int nranks, int rank, int tile_rows, int T, int s, int *out);
int slab_apply_bc(nns_handle *h, int field, double *a, cudaStream_t st);
int slab_unique_id(unsigned char *id128);
int slab_attach(nns_handle *h, int rank, int nranks, const unsigned char *id128);
int slab_exchange(nns_handle *h, double *f, cudaStream_t st);
int slab_last_timing(nns_handle *h, float *sor_ms, int *ticks);
int slab_ipc_export(nns_handle *h, unsigned char *handle64);
int slab_ipc_connect(nns_handle *h, const unsigned char *above64, const unsigned char *below64);
int direct_slab_run(nns_handle *h, double *u, double *v, double *p, int nsteps, cudaStream_t st);
int slab_step(nns_handle *h, const double *u, const double *v, const double *u1, const double *v1, double *p,
              double *un, double *vn, int32_t *sweeps_host, cudaStream_t st);
// direct_fd.cu
int direct_run(nns_handle *h, double *u, double *v, double *p, int nsteps, double *tu, double *tv, double *tp,
               cudaStream_t st);
int launch_apply_bc(nns_handle *h, int field, double *a, cudaStream_t st);
// spectral.cu
int spectral_create(nns_handle *h, const double *const *mats, int n_mats);
void spectral_destroy(nns_handle *h);
int spectral_predictor(nns_handle *h, const double *un, const double *vn, const double *un1, const double *vn1,
                       double *ui, double *vi, cudaStream_t st);
int spectral_correct(nns_handle *h, const double *ui, const double *vi, const double *p, double *uo, double *vo,
                     double *po, double *Qout, cudaStream_t st);
int spectral_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, double *tu, double *tv,
                 double *tp, cudaStream_t st);

// Cached device buffers of the host-buffer entry points: allocated once per handle and grown on demand, so that a
// user loop over step() / simulate() does not pay cudaMalloc / cudaFree (and their device synchronisation) per call.
static int pool_get(nns_handle *h, int k, size_t bytes, void **out) {
    if (h->pool_bytes[k] < bytes) {
        cudaFree(h->d_pool[k]);
        h->d_pool[k] = nullptr;
        h->pool_bytes[k] = 0;
        NNS_CUDA(cudaMalloc(&h->d_pool[k], bytes));
        h->pool_bytes[k] = bytes;
    }
    *out = h->d_pool[k];
    return NNS_OK;
}

// The reference raises per call (warnings are errors): the counter of non-finite values restarts with every host call.
static int reset_nonfinite(nns_handle *h, cudaStream_t st) {
    NNS_CUDA(cudaMemsetAsync(h->d_nonfinite, 0, sizeof(unsigned long long), st));
    return NNS_OK;
}

static int ensure_scratch(nns_handle *h, int k) {
    if (h->d_scratch[k]) return NNS_OK;
    NNS_CUDA(cudaMalloc(&h->d_scratch[k], sizeof(double) * (size_t)h->g.nx * h->g.ny * h->g.batch));
    return NNS_OK;
}

static int chorin_dispatch(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps,
                           int nsteps_total, int step0, int phases, int fixup, double *tu, double *tv,
                           double *tp, int32_t *sweeps, cudaStream_t st, int m0 = 0, int count = -1) {
    if (chorin_stream_eligible(h, phases, nsteps)) {
        // persistent warp-specialised kernel: one launch per step over all members of the range
        const size_t N = (size_t)h->g.nx * h->g.ny;
        const int cnt = count < 0 ? h->g.batch - m0 : count;
        const size_t bytes = sizeof(double) * N * cnt;
        int cur = 0, prev = 1, nxt = 2, rc;
        for (int n = 0; n < nsteps; ++n) {
            rc = chorin_stream_step(h, bufU[cur], bufV[cur], bufU[prev], bufV[prev], bufU[nxt], bufV[nxt], p, tu, tv, tp,
                                    (size_t)nsteps_total * N, (size_t)(step0 + n) * N,
                                    sweeps ? sweeps + (size_t)(step0 + n) * h->g.batch : nullptr, st, m0, cnt);
            if (rc != NNS_OK) return rc;
            const int t = prev; prev = cur; cur = nxt; nxt = t;
        }
        if (fixup && cur != 0) {
            if (cur == 2) {          // (cur, prev) = (2, 0)
                NNS_CUDA(cudaMemcpyAsync(bufU[1], bufU[0], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufV[1], bufV[0], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufU[0], bufU[2], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufV[0], bufV[2], bytes, cudaMemcpyDeviceToDevice, st));
            } else {                 // (cur, prev) = (1, 2)
                NNS_CUDA(cudaMemcpyAsync(bufU[0], bufU[1], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufV[0], bufV[1], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufU[1], bufU[2], bytes, cudaMemcpyDeviceToDevice, st));
                NNS_CUDA(cudaMemcpyAsync(bufV[1], bufV[2], bytes, cudaMemcpyDeviceToDevice, st));
            }
        }
        return NNS_OK;
    }
    if (chorin_chip_fits(h))
        return chorin_chip_run(h, bufU, bufV, p, nsteps, nsteps_total, step0, phases, fixup, tu, tv, tp, sweeps, st,
                               m0, count);
    return chorin_tiled_run(h, bufU, bufV, p, nsteps, nsteps_total, step0, phases, fixup, tu, tv, tp, sweeps, st,
                            m0, count);
}

}  // namespace nns

using namespace nns;

#define NNS_CHECK_HANDLE(h, solver_)                                           \
    do {                                                                       \
        if (!(h)) { set_error("null handle"); return NNS_ERR_INVALID; }        \
        if ((h)->params.solver != (solver_)) {                                 \
            set_error("handle was created for solver %d", (h)->params.solver); \
            return NNS_ERR_INVALID;                                            \
        }                                                                      \
        NNS_CUDA(cudaSetDevice((h)->device));                                  \
    } while (0)

extern "C" {

int32_t nns_abi_version(void) { return NNS_ABI_VERSION; }
const char *nns_last_error(void) { return g_err; }

static int32_t create_impl(const nns_params *P, const nns_bc *bcs, int32_t n_bcs, const double *nu_b,
                           const double *bcval_b, nns_handle **out, nns_handle **partial);

int32_t nns_create(const nns_params *P, const nns_bc *bcs, int32_t n_bcs, const double *nu_b,
                   const double *bcval_b, nns_handle **out) {
    nns_handle *partial = nullptr;
    const int32_t rc = create_impl(P, bcs, n_bcs, nu_b, bcval_b, out, &partial);
    if (rc != NNS_OK && partial) nns_destroy(partial);      // no leak of the handle / its device buffers on a failed create
    return rc;
}

static int32_t create_impl(const nns_params *P, const nns_bc *bcs, int32_t n_bcs, const double *nu_b,
                           const double *bcval_b, nns_handle **out, nns_handle **partial) {
    if (!P || !out || (n_bcs > 0 && !bcs)) { set_error("nns_create: null argument"); return NNS_ERR_INVALID; }
    *out = nullptr;
    if (P->nx < 3 || P->ny < 3 || P->batch < 1 || P->nit < 0) {
        set_error("nns_create: need nx,ny >= 3, batch >= 1, nit >= 0 (got %d,%d,%d,%d)", P->nx, P->ny, P->batch, P->nit);
        return NNS_ERR_INVALID;
    }
    if (P->solver < NNS_SOLVER_CHORIN_FD || P->solver > NNS_SOLVER_CHORIN_SPECTRAL) {
        set_error("nns_create: unknown solver %d", P->solver);
        return NNS_ERR_INVALID;
    }
    if (P->solver == NNS_SOLVER_CHORIN_FD && P->method != NNS_METHOD_EXPLICIT && P->method != NNS_METHOD_SEMI_IMPLICIT) {
        set_error("method not recognized: %d", P->method);   // chorin_fd/simulate.py:60,218
        return NNS_ERR_INVALID;
    }
    if (P->solver == NNS_SOLVER_CHORIN_FD && P->method == NNS_METHOD_SEMI_IMPLICIT && P->nx != P->ny) {
        set_error("semi_implicit needs nx == ny (the reference solves B along axis 0, chorin_fd/simulate.py:159)");
        return NNS_ERR_INVALID;
    }
    if (!(P->dt > 0) || !(P->rho != 0)) { set_error("nns_create: dt must be > 0 and rho != 0"); return NNS_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libnns_b200 has no CPU fallback");
        return NNS_ERR_CUDA;
    }
    nns_handle *h = new (std::nothrow) nns_handle();
    if (!h) { set_error("out of host memory"); return NNS_ERR_NOMEM; }
    memset(h, 0, sizeof(*h));
    *partial = h;
    h->params = *P;
    if (P->device >= 0) h->device = P->device;
    else NNS_CUDA(cudaGetDevice(&h->device));
    NNS_CUDA(cudaSetDevice(h->device));
    NNS_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
    NNS_CUDA(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    Geometry &g = h->g;
    g.nx = P->nx; g.ny = P->ny; g.batch = P->batch; g.nit = P->nit; g.method = P->method;
    g.dt = P->dt; g.rho = P->rho; g.nu = P->nu; g.beta = P->beta;
    g.tol = P->tol > 0 ? P->tol : 5e-6;
    if (P->solver == NNS_SOLVER_CHORIN_SPECTRAL) { g.dx = 2.0 / P->nx; g.dy = 2.0 / P->ny; }   // chorin_spectral:48
    else { g.dx = 2.0 / (P->nx - 1); g.dy = 2.0 / (P->ny - 1); }                               // chorin_fd:58
    h->n_bcs = n_bcs;
    for (int k = 0; k < n_bcs; ++k) {
        const nns_bc &b = bcs[k];
        if (b.field < 0 || b.field > 2 || b.side < 0 || b.side > 3 || b.type < 0 || b.type > 1) {
            set_error("nns_create: bad boundary condition #%d (field %d side %d type %d)", k, b.field, b.side, b.type);
            return NNS_ERR_INVALID;
        }
        BcList &L = h->bc[b.field];
        if (L.n >= NNS_MAX_BC) { set_error("more than %d boundary conditions on one field", NNS_MAX_BC); return NNS_ERR_INVALID; }
        L.side[L.n] = b.side; L.type[L.n] = b.type; L.value[L.n] = b.value; L.slot[L.n] = k;
        L.n++;
    }
    if (P->flags & NNS_FLAG_PERIODIC_X) {
        if (P->solver != NNS_SOLVER_DIRECT_FD) { set_error("NNS_FLAG_PERIODIC_X is a direct_fd extension"); return NNS_ERR_INVALID; }
        for (int f = 0; f < 3; ++f)
            for (int k = 0; k < h->bc[f].n; ++k)
                if (h->bc[f].side[k] == NNS_SIDE_BOTTOM || h->bc[f].side[k] == NNS_SIDE_TOP) {
                    set_error("periodic x: boundary conditions may only name 'left' / 'right' (the channel walls)");
                    return NNS_ERR_INVALID;
                }
    }
    if (nu_b) {
        NNS_CUDA(cudaMalloc(&h->d_nu, sizeof(double) * g.batch));
        NNS_CUDA(cudaMemcpy(h->d_nu, nu_b, sizeof(double) * g.batch, cudaMemcpyHostToDevice));
        h->h_nu = static_cast<double *>(malloc(sizeof(double) * g.batch));
        if (!h->h_nu) { set_error("out of host memory"); return NNS_ERR_NOMEM; }
        memcpy(h->h_nu, nu_b, sizeof(double) * g.batch);
    }
    if (bcval_b && n_bcs > 0) {
        NNS_CUDA(cudaMalloc(&h->d_bcval, sizeof(double) * (size_t)g.batch * n_bcs));
        NNS_CUDA(cudaMemcpy(h->d_bcval, bcval_b, sizeof(double) * (size_t)g.batch * n_bcs, cudaMemcpyHostToDevice));
        h->h_bcval = static_cast<double *>(malloc(sizeof(double) * (size_t)g.batch * n_bcs));
        if (!h->h_bcval) { set_error("out of host memory"); return NNS_ERR_NOMEM; }
        memcpy(h->h_bcval, bcval_b, sizeof(double) * (size_t)g.batch * n_bcs);
    }
    NNS_CUDA(cudaMalloc(&h->d_nonfinite, sizeof(unsigned long long)));
    NNS_CUDA(cudaMemset(h->d_nonfinite, 0, sizeof(unsigned long long)));
    NNS_CUDA(cudaMalloc(&h->d_sweeps, sizeof(int32_t) * g.batch));
    *out = h;
    *partial = nullptr;
    return NNS_OK;
}

int32_t nns_destroy(nns_handle *h) {
    if (!h) return NNS_OK;
    cudaSetDevice(h->device);
    chorin_stream_prof_dump(h);
    cudaFree(h->d_nu); cudaFree(h->d_bcval); cudaFree(h->d_cprime); cudaFree(h->d_b); cudaFree(h->d_p2);
    cudaFree(h->d_sweeps); cudaFree(h->d_nonfinite); cudaFree(h->d_blockdesc);
    chorin_chip_free_plan(h);
    chorin_stream_free(h);
    slab_free(h);
    spectral_destroy(h);
    for (int k = 0; k < 4; ++k) cudaFree(h->d_scratch[k]);
    for (int k = 0; k < 7; ++k) cudaFree(h->d_stage[k]);
    for (int k = 0; k < 10; ++k) cudaFree(h->d_pool[k]);
    free(h->h_nu); free(h->h_bcval);
    for (int k = 0; k < 4; ++k) if (h->streams[k]) cudaStreamDestroy(h->streams[k]);
    delete h;
    return NNS_OK;
}

int64_t nns_launch_count(const nns_handle *h) { return h ? h->launches : 0; }

int32_t nns_nonfinite_count(nns_handle *h, int64_t *count) {
    if (!h || !count) { set_error("null argument"); return NNS_ERR_INVALID; }
    NNS_CUDA(cudaSetDevice(h->device));
    unsigned long long c = 0;
    NNS_CUDA(cudaMemcpy(&c, h->d_nonfinite, sizeof(c), cudaMemcpyDeviceToHost));
    *count = (int64_t)c;
    return NNS_OK;
}

int32_t nns_apply_bc(nns_handle *h, int32_t field, double *a, void *stream) {
    if (!h || !a || field < 0 || field > 2) { set_error("nns_apply_bc: bad argument"); return NNS_ERR_INVALID; }
    NNS_CUDA(cudaSetDevice(h->device));
    return launch_apply_bc(h, field, a, (cudaStream_t)stream);
}

// ---- chorin_fd -------------------------------------------------------------------------------

int32_t nns_chorin_fd_step(nns_handle *h, const double *u, const double *v, const double *u1, const double *v1,
                           double *p, double *u_out, double *v_out, int32_t *sweeps_out, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !p || !u_out || !v_out) { set_error("nns_chorin_fd_step: null field"); return NNS_ERR_INVALID; }
    if (u_out == u || u_out == u1 || v_out == v || v_out == v1) { set_error("nns_chorin_fd_step: outputs alias inputs"); return NNS_ERR_INVALID; }
    double *bu[3] = {const_cast<double *>(u), const_cast<double *>(u1), u_out};
    double *bv[3] = {const_cast<double *>(v), const_cast<double *>(v1), v_out};
    return chorin_dispatch(h, bu, bv, p, 1, 1, 0, 7, 0, nullptr, nullptr, nullptr, sweeps_out, (cudaStream_t)stream);
}

int32_t nns_chorin_fd_run(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p, int32_t nsteps,
                          double *tu, double *tv, double *tp, int32_t *sweeps_out, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !p || nsteps < 0) { set_error("nns_chorin_fd_run: bad argument"); return NNS_ERR_INVALID; }
    if (nsteps == 0) return NNS_OK;
    int rc;
    if ((rc = ensure_scratch(h, 0)) || (rc = ensure_scratch(h, 1))) return rc;
    double *bu[3] = {u, u1, h->d_scratch[0]};
    double *bv[3] = {v, v1, h->d_scratch[1]};
    return chorin_dispatch(h, bu, bv, p, nsteps, nsteps, 0, 7, 1, tu, tv, tp, sweeps_out, (cudaStream_t)stream);
}

int32_t nns_chorin_fd_predictor(nns_handle *h, const double *u, const double *v, const double *u1,
                                const double *v1, double *ui, double *vi, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !ui || !vi) { set_error("nns_chorin_fd_predictor: null field"); return NNS_ERR_INVALID; }
    double *bu[3] = {const_cast<double *>(u), const_cast<double *>(u1), ui};
    double *bv[3] = {const_cast<double *>(v), const_cast<double *>(v1), vi};
    return chorin_dispatch(h, bu, bv, nullptr, 1, 1, 0, 1, 0, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int32_t nns_chorin_fd_pressure(nns_handle *h, const double *ui, const double *vi, double *p, int32_t *sweeps_out,
                               void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!ui || !vi || !p) { set_error("nns_chorin_fd_pressure: null field"); return NNS_ERR_INVALID; }
    double *bu[3] = {nullptr, nullptr, const_cast<double *>(ui)};
    double *bv[3] = {nullptr, nullptr, const_cast<double *>(vi)};
    return chorin_dispatch(h, bu, bv, p, 1, 1, 0, 2, 0, nullptr, nullptr, nullptr, sweeps_out, (cudaStream_t)stream);
}

int32_t nns_chorin_fd_correct(nns_handle *h, const double *ui, const double *vi, double *p, double *u_out,
                              double *v_out, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!ui || !vi || !p || !u_out || !v_out) { set_error("nns_chorin_fd_correct: null field"); return NNS_ERR_INVALID; }
    const size_t bytes = sizeof(double) * (size_t)h->g.nx * h->g.ny * h->g.batch;
    cudaStream_t st = (cudaStream_t)stream;
    if (u_out != ui) NNS_CUDA(cudaMemcpyAsync(u_out, ui, bytes, cudaMemcpyDeviceToDevice, st));
    if (v_out != vi) NNS_CUDA(cudaMemcpyAsync(v_out, vi, bytes, cudaMemcpyDeviceToDevice, st));
    double *bu[3] = {nullptr, nullptr, u_out};
    double *bv[3] = {nullptr, nullptr, v_out};
    return chorin_dispatch(h, bu, bv, p, 1, 1, 0, 4, 0, nullptr, nullptr, nullptr, nullptr, st);
}

// Host-buffer helpers ------------------------------------------------------------------------
namespace {
// Large device -> pageable-host copies (the trajectories of the *_run_host entry points: the reference's simulate()
// returns (nt, nx, ny) numpy arrays) through two pinned staging buffers: the DMA of chunk i + 1 overlaps the host
// memcpy of chunk i.  A plain cudaMemcpy into pageable memory ran at ~2 GB/s (3.1 GB of direct_fd 256^2
// trajectories: 1.5 s of a 1.7 s simulate()).
int d2h_staged(void *dst, const void *src, size_t bytes) {
    constexpr size_t CHUNK = (size_t)32 << 20;
    if (bytes <= CHUNK) { NNS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return NNS_OK; }
    static thread_local void *stage[2] = {nullptr, nullptr};
    cudaStream_t st;
    cudaEvent_t ev[2];
    for (int k = 0; k < 2; ++k)
        if (!stage[k]) NNS_CUDA(cudaHostAlloc(&stage[k], CHUNK, cudaHostAllocDefault));
    NNS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) NNS_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    auto len = [&](size_t i) { return i + 1 < n ? CHUNK : bytes - i * CHUNK; };
    cudaError_t e = cudaMemcpyAsync(stage[0], src, len(0), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(ev[0], st);
    for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
        if (i + 1 < n) {
            e = cudaMemcpyAsync(stage[(i + 1) & 1], static_cast<const char *>(src) + (i + 1) * CHUNK, len(i + 1), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[(i + 1) & 1], st);
        }
        if (e == cudaSuccess) e = cudaEventSynchronize(ev[i & 1]);
        if (e == cudaSuccess) memcpy(static_cast<char *>(dst) + i * CHUNK, stage[i & 1], len(i));
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    for (int k = 0; k < 2; ++k) cudaEventDestroy(ev[k]);
    if (e != cudaSuccess) { set_error("staged device-to-host copy failed: %s", cudaGetErrorString(e)); return NNS_ERR_CUDA; }
    return NNS_OK;
}
}  // namespace

}  // extern "C" (templates need C++ linkage)

// Shared body of the *_run_host entry points: fields (and optional trajectories / sweep counts) through cached device
// buffers; `run` advances the device state.
template <typename Run>
static int32_t run_host_common(nns_handle *h, int nfields, double *const *hostf, int32_t nsteps, double *const *hostt,
                               int32_t *sweeps_out, const char *what, Run &&run) {
    const size_t bytes = sizeof(double) * (size_t)h->g.nx * h->g.ny * h->g.batch;
    int rc;
    double *f[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *t[3] = {nullptr, nullptr, nullptr};
    int32_t *dsw = nullptr;
    if ((rc = reset_nonfinite(h, nullptr))) return rc;
    for (int k = 0; k < nfields; ++k) {
        if ((rc = pool_get(h, k, bytes, reinterpret_cast<void **>(&f[k])))) return rc;
        NNS_CUDA(cudaMemcpyAsync(f[k], hostf[k], bytes, cudaMemcpyHostToDevice, nullptr));
    }
    const bool traj = hostt[0] && hostt[1] && hostt[2] && nsteps > 0;
    if (traj)
        for (int k = 0; k < 3; ++k)
            if ((rc = pool_get(h, 5 + k, bytes * nsteps, reinterpret_cast<void **>(&t[k])))) return rc;
    if (sweeps_out && nsteps > 0 && (rc = pool_get(h, 8, sizeof(int32_t) * (size_t)nsteps * h->g.batch, reinterpret_cast<void **>(&dsw)))) return rc;
    rc = run(f, t, dsw);
    if (rc == NNS_OK) {
        for (int k = 0; k < nfields; ++k) cudaMemcpyAsync(hostf[k], f[k], bytes, cudaMemcpyDeviceToHost, nullptr);
        cudaError_t e = cudaStreamSynchronize(nullptr);
        if (e != cudaSuccess) { set_error("%s: kernel failed: %s", what, cudaGetErrorString(e)); rc = NNS_ERR_CUDA; }
    }
    if (rc == NNS_OK) {
        if (traj) for (int k = 0; k < 3 && rc == NNS_OK; ++k) rc = d2h_staged(hostt[k], t[k], bytes * nsteps);
        if (dsw) cudaMemcpy(sweeps_out, dsw, sizeof(int32_t) * (size_t)nsteps * h->g.batch, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("%s: copy back failed: %s", what, cudaGetErrorString(e)); rc = NNS_ERR_CUDA; }
    }
    if (rc == NNS_OK && (h->params.flags & NNS_FLAG_CHECK_FINITE)) {
        int64_t c = 0;
        if (nns_nonfinite_count(h, &c) == NNS_OK && c > 0) {
            set_error("non-finite values in u/v/p (%lld cells): the reference raises here (warnings are errors)", (long long)c);
            rc = NNS_ERR_NONFINITE;
        }
    }
    return rc;
}

extern "C" {

int32_t nns_chorin_fd_run_host(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p,
                               int32_t nsteps, double *tu, double *tv, double *tp, int32_t *sweeps_out) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !p || nsteps < 0) { set_error("nns_chorin_fd_run_host: bad argument"); return NNS_ERR_INVALID; }
    double *hostf[5] = {u, v, u1, v1, p}, *hostt[3] = {tu, tv, tp};
    return run_host_common(h, 5, hostf, nsteps, hostt, sweeps_out, "nns_chorin_fd_run_host", [&](double **f, double **t, int32_t *dsw) {
        return nns_chorin_fd_run(h, f[0], f[1], f[2], f[3], f[4], nsteps, t[0], t[1], t[2], dsw, nullptr);
    });
}

// One step with HOST buffers, mirroring the reference's step(un, vn, un1, vn1, p) -> (u, v, p):
// 5 fields in, 3 fields out (u_out, v_out, p in place).  The batch is cut into member chunks
// that are pipelined over several streams (H2D of chunk c+1 || kernel of chunk c || D2H of
// chunk c-1), so with pinned host memory the call runs at the PCIe rate.
int32_t nns_chorin_fd_step_host(nns_handle *h, const double *u, const double *v, const double *u1,
                                const double *v1, double *p, double *u_out, double *v_out, int32_t *sweeps_out) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !p || !u_out || !v_out) { set_error("nns_chorin_fd_step_host: null field"); return NNS_ERR_INVALID; }
    const size_t N = (size_t)h->g.nx * h->g.ny, B = h->g.batch;
    const size_t bytes = sizeof(double) * N * B;
    for (int k = 0; k < 7; ++k)
        if (!h->d_stage[k]) NNS_CUDA(cudaMalloc(&h->d_stage[k], bytes));
    for (int k = 0; k < 4; ++k)
        if (!h->streams[k]) NNS_CUDA(cudaStreamCreateWithFlags(&h->streams[k], cudaStreamNonBlocking));
    double **d = h->d_stage;       // u, v, u1, v1, p, u_out, v_out
    const double *hin[5] = {u, v, u1, v1, p};
    double *hout[3] = {u_out, v_out, p};
    const int dout[3] = {5, 6, 4};
    // >= 148 CTAs per launch; many chunks keep the un-overlapped head (first H2D) and tail (last kernel + D2H) short
    int nchunks = B >= 2368 ? 16 : B >= 1184 ? 8 : B >= 296 ? 2 : 1;
    if (const char *e = getenv("NNS_STEP_HOST_CHUNKS")) { const int v = atoi(e); if (v >= 1 && (size_t)v <= B) nchunks = v; }     // experiments
    const size_t per = (B + nchunks - 1) / nchunks;
    int rc = reset_nonfinite(h, nullptr);
    if (rc == NNS_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) { set_error("step_host: stream error"); rc = NNS_ERR_CUDA; }
    for (int c = 0; c < nchunks && rc == NNS_OK; ++c) {
        const size_t m0 = c * per, cnt = m0 + per <= B ? per : B - m0;
        if (m0 >= B) break;
        cudaStream_t st = h->streams[c % 4];
        h->scratch_slot = c % 4;             // launches of different internal streams may overlap: one scratch set each
        const size_t off = m0 * N, cb = sizeof(double) * cnt * N;
        for (int k = 0; k < 5; ++k) NNS_CUDA(cudaMemcpyAsync(d[k] + off, hin[k] + off, cb, cudaMemcpyHostToDevice, st));
        double *bu[3] = {d[0] + off, d[2] + off, d[5] + off};
        double *bv[3] = {d[1] + off, d[3] + off, d[6] + off};
        rc = chorin_dispatch(h, bu, bv, d[4] + off, 1, 1, 0, 7, 0, nullptr, nullptr, nullptr,
                             h->d_sweeps + m0, st, (int)m0, (int)cnt);
        if (rc != NNS_OK) break;
        for (int k = 0; k < 3; ++k) NNS_CUDA(cudaMemcpyAsync(hout[k] + off, d[dout[k]] + off, cb, cudaMemcpyDeviceToHost, st));
        if (sweeps_out) NNS_CUDA(cudaMemcpyAsync(sweeps_out + m0, h->d_sweeps + m0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, st));
    }
    h->scratch_slot = 0;
    for (int k = 0; k < 4; ++k) {
        cudaError_t e = cudaStreamSynchronize(h->streams[k]);
        if (e != cudaSuccess && rc == NNS_OK) { set_error("step_host: %s", cudaGetErrorString(e)); rc = NNS_ERR_CUDA; }
    }
    if (rc == NNS_OK && (h->params.flags & NNS_FLAG_CHECK_FINITE)) {
        int64_t c = 0;
        if (nns_nonfinite_count(h, &c) == NNS_OK && c > 0) { set_error("non-finite values in u/v/p (%lld cells)", (long long)c); rc = NNS_ERR_NONFINITE; }
    }
    return rc;
}

// ---- direct_fd --------------------------------------------------------------------------------

int32_t nns_direct_fd_run(nns_handle *h, double *u, double *v, double *p, int32_t nsteps, double *tu, double *tv,
                          double *tp, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_DIRECT_FD);
    if (!u || !v || !p || nsteps < 0) { set_error("nns_direct_fd_run: bad argument"); return NNS_ERR_INVALID; }
    if ((tu || tv || tp) && !(tu && tv && tp)) { set_error("nns_direct_fd_run: pass all three trajectory buffers or none"); return NNS_ERR_INVALID; }
    if (nsteps == 0) return NNS_OK;
    return direct_run(h, u, v, p, nsteps, tu, tv, tp, (cudaStream_t)stream);
}

int32_t nns_direct_fd_run_host(nns_handle *h, double *u, double *v, double *p, int32_t nsteps, double *tu,
                               double *tv, double *tp) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_DIRECT_FD);
    if (!u || !v || !p || nsteps < 0) { set_error("nns_direct_fd_run_host: bad argument"); return NNS_ERR_INVALID; }
    double *hostf[3] = {u, v, p}, *hostt[3] = {tu, tv, tp};
    return run_host_common(h, 3, hostf, nsteps, hostt, nullptr, "nns_direct_fd_run_host", [&](double **f, double **t, int32_t *) {
        return nns_direct_fd_run(h, f[0], f[1], f[2], nsteps, t[0], t[1], t[2], nullptr);
    });
}

// ---- chorin_spectral --------------------------------------------------------------------------

int32_t nns_spectral_set_operators(nns_handle *h, const double *const *ops, int32_t n_ops) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_SPECTRAL);
    if (!ops) { set_error("nns_spectral_set_operators: null argument"); return NNS_ERR_INVALID; }
    return spectral_create(h, ops, n_ops);
}

#define NNS_CHECK_SPECTRAL(h)                                                                                   \
    do {                                                                                                        \
        NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_SPECTRAL);                                                        \
        if (!(h)->spectral) { set_error("chorin_spectral: call nns_spectral_set_operators first"); return NNS_ERR_INVALID; } \
    } while (0)

int32_t nns_spectral_predictor(nns_handle *h, const double *un, const double *vn, const double *un1,
                               const double *vn1, double *ui, double *vi, void *stream) {
    NNS_CHECK_SPECTRAL(h);
    if (!un || !vn || !un1 || !vn1 || !ui || !vi) { set_error("nns_spectral_predictor: null field"); return NNS_ERR_INVALID; }
    return spectral_predictor(h, un, vn, un1, vn1, ui, vi, (cudaStream_t)stream);
}

int32_t nns_spectral_correct(nns_handle *h, const double *ui, const double *vi, const double *p, double *u_out,
                             double *v_out, double *p_out, double *q_out, void *stream) {
    NNS_CHECK_SPECTRAL(h);
    if (!ui || !vi || !p || !u_out || !v_out || !p_out) { set_error("nns_spectral_correct: null field"); return NNS_ERR_INVALID; }
    if (u_out == ui || v_out == vi) { set_error("nns_spectral_correct: u_out/v_out alias ui/vi"); return NNS_ERR_INVALID; }
    return spectral_correct(h, ui, vi, p, u_out, v_out, p_out, q_out, (cudaStream_t)stream);
}

int32_t nns_spectral_run(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p, int32_t nsteps,
                         double *tu, double *tv, double *tp, void *stream) {
    NNS_CHECK_SPECTRAL(h);
    if (!u || !v || !u1 || !v1 || !p || nsteps < 0) { set_error("nns_spectral_run: bad argument"); return NNS_ERR_INVALID; }
    if ((tu || tv || tp) && !(tu && tv && tp)) { set_error("nns_spectral_run: pass all three trajectory buffers or none"); return NNS_ERR_INVALID; }
    if (nsteps == 0) return NNS_OK;
    int rc;
    if ((rc = ensure_scratch(h, 0)) || (rc = ensure_scratch(h, 1))) return rc;
    double *bu[3] = {u, u1, h->d_scratch[0]};
    double *bv[3] = {v, v1, h->d_scratch[1]};
    return spectral_run(h, bu, bv, p, nsteps, tu, tv, tp, (cudaStream_t)stream);
}

int32_t nns_spectral_run_host(nns_handle *h, double *u, double *v, double *u1, double *v1, double *p,
                              int32_t nsteps, double *tu, double *tv, double *tp) {
    NNS_CHECK_SPECTRAL(h);
    if (!u || !v || !u1 || !v1 || !p || nsteps < 0) { set_error("nns_spectral_run_host: bad argument"); return NNS_ERR_INVALID; }
    double *hostf[5] = {u, v, u1, v1, p}, *hostt[3] = {tu, tv, tp};
    return run_host_common(h, 5, hostf, nsteps, hostt, nullptr, "nns_spectral_run_host", [&](double **f, double **t, int32_t *) {
        return nns_spectral_run(h, f[0], f[1], f[2], f[3], f[4], nsteps, t[0], t[1], t[2], nullptr);
    });
}

// ---- chorin_fd on row slabs (one grid over several GPUs) ---------------------------------------

#define NNS_CHECK_SLAB_HANDLE(h)                                                                                     \
    do {                                                                                                              \
        if (!(h)) { set_error("null handle"); return NNS_ERR_INVALID; }                                               \
        if ((h)->params.solver != NNS_SOLVER_CHORIN_FD && (h)->params.solver != NNS_SOLVER_DIRECT_FD) {               \
            set_error("row slabs exist for chorin_fd and direct_fd (handle was created for solver %d)", (h)->params.solver); \
            return NNS_ERR_INVALID;                                                                                   \
        }                                                                                                             \
        NNS_CUDA(cudaSetDevice((h)->device));                                                                         \
    } while (0)

int32_t nns_slab_partition(int32_t nx, int32_t nranks, int32_t rank, int32_t tile_rows, int32_t *row0, int32_t *nrows) {
    return slab_partition(nx, nranks, rank, tile_rows, row0, nrows, nullptr, nullptr, nullptr);
}

int32_t nns_slab_plan(int32_t nx, int32_t ny, int32_t nranks, int32_t rank, int32_t tile_rows, int32_t tick, int32_t sweep,
                      int32_t *out8) {
    if (!out8) { set_error("nns_slab_plan: null argument"); return NNS_ERR_INVALID; }
    return slab_plan(nx, ny, nranks, rank, tile_rows, tick, sweep, out8);
}

int32_t nns_slab_apply_bc(nns_handle *h, int32_t field, double *a, void *stream) {
    NNS_CHECK_SLAB_HANDLE(h);
    if (!a || field < 0 || field > 2) { set_error("nns_slab_apply_bc: bad argument"); return NNS_ERR_INVALID; }
    return slab_apply_bc(h, field, a, (cudaStream_t)stream);
}

int32_t nns_nccl_unique_id(uint8_t *id128) {
    if (!id128) { set_error("nns_nccl_unique_id: null argument"); return NNS_ERR_INVALID; }
    return slab_unique_id(id128);
}

int32_t nns_slab_attach(nns_handle *h, int32_t rank, int32_t nranks, const uint8_t *id128) {
    NNS_CHECK_SLAB_HANDLE(h);
    if (nranks < 1 || rank < 0 || rank >= nranks) { set_error("nns_slab_attach: bad rank %d of %d", rank, nranks); return NNS_ERR_INVALID; }
    return slab_attach(h, rank, nranks, id128);
}

int32_t nns_slab_exchange(nns_handle *h, double *field, void *stream) {
    NNS_CHECK_SLAB_HANDLE(h);
    if (!field) { set_error("nns_slab_exchange: null field"); return NNS_ERR_INVALID; }
    return slab_exchange(h, field, (cudaStream_t)stream);
}

int32_t nns_slab_ipc_export(nns_handle *h, uint8_t *handle64) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!handle64) { set_error("nns_slab_ipc_export: null argument"); return NNS_ERR_INVALID; }
    return slab_ipc_export(h, handle64);
}

int32_t nns_slab_ipc_connect(nns_handle *h, const uint8_t *above64, const uint8_t *below64) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    return slab_ipc_connect(h, above64, below64);
}

int32_t nns_slab_last_timing(nns_handle *h, float *sor_ms, int32_t *ticks) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!sor_ms || !ticks) { set_error("nns_slab_last_timing: null argument"); return NNS_ERR_INVALID; }
    return slab_last_timing(h, sor_ms, ticks);
}

int32_t nns_chorin_fd_slab_step(nns_handle *h, const double *u, const double *v, const double *u1, const double *v1,
                                double *p, double *u_out, double *v_out, int32_t *sweeps_out_host, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_CHORIN_FD);
    if (!u || !v || !u1 || !v1 || !p || !u_out || !v_out) { set_error("nns_chorin_fd_slab_step: null field"); return NNS_ERR_INVALID; }
    if (u_out == u || u_out == u1 || v_out == v || v_out == v1) { set_error("nns_chorin_fd_slab_step: outputs alias inputs"); return NNS_ERR_INVALID; }
    return slab_step(h, u, v, u1, v1, p, u_out, v_out, sweeps_out_host, (cudaStream_t)stream);
}

int32_t nns_direct_fd_slab_run(nns_handle *h, double *u, double *v, double *p, int32_t nsteps, void *stream) {
    NNS_CHECK_HANDLE(h, NNS_SOLVER_DIRECT_FD);
    if (!u || !v || !p || nsteps < 0) { set_error("nns_direct_fd_slab_run: bad argument"); return NNS_ERR_INVALID; }
    if (h->params.flags & NNS_FLAG_PERIODIC_X) { set_error("nns_direct_fd_slab_run: the periodic-x extension is not available on slabs"); return NNS_ERR_UNSUPPORTED; }
    if (nsteps == 0) return NNS_OK;
    return direct_slab_run(h, u, v, p, nsteps, (cudaStream_t)stream);
}

}  // extern "C"
