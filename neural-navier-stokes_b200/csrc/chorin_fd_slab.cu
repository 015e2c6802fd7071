// chorin_fd_slab.cu -- chorin_fd (explicit) for grids whose pressure field does NOT fit one SM: fields stay
// in HBM, the exact lexicographic SOR runs as a HYPERPLANE of tiles, and one grid can be split into ROW SLABS
// over the GPUs of a box (one process per GPU, NCCL send/recv of single halo rows over NVLink).
//
// Reference semantics (src/chorin_fd/simulate.py of mhw32/neural-navier-stokes): predictor :63-91, u/v BCs
// :221-225, right-hand side :186-188, lexicographic SOR with early exit :183-200, p BCs :230-231, projection
// :204-210.
//
// SOR.  The interior is cut into tiles of TR rows x TC = 32*LC columns.  Tile (I, J) performs sweep s at
// tick T = I + J + 2s: its lexicographic predecessors -- tile (I-1, J) and (I, J-1) at the same sweep, tile
// (I+1, J) and (I, J+1) at the previous sweep -- ran at tick T-1, and no two tiles of one tick touch, so one
// kernel launch per tick executes every tile of the hyperplane (up to nit-1 sweeps in flight) and the result
// is the sequential one.  Inside a tile ONE WARP pipelines the rows across its lanes: lane l owns LC columns
// and works on row k at step k + l; the freshly updated west value arrives from lane l-1 by shuffle, the
// north values are the lane's own previous results (registers), south / east / centre are still old in
// memory.  No shared memory, no intra-tile barrier, exact order.
// The sweeps are not temporally blocked (24 B per cell and sweep from HBM), see DESIGN.md 4.5.
//
// Slabs.  Rank g owns whole tile rows (plus the physical boundary rows at the ends) and stores its rows with
// one halo row above and below.  All ranks run the same global tick loop; after every tick neighbouring
// ranks swap their boundary rows of p (the row above supplies "north, same sweep", the row below "south,
// previous sweep").  u, v, the predictor output and the final p swap halo rows once per step.  The per-sweep
// exit flags are max-reduced over the ranks, so every rank takes the reference's early-exit decision.
#include <cuda.h>
#include <dlfcn.h>
#include <stdlib.h>

#include <algorithm>

#include "nns_common.cuh"

namespace nns {

namespace {

// ---- NCCL through dlopen (the library a torch process already loaded, else the system one) -------------
typedef struct { char internal[128]; } NcclId;
typedef void *NcclComm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclMax = 2;

NcclApi *nccl_api() {
    static NcclApi api;
    if (api.lib) return &api;
    const char *names[] = {getenv("NNS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !n[0]) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return nullptr; }
#define NNS_SYM(f)                                                                                     \
    *reinterpret_cast<void **>(&api.f) = dlsym(api.lib, "nccl" #f);                                    \
    if (!api.f) { set_error("NCCL symbol nccl" #f " not found"); api.lib = nullptr; return nullptr; }
    NNS_SYM(GetUniqueId) NNS_SYM(CommInitRank) NNS_SYM(CommDestroy) NNS_SYM(Send) NNS_SYM(Recv) NNS_SYM(AllReduce)
    NNS_SYM(GroupStart) NNS_SYM(GroupEnd) NNS_SYM(GetErrorString)
#undef NNS_SYM
    return &api;
}

#define NNS_NCCL(call)                                                                                 \
    do {                                                                                               \
        int r__ = (call);                                                                              \
        if (r__ != 0) {                                                                                \
            set_error("%s failed: %s", #call, nccl_api()->GetErrorString(r__));                        \
            return NNS_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

constexpr int LC = 4, TC = 32 * LC;       // columns per lane / per tile

struct SlabState {
    int rank = 0, nranks = 1;
    int row0 = 0, nrows = 0;      // owned global rows [row0, row0 + nrows)
    int TR = 128;                 // tile rows
    int nI = 0, nJ = 0;           // global tile grid over the interior
    int I0 = 0, I1 = 0;           // owned tile rows
    NcclComm comm = nullptr;
    double *d_cprime = nullptr, *d_p0 = nullptr;    // [nrows + 2][ny]
    double *d_thomas = nullptr;                     // semi-implicit: elimination coefficients of the two ADI stages, 2 x 2 nx
    double *d_own[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // single-GPU run(): haloed copies
    int *d_flags = nullptr;       // [64] per-sweep "not converged" flags
    int *h_flags = nullptr;       // pinned
    cudaEvent_t ev[2] = {nullptr, nullptr};   // around the tick loop of the last step
    // peer-memory halo exchange of p inside the tick loop (NVLink stores + flags instead of NCCL launches)
    unsigned char *d_box = nullptr;           // this rank's mailbox: rows [dir][buf][ny] doubles, then flags[2] (uint32)
    unsigned char *peer_box[2] = {nullptr, nullptr};   // mapped mailboxes of the rank above (0) / below (1)
    bool p2p = false;
    unsigned seq = 0;                         // tick sequence number (monotonic over the run)
    cudaStream_t st_x = nullptr;              // direct_fd slabs: exchange stream (halo rows under the interior sweep)
    cudaEvent_t ev_edge = nullptr, ev_xdone = nullptr;
    CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    int last_ticks = 0;
};

struct SlabView {                 // addressing of the local slab of a field: element (i, j) of the GLOBAL grid
    int rowbase;                  // global index of the first stored row (the top halo row)
    int ny;
    __host__ __device__ size_t at(int i, int j) const { return (size_t)(i - rowbase) * ny + j; }
};

int tile_rows_for(int nx) {
    const char *e = getenv("NNS_SLAB_TR");
    if (e && atoi(e) > 0) return atoi(e);
    return nx >= 1024 ? 128 : 16;
}

// ---- kernels ----------------------------------------------------------------------------------------
struct SlabGeom {
    int nx, ny;                   // global grid
    int row0, row1;               // owned rows [row0, row1)
    SlabView v;
    double dt, dx, dy, rho, nu, beta, tol;
};

// _explicit_predictor_step (chorin_fd:63-91): both advection terms difference along axis 0 (:74,76,83,85)
__global__ void slab_predictor_kernel(SlabGeom g, const double *__restrict__ uc, const double *__restrict__ vc,
                                      const double *__restrict__ up, const double *__restrict__ vp,
                                      double *__restrict__ un, double *__restrict__ vn) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    const int ny = g.ny;
    const double u0 = uc[q], v0 = vc[q];
    double ru = u0, rv = v0;
    if (i > 0 && i < g.nx - 1 && j > 0 && j < ny - 1) {
        const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
        const double a0x = 1.5 * g.dt / (2.0 * g.dx), a0y = 1.5 * g.dt / (2.0 * g.dy);
        const double a1x = 0.5 * g.dt / (2.0 * g.dx), a1y = 0.5 * g.dt / (2.0 * g.dy);
        const double c0x = 1.5 * g.dt * g.nu / dx2, c0y = 1.5 * g.dt * g.nu / dy2;
        const double c1x = 0.5 * g.dt * g.nu / dx2, c1y = 0.5 * g.dt * g.nu / dy2;
        const double a0 = up[q], b0 = vp[q];
        const double uS = uc[q + ny], uN = uc[q - ny], uE = uc[q + 1], uW = uc[q - 1];
        const double vS = vc[q + ny], vN = vc[q - ny], vE = vc[q + 1], vW = vc[q - 1];
        const double aS = up[q + ny], aN = up[q - ny], aE = up[q + 1], aW = up[q - 1];
        const double bS = vp[q + ny], bN = vp[q - ny], bE = vp[q + 1], bW = vp[q - 1];
        const double k0 = fma(u0, a0x, v0 * a0y), k1 = fma(a0, a1x, b0 * a1y);
        const double lu = fma(-2.0, u0, uS + uN), mu = fma(-2.0, u0, uE + uW);
        const double la = fma(-2.0, a0, aS + aN), ma = fma(-2.0, a0, aE + aW);
        const double lv = fma(-2.0, v0, vS + vN), mv = fma(-2.0, v0, vE + vW);
        const double lb = fma(-2.0, b0, bS + bN), mb = fma(-2.0, b0, bE + bW);
        ru = fma(-c1y, ma, fma(-c1x, la, fma(c0y, mu, fma(c0x, lu, fma(k1, aS - aN, fma(-k0, uS - uN, u0))))));
        rv = fma(-c1y, mb, fma(-c1x, lb, fma(c0y, mv, fma(c0x, lv, fma(k1, bS - bN, fma(-k0, vS - vN, v0))))));
    }
    un[q] = ru;
    vn[q] = rv;
}

// ---- semi-implicit predictor on a whole grid in HBM (single GPU; chorin_fd:93-167): AB2 advection + Crank-Nicolson ADI,
// all four solves along axis 0 with the matrices of :105-121 (diagonal (2/nu) dx^2 + 2 dt, the reference's precedence),
// the same arithmetic as phase_predictor of chorin_fd_chip.cu.  Columns are independent: one thread per interior
// column marches down the rows (coalesced across the columns), the elimination coefficients are shared by all columns.
__global__ void semi_rhs_kernel(SlabGeom g, const double *__restrict__ uc, const double *__restrict__ vc,
                                const double *__restrict__ up, const double *__restrict__ vp,
                                double *__restrict__ un, double *__restrict__ vn) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    const int ny = g.ny;
    const double u0 = uc[q], v0 = vc[q];
    double ru = u0, rv = v0;
    if (i > 0 && i < g.nx - 1 && j > 0 && j < ny - 1) {
        const double dt = g.dt, nu = g.nu, dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
        const double r2dx = 1.0 / (2.0 * g.dx), r2dy = 1.0 / (2.0 * g.dy), rdx2 = 1.0 / dx2, rdy2 = 1.0 / dy2;
        const double kx = 2.0 / nu * dx2;
        const double u1c = up[q], v1c = vp[q];
        const double uS = uc[q + ny], uN = uc[q - ny], uE = uc[q + 1], uW = uc[q - 1];
        const double vS = vc[q + ny], vN = vc[q - ny], vE = vc[q + 1], vW = vc[q - 1];
        const double pS = up[q + ny], pN = up[q - ny], pE = up[q + 1], pW = up[q - 1];
        const double qS = vp[q + ny], qN = vp[q - ny], qE = vp[q + 1], qW = vp[q - 1];
        const double uHn = u0 * (uS - uN) * r2dx + v0 * (uE - uW) * r2dy;
        const double uHn1 = u1c * (pS - pN) * r2dx + v1c * (pE - pW) * r2dy;
        const double vHn = u0 * (vS - vN) * r2dx + v0 * (vE - vW) * r2dy;
        const double vHn1 = u1c * (qS - qN) * r2dx + v1c * (qE - qW) * r2dy;
        const double uC2 = dt * nu * ((uS - 2.0 * u0 + uN) * rdx2 + (uE - 2.0 * u0 + uW) * rdy2);
        const double vC2 = dt * nu * ((vS - 2.0 * v0 + vN) * rdx2 + (vE - 2.0 * v0 + vW) * rdy2);
        ru = kx * (0.5 * dt * (3.0 * uHn - uHn1) + uC2);
        rv = kx * (0.5 * dt * (3.0 * vHn - vHn1) + vC2);
    }
    un[q] = ru;
    vn[q] = rv;
}

// cpr[i] = c'_i, cpr[nx + i] = 1 / m_i of the constant tridiagonal (off, diag, off): one thread
__global__ void thomas_coef_kernel(int nx, double diag, double off, double *cpr) {
    double c = 0.0;
    for (int i = 1; i < nx - 1; ++i) {
        const double m = diag - off * c;
        c = off / m;
        cpr[i] = c;
        cpr[nx + i] = 1.0 / m;
    }
}

// np.linalg.solve(A, rhs) along axis 0 for the interior columns of two fields (blockIdx.y), in place (A is diagonally
// dominant => LAPACK's partial pivoting never swaps, so this is the same elimination)
__global__ void thomas_axis0_kernel(SlabGeom g, double *x0, double *x1, double off, const double *__restrict__ cpr) {
    const int j = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= g.ny - 1) return;
    double *x = blockIdx.y ? x1 : x0;
    const int nx = g.nx;
    double d = 0.0;
    for (int i = 1; i < nx - 1; ++i) {
        const size_t q = g.v.at(i, j);
        d = (x[q] - off * d) * cpr[nx + i];
        x[q] = d;
    }
    double xn = 0.0;
    for (int i = nx - 2; i >= 1; --i) {
        const size_t q = g.v.at(i, j);
        xn = x[q] - cpr[i] * xn;
        x[q] = xn;
    }
}

// right-hand side of the second ADI stage (chorin_fd:155-165)
__global__ void semi_mid_kernel(SlabGeom g, const double *__restrict__ uc, const double *__restrict__ vc,
                                double *__restrict__ un, double *__restrict__ vn) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j < 1 || j >= g.ny - 1 || i < 1 || i >= g.nx - 1) return;
    const size_t q = g.v.at(i, j);
    const double ky = 2.0 / g.nu * (g.dy * g.dy);
    const double u0 = uc[q], v0 = vc[q];
    un[q] = ky * (un[q] + u0) - g.dt * (uc[q + 1] - 2.0 * u0 + uc[q - 1]);
    vn[q] = ky * (vn[q] + v0) - g.dt * (vc[q + 1] - 2.0 * v0 + vc[q - 1]);
}

// one entry of a BC list on the owned rows (boundary.py:34-86); launched in list order
__global__ void slab_bc_kernel(SlabGeom g, double *A, int side, int neumann, double value) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
        const int i = side == NNS_SIDE_LEFT ? 0 : g.nx - 1, in = side == NNS_SIDE_LEFT ? 1 : g.nx - 2;
        if (i < g.row0 || i >= g.row1 || t >= g.ny) return;
        const double sgn = side == NNS_SIDE_LEFT ? -g.dx : g.dx;
        A[g.v.at(i, t)] = neumann ? A[g.v.at(in, t)] + sgn * value : value;
    } else {
        const int i = g.row0 + t;
        if (i >= g.row1) return;
        const int j = side == NNS_SIDE_BOTTOM ? 0 : g.ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : g.ny - 2;
        const double sgn = side == NNS_SIDE_BOTTOM ? -g.dy : g.dy;
        A[g.v.at(i, j)] = neumann ? A[g.v.at(i, jn)] + sgn * value : value;
    }
}

// pre-scaled right-hand side C' (chorin_fd:186-188)
__global__ void slab_cprime_kernel(SlabGeom g, const double *__restrict__ ui, const double *__restrict__ vi,
                                   double *__restrict__ cp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    double c = 0.0;
    if (i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1) {
        const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy, den = 2.0 * dx2 + 2.0 * dy2;
        const double cc = g.beta / den, cu = g.dx * g.rho * dy2 / g.dt, cv = g.dy * g.rho * dx2 / g.dt;
        c = cc * (cu * (ui[q] - ui[q - g.ny]) + cv * (vi[q] - vi[q - 1]));
    }
    cp[q] = c;
}

// projection (chorin_fd:204-210) in place on (ui, vi), + optional snapshot / non-finite count
__global__ void slab_project_kernel(SlabGeom g, const double *__restrict__ p, double *__restrict__ un,
                                    double *__restrict__ vn, unsigned long long *nonfinite, int check) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    double ru = un[q], rv = vn[q];
    if (i > 0 && i < g.nx - 1 && j > 0 && j < g.ny - 1) {
        ru -= g.dt / (2.0 * g.dx) * (p[q + g.ny] - p[q - g.ny]);
        rv -= g.dt / (2.0 * g.dy) * (p[q + 1] - p[q - 1]);
        un[q] = ru;
        vn[q] = rv;
    }
    if (check && !(isfinite(ru) && isfinite(rv) && isfinite(p[q]))) atomicAdd(nonfinite, 1ull);
}

// Tile rows of sweep s that a rank owning tile rows [I0, I1) works on at tick T (tile (I, J) runs sweep s at
// T = I + J + 2s, 0 <= J < nJ).  Shared by the kernel and the exported plan function.
__host__ __device__ inline bool tick_tile_rows(int T, int s, int I0, int I1, int nJ, int *Ilo, int *Ihi) {
    const int d = T - 2 * s;
    if (d < 0) return false;
    *Ilo = I0 > d - (nJ - 1) ? I0 : d - (nJ - 1);
    *Ihi = I1 - 1 < d ? I1 - 1 : d;
    return *Ilo <= *Ihi;
}

struct SweepArgs {
    SlabGeom g;
    int TR, nJ, I0, I1, cap;
    double *p;
    const double *cp;
    int *flags;        // [cap] set to 1 when sweep s still violates the exit test; null = no tracking
    // peer-memory exchange fused into the sweep (null = the halo rows of p are maintained by the host loop):
    // the rows just outside the slab are READ from this rank's mailbox, and the slab's first / last row are also
    // STORED into the neighbours' mailboxes (NVLink) by the tiles that update them.
    const double *halo_top, *halo_bot;
    double *push_up, *push_dn;
    int first_row, last_row;      // global indices of the slab's first / last owned row
};

// 4 consecutive doubles of a row: one 256-bit access when the address is 32-byte aligned (VEC), else scalars
template <bool VEC>
__device__ __forceinline__ void load4(const double *p, int nvalid, double (&x)[LC]) {
    if (VEC) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
    } else {
#pragma unroll
        for (int jj = 0; jj < LC; ++jj) x[jj] = jj < nvalid ? p[jj] : 0.0;
    }
}
template <bool VEC>
__device__ __forceinline__ void store4(double *p, int nvalid, const double (&x)[LC]) {
    if (VEC) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
    } else {
#pragma unroll
        for (int jj = 0; jj < LC; ++jj)
            if (jj < nvalid) p[jj] = x[jj];
    }
}

// All tiles of hyperplane T (see the header).  blockIdx.y = sweep, 4 warps per CTA = 4 tiles.
// Tile J covers the global columns [J*TC, (J+1)*TC): lane chunks start on multiples of 4, so with ny % 4 == 0
// every row chunk is ONE 256-bit load / store (LDG.E.ENL2.256).  That matters: the lanes of a warp work on 32
// different rows, every access of the warp touches 32 cache lines, and the kernel is bound by the number of
// load / store INSTRUCTIONS (one LSU cycle per line), not by bytes.  Per step and lane: one load of the row
// below, one of C', one store; the east operand comes from lane + 1 by shuffle (it holds that value as its own
// south / first row), the west one from lane - 1 (freshly updated).
template <bool VEC>
__global__ void __launch_bounds__(128) slab_sweep_kernel(const SweepArgs a, int T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y;
    int Ilo, Ihi;
    if (s >= a.cap || !tick_tile_rows(T, s, a.I0, a.I1, a.nJ, &Ilo, &Ihi)) return;
    const int I = Ilo + blockIdx.x * 4 + warp;
    if (I > Ihi) return;
    const int J = T - 2 * s - I;
    const SlabGeom &g = a.g;
    const int ny = g.ny;
    const int i0 = 1 + I * a.TR, i1 = min(i0 + a.TR, g.nx - 1);
    const int nr = i1 - i0;
    const int c0 = J * TC + lane * LC;                    // first column of the lane's chunk (multiple of 4)
    const int nin = max(0, min(LC, ny - c0));             // columns of the chunk inside the grid
    const bool has = nin > 0 && c0 < ny - 1;              // the chunk holds at least one interior column
    bool ok[LC];                                          // interior (updated) cells of the chunk
#pragma unroll
    for (int jj = 0; jj < LC; ++jj) ok[jj] = c0 + jj >= 1 && c0 + jj <= ny - 2;
    const bool east_load = has && (lane == 31 || c0 + LC >= ny - 1) && c0 + LC <= ny - 1;   // no lane + 1 to ask
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy, den = 2.0 * dx2 + 2.0 * dy2;
    const double ca = g.beta * dy2 / den, cb = g.beta * dx2 / den, mbeta = -g.beta, tol = g.tol;
    double *P = a.p;
    const double *CP = a.cp;
    // row i of p for reading: the neighbour's row in the mailbox if i lies just outside the slab
    auto rowp = [&](int i) -> const double * {
        if (a.halo_top && i == a.first_row - 1) return a.halo_top;
        if (a.halo_bot && i == a.last_row + 1) return a.halo_bot;
        return P + g.v.at(i, 0);
    };

    // Software pipeline: the operands of row k are loaded PF steps before they are used, into a register ring
    // indexed by the step number (tau mod PF is the same for the load and the use of a row on every lane).
#ifndef NNS_SLAB_PF
#define NNS_SLAB_PF 4
#endif
    constexpr int PF = NNS_SLAB_PF;
    double pn[LC], pc[LC];
    double qs[PF][LC], qc[PF][LC], qe[PF], qw[PF];
    double wlast = 0.0;            // the lane's last value of the current row (west operand of lane + 1)
    bool viol = false;
#pragma unroll
    for (int jj = 0; jj < LC; ++jj) { pn[jj] = 0.0; pc[jj] = 0.0; }
#pragma unroll
    for (int dd = 0; dd < PF; ++dd) {
        qe[dd] = 0.0; qw[dd] = 0.0;
#pragma unroll
        for (int jj = 0; jj < LC; ++jj) { qs[dd][jj] = 0.0; qc[dd][jj] = 0.0; }
    }
    auto prefetch = [&](int k, int slot) {        // operands of tile row k into ring slot `slot`
        if (has && k >= 0 && k < nr) {
            const size_t q = g.v.at(i0 + k, c0);
            load4<VEC>(rowp(i0 + k + 1) + c0, nin, qs[slot]);
            load4<VEC>(CP + q, nin, qc[slot]);
            if (east_load) qe[slot] = P[q + LC];
            if (lane == 0 && c0 > 0) qw[slot] = P[q - 1];
        }
    };
    // the lane's first row and the row above it, once, before the pipeline starts (a load inside the step loop
    // would stall the WHOLE warp at the next use of pc / pn: the scoreboard tracks registers per warp)
    if (has) {
        const size_t q = g.v.at(i0, c0);
        load4<VEC>(rowp(i0 - 1) + c0, nin, pn);
        load4<VEC>(P + q, nin, pc);
    }
#pragma unroll
    for (int dd = 0; dd < PF; ++dd) prefetch(dd - lane, dd);        // steps tau = 0 .. PF-1 use slots 0 .. PF-1
    for (int tau0 = 0; tau0 < nr + 31; tau0 += PF) {
#pragma unroll
        for (int dd = 0; dd < PF; ++dd) {
            const int tau = tau0 + dd;
            const int k = tau - lane;
            const bool act = has && k >= 0 && k < nr;
            // west: lane - 1 finished row k one step ago.  east: lane + 1 is one row behind and holds row k as the
            // row below its current one (or, before its first step, as its preloaded first row).
            const double wsh = __shfl_up_sync(0xffffffffu, wlast, 1);
            const double esh = __shfl_down_sync(0xffffffffu, k >= 0 ? qs[dd][0] : pc[0], 1);     // evaluated by the SOURCE lane with its own k
            if (act) {
                const size_t q = g.v.at(i0 + k, c0);
                double w = lane == 0 ? qw[dd] : wsh;
                const double east = east_load ? qe[dd] : esh;
                double out[LC];
#pragma unroll
                for (int jj = 0; jj < LC; ++jj) {
                    const double e = jj + 1 < LC ? pc[jj + 1] : east;
                    const double z = fma(ca, qs[dd][jj], fma(cb, e, fma(mbeta, pc[jj], -qc[dd][jj])));
                    const double dd2 = fma(ca, pn[jj], fma(cb, w, z));
                    if (ok[jj]) {
                        viol |= !(fabs(dd2) <= tol);
                        w = pc[jj] + dd2;
                    } else {
                        w = pc[jj];        // boundary column: frozen, still the west operand of the next cell
                    }
                    out[jj] = w;
                    pn[jj] = w;
                    pc[jj] = qs[dd][jj];
                }
                store4<VEC>(P + q, nin, out);
                if (a.push_up && i0 + k == a.first_row) store4<VEC>(a.push_up + c0, nin, out);
                if (a.push_dn && i0 + k == a.last_row) store4<VEC>(a.push_dn + c0, nin, out);
                wlast = w;
            }
            prefetch(k + PF, dd);      // refill the slot just consumed with the row PF steps ahead
            __syncwarp();
        }
    }
    if (a.flags && __any_sync(0xffffffffu, viol) && lane == 0) atomicOr(&a.flags[s], 1);
}

int exchange_rows(nns_handle *h, SlabState *S, double *f, cudaStream_t st) {
    // swap boundary rows with the neighbouring ranks: local layout [nrows + 2][ny], halo rows first and last
    if (S->nranks == 1) return NNS_OK;
    NcclApi *N = nccl_api();
    const int ny = h->g.ny;
    double *top_halo = f, *first = f + ny, *last = f + (size_t)S->nrows * ny, *bot_halo = f + (size_t)(S->nrows + 1) * ny;
    NNS_NCCL(N->GroupStart());
    if (S->rank > 0) {
        NNS_NCCL(N->Send(first, ny, kNcclFloat64, S->rank - 1, S->comm, st));
        NNS_NCCL(N->Recv(top_halo, ny, kNcclFloat64, S->rank - 1, S->comm, st));
    }
    if (S->rank < S->nranks - 1) {
        NNS_NCCL(N->Send(last, ny, kNcclFloat64, S->rank + 1, S->comm, st));
        NNS_NCCL(N->Recv(bot_halo, ny, kNcclFloat64, S->rank + 1, S->comm, st));
    }
    NNS_NCCL(N->GroupEnd());
    return NNS_OK;
}

int apply_bc_list(nns_handle *h, const SlabGeom &g, int field, double *A, cudaStream_t st) {
    const BcList &L = h->bc[field];
    for (int k = 0; k < L.n; ++k) {
        const int n = (L.side[k] == NNS_SIDE_LEFT || L.side[k] == NNS_SIDE_RIGHT) ? g.ny : g.row1 - g.row0;
        slab_bc_kernel<<<(n + 127) / 128, 128, 0, st>>>(g, A, L.side[k], L.type[k] == NNS_BC_NEUMANN, L.value[k]);
        h->launches += 1;
    }
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

// mailbox layout of a rank: rows[dir][buf][ny] doubles (dir 0: written by the rank ABOVE, its last row; dir 1:
// written by the rank BELOW, its first row; buf = tick parity), then flags[dir] (uint32 sequence numbers)
__host__ __device__ inline size_t box_row_off(int dir, int buf, int ny) { return sizeof(double) * (size_t)(dir * 2 + buf) * ny; }
__host__ __device__ inline size_t box_flag_off(int dir, int ny) { return sizeof(double) * (size_t)4 * ny + sizeof(unsigned) * dir; }
inline size_t box_bytes(int ny) { return sizeof(double) * (size_t)4 * ny + 64; }

// One boundary row into the neighbour's mailbox over NVLink, then its flag (system-scope release).
__global__ void __launch_bounds__(1024) slab_push_row_kernel(const double *__restrict__ src, double *dst, int ny,
                                                             volatile unsigned *flag, unsigned seq) {
    for (int j = threadIdx.x; j < ny; j += blockDim.x) dst[j] = src[j];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = seq;
}

// Raise the neighbours' flags after the sweep kernel of a tick (stream order: its remote stores are complete).
__global__ void slab_flag_kernel(volatile unsigned *up, volatile unsigned *dn, unsigned seq) {
    __threadfence_system();
    if (up) *up = seq;
    if (dn) *dn = seq;
}

int launch_sweep(nns_handle *h, SlabState *S, SweepArgs a, int Ia, int Ib, int T, cudaStream_t st) {
    if (Ib <= Ia) return NNS_OK;
    const int dmin = T - 2 * (a.cap - 1), dmax = T;
    if (!(dmax >= Ia && dmin <= (Ib - 1) + (S->nJ - 1))) return NNS_OK;     // no diagonal crosses these tile rows
    a.I0 = Ia; a.I1 = Ib;
    const int ntl = std::min(Ib - Ia, S->nJ);
    const dim3 grid((ntl + 3) / 4, a.cap);
    if (a.g.ny % 4 == 0) slab_sweep_kernel<true><<<grid, 128, 0, st>>>(a, T);
    else slab_sweep_kernel<false><<<grid, 128, 0, st>>>(a, T);
    h->launches += 1;
    return NNS_OK;
}

int run_sweeps(nns_handle *h, SlabState *S, const SlabGeom &g, double *p, int cap, bool track, cudaStream_t st) {
    SweepArgs a{};
    a.g = g; a.TR = S->TR; a.nJ = S->nJ; a.I0 = S->I0; a.I1 = S->I1; a.cap = cap;
    a.p = p; a.cp = S->d_cprime; a.flags = track ? S->d_flags : nullptr;
    const int Tmax = (S->nI - 1) + (S->nJ - 1) + 2 * (cap - 1);
    const int ny = g.ny;
    int rc;
    if (!S->p2p) {
        for (int T = 0; T <= Tmax; ++T) {
            if ((rc = launch_sweep(h, S, a, S->I0, S->I1, T, st))) return rc;
            if ((rc = exchange_rows(h, S, p, st))) return rc;
        }
        NNS_CUDA(cudaGetLastError());
        return NNS_OK;
    }
    // Peer-memory exchange fused into the sweep kernel.  The mailbox holds ONE image of each neighbour's boundary
    // row; the tiles of a tick that update the slab's first / last row store their chunks into the neighbours'
    // mailboxes as well (NVLink stores), and the sweep kernel reads the rows just outside the slab from this
    // rank's mailbox.  Per tick, in stream order: wait until both neighbours have finished the previous tick
    // (stream memory operations on this rank's flags), sweep, raise the neighbours' flags.  A rank therefore runs
    // at most concurrently with its neighbours' SAME tick, and within one tick the row segments a rank writes
    // (tiles J with J = T - I - 2s) and the segments its neighbour reads have opposite parities of J: no chunk is
    // read and written in the same tick, so a single image per row suffices.  No NCCL kernel, no copy kernel and
    // no host synchronisation inside the tick loop.
    const bool up = S->rank > 0, dn = S->rank < S->nranks - 1;
    double *top_halo = p, *first = p + ny, *last = p + (size_t)S->nrows * ny, *bot_halo = p + (size_t)(S->nrows + 1) * ny;
    double *box_top = reinterpret_cast<double *>(S->d_box + box_row_off(0, 0, ny));
    double *box_bot = reinterpret_cast<double *>(S->d_box + box_row_off(1, 0, ny));
    volatile unsigned *flag_up = up ? reinterpret_cast<volatile unsigned *>(S->peer_box[0] + box_flag_off(1, ny)) : nullptr;
    volatile unsigned *flag_dn = dn ? reinterpret_cast<volatile unsigned *>(S->peer_box[1] + box_flag_off(0, ny)) : nullptr;
    static const bool nowait = getenv("NNS_SLAB_ABL_NOWAIT") != nullptr;     // timing ablation only: results are wrong
    auto wait_flags = [&](unsigned v) -> int {
        if (nowait) return NNS_OK;
        if (up && S->StreamWaitValue32((CUstream)st, (CUdeviceptr)(S->d_box + box_flag_off(0, ny)), v, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) { set_error("cuStreamWaitValue32 failed"); return NNS_ERR_CUDA; }
        if (dn && S->StreamWaitValue32((CUstream)st, (CUdeviceptr)(S->d_box + box_flag_off(1, ny)), v, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) { set_error("cuStreamWaitValue32 failed"); return NNS_ERR_CUDA; }
        return NNS_OK;
    };
    // start images: the whole boundary rows once.  The neighbours may still be reading their mailboxes for the
    // previous run of sweeps only if they have not finished it -- they have: every rank raises its last flag after
    // its last tick, and the waits below were passed.
    {
        const unsigned cur = ++S->seq;
        if (up) { slab_push_row_kernel<<<1, 1024, 0, st>>>(first, reinterpret_cast<double *>(S->peer_box[0] + box_row_off(1, 0, ny)), ny, flag_up, cur); h->launches += 1; }
        if (dn) { slab_push_row_kernel<<<1, 1024, 0, st>>>(last, reinterpret_cast<double *>(S->peer_box[1] + box_row_off(0, 0, ny)), ny, flag_dn, cur); h->launches += 1; }
    }
    a.halo_top = up ? box_top : nullptr;
    a.halo_bot = dn ? box_bot : nullptr;
    a.push_up = up ? reinterpret_cast<double *>(S->peer_box[0] + box_row_off(1, 0, ny)) : nullptr;
    a.push_dn = dn ? reinterpret_cast<double *>(S->peer_box[1] + box_row_off(0, 0, ny)) : nullptr;
    a.first_row = S->row0; a.last_row = S->row0 + S->nrows - 1;
    for (int T = 0; T <= Tmax; ++T) {
        if ((rc = wait_flags(S->seq))) return rc;          // the neighbours' rows of the previous tick (or the start images)
        if ((rc = launch_sweep(h, S, a, S->I0, S->I1, T, st))) return rc;
        const unsigned cur = ++S->seq;
        slab_flag_kernel<<<1, 1, 0, st>>>(flag_up, flag_dn, cur);
        h->launches += 1;
    }
    // the halo rows of the final state, for the kernels that follow
    if ((rc = wait_flags(S->seq))) return rc;
    if (up) NNS_CUDA(cudaMemcpyAsync(top_halo, box_top, sizeof(double) * ny, cudaMemcpyDeviceToDevice, st));
    if (dn) NNS_CUDA(cudaMemcpyAsync(bot_halo, box_bot, sizeof(double) * ny, cudaMemcpyDeviceToDevice, st));
    // nobody may overwrite this rank's mailbox (the start images of the next run) before these copies are done:
    // the neighbours wait for this rank's next flag, which is raised after them in stream order
    {
        const unsigned cur = ++S->seq;
        slab_flag_kernel<<<1, 1, 0, st>>>(flag_up, flag_dn, cur);
        h->launches += 1;
    }
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

}  // namespace

// ---- host-side plan (pure host logic, exported through the C ABI and unit-tested on the CPU) ----------
int slab_partition(int nx, int nranks, int rank, int tile_rows, int *row0, int *nrows, int *I0, int *I1, int *nI) {
    if (nx < 3 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("slab partition: bad argument"); return NNS_ERR_INVALID; }
    const int TR = tile_rows > 0 ? tile_rows : tile_rows_for(nx);
    const int n_tiles = (nx - 2 + TR - 1) / TR;
    if (n_tiles < nranks) {
        set_error("slab partition: %d interior rows give %d tile rows of %d, fewer than %d ranks", nx - 2, n_tiles, TR, nranks);
        return NNS_ERR_INVALID;
    }
    const int a = (int)((long long)n_tiles * rank / nranks), b = (int)((long long)n_tiles * (rank + 1) / nranks);
    const int r0 = rank == 0 ? 0 : 1 + a * TR;
    const int r1 = rank == nranks - 1 ? nx : std::min(1 + b * TR, nx - 1);
    if (row0) *row0 = r0;
    if (nrows) *nrows = r1 - r0;
    if (I0) *I0 = a;
    if (I1) *I1 = b;
    if (nI) *nI = n_tiles;
    return NNS_OK;
}

// plan query: tile geometry and, per (tick, sweep), the tile rows of a rank (host logic of the tick loop)
int slab_plan(int nx, int ny, int nranks, int rank, int tile_rows, int T, int s, int *out) {
    int row0, nrows, I0, I1, nI, rc;
    if ((rc = slab_partition(nx, nranks, rank, tile_rows, &row0, &nrows, &I0, &I1, &nI))) return rc;
    const int TR = tile_rows > 0 ? tile_rows : tile_rows_for(nx);
    const int nJ = (ny - 1 + TC - 1) / TC;
    int Ilo = 0, Ihi = -1;
    const bool any = tick_tile_rows(T, s, I0, I1, nJ, &Ilo, &Ihi);
    out[0] = TR; out[1] = TC; out[2] = nI; out[3] = nJ; out[4] = I0; out[5] = I1; out[6] = any ? Ilo : 0; out[7] = any ? Ihi : -1;
    return NNS_OK;
}

int slab_apply_bc(nns_handle *h, int field, double *a, cudaStream_t st) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) { set_error("slab path: call nns_slab_attach first"); return NNS_ERR_INVALID; }
    const Geometry &G = h->g;
    SlabGeom g{};
    g.nx = G.nx; g.ny = G.ny; g.row0 = S->row0; g.row1 = S->row0 + S->nrows;
    g.v.rowbase = S->row0 - 1; g.v.ny = G.ny;
    g.dt = G.dt; g.dx = G.dx; g.dy = G.dy; g.rho = G.rho; g.nu = G.nu; g.beta = G.beta; g.tol = G.tol;
    return apply_bc_list(h, g, field, a, st);
}

void slab_free(nns_handle *h) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) return;
    if (S->comm && nccl_api()) nccl_api()->CommDestroy(S->comm);
    cudaFree(S->d_cprime); cudaFree(S->d_p0); cudaFree(S->d_thomas); cudaFree(S->d_flags);
    for (double *d : S->d_own) cudaFree(d);
    for (int d = 0; d < 2; ++d) if (S->peer_box[d]) cudaIpcCloseMemHandle(S->peer_box[d]);
    cudaFree(S->d_box);
    if (S->h_flags) cudaFreeHost(S->h_flags);
    if (S->ev_edge) cudaEventDestroy(S->ev_edge);
    if (S->ev_xdone) cudaEventDestroy(S->ev_xdone);
    if (S->st_x) cudaStreamDestroy(S->st_x);
    if (S->ev[0]) cudaEventDestroy(S->ev[0]);
    if (S->ev[1]) cudaEventDestroy(S->ev[1]);
    delete S;
    h->slab = nullptr;
}

int slab_unique_id(unsigned char *id128) {
    NcclApi *N = nccl_api();
    if (!N) return NNS_ERR_CUDA;
    NcclId id;
    NNS_NCCL(N->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return NNS_OK;
}

int slab_attach(nns_handle *h, int rank, int nranks, const unsigned char *id128) {
    if (nranks > 1 && (h->g.batch != 1 || (h->params.solver == NNS_SOLVER_CHORIN_FD && h->g.method != NNS_METHOD_EXPLICIT))) {
        set_error("slab path: one grid over several ranks needs batch 1 and the explicit method (the ADI solves of "
                  "semi_implicit run along axis 0, across the slabs)");
        return NNS_ERR_UNSUPPORTED;
    }
    slab_free(h);
    SlabState *S = new SlabState();
    h->slab = S;
    S->rank = rank; S->nranks = nranks;
    S->TR = tile_rows_for(h->g.nx);
    int rc;
    if ((rc = slab_partition(h->g.nx, nranks, rank, S->TR, &S->row0, &S->nrows, &S->I0, &S->I1, &S->nI))) return rc;
    S->nJ = (h->g.ny - 1 + TC - 1) / TC;     // tile J covers the columns [J*TC, (J+1)*TC) up to ny-2
    if (nranks > 1) {
        NcclApi *N = nccl_api();
        if (!N) return NNS_ERR_CUDA;
        if (!id128) { set_error("slab attach: NCCL id missing"); return NNS_ERR_INVALID; }
        NcclId id;
        memcpy(id.internal, id128, 128);
        NNS_NCCL(N->CommInitRank(&S->comm, nranks, id, rank));
    }
    const size_t bytes = sizeof(double) * (size_t)(S->nrows + 2) * h->g.ny;
    NNS_CUDA(cudaMalloc(&S->d_cprime, bytes));
    NNS_CUDA(cudaMalloc(&S->d_p0, bytes));
    NNS_CUDA(cudaMemset(S->d_cprime, 0, bytes));
    NNS_CUDA(cudaMalloc(&S->d_flags, sizeof(int) * 64));
    NNS_CUDA(cudaMallocHost(&S->h_flags, sizeof(int) * 64));
    NNS_CUDA(cudaEventCreate(&S->ev[0]));
    NNS_CUDA(cudaEventCreate(&S->ev[1]));
    return NNS_OK;
}

// Peer-memory mailboxes: every rank allocates one, exports its 64-byte IPC handle, and maps the mailboxes of the
// ranks above and below (handles travel through torch.distributed).  With both mapped, the tick loop of the
// SOR exchanges its rows with NVLink stores + flags (run_sweeps) instead of NCCL launches.
int slab_ipc_export(nns_handle *h, unsigned char *handle64) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) { set_error("slab path: call nns_slab_attach first"); return NNS_ERR_INVALID; }
    if (!S->d_box) {
        NNS_CUDA(cudaMalloc(&S->d_box, box_bytes(h->g.ny)));
        NNS_CUDA(cudaMemset(S->d_box, 0, box_bytes(h->g.ny)));
    }
    cudaIpcMemHandle_t hd;
    NNS_CUDA(cudaIpcGetMemHandle(&hd, S->d_box));
    static_assert(sizeof(hd) == 64, "IPC handle size");
    memcpy(handle64, &hd, 64);
    return NNS_OK;
}

int slab_ipc_connect(nns_handle *h, const unsigned char *above64, const unsigned char *below64) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S || !S->d_box) { set_error("slab path: call nns_slab_ipc_export first"); return NNS_ERR_INVALID; }
    const unsigned char *hs[2] = {above64, below64};
    const bool need[2] = {S->rank > 0, S->rank < S->nranks - 1};
    for (int d = 0; d < 2; ++d) {
        if (!need[d]) continue;
        if (!hs[d]) { set_error("slab ipc connect: missing neighbour handle"); return NNS_ERR_INVALID; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, hs[d], 64);
        void *ptr = nullptr;
        NNS_CUDA(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        S->peer_box[d] = static_cast<unsigned char *>(ptr);
    }
    cudaDriverEntryPointQueryResult qr;
    void *fn = nullptr;
    NNS_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("cuStreamWaitValue32 is not available"); return NNS_ERR_UNSUPPORTED; }
    *reinterpret_cast<void **>(&S->StreamWaitValue32) = fn;
    S->p2p = true;
    return NNS_OK;
}

// device time of the tick loop (tracked sweeps) of the last step and its number of ticks (bench.py's roofline)
int slab_last_timing(nns_handle *h, float *sor_ms, int *ticks) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S || !S->last_ticks) { set_error("slab path: no step has run"); return NNS_ERR_INVALID; }
    NNS_CUDA(cudaEventSynchronize(S->ev[1]));
    NNS_CUDA(cudaEventElapsedTime(sor_ms, S->ev[0], S->ev[1]));
    *ticks = S->last_ticks;
    return NNS_OK;
}

int slab_exchange(nns_handle *h, double *f, cudaStream_t st) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) { set_error("slab path: call nns_slab_attach first"); return NNS_ERR_INVALID; }
    return exchange_rows(h, S, f, st);
}

// One time step on the local slabs ([nrows + 2][ny] each, halo rows valid on entry for u, v, u1, v1, p;
// valid on return for u_out, v_out, p).  Synchronises the stream once (exit-test decision).
int slab_step(nns_handle *h, const double *u, const double *v, const double *u1, const double *v1, double *p,
              double *un, double *vn, int32_t *sweeps_host, cudaStream_t st) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) { set_error("slab path: call nns_slab_attach first"); return NNS_ERR_INVALID; }
    const Geometry &G = h->g;
    if (G.nit - 1 > 64) { set_error("slab path: nit <= 65"); return NNS_ERR_UNSUPPORTED; }
    SlabGeom g{};
    g.nx = G.nx; g.ny = G.ny; g.row0 = S->row0; g.row1 = S->row0 + S->nrows;
    g.v.rowbase = S->row0 - 1; g.v.ny = G.ny;
    g.dt = G.dt; g.dx = G.dx; g.dy = G.dy; g.rho = G.rho; g.nu = G.nu; g.beta = G.beta; g.tol = G.tol;
    const dim3 blk(128), grd((G.ny + 127) / 128, S->nrows);
    const size_t bytes = sizeof(double) * (size_t)(S->nrows + 2) * G.ny;
    int rc;
    if (G.method == NNS_METHOD_SEMI_IMPLICIT) {
        // single GPU only (slab_attach refuses it for several ranks): whole columns are local
        if (!S->d_thomas) NNS_CUDA(cudaMalloc(&S->d_thomas, sizeof(double) * 4 * (size_t)G.nx));
        const double kx = 2.0 / G.nu * (G.dx * G.dx), ky = 2.0 / G.nu * (G.dy * G.dy);
        const dim3 tg((G.ny + 127) / 128, 2);
        semi_rhs_kernel<<<grd, blk, 0, st>>>(g, u, v, u1, v1, un, vn);
        thomas_coef_kernel<<<1, 1, 0, st>>>(G.nx, kx + 2.0 * G.dt, -G.dt, S->d_thomas);
        thomas_coef_kernel<<<1, 1, 0, st>>>(G.nx, ky + 2.0 * G.dt, -G.dt, S->d_thomas + 2 * G.nx);
        thomas_axis0_kernel<<<tg, 128, 0, st>>>(g, un, vn, -G.dt, S->d_thomas);
        semi_mid_kernel<<<grd, blk, 0, st>>>(g, u, v, un, vn);
        thomas_axis0_kernel<<<tg, 128, 0, st>>>(g, un, vn, -G.dt, S->d_thomas + 2 * G.nx);
        h->launches += 6;
    } else {
        slab_predictor_kernel<<<grd, blk, 0, st>>>(g, u, v, u1, v1, un, vn);
        h->launches += 1;
    }
    if ((rc = apply_bc_list(h, g, 0, un, st)) || (rc = apply_bc_list(h, g, 1, vn, st))) return rc;
    if ((rc = exchange_rows(h, S, un, st))) return rc;             // C' needs ui of the row above
    slab_cprime_kernel<<<grd, blk, 0, st>>>(g, un, vn, S->d_cprime);
    h->launches += 1;
    const int cap = G.nit - 1;
    int need = 0;
    if (cap > 0) {
        NNS_CUDA(cudaMemcpyAsync(S->d_p0, p, bytes, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemsetAsync(S->d_flags, 0, sizeof(int) * 64, st));
        NNS_CUDA(cudaEventRecord(S->ev[0], st));
        if ((rc = run_sweeps(h, S, g, p, cap, true, st))) return rc;
        NNS_CUDA(cudaEventRecord(S->ev[1], st));
        S->last_ticks = (S->nI - 1) + (S->nJ - 1) + 2 * (cap - 1) + 1;
        if (S->nranks > 1)
            NNS_NCCL(nccl_api()->AllReduce(S->d_flags, S->d_flags, 64, kNcclInt32, kNcclMax, S->comm, st));
        NNS_CUDA(cudaMemcpyAsync(S->h_flags, S->d_flags, sizeof(int) * 64, cudaMemcpyDeviceToHost, st));
        NNS_CUDA(cudaStreamSynchronize(st));
        need = cap;
        for (int s = 0; s < cap; ++s)
            if (!S->h_flags[s]) { need = s + 1; break; }      // first sweep with max|dp| <= tol: s + 1 sweeps run
        if (need < cap) {
            // the sequential loop would have stopped after `need` sweeps: redo from the saved p, capped
            NNS_CUDA(cudaMemcpyAsync(p, S->d_p0, bytes, cudaMemcpyDeviceToDevice, st));
            if ((rc = run_sweeps(h, S, g, p, need, false, st))) return rc;
        }
    }
    if (sweeps_host) *sweeps_host = need;
    if ((rc = apply_bc_list(h, g, 2, p, st))) return rc;
    if ((rc = exchange_rows(h, S, p, st))) return rc;
    slab_project_kernel<<<grd, blk, 0, st>>>(g, p, un, vn, h->d_nonfinite, h->params.flags & NNS_FLAG_CHECK_FINITE);
    h->launches += 1;
    if ((rc = exchange_rows(h, S, un, st)) || (rc = exchange_rows(h, S, vn, st))) return rc;
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

// Single-GPU run() for grids that do not fit the on-chip paths: plain [nx][ny] buffers from the caller are
// copied into haloed slabs owned by the handle (nranks = 1), stepped, and copied back.
// ---- direct_fd on row slabs (src/direct_fd/simulate.py:56-127; one grid over several GPUs) ------------------------------
// Jacobi has no ordering dependency: every sweep is one kernel on the owned rows, the p BCs, and one halo-row exchange
// with the neighbouring ranks (NCCL send / recv of single rows); u, v swap halo rows once per step.  Same expressions
// per cell as direct_fd.cu (axis 1 <-> dx, axis 0 <-> dy: the transpose of boundary.py, as in the reference).
__global__ void dslab_rhs_kernel(SlabGeom g, const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ b) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    const int ny = g.ny;
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
    const double kb = dx2 * dy2 / (2.0 * (dx2 + dy2));
    double bb = 0.0;
    if (i > 0 && i < g.nx - 1 && j > 0 && j < ny - 1) {
        const double r2dx = 1.0 / (2.0 * g.dx), r2dy = 1.0 / (2.0 * g.dy);
        const double ux = (u[q + 1] - u[q - 1]) * r2dx, vy = (v[q + ny] - v[q - ny]) * r2dy;
        const double uy = (u[q + ny] - u[q - ny]) * r2dy, vx = (v[q + 1] - v[q - 1]) * r2dx;
        bb = g.rho * ((1.0 / g.dt) * (ux + vy)) - ux * ux - 2.0 * (uy * vx) - vy * vy;
    }
    b[q] = kb * bb;
}

__global__ void dslab_jacobi_kernel(SlabGeom g, const double *__restrict__ pc, const double *__restrict__ bs, double *__restrict__ pn) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    const int ny = g.ny;
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
    const double rden = 1.0 / (2.0 * (dx2 + dy2));
    double r = pc[q];
    if (i > 0 && i < g.nx - 1 && j > 0 && j < ny - 1)
        r = (pc[q + 1] + pc[q - 1]) * (dy2 * rden) + (pc[q + ny] + pc[q - ny]) * (dx2 * rden) - bs[q];
    pn[q] = r;
}

// The Jacobi sweep with the p BC list applied in the same launch (the list walk above costs one launch per entry and
// sweep: at 8 GPUs the launches took longer than the sweeps).  Only the four corner cells of the grid depend on the
// order of the list (boundary.py:34-86 applied in list order): a cell of column 0 / ny-1 in another row is written by
// bottom / top entries only and reads the new value next to it, a non-corner cell of row 0 / nx-1 by left / right entries
// only and reads the new value of the adjacent row -- the LAST entry of that side decides, and the thread recomputes the
// neighbour's new value (same device function: same bits).  A corner thread replays the list in registers on the corner
// and its two neighbours (their values before the first entry are the previous sweep's: edge cells are copied through).
struct DirectBcPlan {
    int n;
    int side[NNS_MAX_BC], neu[NNS_MAX_BC];
    double val[NNS_MAX_BC];
    int kind[4];                 // per side (NNS_SIDE_*): 0 none, 1 Dirichlet, 2 Neumann -- the last entry of that side
    double g[4];
};

__device__ __forceinline__ double dslab_jac(const double *__restrict__ pc, const double *__restrict__ bs, size_t q, int ny,
                                            double cx, double cy) {
    return (pc[q + 1] + pc[q - 1]) * cx + (pc[q + ny] + pc[q - ny]) * cy - bs[q];
}

// rows: ra >= 0: the two rows ra, rb (blockIdx.y = 0 / 1: the slab's edge rows, swept first so that their exchange runs
// under the interior sweep); else the rows g.row0 + roff + blockIdx.y
__device__ __forceinline__ double dslab_cell(const SlabGeom &g, const DirectBcPlan &bc, const double *__restrict__ pc,
                                             const double *__restrict__ bs, int i, int j, double cx, double cy) {
    const size_t q = g.v.at(i, j);
    const int ny = g.ny, nx = g.nx;
    const bool rowe = i == 0 || i == nx - 1, cole = j == 0 || j == ny - 1;
    if (!rowe && !cole) return dslab_jac(pc, bs, q, ny, cx, cy);
    const int rside = i == 0 ? NNS_SIDE_LEFT : NNS_SIDE_RIGHT, cside = j == 0 ? NNS_SIDE_BOTTOM : NNS_SIDE_TOP;
    const int ii = i == 0 ? 1 : nx - 2, ji = j == 0 ? 1 : ny - 2;          // the adjacent interior row / column
    const double rsg = i == 0 ? -g.dx : g.dx, csg = j == 0 ? -g.dy : g.dy;
    double r = pc[q];
    if (cole && !rowe) {
        if (bc.kind[cside]) r = bc.kind[cside] == 2 ? dslab_jac(pc, bs, g.v.at(i, ji), ny, cx, cy) + csg * bc.g[cside] : bc.g[cside];
    } else if (rowe && !cole) {
        if (bc.kind[rside]) r = bc.kind[rside] == 2 ? dslab_jac(pc, bs, g.v.at(ii, j), ny, cx, cy) + rsg * bc.g[rside] : bc.g[rside];
    } else {
        const double a11 = dslab_jac(pc, bs, g.v.at(ii, ji), ny, cx, cy);
        double t_adj = pc[g.v.at(ii, j)], t_row = pc[g.v.at(i, ji)];
        for (int k = 0; k < bc.n; ++k) {
            const double gv = bc.val[k];
            if (bc.side[k] == rside) { t_row = bc.neu[k] ? a11 + rsg * gv : gv; r = bc.neu[k] ? t_adj + rsg * gv : gv; }
            else if (bc.side[k] == cside) { t_adj = bc.neu[k] ? a11 + csg * gv : gv; r = bc.neu[k] ? t_row + csg * gv : gv; }
        }
    }
    return r;
}

// VEC: a thread updates the two cells j = 2t, 2t + 1 with 128-bit loads / stores (ny even: rows are 16-byte aligned)
template <bool VEC>
__global__ void __launch_bounds__(128) dslab_jacobi_bc_kernel(SlabGeom g, DirectBcPlan bc, const double *__restrict__ pc,
                                                              const double *__restrict__ bs, double *__restrict__ pn, int ra, int rb,
                                                              int roff, int rend) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, j = VEC ? 2 * t : t;
    const int i = ra >= 0 ? (blockIdx.y == 0 ? ra : rb) : g.row0 + roff + (int)blockIdx.y;
    if (j >= g.ny || i >= (ra >= 0 ? g.row1 : rend)) return;
    const int ny = g.ny, nx = g.nx;
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
    const double rden = 1.0 / (2.0 * (dx2 + dy2));
    const double cx = dy2 * rden, cy = dx2 * rden;
    if (VEC) {
        const size_t q = g.v.at(i, j);
        if (i > 0 && i < nx - 1 && j >= 2 && j + 1 <= ny - 2) {          // both cells interior: the same expression as dslab_jac
            const double2 C = *reinterpret_cast<const double2 *>(pc + q), Nn = *reinterpret_cast<const double2 *>(pc + q - ny);
            const double2 S = *reinterpret_cast<const double2 *>(pc + q + ny), B = *reinterpret_cast<const double2 *>(bs + q);
            const double W = pc[q - 1], E = pc[q + 2];
            double2 r;
            r.x = (C.y + W) * cx + (S.x + Nn.x) * cy - B.x;
            r.y = (E + C.x) * cx + (S.y + Nn.y) * cy - B.y;
            *reinterpret_cast<double2 *>(pn + q) = r;
        } else {
            pn[q] = dslab_cell(g, bc, pc, bs, i, j, cx, cy);
            if (j + 1 < ny) pn[q + 1] = dslab_cell(g, bc, pc, bs, i, j + 1, cx, cy);
        }
    } else {
        pn[g.v.at(i, j)] = dslab_cell(g, bc, pc, bs, i, j, cx, cy);
    }
}

__global__ void dslab_update_kernel(SlabGeom g, const double *__restrict__ uo, const double *__restrict__ vo,
                                    const double *__restrict__ p, double *__restrict__ un, double *__restrict__ vn,
                                    unsigned long long *nonfinite, int check) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = g.row0 + blockIdx.y;
    if (j >= g.ny || i >= g.row1) return;
    const size_t q = g.v.at(i, j);
    const int ny = g.ny;
    const double dt = g.dt, dx = g.dx, dy = g.dy, rho = g.rho, nu = g.nu;
    const double uc = uo[q], vc = vo[q];
    double ru = uc, rv = vc;
    if (i > 0 && i < g.nx - 1 && j > 0 && j < ny - 1) {
        const double kpx = dt / (2.0 * rho * dx), kpy = dt / (2.0 * rho * dy);
        const double kdx = dt / (dx * dx), kdy = dt / (dy * dy), ax = dt / dx, ay = dt / dy;
        const double uW = uo[q - 1], uE = uo[q + 1], uN = uo[q - ny], uS = uo[q + ny];
        const double vW = vo[q - 1], vE = vo[q + 1], vN = vo[q - ny], vS = vo[q + ny];
        ru = uc - uc * ax * (uc - uW) - vc * ay * (uc - uN) - kpx * (p[q + 1] - p[q - 1]) +
             nu * (kdx * (uE - 2.0 * uc + uW) + kdy * (uS - 2.0 * uc + uN));
        rv = vc - uc * ax * (vc - vW) - vc * ay * (vc - vN) - kpy * (p[q + ny] - p[q - ny]) +
             nu * (kdx * (vE - 2.0 * vc + vW) + kdy * (vS - 2.0 * vc + vN));
    }
    un[q] = ru;
    vn[q] = rv;
    if (check && !(isfinite(ru) && isfinite(rv))) atomicAdd(nonfinite, 1ull);
}

// nsteps of step() on the local slabs ([nrows + 2][ny] each; halo rows of u, v, p valid on entry and on return).
int direct_slab_run(nns_handle *h, double *u, double *v, double *p, int nsteps, cudaStream_t st) {
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (!S) { set_error("slab path: call nns_slab_attach first"); return NNS_ERR_INVALID; }
    const Geometry &G = h->g;
    SlabGeom g{};
    g.nx = G.nx; g.ny = G.ny; g.row0 = S->row0; g.row1 = S->row0 + S->nrows;
    g.v.rowbase = S->row0 - 1; g.v.ny = G.ny;
    g.dt = G.dt; g.dx = G.dx; g.dy = G.dy; g.rho = G.rho; g.nu = G.nu; g.beta = G.beta; g.tol = G.tol;
    const dim3 blk(128), grd((G.ny + 127) / 128, S->nrows);
    const size_t bytes = sizeof(double) * (size_t)(S->nrows + 2) * G.ny;
    for (int k = 0; k < 4; ++k)
        if (!S->d_own[k]) { NNS_CUDA(cudaMalloc(&S->d_own[k], bytes)); NNS_CUDA(cudaMemsetAsync(S->d_own[k], 0, bytes, st)); }
    double *uc = u, *vc = v, *un = S->d_own[0], *vn = S->d_own[1], *pn = S->d_own[2], *b = S->d_own[3], *pc = p;
    int rc;
    // p BCs inside the sweep kernel: the rank that owns a global edge row needs the adjacent interior row too (it recomputes
    // that row's new value); NNS_DSLAB_BC=list keeps the list walk (one launch per entry) for tests
    DirectBcPlan plan{};
    {
        const BcList &L = h->bc[2];
        plan.n = L.n;
        for (int k = 0; k < L.n; ++k) {
            plan.side[k] = L.side[k]; plan.neu[k] = L.type[k] == NNS_BC_NEUMANN; plan.val[k] = L.value[k];
            plan.kind[L.side[k]] = plan.neu[k] ? 2 : 1; plan.g[L.side[k]] = L.value[k];
        }
    }
    const bool vec = G.ny % 2 == 0 && !getenv("NNS_DSLAB_SCALAR");        // two cells per thread, 128-bit accesses
    const unsigned grdv = (unsigned)((G.ny / 2 + 127) / 128);
    const char *bcm = getenv("NNS_DSLAB_BC");
    const bool owns_top = S->row0 == 0, owns_bot = S->row0 + S->nrows == G.nx;
    const bool fused_bc = !(bcm && strcmp(bcm, "list") == 0) && G.nx >= 4 && G.ny >= 3 &&
                          (!owns_top || S->nrows >= 2) && (!owns_bot || S->nrows >= 2);
    // several ranks: sweep the slab's first and last row first and exchange them (second stream) under the interior sweep
    const char *ovl = getenv("NNS_DSLAB_OVERLAP");
    const bool overlap = fused_bc && S->nranks > 1 && S->nrows >= 4 && (ovl && atoi(ovl) == 1);      // opt-in until measured on a multi-GPU box
    if (overlap && !S->st_x) {
        NNS_CUDA(cudaStreamCreateWithFlags(&S->st_x, cudaStreamNonBlocking));
        NNS_CUDA(cudaEventCreateWithFlags(&S->ev_edge, cudaEventDisableTiming));
        NNS_CUDA(cudaEventCreateWithFlags(&S->ev_xdone, cudaEventDisableTiming));
    }
    for (int n = 0; n < nsteps; ++n) {
        dslab_rhs_kernel<<<grd, blk, 0, st>>>(g, uc, vc, b);
        for (int s = 0; s < G.nit; ++s) {
            if (fused_bc && overlap) {
                // edge rows, their exchange on the second stream, the interior rows meanwhile on the first
                if (vec) dslab_jacobi_bc_kernel<true><<<dim3(grdv, 2), blk, 0, st>>>(g, plan, pc, b, pn, g.row0, g.row1 - 1, 0, 0);
                else dslab_jacobi_bc_kernel<false><<<dim3(grd.x, 2), blk, 0, st>>>(g, plan, pc, b, pn, g.row0, g.row1 - 1, 0, 0);
                NNS_CUDA(cudaEventRecord(S->ev_edge, st));
                NNS_CUDA(cudaStreamWaitEvent(S->st_x, S->ev_edge, 0));
                if ((rc = exchange_rows(h, S, pn, S->st_x))) return rc;
                NNS_CUDA(cudaEventRecord(S->ev_xdone, S->st_x));
                if (vec) dslab_jacobi_bc_kernel<true><<<dim3(grdv, S->nrows - 2), blk, 0, st>>>(g, plan, pc, b, pn, -1, -1, 1, g.row1 - 1);
                else dslab_jacobi_bc_kernel<false><<<dim3(grd.x, S->nrows - 2), blk, 0, st>>>(g, plan, pc, b, pn, -1, -1, 1, g.row1 - 1);
                NNS_CUDA(cudaStreamWaitEvent(st, S->ev_xdone, 0));
                h->launches += 2;
                double *t = pc; pc = pn; pn = t;
                continue;
            }
            if (fused_bc) {
                if (vec) dslab_jacobi_bc_kernel<true><<<dim3(grdv, grd.y), blk, 0, st>>>(g, plan, pc, b, pn, -1, -1, 0, g.row1);
                else dslab_jacobi_bc_kernel<false><<<grd, blk, 0, st>>>(g, plan, pc, b, pn, -1, -1, 0, g.row1);
                h->launches += 1;
            } else {
                dslab_jacobi_kernel<<<grd, blk, 0, st>>>(g, pc, b, pn);
                h->launches += 1;
                if ((rc = apply_bc_list(h, g, 2, pn, st))) return rc;
            }
            if ((rc = exchange_rows(h, S, pn, st))) return rc;
            double *t = pc; pc = pn; pn = t;
        }
        dslab_update_kernel<<<grd, blk, 0, st>>>(g, uc, vc, pc, un, vn, h->d_nonfinite, h->params.flags & NNS_FLAG_CHECK_FINITE);
        h->launches += 2;
        if ((rc = apply_bc_list(h, g, 0, un, st)) || (rc = apply_bc_list(h, g, 1, vn, st))) return rc;
        if (overlap) {       // one communicator, one stream for its operations
            NNS_CUDA(cudaEventRecord(S->ev_edge, st));
            NNS_CUDA(cudaStreamWaitEvent(S->st_x, S->ev_edge, 0));
            if ((rc = exchange_rows(h, S, un, S->st_x)) || (rc = exchange_rows(h, S, vn, S->st_x))) return rc;
            NNS_CUDA(cudaEventRecord(S->ev_xdone, S->st_x));
            NNS_CUDA(cudaStreamWaitEvent(st, S->ev_xdone, 0));
        } else if ((rc = exchange_rows(h, S, un, st)) || (rc = exchange_rows(h, S, vn, st))) return rc;
        double *t;
        t = uc; uc = un; un = t;
        t = vc; vc = vn; vn = t;
    }
    NNS_CUDA(cudaGetLastError());
    if (pc != p) NNS_CUDA(cudaMemcpyAsync(p, pc, bytes, cudaMemcpyDeviceToDevice, st));
    if (uc != u) {
        NNS_CUDA(cudaMemcpyAsync(u, uc, bytes, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(v, vc, bytes, cudaMemcpyDeviceToDevice, st));
    }
    return NNS_OK;
}

static int tiled_run_member(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                            int step0, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps, int sweeps_stride,
                            cudaStream_t st);

// Single-GPU run() for grids that do not fit the on-chip paths: plain [nx][ny] buffers from the caller are copied into
// haloed slabs owned by the handle (nranks = 1), stepped, and copied back.  The members of a batch are advanced one
// after the other (each with its own nu / BC values): a member of this size fills the GPU by itself.
int chorin_tiled_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                     int step0, int phases, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps,
                     cudaStream_t st, int m0, int count) {
    if (phases != 7) {
        set_error("chorin_fd: grid %dx%d does not fit the on-chip path; the tiled path runs whole steps only (the stage "
                  "entry points are for unit parity on small grids)", h->g.nx, h->g.ny);
        return NNS_ERR_UNSUPPORTED;
    }
    if (count < 0) count = h->g.batch - m0;
    const size_t N = (size_t)h->g.nx * h->g.ny;
    const double nu0 = h->g.nu;
    BcList bc0[3] = {h->bc[0], h->bc[1], h->bc[2]};
    int rc = NNS_OK;
    for (int mm = 0; mm < count && rc == NNS_OK; ++mm) {
        const int m = m0 + mm;
        if (h->h_nu) h->g.nu = h->h_nu[m];
        if (h->h_bcval)
            for (int f = 0; f < 3; ++f)
                for (int k = 0; k < h->bc[f].n; ++k) h->bc[f].value[k] = h->h_bcval[(size_t)m * h->n_bcs + h->bc[f].slot[k]];
        double *U[3] = {bufU[0] + mm * N, bufU[1] + mm * N, bufU[2] + mm * N};
        double *V[3] = {bufV[0] + mm * N, bufV[1] + mm * N, bufV[2] + mm * N};
        rc = tiled_run_member(h, U, V, p + mm * N, nsteps, nsteps_total, step0, fixup,
                              tu ? tu + (size_t)mm * nsteps_total * N : nullptr, tv ? tv + (size_t)mm * nsteps_total * N : nullptr,
                              tp ? tp + (size_t)mm * nsteps_total * N : nullptr, sweeps ? sweeps + mm : nullptr, h->g.batch, st);
    }
    h->g.nu = nu0;
    for (int f = 0; f < 3; ++f) h->bc[f] = bc0[f];
    return rc;
}

static int tiled_run_member(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                            int step0, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps, int sweeps_stride,
                            cudaStream_t st) {
    int rc;
    if (!h->slab && (rc = slab_attach(h, 0, 1, nullptr))) return rc;
    SlabState *S = static_cast<SlabState *>(h->slab);
    if (S->nranks != 1) { set_error("chorin_fd: handle is attached to a multi-rank slab; use nns_chorin_fd_slab_step"); return NNS_ERR_INVALID; }
    const size_t N = (size_t)h->g.nx * h->g.ny, nb = sizeof(double) * N, hb = sizeof(double) * (N + 2 * h->g.ny);
    for (int k = 0; k < 7; ++k)
        if (!S->d_own[k]) { NNS_CUDA(cudaMalloc(&S->d_own[k], hb)); NNS_CUDA(cudaMemsetAsync(S->d_own[k], 0, hb, st)); }
    double *U[3] = {S->d_own[0], S->d_own[1], S->d_own[2]}, *V[3] = {S->d_own[3], S->d_own[4], S->d_own[5]}, *Ps = S->d_own[6];
    const int ny = h->g.ny;
    for (int k = 0; k < 2; ++k) {
        NNS_CUDA(cudaMemcpyAsync(U[k] + ny, bufU[k], nb, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(V[k] + ny, bufV[k], nb, cudaMemcpyDeviceToDevice, st));
    }
    NNS_CUDA(cudaMemcpyAsync(Ps + ny, p, nb, cudaMemcpyDeviceToDevice, st));
    int cur = 0, prev = 1, nxt = 2;
    for (int n = 0; n < nsteps; ++n) {
        int32_t need = 0;
        if ((rc = slab_step(h, U[cur], V[cur], U[prev], V[prev], Ps, U[nxt], V[nxt], &need, st))) return rc;
        if (sweeps) NNS_CUDA(cudaMemcpyAsync(sweeps + (size_t)(step0 + n) * sweeps_stride, &need, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if (tu) {
            const size_t off = (size_t)(step0 + n) * N;
            NNS_CUDA(cudaMemcpyAsync(tu + off, U[nxt] + ny, nb, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(tv + off, V[nxt] + ny, nb, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(tp + off, Ps + ny, nb, cudaMemcpyDeviceToDevice, st));
        }
        NNS_CUDA(cudaStreamSynchronize(st));      // `need` is a host temporary
        const int t = prev; prev = cur; cur = nxt; nxt = t;
    }
    (void)nsteps_total;
    // results back into the caller's buffers with the roles the caller expects
    if (fixup) {
        NNS_CUDA(cudaMemcpyAsync(bufU[0], U[cur] + ny, nb, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(bufV[0], V[cur] + ny, nb, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(bufU[1], U[prev] + ny, nb, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(bufV[1], V[prev] + ny, nb, cudaMemcpyDeviceToDevice, st));
    } else {                                       // single step: new state into buffer 2
        NNS_CUDA(cudaMemcpyAsync(bufU[2], U[cur] + ny, nb, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(bufV[2], V[cur] + ny, nb, cudaMemcpyDeviceToDevice, st));
    }
    NNS_CUDA(cudaMemcpyAsync(p, Ps + ny, nb, cudaMemcpyDeviceToDevice, st));
    return NNS_OK;
}

}  // namespace nns
