// direct_fd.cu -- direct finite-difference Navier-Stokes step (Jacobi pressure Poisson).
//
// Reference semantics reproduced (src/direct_fd/simulate.py of mhw32/neural-navier-stokes):
//   _build_up_b :56-66        RHS b (axis 1 <-> dx, axis 0 <-> dy: the transpose of boundary.py)
//   _pressure_poisson :68-88  exactly nit Jacobi sweeps, p BCs re-applied after EVERY sweep
//   step :90-127              upwind advection, central diffusion, -grad p / rho, then u/v BCs
//
// Two paths:
//   * "chip": one CTA per member, p (ping-pong) and b resident in shared memory, all nit
//     sweeps + BCs + velocity update fused in one launch per run of nsteps (small grids,
//     ensembles).
//   * "cluster": one thread-block CLUSTER per member (2 .. 16 CTAs): every CTA keeps a band of rows of p (ping-pong)
//     and b in its shared memory for the whole run, the band edges travel to the neighbouring CTAs through
//     distributed shared memory after every sweep (one cluster barrier per sweep), RHS and velocity update are fused
//     around the sweeps and all nsteps run in ONE launch: the Jacobi sweeps are temporally blocked on chip, HBM only
//     sees u, v once per step (256 x 256: ~103 launches and 2 x 50 passes over p per step before).
//   * "stream": fields in HBM, one launch per sweep (any grid size).
#include <cooperative_groups.h>

#include "nns_common.cuh"

namespace cg = cooperative_groups;

namespace nns {

struct DirectArgs {
    Geometry g;
    BcList ubc, vbc, pbc;
    const double *nu_b;
    const double *bcval;
    int n_bcs;
    int nsteps, nsteps_total, step0;
    int flags;
    double force_x;             // periodic-x extension: source of u
    double *u, *v, *p;          // [batch][nx][ny], advanced in place
    double *su, *sv;            // scratch (ping-pong partner of u, v)
    double *traj_u, *traj_v, *traj_p;
    unsigned long long *nonfinite;
};

__device__ __forceinline__ void cta_apply_bc_smem(double *A, int nx, int ny, int pitch, const BcList &L,
                                                  const double *bcval, double dx, double dy) {
    for (int k = 0; k < L.n; ++k) {
        const double g = bcval ? bcval[L.slot[k]] : L.value[k];
        const int side = L.side[k];
        const bool neu = L.type[k] == NNS_BC_NEUMANN;
        if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
            const int i = side == NNS_SIDE_LEFT ? 0 : nx - 1, in = side == NNS_SIDE_LEFT ? 1 : nx - 2;
            const double sg = side == NNS_SIDE_LEFT ? -dx : dx;
            for (int j = threadIdx.x; j < ny; j += blockDim.x)
                A[i * pitch + j] = neu ? A[in * pitch + j] + sg * g : g;
        } else {
            const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
            const double sg = side == NNS_SIDE_BOTTOM ? -dy : dy;
            for (int i = threadIdx.x; i < nx; i += blockDim.x)
                A[i * pitch + j] = neu ? A[i * pitch + jn] + sg * g : g;
        }
        __syncthreads();
    }
}

// The same walk with the list staged in shared memory (vals[k], code[k] = side | type << 8): inside the sweep loop a
// per-member value is otherwise a global load, the side / type an indexed parameter load, in front of every entry of every sweep.
__device__ __forceinline__ void cta_apply_bc_smem_staged(double *A, int nx, int ny, int pitch, int nent, const double *vals,
                                                         const int *code, double dx, double dy) {
    for (int k = 0; k < nent; ++k) {
        const double g = vals[k];
        const int side = code[k] & 0xff;
        const bool neu = (code[k] >> 8) == NNS_BC_NEUMANN;
        if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
            const int i = side == NNS_SIDE_LEFT ? 0 : nx - 1, in = side == NNS_SIDE_LEFT ? 1 : nx - 2;
            const double sg = side == NNS_SIDE_LEFT ? -dx : dx;
            for (int j = threadIdx.x; j < ny; j += blockDim.x)
                A[i * pitch + j] = neu ? A[in * pitch + j] + sg * g : g;
        } else {
            const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
            const double sg = side == NNS_SIDE_BOTTOM ? -dy : dy;
            for (int i = threadIdx.x; i < nx; i += blockDim.x)
                A[i * pitch + j] = neu ? A[i * pitch + jn] + sg * g : g;
        }
        __syncthreads();
    }
}

// ---- chip path -------------------------------------------------------------------------
// smem: P0, P1 (ping-pong), Bs: 3 * nx * pitch doubles.
// (NTMAX: the launch's thread count bounds the registers per thread -- 64 at 1024 threads, where ptxas re-reads loop
// invariants from the constant bank in every iteration; smaller grids run with 512 or 256 threads and 128 / 255 registers)
template <int NTMAX>
__global__ void __launch_bounds__(NTMAX, 1) direct_chip_kernel(const DirectArgs a) {
    extern __shared__ double smem[];
    const int nx = a.g.nx, ny = a.g.ny, pitch = ny | 1;
    const size_t N = (size_t)nx * ny;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    double *P0 = smem, *P1 = smem + nx * pitch, *Bs = smem + 2 * nx * pitch;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy, rho = a.g.rho;
    const double nu = a.nu_b ? a.nu_b[b] : a.g.nu;
    const double *bcval = a.bcval ? a.bcval + (size_t)b * a.n_bcs : nullptr;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double rden = 1.0 / (2.0 * (dx2 + dy2));
    const double cx = dy2 * rden, cy = dx2 * rden, kb = dx2 * dy2 * rden;
    const double r2dx = 1.0 / (2.0 * dx), r2dy = 1.0 / (2.0 * dy), rdt = 1.0 / dt;
    double *ug = a.u + (size_t)b * N, *vg = a.v + (size_t)b * N, *pg = a.p + (size_t)b * N;
    double *us = a.su + (size_t)b * N, *vs = a.sv + (size_t)b * N;
    // periodic-x extension: every column is interior, the column neighbours wrap around
    const bool per = a.flags & NNS_FLAG_PERIODIC_X;
    const double fdt = per ? a.force_x * dt : 0.0;

    __shared__ double s_pval[NNS_MAX_BC];
    __shared__ int s_pcode[NNS_MAX_BC];
    if (tid < a.pbc.n) { s_pval[tid] = bcval ? bcval[a.pbc.slot[tid]] : a.pbc.value[tid]; s_pcode[tid] = a.pbc.side[tid] | (a.pbc.type[tid] << 8); }
    for (int i = warp; i < nx; i += nwarps)
        for (int j = lane; j < ny; j += 32) P0[i * pitch + j] = pg[(size_t)i * ny + j];
    __syncthreads();

    for (int n = 0; n < a.nsteps; ++n) {
        const double *uo = (n & 1) ? us : ug, *vo = (n & 1) ? vs : vg;   // u^n, v^n
        double *un = (n & 1) ? ug : us, *vn = (n & 1) ? vg : vs;         // u^{n+1}
        // RHS b (direct_fd:56-66)
        for (int i = warp; i < nx; i += nwarps)
            for (int j = lane; j < ny; j += 32) {
                double bb = 0.0;
                if (i > 0 && i < nx - 1 && (per || (j > 0 && j < ny - 1))) {
                    const size_t q = (size_t)i * ny + j, rq = (size_t)i * ny;
                    const int jp = j < ny - 1 ? j + 1 : 0, jm = j > 0 ? j - 1 : ny - 1;
                    const double ux = (uo[rq + jp] - uo[rq + jm]) * r2dx, vy = (vo[q + ny] - vo[q - ny]) * r2dy;
                    const double uy = (uo[q + ny] - uo[q - ny]) * r2dy, vx = (vo[rq + jp] - vo[rq + jm]) * r2dx;
                    bb = rho * (rdt * (ux + vy)) - ux * ux - 2.0 * (uy * vx) - vy * vy;
                }
                Bs[i * pitch + j] = kb * bb;
            }
        __syncthreads();
        // exactly nit Jacobi sweeps, BCs after every sweep (direct_fd:76-86)
        double *Pc = P0, *Pn = P1;
        for (int s = 0; s < a.g.nit; ++s) {
            for (int i = warp; i < nx; i += nwarps)
                for (int j = lane; j < ny; j += 32) {
                    const int q = i * pitch + j;
                    double r = Pc[q];
                    if (i > 0 && i < nx - 1 && (per || (j > 0 && j < ny - 1))) {
                        const int jp = j < ny - 1 ? j + 1 : 0, jm = j > 0 ? j - 1 : ny - 1;
                        r = (Pc[i * pitch + jp] + Pc[i * pitch + jm]) * cx + (Pc[q + pitch] + Pc[q - pitch]) * cy - Bs[q];
                    }
                    Pn[q] = r;
                }
            __syncthreads();
            cta_apply_bc_smem_staged(Pn, nx, ny, pitch, a.pbc.n, s_pval, s_pcode, dx, dy);
            double *t = Pc; Pc = Pn; Pn = t;
        }
        // velocity update (direct_fd:98-118) into the partner buffers, then u/v BCs (:121-125)
        const double kpx = dt / (2.0 * rho * dx), kpy = dt / (2.0 * rho * dy);
        const double kdx = dt / dx2, kdy = dt / dy2, ax = dt / dx, ay = dt / dy;
        for (int i = warp; i < nx; i += nwarps)
            for (int j = lane; j < ny; j += 32) {
                const size_t q = (size_t)i * ny + j;
                const double uc = uo[q], vc = vo[q];
                double ru = uc, rv = vc;
                if (i > 0 && i < nx - 1 && (per || (j > 0 && j < ny - 1))) {
                    const int s = i * pitch + j;
                    const int jp = j < ny - 1 ? j + 1 : 0, jm = j > 0 ? j - 1 : ny - 1;
                    const size_t rq = (size_t)i * ny;
                    const double uW = uo[rq + jm], uE = uo[rq + jp], uN = uo[q - ny], uS = uo[q + ny];
                    const double vW = vo[rq + jm], vE = vo[rq + jp], vN = vo[q - ny], vS = vo[q + ny];
                    ru = uc - uc * ax * (uc - uW) - vc * ay * (uc - uN) - kpx * (Pc[i * pitch + jp] - Pc[i * pitch + jm]) +
                         nu * (kdx * (uE - 2.0 * uc + uW) + kdy * (uS - 2.0 * uc + uN)) + fdt;
                    rv = vc - uc * ax * (vc - vW) - vc * ay * (vc - vN) - kpy * (Pc[s + pitch] - Pc[s - pitch]) +
                         nu * (kdx * (vE - 2.0 * vc + vW) + kdy * (vS - 2.0 * vc + vN));
                }
                un[q] = ru;
                vn[q] = rv;
            }
        __syncthreads();
        cta_apply_bc_global(un, nx, ny, a.ubc, bcval, dx, dy);
        cta_apply_bc_global(vn, nx, ny, a.vbc, bcval, dx, dy);
        if (a.traj_u || (a.flags & NNS_FLAG_CHECK_FINITE)) {
            const size_t toff = ((size_t)b * a.nsteps_total + (a.step0 + n)) * N;
            unsigned long long bad = 0;
            for (int i = warp; i < nx; i += nwarps)
                for (int j = lane; j < ny; j += 32) {
                    const size_t q = (size_t)i * ny + j;
                    const double x = un[q], y = vn[q], z = Pc[i * pitch + j];
                    if (a.traj_u) { a.traj_u[toff + q] = x; a.traj_v[toff + q] = y; a.traj_p[toff + q] = z; }
                    bad += !(isfinite(x) && isfinite(y) && isfinite(z));
                }
            if ((a.flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(a.nonfinite, bad);
        }
        if (Pc != P0) {     // keep the current pressure in P0 for the next step
            for (int q = tid; q < nx * pitch; q += blockDim.x) P0[q] = Pc[q];
        }
        __syncthreads();
    }
    for (int i = warp; i < nx; i += nwarps)
        for (int j = lane; j < ny; j += 32) pg[(size_t)i * ny + j] = P0[i * pitch + j];
    if (a.nsteps & 1)
        for (size_t q = tid; q < N; q += blockDim.x) { ug[q] = us[q]; vg[q] = vs[q]; }
}

// ---- cluster path -----------------------------------------------------------------------
// Rows [i0, i1) of the member belong to this CTA; shared-memory row index li = i - i0 + 1 (li = 0 and li = nloc + 1
// are the halo rows owned by the neighbouring CTAs).  smem: P0, P1, Bs of (band + 2) * pitch doubles each.
__device__ __forceinline__ void band_apply_bc_global(double *A, int nx, int ny, int i0, int i1, const BcList &L,
                                                     const double *bcval, double dx, double dy) {
    for (int k = 0; k < L.n; ++k) {
        const double g = bcval ? bcval[L.slot[k]] : L.value[k];
        const int side = L.side[k];
        const bool neu = L.type[k] == NNS_BC_NEUMANN;
        if (side == NNS_SIDE_LEFT) {
            if (i0 == 0)
                for (int j = threadIdx.x; j < ny; j += blockDim.x) A[j] = neu ? A[(size_t)ny + j] - dx * g : g;
        } else if (side == NNS_SIDE_RIGHT) {
            if (i1 == nx)
                for (int j = threadIdx.x; j < ny; j += blockDim.x)
                    A[(size_t)(nx - 1) * ny + j] = neu ? A[(size_t)(nx - 2) * ny + j] + dx * g : g;
        } else {
            const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
            const double sg = side == NNS_SIDE_BOTTOM ? -dy : dy;
            for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) A[(size_t)i * ny + j] = neu ? A[(size_t)i * ny + jn] + sg * g : g;
        }
        __syncthreads();
    }
}

// Point-to-point hand-off between neighbouring CTAs of the cluster: the producer's threads store the band edge into the
// consumer's shared memory with st.async, which completes on the consumer's mbarrier; the consumer's warps wait on their
// own mbarrier.  A full cluster barrier per sweep costs several times more.
__device__ __forceinline__ uint32_t dsm_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, unsigned rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(rank));
    return ra;
}
// Asynchronous store into a neighbouring CTA's shared memory that completes (transaction bytes) on an mbarrier of that
// CTA: the consumer that sees the phase complete sees the data -- no fence on either side.  (A release arrive or a fence at
// cluster scope compiles to MEMBAR.ALL.GPU + CCTL.IVALL: three of those per warp and sweep were half of the sweep time.)
#ifdef NNS_CLUSTER_PLAIN_ST
// EXPERIMENT: plain stores through distributed shared memory; after a CTA barrier ONE thread issues fence.release.cluster
// (MEMBAR.ALL.GPU, once per sweep and CTA instead of three times per warp as in the first version) and a relaxed arrive (a
// bare SYNCS.ARRIVE) on each neighbour's mbarrier.
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_async_f64x2(uint32_t raddr, double v0, double v1, uint32_t) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ void mbar_arm(const void *, uint32_t) {}
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t rbar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
#else
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_f64x2(uint32_t raddr, double v0, double v1, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "d"(v0), "d"(v1), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_arm(const void *bar, uint32_t bytes) {      // the one expected arrival + the bytes of the next phase
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dsm_u32(bar)), "r"(bytes) : "memory");
}
#endif
#ifndef NNS_CLUSTER_HINT
#define NNS_CLUSTER_HINT 100000u
#endif
__device__ __forceinline__ void cluster_wait(const void *bar, uint32_t parity) {
    unsigned spins = 0;
    uint32_t ok = 0;
    do {
#ifdef NNS_CLUSTER_POLL
        asm volatile(
            "{\n.reg .pred P1;\n"
            "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n}"
            : "=r"(ok) : "r"(dsm_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
#else
        asm volatile(
            "{\n.reg .pred P1;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n}"
            : "=r"(ok) : "r"(dsm_u32(bar)), "r"(parity), "r"(NNS_CLUSTER_HINT) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();       // a stuck cluster traps instead of hanging the GPU
#endif
    } while (!ok);
}

// One Jacobi sweep of the cluster path is ordered so that the exchange with the neighbouring CTAs hides behind the rest of
// the sweep:
//   1. wait for the neighbours' band edges of the previous sweep (my halo rows of Pc);
//   2. ALL threads compute the two edge rows of the band, store them locally and straight into the neighbours' halo rows
//      (distributed shared memory, st.async completing on the neighbours' mbarriers: no fence, no separate arrival);
//   3. the interior rows of the band; CTA barrier;
//   4. only in the two CTAs that hold a global edge row (i = 0 or nx-1): the cells of that row; CTA barrier.
// The p BC list is applied in list order by the reference (boundary.py:34-86), but only the four corner cells of the grid
// depend on the order: a cell of column 0 / ny-1 in another row is written by bottom / top entries only and reads the new
// value next to it, a non-corner cell of row 0 / nx-1 is written by left / right entries only and reads the new value in
// the adjacent row -- the LAST entry of that side decides.  So the thread of column 1 (ny-2) writes the column cells with
// its row (the edge rows are final before they are pushed), all threads write the edge row after the sweep, and one thread
// per corner replays the list in registers on the corner and its two neighbours (their values before the first entry are
// the previous sweep's: Jacobi copies edge cells through).  Walking the list with a barrier (or a __syncwarp) per entry
// cost 1600-1900 cycles per sweep in the two edge CTAs, which set the pace of the whole cluster.
// Two mbarriers per neighbour, alternating with the sweep parity (arrival count 1 = the owner's re-arm with the expected
// bytes of a halo row): a neighbour's stores of sweep g+2 need my stores of g+1, which follow my wait for g, so a barrier
// is never more than one phase ahead of its waiter.
//
// Row layout in shared memory (pitch even, rows 16-byte aligned): the computed columns j = JLO .. JLO+nj-1 sit at the even
// index 2 + (j - JLO), so a thread updates a PAIR of cells with 128-bit loads of the centre / north / south / b pairs and
// two scalar loads for the outer west / east operands (3.5 shared-memory instructions per cell instead of 7):
//   walls   : JLO = 1, nj = ny-2; index 1 = column 0, index ny = column ny-1 (edge values)        => column j at 1 + j
//   periodic: JLO = 0, nj = ny;   index 1 = copy of column ny-1, index ny+2 = copy of column 0    => column j at 2 + j
// NPT > 0: every thread owns up to NPT fixed pairs of the band (pair index tid + k * blockDim.x, the two edge rows first)
// and keeps their p and b values in registers over the sweeps: a sweep then reads only the north / south pairs and the
// outer west / east cells from shared memory (64 instead of 96 bytes per pair), and the work is spread evenly over the
// threads.  NPT = 0: rows per warp, everything from shared memory (bands with more pairs per thread than the instantiated NPT).
template <bool PER, int NPT>
__global__ void __launch_bounds__(512, 1) direct_cluster_kernel(const DirectArgs a, int band, int pitch) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long hbar[2][2];    // [sweep parity][0: from the CTA above, 1: from the CTA below]
    __shared__ double s_pval[NNS_MAX_BC];                     // values of the p BC entries of this member
    __shared__ int s_pcode[NNS_MAX_BC];                       // side | type << 8 of the p BC entries
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int OFF = PER ? 2 : 1;
    const int NC = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
    const int nx = a.g.nx, ny = a.g.ny;
    const int nj = PER ? ny : ny - 2, xend = 2 + nj;          // computed cells of a row: indices [2, xend)
    const uint32_t row_bytes = 8u * (uint32_t)(nj + 2);      // what a neighbour stores into one of my halo rows per sweep: indices 1 .. xend
    const size_t N = (size_t)nx * ny;
    const int b = blockIdx.x / NC, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int i0 = r * band, i1 = min(nx, i0 + band), nloc = i1 - i0;
    const size_t plane = (size_t)(band + 2) * pitch;
    double *Pc = smem, *Pn = smem + plane, *Bs = smem + 2 * plane;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy, rho = a.g.rho;
    const double nu = a.nu_b ? a.nu_b[b] : a.g.nu;
    const double *bcval = a.bcval ? a.bcval + (size_t)b * a.n_bcs : nullptr;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double rden = 1.0 / (2.0 * (dx2 + dy2));
    const double cx = dy2 * rden, cy = dx2 * rden, kb = dx2 * dy2 * rden;
    const double r2dx = 1.0 / (2.0 * dx), r2dy = 1.0 / (2.0 * dy), rdt = 1.0 / dt;
    double *ug = a.u + (size_t)b * N, *vg = a.v + (size_t)b * N, *pg = a.p + (size_t)b * N;
    double *us = a.su + (size_t)b * N, *vs = a.sv + (size_t)b * N;
    const double fdt = PER ? a.force_x * dt : 0.0;

    if (tid == 0) {
        for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dsm_u32(&hbar[k >> 1][k & 1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int k = 0; k < 4; ++k) mbar_arm(&hbar[k >> 1][k & 1], row_bytes);       // sweeps 0 and 1
    }
    cluster.sync();            // every CTA's mbarriers exist (and are armed) before the first remote store
    const bool has_above = r > 0, has_below = r < NC - 1;
    const bool edge_cta = i0 == 0 || i1 == nx;
    // my edge rows in the neighbours' halo rows: row li = 1 is the lower halo row (band + 1) of the CTA above, row nloc the
    // upper halo row (0) of the CTA below; one pointer per buffer
    // (shared::cluster addresses; 0 = no neighbour.  My barrier [par][1] of the CTA above counts what I, its lower
    // neighbour, store; [par][0] of the CTA below what I, its upper neighbour, store.)
    uint32_t remA_c = has_above ? mapa_u32(dsm_u32(Pc + (size_t)(band + 1) * pitch), r - 1) : 0u;
    uint32_t remA_n = has_above ? mapa_u32(dsm_u32(Pn + (size_t)(band + 1) * pitch), r - 1) : 0u;
    uint32_t remB_c = has_below ? mapa_u32(dsm_u32(Pc), r + 1) : 0u;
    uint32_t remB_n = has_below ? mapa_u32(dsm_u32(Pn), r + 1) : 0u;
    const uint32_t barA0 = has_above ? mapa_u32(dsm_u32(&hbar[0][1]), r - 1) : 0u, barA1 = has_above ? mapa_u32(dsm_u32(&hbar[1][1]), r - 1) : 0u;
    const uint32_t barB0 = has_below ? mapa_u32(dsm_u32(&hbar[0][0]), r + 1) : 0u, barB1 = has_below ? mapa_u32(dsm_u32(&hbar[1][0]), r + 1) : 0u;
    // the last bottom / top entry of the p list: 0 none, 1 Dirichlet, 2 Neumann
    int bk = 0, tk = 0, ek = 0;
    double bgv = 0.0, tgv = 0.0, egv = 0.0;
    for (int k = 0; k < a.pbc.n; ++k) {
        const double g = bcval ? bcval[a.pbc.slot[k]] : a.pbc.value[k];
        const int kind = a.pbc.type[k] == NNS_BC_NEUMANN ? 2 : 1;
        if (a.pbc.side[k] == NNS_SIDE_BOTTOM) { bk = kind; bgv = g; }
        if (a.pbc.side[k] == NNS_SIDE_TOP) { tk = kind; tgv = g; }
        if ((a.pbc.side[k] == NNS_SIDE_LEFT && i0 == 0) || (a.pbc.side[k] == NNS_SIDE_RIGHT && i1 == nx)) { ek = kind; egv = g; }
    }
    // the global edge row of this CTA (if any): its row, the adjacent row, the sign of the Neumann step
    const int erow = i0 == 0 ? 1 : nloc, arow = i0 == 0 ? 2 : nloc - 1, eside = i0 == 0 ? NNS_SIDE_LEFT : NNS_SIDE_RIGHT;
    const double esg = i0 == 0 ? -dx : dx;
    if (tid < a.pbc.n) { s_pval[tid] = bcval ? bcval[a.pbc.slot[tid]] : a.pbc.value[tid]; s_pcode[tid] = a.pbc.side[tid] | (a.pbc.type[tid] << 8); }
    // own rows and the two halo rows of p from global memory (both buffers: the pad / copy cells must be finite everywhere)
    for (size_t q = tid; q < 3 * plane; q += blockDim.x) smem[q] = 0.0;
    __syncthreads();
    for (int li = warp; li < nloc + 2; li += nwarps) {
        const int i = i0 + li - 1;
        if (i < 0 || i >= nx) continue;
        for (int j = lane; j < ny; j += 32) Pc[li * pitch + OFF + j] = pg[(size_t)i * ny + j];
        if (PER && lane == 0) { Pc[li * pitch + 1] = pg[(size_t)i * ny + ny - 1]; Pc[li * pitch + ny + 2] = pg[(size_t)i * ny]; }
    }
    __syncthreads();

    // A pair of cells (indices x, x + 1; x even) of row li: Pc -> Pn, optionally also into a neighbour's halo row.  C, B: the
    // pair's own p and b values; returns its new p values (the second one is the edge cell's if the pair has one computed cell).
    auto pair_core = [&](int li, int x, double2 C, double2 B, uint32_t rem, uint32_t rbar) -> double2 {
        const int q = li * pitch + x;
        const double2 Nn = *reinterpret_cast<const double2 *>(Pc + q - pitch), S = *reinterpret_cast<const double2 *>(Pc + q + pitch);
        const double W = Pc[q - 1], E = Pc[q + 2];
        const double r0 = (C.y + W) * cx + (S.x + Nn.x) * cy - B.x;
        const double r1 = (E + C.x) * cx + (S.y + Nn.y) * cy - B.y;
        const bool v1 = x + 1 < xend;                       // (odd nj: the last pair has one computed cell)
        double e0 = W, e1 = v1 ? E : C.y;                   // walls: the edge cells next to the first / last computed cell
        if (!PER) {
            if (bk) e0 = bk == 2 ? r0 + (-dy) * bgv : bgv;
            if (tk) e1 = tk == 2 ? (v1 ? r1 : r0) + dy * tgv : tgv;
        }
        const bool first = x == 2, last = x + 2 >= xend;
        if (v1) *reinterpret_cast<double2 *>(Pn + q) = make_double2(r0, r1);
        else Pn[q] = r0;
        if (PER) {
            if (first) Pn[li * pitch + xend] = r0;                            // copy of column 0 behind the last column
            if (last) Pn[li * pitch + 1] = v1 ? r1 : r0;                      // copy of column ny-1 in front of the first
        } else {
            if (first) Pn[q - 1] = e0;
            if (last) Pn[li * pitch + xend] = e1;
        }
        if (rem) {
            if (v1) st_async_f64x2(rem + 8u * x, r0, r1, rbar);
            else st_async_f64(rem + 8u * x, r0, rbar);
            if (PER) {
                if (first) st_async_f64(rem + 8u * xend, r0, rbar);
                if (last) st_async_f64(rem + 8u, v1 ? r1 : r0, rbar);
            } else {
                if (first) st_async_f64(rem + 8u * (x - 1), e0, rbar);
                if (last) st_async_f64(rem + 8u * xend, e1, rbar);
            }
        }
        return make_double2(r0, v1 ? r1 : e1);
    };
    auto pair = [&](int li, int x, uint32_t rem, uint32_t rbar) {
        const int q = li * pitch + x;
        pair_core(li, x, *reinterpret_cast<const double2 *>(Pc + q), *reinterpret_cast<const double2 *>(Bs + q), rem, rbar);
    };
    const int npairs = (nj + 1) >> 1;
    // NPT > 0: this thread's pairs.  Pair index -> row: 0 .. npairs-1 row 1, npairs .. 2 npairs - 1 row nloc, then rows 2 .. nloc-1
    constexpr int NPTA = NPT > 0 ? NPT : 1;
    int pli[NPTA], px[NPTA];                 // row (0 = no pair) and first index of the pair
    double2 Ck[NPTA], Bk[NPTA];
#pragma unroll
    for (int k = 0; k < NPTA; ++k) {
        const int idx = tid + k * (int)blockDim.x, rs = idx / npairs;
        const int li = rs == 0 ? 1 : rs == 1 ? nloc : rs;
        const int i = i0 + li - 1;
        pli[k] = (NPT > 0 && rs < nloc && i != 0 && i != nx - 1) ? li : 0;
        px[k] = 2 + 2 * (idx - rs * npairs);
        Ck[k] = make_double2(0.0, 0.0); Bk[k] = make_double2(0.0, 0.0);
    }

    int gs = 0;                // sweeps done since the launch (hand-off phase)
#ifdef NNS_X_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define PF(k) do { const long long now_ = clock64(); pf[k] += now_ - pt; pt = now_; } while (0)
#else
#define PF(k)
#endif
    for (int n = 0; n < a.nsteps; ++n) {
        const double *uo = (n & 1) ? us : ug, *vo = (n & 1) ? vs : vg;   // u^n, v^n
        double *un = (n & 1) ? ug : us, *vn = (n & 1) ? vg : vs;         // u^{n+1}
        // RHS b of the own rows (direct_fd:56-66)
        for (int li = 1 + warp; li <= nloc; li += nwarps) {
            const int i = i0 + li - 1;
            for (int j = lane; j < ny; j += 32) {
                double bb = 0.0;
                if (i > 0 && i < nx - 1 && (PER || (j > 0 && j < ny - 1))) {
                    const size_t q = (size_t)i * ny + j, rq = (size_t)i * ny;
                    const int jp = j < ny - 1 ? j + 1 : 0, jm = j > 0 ? j - 1 : ny - 1;
                    const double ux = (uo[rq + jp] - uo[rq + jm]) * r2dx, vy = (vo[q + ny] - vo[q - ny]) * r2dy;
                    const double uy = (uo[q + ny] - uo[q - ny]) * r2dy, vx = (vo[rq + jp] - vo[rq + jm]) * r2dx;
                    bb = rho * (rdt * (ux + vy)) - ux * ux - 2.0 * (uy * vx) - vy * vy;
                }
                Bs[li * pitch + OFF + j] = kb * bb;
            }
        }
        __syncthreads();
        if (NPT > 0) {
#pragma unroll
            for (int k = 0; k < NPTA; ++k)
                if (pli[k]) {
                    const int q = pli[k] * pitch + px[k];
                    Bk[k] = *reinterpret_cast<const double2 *>(Bs + q);
                    Ck[k] = *reinterpret_cast<const double2 *>(Pc + q);
                }
        }
        PF(0);
        // exactly nit Jacobi sweeps with the p BCs after every sweep (direct_fd:76-86)
        for (int s = 0; s < a.g.nit; ++s, ++gs) {
            // 1. the neighbours' edges of the previous sweep; their barriers are armed again for the sweep after this one
            if (gs > 0) {
                const int hp = (gs - 1) & 1;
                if (lane == 0) {
                    const uint32_t par = (uint32_t)((gs - 1) >> 1) & 1u;
                    if (has_above) cluster_wait(&hbar[hp][0], par);
                    if (has_below) cluster_wait(&hbar[hp][1], par);
                    if (tid == 0) {
                        if (has_above) mbar_arm(&hbar[hp][0], row_bytes);
                        if (has_below) mbar_arm(&hbar[hp][1], row_bytes);
                    }
                }
                __syncwarp();
            }
            PF(1);
            if (NPT > 0) {
                // 2. + 3. the thread's own pairs, p and b from registers (k = 0 holds the edge-row pairs: they go out first)
#pragma unroll
                for (int k = 0; k < NPTA; ++k) {
                    if (k == 1) {
#ifdef NNS_CLUSTER_PLAIN_ST
                        if (has_above || has_below) {
                            asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");           // every thread's edge-row stores are issued
                            if (tid == 0) {
                                asm volatile("fence.release.cluster;" ::: "memory");
                                if (has_above) remote_arrive_relaxed((gs & 1) ? barA1 : barA0);
                                if (has_below) remote_arrive_relaxed((gs & 1) ? barB1 : barB0);
                            }
                        }
#endif
                        PF(2);
                    }
#ifdef NNS_X_EDGEONLY       // timing experiment: only the pushed rows are computed (wrong results): period = hand-off latency
                    if (pli[k] && (pli[k] == 1 || pli[k] == nloc)) {
#else
                    if (pli[k]) {
#endif
                        const int li = pli[k], x = px[k];
                        if (PER && x + 1 >= xend) Ck[k].y = Pc[li * pitch + x + 1];       // (odd ny: the wrap copy next to the last cell)
                        const uint32_t rem = li == 1 ? remA_n : li == nloc ? remB_n : 0u;       // (0 without that neighbour)
                        const uint32_t rbar = li == 1 ? ((gs & 1) ? barA1 : barA0) : ((gs & 1) ? barB1 : barB0);
                        Ck[k] = pair_core(li, x, Ck[k], Bk[k], rem, rbar);
                    }
                }
            } else {
            // 2. edge rows first: li = 1 (to the CTA above) and li = nloc (to the CTA below)
            for (int t = tid; t < 2 * npairs; t += blockDim.x) {
                const bool low = t >= npairs;                          // (nloc >= 3 in every CTA: the launcher checks)
                const int li = low ? nloc : 1, x = 2 + 2 * (low ? t - npairs : t), i = i0 + li - 1;
                if (i == 0 || i == nx - 1) continue;                    // global edge rows: copied below, set by the BC replay
                const uint32_t rem = low ? remB_n : remA_n;
                pair(li, x, rem, low ? ((gs & 1) ? barB1 : barB0) : ((gs & 1) ? barA1 : barA0));
            }
#ifdef NNS_CLUSTER_PLAIN_ST
            if (has_above || has_below) {
                asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");           // every thread's edge-row stores are issued
                if (tid == 0) {
                    asm volatile("fence.release.cluster;" ::: "memory");
                    if (has_above) remote_arrive_relaxed((gs & 1) ? barA1 : barA0);
                    if (has_below) remote_arrive_relaxed((gs & 1) ? barB1 : barB0);
                }
            }
#endif
            PF(2);
            // 3. interior rows of the band (and the copy of a global edge row)
            for (int li = 2 + warp; li <= nloc - 1; li += nwarps)
                for (int x = 2 + 2 * lane; x < xend; x += 64) pair(li, x, 0u, 0u);
            }
            if (i0 == 0)
                for (int k = tid; k < pitch; k += blockDim.x) Pn[1 * pitch + k] = Pc[1 * pitch + k];
            if (i1 == nx)
                for (int k = tid; k < pitch; k += blockDim.x) Pn[nloc * pitch + k] = Pc[nloc * pitch + k];
            PF(3);
            __syncthreads();
            PF(4);
            // 4. the BC list where a global edge row is involved
            if (edge_cta && a.pbc.n > 0) {
                double *er = Pn + erow * pitch + OFF;
                const double *ar = Pn + arow * pitch + OFF, *eo = Pc + erow * pitch + OFF, *ao = Pc + arow * pitch + OFF;
                if (ek)
                    for (int j = (PER ? 0 : 1) + tid; j < (PER ? ny : ny - 1); j += blockDim.x) er[j] = ek == 2 ? ar[j] + esg * egv : egv;
                if (!PER && tid >= blockDim.x - 2) {                  // the two corners of the edge row, list order in registers
                    const bool topc = tid == blockDim.x - 1;
                    const int jc = topc ? ny - 1 : 0, ji = topc ? ny - 2 : 1, cside = topc ? NNS_SIDE_TOP : NNS_SIDE_BOTTOM;
                    const double csg = topc ? dy : -dy, a11 = ar[ji];
                    double t_adj = ao[jc], t_row = eo[ji], c = eo[jc];
                    for (int k = 0; k < a.pbc.n; ++k) {
                        const double g = s_pval[k];
                        const int side = s_pcode[k] & 0xff;
                        const bool neu = (s_pcode[k] >> 8) == NNS_BC_NEUMANN;
                        if (side == eside) { t_row = neu ? a11 + esg * g : g; c = neu ? t_adj + esg * g : g; }
                        else if (side == cside) { t_adj = neu ? a11 + csg * g : g; c = neu ? t_row + csg * g : g; }
                    }
                    er[jc] = c;
                }
                PF(7);
                __syncthreads();
            }
            PF(5);
            double *t = Pc; Pc = Pn; Pn = t;
            uint32_t ta = remA_c; remA_c = remA_n; remA_n = ta;
            ta = remB_c; remB_c = remB_n; remB_n = ta;
        }
        // the last sweep's edges of the neighbours are needed by the pressure gradient of my edge rows
        if (gs > 0) {
            if (tid == 0) {
                const uint32_t par = (uint32_t)((gs - 1) >> 1) & 1u;
                if (has_above) cluster_wait(&hbar[(gs - 1) & 1][0], par);
                if (has_below) cluster_wait(&hbar[(gs - 1) & 1][1], par);
            }
            __syncthreads();
        }
        // velocity update of the own rows (direct_fd:98-118), then u/v BCs (:121-125)
        const double kpx = dt / (2.0 * rho * dx), kpy = dt / (2.0 * rho * dy);
        const double kdx = dt / dx2, kdy = dt / dy2, ax = dt / dx, ay = dt / dy;
        for (int li = 1 + warp; li <= nloc; li += nwarps) {
            const int i = i0 + li - 1;
            for (int j = lane; j < ny; j += 32) {
                const size_t q = (size_t)i * ny + j;
                const double uc = uo[q], vc = vo[q];
                double ru = uc, rv = vc;
                if (i > 0 && i < nx - 1 && (PER || (j > 0 && j < ny - 1))) {
                    const int row = li * pitch + OFF, sq = row + j;
                    const int jp = j < ny - 1 ? j + 1 : 0, jm = j > 0 ? j - 1 : ny - 1;
                    const size_t rq = (size_t)i * ny;
                    const double uW = uo[rq + jm], uE = uo[rq + jp], uN = uo[q - ny], uS = uo[q + ny];
                    const double vW = vo[rq + jm], vE = vo[rq + jp], vN = vo[q - ny], vS = vo[q + ny];
                    ru = uc - uc * ax * (uc - uW) - vc * ay * (uc - uN) - kpx * (Pc[row + jp] - Pc[row + jm]) +
                         nu * (kdx * (uE - 2.0 * uc + uW) + kdy * (uS - 2.0 * uc + uN)) + fdt;
                    rv = vc - uc * ax * (vc - vW) - vc * ay * (vc - vN) - kpy * (Pc[sq + pitch] - Pc[sq - pitch]) +
                         nu * (kdx * (vE - 2.0 * vc + vW) + kdy * (vS - 2.0 * vc + vN));
                }
                un[q] = ru;
                vn[q] = rv;
            }
        }
        __syncthreads();
        band_apply_bc_global(un, nx, ny, i0, i1, a.ubc, bcval, dx, dy);
        band_apply_bc_global(vn, nx, ny, i0, i1, a.vbc, bcval, dx, dy);
        if (a.traj_u || (a.flags & NNS_FLAG_CHECK_FINITE)) {
            const size_t toff = ((size_t)b * a.nsteps_total + (a.step0 + n)) * N;
            unsigned long long bad = 0;
            for (int li = 1 + warp; li <= nloc; li += nwarps) {
                const int i = i0 + li - 1;
                for (int j = lane; j < ny; j += 32) {
                    const size_t q = (size_t)i * ny + j;
                    const double x = un[q], y = vn[q], z = Pc[li * pitch + OFF + j];
                    if (a.traj_u) { a.traj_u[toff + q] = x; a.traj_v[toff + q] = y; a.traj_p[toff + q] = z; }
                    bad += !(isfinite(x) && isfinite(y) && isfinite(z));
                }
            }
            if ((a.flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(a.nonfinite, bad);
        }
        // the neighbouring CTAs read my rows of u^{n+1}, v^{n+1} (global memory) in their next RHS / update
        __threadfence();
        cluster.sync();
        PF(6);
    }
#ifdef NNS_X_PROF
    if (tid == 0 && a.nsteps >= 100) printf("cta %d: per sweep: wait %lld edge %lld interior %lld sync %lld+%lld bc %lld | per step: rhs %lld update %lld\n", r,
        pf[1] / gs, pf[2] / gs, pf[3] / gs, pf[4] / gs, pf[7] / gs, pf[5] / gs, pf[0] / a.nsteps, pf[6] / a.nsteps);
#endif
    for (int li = 1 + warp; li <= nloc; li += nwarps) {
        const int i = i0 + li - 1;
        for (int j = lane; j < ny; j += 32) pg[(size_t)i * ny + j] = Pc[li * pitch + OFF + j];
    }
    if (a.nsteps & 1)
        for (int li = 1 + warp; li <= nloc; li += nwarps) {
            const int i = i0 + li - 1;
            for (int j = lane; j < ny; j += 32) { const size_t q = (size_t)i * ny + j; ug[q] = us[q]; vg[q] = vs[q]; }
        }
}

// ---- stream path ------------------------------------------------------------------------
__global__ void direct_rhs_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                  double *__restrict__ bout, Geometry g) {
    const int nx = g.nx, ny = g.ny;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ny) return;
    const size_t base = (size_t)blockIdx.z * nx * ny, q = base + (size_t)i * ny + j;
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
    const double kb = dx2 * dy2 / (2.0 * (dx2 + dy2));
    double bb = 0.0;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
        const double r2dx = 1.0 / (2.0 * g.dx), r2dy = 1.0 / (2.0 * g.dy);
        const double ux = (u[q + 1] - u[q - 1]) * r2dx, vy = (v[q + ny] - v[q - ny]) * r2dy;
        const double uy = (u[q + ny] - u[q - ny]) * r2dy, vx = (v[q + 1] - v[q - 1]) * r2dx;
        bb = g.rho * ((1.0 / g.dt) * (ux + vy)) - ux * ux - 2.0 * (uy * vx) - vy * vy;
    }
    bout[q] = kb * bb;
}

__global__ void direct_jacobi_kernel(const double *__restrict__ pc, const double *__restrict__ bs,
                                     double *__restrict__ pn, Geometry g) {
    const int nx = g.nx, ny = g.ny;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ny) return;
    const size_t q = (size_t)blockIdx.z * nx * ny + (size_t)i * ny + j;
    const double dx2 = g.dx * g.dx, dy2 = g.dy * g.dy;
    const double rden = 1.0 / (2.0 * (dx2 + dy2));
    double r = pc[q];
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1)
        r = (pc[q + 1] + pc[q - 1]) * (dy2 * rden) + (pc[q + ny] + pc[q - ny]) * (dx2 * rden) - bs[q];
    pn[q] = r;
}

__global__ void direct_update_kernel(const double *__restrict__ uo, const double *__restrict__ vo,
                                     const double *__restrict__ p, double *__restrict__ un,
                                     double *__restrict__ vn, Geometry g, const double *__restrict__ nu_b) {
    const int nx = g.nx, ny = g.ny;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ny) return;
    const size_t q = (size_t)blockIdx.z * nx * ny + (size_t)i * ny + j;
    const double nu = nu_b ? nu_b[blockIdx.z] : g.nu;
    const double dt = g.dt, dx = g.dx, dy = g.dy, rho = g.rho;
    const double uc = uo[q], vc = vo[q];
    double ru = uc, rv = vc;
    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
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
}

__global__ void apply_bc_kernel(double *A, Geometry g, BcList L, const double *bcval, int n_bcs) {
    const size_t N = (size_t)g.nx * g.ny;
    cta_apply_bc_global(A + blockIdx.x * N, g.nx, g.ny, L, bcval ? bcval + (size_t)blockIdx.x * n_bcs : nullptr,
                        g.dx, g.dy);
}

__global__ void snapshot_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                const double *__restrict__ p, double *tu, double *tv, double *tp, size_t N,
                                int nsteps_total, int step, unsigned long long *nonfinite, int flags) {
    const size_t b = blockIdx.y;
    const size_t toff = (b * nsteps_total + step) * N;
    unsigned long long bad = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (size_t)gridDim.x * blockDim.x) {
        const double x = u[b * N + q], y = v[b * N + q], z = p[b * N + q];
        if (tu) { tu[toff + q] = x; tv[toff + q] = y; tp[toff + q] = z; }
        bad += !(isfinite(x) && isfinite(y) && isfinite(z));
    }
    if ((flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(nonfinite, bad);
}

int launch_apply_bc(nns_handle *h, int field, double *a, cudaStream_t st) {
    if (h->bc[field].n == 0) return NNS_OK;
    apply_bc_kernel<<<h->g.batch, 128, 0, st>>>(a, h->g, h->bc[field], h->d_bcval, h->n_bcs);
    NNS_CUDA(cudaGetLastError());
    h->launches += 1;
    return NNS_OK;
}

static int ensure(double **p, size_t bytes) {
    if (*p) return NNS_OK;
    NNS_CUDA(cudaMalloc(p, bytes));
    return NNS_OK;
}

int direct_run(nns_handle *h, double *u, double *v, double *p, int nsteps, double *tu, double *tv, double *tp,
               cudaStream_t st) {
    const Geometry &g = h->g;
    const size_t N = (size_t)g.nx * g.ny, bytes = sizeof(double) * N * g.batch;
    int rc;
    if ((rc = ensure(&h->d_scratch[0], bytes)) || (rc = ensure(&h->d_scratch[1], bytes))) return rc;
    const int pitch = g.ny | 1;
    const size_t smem = sizeof(double) * 3 * (size_t)g.nx * pitch;
    const char *dmode = getenv("NNS_DIRECT_MODE");          // tests: "cluster" / "stream" skip the earlier paths
    const bool skip_chip = dmode && (strcmp(dmode, "cluster") == 0 || strcmp(dmode, "stream") == 0);
    if (smem <= (size_t)h->max_smem_optin && !skip_chip) {
        DirectArgs a{};
        a.g = g; a.ubc = h->bc[0]; a.vbc = h->bc[1]; a.pbc = h->bc[2];
        a.nu_b = h->d_nu; a.bcval = h->d_bcval; a.n_bcs = h->n_bcs;
        a.nsteps = nsteps; a.nsteps_total = nsteps; a.step0 = 0; a.flags = h->params.flags; a.force_x = h->params.force_x;
        a.u = u; a.v = v; a.p = p; a.su = h->d_scratch[0]; a.sv = h->d_scratch[1];
        a.traj_u = tu; a.traj_v = tv; a.traj_p = tp; a.nonfinite = h->d_nonfinite;
        const long cells = (long)g.nx * g.ny;
        const int threads = cells >= 8192 ? 1024 : cells >= 1024 ? 512 : 256;
        auto kern = threads == 1024 ? direct_chip_kernel<1024> : threads == 512 ? direct_chip_kernel<512> : direct_chip_kernel<256>;
        NNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<g.batch, threads, smem, st>>>(a);
        NNS_CUDA(cudaGetLastError());
        h->launches += 1;
        return NNS_OK;
    }
    // cluster path: row bands of one member over the CTAs of a thread-block cluster
    if (!(dmode && strcmp(dmode, "stream") == 0)) {
        // the sweeps are bound by shared-memory bandwidth: the largest cluster whose bands still have >= 8 rows, else
        // the largest one that fits at all
        for (int pass = 0; pass < 2; ++pass)
        for (int nc = 16; nc >= 2; nc /= 2) {
            const int band = (g.nx + nc - 1) / nc;
            const int cpitch = (g.ny + 5) & ~1;             // even, >= ny + 4 (see the kernel's row layout)
            const size_t csmem = sizeof(double) * 3 * (size_t)(band + 2) * cpitch;
            // (the first and the last CTA need >= 3 rows: their pushed row must not touch the global edge row, see the kernel)
            if (band < (pass == 0 ? 8 : 3) || g.nx - (nc - 1) * band < 3 || g.ny < 3 || csmem > (size_t)h->max_smem_optin) continue;
            DirectArgs a{};
            a.g = g; a.ubc = h->bc[0]; a.vbc = h->bc[1]; a.pbc = h->bc[2];
            a.nu_b = h->d_nu; a.bcval = h->d_bcval; a.n_bcs = h->n_bcs;
            a.nsteps = nsteps; a.nsteps_total = nsteps; a.step0 = 0; a.flags = h->params.flags; a.force_x = h->params.force_x;
            a.u = u; a.v = v; a.p = p; a.su = h->d_scratch[0]; a.sv = h->d_scratch[1];
            a.traj_u = tu; a.traj_v = tv; a.traj_p = tp; a.nonfinite = h->d_nonfinite;
            const bool per = h->params.flags & NNS_FLAG_PERIODIC_X;
            const char *thr = getenv("NNS_DIRECT_THREADS");        // experiments
            const int nthreads = thr ? std::min(512, std::max(64, atoi(thr) & ~31)) : 512;
            const int npairs = ((per ? g.ny : g.ny - 2) + 1) / 2;
            // p / b of a thread's pairs in registers: measured 0.079 against 0.085 ms/step on the periodic channel, 0.107
            // against 0.104 on the cavity (256 x 256; the sweep period is set by the neighbour hand-off, not by the
            // shared-memory traffic it saves) -- used for the periodic variant; NNS_DIRECT_NPT = 0 / 4 forces either
            const char *npt_env = getenv("NNS_DIRECT_NPT");
            const bool fits4 = (long)band * npairs <= 4L * nthreads;
            const bool regs4 = fits4 && (npt_env ? atoi(npt_env) == 4 : per);
            auto kern = per ? (regs4 ? direct_cluster_kernel<true, 4> : direct_cluster_kernel<true, 0>)
                            : (regs4 ? direct_cluster_kernel<false, 4> : direct_cluster_kernel<false, 0>);
            NNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
            if (nc > 8) NNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(g.batch * nc));
            cfg.blockDim = dim3((unsigned)nthreads);
            cfg.dynamicSmemBytes = csmem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)nc; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                continue;           // this cluster size cannot be scheduled on the device: try the next one / the stream path
            }
            NNS_CUDA(cudaLaunchKernelEx(&cfg, kern, a, band, cpitch));
            h->launches += 1;
            return NNS_OK;
        }
    }
    // stream path
    if (h->params.flags & NNS_FLAG_PERIODIC_X) {
        set_error("direct_fd periodic-x extension: the grid must fit the chip or the cluster path (nx * ny * 24 B over at most 16 SMs)");
        return NNS_ERR_UNSUPPORTED;
    }
    if ((rc = ensure(&h->d_b, bytes)) || (rc = ensure(&h->d_p2, bytes))) return rc;
    const dim3 blk(128), grd((g.ny + 127) / 128, g.nx, g.batch);
    double *uc = u, *vc = v, *un = h->d_scratch[0], *vn = h->d_scratch[1];
    for (int n = 0; n < nsteps; ++n) {
        direct_rhs_kernel<<<grd, blk, 0, st>>>(uc, vc, h->d_b, g);
        double *pc = p, *pn = h->d_p2;
        for (int s = 0; s < g.nit; ++s) {
            direct_jacobi_kernel<<<grd, blk, 0, st>>>(pc, h->d_b, pn, g);
            h->launches += 1;
            if ((rc = launch_apply_bc(h, 2, pn, st))) return rc;
            double *t = pc; pc = pn; pn = t;
        }
        if (pc != p) NNS_CUDA(cudaMemcpyAsync(p, pc, bytes, cudaMemcpyDeviceToDevice, st));
        direct_update_kernel<<<grd, blk, 0, st>>>(uc, vc, p, un, vn, g, h->d_nu);
        h->launches += 2;
        if ((rc = launch_apply_bc(h, 0, un, st)) || (rc = launch_apply_bc(h, 1, vn, st))) return rc;
        if (tu || (h->params.flags & NNS_FLAG_CHECK_FINITE)) {
            snapshot_kernel<<<dim3(64, g.batch), 256, 0, st>>>(un, vn, p, tu, tv, tp, N, nsteps, n,
                                                               h->d_nonfinite, h->params.flags);
            h->launches += 1;
        }
        double *t;
        t = uc; uc = un; un = t;
        t = vc; vc = vn; vn = t;
    }
    NNS_CUDA(cudaGetLastError());
    if (uc != u) {
        NNS_CUDA(cudaMemcpyAsync(u, uc, bytes, cudaMemcpyDeviceToDevice, st));
        NNS_CUDA(cudaMemcpyAsync(v, vc, bytes, cudaMemcpyDeviceToDevice, st));
    }
    return NNS_OK;
}

}  // namespace nns
