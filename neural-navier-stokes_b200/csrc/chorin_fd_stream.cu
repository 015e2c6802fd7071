// chorin_fd_stream.cu -- chorin_fd (explicit) ensemble step as a PERSISTENT, WARP-SPECIALISED kernel:
// one CTA per SM loops over its ensemble members; while the 8 "SOR warps" run the 49 exact-order
// Gauss-Seidel sweeps of member k out of REGISTERS, the 4 "stencil warps" of the same CTA stream the
// predictor of member k+1 and the projection of member k-1 through a TMA-fed (cp.async.bulk)
// shared-memory row ring.  The memory-bound phases therefore overlap the FP64-bound phase on the same
// SM (the one-CTA-per-member kernel of chorin_fd_chip.cu serialises them: 43% of its time).
//
// Reference semantics (src/chorin_fd/simulate.py of mhw32/neural-navier-stokes), per member:
//   stencil pass 1  _explicit_predictor_step :63-91 (x-only advection differences kept), u_bc / v_bc in
//                   list order :221-225, pressure right-hand side :186-188 (pre-scaled: C')
//   SOR             _get_pressure :169-202: lexicographic SOR, <= nit-1 sweeps, exit at max|dp| <= tol,
//                   executed as the block Gauss-Seidel wavefront of chorin_fd_chip.cu (thread-owned
//                   BR x BC blocks of p in registers, one named barrier per super-stage)
//   stencil pass 2  p_bc in list order :230-231, _correction_step :204-210, trajectory snapshot :263-265
//
// Register budget (setmaxnreg): 256 SOR threads x 208 + 128 stencil threads x 88 = 64512 registers.
// Shared memory: C' (thread-private 16-byte chunks, 128 KiB) + halo slots (64 KiB) + row ring (24 KiB).
// Hand-offs inside the CTA use named barriers (bar.arrive / bar.sync) in a full/empty protocol; the
// C' image of the next member and the intermediate velocities travel through L2-resident global memory.
#include <algorithm>
#include <type_traits>
#include <vector>

#include "nns_common.cuh"
#include "sor_block.cuh"

namespace nns {

namespace {

constexpr int NT_ST = 128;      // threads of the stencil role (1 warpgroup)
constexpr int GR = 4;           // rows per ring group of the legacy kernel (one bulk copy per field, one full/empty mbarrier pair)
constexpr int GRW = 8;          // rows per ring group of the wave kernel (measured: 4 -> 8 rows per group: 5.29 -> 4.71 ms/step)
constexpr int NG = 2;           // groups in the ring of the legacy kernel (double buffer)
constexpr int NGW = 4;          // groups in the ring of the wave kernel (C' lives in Tensor Memory: room for a deep ring)
constexpr int RING = GR * NG;   // rows per field in the stencil ring
#ifndef NNS_REGS_SOR
#define NNS_REGS_SOR 192     // 8 x 32 x 192 + 4 x 32 x 120 = 64512 = 384 x 168 (the CTA's allocation)
#define NNS_REGS_ST 120
#endif
constexpr int REGS_SOR = NNS_REGS_SOR, REGS_ST = NNS_REGS_ST;
constexpr int REGS_SOR_W = NNS_REGS_SOR, REGS_ST_W = NNS_REGS_ST;     // setmaxnreg only moves registers inside the CTA's launch allocation (384 x 168)
constexpr int NW_SOR = NT_SOR / 32;
constexpr int N_SCRATCH = 4;    // per-CTA scratch sets: launches on different internal streams (nns_chorin_fd_step_host) may overlap

// named barrier ids (0 is __syncthreads)
enum { BAR_SOR = 1, BAR_ST = 2, BAR_READY = 3, BAR_CONSUMED = 5, BAR_DONE = 7 };   // +0/+1 by member parity

struct SBlock {          // one per SOR thread (host-built)
    short r0, c0;        // first interior row / column of the block
    short nN, nS, nW, nE;   // thread ids of the neighbouring blocks, -1 = physical boundary
    short bd;            // anti-diagonal of the top sub-block: 2 * bi + bj
    short pad;
};

struct StreamArgs {
    Geometry g;
    BcList ubc, vbc, pbc;
    const double *nu_b;
    const double *bcval;
    int n_bcs;
    int count;               // members handled by this launch
    int flags;
    const SBlock *desc;      // [NT_SOR]
    const short *tidmap;     // [NBR*NBC] thread id of block (bi, bj)
    const double *uc, *vc, *up, *vp;   // u^n, v^n, u^{n-1}, v^{n-1}
    double *un, *vn;         // u^{n+1}, v^{n+1} (hold ui, vi in between)
    double *p;
    double *cimg;            // [gridDim.x][NCH*NT_SOR] double2 images of C' in the smem layout
    double *traj_u, *traj_v, *traj_p;
    size_t traj_member_stride, traj_off;   // element offsets: member stride, offset of this step
    int32_t *sweeps;         // [count] or null
    unsigned long long *nonfinite;
    long long *prof;         // optional [gridDim.x][NPROF] phase cycle counters (NNS_STREAM_PROF=1)
    // legacy kernel as the re-run pass of the wave kernel: members list[0 .. *list_count) instead of 0 .. count
    const int *list;
    const int *list_count;
    // wave kernel
    int *redo_list;          // members whose SOR loop stops early (or undecided by the fast test): re-run pass
    int *redo_count;
    double *pscr;            // [gridDim.x][2][NX*NY] SOR results before p_bc (L2-resident scratch)
    unsigned depmask[NW_SOR];            // per SOR warp: the warps that hold neighbours of its blocks (stage hand-off)
    int wamin[NW_SOR], wrange[NW_SOR];   // per SOR warp: first block diagonal and spread of its lanes
    int rmax;                // largest spread
    long long *trace;        // optional (NNS_WAVE_TRACE=1): clock64 of CTA 0 per SOR warp at sweep start / barrier arrival / release
};

// phase timers: thread `lead` of a role accumulates clock64() deltas into prof[slot]
constexpr int NPROF = 32;
constexpr int TRACE_STAGES = 160;   // NNS_STREAM_TRACE builds: stage trace of CTA 0, second member
#ifdef NNS_STREAM_PROF_FINE      // timers inside the stencil passes: they cost registers even when switched off at run time
#define NNS_FINE(...) __VA_ARGS__
#else
#define NNS_FINE(...)
#endif
#define NNS_PROF_T() (a.prof ? clock64() : 0ll)
#define NNS_PROF_ADD(slot, t0) do { if (a.prof && lead) a.prof[(size_t)blockIdx.x * NPROF + (slot)] += clock64() - (t0); } while (0)

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
#ifdef NNS_DEBUG_TRAP        // a stuck pipeline traps instead of hanging the GPU (experiments)
    unsigned spins = 0;
    uint32_t ok = 0;
    do {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 22)) { printf("mbar_wait stuck: block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, (void *)b, parity); __trap(); }
    } while (!ok);
#else
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
#endif
}
// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
#ifdef NNS_SOR_STAGE_HANDOFF
__device__ __forceinline__ void mbar_arrive_release(uint64_t *b) {      // SASS: a bare SYNCS.ARRIVE (no MEMBAR)
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// blocking wait with a hardware suspend hint: the waiting warp sleeps instead of polling through the LSU
__device__ __forceinline__ void mbar_wait_acquire(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP_A:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P1, [%0], %1, 0x989680;\n"
        "@P1 bra WAIT_DONE_A;\n"
        "bra WAIT_LOOP_A;\n"
        "WAIT_DONE_A:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
#endif
#ifdef NNS_SOR_TMEM
// 16 consecutive 32-bit Tensor Memory columns of the thread's own lane (4 chunks of C')
__device__ __forceinline__ void tm_st16_top(uint32_t taddr, const double2 (&v)[4]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)), "r"(__double2hiint(v[0].y)),
        "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)),
        "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
        "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)), "r"(__double2hiint(v[3].y))
        : "memory");
}
#endif
// 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__constant__ short c_ord[128];     // diag_ord of cell q = li * BC + lj (host-filled; used by the stencil role)

// All sweeps of one member as a wavefront of SUB-blocks.  Every thread owns a BR x BC block of p and
// sweeps it as two sub-blocks, rows [0, RS) and [RS, BR), in alternate super-stages: with sub-block row
// index sbi = 2*bi + {0, 1}, sub-block (sbi, bj) performs sweep s at super-stage T = sbi + bj + 2s.  All
// four lexicographic dependencies of a sub-block are then one super-stage old (north / west: same sweep,
// south / east: previous sweep), the boundary between the two sub-blocks of a thread stays in registers,
// and -- unlike a one-block-per-thread wavefront, where a thread idles every other super-stage -- every
// thread works in every super-stage, so each SM sub-partition always has TWO warps to issue from.
// sd = 2*bi + bj.  mask: bit s set = sweep s still violates the exit test; amb: bit s set = undecided by
// the fast test (TRACK 1 only).
template <int BR, int BC, int RS, int TRACK>
__device__ __forceinline__ void wavefront(double (&P)[BR][BC], const double2 *Cme, const SHalo<BR, BC> &h, bool owner,
                                          int sd, int tmax, int cap, const Coef &k, unsigned long long tolbits,
                                          unsigned long long &mask, unsigned long long &amb, uint64_t (*hb)[2], unsigned dep, int &gbase,
                                          uint32_t tmc, long long *prof = nullptr, long long *trace = nullptr) {
    // One CTA-wide (SOR role) barrier per stage.  EXPERIMENT (-DNNS_SOR_STAGE_HANDOFF, measured slower: 3.94 against 3.62
    // ms/step): point-to-point hand-off instead.  A sub-block sweep of stage T reads what its four neighbours published
    // in stage T-1 and overwrites what they read in stage T-1, so a warp may enter stage T as soon as the warps that hold
    // neighbours of its blocks (bit mask `dep`, host-built, symmetric) have completed T-1.  Every warp owns two mbarriers
    // (stages of even / odd global index G, arrival count = number of its neighbour warps): after stage G a warp arrives
    // (release: a bare SYNCS.ARRIVE) on hb[c][G & 1] of each neighbour warp c, before stage G it waits (acquire,
    // hardware-suspended) for phase (G-1)/2 of its own hb[w][(G-1) & 1]; a neighbour cannot complete G+1 -- the next
    // arrival on the same barrier -- before this warp has completed G, so a barrier is never more than one phase ahead
    // of its waiter.  Bit-identical results, but every warp has 5 of the 7 others as neighbours: it is a CTA barrier
    // built from slower parts (~270 cycles of wait per stage in the trace against ~130 for BAR.SYNC).  Polling progress
    // counters through the LSU instead of mbarriers: 7.6 ms/step (the pollers starve the warps they wait for).
#if defined(NNS_SOR_STAGE_HANDOFF) || defined(NNS_STREAM_TRACE)
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#endif
    if (owner) publish<BR, BC, 0, BR>(P, h);   // the whole perimeter once
    named_sync(BAR_SOR, NT_SOR);
    const unsigned tolhi = (unsigned)(tolbits >> 32);
    // (Splitting this loop into top / bottom stage pairs with the sub-block kind fixed at compile time, as the wave
    // kernel does, is 2 % faster here too -- 4.42 ms/step -- but made the results of this kernel NON-DETERMINISTIC
    // (about 1 % of the members of later rounds off by 1e-6 .. 1e-2 from run to run, scripts/determinism_check.py); the
    // cause was not found, so the single loop stays.)
    for (int T = 0; T <= tmax; ++T) {
        const int q = T - sd;                       // 2s for the top sub-block, 2s + 1 for the bottom one
        const bool work = owner && q >= 0 && q <= 2 * (cap - 1) + 1;
#ifdef NNS_SOR_STAGE_HANDOFF
#ifdef NNS_STREAM_TRACE
        if (trace && lane == 0 && T < TRACE_STAGES) trace[T * 4 + 3] = clock64();
#endif
        const int G = gbase + T;
        if (G > 0) mbar_wait_acquire(&hb[wid][(G - 1) & 1], (unsigned)((G - 1) >> 1) & 1u);
#endif
#ifdef NNS_STREAM_TRACE
        const unsigned nact = __popc(__ballot_sync(0xffffffffu, work));
        if (trace && (threadIdx.x & 31) == 0 && T < TRACE_STAGES) { trace[T * 4 + 0] = clock64(); trace[T * 4 + 2] = nact; }
#endif
#ifdef NNS_SOR_TMEM
        // C' in Tensor Memory: tcgen05.ld is warp-collective, so the whole warp runs the sweep as soon as one lane has work;
        // the other lanes commit nothing (block_sweep_tmt)
        if (__any_sync(0xffffffffu, work)) {
            unsigned mhi = 0u;
            bool v = false;
            if (!(q & 1)) block_sweep_tmt<BR, BC, RS, 0, RS, TRACK>(P, tmc, h, k, work, tolbits, mhi, v);
            else block_sweep_tmt<BR, BC, RS, RS, BR, TRACK>(P, tmc, h, k, work, tolbits, mhi, v);
            if (work) {
                if (TRACK == 1) {
                    mask |= (unsigned long long)(mhi > tolhi) << (q >> 1);
                    amb |= (unsigned long long)(mhi == tolhi) << (q >> 1);
                } else if (TRACK == 2) {
                    mask |= (unsigned long long)v << (q >> 1);
                }
            }
        }
#else
        if (work) {
            unsigned mhi = 0u;
            bool v = false;
#ifdef NNS_STREAM_PROF_SWEEP
            const long long ts0 = prof ? clock64() : 0ll;
#endif
#ifdef NNS_SOR_PFD       // experiment: explicit C' prefetch distance (anti-diagonals)
            if (!(q & 1)) block_sweep_pf<BR, BC, RS, 0, RS, TRACK, NNS_SOR_PFD>(P, Cme, h, k, tolbits, mhi, v);
            else block_sweep_pf<BR, BC, RS, RS, BR, TRACK, NNS_SOR_PFD>(P, Cme, h, k, tolbits, mhi, v);
#else
            if (!(q & 1)) block_sweep<BR, BC, RS, 0, RS, TRACK>(P, Cme, h, k, tolbits, mhi, v);
            else block_sweep<BR, BC, RS, RS, BR, TRACK>(P, Cme, h, k, tolbits, mhi, v);
#endif
#ifdef NNS_STREAM_PROF_SWEEP
            if (prof) { prof[0] += clock64() - ts0; prof[1] += 1; }
#endif
            if (TRACK == 1) {
                mask |= (unsigned long long)(mhi > tolhi) << (q >> 1);
                amb |= (unsigned long long)(mhi == tolhi) << (q >> 1);
            } else if (TRACK == 2) {
                mask |= (unsigned long long)v << (q >> 1);
            }
        }
#endif
#ifdef NNS_STREAM_TRACE
        __syncwarp();
        if (trace && (threadIdx.x & 31) == 0 && T < TRACE_STAGES) trace[T * 4 + 1] = clock64();
#endif
#ifndef NNS_SOR_STAGE_HANDOFF
        named_sync(BAR_SOR, NT_SOR);
#ifdef NNS_STREAM_TRACE
        if (trace && (threadIdx.x & 31) == 0 && T < TRACE_STAGES) trace[T * 4 + 3] = clock64();
#endif
#else
        __syncwarp();
        if (lane < NW_SOR && ((dep >> lane) & 1u)) mbar_arrive_release(&hb[lane][G & 1]);
#endif
    }
#ifdef NNS_SOR_STAGE_HANDOFF
    gbase += tmax + 1;
    named_sync(BAR_SOR, NT_SOR);
#endif
}

// Sequential BC list on a row-major GLOBAL field by the 128 stencil threads (boundary.py:34-86).
__device__ __forceinline__ void st_apply_bc_global(double *A, int nx, int ny, const BcList &L, const double *bcval,
                                                   double dx, double dy, int ts) {
    for (int kk = 0; kk < L.n; ++kk) {
        const double g = bcval ? bcval[L.slot[kk]] : L.value[kk];
        const int side = L.side[kk];
        const bool neu = L.type[kk] == NNS_BC_NEUMANN;
        if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
            const int i = side == NNS_SIDE_LEFT ? 0 : nx - 1, in = side == NNS_SIDE_LEFT ? 1 : nx - 2;
            const double sgn = side == NNS_SIDE_LEFT ? -dx : dx;
            for (int j = ts; j < ny; j += NT_ST) A[(size_t)i * ny + j] = neu ? A[(size_t)in * ny + j] + sgn * g : g;
        } else {
            const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
            const double sgn = side == NNS_SIDE_BOTTOM ? -dy : dy;
            for (int i = ts; i < nx; i += NT_ST) A[(size_t)i * ny + j] = neu ? A[(size_t)i * ny + jn] + sgn * g : g;
        }
        __threadfence_block();
        named_sync(BAR_ST, NT_ST);
    }
}

template <int BR, int BC, int NBR, int NBC>
struct Cfg {
    static constexpr int NX = NBR * BR + 2, NY = NBC * BC + 2;
    static constexpr int NB = NBR * NBC;
    static constexpr int NCELL = BR * BC, NCH = (NCELL + 1) / 2, NSLOT = 2 * BC + 2 * BR;
    static constexpr size_t CS_BYTES = sizeof(double2) * NCH * NT_SOR;
    static constexpr size_t H_BYTES = sizeof(double) * NSLOT * NT_SOR;
    static constexpr size_t RING_BYTES = sizeof(double) * RING * 4 * NY;
    static constexpr size_t SMEM_BYTES = CS_BYTES + H_BYTES + RING_BYTES;
    static_assert(NB <= NT_SOR, "one SOR thread per block");
    static_assert(NY == NT_ST, "the stencil role maps one thread to one column");
    static_assert((NY * sizeof(double)) % 16 == 0, "bulk copies need 16-byte rows");
};

// ----------------------------------------------------------------------------------------------
// stencil role: row streamer.  Rows of F fields travel global -> shared memory in groups of GR rows;
// a group of one field is contiguous in global memory, so it is ONE bulk copy (TMA unit) per field and
// group (measured on B200: ~75 cycles of issue per bulk copy and ~200 per mbarrier round trip, so the
// copies must be few and large -- scripts/micro/tma_ring.cu).  The ring holds NG = 2 groups (double
// buffer); each group has a "full" mbarrier (transaction bytes) and an "empty" mbarrier (one arrival per
// stencil warp).  Group g of a pass has sequence number g0 + g: slot = seq % NG, parity = (seq / NG) & 1.
// There is NO CTA-wide barrier in the row loops: the four stencil warps only meet through the ring.
// A row needs the row below it, so a step works on rows [GR*g - 1, GR*g + GR - 1): the east / west
// operands of the lagging row GR*g - 1 are saved in registers before its group is released.
// ----------------------------------------------------------------------------------------------
template <int NY, int NG, int GR>
struct Ring {
    double *buf;          // [NG][4][GR][NY]
    uint64_t *full;       // [NG]
    uint64_t *empty;      // [NG]
    unsigned g0;          // sequence number of group 0 of the current pass

    // row r (0 <= r < GR) of field f of group g
    __device__ __forceinline__ const double *row(unsigned g, int f, int r) const {
        return buf + ((((g0 + g) % NG) * 4 + f) * GR + r) * NY;
    }
    // producer lane: refill the slot of group g with rows [g*GR, g*GR+GR) of F fields
    template <int F>
    __device__ __forceinline__ void issue(unsigned g, const double *const (&src)[F]) const {
        const unsigned seq = g0 + g;
        if (seq >= (unsigned)NG) mbar_wait(empty + seq % NG, ((seq / NG) - 1u) & 1u);   // previous use consumed by all warps
        uint64_t *b = full + seq % NG;
        constexpr uint32_t bytes = (uint32_t)(GR * NY * sizeof(double));
        mbar_expect_tx(b, F * bytes);
#pragma unroll
        for (int f = 0; f < F; ++f)
            bulk_g2s(const_cast<double *>(row(g, f, 0)), src[f] + (size_t)g * GR * NY, bytes, b);
    }
    __device__ __forceinline__ void wait_full(unsigned g) const {
        const unsigned seq = g0 + g;
        mbar_wait(full + seq % NG, (seq / NG) & 1u);
    }
    __device__ __forceinline__ void release(unsigned g) const {     // whole warp: its reads of group g are done
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + (g0 + g) % NG);
    }
    // every stencil thread pulls one 128-byte line of group g of F fields towards L2 (F*GR*NY*8/128 <= 128 lines)
    template <int F>
    __device__ __forceinline__ void prefetch_l2(unsigned g, const double *const (&src)[F], int ts) const {
        constexpr int LPF = GR * NY * (int)sizeof(double) / 128;      // lines per field and group
#pragma unroll
        for (int t = ts; t < F * LPF; t += NT_ST) {
            const int f = t / LPF, l = t - f * LPF;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(src[f] + (size_t)g * GR * NY) + l * 128));
        }
    }
};

// C' = beta/den * (dx rho dy^2/dt (ui[i,j]-ui[i-1,j]) + dy rho dx^2/dt (vi[i,j]-vi[i,j-1]))  (chorin_fd:186-188)
// stored at the position the owning SOR thread expects (image of its shared-memory chunks).
template <typename C>
__device__ __forceinline__ void store_cprime(double *img, const short *tidmap, const short *s_ord, int i, int j, double c) {
    const int bi = (i - 1) / C::BRc, li = (i - 1) - bi * C::BRc;
    const int bj = (j - 1) / C::BCc, lj = (j - 1) - bj * C::BCc;
    const int t = tidmap[bi * C::NBCc + bj], q = s_ord[li * C::BCc + lj];
    img[((size_t)(q >> 1) * NT_SOR + t) * 2 + (q & 1)] = c;
}

// pass 1: predictor of member m -> un, vn (global) and the C' image.
template <typename C, int NGR, int GR>
__device__ void stencil_pass1(const StreamArgs &a, Ring<C::NY, NGR, GR> &ring, const short *s_ord, const short *s_tid, int m, int mnext, int ts,
                              double *img) {
#ifdef NNS_ABL_NOSTENCIL      // timing ablation: the SOR role alone on the SM
    return;
#endif
    constexpr int NX = C::NX, NY = C::NY, NGROUPS = NX / GR;
    static_assert(NX % GR == 0, "rows must fill whole ring groups");
    const size_t N = (size_t)NX * NY;
    const double *src[4] = {a.uc + m * N, a.vc + m * N, a.up + m * N, a.vp + m * N};
    double *un = a.un + m * N, *vn = a.vn + m * N;
    const double nu = a.nu_b ? a.nu_b[m] : a.g.nu;
    const double *bcval = a.bcval ? a.bcval + (size_t)m * a.n_bcs : nullptr;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy, rho = a.g.rho, beta = a.g.beta;
    const double dx2 = dx * dx, dy2 = dy * dy;
    // u' = u - dt (3/2 Adv(u^n) - 1/2 Adv(u^{n-1})) + dt nu (3/2 Lap(u^n) - 1/2 Lap(u^{n-1}))  (chorin_fd:63-91)
    // with every constant folded into one coefficient per term (36 FP64 operations per cell)
    const double a0x = 1.5 * dt / (2.0 * dx), a0y = 1.5 * dt / (2.0 * dy);     // AB2 weights of the advection speeds
    const double a1x = 0.5 * dt / (2.0 * dx), a1y = 0.5 * dt / (2.0 * dy);
    const double c0x = 1.5 * dt * nu / dx2, c0y = 1.5 * dt * nu / dy2;         // AB2 weights of the Laplacians
    const double c1x = 0.5 * dt * nu / dx2, c1y = 0.5 * dt * nu / dy2;
    const double den = 2.0 * dx2 + 2.0 * dy2;
    const double cc = beta / den, cu = dx * rho * dy2 / dt, cv = dy * rho * dx2 / dt;
    const int j = ts, lane = ts & 31;
    const bool jin = j > 0 && j < NY - 1;
    const int bj = jin ? (j - 1) / C::BCc : 0, lj = jin ? (j - 1) - bj * C::BCc : 0;

    NNS_FINE(const bool lead = ts == 0; long long tp0 = NNS_PROF_T(), twait = 0;)
    // pull the next member's pressure towards L2 while we are at it (the SOR role loads it soon)
    if (mnext >= 0) {
        const char *pp = reinterpret_cast<const char *>(a.p + (size_t)mnext * N);
        for (size_t off = (size_t)ts * 128; off < N * sizeof(double); off += (size_t)NT_ST * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + off));
    }
    if (ts == 0) {
#pragma unroll
        for (int g = 0; g < NGR; ++g) ring.template issue<4>(g, src);
    }
    ring.template prefetch_l2<4>(NGR, src, ts);
    ring.template prefetch_l2<4>(NGR + 1, src, ts);

    double uN = 0, uC = 0, uS = 0, vN = 0, vC = 0, vS = 0, aN = 0, aC = 0, aS = 0, bN = 0, bC = 0, bS = 0;
    double uE = 0, uW = 0, vE = 0, vW = 0, aE = 0, aW = 0, bE = 0, bW = 0;   // east / west operands of the current row
    double ru_prev = 0.0;
    int bi = 0, li = -1, tsor = 0;           // block row / row inside the block of the current row, owning SOR thread

    // one row: (uC..) is row i, (uS..) row i+1, (uE, uW..) the east / west neighbours of row i
    // (a branch-free variant -- selects instead of the interior test -- measured slower: 4.79 vs 4.60 ms/step)
    auto do_row = [&](int i) {
        double ru = uC, rv = vC;
        const bool interior = jin && i > 0 && i < NX - 1;
        if (interior) {
            // both advection terms difference along axis 0 (chorin_fd:74,76,83,85)
            const double k0 = fma(uC, a0x, vC * a0y), k1 = fma(aC, a1x, bC * a1y);
            const double lu = fma(-2.0, uC, uS + uN), mu = fma(-2.0, uC, uE + uW);
            const double la = fma(-2.0, aC, aS + aN), ma = fma(-2.0, aC, aE + aW);
            const double lv = fma(-2.0, vC, vS + vN), mv = fma(-2.0, vC, vE + vW);
            const double lb = fma(-2.0, bC, bS + bN), mb = fma(-2.0, bC, bE + bW);
            ru = fma(-c1y, ma, fma(-c1x, la, fma(c0y, mu, fma(c0x, lu, fma(k1, aS - aN, fma(-k0, uS - uN, uC))))));
            rv = fma(-c1y, mb, fma(-c1x, lb, fma(c0y, mv, fma(c0x, lv, fma(k1, bS - bN, fma(-k0, vS - vN, vC))))));
        }
        un[(size_t)i * NY + j] = ru;
        vn[(size_t)i * NY + j] = rv;
        // vi of the western neighbour: the lane below (lane 0 of a warp has no such lane: its column, like
        // row 1 and column 1 whose operands are boundary lines, is patched after the BC pass)
        const double rv_w = __shfl_up_sync(0xffffffffu, rv, 1);
        if (i >= 1 && i < NX - 1) {
            if (++li == C::BRc) { li = 0; ++bi; }
            if (li == 0 && jin) tsor = s_tid[bi * C::NBCc + bj];
            if (jin && lane > 0) {
                const double c = cc * (cu * (ru - ru_prev) + cv * (rv - rv_w));
                const int q = s_ord[li * C::BCc + lj];
                img[((size_t)(q >> 1) * NT_SOR + tsor) * 2 + (q & 1)] = c;
            }
        }
        ru_prev = ru;
    };
#ifndef NNS_ST_BRANCHY
    // rows 1 .. NX-2 without a branch (every group but the first holds interior rows only): the unrolled rows of a
    // group form ONE basic block, so that ptxas can interleave their dependency chains
    auto do_row_flat = [&](int i) {
        const double k0 = fma(uC, a0x, vC * a0y), k1 = fma(aC, a1x, bC * a1y);
        const double lu = fma(-2.0, uC, uS + uN), mu = fma(-2.0, uC, uE + uW);
        const double la = fma(-2.0, aC, aS + aN), ma = fma(-2.0, aC, aE + aW);
        const double lv = fma(-2.0, vC, vS + vN), mv = fma(-2.0, vC, vE + vW);
        const double lb = fma(-2.0, bC, bS + bN), mb = fma(-2.0, bC, bE + bW);
        const double ruc = fma(-c1y, ma, fma(-c1x, la, fma(c0y, mu, fma(c0x, lu, fma(k1, aS - aN, fma(-k0, uS - uN, uC))))));
        const double rvc = fma(-c1y, mb, fma(-c1x, lb, fma(c0y, mv, fma(c0x, lv, fma(k1, bS - bN, fma(-k0, vS - vN, vC))))));
        const double ru = jin ? ruc : uC, rv = jin ? rvc : vC;
        un[(size_t)i * NY + j] = ru;
        vn[(size_t)i * NY + j] = rv;
        const double rv_w = __shfl_up_sync(0xffffffffu, rv, 1);
        const bool wrap = li == C::BRc - 1;
        li = wrap ? 0 : li + 1;
        bi += wrap;
        if (wrap) tsor = s_tid[bi * C::NBCc + bj];
        const double c = cc * (cu * (ru - ru_prev) + cv * (rv - rv_w));
        const int q = s_ord[li * C::BCc + lj];
        if (jin && lane > 0) img[((size_t)(q >> 1) * NT_SOR + tsor) * 2 + (q & 1)] = c;
        ru_prev = ru;
    };
#endif
    const int jw = j > 0 ? j - 1 : j, je = j < NY - 1 ? j + 1 : j;
    NNS_FINE(NNS_PROF_ADD(16, tp0); tp0 = NNS_PROF_T();)
    auto do_group = [&](int g, auto flat) {
        NNS_FINE(const long long tw0 = NNS_PROF_T();)
        ring.wait_full(g);
        NNS_FINE(if (a.prof) twait += clock64() - tw0;)
#pragma unroll
        for (int r = -1; r < GR - 1; ++r) {              // rows GR*g - 1 .. GR*g + GR - 2
            const int i = g * GR + r;
            // shift the window: row i becomes current, row i+1 (slot row r+1 of this group) is loaded
            uN = uC; vN = vC; aN = aC; bN = bC;
            uC = uS; vC = vS; aC = aS; bC = bS;
            uS = ring.row(g, 0, r + 1)[j]; vS = ring.row(g, 1, r + 1)[j]; aS = ring.row(g, 2, r + 1)[j]; bS = ring.row(g, 3, r + 1)[j];
            if (r >= 0) {                                  // east / west of row i: in this group (r == -1: saved registers)
                uE = ring.row(g, 0, r)[je]; uW = ring.row(g, 0, r)[jw]; vE = ring.row(g, 1, r)[je]; vW = ring.row(g, 1, r)[jw];
                aE = ring.row(g, 2, r)[je]; aW = ring.row(g, 2, r)[jw]; bE = ring.row(g, 3, r)[je]; bW = ring.row(g, 3, r)[jw];
            }
#ifndef NNS_ST_BRANCHY
            if (decltype(flat)::value) do_row_flat(i);
            else
#endif
            if (i >= 0) do_row(i);
        }
        // east / west of the group's last row, needed by the next step after this slot is refilled
        uE = ring.row(g, 0, GR - 1)[je]; uW = ring.row(g, 0, GR - 1)[jw]; vE = ring.row(g, 1, GR - 1)[je]; vW = ring.row(g, 1, GR - 1)[jw];
        aE = ring.row(g, 2, GR - 1)[je]; aW = ring.row(g, 2, GR - 1)[jw]; bE = ring.row(g, 3, GR - 1)[je]; bW = ring.row(g, 3, GR - 1)[jw];
        ring.release(g);
        if (g + NGR < NGROUPS) {
            NNS_FINE(const long long ti0 = NNS_PROF_T();)
            if (ts == 0) ring.template issue<4>(g + NGR, src);
            NNS_FINE(if (a.prof && ts == 0) a.prof[(size_t)blockIdx.x * NPROF + 6] += clock64() - ti0;)
            if (g + NGR + 2 < NGROUPS) ring.template prefetch_l2<4>(g + NGR + 2, src, ts);
        }
    };
#ifndef NNS_ST_BRANCHY
    do_group(0, std::false_type{});
    for (int g = 1; g < NGROUPS; ++g) do_group(g, std::true_type{});
#else
    for (int g = 0; g < NGROUPS; ++g) do_group(g, std::false_type{});
#endif
    uN = uC; vN = vC; uC = uS; vC = vS;                    // last row (an edge: copied)
    do_row(NX - 1);
    ring.g0 += NGROUPS;
    NNS_FINE(if (a.prof && ts == 32) a.prof[(size_t)blockIdx.x * NPROF + 7] += clock64() - tp0;)
    NNS_FINE(NNS_PROF_ADD(18, tp0); if (a.prof && lead) a.prof[(size_t)blockIdx.x * NPROF + 17] += twait; tp0 = NNS_PROF_T();)
    __threadfence_block();
    named_sync(BAR_ST, NT_ST);
    st_apply_bc_global(un, NX, NY, a.ubc, bcval, dx, dy, ts);
    st_apply_bc_global(vn, NX, NY, a.vbc, bcval, dx, dy, ts);
    NNS_FINE(NNS_PROF_ADD(19, tp0); tp0 = NNS_PROF_T();)
    // patch C' where an operand is a boundary line (row 1, column 1) or belongs to another warp's lane 31
    auto patch = [&](int i, int jj) {
        const size_t gq = (size_t)i * NY + jj;
        const double c = cc * (cu * (un[gq] - un[gq - NY]) + cv * (vn[gq] - vn[gq - 1]));
        store_cprime<C>(img, s_tid, s_ord, i, jj, c);
    };
    if (jin) patch(1, j);
    for (int i = 2 + ts; i < NX - 1; i += NT_ST) {
        patch(i, 1);
#pragma unroll
        for (int w = 1; w < NT_ST / 32; ++w)
            if (32 * w < NY - 1) patch(i, 32 * w);
    }
    __threadfence();                      // the image is pulled through L2 (cp.async.cg) by the SOR role
    NNS_FINE(NNS_PROF_ADD(20, tp0);)
}

// pass 2: p_bc, projection and trajectory snapshot of member m.
// psrc: where the SOR role left its result -- the member's own p (legacy kernel: in place) or the CTA's scratch
// image (wave kernel: the edges are copied in from p first, and every row of the final p is written back here,
// coalesced, so that the member's p stays untouched until its sweep count is known).
template <typename C, int NGR, int GR, bool SCR>
__device__ void stencil_pass2(const StreamArgs &a, Ring<C::NY, NGR, GR> &ring, int m, int ts, double *pscr = nullptr) {
#ifdef NNS_ABL_NOSTENCIL
    return;
#endif
    constexpr int NX = C::NX, NY = C::NY, NGROUPS = NX / GR;
    const size_t N = (size_t)NX * NY;
    double *pout = a.p + m * N, *un = a.un + m * N, *vn = a.vn + m * N;
    double *pg = SCR ? pscr : pout;
    NNS_FINE(const bool lead = ts == 0; long long tp0 = NNS_PROF_T(), twait = 0;)
    const double *bcval = a.bcval ? a.bcval + (size_t)m * a.n_bcs : nullptr;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy;
    if (SCR) {
        static_assert(NX == NT_ST && NY == NT_ST, "edge copy maps one thread to one edge cell per side");
        pg[ts] = pout[ts];
        pg[(size_t)(NX - 1) * NY + ts] = pout[(size_t)(NX - 1) * NY + ts];
        pg[(size_t)ts * NY] = pout[(size_t)ts * NY];
        pg[(size_t)ts * NY + NY - 1] = pout[(size_t)ts * NY + NY - 1];
        __threadfence_block();
        named_sync(BAR_ST, NT_ST);
    }
    st_apply_bc_global(pg, NX, NY, a.pbc, bcval, dx, dy, ts);
    __threadfence();                      // generic-proxy writes of p / un / vn (this CTA) before the bulk reads
    named_sync(BAR_ST, NT_ST);
    const double *src[3] = {pg, un, vn};
    const int j = ts;
    const bool jin = j > 0 && j < NY - 1;
    const int jw = j > 0 ? j - 1 : j, je = j < NY - 1 ? j + 1 : j;
    const double kx = dt / (2.0 * dx), ky = dt / (2.0 * dy);
    const size_t toff = (size_t)m * a.traj_member_stride + a.traj_off;
    if (ts == 0) {
        fence_proxy_async();
#pragma unroll
        for (int g = 0; g < NGR; ++g) ring.template issue<3>(g, src);
    }
    double pN = 0, pC = 0, pS = 0, pE = 0, pW = 0, ruC = 0, rvC = 0, ruS = 0, rvS = 0;
    unsigned long long bad = 0;
    auto do_row = [&](int i) {
        double ru = ruC, rv = rvC;
        const size_t q = (size_t)i * NY + j;
        if (jin && i > 0 && i < NX - 1) {
            ru -= kx * (pS - pN);
            rv -= ky * (pE - pW);
            un[q] = ru;
            vn[q] = rv;
        }
        if (SCR) pout[q] = pC;
        if (a.traj_u) { a.traj_u[toff + q] = ru; a.traj_v[toff + q] = rv; a.traj_p[toff + q] = pC; }
        if (a.flags & NNS_FLAG_CHECK_FINITE) bad += !(isfinite(ru) && isfinite(rv) && isfinite(pC));
    };
    NNS_FINE(NNS_PROF_ADD(21, tp0); tp0 = NNS_PROF_T();)
    for (int g = 0; g < NGROUPS; ++g) {
        NNS_FINE(const long long tw0 = NNS_PROF_T();)
        ring.wait_full(g);
        NNS_FINE(if (a.prof) twait += clock64() - tw0;)
#pragma unroll
        for (int r = -1; r < GR - 1; ++r) {
            const int i = g * GR + r;
            pN = pC; pC = pS; ruC = ruS; rvC = rvS;
            pS = ring.row(g, 0, r + 1)[j]; ruS = ring.row(g, 1, r + 1)[j]; rvS = ring.row(g, 2, r + 1)[j];
            if (r >= 0) { pE = ring.row(g, 0, r)[je]; pW = ring.row(g, 0, r)[jw]; }
            if (i >= 0) do_row(i);
        }
        pE = ring.row(g, 0, GR - 1)[je]; pW = ring.row(g, 0, GR - 1)[jw];
        ring.release(g);
        if (ts == 0 && g + NGR < NGROUPS) ring.template issue<3>(g + NGR, src);
    }
    pN = pC; pC = pS; ruC = ruS; rvC = rvS;
    do_row(NX - 1);
    ring.g0 += NGROUPS;
    NNS_FINE(NNS_PROF_ADD(23, tp0); if (a.prof && lead) a.prof[(size_t)blockIdx.x * NPROF + 22] += twait;)
    if ((a.flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(a.nonfinite, bad);
}

// ----------------------------------------------------------------------------------------------
template <int BR, int BC, int NBR, int NBC>
struct CfgX : Cfg<BR, BC, NBR, NBC> {
    static constexpr int BRc = BR, BCc = BC, NBRc = NBR, NBCc = NBC;
    static constexpr int RSc = (BR + 1) / 2;       // rows of the top sub-block
};

// LIST: the members are list[0 .. *list_count) (re-run pass of the wave kernel) instead of 0 .. count
template <typename C, bool LIST>
__global__ void __launch_bounds__(NT_SOR + NT_ST, 1) chorin_stream_kernel(const StreamArgs a) {
    constexpr int BR = C::BRc, BC = C::BCc, RS = C::RSc, NX = C::NX, NY = C::NY, NCH = C::NCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_mask[2];
    __shared__ int s_need;
    __shared__ __align__(8) uint64_t s_full[NG], s_empty[NG];
    __shared__ short s_ord[128];          // split_ord of cell li * BC + lj (copy of c_ord: per-lane indices)
    __shared__ __align__(8) uint64_t s_hb[NW_SOR][2];   // SOR role: stage hand-off barriers (wavefront())
    __shared__ __align__(8) uint64_t s_pfull;      // SOR role: the member's p has landed in the (idle) C' region

    double2 *Cs = reinterpret_cast<double2 *>(smem_raw);
    double *H = reinterpret_cast<double *>(smem_raw + C::CS_BYTES);
    double *ringbuf = reinterpret_cast<double *>(smem_raw + C::CS_BYTES + C::H_BYTES);

    const int tid = threadIdx.x;
    const size_t N = (size_t)NX * NY;
    const int total = LIST ? *a.list_count : a.count;
    const int nmine = total > (int)blockIdx.x ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto member = [&](int k) { const int idx = (int)blockIdx.x + k * (int)gridDim.x; return LIST ? a.list[idx] : idx; };
    double *img = a.cimg + (size_t)blockIdx.x * 2 * NCH * NT_SOR;

    if (tid == 0) {
        for (int r = 0; r < NG; ++r) { mbar_init(&s_full[r], 1); mbar_init(&s_empty[r], NT_ST / 32); }
        mbar_init(&s_pfull, 1);
        for (int w = 0; w < NW_SOR; ++w) { mbar_init(&s_hb[w][0], __popc(a.depmask[w])); mbar_init(&s_hb[w][1], __popc(a.depmask[w])); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_mask[0] = 0ull; s_mask[1] = 0ull;
    }
    if (tid < 128) s_ord[tid] = c_ord[tid];
    __syncthreads();
    if (nmine == 0) return;
#ifdef NNS_SOR_TMEM
    __shared__ uint32_t s_tmem;
    if (tid < 32) {         // the whole Tensor Memory of the SM (one CTA per SM): C' of the SOR threads, one lane per thread
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // lanes 32 * (warp % 4) .. + 31 belong to the warp; SOR warps w and w + 4 share a lane quarter: columns 0 / 128
    const uint32_t tm_mine = s_tmem + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(128 * ((tid >> 5) >> 2));
#else
    const uint32_t tm_mine = 0u;
#endif

    if (tid >= NT_SOR) {
        // =============================== stencil role ===========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
        const int ts = tid - NT_SOR;
        Ring<NY, NG, GR> ring{ringbuf, s_full, s_empty, 0u};
        const bool lead = ts == 0;
        long long t0 = NNS_PROF_T();
        stencil_pass1<C, NG, GR>(a, ring, s_ord, a.tidmap, member(0), nmine > 1 ? member(1) : -1, ts, img);
        NNS_PROF_ADD(8, t0);
        named_arrive(BAR_READY + 0, NT_SOR + NT_ST);
        for (int k = 0; k < nmine; ++k) {
            const int m = member(k);
            if (k + 1 < nmine) {
                t0 = NNS_PROF_T();
                named_sync(BAR_CONSUMED + (k & 1), NT_SOR + NT_ST);       // the SOR role has pulled image k
                NNS_PROF_ADD(9, t0);
                t0 = NNS_PROF_T();
                stencil_pass1<C, NG, GR>(a, ring, s_ord, a.tidmap, member(k + 1), k + 2 < nmine ? member(k + 2) : -1, ts, img);
                NNS_PROF_ADD(8, t0);
                named_arrive(BAR_READY + ((k + 1) & 1), NT_SOR + NT_ST);
            }
            t0 = NNS_PROF_T();
            named_sync(BAR_DONE + (k & 1), NT_SOR + NT_ST);               // SOR of member k finished, p written
            NNS_PROF_ADD(10, t0);
            t0 = NNS_PROF_T();
            stencil_pass2<C, NG, GR, false>(a, ring, m, ts);
            NNS_PROF_ADD(11, t0);
        }
    } else {
        // ================================= SOR role =============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR));
        SBlock ds = a.desc[tid];
        const bool owner = ds.r0 > 0;
        if (!owner) { ds.r0 = 1; ds.c0 = 1; ds.bd = (short)(tid >= NT_SOR / 2); }     // (diagonal parity of the warp: uniform top / bottom choice)
        const int r0 = ds.r0, c0 = ds.c0;
        SHalo<BR, BC> h;
        h.Hme = H + tid;
        h.pubT = ds.nN >= 0; h.pubB = ds.nS >= 0; h.pubL = ds.nW >= 0; h.pubR = ds.nE >= 0;
        h.hN = h.pubT ? H + BC * NT_SOR + ds.nN : H + tid;
        h.hS = h.pubB ? H + ds.nS : H + BC * NT_SOR + tid;
        h.hW = h.pubL ? H + (2 * BC + BR) * NT_SOR + ds.nW : H + 2 * BC * NT_SOR + tid;
        h.hE = h.pubR ? H + 2 * BC * NT_SOR + ds.nE : H + (2 * BC + BR) * NT_SOR + tid;
        const double2 *Cme = Cs + tid;

        const double dx = a.g.dx, dy = a.g.dy, beta = a.g.beta;
        const double dx2 = dx * dx, dy2 = dy * dy, den = 2.0 * dx2 + 2.0 * dy2;
        Coef k;
        k.ca = beta * dy2 / den; k.cb = beta * dx2 / den; k.cu = 0; k.cv = 0; k.beta = beta; k.tol = a.g.tol;
#ifdef NNS_SOR_FORM_PN
        k.cc = 1.0 - beta;
#else
        k.cc = -beta;
#endif
        const unsigned long long tolbits = (unsigned long long)__double_as_longlong(a.g.tol);
        const int cap = a.g.nit - 1;
        const int tmax = 2 * C::NBRc + C::NBCc - 2 + 2 * (cap - 1);      // last sub-block diagonal + 2 (cap - 1)

        int gbase = 0;                     // global stage index of the hand-off barriers (wavefront())
        for (int kk = 0; kk < nmine; ++kk) {
            const int m = member(kk);
            double *pg = a.p + (size_t)m * N;
            double P[BR][BC];
            auto load_block = [&]() {
#pragma unroll
                for (int li = 0; li < BR; ++li)
#pragma unroll
                    for (int lj = 0; lj < BC; ++lj) P[li][lj] = owner ? pg[(size_t)(r0 + li) * NY + c0 + lj] : 0.0;
            };
            const bool lead = tid == 0;
            long long t0 = NNS_PROF_T();
#ifndef NNS_SOR_DIRECT_P
            // The member's p travels global -> shared memory as ONE bulk copy into the C' region (idle between two
            // members' sweeps) and from there into the register blocks: a thread's block is 9 rows x 56 bytes, so the
            // direct loads are one 32-byte sector per lane and instruction (63 x 256 sectors, ~16 k cycles per member).
            static_assert(sizeof(double) * (size_t)NX * NY <= C::CS_BYTES, "p must fit the C' region");
            const double *ps = reinterpret_cast<const double *>(Cs);
            if (tid == 0 && kk == 0) {             // (later members: issued behind the previous member's store, below)
                fence_proxy_async();               // earlier generic accesses of the region before the async-proxy write
                mbar_expect_tx(&s_pfull, (uint32_t)(sizeof(double) * N));
                bulk_g2s(Cs, pg, (uint32_t)(sizeof(double) * N), &s_pfull);
            }
#ifdef NNS_SOR_TMEM
            mbar_wait(&s_pfull, kk & 1);           // (C' does not pass through the region: one phase per member)
#else
            mbar_wait(&s_pfull, 0);                // phases of s_pfull alternate: p (parity 0), C' image (parity 1)
#endif
#pragma unroll
            for (int li = 0; li < BR; ++li)
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) P[li][lj] = owner ? ps[(r0 + li) * NY + c0 + lj] : 0.0;
#else
            const double *ps = pg;
            load_block();                          // p is not touched by the stencil role before SOR finishes
#endif
            if (owner) {                           // frozen boundary values into the unread own slots
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) {
                    if (!h.pubT) h.Hme[lj * NT_SOR] = ps[(size_t)(r0 - 1) * NY + c0 + lj];
                    if (!h.pubB) h.Hme[(BC + lj) * NT_SOR] = ps[(size_t)(r0 + BR) * NY + c0 + lj];
                }
#pragma unroll
                for (int li = 0; li < BR; ++li) {
                    if (!h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = ps[(size_t)(r0 + li) * NY + c0 - 1];
                    if (!h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = ps[(size_t)(r0 + li) * NY + c0 + BC];
                }
            }
            NNS_PROF_ADD(0, t0);
            t0 = NNS_PROF_T();
            named_sync(BAR_READY + (kk & 1), NT_SOR + NT_ST);            // C' image of member kk is complete
            NNS_PROF_ADD(1, t0);
            t0 = NNS_PROF_T();
#ifdef NNS_SOR_TMEM
            {   // the thread's 32 chunks of the image -> its Tensor Memory lane (coalesced 16-byte loads across the warp)
                const double2 *gi = reinterpret_cast<const double2 *>(img) + tid;
#pragma unroll 2
                for (int c = 0; c < NCH; c += 4) {
                    double2 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = __ldcg(gi + (size_t)(c + e) * NT_SOR);
                    tm_st16_top(tm_mine + 4 * c, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            named_sync(BAR_SOR, NT_SOR);
#elif !defined(NNS_SOR_DIRECT_P)
            // the image has the layout of the C' region: one bulk copy (every SOR thread has left the region: BAR_READY)
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&s_pfull, (uint32_t)C::CS_BYTES);
                bulk_g2s(Cs, img, (uint32_t)C::CS_BYTES, &s_pfull);
            }
            mbar_wait(&s_pfull, 1);
#else
            {
                const double2 *gi = reinterpret_cast<const double2 *>(img) + tid;
#pragma unroll
                for (int c = 0; c < NCH; ++c) cp_async16(Cs + c * NT_SOR + tid, gi + c * NT_SOR);
                cp_async_wait_all();
            }
            named_sync(BAR_SOR, NT_SOR);
#endif
            if (kk + 1 < nmine) named_arrive(BAR_CONSUMED + (kk & 1), NT_SOR + NT_ST);
            NNS_PROF_ADD(2, t0);
            t0 = NNS_PROF_T();

            int need = 0;
            if (cap > 0) {
                // reduce the per-thread sweep flags to the number of sweeps the reference loop runs;
                // returns -1 if the fast test left the deciding sweep undecided
                auto sweeps_needed = [&](unsigned long long mask, unsigned long long amb) -> int {
                    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mask);
                    unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mask >> 32));
                    unsigned alo = __reduce_or_sync(0xffffffffu, (unsigned)amb);
                    unsigned ahi = __reduce_or_sync(0xffffffffu, (unsigned)(amb >> 32));
                    if ((tid & 31) == 0) {
                        atomicOr(&s_mask[0], ((unsigned long long)hi << 32) | lo);
                        atomicOr(&s_mask[1], ((unsigned long long)ahi << 32) | alo);
                    }
                    named_sync(BAR_SOR, NT_SOR);
                    if (tid == 0) {
                        const unsigned long long full = cap >= 64 ? ~0ull : ((1ull << cap) - 1ull);
                        const unsigned long long viol = s_mask[0], und = s_mask[1] & ~viol & full;
                        const unsigned long long clr = ~viol & full;       // sweeps without a certain violation
                        int nd = clr ? __ffsll((long long)clr) : cap;      // first such sweep: s + 1 sweeps run
                        // the first sweep without a certain violation decides; if it is merely undecided, redo exactly
                        if (clr && ((und >> (nd - 1)) & 1ull)) nd = -1;
                        s_need = nd;
                        s_mask[0] = 0ull; s_mask[1] = 0ull;
                    }
                    named_sync(BAR_SOR, NT_SOR);
                    return s_need;
                };
                unsigned long long mask = 0ull, amb = 0ull;
                wavefront<BR, BC, RS, 1>(P, Cme, h, owner, ds.bd, tmax, cap, k, tolbits, mask, amb, s_hb, a.depmask[tid >> 5], gbase, tm_mine,
                                     a.prof && (tid == 0 || tid == 128) ? a.prof + (size_t)blockIdx.x * NPROF + 12 + (tid >> 6) : nullptr,
                                     a.trace && blockIdx.x == 0 && kk == 1 ? a.trace + (size_t)(tid >> 5) * TRACE_STAGES * 4 : nullptr);
                NNS_PROF_ADD(3, t0);
                t0 = NNS_PROF_T();
                need = sweeps_needed(mask, amb);
#ifdef NNS_ABL_NOSTENCIL
                need = cap;
#endif
                if (need < 0) {
                    // max|dp| of the deciding sweep shares its high word with tol: repeat with the exact test
                    load_block();
                    mask = 0ull; amb = 0ull;
                    wavefront<BR, BC, RS, 2>(P, Cme, h, owner, ds.bd, tmax, cap, k, tolbits, mask, amb, s_hb, a.depmask[tid >> 5], gbase, tm_mine);
                    need = sweeps_needed(mask, 0ull);
                }
                if (need < cap) {
                    // the sequential loop would have stopped after `need` sweeps: redo from the
                    // untouched global p with the sweep count capped (rare: near steady state)
                    load_block();
                    unsigned long long d0 = 0ull, d1 = 0ull;
                    wavefront<BR, BC, RS, 0>(P, Cme, h, owner, ds.bd, 2 * C::NBRc + C::NBCc - 2 + 2 * (need - 1), need, k, tolbits,
                                         d0, d1, s_hb, a.depmask[tid >> 5], gbase, tm_mine);
                }
            }
#ifndef NNS_SOR_DIRECT_P
            NNS_PROF_ADD(4, t0);
            t0 = NNS_PROF_T();
            {
                // blocks -> image of rows 1 .. NX-2 in the C' region (dead after the sweeps) -> ONE bulk copy to global.
                // The image rows are whole rows: the frozen columns 0 and NY-1 come from the boundary blocks' own slots.
                double *pw = reinterpret_cast<double *>(Cs);
                if (owner) {
#pragma unroll
                    for (int li = 0; li < BR; ++li) {
#pragma unroll
                        for (int lj = 0; lj < BC; ++lj) pw[(r0 + li) * NY + c0 + lj] = P[li][lj];
                        if (!h.pubL) pw[(r0 + li) * NY + c0 - 1] = h.Hme[(2 * BC + li) * NT_SOR];
                        if (!h.pubR) pw[(r0 + li) * NY + c0 + BC] = h.Hme[(2 * BC + BR + li) * NT_SOR];
                    }
                }
                fence_proxy_async();               // generic-proxy writes before the async-proxy read of the bulk copy
                named_sync(BAR_SOR, NT_SOR);
                if (tid == 0) {
                    bulk_s2g(pg + NY, pw + NY, (uint32_t)(sizeof(double) * (NX - 2) * NY));
                    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
                    if (kk + 1 < nmine) {          // the region is free again: the next member's p is already on its way
                        mbar_expect_tx(&s_pfull, (uint32_t)(sizeof(double) * N));
                        bulk_g2s(Cs, a.p + (size_t)member(kk + 1) * N, (uint32_t)(sizeof(double) * N), &s_pfull);
                    }
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // complete (not only read): the stencil role re-reads p after BAR_DONE
                }
            }
#else
            if (owner) {
#pragma unroll
                for (int li = 0; li < BR; ++li)
#pragma unroll
                    for (int lj = 0; lj < BC; ++lj) pg[(size_t)(r0 + li) * NY + c0 + lj] = P[li][lj];
            }
            NNS_PROF_ADD(4, t0);
            t0 = NNS_PROF_T();
#endif
            if (a.sweeps && tid == 0) a.sweeps[m] = need;
            __threadfence();              // p is re-read by the stencil role, partly through the TMA unit
            named_arrive(BAR_DONE + (kk & 1), NT_SOR + NT_ST);
            NNS_PROF_ADD(5, t0);
        }
    }
#ifdef NNS_SOR_TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(512));
#endif
}


#ifdef NNS_ENABLE_WAVE      // experiment of round 1: not reproducible bit for bit, not part of the product build
// ----------------------------------------------------------------------------------------------
// Wave kernel: the same two roles, but the SOR wavefront runs CONTINUOUSLY across the members of a CTA.
//
//  * Sub-block (sbi, bj) performs sweep s of the CTA's k-th member at super-stage sbi + bj + 2s + PER*k with
//    PER = 2*(nit-1) + R + 1 (R = largest spread of block diagonals inside one warp): while the blocks at the
//    bottom right finish member k, the blocks at the top left already sweep member k+1.  Members change per
//    WARP (all lanes of a warp work on the same member, lanes outside the band are predicated off), so every
//    branch and every Tensor Memory access is warp-uniform.  113 super-stages per member instead of 141, and
//    no serial load / pull / reduce / store phases.
//  * The right-hand side C' lives in TENSOR MEMORY (256 KB per SM, otherwise idle in this kernel): 128 columns
//    per thread and member, double-buffered; tcgen05.ld in the sweeps (block_sweep_tm), tcgen05.st by the
//    stencil warp of the same SM sub-partition (lane quarter), which forwards the C' image of the next member.
//    The SOR role never touches C' traffic, and the single shared-memory pipe only carries the block halos.
//  * Shared memory: halo slots (64 KiB) + an 8-deep TMA row ring (128 KiB) for the stencil role.
//  * A warp that has finished its member stores its p blocks to an L2-resident scratch image, starts the loads
//    of the next member's blocks and start halos (cp.async straight into the neighbours' slots, which nobody
//    reads any more), and one idle super-stage later publishes its exit-test flags.  The stencil role turns
//    the scratch image into the final p (p_bc, projection) once all eight warps have reported; if the flags say
//    that the reference loop would have stopped early (or the fast test is undecided), the member's p is still
//    untouched and the member is queued for the legacy kernel, which re-runs it with the exact sweep count.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void spin_until_ge(const volatile int *p, int v) {
    while (*p < v) __nanosleep(32);
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const double2 (&v)[4]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)), "r"(__double2hiint(v[0].y)),
        "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)),
        "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
        "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)), "r"(__double2hiint(v[3].y))
        : "memory");
}
__device__ __forceinline__ double2 ld_cg_f64x2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <typename C>
struct WaveSmem {
    static constexpr size_t H_BYTES = C::H_BYTES;
    static constexpr size_t RING_BYTES = sizeof(double) * GRW * NGW * 4 * C::NY;
    static constexpr size_t SMEM_BYTES = H_BYTES + RING_BYTES;
};

template <typename C>
__global__ void __launch_bounds__(NT_SOR + NT_ST, 1) chorin_wave_kernel(const StreamArgs a) {
    constexpr int BR = C::BRc, BC = C::BCc, RS = C::RSc, NX = C::NX, NY = C::NY, NCH = C::NCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_mask[2][2];       // [member parity][violated, undecided] sweep bit sets
    __shared__ volatile int s_cready[NW_SOR];          // members whose C' has been delivered to the warp's Tensor Memory
    __shared__ volatile int s_wdone[NW_SOR];           // members the warp has finished sweeping (its C' buffer is free)
    __shared__ volatile int s_wstored[NW_SOR];         // members whose p blocks and exit flags of the warp are visible
    __shared__ int s_need;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[NGW], s_empty[NGW];
    __shared__ short s_ord[128];
    __shared__ short s_tid[C::NB];

    double *H = reinterpret_cast<double *>(smem_raw);
    double *ringbuf = reinterpret_cast<double *>(smem_raw + WaveSmem<C>::H_BYTES);

    const int tid = threadIdx.x;
    const size_t N = (size_t)NX * NY;
    const int nmine = a.count > (int)blockIdx.x ? (a.count - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    double *img = a.cimg + (size_t)blockIdx.x * 2 * NCH * NT_SOR;
    double *pscr = a.pscr + (size_t)blockIdx.x * 2 * N;
    if (nmine == 0) return;

    if (tid == 0) {
        for (int r = 0; r < NGW; ++r) { mbar_init(&s_full[r], 1); mbar_init(&s_empty[r], NT_ST / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_mask[0][0] = s_mask[0][1] = s_mask[1][0] = s_mask[1][1] = 0ull;
    }
    if (tid < NW_SOR) { s_cready[tid] = 0; s_wdone[tid] = 0; s_wstored[tid] = 0; }
    if (tid < 128) s_ord[tid] = c_ord[tid];
    if (tid < C::NB) s_tid[tid] = a.tidmap[tid];
    if (tid < 32) {         // the whole Tensor Memory of the SM (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const int cap = a.g.nit - 1;

    if (tid >= NT_SOR) {
        // =============================== stencil role ===========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST_W));
        const int ts = tid - NT_SOR, sw = ts >> 5, lane = ts & 31;
        Ring<NY, NGW, GRW> ring{ringbuf, s_full, s_empty, 0u};
        const bool lead = ts == 0;
        // forward the C' image of the CTA's k-th member into the Tensor Memory of the two SOR warps that share
        // this warp's lane quarter (SOR warps sw and sw + 4)
        auto fill = [&](int k) {
            named_sync(BAR_ST, NT_ST);          // the image is complete (global writes of all stencil threads)
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                const int hw = sw + 4 * hh;
                spin_until_ge(&s_wdone[hw], k - 1);         // buffer k & 1 was last read for member k - 2
                const uint32_t taddr = tmem + ((uint32_t)(32 * sw) << 16) + (uint32_t)(256 * hh + 128 * (k & 1));
                const double2 *gi = reinterpret_cast<const double2 *>(img) + hw * 32 + lane;
#pragma unroll 2
                for (int c = 0; c < NCH; c += 4) {
                    double2 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = ld_cg_f64x2(gi + (size_t)(c + e) * NT_SOR);
                    tm_st16(taddr + 4 * c, v);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __threadfence_block();
                __syncwarp();
                if (lane == 0) s_cready[hw] = k + 1;
            }
            named_sync(BAR_ST, NT_ST);          // the image may be overwritten by the next pass 1
        };
        long long t0 = NNS_PROF_T();
        stencil_pass1<C, NGW, GRW>(a, ring, s_ord, s_tid, blockIdx.x, nmine > 1 ? (int)(blockIdx.x + gridDim.x) : -1, ts, img);
        NNS_PROF_ADD(8, t0);
        fill(0);
        for (int k = 0; k < nmine; ++k) {
            const int m = blockIdx.x + k * gridDim.x;
            if (k + 1 < nmine) {
                t0 = NNS_PROF_T();
                stencil_pass1<C, NGW, GRW>(a, ring, s_ord, s_tid, m + gridDim.x, k + 2 < nmine ? (int)(m + 2 * gridDim.x) : -1, ts, img);
                NNS_PROF_ADD(8, t0);
                t0 = NNS_PROF_T();
                fill(k + 1);
                NNS_PROF_ADD(9, t0);
            }
            t0 = NNS_PROF_T();
            if (ts < NW_SOR) spin_until_ge(&s_wstored[ts], k + 1);       // all eight warps have reported member k
            __threadfence();
            named_sync(BAR_ST, NT_ST);
            if (ts == 0) {
                // number of sweeps the reference loop runs (chorin_fd:183-199): the first sweep without a certain
                // violation decides; if it is merely undecided by the fast test the re-run pass decides exactly
                const unsigned long long full = cap >= 64 ? ~0ull : ((1ull << cap) - 1ull);
                const unsigned long long viol = s_mask[k & 1][0], und = s_mask[k & 1][1] & ~viol & full;
                const unsigned long long clr = ~viol & full;
                int nd = clr ? __ffsll((long long)clr) : cap;
                if (clr && ((und >> (nd - 1)) & 1ull)) nd = -1;
                s_need = nd;
                s_mask[k & 1][0] = 0ull; s_mask[k & 1][1] = 0ull;
            }
            named_sync(BAR_ST, NT_ST);
            const int need = s_need;
            NNS_PROF_ADD(10, t0);
            t0 = NNS_PROF_T();
#if defined(NNS_ABL_NOSWEEP) || defined(NNS_ABL_NOSTENCIL)
            if (true) {
#else
            if (need == cap) {
#endif
                stencil_pass2<C, NGW, GRW, true>(a, ring, m, ts, pscr + (size_t)(k & 1) * N);
                if (a.sweeps && ts == 0) a.sweeps[m] = need;
            } else if (ts == 0) {
                a.redo_list[atomicAdd(a.redo_count, 1)] = m;
            }
            NNS_PROF_ADD(11, t0);
        }
    } else {
        // ================================= SOR role =============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR_W));
        SBlock ds = a.desc[tid];
        const bool owner = ds.r0 > 0;
        const int w = tid >> 5, lane = tid & 31;
        if (!owner) { ds.r0 = 1; ds.c0 = 1; }
        const int r0 = ds.r0, c0 = ds.c0;
        SHalo<BR, BC> h;
        h.Hme = H + tid;
        h.pubT = ds.nN >= 0; h.pubB = ds.nS >= 0; h.pubL = ds.nW >= 0; h.pubR = ds.nE >= 0;
        h.hN = h.pubT ? H + BC * NT_SOR + ds.nN : H + tid;
        h.hS = h.pubB ? H + ds.nS : H + BC * NT_SOR + tid;
        h.hW = h.pubL ? H + (2 * BC + BR) * NT_SOR + ds.nW : H + 2 * BC * NT_SOR + tid;
        h.hE = h.pubR ? H + 2 * BC * NT_SOR + ds.nE : H + (2 * BC + BR) * NT_SOR + tid;
        if (!owner) { h.pubT = h.pubB = h.pubL = h.pubR = false; }

        const double dx = a.g.dx, dy = a.g.dy, beta = a.g.beta;
        const double dx2 = dx * dx, dy2 = dy * dy, den = 2.0 * dx2 + 2.0 * dy2;
        Coef k;
        k.ca = beta * dy2 / den; k.cb = beta * dx2 / den; k.cc = -beta; k.cu = 0; k.cv = 0; k.beta = beta; k.tol = a.g.tol;
        const unsigned tolhi = (unsigned)((unsigned long long)__double_as_longlong(a.g.tol) >> 32);
        const int amin = a.wamin[w], last = 2 * cap - 1 + a.wrange[w];      // last local stage with an active lane
        const int per = 2 * cap + a.rmax + 1;
        const int delta = owner ? ds.bd - amin : (1 << 20);                 // the lane's band starts delta stages later
        int tend = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW_SOR; ++w2) tend = max(tend, a.wamin[w2] + 2 * cap + a.wrange[w2]);   // last report stage of a member
        tend += per * (nmine - 1);
        const uint32_t tm_mine = tmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(256 * (w >> 2));
        const bool lead = tid == 0;

        double P[BR][BC];
        // loads of a member's blocks and of the start values on the far side of the block's south and east edges
        // (and of the frozen boundary values for blocks at the wall): registers / cp.async into the halo slots
        auto fetch = [&](int m) {
            const double *pg = a.p + (size_t)m * N;
#pragma unroll
            for (int li = 0; li < BR; ++li)
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) P[li][lj] = pg[(size_t)(r0 + li) * NY + c0 + lj];     // (16-byte accesses cost more in register moves than they save)
            if (owner) {
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) {
                    cp_async8(const_cast<double *>(h.hS) + lj * NT_SOR, pg + (size_t)(r0 + BR) * NY + c0 + lj);
                    if (!h.pubT) cp_async8(const_cast<double *>(h.hN) + lj * NT_SOR, pg + (size_t)(r0 - 1) * NY + c0 + lj);
                }
#pragma unroll
                for (int li = 0; li < BR; ++li) {
                    cp_async8(const_cast<double *>(h.hE) + li * NT_SOR, pg + (size_t)(r0 + li) * NY + c0 + BC);
                    if (!h.pubL) cp_async8(const_cast<double *>(h.hW) + li * NT_SOR, pg + (size_t)(r0 + li) * NY + c0 - 1);
                }
            }
        };
        unsigned long long mask = 0ull, amb = 0ull;
        long long tsw = 0;
        const long long tk0 = NNS_PROF_T();
        // Every warp passes tend + 1 stage barriers: amin before its first member, per for every member (last + 1
        // sweep stages -- an even number: top / bottom pairs --, one report stage, idle stages), the rest at the end.
        // The member changes sit outside the hot loop, which only holds the two sweeps and their barriers.
        auto sweep = [&](int kw, int tl, auto top) {
#ifndef NNS_ABL_NOSWEEP      // timing ablation: the stencil role alone on the SM
            const long long t0 = NNS_PROF_T();
            const int q = tl - delta;
            const bool act = q >= 0 && q <= 2 * cap - 1;
            const int sidx = q >> 1;
            const uint32_t tmc = tm_mine + (uint32_t)(128 * (kw & 1));
            unsigned mhi = 0u;
            if (decltype(top)::value) block_sweep_tm<BR, BC, RS, 0, RS>(P, tmc, h, k, act, sidx != cap - 1, mhi);
            else block_sweep_tm<BR, BC, RS, RS, BR>(P, tmc, h, k, act, sidx != cap - 1, mhi);
            if (act) {
                mask |= (unsigned long long)(mhi > tolhi) << sidx;
                amb |= (unsigned long long)(mhi == tolhi) << sidx;
            }
            if (a.prof && lane == 0) tsw += clock64() - t0;
#endif
        };
        using TopT = std::integral_constant<bool, true>;
        using BotT = std::integral_constant<bool, false>;
        int passed = 0;
        for (; passed < amin; ++passed) named_sync(BAR_SOR, NT_SOR);
        for (int kw = 0; kw < nmine; ++kw) {
            const int m = blockIdx.x + kw * gridDim.x;
            // ---- start of the member (local stage 0)
            if (kw == 0) fetch(m);
            cp_async_wait_all();
            {
                const long long t0 = NNS_PROF_T();
                spin_until_ge(&s_cready[w], kw + 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                NNS_PROF_ADD(1, t0);
                __syncwarp();
            }
            // ---- hot loop: local stages 0 .. last - 2 in top / bottom pairs
#pragma unroll 1
            for (int tl = 0; tl + 1 < last; tl += 2) {
                sweep(kw, tl, TopT{});
                named_sync(BAR_SOR, NT_SOR);
                sweep(kw, tl + 1, BotT{});
                named_sync(BAR_SOR, NT_SOR);
            }
            sweep(kw, last - 1, TopT{});
            named_sync(BAR_SOR, NT_SOR);
            sweep(kw, last, BotT{});
            // ---- local stage `last`: the warp has finished member kw: results to the scratch image, next member's
            // loads in flight
            {
                double *ps = pscr + (size_t)(kw & 1) * N;
#ifdef NNS_ABL_NOTRANS       // timing ablation: no block stores / loads at the member changes
                if (false) {
#else
                if (owner) {
#endif
#pragma unroll
                    for (int li = 0; li < BR; ++li)
#pragma unroll
                        for (int lj = 0; lj < BC; ++lj) ps[(size_t)(r0 + li) * NY + c0 + lj] = P[li][lj];
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) s_wdone[w] = kw + 1;
#ifndef NNS_ABL_NOTRANS
                if (kw + 1 < nmine) fetch(m + gridDim.x);
#endif
            }
            named_sync(BAR_SOR, NT_SOR);
            // ---- local stage last + 1: the warp reports
            {
                const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mask), hi = __reduce_or_sync(0xffffffffu, (unsigned)(mask >> 32));
                const unsigned alo = __reduce_or_sync(0xffffffffu, (unsigned)amb), ahi = __reduce_or_sync(0xffffffffu, (unsigned)(amb >> 32));
                mask = 0ull; amb = 0ull;
                // the p blocks of every lane before the warp reports: the readers are threads of this CTA (generic
                // proxy) and bulk copies issued by them (async proxy)
                __threadfence_block();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    atomicOr(&s_mask[kw & 1][0], ((unsigned long long)hi << 32) | lo);
                    atomicOr(&s_mask[kw & 1][1], ((unsigned long long)ahi << 32) | alo);
                    __threadfence_block();
                    s_wstored[w] = kw + 1;
                }
            }
            named_sync(BAR_SOR, NT_SOR);
            // ---- idle stages up to the end of the member's period
            for (int tl = last + 2; tl < per; ++tl) named_sync(BAR_SOR, NT_SOR);
            passed += per;
        }
        for (; passed <= tend; ++passed) named_sync(BAR_SOR, NT_SOR);
        if (a.prof && lead) {
            a.prof[(size_t)blockIdx.x * NPROF + 3] += clock64() - tk0;
            a.prof[(size_t)blockIdx.x * NPROF + 12] += tsw;
        }
        if (a.prof && lane == 0) a.prof[(size_t)blockIdx.x * NPROF + 24 + w] += tsw;
    }
    // every Tensor Memory access of the CTA is complete
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

#endif  // NNS_ENABLE_WAVE

using Cfg128 = CfgX<9, 7, 14, 18>;      // 128 x 128: 126 = 14*9 = 18*7, 252 blocks of 63 cells

struct StreamPlan {
    std::vector<SBlock> desc;
    std::vector<short> tidmap;
    void *d_tab = nullptr;      // desc then tidmap
    double *d_img = nullptr;
    long long *d_prof = nullptr;
    long long *d_trace = nullptr;
    double *d_pscr = nullptr;   // wave kernel: [grid][2][NX*NY] SOR results
    int *d_redo = nullptr;      // wave kernel: [1 + batch] counter, member list of the re-run pass
    unsigned depmask[NW_SOR] = {0};
    int wamin[NW_SOR] = {0}, wrange[NW_SOR] = {0}, rmax = 0;
    int grid = 0;
    bool ok = false;
    bool wave = false;
};

template <typename C>
static void build_tables(StreamPlan &pl) {
    constexpr int BR = C::BRc, BC = C::BCc, NBR = C::NBRc, NBC = C::NBCc, NB = NBR * NBC;
    struct Item { int key, bi, bj; };
    std::vector<Item> items;
    for (int bi = 0; bi < NBR; ++bi)
        for (int bj = 0; bj < NBC; ++bj) items.push_back({2 * bi + bj, bi, bj});     // key = top sub-block diagonal
    // threads whose top sub-blocks work in the same super-stages (same parity of the key, i.e. of bj) share
    // warps, ordered by diagonal so that a warp's lanes enter and leave the active band together
    std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) {
        if ((x.key & 1) != (y.key & 1)) return (x.key & 1) < (y.key & 1);
        return x.key != y.key ? x.key < y.key : x.bj < y.bj;
    });
    // even-parity blocks occupy threads [0, n_even), odd-parity blocks start at thread NT_SOR/2: a warp never
    // mixes parities, and each SM sub-partition hosts one warp of either parity (top and bottom sub-block
    // sweeps have different lengths; together they balance)
    int n_even = 0;
    for (const Item &it : items) n_even += !(it.key & 1);
    std::vector<int> tid_of((size_t)NB), thr((size_t)NB);
    for (int t = 0; t < NB; ++t) {
#ifndef NNS_SOR_SAME_PAIR
        // the warps of the odd class in reverse order: an SM sub-partition (warp % 4) then hosts an early-diagonal warp of
        // one class and a late-diagonal warp of the other, so that during the fill / drain stages of the wavefront the
        // active warps sit on different sub-partitions instead of sharing one
        thr[t] = t < n_even ? t : NT_SOR / 2 + 32 * (NW_SOR / 2 - 1 - (t - n_even) / 32) + (t - n_even) % 32;
#else
        thr[t] = t < n_even ? t : NT_SOR / 2 + (t - n_even);
#endif
        tid_of[(size_t)items[t].bi * NBC + items[t].bj] = thr[t];
    }
    pl.desc.assign(NT_SOR, SBlock{0, 0, -1, -1, -1, -1, 0, 0});     // r0 == 0 marks a thread without a block
    pl.tidmap.resize(NB);
    for (int t = 0; t < NB; ++t) pl.tidmap[t] = (short)tid_of[t];
    pl.ok = n_even <= NT_SOR / 2 && NB - n_even <= NT_SOR / 2;
    for (int it = 0; it < NB; ++it) {
        const int t = thr[it];
        const int bi = items[it].bi, bj = items[it].bj;
        SBlock d;
        d.r0 = (short)(1 + BR * bi);
        d.c0 = (short)(1 + BC * bj);
        d.nN = bi > 0 ? (short)tid_of[(size_t)(bi - 1) * NBC + bj] : (short)-1;
        d.nS = bi < NBR - 1 ? (short)tid_of[(size_t)(bi + 1) * NBC + bj] : (short)-1;
        d.nW = bj > 0 ? (short)tid_of[(size_t)bi * NBC + bj - 1] : (short)-1;
        d.nE = bj < NBC - 1 ? (short)tid_of[(size_t)bi * NBC + bj + 1] : (short)-1;
        d.bd = (short)(2 * bi + bj);
        d.pad = 0;
        pl.desc[t] = d;
    }
    // stage hand-off: warp w waits for the warps that hold a neighbour of one of its blocks
    for (int w = 0; w < NW_SOR; ++w) pl.depmask[w] = 0u;
    for (int t = 0; t < NT_SOR; ++t) {
        const SBlock &d = pl.desc[t];
        if (d.r0 <= 0) continue;
        for (short nb : {d.nN, d.nS, d.nW, d.nE})
            if (nb >= 0 && (nb >> 5) != (t >> 5)) pl.depmask[t >> 5] |= 1u << (nb >> 5);
    }
    // wave kernel: first block diagonal and spread of every SOR warp (members change per warp)
    pl.rmax = 0;
    for (int w = 0; w < NW_SOR; ++w) {
        int lo = 1 << 20, hi = -1;
        for (int l = 0; l < 32; ++l) {
            const SBlock &d = pl.desc[(size_t)w * 32 + l];
            if (d.r0 > 0) { lo = std::min(lo, (int)d.bd); hi = std::max(hi, (int)d.bd); }
        }
        pl.wamin[w] = hi < 0 ? 0 : lo;
        pl.wrange[w] = hi < 0 ? 0 : hi - lo;
        pl.rmax = std::max(pl.rmax, pl.wrange[w]);
    }
}

}  // namespace

// Debug aid: print and reset the phase counters (NNS_STREAM_PROF=1), averaged over CTAs.
void chorin_stream_prof_dump(nns_handle *h) {
    StreamPlan *pl = static_cast<StreamPlan *>(h->stream_plan);
#ifdef NNS_STREAM_TRACE
    if (pl && pl->d_trace && !pl->wave) {
        std::vector<long long> t((size_t)NW_SOR * TRACE_STAGES * 4);
        cudaDeviceSynchronize();
        cudaMemcpy(t.data(), pl->d_trace, sizeof(long long) * t.size(), cudaMemcpyDeviceToHost);
        // per stage and warp: start (after the previous release) relative to warp 0, cycles start -> arrival, arrival -> release, active lanes
        for (int st = 0; st < TRACE_STAGES; ++st) {
            if (!t[(size_t)st * 4]) continue;
            fprintf(stderr, "[nns stream trace] T %3d len %5lld:", st, st + 1 < TRACE_STAGES && t[(size_t)(st + 1) * 4] ? t[(size_t)(st + 1) * 4] - t[(size_t)st * 4] : 0ll);
            for (int w = 0; w < NW_SOR; ++w) {
                const long long *e = &t[((size_t)w * TRACE_STAGES + st) * 4];
                fprintf(stderr, " | %2lld %4lld+%4lld", e[2], e[1] - e[0], e[3] - e[1]);
            }
            fprintf(stderr, "\n");
        }
        cudaMemset(pl->d_trace, 0, sizeof(long long) * t.size());
    }
#endif
    if (pl && pl->d_trace && pl->wave) {
        std::vector<long long> t((size_t)NW_SOR * 16 * 4);
        cudaDeviceSynchronize();
        cudaMemcpy(t.data(), pl->d_trace, sizeof(long long) * t.size(), cudaMemcpyDeviceToHost);
        const long long base = t[0];
        for (int st = 0; st < 16; ++st) {
            fprintf(stderr, "[nns wave trace] stage %2d:", st);
            for (int w = 0; w < NW_SOR; ++w) {
                const long long *e = &t[((size_t)w * 16 + st) * 4];
                fprintf(stderr, " w%d %lld/%lld/%lld", w, e[0] - base, e[1] - base, e[2] - base);
            }
            fprintf(stderr, "\n");
        }
    }
    if (!pl || !pl->d_prof) return;
    std::vector<long long> v((size_t)NPROF * pl->grid);
    cudaDeviceSynchronize();
    cudaMemcpy(v.data(), pl->d_prof, sizeof(long long) * v.size(), cudaMemcpyDeviceToHost);
    cudaMemset(pl->d_prof, 0, sizeof(long long) * v.size());
    static const char *nm[NPROF] = {"sor.load_p", "sor.wait_ready", "sor.pull_cimg", "sor.wavefront", "sor.reduce_redo",
                                    "sor.store_p", "p1.producer_issue(w0)", "p1.row_loop(w1)", "st.pass1", "st.wait_consumed/fill", "st.wait_done", "st.pass2",
                                    "sor.sweep_cyc(t0)", "sor.sweeps(t0)", "sor.sweep_cyc(t128)", "sor.sweeps(t128)",
                                    "p1.prologue", "p1.wait_full", "p1.row_loop", "p1.bcs", "p1.patch", "p2.edges_bcs",
                                    "p2.wait_full", "p2.row_loop", "sor.sweep_cyc(w0)", "sor.sweep_cyc(w1)", "sor.sweep_cyc(w2)",
                                    "sor.sweep_cyc(w3)", "sor.sweep_cyc(w4)", "sor.sweep_cyc(w5)", "sor.sweep_cyc(w6)", "sor.sweep_cyc(w7)"};
    for (int k = 0; k < NPROF; ++k) {
        if (!nm[k][0]) continue;
        double s = 0;
        for (int b = 0; b < pl->grid; ++b) s += (double)v[(size_t)b * NPROF + k];
        fprintf(stderr, "[nns stream prof] %-22s %12.0f cycles/CTA\n", nm[k], s / pl->grid);
    }
}

bool chorin_stream_eligible(const nns_handle *h, int phases, int nsteps) {
    const char *e = getenv("NNS_CHIP_MODE");
    if (e && strcmp(e, "stream") != 0 && e[0]) return false;     // tests force the other paths
    return h->g.method == NNS_METHOD_EXPLICIT && phases == 7 && nsteps >= 1 && h->g.nx == Cfg128::NX &&
           h->g.ny == Cfg128::NY && h->g.nit >= 2 && h->g.nit <= 65 &&
           (size_t)h->max_smem_optin >= Cfg128::SMEM_BYTES + 256;
}

void chorin_stream_free(nns_handle *h) {
    StreamPlan *pl = static_cast<StreamPlan *>(h->stream_plan);
    if (!pl) return;
    cudaFree(pl->d_tab);
    cudaFree(pl->d_img);
    cudaFree(pl->d_prof);
    cudaFree(pl->d_pscr);
    cudaFree(pl->d_redo);
    delete pl;
    h->stream_plan = nullptr;
}

// One step of members [m0, m0+count); field pointers already offset to member m0.
int chorin_stream_step(nns_handle *h, const double *uc, const double *vc, const double *up, const double *vp,
                       double *un, double *vn, double *p, double *tu, double *tv, double *tp,
                       size_t traj_member_stride, size_t traj_off, int32_t *sweeps, cudaStream_t st, int m0,
                       int count) {
    using C = Cfg128;
    if (count < 0) count = h->g.batch - m0;
    StreamPlan *pl = static_cast<StreamPlan *>(h->stream_plan);
    if (!pl) {
        pl = new StreamPlan();
        h->stream_plan = pl;
        build_tables<C>(*pl);
        const size_t dbytes = sizeof(SBlock) * pl->desc.size(), tbytes = sizeof(short) * pl->tidmap.size();
        NNS_CUDA(cudaMalloc(&pl->d_tab, dbytes + tbytes));
        NNS_CUDA(cudaMemcpy(pl->d_tab, pl->desc.data(), dbytes, cudaMemcpyHostToDevice));
        NNS_CUDA(cudaMemcpy(static_cast<char *>(pl->d_tab) + dbytes, pl->tidmap.data(), tbytes, cudaMemcpyHostToDevice));
        short ord[128] = {0};
        for (int li = 0; li < C::BRc; ++li)
            for (int lj = 0; lj < C::BCc; ++lj) ord[li * C::BCc + lj] = (short)split_ord<C::BRc, C::BCc, C::RSc>(li, lj);
        NNS_CUDA(cudaMemcpyToSymbol(c_ord, ord, sizeof(ord)));
        if (getenv("NNS_STREAM_PROF")) {
            NNS_CUDA(cudaMalloc(&pl->d_prof, sizeof(long long) * NPROF * h->sm_count));
            NNS_CUDA(cudaMemset(pl->d_prof, 0, sizeof(long long) * NPROF * h->sm_count));
        }
        pl->grid = h->sm_count;
        NNS_CUDA(cudaMalloc(&pl->d_img, sizeof(double2) * C::NCH * NT_SOR * (size_t)pl->grid * N_SCRATCH));
        NNS_CUDA(cudaMemset(pl->d_img, 0, sizeof(double2) * C::NCH * NT_SOR * (size_t)pl->grid * N_SCRATCH));
        NNS_CUDA(cudaFuncSetAttribute(chorin_stream_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::SMEM_BYTES));
#ifdef NNS_STREAM_TRACE
        if (getenv("NNS_STREAM_TRACE")) {
            NNS_CUDA(cudaMalloc(&pl->d_trace, sizeof(long long) * NW_SOR * TRACE_STAGES * 4));
            NNS_CUDA(cudaMemset(pl->d_trace, 0, sizeof(long long) * NW_SOR * TRACE_STAGES * 4));
        }
#endif
#ifdef NNS_ENABLE_WAVE
        const char *mode = getenv("NNS_STREAM_MODE");
        pl->wave = mode && strcmp(mode, "wave") == 0 && h->g.nit - 1 <= 64;
        if (pl->wave) fprintf(stderr, "[nns_b200] WARNING: NNS_STREAM_MODE=wave selects an experimental kernel whose results are not reproducible bit for bit\n");
        if (pl->wave && getenv("NNS_WAVE_TRACE")) {
            NNS_CUDA(cudaMalloc(&pl->d_trace, sizeof(long long) * NW_SOR * 16 * 4));
            NNS_CUDA(cudaMemset(pl->d_trace, 0, sizeof(long long) * NW_SOR * 16 * 4));
        }
        if (pl->wave) {
            NNS_CUDA(cudaFuncSetAttribute(chorin_stream_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)C::SMEM_BYTES));
            NNS_CUDA(cudaMalloc(&pl->d_pscr, sizeof(double) * 2 * C::NX * C::NY * (size_t)pl->grid));
            NNS_CUDA(cudaMemset(pl->d_pscr, 0, sizeof(double) * 2 * C::NX * C::NY * (size_t)pl->grid));
            NNS_CUDA(cudaMalloc(&pl->d_redo, sizeof(int) * (1 + (size_t)h->g.batch)));
            NNS_CUDA(cudaFuncSetAttribute(chorin_wave_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)WaveSmem<C>::SMEM_BYTES));
        }
#endif
    }
    StreamArgs a{};
    a.g = h->g;
    a.ubc = h->bc[0]; a.vbc = h->bc[1]; a.pbc = h->bc[2];
    a.nu_b = h->d_nu ? h->d_nu + m0 : nullptr;
    a.bcval = h->d_bcval ? h->d_bcval + (size_t)m0 * h->n_bcs : nullptr;
    a.n_bcs = h->n_bcs;
    a.count = count;
    a.flags = h->params.flags;
    a.desc = static_cast<const SBlock *>(pl->d_tab);
    a.tidmap = reinterpret_cast<const short *>(static_cast<const char *>(pl->d_tab) + sizeof(SBlock) * pl->desc.size());
    a.uc = uc; a.vc = vc; a.up = up; a.vp = vp; a.un = un; a.vn = vn; a.p = p;
    a.cimg = pl->d_img + (size_t)(h->scratch_slot % N_SCRATCH) * 2 * C::NCH * NT_SOR * (size_t)pl->grid;     // a set per internal stream: concurrent launches never share a C' image
    a.traj_u = tu; a.traj_v = tv; a.traj_p = tp;
    a.traj_member_stride = traj_member_stride; a.traj_off = traj_off;
    a.sweeps = sweeps;
    a.nonfinite = h->d_nonfinite;
    a.prof = pl->d_prof;
    a.trace = pl->d_trace;
    for (int w = 0; w < NW_SOR; ++w) a.depmask[w] = pl->depmask[w];
    const int grid = count < pl->grid ? count : pl->grid;
#ifdef NNS_ENABLE_WAVE
    if (pl->wave) {
        a.redo_count = pl->d_redo;
        a.redo_list = pl->d_redo + 1;
        a.pscr = pl->d_pscr;
        for (int w = 0; w < NW_SOR; ++w) { a.wamin[w] = pl->wamin[w]; a.wrange[w] = pl->wrange[w]; }
        a.rmax = pl->rmax;
        NNS_CUDA(cudaMemsetAsync(pl->d_redo, 0, sizeof(int), st));
        chorin_wave_kernel<C><<<grid, NT_SOR + NT_ST, WaveSmem<C>::SMEM_BYTES, st>>>(a);
        NNS_CUDA(cudaGetLastError());
        // re-run pass: members whose SOR loop stops before nit - 1 sweeps (none in the common case: the CTAs
        // read an empty list and leave)
        a.list = a.redo_list;
        a.list_count = a.redo_count;
        chorin_stream_kernel<C, true><<<grid, NT_SOR + NT_ST, C::SMEM_BYTES, st>>>(a);
        NNS_CUDA(cudaGetLastError());
        h->launches += 2;
        return NNS_OK;
    }
#endif
    chorin_stream_kernel<C, false><<<grid, NT_SOR + NT_ST, C::SMEM_BYTES, st>>>(a);
    NNS_CUDA(cudaGetLastError());
    h->launches += 1;
    return NNS_OK;
}

}  // namespace nns
