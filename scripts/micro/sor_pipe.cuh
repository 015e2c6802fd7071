// sor_pipe.cuh -- EXPERIMENT (not part of libnns_b200; see DESIGN.md 4.1c and profiles/r2_pipe_*): a row-streamed SOR role: exact lexicographic Gauss-Seidel/SOR sweeps of
// src/chorin_fd/simulate.py:190-200 executed as a PIPELINE OVER SWEEPS through which the rows of the
// ensemble members stream.
//
//   * The grid rows of the CTA's members form one long stream (member k row i = stream row k*NX + i).
//   * One warp owns a whole 128-column row: lane l owns columns 4l .. 4l+3 and runs l steps behind lane
//     l-1, so that the west operand (new value of column 4l-1) arrives by one warp shuffle per step and
//     the east operand (old value of column 4l+4) by another: inside a row the update order is exactly
//     the lexicographic one.
//   * Sweep level s+1 follows sweep level s two rows behind (it needs the level-s values of the row below).
//     A warp holds K consecutive levels in registers (two rows of four doubles per level); the K levels
//     of a step are independent instruction streams (K-fold ILP for the 8-cycle DFMA latency), and the
//     eight SOR warps of the CTA form a chain over all nit-1 levels: warp w hands the rows leaving its
//     last level to warp w+1 through a small shared-memory ring (lane-private 16-byte chunks, release /
//     acquire step counters, no CTA barrier anywhere).
//   * The right-hand side C' of a row travels with the row through the rings and waits in the warp's
//     Tensor Memory lanes (tcgen05.st / tcgen05.ld, 32x32b shape: 8 columns per row) until each of the
//     warp's levels has used it.
//   * Frozen cells (boundary rows 0 and NX-1 of every member, boundary columns 0 and 127) pass through
//     unchanged: boundary rows by zeroed coefficients (their C' is zero), boundary columns by a select.
//     Members are therefore isolated from each other and from the dummy rows in front of and behind the
//     stream, and the stream never stops between members: no fill / drain per member.
//   * Exit test of the reference loop (max|dp| <= tol per sweep): running maxima of the high words of
//     the per-cell increments on the integer pipe, per level and member; a member whose flags say that
//     the reference loop stops early (or that the fast test is undecided) is queued for the exact re-run
//     pass (chorin_stream_kernel<.., LIST>).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nns {
namespace pipe {

constexpr int NW = 7;                 // SOR warps = pipeline stages
constexpr int KW = 7;                 // levels per warp: NW * KW = 49 >= nit - 1 (levels beyond nit - 1 pass rows through)
constexpr int RSLOT = 4;              // slots of an inter-warp ring
constexpr int ROWB = 1024;            // one row of 128 doubles
constexpr int LG = 2;                 // levels per group (one Tensor Memory load, interleaved dependent chains)
constexpr int SLOTB = 2 * ROWB;       // p row + C'' row (lane-private 16-byte chunks: [p lo][p hi][C'' lo][C'' hi] x 32 lanes)
constexpr int R0S = 48;               // slots of the entry ring (same skewed format: the lane skew alone keeps 32 slots alive)
constexpr int TM_SLOTS = 32;          // C'' rows per warp in Tensor Memory (8 columns each): 256 columns per warp
constexpr unsigned SPIN_LIMIT = 1u << 15;     // x 20 us of hardware sleep per try

struct Coef {
    // p' = p + om * r,  r = a (N + S) + b (E + W) - p - C''   with a = dy^2/den, b = dx^2/den, den = 2 dx^2 + 2 dy^2,
    // om = beta and C'' = C / den (chorin_fd/simulate.py:186-196 with the constants folded)
    double a, b, om;
    unsigned thi;                     // high word of tol / om: the fast exit test compares the high words of |r|
};

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier full / empty protocol of the rings: mbarrier.arrive (release, CTA scope) is a bare SYNCS.ARRIVE behind the
// shared-memory stores of the same thread (no MEMBAR, unlike st.release), try_wait (acquire) needs no fence either.
__device__ __forceinline__ void mbar_init(uint32_t b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t b) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred P1;\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P1, [%1], %2, %3;\n"     // %3: hardware sleep until the phase completes, at most this many ns (a waiting warp must not steal issue slots)
        "selp.u32 %0, 1, 0, P1;\n}"
        : "=r"(ok) : "r"(b), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
// a stuck pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity) {
#ifdef PIPE_ABL_NOSYNC
    return;
#endif
    if (mbar_try(b, parity)) return;
    unsigned spins = 0;
    while (!mbar_try(b, parity))
        if (++spins > SPIN_LIMIT) __trap();
}
__device__ __forceinline__ double2 lds2(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
// tcgen05.ld, 32x32b shape: thread i of the warp reads NC consecutive 32-bit columns of lane 32*(warp%4)+i
template <int NC> __device__ __forceinline__ void tm_ld(uint32_t taddr, uint32_t *r);
template <> __device__ __forceinline__ void tm_ld<8>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld<16>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void tm_ld<32>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// wait for the outstanding tcgen05.ld of the thread; the registers are tied to the statement so that no use moves above it
template <int NC> __device__ __forceinline__ void tm_wait_ld(uint32_t *r);
template <> __device__ __forceinline__ void tm_wait_ld<8>(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
template <> __device__ __forceinline__ void tm_wait_ld<16>(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
template <> __device__ __forceinline__ void tm_wait_ld<32>(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const double (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                 "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory control block of the pipeline: full / empty mbarrier per ring slot.
struct Ctl {
    unsigned long long full0[R0S], empty0[R0S];              // entry ring (feeder -> SOR warp 0)
    unsigned long long full[NW][RSLOT], empty[NW][RSLOT];    // ring w (SOR warp w-1 -> w), w = 1 .. NW-1: 32 arrivals each
    int nrep[4];           // warps that have reported member k (slot k & 3)
    unsigned long long viol[4], amb[4];
};

template <int K>
struct WarpState {
    double B[K + 1][2][4];    // B[s]: the two newest rows entering level s (B[K]: leaving the warp), slot = step parity
    float amax[K > 0 ? K : 1];     // running maxima of the high words of |r| per level (as float bit patterns)
    double cprev[4];          // C'' of the row that entered in the previous step (stored to Tensor Memory at the next step's top)
    uint32_t cq[8 * LG];      // C'' rows in flight from Tensor Memory (group 1 of the next step between steps)
    unsigned flags;           // exit-test results of the member that just ended: bit s violated, bit 8+s undecided
    unsigned F;               // bit k: the row that entered the warp k steps ago is a boundary row (0 or NX-1) of its member
};

// Per-warp constants and counters of the SOR role.
struct WarpCtx {
    Coef k;
    double a0, b0, one0, a3, b3, one3;   // coefficients of the lane's first / last cell: zero in a boundary column (lane 0 / 31)
    int nx;                   // rows per member
    int lane;
    uint32_t tm;              // Tensor Memory address of the warp's C'' window (lane quarter included; 256 columns)
    uint32_t in_ring, out_ring;          // shared-memory addresses of the input / output ring (+ lane offset)
    uint32_t full_in, empty_in, full_out, empty_out;    // mbarrier arrays of the two rings
    int in_slots;             // slots of the input ring (R0S for the first warp, RSLOT otherwise)
    int in_slot, out_slot;    // slot of this step
    uint32_t in_par, out_par; // phase parities: full barrier of the input slot, empty barrier of the output slot
    int iin;                  // row-in-member of the row entering the warp in this step (lane-private)
};

// Loop invariants of the role are made opaque to the compiler: otherwise it re-derives them from S2R / constant-bank reads
// in every step (dozens of instructions per step) instead of keeping them in registers.
__device__ __forceinline__ uint32_t opaque(uint32_t x) { uint32_t y; asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ int opaque(int x) { int y; asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ double opaque(double x) { double y; asm volatile("mov.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ void pin_ctx(WarpCtx &c) {
    c.a0 = opaque(c.a0); c.b0 = opaque(c.b0); c.one0 = opaque(c.one0); c.a3 = opaque(c.a3); c.b3 = opaque(c.b3); c.one3 = opaque(c.one3);
    c.k.a = opaque(c.k.a); c.k.b = opaque(c.k.b); c.k.om = opaque(c.k.om);
    c.tm = opaque(c.tm); c.in_ring = opaque(c.in_ring); c.out_ring = opaque(c.out_ring);
    c.full_in = opaque(c.full_in); c.empty_in = opaque(c.empty_in); c.full_out = opaque(c.full_out); c.empty_out = opaque(c.empty_out);
    c.in_slots = opaque(c.in_slots); c.nx = opaque(c.nx);
}

__device__ __forceinline__ void init_lane_coef(WarpCtx &c) {
    c.a0 = c.lane == 0 ? 0.0 : c.k.a;  c.b0 = c.lane == 0 ? 0.0 : c.k.b;  c.one0 = c.lane == 0 ? 0.0 : 1.0;
    c.a3 = c.lane == 31 ? 0.0 : c.k.a; c.b3 = c.lane == 31 ? 0.0 : c.k.b; c.one3 = c.lane == 31 ? 0.0 : 1.0;
}

// Tensor Memory layout of a warp's C'' window (256 columns): the rows that entered in steps of parity h live in columns
// [128 h, 128 h + 128) as 16 slots of 8 columns; the row of step n goes to slots j and j + 8 with j = (n >> 1) & 7, so
// that the rows of up to eight consecutive same-parity steps are always one contiguous run of slots (no wrap) and one
// wide tcgen05.ld fetches the C'' rows of several levels.  Level s computes, in step n, the row that entered in step
// n - 2 - 2s: slot ((n >> 1) - 1 - s) & 7 of the same parity.  (K <= 7: eight live rows per parity.)
template <int K, int PAR>
__device__ __forceinline__ uint32_t tm_slot_addr(const WarpCtx &c, int n, int s_hi) {
    return c.tm + 128u * PAR + 8u * (uint32_t)(((n >> 1) - 1 - s_hi) & 7);
}

// Levels S_HI, S_HI-1, .., S_HI-L+1 of one step, phase A: u = -p - C'' (consumes the C'' rows cq[0..8), cq[8..16), ...;
// the registers of cq are free for the next Tensor Memory load afterwards)
template <int K, int PAR, int S_HI, int L>
__device__ __forceinline__ void levels_a(const WarpState<K> &st, const WarpCtx &c, const uint32_t *cq, double (&u)[LG][4], double (&cnew)[4]) {
#pragma unroll
    for (int g = 0; g < L; ++g) {
        const int s = S_HI - g;
        double cp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cp[q] = __hiloint2double((int)cq[8 * g + 2 * q + 1], (int)cq[8 * g + 2 * q]);
#ifdef PIPE_ABL_NOTMEM
#pragma unroll
        for (int q = 0; q < 4; ++q) cp[q] = 1e-3 * (s + 1) + q;
#endif
        if (s == K - 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cnew[q] = cp[q];
        }
        const double(&Cn)[4] = st.B[s][PAR];
        u[g][0] = fma(-c.one0, Cn[0], -cp[0]);
        u[g][1] = -Cn[1] - cp[1];
        u[g][2] = -Cn[2] - cp[2];
        u[g][3] = fma(-c.one3, Cn[3], -cp[3]);
    }
}

// phase B: the rest of the update; the dependent chains (west operand) of the L levels are interleaved
template <int K, int PAR, int S_HI, int L>
__device__ __forceinline__ void levels_b(WarpState<K> &st, const WarpCtx &c, const double (&u)[LG][4]) {
    double t[L][4], w[L], om[L];
#pragma unroll
    for (int g = 0; g < L; ++g) {
        const int s = S_HI - g;
        const bool frozen = (st.F >> (2 + 2 * s)) & 1u;
        om[g] = frozen ? 0.0 : c.k.om;
        const double(&Cn)[4] = st.B[s][PAR];
        const double(&S)[4] = st.B[s][PAR ^ 1];
        const double(&N)[4] = st.B[s + 1][PAR ^ 1];
#ifndef PIPE_ABL_NOSHFL
        // lane l+1 runs one row behind: the row this lane computes is its SOUTH row
        const double e3 = __shfl_down_sync(0xffffffffu, S[0], 1);      // old value of column 4l+4 (lane 31: unused)
        w[g] = __shfl_up_sync(0xffffffffu, N[3], 1);                   // new value of column 4l-1 (lane 0: unused)
#else
        const double e3 = S[0];
        w[g] = N[3];
#endif
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double e = q < 3 ? Cn[q + 1] : e3;
            const double aq = q == 0 ? c.a0 : q == 3 ? c.a3 : c.k.a, bq = q == 0 ? c.b0 : q == 3 ? c.b3 : c.k.b;
            t[g][q] = fma(aq, N[q] + S[q], fma(bq, e, u[g][q]));
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int g = 0; g < L; ++g) {
            const int s = S_HI - g;
            const double bq = q == 0 ? c.b0 : q == 3 ? c.b3 : c.k.b;
            const double r = fma(bq, w[g], t[g][q]);
            w[g] = fma(om[g], r, st.B[s][PAR][q]);
            st.B[s + 1][PAR][q] = w[g];
#ifndef PIPE_ABL_NOEXIT
            // running maximum of the high words of |r| as FLOAT bit patterns (same order for positive patterns; FMNMX with
            // the |x| operand modifier is one ALU-pipe instruction, the integer VIMNMX is a slow-pipe instruction on B200).
            // Patterns that are float NaNs (|r| >= 2^1017) or denormals are ignored / flushed: the maximum can only come
            // out too small, i.e. claim an early exit that the exact re-run pass then refutes.
            st.amax[s] = fmaxf(st.amax[s], fabsf(__int_as_float(__double2hiint(r))));
#endif
        }
}

// The K levels of a step in groups of LG (K-1, K-2 | K-3, K-4 | ...): one Tensor Memory load per group, software-
// pipelined -- on entry st.cq holds the C'' rows of group 0 (issued at the bottom of the previous step), the load of
// group g+1 is issued as soon as group g has consumed its rows (phase A), and the load of the next step's group 0 after
// the last group's phase A.
template <int K, int PAR, int G>
__device__ __forceinline__ void level_groups(WarpState<K> &st, const WarpCtx &c, int n, double (&cnew)[4]) {
    constexpr int S_HI = K - 1 - LG * G, L = S_HI + 1 >= LG ? LG : S_HI + 1, NC = 8 * L;
    constexpr int S_NEXT = S_HI - L, LN = S_NEXT + 1 >= LG ? LG : S_NEXT + 1;       // next group of this step (S_NEXT < 0: none)
    constexpr int L0 = K >= LG ? LG : K;                                             // group 0 of the next step
    double u[LG][4];
#ifndef PIPE_ABL_NOTMEM
    tm_wait_ld<NC>(st.cq);
#endif
    levels_a<K, PAR, S_HI, L>(st, c, st.cq, u, cnew);
#ifndef PIPE_ABL_NOTMEM
    if (S_NEXT >= 0) {
        tm_ld<8 * (LN > 0 ? LN : 1)>(tm_slot_addr<K, PAR>(c, n, S_NEXT >= 0 ? S_NEXT : 0), st.cq);
    } else {
        tm_wait_st();
        tm_ld<8 * L0>(tm_slot_addr<K, PAR ^ 1>(c, n + 1, K - 1), st.cq);
    }
#endif
    levels_b<K, PAR, S_HI, L>(st, c, u);
    if constexpr (S_NEXT >= 0) level_groups<K, PAR, G + 1>(st, c, n, cnew);
}

// One step of a warp with K levels.  PAR = n & 1 (compile time through 2x unrolling).
//   level s:  centre (old) = B[s][PAR], south (old) = B[s][PAR^1], north (new) = B[s+1][PAR^1], result -> B[s+1][PAR]
//   (descending s: level s+1 has read B[s+1][PAR] before level s overwrites it), then the row read from the input
//   ring in this step -> B[0][PAR].  Level s therefore computes, in step n, the row that entered the warp in step
//   n - 2 - 2s, and the row leaving the warp in step n entered it in step n - 2K.
//   Frozen rows: om = 0 (p' = p + 0 * r; a non-finite r -- a diverged member -- would leak into the frozen row and from
//   there into the next member of the stream: members whose maxima are non-finite are reported, and they and their
//   successors go through the isolated re-run pass).  Frozen columns: zero coefficients and zero C'' give r = 0.
//   The levels run as two groups (K-1 .. K-L1 and L2-1 .. 0) with one wide Tensor Memory load each; the loads are
//   software-pipelined: on entry st.cq holds the C'' rows of group 1 (issued at the bottom of the previous step), the
//   group-2 load is issued as soon as group 1 has consumed its rows, and the next step's group-1 load as soon as group 2 has.
template <int K, int PAR, typename ExitFn>
__device__ __forceinline__ void step(WarpState<K> &st, WarpCtx &c, const bool LAST, int n, ExitFn &&exit_row) {
    static_assert(K >= 1 && K <= 7, "eight live C'' rows per parity in Tensor Memory");
    // ---- top: the C'' row that entered in the previous step into Tensor Memory (slots j and j + 8, parity PAR ^ 1)
    {
        const uint32_t ta = c.tm + 128u * (PAR ^ 1) + 8u * (uint32_t)(((n - 1) >> 1) & 7);
        tm_st8(ta, st.cprev);
        tm_st8(ta + 64u, st.cprev);
    }
    // early probes of the two ring barriers of this step (the answers are needed at the bottom)
    const uint32_t bar_in = c.full_in + 8u * (uint32_t)c.in_slot, bar_out = c.empty_out + 8u * (uint32_t)c.out_slot;
    const bool in_ok = mbar_try(bar_in, c.in_par);
    const bool out_ok = (LAST || n < RSLOT) ? true : mbar_try(bar_out, c.out_par);
    st.F = (st.F << 1) | (unsigned)(c.iin == 0 || c.iin == c.nx - 1);
    if (++c.iin == c.nx) c.iin = 0;
    double cnew[4] = {0.0, 0.0, 0.0, 0.0};     // C'' of the row leaving the warp
#ifndef PIPE_ABL_NOEXIT
    // rare path: a member boundary passes some level.  Row NX-1 (frozen, previous row not): the maxima of that level
    // belong to a complete member -> flags.  Row 1 (not frozen, previous row frozen): the maxima hold the garbage
    // residuals of the two frozen rows -> reset.
    constexpr unsigned EV = 0xfffffffcu & ((1u << (2 * K + 2)) - 1u);
    if (st.F & EV) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            const unsigned now = (st.F >> (2 + 2 * s)) & 1u, prev = (st.F >> (3 + 2 * s)) & 1u;
            if (now && !prev) {
                const unsigned mhi = (unsigned)__float_as_int(st.amax[s]);
                // |om r| <= tol decided on the high words of r and tol/om: certain unless they are within one
                // high-word step of each other (rounding of the product), then the exact re-run decides
                const bool und = mhi + 1u >= c.k.thi && mhi <= c.k.thi + 1u;
                st.flags |= (unsigned)(mhi > c.k.thi && !und) << s | (unsigned)und << (8 + s);
            }
            if (prev && !now) st.amax[s] = 0.0f;
        }
    }
#endif
    level_groups<K, PAR, 0>(st, c, n, cnew);
    // ---- bottom: the input row of this step enters the warp
    if (!in_ok) mbar_wait(bar_in, c.in_par);
    const uint32_t ai = c.in_ring + (uint32_t)c.in_slot * SLOTB;
    const double2 pin0 = lds2(ai), pin1 = lds2(ai + 512), cin0 = lds2(ai + 1024), cin1 = lds2(ai + 1536);
    mbar_arrive(c.empty_in + 8u * (uint32_t)c.in_slot);
    if (++c.in_slot == c.in_slots) { c.in_slot = 0; c.in_par ^= 1u; }
    // ---- the row leaving the warp
    if (LAST) {
        exit_row(st.B[K][PAR], n);
    } else {
        if (!out_ok) mbar_wait(bar_out, c.out_par);
        const uint32_t a = c.out_ring + (uint32_t)c.out_slot * SLOTB;
        sts2(a, st.B[K][PAR][0], st.B[K][PAR][1]);
        sts2(a + 512, st.B[K][PAR][2], st.B[K][PAR][3]);
        sts2(a + 1024, cnew[0], cnew[1]);
        sts2(a + 1536, cnew[2], cnew[3]);
        mbar_arrive(c.full_out + 8u * (uint32_t)c.out_slot);
        if (++c.out_slot == RSLOT) { c.out_slot = 0; if (n >= RSLOT) c.out_par ^= 1u; }
    }
    st.B[0][PAR][0] = pin0.x; st.B[0][PAR][1] = pin0.y; st.B[0][PAR][2] = pin1.x; st.B[0][PAR][3] = pin1.y;
    st.cprev[0] = cin0.x; st.cprev[1] = cin0.y; st.cprev[2] = cin1.x; st.cprev[3] = cin1.y;
}

// Before step 0: the C'' window is zero (the dummy rows in front of the stream stay finite) and the group-1 rows of
// step 0 are on their way.
template <int K>
__device__ __forceinline__ void prologue(WarpState<K> &st, const WarpCtx &c) {
    const double z[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < TM_SLOTS; ++q) tm_st8(c.tm + 8u * q, z);
    tm_wait_st();
    constexpr int L0 = K >= LG ? LG : K;
#ifndef PIPE_ABL_NOTMEM
    tm_ld<8 * L0>(tm_slot_addr<K, 0>(c, 0, K - 1), st.cq);
#endif
}

}  // namespace pipe
}  // namespace nns
