// sor_block.cuh -- the register-block SOR sweep shared by chorin_fd_stream.cu and the sweep
// microbenchmark (scripts/micro/sweep_bench.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nns {

constexpr int NT_SOR = 256;     // threads of the SOR role (2 warpgroups); stride of the smem tables

struct Coef {
    double ca, cb, cc, cu, cv, beta, tol;
};

// |d| <= tol decided on the INTEGER pipe (the FP64 pipe is the bottleneck of the sweeps): for finite
// doubles the magnitude order equals the order of the bit patterns; NaN patterns exceed every finite
// one, so NaN counts as "not converged", like !(fabs(d) <= tol).
__device__ __forceinline__ bool exceeds_bits(double d, unsigned long long tolbits) {
    return ((unsigned long long)__double_as_longlong(d) & 0x7fffffffffffffffull) > tolbits;
}

template <int BR, int BC>
struct SHalo {
    const double *hN, *hS, *hW, *hE;
    double *Hme;
    bool pubT, pubB, pubL, pubR;
};

// Publish the perimeter cells of rows [R0, R1) of the block into the thread's halo slots
// ([0,BC) top row, [BC,2BC) bottom row, [2BC,2BC+BR) left column, [2BC+BR,2BC+2BR) right column);
// slots facing a physical boundary hold the frozen boundary values of p and are never republished.
template <int BR, int BC, int R0, int R1>
__device__ __forceinline__ void publish(const double (&P)[BR][BC], const SHalo<BR, BC> &h) {
#pragma unroll
    for (int lj = 0; lj < BC; ++lj) {
        if (R0 == 0 && h.pubT) h.Hme[lj * NT_SOR] = P[0][lj];
        if (R1 == BR && h.pubB) h.Hme[(BC + lj) * NT_SOR] = P[BR - 1][lj];
    }
#pragma unroll
    for (int li = R0; li < R1; ++li) {
        if (h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][0];
        if (h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][BC - 1];
    }
}

// Position of cell (li, lj) of an R x BC (sub-)block in ANTI-DIAGONAL-major order (diagonal li + lj, then
// li): the order in which a sweep consumes the right-hand side.
template <int R, int BC>
__host__ __device__ constexpr int diag_ord(int li, int lj) {
    const int kd = li + lj;
    int o = 0;
    for (int t = 0; t < kd; ++t) {
        const int lo = t - (BC - 1) > 0 ? t - (BC - 1) : 0, hi = t < R - 1 ? t : R - 1;
        o += hi - lo + 1;
    }
    const int lo = kd - (BC - 1) > 0 ? kd - (BC - 1) : 0;
    return o + (li - lo);
}

// Layout of the C' chunks of a thread whose BR x BC block is swept as two sub-blocks, rows [0, RS) and
// [RS, BR): the cells of the top sub-block in its diagonal order, then those of the bottom one.
template <int BR, int BC, int RS>
__host__ __device__ constexpr int split_ord(int li, int lj) {
    if (RS < 0) {       // three sub-blocks of -RS rows each (BR == -3 * RS)
        constexpr int H = RS < 0 ? -RS : 1;
        const int kb = li / H;
        return kb * H * BC + diag_ord<H, BC>(li - kb * H, lj);
    }
    return li < RS ? diag_ord<(RS > 0 ? RS : 1), BC>(li, lj) : RS * BC + diag_ord<(RS > 0 ? BR - RS : 1), BC>(li - RS, lj);
}

// One sweep over rows [R0, R1) of the thread's own BR x BC block (the block is swept as two sub-blocks
// in alternate super-stages, see chorin_fd_stream.cu), all operands in registers, then publication of
// that sub-block's part of the perimeter for the neighbouring threads.
// Inside the sub-block the lexicographic order of the reference is executed as an anti-diagonal
// wavefront: cells of one diagonal only depend on the previous diagonal (north, west: new values)
// and on later diagonals (south, east: old values), so the result is the sequential one, while the
// instruction stream carries up to min(R1-R0, BC) independent dependency chains (the FP64 pipe has an
// 8-cycle latency and a 2-cycle issue interval per warp).
//   stage A (old operands):  base = ca*s + cb*e - beta*c - C'
//   stage B (new operands):  d = ca*n + cb*w + base, p += d     (2 dependent DFMAs + 1 DADD per cell)
// North of row R0 is the thread's own row R0-1 (already swept this sweep) or the halo of the block above;
// south of row R1-1 is the own row R1 (not yet swept) or the halo of the block below.
// TRACK: 0 none; 1 fast: running maximum of the high words of |p' - p| on the integer pipe (decides
// "max|dp| <= tol" unless the maximum shares its high word with tol: ambiguous, resolved by an exact
// re-run); 2 exact 64-bit comparison per cell.
// Running maximum of the high words of |d| for the fast exit test.  The high word of a non-negative double orders like
// the same bits read as a float, so the maximum runs on the FP32 pipe (FMNMX with the |x| operand modifier); the
// integer form (LOP3 + VIMNMX3) issues on a slow pipe on B200 (scripts/micro/pipe_bench.cu: a third of the step).
// max.NaN keeps a high word >= 0x7f800001 (|d| >= 2^1017, inf or NaN: a float NaN pattern) as a NaN, which compares
// above every tolerance as an unsigned integer -- "violated", as the exact test says.  Denormal patterns are kept
// (the library is built without -ftz).
__device__ __forceinline__ unsigned track_hi(unsigned m, double d) {
#ifdef NNS_TRACK_INT
    return max(m, (unsigned)__double2hiint(d) & 0x7fffffffu);
#else
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(__uint_as_float(m)), "f"(fabsf(__int_as_float(__double2hiint(d)))));
    return __float_as_uint(r);
#endif
}

template <int BR, int BC, int RS, int R0, int R1, int TRACK>
__device__ __forceinline__ void block_sweep(double (&P)[BR][BC], const double2 *__restrict__ Cme, const SHalo<BR, BC> &h,
                                            const Coef &k, unsigned long long tolbits, unsigned &mhi, bool &viol) {
#ifdef NNS_ABL_NOCOMPUTE
    publish<BR, BC, R0, R1>(P, h);
    return;
#endif
    constexpr int NR = R1 - R0, ND = NR + BC - 1;
    double base[2][NR];
    auto stageA = [&](int kd, double (&out)[NR]) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const int o = split_ord<BR, BC, RS>(li, lj);
#ifdef NNS_ABL_NOCPRIME      // timing ablations (scripts/ablate.sh): results are wrong on purpose
                const double cp = 1e-3 * o;
#else
                const double2 cc2 = Cme[(o >> 1) * NT_SOR];
                const double cp = (o & 1) ? cc2.y : cc2.x;
#endif
#ifdef NNS_ABL_NOHALO
                const double s = li < BR - 1 ? P[li + 1][lj] : 0.5;
                const double e = lj < BC - 1 ? P[li][lj + 1] : 0.25;
#else
                const double s = li < BR - 1 ? P[li + 1][lj] : h.hS[lj * NT_SOR];
                const double e = lj < BC - 1 ? P[li][lj + 1] : h.hE[li * NT_SOR];
#endif
                out[r] = fma(k.ca, s, fma(k.cb, e, fma(k.cc, P[li][lj], -cp)));      // k.cc = -beta (1 - beta with NNS_SOR_FORM_PN)
            }
        }
    };
    stageA(0, base[0]);
#pragma unroll
    for (int kd = 0; kd < ND; ++kd) {
        if (kd + 1 < ND) stageA(kd + 1, base[(kd + 1) & 1]);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
#ifdef NNS_ABL_NOHALO
                const double n = li > 0 ? P[li - 1][lj] : 0.125;
                const double w = lj > 0 ? P[li][lj - 1] : 0.375;
#else
                const double n = li > 0 ? P[li - 1][lj] : h.hN[lj * NT_SOR];
                const double w = lj > 0 ? P[li][lj - 1] : h.hW[li * NT_SOR];
#endif
#ifdef NNS_SOR_FORM_PN       // p' in 5 DFMAs, d = p' - p for the exit test (shorter chain, but needs register moves)
                const double pn = fma(k.ca, n, fma(k.cb, w, base[kd & 1][r]));
                if (TRACK) {
                    const double d = pn - P[li][lj];
                    if (TRACK == 1) mhi = track_hi(mhi, d);
                    else viol |= exceeds_bits(d, tolbits);
                }
                P[li][lj] = pn;
#else                       // d in 5 DFMAs, p += d in place (no register renaming at the loop back-edge)
                const double d = fma(k.ca, n, fma(k.cb, w, base[kd & 1][r]));
                if (TRACK == 1) mhi = track_hi(mhi, d);
                if (TRACK == 2) viol |= exceeds_bits(d, tolbits);
                P[li][lj] += d;
#endif
#if !defined(NNS_ABL_NOPUBLISH) && defined(NNS_SOR_PUBLISH_EARLY)     // measured slower on B200 (4.81 vs 4.58 ms/step)
                // publish perimeter cells as soon as they are final: the stores drain under the remaining
                // diagonals instead of in front of the stage barrier
                if (li == 0 && h.pubT) h.Hme[lj * NT_SOR] = P[li][lj];
                if (li == BR - 1 && h.pubB) h.Hme[(BC + lj) * NT_SOR] = P[li][lj];
                if (lj == 0 && h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][lj];
                if (lj == BC - 1 && h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][lj];
#endif
            }
        }
    }
#if !defined(NNS_ABL_NOPUBLISH) && !defined(NNS_SOR_PUBLISH_EARLY)
    publish<BR, BC, R0, R1>(P, h);
#endif
}

// number of cells of an NR x BC sub-block on anti-diagonals < t
template <int NR, int BC>
__host__ __device__ constexpr int diag_cum(int t) {
    int o = 0;
    for (int d = 0; d < t; ++d) {
        const int lo = d - (BC - 1) > 0 ? d - (BC - 1) : 0, hi = d < NR - 1 ? d : NR - 1;
        if (hi >= lo) o += hi - lo + 1;
    }
    return o;
}

// block_sweep with the loads of the right-hand side issued PFD anti-diagonals ahead of their use (explicit chunk
// schedule, volatile loads).  With C' re-read from shared memory in every sweep the plain block_sweep is bound by
// the exposed load latency at every diagonal step (890 cycles per 31.5-cell sweep against 536 when C' sits in
// registers, scripts/micro/lock_bench.cu): there are no spare registers for the compiler to hoist the loads.
__device__ __forceinline__ double2 lds_f64x2_volatile(const double2 *p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
template <int BR, int BC, int RS, int R0, int R1, int TRACK, int PFD>
__device__ __forceinline__ void block_sweep_pf(double (&P)[BR][BC], const double2 *__restrict__ Cme, const SHalo<BR, BC> &h,
                                               const Coef &k, unsigned long long tolbits, unsigned &mhi, bool &viol) {
    constexpr int NR = R1 - R0, ND = NR + BC - 1;
    constexpr int OB = R0 == 0 ? 0 : RS * BC;
    constexpr int CLO = OB >> 1, CHI = (OB + NR * BC - 1) >> 1, NCQ = CHI - CLO + 1;
    double2 cq[NCQ];
    auto chi = [](int kd) { return kd < 0 ? CLO - 1 : kd >= ND ? CHI : (OB + diag_cum<NR, BC>(kd + 1) - 1) >> 1; };
    auto issue = [&](int from, int to) {
#pragma unroll
        for (int c = CLO; c <= CHI; ++c)
            if (c > from && c <= to) cq[c - CLO] = lds_f64x2_volatile(Cme + c * NT_SOR);
    };
    double base[2][NR];
    auto stageA = [&](int kd, double (&out)[NR]) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const int o = split_ord<BR, BC, RS>(li, lj);
                const double cp = (o & 1) ? cq[(o >> 1) - CLO].y : cq[(o >> 1) - CLO].x;
                const double s = li < BR - 1 ? P[li + 1][lj] : h.hS[lj * NT_SOR];
                const double e = lj < BC - 1 ? P[li][lj + 1] : h.hE[li * NT_SOR];
                out[r] = fma(k.ca, s, fma(k.cb, e, fma(k.cc, P[li][lj], -cp)));
            }
        }
    };
    issue(CLO - 1, chi(PFD));                  // diagonals 0 .. PFD
    stageA(0, base[0]);
#pragma unroll
    for (int kd = 0; kd < ND; ++kd) {
        if (kd + 1 < ND) {
            issue(chi(kd + PFD), chi(kd + 1 + PFD));
            stageA(kd + 1, base[(kd + 1) & 1]);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const double n = li > 0 ? P[li - 1][lj] : h.hN[lj * NT_SOR];
                const double w = lj > 0 ? P[li][lj - 1] : h.hW[li * NT_SOR];
                const double d = fma(k.ca, n, fma(k.cb, w, base[kd & 1][r]));
                if (TRACK == 1) mhi = track_hi(mhi, d);
                if (TRACK == 2) viol |= exceeds_bits(d, tolbits);
                P[li][lj] += d;
            }
        }
    }
    publish<BR, BC, R0, R1>(P, h);
}

// ----------------------------------------------------------------------------------------------
// Tensor Memory as a thread-private scratchpad (chorin_wave_kernel of chorin_fd_stream.cu): the right-hand
// side C' of a thread lives in the thread's own TMEM lane (tcgen05.ld/st shape 32x32b: thread i of the warp
// accesses lane 32*(warp%4)+i, consecutive 32-bit columns), in the same 16-byte chunks and the same
// consumption order as the shared-memory layout above.  Measured on B200 (scripts/micro/tmem_bench.cu): x4
// loads sustain ~50 B/clk per SM sub-partition on a pipe of their own, latency ~35 cycles, so the sweeps no
// longer compete with the stencil role for the single shared-memory pipe of the SM.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
// wait for every outstanding tcgen05.ld of the thread; the registers are tied to the statement so that the
// compiler cannot move a use in front of the wait
__device__ __forceinline__ void tm_wait_ld(uint32_t (&r)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
}
__device__ __forceinline__ void tm_tie(uint32_t (&r)[4]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
}


// block_sweep with the right-hand side in Tensor Memory and WARP-UNIFORM control flow (tcgen05.ld is a
// warp-collective instruction): every lane executes the sweep, lanes outside the wavefront band (`act`
// false) compute on whatever their registers hold and commit nothing.  tmc: TMEM address of chunk 0 of the
// member's C' buffer of this thread.  pubTL false: the top row / left column are not published (last sweep
// of a member: nobody reads them, and the neighbour may already have deposited the next member's start
// values in those slots).  The chunk loads run one anti-diagonal ahead of their use.
template <int BR, int BC, int RS, int R0, int R1>
__device__ __forceinline__ void block_sweep_tm(double (&P)[BR][BC], uint32_t tmc, const SHalo<BR, BC> &h, const Coef &k,
                                               bool act, bool pubTL, unsigned &mhi) {
    constexpr int NR = R1 - R0, ND = NR + BC - 1;
    constexpr int OB = R0 == 0 ? 0 : RS * BC;                  // first cell (consumption order) of this sub-block
    constexpr int CLO = OB >> 1, CHI = (OB + NR * BC - 1) >> 1, NCQ = CHI - CLO + 1;
    uint32_t cq[NCQ][4];
    // last chunk needed by the cells of diagonal kd
    auto chi = [](int kd) { return (OB + diag_cum<NR, BC>(kd + 1) - 1) >> 1; };
    auto issue = [&](int from, int to) {       // chunks (from, to]
#pragma unroll
        for (int c = CLO; c <= CHI; ++c)
            if (c > from && c <= to) tm_ld4(tmc + 4 * c, cq[c - CLO]);
    };
    auto wait = [&](int from, int to) {
        bool first = true;
#pragma unroll
        for (int c = CLO; c <= CHI; ++c)
            if (c > from && c <= to) {
                if (first) tm_wait_ld(cq[c - CLO]);
                else tm_tie(cq[c - CLO]);
                first = false;
            }
    };
    const double actf = act ? 1.0 : 0.0;
    double base[2][NR];
    auto stageA = [&](int kd, double (&out)[NR]) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const int o = split_ord<BR, BC, RS>(li, lj);
                const uint32_t(&c4)[4] = cq[(o >> 1) - CLO];
                const double cp = (o & 1) ? __hiloint2double((int)c4[3], (int)c4[2]) : __hiloint2double((int)c4[1], (int)c4[0]);
                const double s = li < BR - 1 ? P[li + 1][lj] : h.hS[lj * NT_SOR];
                const double e = lj < BC - 1 ? P[li][lj + 1] : h.hE[li * NT_SOR];
                out[r] = fma(k.ca, s, fma(k.cb, e, fma(k.cc, P[li][lj], -cp)));
            }
        }
    };
    issue(CLO - 1, chi(1));
    wait(CLO - 1, chi(1));
    if (ND > 2) issue(chi(1), chi(2));
    stageA(0, base[0]);
#pragma unroll
    for (int kd = 0; kd < ND; ++kd) {
        if (kd + 1 < ND) {
            if (kd >= 1) wait(chi(kd), chi(kd + 1));
            if (kd >= 1 && kd + 2 < ND) issue(chi(kd + 1), chi(kd + 2));
            stageA(kd + 1, base[(kd + 1) & 1]);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const double n = li > 0 ? P[li - 1][lj] : h.hN[lj * NT_SOR];
                const double w = lj > 0 ? P[li][lj - 1] : h.hW[li * NT_SOR];
                const double d = fma(k.ca, n, fma(k.cb, w, base[kd & 1][r]));
                mhi = track_hi(mhi, d);
                // p += d for the lanes inside the band, p unchanged (+ 0 * d) for the others: one DFMA, like the
                // DADD of the legacy sweep and bit-identical to it for actf == 1 (a select would cost two extra
                // instructions per cell)
                P[li][lj] = fma(d, actf, P[li][lj]);
            }
        }
    }
    const bool pT = h.pubT && act && pubTL, pB = h.pubB && act, pL = h.pubL && act && pubTL, pR = h.pubR && act;
#pragma unroll
    for (int lj = 0; lj < BC; ++lj) {
        if (R0 == 0 && pT) h.Hme[lj * NT_SOR] = P[0][lj];
        if (R1 == BR && pB) h.Hme[(BC + lj) * NT_SOR] = P[BR - 1][lj];
    }
#pragma unroll
    for (int li = R0; li < R1; ++li) {
        if (pL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][0];
        if (pR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][BC - 1];
    }
}

// The same with the exit test of block_sweep (TRACK 0 / 1 / 2): the member-at-a-time kernel with C' in Tensor Memory.
template <int BR, int BC, int RS, int R0, int R1, int TRACK>
__device__ __forceinline__ void block_sweep_tmt(double (&P)[BR][BC], uint32_t tmc, const SHalo<BR, BC> &h, const Coef &k,
                                                bool act, unsigned long long tolbits, unsigned &mhi, bool &viol) {
    constexpr bool pubTL = true;
    constexpr int NR = R1 - R0, ND = NR + BC - 1;
    constexpr int OB = R0 == 0 ? 0 : RS * BC;                  // first cell (consumption order) of this sub-block
    constexpr int CLO = OB >> 1, CHI = (OB + NR * BC - 1) >> 1, NCQ = CHI - CLO + 1;
    uint32_t cq[NCQ][4];
    // last chunk needed by the cells of diagonal kd
    auto chi = [](int kd) { return (OB + diag_cum<NR, BC>(kd + 1) - 1) >> 1; };
    auto issue = [&](int from, int to) {       // chunks (from, to]
#pragma unroll
        for (int c = CLO; c <= CHI; ++c)
            if (c > from && c <= to) tm_ld4(tmc + 4 * c, cq[c - CLO]);
    };
    auto wait = [&](int from, int to) {
        bool first = true;
#pragma unroll
        for (int c = CLO; c <= CHI; ++c)
            if (c > from && c <= to) {
                if (first) tm_wait_ld(cq[c - CLO]);
                else tm_tie(cq[c - CLO]);
                first = false;
            }
    };
    const double actf = act ? 1.0 : 0.0;
    double base[2][NR];
    auto stageA = [&](int kd, double (&out)[NR]) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const int o = split_ord<BR, BC, RS>(li, lj);
                const uint32_t(&c4)[4] = cq[(o >> 1) - CLO];
                const double cp = (o & 1) ? __hiloint2double((int)c4[3], (int)c4[2]) : __hiloint2double((int)c4[1], (int)c4[0]);
                const double s = li < BR - 1 ? P[li + 1][lj] : h.hS[lj * NT_SOR];
                const double e = lj < BC - 1 ? P[li][lj + 1] : h.hE[li * NT_SOR];
                out[r] = fma(k.ca, s, fma(k.cb, e, fma(k.cc, P[li][lj], -cp)));
            }
        }
    };
    issue(CLO - 1, chi(1));
    wait(CLO - 1, chi(1));
    if (ND > 2) issue(chi(1), chi(2));
    stageA(0, base[0]);
#pragma unroll
    for (int kd = 0; kd < ND; ++kd) {
        if (kd + 1 < ND) {
            if (kd >= 1) wait(chi(kd), chi(kd + 1));
            if (kd >= 1 && kd + 2 < ND) issue(chi(kd + 1), chi(kd + 2));
            stageA(kd + 1, base[(kd + 1) & 1]);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int lj = kd - r, li = R0 + r;
            if (lj >= 0 && lj < BC) {
                const double n = li > 0 ? P[li - 1][lj] : h.hN[lj * NT_SOR];
                const double w = lj > 0 ? P[li][lj - 1] : h.hW[li * NT_SOR];
                const double d = fma(k.ca, n, fma(k.cb, w, base[kd & 1][r]));
                if (TRACK == 1) mhi = track_hi(mhi, d);
                if (TRACK == 2) viol |= exceeds_bits(d, tolbits);
                // p += d for the lanes inside the band, p unchanged (+ 0 * d) for the others: one DFMA, like the
                // DADD of the legacy sweep and bit-identical to it for actf == 1 (a select would cost two extra
                // instructions per cell)
                P[li][lj] = fma(d, actf, P[li][lj]);
            }
        }
    }
    const bool pT = h.pubT && act && pubTL, pB = h.pubB && act, pL = h.pubL && act && pubTL, pR = h.pubR && act;
#pragma unroll
    for (int lj = 0; lj < BC; ++lj) {
        if (R0 == 0 && pT) h.Hme[lj * NT_SOR] = P[0][lj];
        if (R1 == BR && pB) h.Hme[(BC + lj) * NT_SOR] = P[BR - 1][lj];
    }
#pragma unroll
    for (int li = R0; li < R1; ++li) {
        if (pL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][0];
        if (pR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][BC - 1];
    }
}


}  // namespace nns
