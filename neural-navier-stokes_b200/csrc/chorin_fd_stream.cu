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
#include <vector>

#include "nns_common.cuh"

namespace nns {

namespace {

constexpr int NT_SOR = 256;     // threads of the SOR role (2 warpgroups)
constexpr int NT_ST = 128;      // threads of the stencil role (1 warpgroup)
constexpr int RING = 6;         // rows in flight per field in the stencil ring
constexpr int REGS_SOR = 208, REGS_ST = 88;

// named barrier ids (0 is __syncthreads)
enum { BAR_SOR = 1, BAR_ST = 2, BAR_READY = 3, BAR_CONSUMED = 5, BAR_DONE = 7 };   // +0/+1 by member parity

struct SBlock {          // one per SOR thread (host-built)
    short r0, c0;        // first interior row / column of the block
    short nN, nS, nW, nE;   // thread ids of the neighbouring blocks, -1 = physical boundary
    short bd;            // block anti-diagonal
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
};

struct Coef {
    double ca, cb, cc, cu, cv, beta, tol;
};

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
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
}
// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

// One lexicographic sweep over the thread's own block (all operands in registers), then publication
// of the block perimeter for the neighbours (see chorin_fd_chip.cu for the ordering argument).
template <int BR, int BC, bool TRACK>
__device__ __forceinline__ bool block_sweep(double (&P)[BR][BC], const double2 *__restrict__ Cme, const SHalo<BR, BC> &h,
                                            const Coef &k, unsigned long long tolbits) {
    double hn[BC], hs[BC], hw[BR], he[BR];
#pragma unroll
    for (int lj = 0; lj < BC; ++lj) { hn[lj] = h.hN[lj * NT_SOR]; hs[lj] = h.hS[lj * NT_SOR]; }
#pragma unroll
    for (int li = 0; li < BR; ++li) { hw[li] = h.hW[li * NT_SOR]; he[li] = h.hE[li * NT_SOR]; }
    bool viol = false;
#pragma unroll
    for (int li = 0; li < BR; ++li) {
#pragma unroll
        for (int lj = 0; lj < BC; ++lj) {
            const int q = li * BC + lj;
            const double2 cc2 = Cme[(q >> 1) * NT_SOR];
            const double cp = (q & 1) ? cc2.y : cc2.x;
            const double n = li > 0 ? P[li - 1][lj] : hn[lj];
            const double w = lj > 0 ? P[li][lj - 1] : hw[li];
            const double s = li < BR - 1 ? P[li + 1][lj] : hs[lj];
            const double e = lj < BC - 1 ? P[li][lj + 1] : he[li];
            const double c = P[li][lj];
            const double z = fma(k.ca, s, fma(k.cb, e, fma(-k.beta, c, -cp)));
            const double d = fma(k.ca, n, fma(k.cb, w, z));
            P[li][lj] = c + d;
            if (TRACK) viol |= exceeds_bits(d, tolbits);
        }
    }
#pragma unroll
    for (int lj = 0; lj < BC; ++lj) {
        if (h.pubT) h.Hme[lj * NT_SOR] = P[0][lj];
        if (h.pubB) h.Hme[(BC + lj) * NT_SOR] = P[BR - 1][lj];
    }
#pragma unroll
    for (int li = 0; li < BR; ++li) {
        if (h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][0];
        if (h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][BC - 1];
    }
    return viol;
}

template <int BR, int BC, bool TRACK>
__device__ __forceinline__ void wavefront(double (&P)[BR][BC], const double2 *Cme, const SHalo<BR, BC> &h, bool owner,
                                          int bd, int tmax, int cap, const Coef &k, unsigned long long tolbits,
                                          unsigned long long &mask) {
    if (owner) {   // publish the whole perimeter once
#pragma unroll
        for (int lj = 0; lj < BC; ++lj) {
            if (h.pubT) h.Hme[lj * NT_SOR] = P[0][lj];
            if (h.pubB) h.Hme[(BC + lj) * NT_SOR] = P[BR - 1][lj];
        }
#pragma unroll
        for (int li = 0; li < BR; ++li) {
            if (h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = P[li][0];
            if (h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = P[li][BC - 1];
        }
    }
    named_sync(BAR_SOR, NT_SOR);
    for (int T = 0; T <= tmax; ++T) {
        const int two_s = T - bd;
        const bool work = owner && !(two_s & 1) && (unsigned)two_s <= (unsigned)(2 * (cap - 1));
        if (work) {
            const bool v = block_sweep<BR, BC, TRACK>(P, Cme, h, k, tolbits);
            if (TRACK) mask |= (unsigned long long)v << (two_s >> 1);
        }
        named_sync(BAR_SOR, NT_SOR);
    }
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
    static constexpr size_t ROWBUF_BYTES = sizeof(double) * 2 * NY;
    static constexpr size_t SMEM_BYTES = CS_BYTES + H_BYTES + RING_BYTES + ROWBUF_BYTES;
    static_assert(NB <= NT_SOR, "one SOR thread per block");
    static_assert(NY == NT_ST, "the stencil role maps one thread to one column");
    static_assert((NY * sizeof(double)) % 16 == 0, "bulk copies need 16-byte rows");
};

// ----------------------------------------------------------------------------------------------
// stencil role: row streamer.  Rows of F fields travel global -> smem ring by bulk copies issued by
// one thread; row r of a pass has sequence number seq0 + r, slot = seq % RING, parity = (seq/RING)&1.
// ----------------------------------------------------------------------------------------------
template <int NY>
struct Ring {
    double *buf;          // [RING][4][NY]
    uint64_t *full;       // [RING]
    unsigned seq0;        // sequence number of row 0 of the current pass

    __device__ __forceinline__ double *row(unsigned r, int f) const { return buf + (((seq0 + r) % RING) * 4 + f) * NY; }
    template <int F>
    __device__ __forceinline__ void issue(unsigned r, const double *const (&src)[F]) const {
        uint64_t *b = full + (seq0 + r) % RING;
        mbar_expect_tx(b, F * NY * (uint32_t)sizeof(double));
#pragma unroll
        for (int f = 0; f < F; ++f) bulk_g2s(row(r, f), src[f] + (size_t)r * NY, NY * (uint32_t)sizeof(double), b);
    }
    __device__ __forceinline__ void wait(unsigned r) const {
        const unsigned seq = seq0 + r;
        mbar_wait(full + seq % RING, (seq / RING) & 1u);
    }
};

template <typename C>
__device__ void stencil_pass1(const StreamArgs &a, Ring<C::NY> &ring, double *rowbuf, int m, int ts, double *img) {
    constexpr int NX = C::NX, NY = C::NY;
    const size_t N = (size_t)NX * NY;
    const double *src[4] = {a.uc + m * N, a.vc + m * N, a.up + m * N, a.vp + m * N};
    double *un = a.un + m * N, *vn = a.vn + m * N;
    const double nu = a.nu_b ? a.nu_b[m] : a.g.nu;
    const double *bcval = a.bcval ? a.bcval + (size_t)m * a.n_bcs : nullptr;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy, rho = a.g.rho, beta = a.g.beta;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double r2dx = 1.0 / (2.0 * dx), r2dy = 1.0 / (2.0 * dy), rdx2 = 1.0 / dx2, rdy2 = 1.0 / dy2;
    const double den = 2.0 * dx2 + 2.0 * dy2;
    const double cc = beta / den, cu = dx * rho * dy2 / dt, cv = dy * rho * dx2 / dt;
    const int j = ts;
    const bool jin = j > 0 && j < NY - 1;
    // column part of the C' image address (block column, column inside the block)
    const int bj = jin ? (j - 1) / C::BCc : 0, lj = jin ? (j - 1) - bj * C::BCc : 0;

    // pull the next member's pressure towards L2 while we are at it (the SOR role loads it soon)
    if (m + (int)gridDim.x < a.count) {
        const char *pp = reinterpret_cast<const char *>(a.p + (size_t)(m + gridDim.x) * N);
        for (size_t off = (size_t)ts * 128; off < N * sizeof(double); off += (size_t)NT_ST * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + off));
    }

    if (ts == 0)
        for (unsigned r = 0; r < (unsigned)RING && r < (unsigned)NX; ++r) ring.template issue<4>(r, src);

    double uN = 0, uC = 0, uS = 0, vN = 0, vC = 0, vS = 0, aN = 0, aC = 0, aS = 0, bN = 0, bC = 0, bS = 0;
    double ru_prev = 0.0;
    ring.wait(0);
    uS = ring.row(0, 0)[j]; vS = ring.row(0, 1)[j]; aS = ring.row(0, 2)[j]; bS = ring.row(0, 3)[j];
    for (int i = 0; i < NX; ++i) {
        uN = uC; vN = vC; aN = aC; bN = bC;
        uC = uS; vC = vS; aC = aS; bC = bS;
        if (i + 1 < NX) {
            ring.wait(i + 1);
            uS = ring.row(i + 1, 0)[j]; vS = ring.row(i + 1, 1)[j]; aS = ring.row(i + 1, 2)[j]; bS = ring.row(i + 1, 3)[j];
        }
        double ru = uC, rv = vC;
        const bool interior = jin && i > 0 && i < NX - 1;
        if (interior) {
            const double *ur = ring.row(i, 0), *vr = ring.row(i, 1), *ar = ring.row(i, 2), *br = ring.row(i, 3);
            const double uE = ur[j + 1], uW = ur[j - 1], vE = vr[j + 1], vW = vr[j - 1];
            const double pE = ar[j + 1], pW = ar[j - 1], qE = br[j + 1], qW = br[j - 1];
            // both advection terms difference along axis 0 (chorin_fd:74,76,83,85)
            const double k0 = uC * r2dx + vC * r2dy, k1 = aC * r2dx + bC * r2dy;
            const double advu = 1.5 * (k0 * (uS - uN)) - 0.5 * (k1 * (aS - aN));
            const double advv = 1.5 * (k0 * (vS - vN)) - 0.5 * (k1 * (bS - bN));
            const double lapu = 1.5 * ((uS - 2.0 * uC + uN) * rdx2 + (uE - 2.0 * uC + uW) * rdy2) -
                                0.5 * ((aS - 2.0 * aC + aN) * rdx2 + (pE - 2.0 * aC + pW) * rdy2);
            const double lapv = 1.5 * ((vS - 2.0 * vC + vN) * rdx2 + (vE - 2.0 * vC + vW) * rdy2) -
                                0.5 * ((bS - 2.0 * bC + bN) * rdx2 + (qE - 2.0 * bC + qW) * rdy2);
            ru = uC - dt * advu + (dt * nu) * lapu;
            rv = vC - dt * advv + (dt * nu) * lapv;
        }
        un[(size_t)i * NY + j] = ru;
        vn[(size_t)i * NY + j] = rv;
        rowbuf[(i & 1) * NY + j] = rv;
        named_sync(BAR_ST, NT_ST);          // row i consumed by everyone: its ring slot is free, rowbuf is complete
        if (ts == 0 && i + RING < NX) ring.template issue<4>(i + RING, src);
        if (interior) {
            // C' = beta/den * (dx rho dy^2/dt (ui[i,j]-ui[i-1,j]) + dy rho dx^2/dt (vi[i,j]-vi[i,j-1]))   (:186-188);
            // row 1 / column 1 use pre-BC edge values here and are patched after the BC pass
            const double rv_w = rowbuf[(i & 1) * NY + j - 1];
            const double c = cc * (cu * (ru - ru_prev) + cv * (rv - rv_w));
            const int bi = (i - 1) / C::BRc, li = (i - 1) - bi * C::BRc;
            const int t = a.tidmap[bi * C::NBCc + bj], q = li * C::BCc + lj;
            img[((size_t)(q >> 1) * NT_SOR + t) * 2 + (q & 1)] = c;
        }
        ru_prev = ru;
    }
    ring.seq0 += NX;
    __threadfence_block();
    named_sync(BAR_ST, NT_ST);
    st_apply_bc_global(un, NX, NY, a.ubc, bcval, dx, dy, ts);
    st_apply_bc_global(vn, NX, NY, a.vbc, bcval, dx, dy, ts);
    // patch C' on row 1 (thread j) and column 1 (thread i), which read the boundary lines
    if (jin) {
        const double c = cc * (cu * (un[(size_t)NY + j] - un[j]) + cv * (vn[(size_t)NY + j] - vn[(size_t)NY + j - 1]));
        const int t = a.tidmap[bj], q = lj;
        img[((size_t)(q >> 1) * NT_SOR + t) * 2 + (q & 1)] = c;
    }
    for (int i = 2 + ts; i < NX - 1; i += NT_ST) {
        const size_t g = (size_t)i * NY + 1;
        const double c = cc * (cu * (un[g] - un[g - NY]) + cv * (vn[g] - vn[g - 1]));
        const int bi = (i - 1) / C::BRc, li = (i - 1) - bi * C::BRc;
        const int t = a.tidmap[bi * C::NBCc], q = li * C::BCc;
        img[((size_t)(q >> 1) * NT_SOR + t) * 2 + (q & 1)] = c;
    }
    __threadfence();                      // the image is pulled through L2 (cp.async.cg) by the SOR role
}

template <typename C>
__device__ void stencil_pass2(const StreamArgs &a, Ring<C::NY> &ring, int m, int ts) {
    constexpr int NX = C::NX, NY = C::NY;
    const size_t N = (size_t)NX * NY;
    double *pg = a.p + m * N, *un = a.un + m * N, *vn = a.vn + m * N;
    const double *bcval = a.bcval ? a.bcval + (size_t)m * a.n_bcs : nullptr;
    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy;
    st_apply_bc_global(pg, NX, NY, a.pbc, bcval, dx, dy, ts);
    __threadfence();                      // generic-proxy writes of p / un / vn (this CTA) before the bulk reads
    named_sync(BAR_ST, NT_ST);
    const double *src[3] = {pg, un, vn};
    const int j = ts;
    const bool jin = j > 0 && j < NY - 1;
    const double kx = dt / (2.0 * dx), ky = dt / (2.0 * dy);
    const size_t toff = (size_t)m * a.traj_member_stride + a.traj_off;
    if (ts == 0) {
        fence_proxy_async();
        for (unsigned r = 0; r < (unsigned)RING && r < (unsigned)NX; ++r) ring.template issue<3>(r, src);
    }
    double pN = 0, pC = 0, pS = 0;
    unsigned long long bad = 0;
    ring.wait(0);
    pS = ring.row(0, 0)[j];
    for (int i = 0; i < NX; ++i) {
        pN = pC; pC = pS;
        if (i + 1 < NX) {
            ring.wait(i + 1);
            pS = ring.row(i + 1, 0)[j];
        }
        double ru = ring.row(i, 1)[j], rv = ring.row(i, 2)[j];
        const size_t q = (size_t)i * NY + j;
        if (jin && i > 0 && i < NX - 1) {
            const double *pr = ring.row(i, 0);
            ru -= kx * (pS - pN);
            rv -= ky * (pr[j + 1] - pr[j - 1]);
            un[q] = ru;
            vn[q] = rv;
        }
        if (a.traj_u) { a.traj_u[toff + q] = ru; a.traj_v[toff + q] = rv; a.traj_p[toff + q] = pC; }
        if (a.flags & NNS_FLAG_CHECK_FINITE) bad += !(isfinite(ru) && isfinite(rv) && isfinite(pC));
        named_sync(BAR_ST, NT_ST);
        if (ts == 0 && i + RING < NX) ring.template issue<3>(i + RING, src);
    }
    ring.seq0 += NX;
    if ((a.flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(a.nonfinite, bad);
}

// ----------------------------------------------------------------------------------------------
template <int BR, int BC, int NBR, int NBC>
struct CfgX : Cfg<BR, BC, NBR, NBC> {
    static constexpr int BRc = BR, BCc = BC, NBRc = NBR, NBCc = NBC;
};

template <typename C>
__global__ void __launch_bounds__(NT_SOR + NT_ST, 1) chorin_stream_kernel(const StreamArgs a) {
    constexpr int BR = C::BRc, BC = C::BCc, NX = C::NX, NY = C::NY, NCH = C::NCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_mask;
    __shared__ int s_need;
    __shared__ __align__(8) uint64_t s_full[RING];

    double2 *Cs = reinterpret_cast<double2 *>(smem_raw);
    double *H = reinterpret_cast<double *>(smem_raw + C::CS_BYTES);
    double *ringbuf = reinterpret_cast<double *>(smem_raw + C::CS_BYTES + C::H_BYTES);
    double *rowbuf = reinterpret_cast<double *>(smem_raw + C::CS_BYTES + C::H_BYTES + C::RING_BYTES);

    const int tid = threadIdx.x;
    const size_t N = (size_t)NX * NY;
    const int nmine = a.count > (int)blockIdx.x ? (a.count - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    double *img = a.cimg + (size_t)blockIdx.x * 2 * NCH * NT_SOR;

    if (tid == 0) {
        for (int r = 0; r < RING; ++r) mbar_init(&s_full[r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_mask = 0ull;
    }
    __syncthreads();
    if (nmine == 0) return;

    if (tid >= NT_SOR) {
        // =============================== stencil role ===========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
        const int ts = tid - NT_SOR;
        Ring<NY> ring{ringbuf, s_full, 0u};
        stencil_pass1<C>(a, ring, rowbuf, blockIdx.x, ts, img);
        named_arrive(BAR_READY + 0, NT_SOR + NT_ST);
        for (int k = 0; k < nmine; ++k) {
            const int m = blockIdx.x + k * gridDim.x;
            if (k + 1 < nmine) {
                named_sync(BAR_CONSUMED + (k & 1), NT_SOR + NT_ST);       // the SOR role has pulled image k
                stencil_pass1<C>(a, ring, rowbuf, m + gridDim.x, ts, img);
                named_arrive(BAR_READY + ((k + 1) & 1), NT_SOR + NT_ST);
            }
            named_sync(BAR_DONE + (k & 1), NT_SOR + NT_ST);               // SOR of member k finished, p written
            stencil_pass2<C>(a, ring, m, ts);
        }
    } else {
        // ================================= SOR role =============================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR));
        const bool owner = tid < C::NB;
        SBlock ds{1, 1, -1, -1, -1, -1, 0, 0};
        if (owner) ds = a.desc[tid];
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
        k.ca = beta * dy2 / den; k.cb = beta * dx2 / den; k.cc = 0; k.cu = 0; k.cv = 0; k.beta = beta; k.tol = a.g.tol;
        const unsigned long long tolbits = (unsigned long long)__double_as_longlong(a.g.tol);
        const int cap = a.g.nit - 1;
        const int tmax = C::NBRc + C::NBCc - 2 + 2 * (cap - 1);

        for (int kk = 0; kk < nmine; ++kk) {
            const int m = blockIdx.x + kk * gridDim.x;
            double *pg = a.p + (size_t)m * N;
            double P[BR][BC];
            auto load_block = [&]() {
#pragma unroll
                for (int li = 0; li < BR; ++li)
#pragma unroll
                    for (int lj = 0; lj < BC; ++lj) P[li][lj] = owner ? pg[(size_t)(r0 + li) * NY + c0 + lj] : 0.0;
            };
            load_block();                          // p is not touched by the stencil role before SOR finishes
            if (owner) {                           // frozen boundary values into the unread own slots
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) {
                    if (!h.pubT) h.Hme[lj * NT_SOR] = pg[(size_t)(r0 - 1) * NY + c0 + lj];
                    if (!h.pubB) h.Hme[(BC + lj) * NT_SOR] = pg[(size_t)(r0 + BR) * NY + c0 + lj];
                }
#pragma unroll
                for (int li = 0; li < BR; ++li) {
                    if (!h.pubL) h.Hme[(2 * BC + li) * NT_SOR] = pg[(size_t)(r0 + li) * NY + c0 - 1];
                    if (!h.pubR) h.Hme[(2 * BC + BR + li) * NT_SOR] = pg[(size_t)(r0 + li) * NY + c0 + BC];
                }
            }
            named_sync(BAR_READY + (kk & 1), NT_SOR + NT_ST);            // C' image of member kk is complete
            {
                const double2 *gi = reinterpret_cast<const double2 *>(img) + tid;
#pragma unroll
                for (int c = 0; c < NCH; ++c) cp_async16(Cs + c * NT_SOR + tid, gi + c * NT_SOR);
                cp_async_wait_all();
            }
            if (tid == 0) s_mask = 0ull;
            named_sync(BAR_SOR, NT_SOR);
            if (kk + 1 < nmine) named_arrive(BAR_CONSUMED + (kk & 1), NT_SOR + NT_ST);

            int need = 0;
            if (cap > 0) {
                unsigned long long mask = 0ull;
                wavefront<BR, BC, true>(P, Cme, h, owner, ds.bd, tmax, cap, k, tolbits, mask);
                unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mask);
                unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mask >> 32));
                if ((tid & 31) == 0) atomicOr(&s_mask, ((unsigned long long)hi << 32) | lo);
                named_sync(BAR_SOR, NT_SOR);
                if (tid == 0) {
                    const unsigned long long full = cap >= 64 ? ~0ull : ((1ull << cap) - 1ull);
                    const unsigned long long clr = ~s_mask & full;
                    s_need = clr ? __ffsll((long long)clr) : cap;
                }
                named_sync(BAR_SOR, NT_SOR);
                need = s_need;
                if (need < cap) {
                    // the sequential loop would have stopped after `need` sweeps: redo from the
                    // untouched global p with the sweep count capped (rare: near steady state)
                    load_block();
                    unsigned long long dummy = 0ull;
                    wavefront<BR, BC, false>(P, Cme, h, owner, ds.bd, C::NBRc + C::NBCc - 2 + 2 * (need - 1), need, k,
                                             tolbits, dummy);
                }
            }
            if (owner) {
#pragma unroll
                for (int li = 0; li < BR; ++li)
#pragma unroll
                    for (int lj = 0; lj < BC; ++lj) pg[(size_t)(r0 + li) * NY + c0 + lj] = P[li][lj];
            }
            if (a.sweeps && tid == 0) a.sweeps[m] = need;
            __threadfence();              // p is re-read by the stencil role, partly through the TMA unit
            named_arrive(BAR_DONE + (kk & 1), NT_SOR + NT_ST);
        }
    }
}

using Cfg128 = CfgX<9, 7, 14, 18>;      // 128 x 128: 126 = 14*9 = 18*7, 252 blocks of 63 cells

struct StreamPlan {
    std::vector<SBlock> desc;
    std::vector<short> tidmap;
    void *d_tab = nullptr;      // desc then tidmap
    double *d_img = nullptr;
    int grid = 0;
};

template <typename C>
static void build_tables(StreamPlan &pl) {
    constexpr int BR = C::BRc, BC = C::BCc, NBR = C::NBRc, NBC = C::NBCc, NB = NBR * NBC;
    struct Item { int key, bi, bj; };
    std::vector<Item> items;
    for (int bi = 0; bi < NBR; ++bi)
        for (int bj = 0; bj < NBC; ++bj) items.push_back({bi + bj, bi, bj});
    // blocks of one parity work in the same super-stage: keep them in the same warps, ordered by diagonal
    std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) {
        if ((x.key & 1) != (y.key & 1)) return (x.key & 1) < (y.key & 1);
        return x.key != y.key ? x.key < y.key : x.bj < y.bj;
    });
    std::vector<int> tid_of((size_t)NB);
    for (int t = 0; t < NB; ++t) tid_of[(size_t)items[t].bi * NBC + items[t].bj] = t;
    pl.desc.assign(NT_SOR, SBlock{1, 1, -1, -1, -1, -1, 0, 0});
    pl.tidmap.resize(NB);
    for (int t = 0; t < NB; ++t) pl.tidmap[t] = (short)tid_of[t];
    for (int t = 0; t < NB; ++t) {
        const int bi = items[t].bi, bj = items[t].bj;
        SBlock d;
        d.r0 = (short)(1 + BR * bi);
        d.c0 = (short)(1 + BC * bj);
        d.nN = bi > 0 ? (short)tid_of[(size_t)(bi - 1) * NBC + bj] : (short)-1;
        d.nS = bi < NBR - 1 ? (short)tid_of[(size_t)(bi + 1) * NBC + bj] : (short)-1;
        d.nW = bj > 0 ? (short)tid_of[(size_t)bi * NBC + bj - 1] : (short)-1;
        d.nE = bj < NBC - 1 ? (short)tid_of[(size_t)bi * NBC + bj + 1] : (short)-1;
        d.bd = (short)(bi + bj);
        d.pad = 0;
        pl.desc[t] = d;
    }
}

}  // namespace

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
        pl->grid = h->sm_count;
        NNS_CUDA(cudaMalloc(&pl->d_img, sizeof(double2) * C::NCH * NT_SOR * (size_t)pl->grid));
        NNS_CUDA(cudaMemset(pl->d_img, 0, sizeof(double2) * C::NCH * NT_SOR * (size_t)pl->grid));
        NNS_CUDA(cudaFuncSetAttribute(chorin_stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)C::SMEM_BYTES));
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
    a.cimg = pl->d_img;
    a.traj_u = tu; a.traj_v = tv; a.traj_p = tp;
    a.traj_member_stride = traj_member_stride; a.traj_off = traj_off;
    a.sweeps = sweeps;
    a.nonfinite = h->d_nonfinite;
    const int grid = count < pl->grid ? count : pl->grid;
    chorin_stream_kernel<C><<<grid, NT_SOR + NT_ST, C::SMEM_BYTES, st>>>(a);
    NNS_CUDA(cudaGetLastError());
    h->launches += 1;
    return NNS_OK;
}

}  // namespace nns
