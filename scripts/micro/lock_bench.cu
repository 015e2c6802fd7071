// lock_bench.cu -- steady state of the SOR wavefront: all 8 warps sweep one sub-block per stage (even-parity warps
// the top sub-block on even stages, odd-parity warps the bottom one, and vice versa), with or without the stage
// barrier, right-hand side in shared memory (block_sweep) or in Tensor Memory (block_sweep_tm).  Separates the cost
// of running the two warps of an SM sub-partition in lockstep from everything else in chorin_fd_stream.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o lock_bench lock_bench.cu
#include "../../neural-navier-stokes_b200/csrc/chorin_fd_stream.cu"

namespace nns { void set_error(const char *, ...) {} }
using namespace nns;

template <bool TM, int BARRIER, int PMAP = 0>      // PMAP 1: both warps of an SM sub-partition sweep the same sub-block kind; BARRIER: 0 none, 1 bar.sync of the 8 warps per stage, 2 per-SMSP pair barrier only
__global__ void __launch_bounds__(384, 1) lock_kernel(const SBlock *desc, long long *cyc, int stages, int nwarps) {
    using C = Cfg128;
    constexpr int BR = C::BRc, BC = C::BCc, RS = C::RSc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    double *H = reinterpret_cast<double *>(smem_raw);
    double2 *Cs = reinterpret_cast<double2 *>(smem_raw + C::H_BYTES);
    const int tid = threadIdx.x;
    for (int q = tid; q < C::NSLOT * NT_SOR; q += 384) H[q] = 1e-4 * (q % 101);
    if (!TM) for (int q = tid; q < C::NCH * NT_SOR; q += 384) Cs[q] = make_double2(1e-3 * (q % 97), 2e-3 * (q % 89));
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    if (tid < NT_SOR) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR));
        const int w = tid >> 5;
        const uint32_t tm_mine = tmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(256 * (w >> 2));
        if (TM || PMAP == 6) {
            for (int c = 0; c < 32; c += 4) {
                double2 v[4];
                for (int e = 0; e < 4; ++e) v[e] = make_double2(1e-3 * ((c + e) * 7 + tid % 13), 2e-3 * ((c + e) * 5 + tid % 11));
                tm_st16(tm_mine + 4 * c, v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        SBlock ds = desc[tid];
        const bool owner = ds.r0 > 0;
        SHalo<BR, BC> h;
        h.Hme = H + tid;
        h.pubT = ds.nN >= 0; h.pubB = ds.nS >= 0; h.pubL = ds.nW >= 0; h.pubR = ds.nE >= 0;
        h.hN = h.pubT ? H + BC * NT_SOR + ds.nN : H + tid;
        h.hS = h.pubB ? H + ds.nS : H + BC * NT_SOR + tid;
        h.hW = h.pubL ? H + (2 * BC + BR) * NT_SOR + ds.nW : H + 2 * BC * NT_SOR + tid;
        h.hE = h.pubR ? H + 2 * BC * NT_SOR + ds.nE : H + (2 * BC + BR) * NT_SOR + tid;
        if (!owner) { h.pubT = h.pubB = h.pubL = h.pubR = false; }
        Coef k;
        k.ca = 0.3125; k.cb = 0.3125; k.cc = -1.25; k.cu = 0; k.cv = 0; k.beta = 1.25; k.tol = 5e-6;
        const unsigned long long tolbits = (unsigned long long)__double_as_longlong(k.tol);
        double P[BR][BC];
#pragma unroll
        for (int li = 0; li < BR; ++li)
#pragma unroll
            for (int lj = 0; lj < BC; ++lj) P[li][lj] = 0.01 * (li + lj) + tid;
        unsigned acc = 0;
        named_sync(BAR_SOR, NT_SOR);
        const long long t0 = clock64();
        if (w < nwarps || (w >= 4 && w - 4 < nwarps - 4)) {
            for (int T = 0; T < stages; ++T) {
                unsigned mhi = 0u;
                bool v = false;
                const bool top = !((T + (PMAP ? (w >> 1) : (w >> 2))) & 1);
                if (TM) {
                    if (top) block_sweep_tm<BR, BC, RS, 0, RS>(P, tm_mine, h, k, owner, true, mhi);
                    else block_sweep_tm<BR, BC, RS, RS, BR>(P, tm_mine, h, k, owner, true, mhi);
                } else if (PMAP >= 10) {     // explicit C' prefetch PMAP - 10 diagonals ahead
                    if (top) block_sweep_pf<BR, BC, RS, 0, RS, 1, PMAP - 10>(P, Cs + tid, h, k, tolbits, mhi, v);
                    else block_sweep_pf<BR, BC, RS, RS, BR, 1, PMAP - 10>(P, Cs + tid, h, k, tolbits, mhi, v);
                } else if (PMAP == 6) {      // hybrid: top sub-block's C' from shared memory, bottom sub-block's from Tensor Memory
                    if (top) block_sweep<BR, BC, RS, 0, RS, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                    else block_sweep_tm<BR, BC, RS, RS, BR>(P, tm_mine, h, k, owner, true, mhi);
                } else if (PMAP == 5) {      // three sub-blocks of 3 rows: even stages sweep sub-blocks 0 and 2, odd stages sub-block 1
                    if (top) {
                        block_sweep<BR, BC, -3, 0, 3, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                        block_sweep<BR, BC, -3, 6, 9, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                    } else {
                        block_sweep<BR, BC, -3, 3, 6, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                    }
                } else if (PMAP == 2) {      // same sub-block every stage, two distinct copies of the code (TRACK 1 / TRACK 2)
                    if (top) block_sweep<BR, BC, RS, 0, RS, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                    else block_sweep<BR, BC, RS, 0, RS, 2>(P, Cs + tid, h, k, tolbits, mhi, v);
                } else if (PMAP == 3) {      // same sub-block, one copy of the code
                    block_sweep<BR, BC, RS, 0, RS, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                } else if (PMAP == 4) {      // same sub-block, one copy of the code (TRACK 2)
                    block_sweep<BR, BC, RS, 0, RS, 2>(P, Cs + tid, h, k, tolbits, mhi, v);
                } else {
                    if (top) block_sweep<BR, BC, RS, 0, RS, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                    else block_sweep<BR, BC, RS, RS, BR, 1>(P, Cs + tid, h, k, tolbits, mhi, v);
                }
                acc += mhi;
                if (BARRIER == 1) named_sync(BAR_SOR, 32 * (nwarps > 4 ? 8 : nwarps));
                if (BARRIER == 2) named_sync(8 + (w & 3), 64);
            }
        }
        const long long t1 = clock64();
        double s = acc;
#pragma unroll
        for (int li = 0; li < BR; ++li)
#pragma unroll
            for (int lj = 0; lj < BC; ++lj) s += P[li][lj];
        if (s == 1.2345) cyc[3] = (long long)s;
        if ((tid & 31) == 0) cyc[blockIdx.x * 8 + w] = t1 - t0;
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <bool TM, int BARRIER, int PMAP = 0>
static void run(const SBlock *d_desc, long long *cyc, int nwarps, const char *label) {
    using C = Cfg128;
    const size_t smem = C::H_BYTES + C::CS_BYTES;
    const int stages = 4000;
    cudaFuncSetAttribute(lock_kernel<TM, BARRIER, PMAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(cyc, 0, sizeof(long long) * 8 * 148);
    lock_kernel<TM, BARRIER, PMAP><<<148, 384, smem>>>(d_desc, cyc, stages, nwarps);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-44s warps %d: %.0f cycles/stage (%s)\n", label, nwarps, (double)mx / stages, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    StreamPlan pl;
    build_tables<Cfg128>(pl);
    SBlock *d_desc; long long *cyc;
    cudaMalloc(&d_desc, sizeof(SBlock) * pl.desc.size());
    cudaMemcpy(d_desc, pl.desc.data(), sizeof(SBlock) * pl.desc.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&cyc, sizeof(long long) * 8 * 148);
    run<false, 0>(d_desc, cyc, 8, "smem C', free running");
    run<false, 1>(d_desc, cyc, 8, "smem C', barrier per stage");
    run<false, 2>(d_desc, cyc, 8, "smem C', pair barrier per stage");
    run<true, 0>(d_desc, cyc, 8, "TMEM C', free running");
    run<true, 1>(d_desc, cyc, 8, "TMEM C', barrier per stage");
    run<true, 2>(d_desc, cyc, 8, "TMEM C', pair barrier per stage");
    run<false, 0, 1>(d_desc, cyc, 8, "smem C', free running, same kind per SMSP");
    run<false, 1, 1>(d_desc, cyc, 8, "smem C', barrier, same kind per SMSP");
    run<true, 0, 1>(d_desc, cyc, 8, "TMEM C', free running, same kind per SMSP");
    run<true, 1, 1>(d_desc, cyc, 8, "TMEM C', barrier, same kind per SMSP");
    run<false, 0, 11>(d_desc, cyc, 4, "C' prefetch 1 diagonal ahead, 1 warp/SMSP");
    run<false, 0, 12>(d_desc, cyc, 4, "C' prefetch 2 diagonals ahead, 1 warp/SMSP");
    run<false, 0, 13>(d_desc, cyc, 4, "C' prefetch 3 diagonals ahead, 1 warp/SMSP");
    run<false, 0, 11>(d_desc, cyc, 8, "C' prefetch 1 diagonal ahead, free running");
    run<false, 0, 12>(d_desc, cyc, 8, "C' prefetch 2 diagonals ahead, free running");
    run<false, 0, 13>(d_desc, cyc, 8, "C' prefetch 3 diagonals ahead, free running");
    run<false, 1, 12>(d_desc, cyc, 8, "C' prefetch 2 diagonals ahead, barrier");
    run<false, 1, 13>(d_desc, cyc, 8, "C' prefetch 3 diagonals ahead, barrier");
    run<false, 0, 6>(d_desc, cyc, 8, "hybrid top smem / bottom TMEM, free running");
    run<false, 1, 6>(d_desc, cyc, 8, "hybrid top smem / bottom TMEM, barrier");
    run<false, 0, 6>(d_desc, cyc, 4, "hybrid, free running, 1 warp/SMSP");
    run<false, 0, 5>(d_desc, cyc, 8, "three sub-blocks, free running");
    run<false, 1, 5>(d_desc, cyc, 8, "three sub-blocks, barrier per stage");
    run<false, 0, 5>(d_desc, cyc, 4, "three sub-blocks, free running, 1 warp/SMSP");
    run<false, 1, 5>(d_desc, cyc, 4, "three sub-blocks, barrier, 1 warp/SMSP");
    run<false, 0, 3>(d_desc, cyc, 4, "top only, one code copy, 1 warp/SMSP");
    run<false, 0, 4>(d_desc, cyc, 4, "top only (exact test), one copy, 1 warp/SMSP");
    run<false, 0, 2>(d_desc, cyc, 4, "top only, two code copies, 1 warp/SMSP");
    run<false, 0, 3>(d_desc, cyc, 8, "top only, one code copy, 2 warps/SMSP");
    run<false, 0, 2>(d_desc, cyc, 8, "top only, two code copies, 2 warps/SMSP");
    run<false, 0>(d_desc, cyc, 4, "smem C', free running, 1 warp/SMSP");
    run<false, 1>(d_desc, cyc, 4, "smem C', barrier, 1 warp/SMSP");
    run<true, 0>(d_desc, cyc, 4, "TMEM C', free running, 1 warp/SMSP");
    run<true, 1>(d_desc, cyc, 4, "TMEM C', barrier, 1 warp/SMSP");
    return 0;
}
