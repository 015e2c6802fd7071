// wave_bench.cu -- isolates the SOR role of chorin_fd_stream.cu: the sub-block wavefront (all 49 sweeps of a
// 128x128 member, one named barrier per super-stage) with the real block tables, but without the stencil role and
// without global memory traffic.  Variants: 256 threads / 255 registers, or 384 threads with setmaxnreg like the
// real kernel; zero or non-zero data; short or long runs (power management).  Reports cycles per super-stage and
// per sub-block sweep.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o wave_bench wave_bench.cu
#include "../../neural-navier-stokes_b200/csrc/chorin_fd_stream.cu"

namespace nns { void set_error(const char *, ...) {} }
using namespace nns;

template <int NTH, bool SETREG>
__global__ void __launch_bounds__(NTH, 1) wave_kernel(const SBlock *desc, long long *cyc, int members, double scale) {
    using C = Cfg128;
    constexpr int BR = C::BRc, BC = C::BCc, RS = C::RSc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *Cs = reinterpret_cast<double2 *>(smem_raw);
    double *H = reinterpret_cast<double *>(smem_raw + C::CS_BYTES);
    const int tid = threadIdx.x;
    for (int q = tid; q < C::NCH * NT_SOR; q += NTH) Cs[q] = make_double2(scale * 1e-3 * (q % 97), scale * 2e-3 * (q % 89));
    for (int q = tid; q < C::NSLOT * NT_SOR; q += NTH) H[q] = scale * 1e-4 * (q % 101);
    __syncthreads();
    if (tid >= NT_SOR) {
        if (SETREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
        return;
    }
    if (SETREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR));
    SBlock ds = desc[tid];
    const bool owner = ds.r0 > 0;
    SHalo<BR, BC> h;
    h.Hme = H + tid;
    h.pubT = ds.nN >= 0; h.pubB = ds.nS >= 0; h.pubL = ds.nW >= 0; h.pubR = ds.nE >= 0;
    h.hN = h.pubT ? H + BC * NT_SOR + ds.nN : H + tid;
    h.hS = h.pubB ? H + ds.nS : H + BC * NT_SOR + tid;
    h.hW = h.pubL ? H + (2 * BC + BR) * NT_SOR + ds.nW : H + 2 * BC * NT_SOR + tid;
    h.hE = h.pubR ? H + 2 * BC * NT_SOR + ds.nE : H + (2 * BC + BR) * NT_SOR + tid;
    Coef k;
    k.ca = 0.3125; k.cb = 0.3125; k.cc = -1.25; k.cu = 0; k.cv = 0; k.beta = 1.25; k.tol = 5e-6;
    const unsigned long long tolbits = (unsigned long long)__double_as_longlong(k.tol);
    const int cap = 49, tmax = 2 * C::NBRc + C::NBCc - 2 + 2 * (cap - 1);
    double P[BR][BC];
    long long prof[2] = {0, 0};
    unsigned long long mask = 0, amb = 0;
    const long long t0 = clock64();
    for (int m = 0; m < members; ++m) {
#pragma unroll
        for (int li = 0; li < BR; ++li)
#pragma unroll
            for (int lj = 0; lj < BC; ++lj) P[li][lj] = scale * (0.01 * (li + lj) + tid + m);
        wavefront<BR, BC, RS, 1>(P, Cs + tid, h, owner, ds.bd, tmax, cap, k, tolbits, mask, amb, tid == 0 ? prof : nullptr);
    }
    const long long t1 = clock64();
    double s = (double)(mask + amb);
#pragma unroll
    for (int li = 0; li < BR; ++li)
#pragma unroll
        for (int lj = 0; lj < BC; ++lj) s += P[li][lj];
    if (s == 1.2345) cyc[3] = (long long)s;
    if (tid == 0) {
        cyc[blockIdx.x * 4 + 0] = t1 - t0;
        cyc[blockIdx.x * 4 + 1] = prof[0];
        cyc[blockIdx.x * 4 + 2] = prof[1];
    }
}

template <int NTH, bool SETREG>
static void run(const SBlock *d_desc, long long *cyc, int members, double scale, const char *label) {
    using C = Cfg128;
    const size_t smem = C::CS_BYTES + C::H_BYTES;
    cudaFuncSetAttribute(wave_kernel<NTH, SETREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(cyc, 0, sizeof(long long) * 4 * 148);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    wave_kernel<NTH, SETREG><<<148, NTH, smem>>>(d_desc, cyc, members, scale);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long h[4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double stages = 141.0 * members;
    printf("%-28s members %4d scale %g: %.3f ms, %.0f cycles/stage, %.0f cycles per sub-block sweep of thread 0 (%lld sweeps), %.3f GHz\n",
           label, members, scale, ms, h[0] / stages, h[2] ? (double)h[1] / h[2] : 0.0, h[2], h[0] / (ms * 1e6));
}

// the same wavefront with the right-hand side in Tensor Memory and warp-uniform control flow (block_sweep_tm)
__global__ void __launch_bounds__(384, 1) wave_kernel_tm(const SBlock *desc, long long *cyc, int members, double scale) {
    using C = Cfg128;
    constexpr int BR = C::BRc, BC = C::BCc, RS = C::RSc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    double *H = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    for (int q = tid; q < C::NSLOT * NT_SOR; q += 384) H[q] = scale * 1e-4 * (q % 101);
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
        for (int c = 0; c < 32; c += 4) {
            double2 v[4];
            for (int e = 0; e < 4; ++e) v[e] = make_double2(scale * 1e-3 * ((c + e) * 7 + tid % 13), scale * 2e-3 * ((c + e) * 5 + tid % 11));
            tm_st16(tm_mine + 4 * c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
        const unsigned tolhi = (unsigned)((unsigned long long)__double_as_longlong(k.tol) >> 32);
        const int cap = 49, tmax = 2 * C::NBRc + C::NBCc - 2 + 2 * (cap - 1);
        double P[BR][BC];
        unsigned long long mask = 0, amb = 0;
        long long tsw = 0, nsw = 0;
        const long long t0 = clock64();
        for (int m = 0; m < members; ++m) {
#pragma unroll
            for (int li = 0; li < BR; ++li)
#pragma unroll
                for (int lj = 0; lj < BC; ++lj) P[li][lj] = scale * (0.01 * (li + lj) + tid + m);
            named_sync(BAR_SOR, NT_SOR);
            for (int T = 0; T <= tmax; ++T) {
                const int q = T - ds.bd;
                const bool act = owner && q >= 0 && q <= 2 * cap - 1;
                if (__any_sync(0xffffffffu, act)) {
                    const long long ts0 = clock64();
                    unsigned mhi = 0u;
                    if (!((T - (w >> 2)) & 1)) block_sweep_tm<BR, BC, RS, 0, RS>(P, tm_mine, h, k, act, (q >> 1) != cap - 1, mhi);
                    else block_sweep_tm<BR, BC, RS, RS, BR>(P, tm_mine, h, k, act, (q >> 1) != cap - 1, mhi);
                    if (act) { mask |= (unsigned long long)(mhi > tolhi) << (q >> 1); amb |= (unsigned long long)(mhi == tolhi) << (q >> 1); }
                    if (tid == 0) { tsw += clock64() - ts0; nsw += 1; }
                }
                named_sync(BAR_SOR, NT_SOR);
            }
        }
        const long long t1 = clock64();
        double s = (double)(mask + amb);
#pragma unroll
        for (int li = 0; li < BR; ++li)
#pragma unroll
            for (int lj = 0; lj < BC; ++lj) s += P[li][lj];
        if (s == 1.2345) cyc[3] = (long long)s;
        if (tid == 0) { cyc[blockIdx.x * 4 + 0] = t1 - t0; cyc[blockIdx.x * 4 + 1] = tsw; cyc[blockIdx.x * 4 + 2] = nsw; }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static void run_tm(const SBlock *d_desc, long long *cyc, int members, double scale, const char *label) {
    using C = Cfg128;
    const size_t smem = C::H_BYTES + 100 * 1024;
    cudaFuncSetAttribute(wave_kernel_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(cyc, 0, sizeof(long long) * 4 * 148);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    wave_kernel_tm<<<148, 384, smem>>>(d_desc, cyc, members, scale);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long h[4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double stages = 141.0 * members;
    printf("%-28s members %4d scale %g: %.3f ms, %.0f cycles/stage, %.0f cycles per sub-block sweep of thread 0 (%lld sweeps), %.3f GHz\n",
           label, members, scale, ms, h[0] / stages, h[2] ? (double)h[1] / h[2] : 0.0, h[2], h[0] / (ms * 1e6));
}

int main() {
    StreamPlan pl;
    build_tables<Cfg128>(pl);
    SBlock *d_desc; long long *cyc;
    cudaMalloc(&d_desc, sizeof(SBlock) * pl.desc.size());
    cudaMemcpy(d_desc, pl.desc.data(), sizeof(SBlock) * pl.desc.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&cyc, sizeof(long long) * 4 * 148);
    run_tm(d_desc, cyc, 28, 1.0, "TMEM sweep, 384thr/setmaxnreg");
    run_tm(d_desc, cyc, 28, 1.0, "TMEM sweep, 384thr/setmaxnreg");
    for (int rep = 0; rep < 1; ++rep) {
        run<256, false>(d_desc, cyc, 28, 1.0, "256thr/255regs");
        run<384, true>(d_desc, cyc, 28, 1.0, "384thr/setmaxnreg");
        run<384, false>(d_desc, cyc, 28, 1.0, "384thr/168regs");
        run<256, false>(d_desc, cyc, 28, 0.0, "256thr/255regs zero data");
        run<384, true>(d_desc, cyc, 28, 0.0, "384thr/setmaxnreg zero data");
        run<256, false>(d_desc, cyc, 28 * 20, 1.0, "256thr/255regs long");
        run<384, true>(d_desc, cyc, 28 * 20, 1.0, "384thr/setmaxnreg long");
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
