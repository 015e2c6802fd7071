// tma_ring.cu -- cost of streaming rows global -> shared with cp.async.bulk (TMA unit, 1-D bulk copies),
// as chorin_fd_stream.cu's stencil role does: one producer lane, 4 consumer warps, NG groups in flight,
// F copies of CB bytes per group.  Prints cycles per group and GB/s per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const double *src, size_t per_cta, int ngroups, int NG, int F, int CB, double *out, long long *cyc) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[8], empty[8];
    const int t = threadIdx.x;
    if (t == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); } asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const char *base = reinterpret_cast<const char *>(src) + (size_t)blockIdx.x * per_cta;
    const int gbytes = F * CB;
    auto issue = [&](int g) {
        const int s = g % NG;
        if (g >= NG) mbar_wait(&empty[s], ((g / NG) - 1) & 1);
        mbar_expect(&full[s], gbytes);
        for (int f = 0; f < F; ++f) bulk(smem + (size_t)s * gbytes + (size_t)f * CB, base + (size_t)g * gbytes + (size_t)f * CB, CB, &full[s]);
    };
    if (t == 0) for (int g = 0; g < NG - 1 && g < ngroups; ++g) issue(g);
    double acc = 0;
    long long c_issue = 0, c_wait = 0, c_use = 0;
    const long long t0 = clock64();
    for (int g = 0; g < ngroups; ++g) {
        long long a0 = clock64();
        if (t == 0 && g + NG - 1 < ngroups) issue(g + NG - 1);
        long long a1 = clock64();
        mbar_wait(&full[g % NG], (g / NG) & 1);
        long long a2 = clock64();
        const double *row = reinterpret_cast<const double *>(smem + (size_t)(g % NG) * gbytes);
        for (int q = t; q < gbytes / 8; q += 128) acc += row[q];
        __syncwarp();
        if ((t & 31) == 0) mbar_arrive(&empty[g % NG]);
        long long a3 = clock64();
        c_issue += a1 - a0; c_wait += a2 - a1; c_use += a3 - a2;
    }
    const long long t1 = clock64();
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0 && (t == 0 || t == 64)) printf("  thread %d: issue %lld wait_full %lld use+arrive %lld cycles/group\n", t, c_issue / ngroups, c_wait / ngroups, c_use / ngroups);
    if (acc == 1.2345) out[0] = acc;
}

int main() {
    const size_t per_cta = 8u << 20;     // 8 MiB per CTA, 148 CTAs: 1.2 GB streamed (HBM)
    double *src, *out; long long *cyc;
    cudaMalloc(&src, per_cta * 148); cudaMemset(src, 0, per_cta * 148);
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 148);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int NG, F, CB; } cfgs[] = {{4, 8, 1024}, {4, 4, 2048}, {4, 2, 4096}, {4, 1, 8192}, {2, 4, 4096}, {8, 8, 1024}, {8, 4, 2048},
                                            {4, 4, 4096}};
    for (auto c : cfgs) {
        const int gbytes = c.F * c.CB, ngroups = (int)(per_cta / gbytes) / 4;
        k<<<148, 128, (size_t)c.NG * gbytes>>>(src, per_cta, ngroups, c.NG, c.F, c.CB, out, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("NG %d groups in ring, %d copies x %5d B per group: %7.0f cycles/group, %6.1f B/clk/SM (%s)\n", c.NG, c.F, c.CB,
               mx / ngroups, (double)gbytes * ngroups / mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
