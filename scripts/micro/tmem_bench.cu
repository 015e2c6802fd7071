// tmem_bench.cu -- can Tensor Memory serve as a thread-private scratchpad for the SOR right-hand side?
// Measures tcgen05.ld (32x32b: every thread reads N consecutive 32-bit columns of its own TMEM lane) latency and
// throughput per SM for 1..8 warps, next to a round-trip correctness check of tcgen05.st -> tcgen05.ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_bench tmem_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: x4 load + wait, dependent (latency); 1: 8 x4 loads in flight per wait; 2: 2 x16 loads per wait
__global__ void __launch_bounds__(256, 1) bench(long long *cyc, unsigned *errs, int iters, int nwarps, int mode) {
    extern __shared__ unsigned char pad[];
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = s_base;
    // thread-private window: lane quarter of the warp, 256 columns per warp of a pair
    const uint32_t mine = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(256 * (warp >> 2));
    // fill with a pattern and check the round trip
    for (int c = 0; c < 256; c += 4) {
        uint32_t v[4] = {(uint32_t)(tid * 1000 + c), (uint32_t)(tid * 1000 + c + 1), (uint32_t)(tid * 1000 + c + 2), (uint32_t)(tid * 1000 + c + 3)};
        tm_st4(mine + c, v);
    }
    tm_wait_st();
    unsigned bad = 0;
    for (int c = 0; c < 256; c += 4) {
        uint32_t v[4];
        tm_ld4(mine + c, v);
        tm_wait_ld();
        for (int k = 0; k < 4; ++k) bad += v[k] != (uint32_t)(tid * 1000 + c + k);
    }
    if (bad) atomicAdd(errs, bad);
    __syncthreads();
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        t0 = clock64();
        if (mode == 0) {
            uint32_t col = 0;
            for (int it = 0; it < iters; ++it) {
                uint32_t v[4];
                tm_ld4(mine + (col & 252), v);
                tm_wait_ld();
                col = v[0] & 4;          // dependent address
                acc += v[1];
            }
        } else if (mode == 1) {
            for (int it = 0; it < iters; ++it) {
                uint32_t v[8][4];
#pragma unroll
                for (int k = 0; k < 8; ++k) tm_ld4(mine + ((it * 32 + k * 4) & 252), v[k]);
                tm_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k][0] ^ v[k][1] ^ v[k][2] ^ v[k][3];
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                uint32_t v[2][16];
#pragma unroll
                for (int k = 0; k < 2; ++k) tm_ld16(mine + ((it * 32 + k * 16) & 240), v[k]);
                tm_wait_ld();
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int q = 0; q < 16; ++q) acc += v[k][q];
            }
        }
        t1 = clock64();
    }
    if (acc == 0x12345678u) errs[1] = acc;
    if ((tid & 31) == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}

int main() {
    long long *cyc; unsigned *errs;
    cudaMalloc(&cyc, sizeof(long long) * 8 * 148);
    cudaMalloc(&errs, 8);
    cudaMemset(errs, 0, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4000;
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {1, 4, 8}) {
            cudaMemset(cyc, 0, sizeof(long long) * 8 * 148);
            bench<<<148, 256, 200 * 1024>>>(cyc, errs, iters, nw, mode);
            cudaDeviceSynchronize();
            long long h[8];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes_per_iter = mode == 0 ? 512.0 : 4096.0;     // per warp
            printf("mode %d (%s) warps/SM %d: %.1f cycles per iteration, %.1f B/clk/SM\n", mode,
                   mode == 0 ? "x4 dependent" : mode == 1 ? "8 x4 per wait" : "2 x16 per wait", nw, (double)mx / iters,
                   bytes_per_iter * nw * iters / (double)mx);
        }
    unsigned he[2];
    cudaMemcpy(he, errs, 8, cudaMemcpyDeviceToHost);
    printf("round-trip mismatches: %u; last error: %s\n", he[0], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
