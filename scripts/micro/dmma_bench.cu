// dmma_bench.cu -- FP64 contraction on B200: legacy tensor path (mma.sync.aligned.m8n8k4.f64, SASS DMMA) against the
// FP64 FMA pipe (DFMA).  tcgen05 has no f64 kind, so DMMA is the only tensor-core route for the chorin_spectral
// products (src/chorin_spectral/simulate.py:264-298, 367-380).  Issue-rate kernels (independent accumulators, operands in
// registers) give the pipe peaks; run under ncu for sm__inst_executed_pipe_fp64 / tensor-pipe counters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// D(8x8) += A(8x4) * B(4x8): per lane one element of A and B, two of C/D
template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 4, iters = 20000;
    double *out; cudaMalloc(&out, sizeof(double) * grid * 256);
    constexpr int ILP = 8;
    const float t1 = time_ms([&] { dfma_kernel<ILP><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    const float t2 = time_ms([&] { dmma_kernel<ILP><<<grid, 256>>>(out, iters, 0.999, 1e-3); });
    const double warps = (double)grid * 8;
    const double f1 = warps * iters * ILP * 64.0 / (t1 * 1e-3) / 1e12;        // 32 lanes x 2 flops
    const double f2 = warps * iters * ILP * 512.0 / (t2 * 1e-3) / 1e12;       // 8 x 8 x 4 x 2 flops
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    printf("DFMA  (FP64 FMA pipe)             : %7.2f TFLOP/s  (%.3f ms)\n", f1, t1);
    printf("DMMA  (mma.sync m8n8k4 f64)       : %7.2f TFLOP/s  (%.3f ms)\n", f2, t2);
    printf("DMMA / DFMA = %.2f\n", f2 / f1);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
