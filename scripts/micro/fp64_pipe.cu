// fp64_pipe.cu -- B200 microbenchmarks behind DESIGN.md's FP64 ceiling: DFMA issue rate per SM as a
// function of resident warps and independent chains, dependent-chain latency, and shared-memory
// LDS.64 / LDS.128 bandwidth.  Build: nvcc -arch=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void dfma_chain(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}

__global__ void lat_chain(double *out, long long *cyc, int iters, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x = fma(x, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}

__global__ void lat_add_chain(double *out, long long *cyc, int iters, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x = x + b;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (x == 123.456) out[0] = x;
}

template <int VEC>   // 1: LDS.64, 2: LDS.128
__global__ void lds_bw(double *out, int iters) {
    extern __shared__ double sm[];
    for (int q = threadIdx.x; q < 4096; q += blockDim.x) sm[q] = q;
    __syncthreads();
    double s = 0;
    const int base = (threadIdx.x * VEC) & 4095;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int o = (base + r * 512 + it) & (4095 & ~(VEC - 1));
            if (VEC == 1) s += sm[o];
            else { double2 v = *reinterpret_cast<double2 *>(sm + o); s += v.x + v.y; }
        }
    }
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, clock attr %d kHz\n", pr.name, sms, clk_khz);
    double *out; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    const int iters = 4096;
    // DFMA throughput: 1 CTA per SM, vary warps and ILP
    for (int warps : {4, 8, 12, 16, 32}) {
        float m1 = timeit([&] { dfma_chain<1><<<sms, warps * 32, 0>>>(out, iters, 1.0000001, 1e-9); });
        float m2 = timeit([&] { dfma_chain<2><<<sms, warps * 32, 0>>>(out, iters, 1.0000001, 1e-9); });
        float m4 = timeit([&] { dfma_chain<4><<<sms, warps * 32, 0>>>(out, iters, 1.0000001, 1e-9); });
        float m8 = timeit([&] { dfma_chain<8><<<sms, warps * 32, 0>>>(out, iters, 1.0000001, 1e-9); });
        auto rate = [&](float ms, int ilp) { return (double)sms * warps * 32 * iters * 8.0 * ilp / (ms * 1e-3) / 1e12; };
        printf("warps/SM %2d: DFMA Tinstr/s ILP1 %.2f ILP2 %.2f ILP4 %.2f ILP8 %.2f  (x2 = TFLOP/s)\n", warps,
               rate(m1, 1), rate(m2, 2), rate(m4, 4), rate(m8, 8));
    }
    long long c;
    lat_chain<<<1, 32>>>(out, cyc, 1024, 1.0000001, 1e-9);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.2f cycles\n", (double)c / (1024 * 16));
    lat_add_chain<<<1, 32>>>(out, cyc, 1024, 1e-9);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DADD dependent latency: %.2f cycles\n", (double)c / (1024 * 16));
    for (int warps : {4, 8, 16, 32}) {
        float a = timeit([&] { lds_bw<1><<<sms, warps * 32, 32768>>>(out, iters); });
        float b = timeit([&] { lds_bw<2><<<sms, warps * 32, 32768>>>(out, iters); });
        printf("warps/SM %2d: LDS.64 %.1f GB/s/SM  LDS.128 %.1f GB/s/SM\n", warps,
               (double)warps * 32 * iters * 8 * 8.0 / (a * 1e-3) / 1e9, (double)warps * 32 * iters * 8 * 16.0 / (b * 1e-3) / 1e9);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
