// coissue.cu -- do integer / shared-memory instructions issue in the shadow of DFMA on B200?  One or two warps per
// SM sub-partition run G groups of 6 independent DFMAs, each group followed by K independent integer instructions
// (or K LDS.64).  If cycles per group = 6 * 2.3 + K the dispatch port is blocked while a DFMA issues.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o coissue coissue.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int K, int KIND>     // KIND 0: LOP3/IADD mix, 1: LDS.64
__global__ void __launch_bounds__(256, 1) kern(double *out, long long *cyc, int iters, double a, double b, unsigned m) {
    __shared__ double sm[2048];
    for (int q = threadIdx.x; q < 2048; q += blockDim.x) sm[q] = q;
    __syncthreads();
    double x[12];
    unsigned u[12];
    double ls = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = threadIdx.x + i; u[i] = threadIdx.x * 7 + i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
#pragma unroll
            for (int i = 0; i < 6; ++i) x[(g * 6 + i) % 12] = fma(x[(g * 6 + i) % 12], a, b);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (KIND == 0) u[(g + i) % 12] = (u[(g + i) % 12] ^ m) + (unsigned)i;
                else ls += sm[(threadIdx.x + 32 * ((g * K + i) & 31) + it) & 2047];
            }
        }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
    double s = ls;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += x[i] + u[i];
    if (s == 1.2345) out[0] = s;
}

template <int K, int KIND>
static void run(double *out, long long *cyc, int warps) {
    const int iters = 2000;
    kern<K, KIND><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9, 0x5a5a5a5au);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%s K=%2d warps/SM %d: %.1f cycles per group of 6 DFMA + K per warp, %.1f per SM sub-partition-group\n", KIND ? "LDS.64" : "int   ", K,
           warps, (double)mx / (iters * 8.0), (double)mx / (iters * 8.0) / ((warps + 3) / 4));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 8 * 148);
    for (int w : {4, 8}) {
        run<0, 0>(out, cyc, w); run<3, 0>(out, cyc, w); run<6, 0>(out, cyc, w); run<12, 0>(out, cyc, w); run<24, 0>(out, cyc, w);
        run<3, 1>(out, cyc, w); run<6, 1>(out, cyc, w);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
