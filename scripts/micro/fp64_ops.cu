// fp64_ops.cu -- DFMA issue rate of ONE warp per SM sub-partition as a function of operand pattern
// (distinct register operands vs reused ones) and chain structure.  Cycles per DFMA (warp-level).
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(128, 1) k(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); z[i] = 1e-9 * i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);                 // 1 register operand + 2 uniform
                if (MODE == 1) x[i] = fma(y[i], a, x[i]);              // 2 distinct registers
                if (MODE == 2) x[i] = fma(y[i], z[i], x[i]);           // 3 distinct registers
                if (MODE == 3) x[i] = fma(y[(i + r) % ILP], z[(i + 2 * r + 1) % ILP], x[i]);   // 3 distinct, shuffled
                if (MODE == 4) x[i] = fma(a, x[(i + 1) % ILP], x[i]);  // neighbour-coupled like the stencil
            }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    if (s == 1.2345) out[0] = s;
}

template <int MODE, int ILP>
static void run(double *out, long long *cyc, int warps) {
    const int iters = 4096;
    k<MODE, ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d ILP %2d warps/SM %d: %.2f cycles per warp-DFMA (per sub-partition)\n", MODE, ILP, warps,
           (double)c / (iters * 4.0 * ILP) / (warps > 4 ? 1.0 : 1.0));
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 148);
    run<0, 8>(out, cyc, 4); run<1, 8>(out, cyc, 4); run<2, 8>(out, cyc, 4); run<3, 8>(out, cyc, 4); run<4, 8>(out, cyc, 4);
    run<0, 16>(out, cyc, 4); run<2, 16>(out, cyc, 4); run<3, 16>(out, cyc, 4); run<4, 16>(out, cyc, 4);
    run<2, 32>(out, cyc, 4); run<3, 32>(out, cyc, 4);
    run<2, 8>(out, cyc, 1); run<3, 16>(out, cyc, 1);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
