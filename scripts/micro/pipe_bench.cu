// pipe_bench.cu -- the SOR pipeline of sor_pipe.cuh alone on the SMs: 8 warps, 7+6*7 = 49 levels, rows read from a
// pre-filled entry ring (always ready), exit rows written to global memory.  Measures cycles per step (one step =
// one row through all 49 levels) and checks the result against a sequential lexicographic SOR on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "sor_pipe.cuh"

using namespace nns::pipe;

constexpr int NT = 384, REGS_SOR = 208, REGS_ST = 88;     // setmaxnreg works per warpgroup: warps 0-7 (7 SOR + 1 spare) / warps 8-11

struct Args {
    const double *p_in, *c_in;     // [members][NX][128] of this CTA: blockIdx.x * members
    double *p_out;
    long long *cyc;
    int nx, members, cap;
    double ca, cb, om;
};

template <int K>
__device__ void sor_role(const Args &a, Ctl *ctl, unsigned char *rings, uint32_t tmem, int w, int lane, int off, int nsteps) {
    WarpState<K> st;
#pragma unroll
    for (int s = 0; s <= K; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) st.B[s][h][q] = 0.0;
#pragma unroll
    for (int s = 0; s < K; ++s) st.amax[s] = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) st.cprev[q] = 0.0;
    st.flags = 0;
    st.F = 0;
    WarpCtx c;
    c.k.a = a.ca; c.k.b = a.cb; c.k.om = a.om; c.k.thi = 0x3ed0c6f7u;
    c.nx = a.nx; c.lane = lane;
    init_lane_coef(c);
    c.tm = tmem + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(256 * (w >> 2));
    prologue<K>(st, c);
    const bool FIRST = w == 0, LAST = w == NW - 1;
    const uint32_t ring0 = s_u32(rings), iw = ring0 + R0S * SLOTB;
    c.in_ring = (FIRST ? ring0 : iw + (w - 1) * RSLOT * SLOTB) + lane * 16;
    c.out_ring = iw + w * RSLOT * SLOTB + lane * 16;
    c.full_in = FIRST ? s_u32(&ctl->full0[0]) : s_u32(&ctl->full[w][0]);
    c.empty_in = FIRST ? s_u32(&ctl->empty0[0]) : s_u32(&ctl->empty[w][0]);
    c.full_out = s_u32(&ctl->full[(w + 1) % NW][0]);
    c.empty_out = s_u32(&ctl->empty[(w + 1) % NW][0]);
    c.in_slots = FIRST ? R0S : RSLOT;
    c.in_slot = 0; c.out_slot = 0; c.in_par = 0; c.out_par = 0;
    const int r0 = -off - lane;          // stream row entering the warp in step 0
    c.iin = ((r0 % a.nx) + a.nx) % a.nx;
    pin_ctx(c);
    const int total = a.members * a.nx;
    double *pout = a.p_out + (size_t)blockIdx.x * total * 128 + 4 * lane;
    auto exit_row = [&](const double *row, int n) {
        const int r = n - off - 2 * K - lane;
        if (r >= 0 && r < total) {
            double2 *q = reinterpret_cast<double2 *>(pout + (size_t)r * 128);
            q[0] = make_double2(row[0], row[1]);
            q[1] = make_double2(row[2], row[3]);
        }
    };
    const long long t0 = clock64();
    for (int n = 0; n < nsteps; n += 2) {
        step<K, 0>(st, c, LAST, n, exit_row);
        step<K, 1>(st, c, LAST, n + 1, exit_row);
    }
    if (lane == 0) a.cyc[blockIdx.x * NW + w] = clock64() - t0;
    if (st.flags == (unsigned)a.cap) a.cyc[0] = 0;     // keeps the exit-test code alive
}

__global__ void __launch_bounds__(NT, 1) pipe_kernel(const Args a) {
    extern __shared__ __align__(16) unsigned char rings[];
    __shared__ Ctl ctl;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int g = 0; g < R0S; ++g) { mbar_init(s_u32(&ctl.full0[g]), 128); mbar_init(s_u32(&ctl.empty0[g]), 32); }
        for (int i = 0; i < NW; ++i)
            for (int q = 0; q < RSLOT; ++q) { mbar_init(s_u32(&ctl.full[i][q]), 32); mbar_init(s_u32(&ctl.empty[i][q]), 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int q = tid; q < R0S * SLOTB / 8; q += NT) reinterpret_cast<double *>(rings)[q] = 0.0;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const int total = a.members * a.nx;
    const int nsteps = (total + 2 * KW * NW + 31 + 2 + 1) & ~1;
    if (w >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ST));
        // feeder: thread ts owns column ts; row r goes to the skewed slots (r + l) % R0S of SOR lane l = ts / 4
        const int ts = tid - 256, l = ts >> 2, q = ts & 3;
        const double *pin = a.p_in + (size_t)blockIdx.x * total * 128, *cin = a.c_in + (size_t)blockIdx.x * total * 128;
        double *ring = reinterpret_cast<double *>(rings);
        const int inslot = (q >> 1) * 64 + l * 2 + (q & 1);        // doubles: chunk (q >> 1) * 512 B + lane * 16 B + element
        constexpr int FB = 4;
        double bp[FB], bc[FB];
        auto load = [&](int r) {
#pragma unroll
            for (int k = 0; k < FB; ++k) {
                bp[k] = r + k < total ? pin[(size_t)(r + k) * 128 + ts] : 0.0;
                bc[k] = r + k < total ? cin[(size_t)(r + k) * 128 + ts] : 0.0;
            }
        };
        load(0);
        for (int r0 = 0; r0 < nsteps + 1; r0 += FB) {
#pragma unroll
            for (int k = 0; k < FB; ++k) {
                const int r = r0 + k;
                // slot (r + 31) % R0S is touched for the first time since its previous use (step r + 31 - R0S)
                if (r + 31 >= R0S) mbar_wait(s_u32(&ctl.empty0[(r + 31) % R0S]), (uint32_t)((r + 31) / R0S - 1) & 1u);
                const int slot = (r + l) % R0S;
                ring[slot * 256 + inslot] = bp[k];
                ring[slot * 256 + 128 + inslot] = bc[k];
                mbar_arrive(s_u32(&ctl.full0[r % R0S]));      // slot r is complete once every thread has written row r
            }
            load(r0 + FB);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOR));
        if (w < NW) sor_role<KW>(a, &ctl, rings, tmem, w, lane, 2 * KW * w, nsteps);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main(int argc, char **argv) {
    const int nx = 128, members = argc > 1 ? atoi(argv[1]) : 8, cap = 49, grid = argc > 2 ? atoi(argv[2]) : 148;
    const size_t N = (size_t)nx * 128, tot = N * members * grid;
    std::vector<double> hp(tot), hc(tot);
    srand(1);
    for (size_t i = 0; i < tot; ++i) { hp[i] = (rand() % 2001 - 1000) * 1e-3; hc[i] = (rand() % 2001 - 1000) * 1e-4; }
    for (size_t m = 0; m < (size_t)members * grid; ++m)         // C' is zero on the frozen lines (the feeder of the real kernel guarantees it)
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < 128; ++j)
                if (i == 0 || i == nx - 1 || j == 0 || j == 127) hc[m * N + (size_t)i * 128 + j] = 0.0;
    double *dp, *dc, *dq; long long *dcyc;
    cudaMalloc(&dp, tot * 8); cudaMalloc(&dc, tot * 8); cudaMalloc(&dq, tot * 8); cudaMalloc(&dcyc, grid * NW * 8);
    cudaMemcpy(dp, hp.data(), tot * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, hc.data(), tot * 8, cudaMemcpyHostToDevice);
    cudaMemset(dq, 0, tot * 8);
    const double beta = 1.25, dx = 2.0 / 127, dy = 2.0 / 127, dx2 = dx * dx, dy2 = dy * dy, den = 2 * dx2 + 2 * dy2;
    Args a{dp, dc, dq, dcyc, nx, members, cap, dy2 / den, dx2 / den, beta};
    const size_t smem = R0S * SLOTB + NW * RSLOT * SLOTB;
    cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe_kernel<<<grid, NT, smem>>>(a);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(err)); return 1; }
    cudaEventRecord(e0);
    pipe_kernel<<<grid, NT, smem>>>(a);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(grid * NW);
    cudaMemcpy(cyc.data(), dcyc, grid * NW * 8, cudaMemcpyDeviceToHost);
    const int total = members * nx;
    printf("grid %d, %d members per CTA: %.3f ms; warp cycles of CTA 0:", grid, members, ms);
    for (int w = 0; w < NW; ++w) printf(" %lld", cyc[w]);
    printf("\n  = %.1f cycles per step (%d steps); ideal FP64 issue: 14 levels x 24 x 2.2 = 740\n", (double)cyc[NW - 1] / (total + 134), total + 134);
    // host check of member 0 of CTA 0 and of the last member of the last CTA
    std::vector<double> out(tot);
    cudaMemcpy(out.data(), dq, tot * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (size_t m : {(size_t)0, (size_t)members * grid - 1}) {
        std::vector<double> p(hp.begin() + m * N, hp.begin() + (m + 1) * N);
        const double *c = &hc[m * N];
        for (int s = 0; s < cap; ++s)
            for (int i = 1; i < nx - 1; ++i)
                for (int j = 1; j < 127; ++j) {
                    const size_t q = (size_t)i * 128 + j;
                    const double t = fma(a.ca, p[q - 128] + p[q + 128], fma(a.cb, p[q + 1], -p[q] - c[q]));
                    p[q] = fma(a.om, fma(a.cb, p[q - 1], t), p[q]);
                }
        double e = 0, nn = 0;
        for (size_t q = 0; q < N; ++q) { e += (p[q] - out[m * N + q]) * (p[q] - out[m * N + q]); nn += p[q] * p[q]; }
        printf("  member %zu: rel L2 error vs sequential SOR %.3e\n", m, sqrt(e / nn));
        worst = fmax(worst, sqrt(e / nn));
    }
    printf(worst < 1e-12 ? "PASS\n" : "FAIL\n");
    return worst < 1e-12 ? 0 : 1;
}
