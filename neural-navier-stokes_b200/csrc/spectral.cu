// spectral.cu -- Chebyshev pseudo-spectral Chorin step (src/chorin_spectral/simulate.py of the
// reference): every derivative / Helmholtz / Uzawa operator is a dense (N-2)x(N-2) fp64 matrix
// product, 28 of them per step (the reference issues 32; 4 of its second-derivative products
// are never used, :271,:274).
//
// The operators are built on the HOST exactly as the reference builds them (numpy loops +
// LAPACK eig/inv, one-time, :59-199) and uploaded verbatim, so only the GEMM summation order
// differs from the reference.  The device side is a table-driven batched fp64 GEMM kernel
// (all products of one dependency level in ONE launch) plus a few fused pointwise kernels.
//
// Tensor cores: tcgen05 has no f64 kind; the legacy tensor path mma.sync.aligned.m8n8k4.f64 (SASS DMMA) does, and on
// B200 it runs on the tensor pipe at 37.1 TFLOP/s against 33.9 TFLOP/s of the FP64 FMA pipe (measured:
// scripts/micro/dmma_bench.cu, profiles/r2_micro_dmma_bench.txt -- ncu: sm__pipe_tensor_cycles_active 99.99 %,
// sm__pipe_fp64 idle).  The products are dense contractions, so they run on DMMA: 8 FMAs per lane and instruction
// from two operand registers, i.e. a quarter of the shared-memory operand traffic of a 4 x 4 register-tiled DFMA
// kernel, and the FP64 pipe stays free for the pointwise kernels.
#include "nns_common.cuh"

namespace nns {

// C[b] (m x n, ldc) = op(A)[b] (m x k) * op(B)[b] (k x n), optional epilogue
//   transB: B is stored n x k (row-major) and used as B^T
//   epi 1 : C /= (e0 + e1 * lx[i] + e2 * ly[j])
struct GemmDesc {
    const double *A, *B;
    double *C;
    long long sA, sB, sC;     // batch strides (0 = shared operand)
    int lda, ldb, ldc;
    int m, n, k;
    int transB;
    int epi;
    const double *lx, *ly;
    double e0, e1, e2;
};

#define NNS_MAX_GEMM 12
struct GemmTable {
    GemmDesc g[NNS_MAX_GEMM];
    int n;
};


__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// TM x TN tile per CTA, 8 warps: warp (wy, wx) of a 2 x 4 grid computes (TM / 2) x (TN / 4) = WM x WN DMMA tiles of 8 x 8
// (64 x 64: 4 x 2 tiles per warp, for ensembles; 32 x 32 and 16 x 32 for a few simulations, whose 125 x 125 products would
// otherwise occupy 4 CTAs each: 0.177 -> 0.117 -> 0.111 ms/step for a single N = 127 simulation).
// m8n8k4 fragments: lane l holds A[l / 4][l % 4], B[l % 4][l / 4], C[l / 4][2 (l % 4) + {0, 1}].
// The next k-slab travels global -> registers while the current one is multiplied out of shared memory.
template <int TM, int TN, int TK>
__global__ void __launch_bounds__(256) spectral_gemm_kernel(const GemmTable tab, int batch) {
    constexpr int SA = TK + 4, SB = TN + 4;     // shared-memory strides (doubles): conflict-free fragment loads (stride = 4 mod 16)
    constexpr int WM = TM / 16, WN = TN / 32;   // DMMA tiles per warp
    constexpr int QA = TM * TK / 256, QB = TN * TK / 256;       // elements per thread and slab
    const int gi = blockIdx.z / batch, b = blockIdx.z - gi * batch;
    const GemmDesc &d = tab.g[gi];
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    if (row0 >= d.m || col0 >= d.n) return;
    __shared__ double As[TM * SA];
    __shared__ double Bs[TK * SB];
    const double *A = d.A + (long long)b * d.sA, *B = d.B + (long long)b * d.sB;
    double *C = d.C + (long long)b * d.sC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wy = warp >> 2, wx = warp & 3;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double ra[QA], rb[QB];
    // A slab: TM rows x 16 k (thread: QA consecutive k of one row); B slab: 16 k x TN cols
    auto gload = [&](int k0) {
        {
            const int r = tid / (TK / QA), kk = (tid % (TK / QA)) * QA, gr = row0 + r;
#pragma unroll
            for (int q = 0; q < QA; ++q) {
                const int gk = k0 + kk + q;
                ra[q] = (gr < d.m && gk < d.k) ? A[(long long)gr * d.lda + gk] : 0.0;
            }
        }
        if (d.transB) {       // B stored n x k: thread reads QB consecutive k of one column
            const int c = tid / (TK / QB), kk = (tid % (TK / QB)) * QB, gc = col0 + c;
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                const int gk = k0 + kk + q;
                rb[q] = (gc < d.n && gk < d.k) ? B[(long long)gc * d.ldb + gk] : 0.0;
            }
        } else {              // B stored k x n: thread reads QB consecutive columns of one k
            const int kk = tid / (TN / QB), c = (tid % (TN / QB)) * QB, gk = k0 + kk;
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                const int gc = col0 + c + q;
                rb[q] = (gc < d.n && gk < d.k) ? B[(long long)gk * d.ldb + gc] : 0.0;
            }
        }
    };
    auto sstore = [&]() {
        {
            const int r = tid / (TK / QA), kk = (tid % (TK / QA)) * QA;
#pragma unroll
            for (int q = 0; q < QA; ++q) As[r * SA + kk + q] = ra[q];
        }
        if (d.transB) {
            const int c = tid / (TK / QB), kk = (tid % (TK / QB)) * QB;
#pragma unroll
            for (int q = 0; q < QB; ++q) Bs[(kk + q) * SB + c] = rb[q];
        } else {
            const int kk = tid / (TN / QB), c = (tid % (TN / QB)) * QB;
#pragma unroll
            for (int q = 0; q < QB; ++q) Bs[kk * SB + c + q] = rb[q];
        }
    };
    gload(0);
    for (int k0 = 0; k0 < d.k; k0 += TK) {
        __syncthreads();          // the previous slab has been consumed
        sstore();
        __syncthreads();
        if (k0 + TK < d.k) gload(k0 + TK);
#pragma unroll
        for (int k4 = 0; k4 < TK; k4 += 4) {
            double af[WM], bf[WN];
#pragma unroll
            for (int i = 0; i < WM; ++i) af[i] = As[(wy * (TM / 2) + i * 8 + fr) * SA + k4 + fk];
#pragma unroll
            for (int j = 0; j < WN; ++j) bf[j] = Bs[(k4 + fk) * SB + wx * (TN / 4) + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < WM; ++i)
#pragma unroll
                for (int j = 0; j < WN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gr = row0 + wy * (TM / 2) + i * 8 + fr, gc = col0 + wx * (TN / 4) + j * 8 + 2 * fk + e;
                if (gr < d.m && gc < d.n) {
                    double v = acc[i][j][e];
                    if (d.epi == 1) v = v / (d.e0 + d.e1 * d.lx[gr] + d.e2 * d.ly[gc]);
                    C[(long long)gr * d.ldc + gc] = v;
                }
            }
}

// F = 2 f - 3dt (u f_x + v f_y) + dt (u1 f1_x + v1 f1_y) + dt (f_xx + f_yy)   (chorin_spectral:277-282)
__global__ void spectral_rhs_kernel(const double *__restrict__ un, const double *__restrict__ vn,
                                    const double *__restrict__ un1, const double *__restrict__ vn1,
                                    const double *__restrict__ W, double *__restrict__ Fu, double *__restrict__ Fv,
                                    int nx, int ny, double dt, long long wstride) {
    const int n = nx - 2, m = ny - 2;
    const long long b = blockIdx.y;
    const size_t N = (size_t)nx * ny, NI = (size_t)n * m;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < NI; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q / m), j = (int)(q - (size_t)i * m);
        const size_t g = b * N + (size_t)(i + 1) * ny + (j + 1);
        const double u = un[g], v = vn[g], u1 = un1[g], v1 = vn1[g];
        const double *w = W + b * NI + q;       // 12 derivative planes, plane stride wstride
        const double un_dx = w[0], un_dy = w[wstride], un1_dx = w[2 * wstride], un1_dy = w[3 * wstride];
        const double vn_dx = w[4 * wstride], vn_dy = w[5 * wstride], vn1_dx = w[6 * wstride], vn1_dy = w[7 * wstride];
        const double un_ddx = w[8 * wstride], un_ddy = w[9 * wstride], vn_ddx = w[10 * wstride], vn_ddy = w[11 * wstride];
        Fu[b * NI + q] = 2.0 * u - 3.0 * dt * (u * un_dx + v * un_dy) + dt * (u1 * un1_dx + v1 * un1_dy) +
                         dt * (un_ddx + un_ddy);
        Fv[b * NI + q] = 2.0 * v - 3.0 * dt * (u * vn_dx + v * vn_dy) + dt * (u1 * vn1_dx + v1 * vn1_dy) +
                         dt * (vn_ddx + vn_ddy);
    }
}

// Boundary rows / columns of the predictor from the interior solution (chorin_spectral:249-254,
// 322-334), corners zero.  bvec = [b0_x(n) | bN_x(n) | b0_y(m) | bN_y(m)], sc = [1/e_x, const_x0, 1/e_y, const_y0].
__global__ void __launch_bounds__(256) spectral_boundary_kernel(double *__restrict__ A0, const double *__restrict__ bvec0,
                                                                const double *__restrict__ sc0, double *__restrict__ A1,
                                                                const double *__restrict__ bvec1, const double *__restrict__ sc1,
                                                                int nx, int ny) {
    // blockDim = (32, 8).  blockIdx.x < ceil(m / 32): 32 columns of the x0 / xN rows, the reduction over i split over the 8
    // thread rows (coalesced loads, chains of ~n / 8 terms, combined through shared memory); the other blocks: 8 rows of the
    // y0 / yN columns, one warp per row, lanes over j (coalesced), shuffle reduction.  (One thread per output with a serial
    // loop over the 125 terms took ~19 us per launch -- a sixth of a single-simulation step.)
    const int n = nx - 2, m = ny - 2;
    double *A = blockIdx.z ? A1 : A0;                    // both velocity components in one launch
    const double *bvec = blockIdx.z ? bvec1 : bvec0, *sc = blockIdx.z ? sc1 : sc0;
    double *F = A + (size_t)blockIdx.y * nx * ny;
    const double *b0x = bvec, *bNx = bvec + n, *b0y = bvec + 2 * n, *bNy = bvec + 2 * n + m;
    const int tx = threadIdx.x, ty = threadIdx.y, ncb = (m + 31) / 32;
    __shared__ double red[2][8][33];
    if ((int)blockIdx.x < ncb) {
        const int t = blockIdx.x * 32 + tx;
        double s0 = 0.0, sN = 0.0;
        if (t < m)
            for (int i = ty; i < n; i += 8) {
                const double v = F[(size_t)(i + 1) * ny + t + 1];
                s0 = fma(b0x[i], v, s0);
                sN = fma(bNx[i], v, sN);
            }
        red[0][ty][tx] = s0; red[1][ty][tx] = sN;
        __syncthreads();
        if (ty == 0 && t < m) {
            double a0 = 0.0, aN = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { a0 += red[0][k][tx]; aN += red[1][k][tx]; }
            F[t + 1] = sc[0] * a0 + sc[1];
            F[(size_t)(nx - 1) * ny + t + 1] = sc[0] * aN;
        }
        if (blockIdx.x == 0 && tx == 0 && ty == 1) {
            F[0] = 0.0; F[ny - 1] = 0.0; F[(size_t)(nx - 1) * ny] = 0.0; F[(size_t)(nx - 1) * ny + ny - 1] = 0.0;
        }
    } else {
        const int i = ((int)blockIdx.x - ncb) * 8 + ty;
        if (i < n) {
            double s0 = 0.0, sN = 0.0;
            for (int j = tx; j < m; j += 32) {
                const double v = F[(size_t)(i + 1) * ny + j + 1];
                s0 = fma(b0y[j], v, s0);
                sN = fma(bNy[j], v, sN);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); sN += __shfl_xor_sync(0xffffffffu, sN, o); }
            if (tx == 0) {
                F[(size_t)(i + 1) * ny] = sc[2] * s0 + sc[3];
                F[(size_t)(i + 1) * ny + ny - 1] = sc[2] * sN;
            }
        }
    }
}

// H = -rho/dt (S - G1 - G2)   (chorin_spectral:367)
__global__ void spectral_uzawa_rhs_kernel(const double *__restrict__ S, const double *__restrict__ G1,
                                          const double *__restrict__ G2, double *__restrict__ H, size_t NI,
                                          double rho_dt) {
    const size_t b = blockIdx.y;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < NI; q += (size_t)gridDim.x * blockDim.x)
        H[b * NI + q] = -rho_dt * (S[q] - G1[b * NI + q] - G2[b * NI + q]);
}

// u = ui - (DxDPx Q) dt/rho, v = vi - (Q DyDPy^T) dt/rho, p[interior] = Q, edges copied
// (chorin_spectral:378-381)
__global__ void spectral_project_kernel(const double *__restrict__ ui, const double *__restrict__ vi,
                                        const double *p, const double *__restrict__ G3,
                                        const double *__restrict__ G4, const double *__restrict__ Q,
                                        double *__restrict__ uo, double *__restrict__ vo, double *po,
                                        int nx, int ny, double dt, double rho) {
    const int n = nx - 2, m = ny - 2;
    const size_t b = blockIdx.y, N = (size_t)nx * ny, NI = (size_t)n * m;
    (void)n;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q / ny), j = (int)(q - (size_t)i * ny);
        double u = ui[b * N + q], v = vi[b * N + q], pp = p[b * N + q];
        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
            const size_t qi = b * NI + (size_t)(i - 1) * m + (j - 1);
            u = u - G3[qi] * dt / rho;
            v = v - G4[qi] * dt / rho;
            pp = Q[qi];
        }
        uo[b * N + q] = u;
        vo[b * N + q] = v;
        po[b * N + q] = pp;
    }
}

// trajectory snapshot (chorin_spectral:560-562) + optional non-finite count
__global__ void spectral_snapshot_kernel(const double *__restrict__ u, const double *__restrict__ v,
                                         const double *__restrict__ p, double *tu, double *tv, double *tp, size_t N,
                                         int nsteps_total, int step, unsigned long long *nonfinite, int flags) {
    const size_t b = blockIdx.y;
    const size_t toff = (b * nsteps_total + step) * N;
    unsigned long long bad = 0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (size_t)gridDim.x * blockDim.x) {
        const double x = u[b * N + q], y = v[b * N + q], z = p[b * N + q];
        if (tu) { tu[toff + q] = x; tv[toff + q] = y; tp[toff + q] = z; }
        bad += !(isfinite(x) && isfinite(y) && isfinite(z));
    }
    if ((flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(nonfinite, bad);
}

struct SpectralPlan {
    int nx, ny, n, m, batch;
    double dt, rho;
    double *mats[40];
    double *W;        // 12 derivative planes + scratch, each batch*n*m
    double *scratch[8];
    double *ui, *vi;  // predictor outputs of the run loop
    // CUDA graph of one full buffer rotation (three steps) of the run loop: a step is a chain of 15 small dependent
    // launches (126^2 problems), i.e. launch-latency-bound when issued one by one
    cudaGraphExec_t gexec;
    cudaStream_t cap;
    const void *gkey[8];
    long long gnodes;
};

#ifndef NNS_SPECTRAL_TK_LARGE
#define NNS_SPECTRAL_TK_LARGE 16
#endif
#ifndef NNS_SPECTRAL_TK_SMALL
#define NNS_SPECTRAL_TK_SMALL 32      // k-slab of the 16 x 32-tile instance (a single simulation: fewer barriers per product)
#endif
static int run_table(nns_handle *h, GemmTable &t, int batch, cudaStream_t st) {
    int mm = 0, nn = 0;
    for (int i = 0; i < t.n; ++i) { mm = t.g[i].m > mm ? t.g[i].m : mm; nn = t.g[i].n > nn ? t.g[i].n : nn; }
    // few products in the launch (a single simulation): smaller tiles, 4 or 8 times the CTAs
    const long ctas64 = (long)((nn + 63) / 64) * ((mm + 63) / 64) * t.n * batch;
    const char *tile = getenv("NNS_SPECTRAL_TILE");          // experiments: 32 / 64
    const bool small = tile ? atoi(tile) == 32 : ctas64 < 2L * h->sm_count;
    if (tile ? atoi(tile) == 16 : 2 * ctas64 < h->sm_count) {
        dim3 grid((nn + 31) / 32, (mm + 15) / 16, t.n * batch);
        spectral_gemm_kernel<16, 32, NNS_SPECTRAL_TK_SMALL><<<grid, 256, 0, st>>>(t, batch);
    } else if (small) {
        dim3 grid((nn + 31) / 32, (mm + 31) / 32, t.n * batch);
        spectral_gemm_kernel<32, 32, 16><<<grid, 256, 0, st>>>(t, batch);
    } else {
        dim3 grid((nn + 63) / 64, (mm + 63) / 64, t.n * batch);
        spectral_gemm_kernel<64, 64, NNS_SPECTRAL_TK_LARGE><<<grid, 256, 0, st>>>(t, batch);
    }
    NNS_CUDA(cudaGetLastError());
    h->launches += 1;
    return NNS_OK;
}

static GemmDesc mk(const double *A, int lda, long long sA, const double *B, int ldb, long long sB, int transB,
                   double *C, int ldc, long long sC, int m, int n, int k) {
    GemmDesc d{};
    d.A = A; d.B = B; d.C = C; d.sA = sA; d.sB = sB; d.sC = sC; d.lda = lda; d.ldb = ldb; d.ldc = ldc;
    d.m = m; d.n = n; d.k = k; d.transB = transB; d.epi = 0;
    return d;
}

enum {  // indices into SpectralPlan::mats == the order of the pointers passed to nns_spectral_create
    M_DX = 0, M_DY, M_DX2, M_DY2, M_UPINV, M_UQINV, M_UP, M_UQ, M_VPINV, M_VQINV, M_VP, M_VQ,
    M_ULX, M_ULY, M_VLX, M_VLY, M_PPINV, M_PQINV, M_PP, M_PQ, M_PLX, M_PLY, M_DXDPX, M_DYDPY, M_S,
    M_BVEC_U, M_BVEC_V, M_SC_U, M_SC_V, M_COUNT
};

void spectral_destroy(nns_handle *h);

int spectral_create(nns_handle *h, const double *const *mats, int n_mats) {
    spectral_destroy(h);
    if (n_mats != M_COUNT) { set_error("nns_spectral_create: expected %d operator arrays, got %d", (int)M_COUNT, n_mats); return NNS_ERR_INVALID; }
    SpectralPlan *sp = new SpectralPlan();
    memset(sp, 0, sizeof(*sp));
    sp->nx = h->g.nx; sp->ny = h->g.ny; sp->n = h->g.nx - 2; sp->m = h->g.ny - 2; sp->batch = h->g.batch;
    sp->dt = h->g.dt; sp->rho = h->g.rho;
    const int n = sp->n, m = sp->m;
    const size_t sizes[M_COUNT] = {
        (size_t)n * n, (size_t)m * m, (size_t)n * n, (size_t)m * m,
        (size_t)n * n, (size_t)m * m, (size_t)n * n, (size_t)m * m, (size_t)n * n, (size_t)m * m, (size_t)n * n, (size_t)m * m,
        (size_t)n, (size_t)m, (size_t)n, (size_t)m,
        (size_t)n * n, (size_t)m * m, (size_t)n * n, (size_t)m * m, (size_t)n, (size_t)m,
        (size_t)n * n, (size_t)m * m, (size_t)n * m,
        (size_t)2 * n + 2 * m, (size_t)2 * n + 2 * m, 4, 4};
    h->spectral = sp;
    for (int i = 0; i < M_COUNT; ++i) {
        if (!mats[i]) { set_error("nns_spectral_create: operator %d is null", i); return NNS_ERR_INVALID; }
        NNS_CUDA(cudaMalloc(&sp->mats[i], sizeof(double) * sizes[i]));
        NNS_CUDA(cudaMemcpy(sp->mats[i], mats[i], sizeof(double) * sizes[i], cudaMemcpyHostToDevice));
    }
    const size_t NI = (size_t)n * m * sp->batch;
    NNS_CUDA(cudaMalloc(&sp->W, sizeof(double) * NI * 12));
    for (int i = 0; i < 8; ++i) NNS_CUDA(cudaMalloc(&sp->scratch[i], sizeof(double) * NI));
    return NNS_OK;
}

void spectral_destroy(nns_handle *h) {
    SpectralPlan *sp = static_cast<SpectralPlan *>(h->spectral);
    if (!sp) return;
    for (int i = 0; i < 40; ++i) cudaFree(sp->mats[i]);
    cudaFree(sp->W);
    for (int i = 0; i < 8; ++i) cudaFree(sp->scratch[i]);
    cudaFree(sp->ui); cudaFree(sp->vi);
    if (sp->gexec) cudaGraphExecDestroy(sp->gexec);
    if (sp->cap) cudaStreamDestroy(sp->cap);
    delete sp;
    h->spectral = nullptr;
}

// _predictor_step (chorin_spectral:232-337): un,vn,un1,vn1 -> ui, vi   (all [batch][nx][ny])
int spectral_predictor(nns_handle *h, const double *un, const double *vn, const double *un1, const double *vn1,
                       double *ui, double *vi, cudaStream_t st) {
    SpectralPlan *sp = static_cast<SpectralPlan *>(h->spectral);
    const int nx = sp->nx, ny = sp->ny, n = sp->n, m = sp->m, B = sp->batch;
    const long long N = (long long)nx * ny, NI = (long long)n * m, WS = NI * B;
    const int off = ny + 1;                     // interior view of a [nx][ny] field
    double **M = sp->mats;
    int rc;
    GemmTable t{};
    const double *f[4] = {un, un1, vn, vn1};
    for (int q = 0; q < 4; ++q) {               // planes 0..7: Dx f, f Dy^T for f = un, un1, vn, vn1
        t.g[t.n++] = mk(M[M_DX], n, 0, f[q] + off, ny, N, 0, sp->W + (2 * q) * WS, m, NI, n, m, n);
        t.g[t.n++] = mk(f[q] + off, ny, N, M[M_DY], m, 0, 1, sp->W + (2 * q + 1) * WS, m, NI, n, m, m);
    }
    t.g[t.n++] = mk(M[M_DX2], n, 0, un + off, ny, N, 0, sp->W + 8 * WS, m, NI, n, m, n);
    t.g[t.n++] = mk(un + off, ny, N, M[M_DY2], m, 0, 1, sp->W + 9 * WS, m, NI, n, m, m);
    t.g[t.n++] = mk(M[M_DX2], n, 0, vn + off, ny, N, 0, sp->W + 10 * WS, m, NI, n, m, n);
    t.g[t.n++] = mk(vn + off, ny, N, M[M_DY2], m, 0, 1, sp->W + 11 * WS, m, NI, n, m, m);
    if ((rc = run_table(h, t, B, st))) return rc;
    double *Fu = sp->scratch[0], *Fv = sp->scratch[1], *T0 = sp->scratch[2], *T1 = sp->scratch[3];
    spectral_rhs_kernel<<<dim3(64, B), 256, 0, st>>>(un, vn, un1, vn1, sp->W, Fu, Fv, nx, ny, sp->dt, WS);
    h->launches += 1;
    // Helmholtz solves by diagonalisation (:285-298), u and v side by side in every launch
    t = GemmTable{};
    t.g[t.n++] = mk(M[M_UPINV], n, 0, Fu, m, NI, 0, T0, m, NI, n, m, n);
    t.g[t.n++] = mk(M[M_VPINV], n, 0, Fv, m, NI, 0, T1, m, NI, n, m, n);
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(T0, m, NI, M[M_UQINV], m, 0, 1, Fu, m, NI, n, m, m);
    t.g[t.n++] = mk(T1, m, NI, M[M_VQINV], m, 0, 1, Fv, m, NI, n, m, m);
    t.g[0].epi = 1; t.g[0].lx = M[M_ULX]; t.g[0].ly = M[M_ULY]; t.g[0].e0 = 2.0; t.g[0].e1 = -sp->dt; t.g[0].e2 = -sp->dt;
    t.g[1].epi = 1; t.g[1].lx = M[M_VLX]; t.g[1].ly = M[M_VLY]; t.g[1].e0 = 2.0; t.g[1].e1 = -sp->dt; t.g[1].e2 = -sp->dt;
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(Fu, m, NI, M[M_UQ], m, 0, 1, T0, m, NI, n, m, m);
    t.g[t.n++] = mk(Fv, m, NI, M[M_VQ], m, 0, 1, T1, m, NI, n, m, m);
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};                            // solution straight into the interiors of ui, vi
    t.g[t.n++] = mk(M[M_UP], n, 0, T0, m, NI, 0, ui + off, ny, N, n, m, n);
    t.g[t.n++] = mk(M[M_VP], n, 0, T1, m, NI, 0, vi + off, ny, N, n, m, n);
    if ((rc = run_table(h, t, B, st))) return rc;
    const int tb = (m + 31) / 32 + (n + 7) / 8;
    spectral_boundary_kernel<<<dim3(tb, B, 2), dim3(32, 8), 0, st>>>(ui, M[M_BVEC_U], M[M_SC_U], vi, M[M_BVEC_V], M[M_SC_V], nx, ny);
    h->launches += 1;
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

// _correction_step (chorin_spectral:339-383): ui, vi, p -> u, v, p_out (+ optional Q [batch][n][m])
int spectral_correct(nns_handle *h, const double *ui, const double *vi, const double *p, double *uo, double *vo,
                     double *po, double *Qout, cudaStream_t st) {
    SpectralPlan *sp = static_cast<SpectralPlan *>(h->spectral);
    const int nx = sp->nx, ny = sp->ny, n = sp->n, m = sp->m, B = sp->batch;
    const long long N = (long long)nx * ny, NI = (long long)n * m;
    const int off = ny + 1;
    double **M = sp->mats;
    double *G1 = sp->scratch[0], *G2 = sp->scratch[1], *Hh = sp->scratch[2], *T0 = sp->scratch[3];
    double *Q = Qout ? Qout : sp->scratch[4], *G3 = sp->scratch[5], *G4 = sp->scratch[6];
    int rc;
    GemmTable t{};
    t.g[t.n++] = mk(M[M_DX], n, 0, ui + off, ny, N, 0, G1, m, NI, n, m, n);
    t.g[t.n++] = mk(vi + off, ny, N, M[M_DY], m, 0, 1, G2, m, NI, n, m, m);
    if ((rc = run_table(h, t, B, st))) return rc;
    spectral_uzawa_rhs_kernel<<<dim3(64, B), 256, 0, st>>>(M[M_S], G1, G2, Hh, (size_t)NI, sp->rho / sp->dt);
    h->launches += 1;
    t = GemmTable{};
    t.g[t.n++] = mk(M[M_PPINV], n, 0, Hh, m, NI, 0, T0, m, NI, n, m, n);
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(T0, m, NI, M[M_PQINV], m, 0, 1, Hh, m, NI, n, m, m);
    t.g[0].epi = 1; t.g[0].lx = M[M_PLX]; t.g[0].ly = M[M_PLY]; t.g[0].e0 = 0.0; t.g[0].e1 = 1.0; t.g[0].e2 = 1.0;
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(Hh, m, NI, M[M_PQ], m, 0, 1, T0, m, NI, n, m, m);
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(M[M_PP], n, 0, T0, m, NI, 0, Q, m, NI, n, m, n);
    if ((rc = run_table(h, t, B, st))) return rc;
    t = GemmTable{};
    t.g[t.n++] = mk(M[M_DXDPX], n, 0, Q, m, NI, 0, G3, m, NI, n, m, n);
    t.g[t.n++] = mk(Q, m, NI, M[M_DYDPY], m, 0, 1, G4, m, NI, n, m, m);
    if ((rc = run_table(h, t, B, st))) return rc;
    spectral_project_kernel<<<dim3(64, B), 256, 0, st>>>(ui, vi, p, G3, G4, Q, uo, vo, po, nx, ny, sp->dt, sp->rho);
    h->launches += 1;
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

// nsteps of NavierStokesSystem.step + the rotation of simulate (chorin_spectral:54-57, :547-570).
// bufU/bufV: roles at entry 0 = cur, 1 = prev, 2 = scratch; on return buffer 0 holds step n and
// buffer 1 step n-1.
int spectral_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, double *tu, double *tv,
                 double *tp, cudaStream_t st) {
    SpectralPlan *sp = static_cast<SpectralPlan *>(h->spectral);
    if (!sp) { set_error("chorin_spectral: operators not uploaded (call nns_spectral_set_operators first)"); return NNS_ERR_INVALID; }
    const size_t N = (size_t)sp->nx * sp->ny, bytes = sizeof(double) * N * sp->batch;
    if (!sp->ui) NNS_CUDA(cudaMalloc(&sp->ui, bytes));
    if (!sp->vi) NNS_CUDA(cudaMalloc(&sp->vi, bytes));
    int cur = 0, prev = 1, nxt = 2, rc;
    int n0 = 0;
    const char *nog = getenv("NNS_SPECTRAL_NOGRAPH");
    if (!tu && nsteps >= 6 && !(nog && nog[0] == '1')) {
        // replay whole rotations (3 steps: the buffer roles are back where they started) from a captured graph
        const void *key[8] = {bufU[0], bufU[1], bufU[2], bufV[0], bufV[1], bufV[2], p, (const void *)(size_t)h->params.flags};
        if (!sp->gexec || memcmp(key, sp->gkey, sizeof(key)) != 0) {
            if (sp->gexec) { cudaGraphExecDestroy(sp->gexec); sp->gexec = nullptr; }
            if (!sp->cap) NNS_CUDA(cudaStreamCreateWithFlags(&sp->cap, cudaStreamNonBlocking));
            const long long l0 = h->launches;
            NNS_CUDA(cudaStreamBeginCapture(sp->cap, cudaStreamCaptureModeThreadLocal));
            int c = 0, pv = 1, nx2 = 2;
            rc = NNS_OK;
            for (int k = 0; k < 3 && rc == NNS_OK; ++k) {
                rc = spectral_predictor(h, bufU[c], bufV[c], bufU[pv], bufV[pv], sp->ui, sp->vi, sp->cap);
                if (rc == NNS_OK) rc = spectral_correct(h, sp->ui, sp->vi, p, bufU[nx2], bufV[nx2], p, nullptr, sp->cap);
                if (rc == NNS_OK && (h->params.flags & NNS_FLAG_CHECK_FINITE)) {
                    spectral_snapshot_kernel<<<dim3(64, sp->batch), 256, 0, sp->cap>>>(bufU[nx2], bufV[nx2], p, nullptr, nullptr, nullptr, N,
                                                                                    nsteps, 0, h->d_nonfinite, h->params.flags);
                    h->launches += 1;
                }
                const int t = pv; pv = c; c = nx2; nx2 = t;
            }
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(sp->cap, &graph);
            sp->gnodes = h->launches - l0;
            h->launches = l0;
            if (rc != NNS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) { set_error("chorin_spectral: graph capture failed: %s", cudaGetErrorString(e)); return NNS_ERR_CUDA; }
            e = cudaGraphInstantiate(&sp->gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { set_error("chorin_spectral: graph instantiation failed: %s", cudaGetErrorString(e)); return NNS_ERR_CUDA; }
            memcpy(sp->gkey, key, sizeof(key));
        }
        const int reps = nsteps / 3;
        for (int k = 0; k < reps; ++k) NNS_CUDA(cudaGraphLaunch(sp->gexec, st));
        h->launches += sp->gnodes * reps;
        n0 = reps * 3;
    }
    for (int n = n0; n < nsteps; ++n) {
        if ((rc = spectral_predictor(h, bufU[cur], bufV[cur], bufU[prev], bufV[prev], sp->ui, sp->vi, st))) return rc;
        if ((rc = spectral_correct(h, sp->ui, sp->vi, p, bufU[nxt], bufV[nxt], p, nullptr, st))) return rc;
        if (tu || (h->params.flags & NNS_FLAG_CHECK_FINITE)) {
            spectral_snapshot_kernel<<<dim3(64, sp->batch), 256, 0, st>>>(bufU[nxt], bufV[nxt], p, tu, tv, tp, N, nsteps, n,
                                                                           h->d_nonfinite, h->params.flags);
            h->launches += 1;
        }
        const int t = prev; prev = cur; cur = nxt; nxt = t;
    }
    NNS_CUDA(cudaGetLastError());
    if (cur != 0) {
        // final roles -> caller's buffers (buffer 0 = step n, buffer 1 = step n-1)
        if (cur == 2) {          // (cur, prev) = (2, 0)
            NNS_CUDA(cudaMemcpyAsync(bufU[1], bufU[0], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufV[1], bufV[0], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufU[0], bufU[2], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufV[0], bufV[2], bytes, cudaMemcpyDeviceToDevice, st));
        } else {                 // (cur, prev) = (1, 2)
            NNS_CUDA(cudaMemcpyAsync(bufU[0], bufU[1], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufV[0], bufV[1], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufU[1], bufU[2], bytes, cudaMemcpyDeviceToDevice, st));
            NNS_CUDA(cudaMemcpyAsync(bufV[1], bufV[2], bytes, cudaMemcpyDeviceToDevice, st));
        }
    }
    return NNS_OK;
}

}  // namespace nns
