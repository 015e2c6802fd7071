// chorin_fd_chip.cu -- chorin_fd time step with the whole pressure grid of one member resident
// in one SM's shared memory ("chip" path): one CTA per ensemble member, all phases of the
// step fused in one launch, nsteps steps per launch.
//
// Reference semantics reproduced (src/chorin_fd/simulate.py of mhw32/neural-navier-stokes):
//   phase A  _explicit_predictor_step :63-91 (x-only advection differences kept) or
//            _semi_implicit_predictor_step :93-167, then u_bc / v_bc in list order :221-225
//   phase B  _get_pressure :169-202 -- lexicographic Gauss-Seidel SOR, <= nit-1 sweeps, exit
//            when max|p - pPrev| <= tol.  Executed as a hyperplane wavefront t = i + j + 2s
//            (cell (i,j) of sweep s): every dependency of the lexicographic order lies at
//            t-1 or t-2, so all sweeps are pipelined through ONE pass over the grid with a
//            single __syncthreads per stage and the result equals the sequential order.
//   phase C  p_bc in list order :230-231, _correction_step :204-210, trajectory snapshot
//            :263-265.
//
// Shared-memory layout of p ("split rows"): two half arrays by column parity,
//   P[h][i][jh],  h = j & 1, jh = j >> 1, pitch PH, half stride HS.
// At stage t the active cells of row i are j = (t-i) - 2s, s = s_lo..s_hi: one parity, i.e. a
// CONTIGUOUS run of jh in one half array, so a warp's 32 lanes (= 32 consecutive sweeps)
// read and write consecutive 8-byte words: conflict-free LDS/STS without padding.
#include "nns_common.cuh"

namespace nns {

struct ChipArgs {
    Geometry g;
    BcList ubc, vbc, pbc;
    const double *nu_b;      // [batch] or null
    const double *bcval;     // [batch][n_bcs] or null
    int n_bcs;
    int PH, HS;              // half-row pitch, half-array stride (doubles)
    int nsteps, nsteps_total, step0;
    int phases;              // bit0 A, bit1 B, bit2 C
    int fixup;               // copy final cur/prev into buffers 0/1
    int flags;
    double *bufU[3], *bufV[3];   // roles at entry: 0 = cur (u^n), 1 = prev (u^{n-1}), 2 = next
    double *p;
    double *cprime;          // global C' scratch [batch][2*HS] when it does not fit in smem
    double *traj_u, *traj_v, *traj_p;   // [batch][nsteps_total][nx][ny] or null
    int32_t *sweeps;         // [nsteps_total][batch] or null
    unsigned long long *nonfinite;
};

__device__ __forceinline__ int split_off(int i, int j, int PH, int HS) {
    return (j & 1) * HS + i * PH + (j >> 1);
}

// Thomas solve along axis 0 for all interior columns, constant tridiagonal (-off, diag, -off)
// i.e. np.linalg.solve(A, rhs) of chorin_fd/simulate.py:137,153,159,165 (A diagonally
// dominant => LAPACK's partial pivoting never swaps, so this is the same elimination).
// rhs/x are row-major [nx][ny] interiors; cpr holds the nx forward-sweep coefficients.
__device__ void cta_thomas_axis0(double *x, int nx, int ny, double diag, double off, double *cpr) {
    // forward coefficients (same for every column); thread 0 builds them once per call
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int i = 1; i < nx - 1; ++i) {
            const double m = diag - off * c;   // pivot after eliminating the sub-diagonal
            c = off / m;
            cpr[i] = c;                        // c_i' = off / m_i
            cpr[nx + i] = 1.0 / m;             // 1 / m_i
        }
    }
    __syncthreads();
    for (int j = 1 + threadIdx.x; j < ny - 1; j += blockDim.x) {
        double d = 0.0;
        for (int i = 1; i < nx - 1; ++i) {     // forward: d_i' = (d_i - off*d_{i-1}') / m_i
            d = (x[(size_t)i * ny + j] - off * d) * cpr[nx + i];
            x[(size_t)i * ny + j] = d;
        }
        double xn = 0.0;
        for (int i = nx - 2; i >= 1; --i) {    // backward: x_i = d_i' - c_i' x_{i+1}
            xn = x[(size_t)i * ny + j] - cpr[i] * xn;
            x[(size_t)i * ny + j] = xn;
        }
    }
    __syncthreads();
}

template <bool TRACK>
__device__ __forceinline__ void sor_wavefront(double *Ps, const double *Cs, int nx, int ny, int PH, int HS,
                                              int cap, double ca, double cb, double beta, double tol,
                                              unsigned long long &mask, int *viol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int tmax = (nx - 2) + (ny - 2) + 2 * (cap - 1);
    const int span = (ny - 2) + 2 * (cap - 1);
    for (int t = 2; t <= tmax; ++t) {
        const int ilo = max(1, t - span), ihi = min(nx - 2, t - 1);
        for (int i = ilo + warp; i <= ihi; i += nwarps) {
            const int a = t - i;                              // j + 2s
            const int over = a - (ny - 2);
            const int s_lo = over > 0 ? (over + 1) >> 1 : 0;
            const int s_hi = min(cap - 1, (a - 1) >> 1);
            const int h = a & 1, jh0 = a >> 1;
            const int rowc = h * HS + i * PH;
            const double *Po = Ps + (1 - h) * HS + i * PH - (1 - h);   // W = Po[jh], E = Po[jh+1]
            for (int s = s_lo + lane; s <= s_hi; s += 32) {
                const int jh = jh0 - s;
                const double c = Ps[rowc + jh];
                const double n = Ps[rowc - PH + jh], so = Ps[rowc + PH + jh];
                const double w = Po[jh], e = Po[jh + 1];
                const double cp = Cs[rowc + jh];
                const double d = fma(ca, n + so, fma(cb, e + w, fma(-beta, c, -cp)));
                Ps[rowc + jh] = c + d;
                if (TRACK) {
                    if (!(fabs(d) <= tol)) {
                        if (s < 64) mask |= 1ull << s;
                        else viol[s - 64] = 1;
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <bool CP_SMEM>
__global__ void __launch_bounds__(1024, 1) chorin_chip_kernel(const ChipArgs a) {
    extern __shared__ double smem[];
    __shared__ unsigned long long s_mask;
    __shared__ int s_need;

    const int nx = a.g.nx, ny = a.g.ny, PH = a.PH, HS = a.HS;
    const size_t N = (size_t)nx * ny;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    double *Ps = smem;
    double *Cs = CP_SMEM ? smem + 2 * HS : a.cprime + (size_t)b * 2 * HS;
    double *aux = smem + (CP_SMEM ? 4 : 2) * HS;           // 2*nx doubles (Thomas coefficients)
    int *viol = reinterpret_cast<int *>(aux + 2 * nx);     // max(0, nit-65) ints

    const double dt = a.g.dt, dx = a.g.dx, dy = a.g.dy, rho = a.g.rho, beta = a.g.beta;
    const double nu = a.nu_b ? a.nu_b[b] : a.g.nu;
    const double *bcval = a.bcval ? a.bcval + (size_t)b * a.n_bcs : nullptr;
    const double dx2 = dx * dx, dy2 = dy * dy;
    const double den = 2.0 * dx2 + 2.0 * dy2;
    const double ca = beta * dy2 / den, cb = beta * dx2 / den, cc = beta / den;
    const double cu = dx * rho * dy2 / dt, cv = dy * rho * dx2 / dt;
    const double r2dx = 1.0 / (2.0 * dx), r2dy = 1.0 / (2.0 * dy), rdx2 = 1.0 / dx2, rdy2 = 1.0 / dy2;
    const int cap = a.g.nit - 1;

    int cur = 0, prev = 1, nxt = 2;
    double *pg = a.p + (size_t)b * N;

    for (int n = 0; n < a.nsteps; ++n) {
        const double *uc = a.bufU[cur] + (size_t)b * N, *vc = a.bufV[cur] + (size_t)b * N;
        const double *up = a.bufU[prev] + (size_t)b * N, *vp = a.bufV[prev] + (size_t)b * N;
        double *un = a.bufU[nxt] + (size_t)b * N, *vn = a.bufV[nxt] + (size_t)b * N;

        // ------------------------------ phase A: predictor + u/v BCs --------------------
        if (a.phases & 1) {
            if (a.g.method == NNS_METHOD_EXPLICIT) {
                for (int i = warp; i < nx; i += nwarps)
                    for (int j = lane; j < ny; j += 32) {
                        const size_t q = (size_t)i * ny + j;
                        const double u0 = uc[q], v0 = vc[q];
                        double ru = u0, rv = v0;
                        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
                            const double u1c = up[q], v1c = vp[q];
                            const double uS = uc[q + ny], uN = uc[q - ny], uE = uc[q + 1], uW = uc[q - 1];
                            const double vS = vc[q + ny], vN = vc[q - ny], vE = vc[q + 1], vW = vc[q - 1];
                            const double pS = up[q + ny], pN = up[q - ny], pE = up[q + 1], pW = up[q - 1];
                            const double qS = vp[q + ny], qN = vp[q - ny], qE = vp[q + 1], qW = vp[q - 1];
                            // both advection terms difference along axis 0 (chorin_fd:74,76,83,85)
                            const double k0 = u0 * r2dx + v0 * r2dy, k1 = u1c * r2dx + v1c * r2dy;
                            const double advu = 1.5 * (k0 * (uS - uN)) - 0.5 * (k1 * (pS - pN));
                            const double advv = 1.5 * (k0 * (vS - vN)) - 0.5 * (k1 * (qS - qN));
                            const double lapu = 1.5 * ((uS - 2.0 * u0 + uN) * rdx2 + (uE - 2.0 * u0 + uW) * rdy2) -
                                                0.5 * ((pS - 2.0 * u1c + pN) * rdx2 + (pE - 2.0 * u1c + pW) * rdy2);
                            const double lapv = 1.5 * ((vS - 2.0 * v0 + vN) * rdx2 + (vE - 2.0 * v0 + vW) * rdy2) -
                                                0.5 * ((qS - 2.0 * v1c + qN) * rdx2 + (qE - 2.0 * v1c + qW) * rdy2);
                            ru = u0 - dt * advu + (dt * nu) * lapu;
                            rv = v0 - dt * advv + (dt * nu) * lapv;
                        }
                        un[q] = ru;
                        vn[q] = rv;
                    }
                __syncthreads();
            } else {
                // semi-implicit: AB2 advection + Crank-Nicolson ADI, all four solves along axis 0
                // (chorin_fd:93-167; diagonal (2/nu)dx^2+2dt :108, vC scaled by dx^2 :150).
                const double kx = 2.0 / nu * dx2, ky = 2.0 / nu * dy2;
                // stage 1 right-hand sides uC, vC -> un, vn interiors (edges = copies of u^n)
                for (int i = warp; i < nx; i += nwarps)
                    for (int j = lane; j < ny; j += 32) {
                        const size_t q = (size_t)i * ny + j;
                        const double u0 = uc[q], v0 = vc[q];
                        double ru = u0, rv = v0;
                        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
                            const double u1c = up[q], v1c = vp[q];
                            const double uS = uc[q + ny], uN = uc[q - ny], uE = uc[q + 1], uW = uc[q - 1];
                            const double vS = vc[q + ny], vN = vc[q - ny], vE = vc[q + 1], vW = vc[q - 1];
                            const double pS = up[q + ny], pN = up[q - ny], pE = up[q + 1], pW = up[q - 1];
                            const double qS = vp[q + ny], qN = vp[q - ny], qE = vp[q + 1], qW = vp[q - 1];
                            const double uHn = u0 * (uS - uN) * r2dx + v0 * (uE - uW) * r2dy;
                            const double uHn1 = u1c * (pS - pN) * r2dx + v1c * (pE - pW) * r2dy;
                            const double vHn = u0 * (vS - vN) * r2dx + v0 * (vE - vW) * r2dy;
                            const double vHn1 = u1c * (qS - qN) * r2dx + v1c * (qE - qW) * r2dy;
                            const double uC2 = dt * nu * ((uS - 2.0 * u0 + uN) * rdx2 + (uE - 2.0 * u0 + uW) * rdy2);
                            const double vC2 = dt * nu * ((vS - 2.0 * v0 + vN) * rdx2 + (vE - 2.0 * v0 + vW) * rdy2);
                            ru = kx * (0.5 * dt * (3.0 * uHn - uHn1) + uC2);
                            rv = kx * (0.5 * dt * (3.0 * vHn - vHn1) + vC2);
                        }
                        un[q] = ru;
                        vn[q] = rv;
                    }
                __syncthreads();
                cta_thomas_axis0(un, nx, ny, kx + 2.0 * dt, -dt, aux);   // ut
                cta_thomas_axis0(vn, nx, ny, kx + 2.0 * dt, -dt, aux);   // vt
                // stage 2 right-hand sides uS, vS (in place on the interiors)
                for (int i = 1 + warp; i < nx - 1; i += nwarps)
                    for (int j = 1 + lane; j < ny - 1; j += 32) {
                        const size_t q = (size_t)i * ny + j;
                        const double u0 = uc[q], v0 = vc[q];
                        un[q] = ky * (un[q] + u0) - dt * (uc[q + 1] - 2.0 * u0 + uc[q - 1]);
                        vn[q] = ky * (vn[q] + v0) - dt * (vc[q + 1] - 2.0 * v0 + vc[q - 1]);
                    }
                __syncthreads();
                cta_thomas_axis0(un, nx, ny, ky + 2.0 * dt, -dt, aux);   // B is applied along axis 0 (:159)
                cta_thomas_axis0(vn, nx, ny, ky + 2.0 * dt, -dt, aux);
            }
            cta_apply_bc_global(un, nx, ny, a.ubc, bcval, dx, dy);
            cta_apply_bc_global(vn, nx, ny, a.vbc, bcval, dx, dy);
        }

        // ------------------------------ phase B: SOR pressure ----------------------------
        int need = 0;
        if (a.phases & 2) {
            if (tid == 0) s_mask = 0ull;
            for (int k = tid; k < cap - 64; k += blockDim.x) viol[k] = 0;
            for (int i = warp; i < nx; i += nwarps)
                for (int j = lane; j < ny; j += 32) {
                    const size_t q = (size_t)i * ny + j;
                    const int so = split_off(i, j, PH, HS);
                    Ps[so] = pg[q];
                    double c = 0.0;
                    if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1)
                        c = cc * (cu * (un[q] - un[q - ny]) + cv * (vn[q] - vn[q - 1]));
                    Cs[so] = c;
                }
            __syncthreads();
            need = cap > 0 ? cap : 0;
            if (cap > 0) {
                unsigned long long mask = 0ull;
                sor_wavefront<true>(Ps, Cs, nx, ny, PH, HS, cap, ca, cb, beta, a.g.tol, mask, viol);
                unsigned lo = (unsigned)mask, hi = (unsigned)(mask >> 32);
                lo = __reduce_or_sync(0xffffffffu, lo);
                hi = __reduce_or_sync(0xffffffffu, hi);
                if (lane == 0) atomicOr(&s_mask, ((unsigned long long)hi << 32) | lo);
                __syncthreads();
                if (tid == 0) {
                    int nd = cap;
                    const int c64 = cap < 64 ? cap : 64;
                    const unsigned long long full = c64 == 64 ? ~0ull : ((1ull << c64) - 1ull);
                    const unsigned long long clr = ~s_mask & full;
                    if (clr) nd = __ffsll((long long)clr);      // first sweep s with max|dp| <= tol: s+1 sweeps run
                    else
                        for (int s = 64; s < cap; ++s)
                            if (!viol[s - 64]) { nd = s + 1; break; }
                    s_need = nd;
                }
                __syncthreads();
                need = s_need;
                if (need < cap) {
                    // the sequential loop would have stopped after `need` sweeps: redo from the
                    // untouched global p with the sweep count capped (rare: near steady state)
                    for (int i = warp; i < nx; i += nwarps)
                        for (int j = lane; j < ny; j += 32)
                            Ps[split_off(i, j, PH, HS)] = pg[(size_t)i * ny + j];
                    __syncthreads();
                    unsigned long long dummy = 0ull;
                    sor_wavefront<false>(Ps, Cs, nx, ny, PH, HS, need, ca, cb, beta, a.g.tol, dummy, viol);
                }
            }
            if (a.sweeps && tid == 0) a.sweeps[(size_t)(a.step0 + n) * a.g.batch + b] = need;
        } else if (a.phases & 4) {
            for (int i = warp; i < nx; i += nwarps)
                for (int j = lane; j < ny; j += 32) Ps[split_off(i, j, PH, HS)] = pg[(size_t)i * ny + j];
            __syncthreads();
        }

        // ------------------------------ phase C: p BCs, projection, snapshot --------------
        if (a.phases & 4) {
            for (int k = 0; k < a.pbc.n; ++k) {
                const double g = bcval ? bcval[a.pbc.slot[k]] : a.pbc.value[k];
                const int side = a.pbc.side[k];
                const bool neu = a.pbc.type[k] == NNS_BC_NEUMANN;
                if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
                    const int i = side == NNS_SIDE_LEFT ? 0 : nx - 1, in = side == NNS_SIDE_LEFT ? 1 : nx - 2;
                    const double sg = side == NNS_SIDE_LEFT ? -dx : dx;
                    for (int j = tid; j < ny; j += blockDim.x)
                        Ps[split_off(i, j, PH, HS)] = neu ? Ps[split_off(in, j, PH, HS)] + sg * g : g;
                } else {
                    const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1, jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
                    const double sg = side == NNS_SIDE_BOTTOM ? -dy : dy;
                    for (int i = tid; i < nx; i += blockDim.x)
                        Ps[split_off(i, j, PH, HS)] = neu ? Ps[split_off(i, jn, PH, HS)] + sg * g : g;
                }
                __syncthreads();
            }
        }
        if (a.phases & 6) {
            const size_t toff = ((size_t)b * a.nsteps_total + (a.step0 + n)) * N;
            const double kx = dt / (2.0 * dx), ky = dt / (2.0 * dy);
            unsigned long long bad = 0;
            for (int i = warp; i < nx; i += nwarps)
                for (int j = lane; j < ny; j += 32) {
                    const size_t q = (size_t)i * ny + j;
                    const double pc = Ps[split_off(i, j, PH, HS)];
                    pg[q] = pc;
                    if (a.traj_p) a.traj_p[toff + q] = pc;
                    if (a.phases & 4) {
                        double ru = un[q], rv = vn[q];
                        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
                            ru -= kx * (Ps[split_off(i + 1, j, PH, HS)] - Ps[split_off(i - 1, j, PH, HS)]);
                            rv -= ky * (Ps[split_off(i, j + 1, PH, HS)] - Ps[split_off(i, j - 1, PH, HS)]);
                            un[q] = ru;
                            vn[q] = rv;
                        }
                        if (a.traj_u) a.traj_u[toff + q] = ru;
                        if (a.traj_v) a.traj_v[toff + q] = rv;
                        if (a.flags & NNS_FLAG_CHECK_FINITE) bad += !(isfinite(ru) && isfinite(rv) && isfinite(pc));
                    }
                }
            if ((a.flags & NNS_FLAG_CHECK_FINITE) && bad) atomicAdd(a.nonfinite, bad);
        }
        __syncthreads();
        const int t = prev; prev = cur; cur = nxt; nxt = t;
    }

    if (a.fixup && cur != 0) {
        // final roles -> caller's buffers: buffer 0 must hold step n, buffer 1 step n-1
        double *U0 = a.bufU[0] + (size_t)b * N, *U1 = a.bufU[1] + (size_t)b * N, *U2 = a.bufU[2] + (size_t)b * N;
        double *V0 = a.bufV[0] + (size_t)b * N, *V1 = a.bufV[1] + (size_t)b * N, *V2 = a.bufV[2] + (size_t)b * N;
        if (cur == 2) {          // (cur,prev) = (2,0): prev first, then cur
            for (size_t q = tid; q < N; q += blockDim.x) { U1[q] = U0[q]; V1[q] = V0[q]; }
            __syncthreads();
            for (size_t q = tid; q < N; q += blockDim.x) { U0[q] = U2[q]; V0[q] = V2[q]; }
        } else {                 // (cur,prev) = (1,2)
            for (size_t q = tid; q < N; q += blockDim.x) { U0[q] = U1[q]; V0[q] = V1[q]; }
            __syncthreads();
            for (size_t q = tid; q < N; q += blockDim.x) { U1[q] = U2[q]; V1[q] = V2[q]; }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------

struct ChipPlan {
    int PH, HS, threads;
    size_t smem_bytes;
    bool cp_smem;
    bool fits;
};

ChipPlan chorin_chip_plan(const nns_handle *h) {
    ChipPlan pl{};
    const int nx = h->g.nx, ny = h->g.ny;
    pl.PH = (ny + 1) / 2;
    int hs = nx * pl.PH;
    hs = ((hs + 15) / 16) * 16 + 8;          // half stride = 8 mod 16 doubles: the halves sit 16 banks apart
    pl.HS = hs;
    const size_t extra = sizeof(double) * 2 * nx + sizeof(int) * (size_t)(h->g.nit > 65 ? h->g.nit - 65 : 0) + 64;
    const size_t one = sizeof(double) * 2 * (size_t)hs;
    const size_t lim = (size_t)h->max_smem_optin;
    pl.cp_smem = 2 * one + extra <= lim;
    pl.smem_bytes = (pl.cp_smem ? 2 : 1) * one + extra;
    pl.fits = pl.smem_bytes <= lim;
    const long cells = (long)nx * ny;
    pl.threads = cells >= 8192 ? 1024 : cells >= 1024 ? 512 : 256;
    return pl;
}

int chorin_chip_launch(nns_handle *h, ChipArgs &a, cudaStream_t st, int m0, int count) {
    const ChipPlan pl = chorin_chip_plan(h);
    if (!pl.fits) {
        set_error("chorin_fd chip path: grid %dx%d needs %zu B of shared memory (> %d)", h->g.nx, h->g.ny,
                  pl.smem_bytes, h->max_smem_optin);
        return NNS_ERR_UNSUPPORTED;
    }
    a.PH = pl.PH;
    a.HS = pl.HS;
    if (!pl.cp_smem && !h->d_cprime)
        NNS_CUDA(cudaMalloc(&h->d_cprime, sizeof(double) * 2 * (size_t)pl.HS * h->g.batch));
    a.cprime = h->d_cprime ? h->d_cprime + (size_t)m0 * 2 * pl.HS : nullptr;
    if (pl.cp_smem) {
        NNS_CUDA(cudaFuncSetAttribute(chorin_chip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        chorin_chip_kernel<true><<<count, pl.threads, pl.smem_bytes, st>>>(a);
    } else {
        NNS_CUDA(cudaFuncSetAttribute(chorin_chip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        chorin_chip_kernel<false><<<count, pl.threads, pl.smem_bytes, st>>>(a);
    }
    NNS_CUDA(cudaGetLastError());
    h->launches += 1;
    return NNS_OK;
}

bool chorin_chip_fits(const nns_handle *h) { return chorin_chip_plan(h).fits; }

int chorin_chip_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                    int step0, int phases, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps,
                    cudaStream_t st, int m0, int count) {
    // Members [m0, m0+count): field / trajectory / sweeps pointers are already offset by the
    // caller; the per-member parameter tables are offset here.
    if (count < 0) count = h->g.batch - m0;
    ChipArgs a{};
    a.g = h->g;
    a.ubc = h->bc[0]; a.vbc = h->bc[1]; a.pbc = h->bc[2];
    a.nu_b = h->d_nu ? h->d_nu + m0 : nullptr;
    a.bcval = h->d_bcval ? h->d_bcval + (size_t)m0 * h->n_bcs : nullptr;
    a.n_bcs = h->n_bcs;
    a.nsteps = nsteps; a.nsteps_total = nsteps_total; a.step0 = step0;
    a.phases = phases; a.fixup = fixup; a.flags = h->params.flags;
    for (int k = 0; k < 3; ++k) { a.bufU[k] = bufU[k]; a.bufV[k] = bufV[k]; }
    a.p = p;
    a.traj_u = tu; a.traj_v = tv; a.traj_p = tp;
    a.sweeps = sweeps;
    a.nonfinite = h->d_nonfinite;
    return chorin_chip_launch(h, a, st, m0, count);
}

}  // namespace nns
