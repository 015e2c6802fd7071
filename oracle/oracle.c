/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * finite-difference Navier-Stokes step (mhw32/neural-navier-stokes), used as the
 * parity checker for the CUDA path and as the timed CPU baseline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (neural-navier-stokes_b200/) never does.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the reference's own
 * Python classes from /root/reference (in the build container) and checks this
 * restatement against them; the resulting fixtures are committed under tests/golden/.
 * The arithmetic below keeps the reference's operation order on purpose (no FMA
 * contraction: compile with -ffp-contract=off), so for the explicit chorin_fd and the
 * direct_fd paths it reproduces the numpy results bit for bit.
 *
 * Conventions (reference: src/boundary.py:34-86, src/chorin_fd/simulate.py,
 * src/direct_fd/simulate.py): arrays are C-contiguous double [nx][ny], index [i][j],
 * j fastest.  Sides: 0=left A[0,:], 1=right A[-1,:], 2=bottom A[:,0], 3=top A[:,-1].
 * Types: 0=dirichlet, 1=neumann.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int side;     /* 0 left, 1 right, 2 bottom, 3 top */
    int type;     /* 0 dirichlet, 1 neumann */
    double value;
} orc_bc;

#define IDX(i, j) ((size_t)(i) * (size_t)ny + (size_t)(j))

/* Python's float ** 2 goes through libm pow(); keep the call (no folding to x*x). */
static double py_sq(double x) {
    volatile double two = 2.0;
    return pow(x, two);
}

/* ---- src/boundary.py:34-48 (Dirichlet.apply) and :56-86 (Neumann.apply) ---------- */
void orc_bc_apply(double *A, int nx, int ny, const orc_bc *bc, double dx, double dy) {
    const double g = bc->value;
    if (bc->type == 0) {
        switch (bc->side) {
        case 0: for (int j = 0; j < ny; ++j) A[IDX(0, j)] = g; break;
        case 1: for (int j = 0; j < ny; ++j) A[IDX(nx - 1, j)] = g; break;
        case 2: for (int i = 0; i < nx; ++i) A[IDX(i, 0)] = g; break;
        case 3: for (int i = 0; i < nx; ++i) A[IDX(i, ny - 1)] = g; break;
        }
    } else {
        switch (bc->side) {
        case 0: for (int j = 0; j < ny; ++j) A[IDX(0, j)] = A[IDX(1, j)] - dx * g; break;
        case 1: for (int j = 0; j < ny; ++j) A[IDX(nx - 1, j)] = A[IDX(nx - 2, j)] + dx * g; break;
        case 2: for (int i = 0; i < nx; ++i) A[IDX(i, 0)] = A[IDX(i, 1)] - dy * g; break;
        case 3: for (int i = 0; i < nx; ++i) A[IDX(i, ny - 1)] = A[IDX(i, ny - 2)] + dy * g; break;
        }
    }
}

void orc_bc_apply_list(double *A, int nx, int ny, const orc_bc *bcs, int n, double dx, double dy) {
    for (int k = 0; k < n; ++k) orc_bc_apply(A, nx, ny, &bcs[k], dx, dy);
}

/* ---- src/chorin_fd/simulate.py:63-91 (_explicit_predictor_step) ------------------
 * Quirk kept: both advection terms difference along axis 0 (:74,76,83,85).          */
void orc_chorin_explicit_predictor(const double *un, const double *vn, const double *un1,
                                   const double *vn1, double *ui, double *vi, int nx, int ny,
                                   double dt, double dx, double dy, double nu) {
    const double dx2 = py_sq(dx), dy2 = py_sq(dy);
    const double twodx = 2 * dx, twody = 2 * dy, dtnu = dt * nu;
    memcpy(ui, un, sizeof(double) * (size_t)nx * ny);
    memcpy(vi, vn, sizeof(double) * (size_t)nx * ny);
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j) {
            const double uc = un[IDX(i, j)], vc = vn[IDX(i, j)];
            const double u1c = un1[IDX(i, j)], v1c = vn1[IDX(i, j)];
            {   /* u momentum */
                const double d = un[IDX(i + 1, j)] - un[IDX(i - 1, j)];
                const double d1 = un1[IDX(i + 1, j)] - un1[IDX(i - 1, j)];
                const double adv = uc * d / twodx + vc * d / twody;
                const double adv1 = u1c * d1 / twodx + v1c * d1 / twody;
                const double lap = (un[IDX(i + 1, j)] - 2 * uc + un[IDX(i - 1, j)]) / dx2 +
                                   (un[IDX(i, j + 1)] - 2 * uc + un[IDX(i, j - 1)]) / dy2;
                const double lap1 = (un1[IDX(i + 1, j)] - 2 * u1c + un1[IDX(i - 1, j)]) / dx2 +
                                    (un1[IDX(i, j + 1)] - 2 * u1c + un1[IDX(i, j - 1)]) / dy2;
                ui[IDX(i, j)] = uc - dt * (1.5 * adv - 0.5 * adv1) + dtnu * (1.5 * lap - 0.5 * lap1);
            }
            {   /* v momentum */
                const double d = vn[IDX(i + 1, j)] - vn[IDX(i - 1, j)];
                const double d1 = vn1[IDX(i + 1, j)] - vn1[IDX(i - 1, j)];
                const double adv = uc * d / twodx + vc * d / twody;
                const double adv1 = u1c * d1 / twodx + v1c * d1 / twody;
                const double lap = (vn[IDX(i + 1, j)] - 2 * vc + vn[IDX(i - 1, j)]) / dx2 +
                                   (vn[IDX(i, j + 1)] - 2 * vc + vn[IDX(i, j - 1)]) / dy2;
                const double lap1 = (vn1[IDX(i + 1, j)] - 2 * v1c + vn1[IDX(i - 1, j)]) / dx2 +
                                    (vn1[IDX(i, j + 1)] - 2 * v1c + vn1[IDX(i, j - 1)]) / dy2;
                vi[IDX(i, j)] = vc - dt * (1.5 * adv - 0.5 * adv1) + dtnu * (1.5 * lap - 0.5 * lap1);
            }
        }
}

/* Dense LU with partial pivoting, many right-hand sides: restates what
 * np.linalg.solve (LAPACK dgesv) does for the (n x n) @ X = (n x m) systems at
 * src/chorin_fd/simulate.py:137,153,159,165.  The matrix is the dense tridiagonal
 * one built at :105-121.  LAPACK's blocked update order differs, so this leg is
 * pinned to ~1e-13, not bit for bit.                                               */
static int dense_solve(int n, int m, double diag, double off, const double *rhs, double *x) {
    double *a = (double *)calloc((size_t)n * n, sizeof(double));
    double *b = (double *)malloc(sizeof(double) * (size_t)n * m);
    if (!a || !b) { free(a); free(b); return -1; }
    for (int i = 0; i < n; ++i) {
        a[(size_t)i * n + i] = diag;
        if (i > 0) a[(size_t)i * n + i - 1] = off;
        if (i + 1 < n) a[(size_t)i * n + i + 1] = off;
    }
    memcpy(b, rhs, sizeof(double) * (size_t)n * m);
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(a[(size_t)k * n + k]);
        for (int r = k + 1; r < n; ++r)
            if (fabs(a[(size_t)r * n + k]) > best) { best = fabs(a[(size_t)r * n + k]); piv = r; }
        if (piv != k) {
            for (int c = 0; c < n; ++c) { double t = a[(size_t)k * n + c]; a[(size_t)k * n + c] = a[(size_t)piv * n + c]; a[(size_t)piv * n + c] = t; }
            for (int c = 0; c < m; ++c) { double t = b[(size_t)k * m + c]; b[(size_t)k * m + c] = b[(size_t)piv * m + c]; b[(size_t)piv * m + c] = t; }
        }
        const double pk = a[(size_t)k * n + k];
        for (int r = k + 1; r < n; ++r) {
            const double arc = a[(size_t)r * n + k];
            if (arc == 0.0) continue;
            const double l = arc / pk;
            for (int c = k + 1; c < n; ++c) a[(size_t)r * n + c] -= l * a[(size_t)k * n + c];
            for (int c = 0; c < m; ++c) b[(size_t)r * m + c] -= l * b[(size_t)k * m + c];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        for (int c = 0; c < m; ++c) {
            double s = b[(size_t)k * m + c];
            for (int q = k + 1; q < n; ++q) {
                const double akq = a[(size_t)k * n + q];
                if (akq != 0.0) s -= akq * x[(size_t)q * m + c];
            }
            x[(size_t)k * m + c] = s / a[(size_t)k * n + k];
        }
    }
    free(a); free(b);
    return 0;
}

/* ---- src/chorin_fd/simulate.py:93-167 (_semi_implicit_predictor_step) -------------
 * Quirks kept: diagonal (2/nu)*dx^2 + 2dt (:108,117), vC scaled by dx^2 (:150), all
 * four solves along axis 0 (:137,153,159,165) => needs nx == ny.                     */
int orc_chorin_semi_implicit_predictor(const double *un, const double *vn, const double *un1,
                                       const double *vn1, double *ui, double *vi, int nx, int ny,
                                       double dt, double dx, double dy, double nu) {
    if (nx != ny) return -2;
    const int n = nx - 2, m = ny - 2;
    const double dx2 = py_sq(dx), dy2 = py_sq(dy);
    const double twodx = 2 * dx, twody = 2 * dy;
    const double kx = 2 / nu * dx2, ky = 2 / nu * dy2;
    const double diagA = kx + 2 * dt, diagB = ky + 2 * dt;
    const size_t nm = (size_t)n * m;
    double *uC = (double *)malloc(sizeof(double) * nm), *vC = (double *)malloc(sizeof(double) * nm);
    double *ut = (double *)malloc(sizeof(double) * nm), *vt = (double *)malloc(sizeof(double) * nm);
    double *uS = (double *)malloc(sizeof(double) * nm), *vS = (double *)malloc(sizeof(double) * nm);
    double *uo = (double *)malloc(sizeof(double) * nm), *vo = (double *)malloc(sizeof(double) * nm);
    memcpy(ui, un, sizeof(double) * (size_t)nx * ny);
    memcpy(vi, vn, sizeof(double) * (size_t)nx * ny);
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j) {
            const size_t q = (size_t)(i - 1) * m + (j - 1);
            const double uc = un[IDX(i, j)], vc = vn[IDX(i, j)];
            const double u1c = un1[IDX(i, j)], v1c = vn1[IDX(i, j)];
            const double uHn = uc * (un[IDX(i + 1, j)] - un[IDX(i - 1, j)]) / twodx +
                               vc * (un[IDX(i, j + 1)] - un[IDX(i, j - 1)]) / twody;
            const double uHn1 = u1c * (un1[IDX(i + 1, j)] - un1[IDX(i - 1, j)]) / twodx +
                                v1c * (un1[IDX(i, j + 1)] - un1[IDX(i, j - 1)]) / twody;
            const double uC1 = dt / 2. * (3 * uHn - uHn1);
            const double uC2 = dt * nu * ((un[IDX(i + 1, j)] - 2 * uc + un[IDX(i - 1, j)]) / dx2 +
                                          (un[IDX(i, j + 1)] - 2 * uc + un[IDX(i, j - 1)]) / dy2);
            uC[q] = kx * (uC1 + uC2);
            const double vHn = uc * (vn[IDX(i + 1, j)] - vn[IDX(i - 1, j)]) / twodx +
                               vc * (vn[IDX(i, j + 1)] - vn[IDX(i, j - 1)]) / twody;
            const double vHn1 = u1c * (vn1[IDX(i + 1, j)] - vn1[IDX(i - 1, j)]) / twodx +
                                v1c * (vn1[IDX(i, j + 1)] - vn1[IDX(i, j - 1)]) / twody;
            const double vC1 = dt / 2. * (3 * vHn - vHn1);
            const double vC2 = dt * nu * ((vn[IDX(i + 1, j)] - 2 * vc + vn[IDX(i - 1, j)]) / dx2 +
                                          (vn[IDX(i, j + 1)] - 2 * vc + vn[IDX(i, j - 1)]) / dy2);
            vC[q] = kx * (vC1 + vC2);
        }
    int rc = dense_solve(n, m, diagA, -dt, uC, ut);
    if (!rc) rc = dense_solve(n, m, diagA, -dt, vC, vt);
    if (!rc) {
        for (int i = 1; i < nx - 1; ++i)
            for (int j = 1; j < ny - 1; ++j) {
                const size_t q = (size_t)(i - 1) * m + (j - 1);
                const double uc = un[IDX(i, j)], vc = vn[IDX(i, j)];
                uS[q] = ky * (ut[q] + uc) - dt * (un[IDX(i, j + 1)] - 2 * uc + un[IDX(i, j - 1)]);
                vS[q] = ky * (vt[q] + vc) - dt * (vn[IDX(i, j + 1)] - 2 * vc + vn[IDX(i, j - 1)]);
            }
        rc = dense_solve(m, m, diagB, -dt, uS, uo);   /* B is (ny-2)^2, applied along axis 0 */
        if (!rc) rc = dense_solve(m, m, diagB, -dt, vS, vo);
    }
    if (!rc)
        for (int i = 1; i < nx - 1; ++i)
            for (int j = 1; j < ny - 1; ++j) {
                const size_t q = (size_t)(i - 1) * m + (j - 1);
                ui[IDX(i, j)] = uo[q];
                vi[IDX(i, j)] = vo[q];
            }
    free(uC); free(vC); free(ut); free(vt); free(uS); free(vS); free(uo); free(vo);
    return rc;
}

/* ---- src/chorin_fd/simulate.py:169-202 (_get_pressure) ----------------------------
 * Lexicographic in-place SOR, at most nit-1 sweeps, err = max|p - pPrev| <= 5e-6 exits.
 * Returns the number of sweeps executed.                                             */
int orc_chorin_pressure(const double *ui, const double *vi, double *p, int nx, int ny, int nit,
                        double dt, double dx, double dy, double rho, double beta) {
    const double tol = 5e-6;
    const double dx2 = py_sq(dx), dy2 = py_sq(dy);
    const size_t N = (size_t)nx * ny;
    double *C = (double *)calloc(N, sizeof(double));
    double *pPrev = (double *)malloc(sizeof(double) * N);
    memcpy(pPrev, p, sizeof(double) * N);
    const double cu = dx * rho * dy2 / dt, cv = dy * rho * dx2 / dt;
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j)
            C[IDX(i, j)] = cu * (ui[IDX(i, j)] - ui[IDX(i - 1, j)]) + cv * (vi[IDX(i, j)] - vi[IDX(i, j - 1)]);
    const double den = 2 * dx2 + 2 * dy2, omb = 1 - beta;
    double err = 1;
    int it = 1, sweeps = 0;
    while (err > tol && it < nit) {
        for (int i = 1; i < nx - 1; ++i)
            for (int j = 1; j < ny - 1; ++j)
                p[IDX(i, j)] = beta * (dy2 * p[IDX(i + 1, j)] + dy2 * p[IDX(i - 1, j)] +
                                       dx2 * p[IDX(i, j + 1)] + dx2 * p[IDX(i, j - 1)] - C[IDX(i, j)]) / den +
                               omb * p[IDX(i, j)];
        err = 0;   /* np.max(np.abs(p - pPrev)) over the full array (edges contribute 0) */
        for (size_t q = 0; q < N; ++q) {
            const double d = fabs(p[q] - pPrev[q]);
            if (d > err || d != d) err = d;
        }
        memcpy(pPrev, p, sizeof(double) * N);
        ++it; ++sweeps;
    }
    free(C); free(pPrev);
    return sweeps;
}

/* ---- src/chorin_fd/simulate.py:204-210 (_correction_step) ------------------------- */
void orc_chorin_correction(const double *ui, const double *vi, const double *p, double *u, double *v,
                           int nx, int ny, double dt, double dx, double dy) {
    const double kx = dt / (2 * dx), ky = dt / (2 * dy);
    memcpy(u, ui, sizeof(double) * (size_t)nx * ny);
    memcpy(v, vi, sizeof(double) * (size_t)nx * ny);
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j) {
            u[IDX(i, j)] = ui[IDX(i, j)] - kx * (p[IDX(i + 1, j)] - p[IDX(i - 1, j)]);
            v[IDX(i, j)] = vi[IDX(i, j)] - ky * (p[IDX(i, j + 1)] - p[IDX(i, j - 1)]);
        }
}

typedef struct {
    int nx, ny, nit, method;   /* method: 0 explicit, 1 semi_implicit */
    double dt, rho, nu, beta;
    int n_ubc, n_vbc, n_pbc;
} orc_chorin_params;

/* ---- src/chorin_fd/simulate.py:212-234 (step): u_out/v_out are new arrays, p in place */
int orc_chorin_step(const orc_chorin_params *P, const orc_bc *ubc, const orc_bc *vbc, const orc_bc *pbc,
                    const double *un, const double *vn, const double *un1, const double *vn1, double *p,
                    double *u_out, double *v_out, int *sweeps) {
    const int nx = P->nx, ny = P->ny;
    const double dx = 2. / (nx - 1), dy = 2. / (ny - 1);
    const size_t N = (size_t)nx * ny;
    double *ui = (double *)malloc(sizeof(double) * N), *vi = (double *)malloc(sizeof(double) * N);
    int rc = 0;
    if (P->method == 0) orc_chorin_explicit_predictor(un, vn, un1, vn1, ui, vi, nx, ny, P->dt, dx, dy, P->nu);
    else rc = orc_chorin_semi_implicit_predictor(un, vn, un1, vn1, ui, vi, nx, ny, P->dt, dx, dy, P->nu);
    if (!rc) {
        orc_bc_apply_list(ui, nx, ny, ubc, P->n_ubc, dx, dy);
        orc_bc_apply_list(vi, nx, ny, vbc, P->n_vbc, dx, dy);
        const int s = orc_chorin_pressure(ui, vi, p, nx, ny, P->nit, P->dt, dx, dy, P->rho, P->beta);
        if (sweeps) *sweeps = s;
        orc_bc_apply_list(p, nx, ny, pbc, P->n_pbc, dx, dy);
        orc_chorin_correction(ui, vi, p, u_out, v_out, nx, ny, P->dt, dx, dy);
    }
    free(ui); free(vi);
    return rc;
}

/* ---- src/chorin_fd/simulate.py:236-271 (_init_variables + simulate) ---------------
 * traj_* may be NULL; otherwise [nt][nx][ny].  state_* (optional) receive the final
 * (u, v, u1, v1, p).  sweeps (optional) is [nt].                                     */
int orc_chorin_simulate(const orc_chorin_params *P, const orc_bc *ubc, const orc_bc *vbc, const orc_bc *pbc,
                        const double *u_ic, const double *v_ic, const double *p_ic, int nt,
                        double *traj_u, double *traj_v, double *traj_p, int *sweeps,
                        double *fin_u, double *fin_v, double *fin_u1, double *fin_v1, double *fin_p) {
    const int nx = P->nx, ny = P->ny;
    const double dx = 2. / (nx - 1), dy = 2. / (ny - 1);
    const size_t N = (size_t)nx * ny, B = sizeof(double) * N;
    double *u = (double *)malloc(B), *v = (double *)malloc(B), *p = (double *)malloc(B);
    double *u1 = (double *)malloc(B), *v1 = (double *)malloc(B);
    double *un = (double *)malloc(B), *vn = (double *)malloc(B);
    memcpy(u, u_ic, B); memcpy(v, v_ic, B); memcpy(p, p_ic, B);
    orc_bc_apply_list(u, nx, ny, ubc, P->n_ubc, dx, dy);
    orc_bc_apply_list(v, nx, ny, vbc, P->n_vbc, dx, dy);
    orc_bc_apply_list(p, nx, ny, pbc, P->n_pbc, dx, dy);
    memcpy(u1, u, B); memcpy(v1, v, B);
    int rc = 0;
    for (int n = 0; n < nt && !rc; ++n) {
        rc = orc_chorin_step(P, ubc, vbc, pbc, u, v, u1, v1, p, un, vn, sweeps ? &sweeps[n] : NULL);
        double *t;
        t = u1; u1 = u; u = un; un = t;
        t = v1; v1 = v; v = vn; vn = t;
        if (traj_u) memcpy(traj_u + (size_t)n * N, u, B);
        if (traj_v) memcpy(traj_v + (size_t)n * N, v, B);
        if (traj_p) memcpy(traj_p + (size_t)n * N, p, B);
    }
    if (fin_u) memcpy(fin_u, u, B);
    if (fin_v) memcpy(fin_v, v, B);
    if (fin_u1) memcpy(fin_u1, u1, B);
    if (fin_v1) memcpy(fin_v1, v1, B);
    if (fin_p) memcpy(fin_p, p, B);
    free(u); free(v); free(p); free(u1); free(v1); free(un); free(vn);
    return rc;
}

/* Ensemble of independent chorin_fd simulations (no reference counterpart; every member
 * is one orc_chorin_simulate).  Per-member nu[b] and BC values: bc arrays are laid out
 * [batch][n_bc].  Fields are [batch][nx][ny]; state is advanced IN PLACE by nt steps
 * (u1/v1 carry the previous step, as in simulate()); no initial BC application here.
 * Runs members in parallel with OpenMP when available.  Returns threads used (>0).   */
int orc_chorin_ensemble_run(const orc_chorin_params *P, const double *nu, const orc_bc *ubc,
                            const orc_bc *vbc, const orc_bc *pbc, int batch, int nt, double *u, double *v,
                            double *u1, double *v1, double *p, int *sweeps /* [nt][batch] or NULL */) {
    const size_t N = (size_t)P->nx * P->ny, Bb = sizeof(double) * N;
    int used = 1;
#ifdef _OPENMP
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < batch; ++b) {
        orc_chorin_params Q = *P;
        if (nu) Q.nu = nu[b];
        double *un = (double *)malloc(Bb), *vn = (double *)malloc(Bb);
        double *ub = u + b * N, *vb = v + b * N, *u1b = u1 + b * N, *v1b = v1 + b * N, *pb = p + b * N;
        for (int n = 0; n < nt; ++n) {
            int s = 0;
            orc_chorin_step(&Q, ubc + (size_t)b * P->n_ubc, vbc + (size_t)b * P->n_vbc,
                            pbc + (size_t)b * P->n_pbc, ub, vb, u1b, v1b, pb, un, vn, &s);
            if (sweeps) sweeps[(size_t)n * batch + b] = s;
            memcpy(u1b, ub, Bb); memcpy(v1b, vb, Bb);
            memcpy(ub, un, Bb); memcpy(vb, vn, Bb);
        }
        free(un); free(vn);
    }
    return used;
}

/* ================================ direct_fd ======================================== */

/* ---- src/direct_fd/simulate.py:56-66 (_build_up_b): axis 1 <-> dx, axis 0 <-> dy --- */
void orc_direct_build_b(const double *u, const double *v, double *b, int nx, int ny, double dt,
                        double dx, double dy, double rho) {
    memset(b, 0, sizeof(double) * (size_t)nx * ny);
    const double twodx = 2 * dx, twody = 2 * dy;
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j) {
            const double ux = (u[IDX(i, j + 1)] - u[IDX(i, j - 1)]) / twodx;
            const double vy = (v[IDX(i + 1, j)] - v[IDX(i - 1, j)]) / twody;
            const double uy = (u[IDX(i + 1, j)] - u[IDX(i - 1, j)]) / twody;
            const double t1 = rho * (1 / dt * (ux + vy));
            const double t2 = ux * ux;                     /* numpy: x**2 on arrays == x*x */
            const double t3 = 2 * (uy * (v[IDX(i, j + 1)] - v[IDX(i, j - 1)]) / twodx);
            const double t4 = vy * vy;
            b[IDX(i, j)] = t1 - t2 - t3 - t4;
        }
}

/* ---- src/direct_fd/simulate.py:68-88 (_pressure_poisson): exactly nit Jacobi sweeps,
 * p BCs re-applied after every sweep (:85-86).                                        */
void orc_direct_pressure(double *p, const double *b, int nx, int ny, int nit, double dx, double dy,
                         const orc_bc *pbc, int n_pbc) {
    const size_t N = (size_t)nx * ny;
    double *pn = (double *)malloc(sizeof(double) * N);
    const double dx2 = py_sq(dx), dy2 = py_sq(dy);
    const double den = 2 * (dx2 + dy2);
    const double kb = dx2 * dy2 / den;
    for (int q = 0; q < nit; ++q) {
        memcpy(pn, p, sizeof(double) * N);
        for (int i = 1; i < nx - 1; ++i)
            for (int j = 1; j < ny - 1; ++j)
                p[IDX(i, j)] = ((pn[IDX(i, j + 1)] + pn[IDX(i, j - 1)]) * dy2 +
                                (pn[IDX(i + 1, j)] + pn[IDX(i - 1, j)]) * dx2) / den -
                               kb * b[IDX(i, j)];
        orc_bc_apply_list(p, nx, ny, pbc, n_pbc, dx, dy);
    }
    free(pn);
}

typedef struct {
    int nx, ny, nit;
    double dt, rho, nu;
    int n_ubc, n_vbc, n_pbc;
} orc_direct_params;

/* ---- src/direct_fd/simulate.py:90-127 (step): u, v, p updated in place ------------- */
void orc_direct_step(const orc_direct_params *P, const orc_bc *ubc, const orc_bc *vbc, const orc_bc *pbc,
                     double *u, double *v, double *p) {
    const int nx = P->nx, ny = P->ny;
    const double dx = 2. / (nx - 1), dy = 2. / (ny - 1);
    const double dt = P->dt, rho = P->rho, nu = P->nu;
    const size_t N = (size_t)nx * ny;
    double *un = (double *)malloc(sizeof(double) * N), *vn = (double *)malloc(sizeof(double) * N);
    double *b = (double *)malloc(sizeof(double) * N);
    memcpy(un, u, sizeof(double) * N); memcpy(vn, v, sizeof(double) * N);
    orc_direct_build_b(u, v, b, nx, ny, dt, dx, dy, rho);
    orc_direct_pressure(p, b, nx, ny, P->nit, dx, dy, pbc, P->n_pbc);
    const double dx2 = py_sq(dx), dy2 = py_sq(dy);
    const double kpx = dt / (2 * rho * dx), kpy = dt / (2 * rho * dy);
    const double kdx = dt / dx2, kdy = dt / dy2;
    for (int i = 1; i < nx - 1; ++i)
        for (int j = 1; j < ny - 1; ++j) {
            const double uc = un[IDX(i, j)], vc = vn[IDX(i, j)];
            u[IDX(i, j)] = uc - uc * dt / dx * (uc - un[IDX(i, j - 1)]) -
                           vc * dt / dy * (uc - un[IDX(i - 1, j)]) -
                           kpx * (p[IDX(i, j + 1)] - p[IDX(i, j - 1)]) +
                           nu * (kdx * (un[IDX(i, j + 1)] - 2 * uc + un[IDX(i, j - 1)]) +
                                 kdy * (un[IDX(i + 1, j)] - 2 * uc + un[IDX(i - 1, j)]));
            v[IDX(i, j)] = vc - uc * dt / dx * (vc - vn[IDX(i, j - 1)]) -
                           vc * dt / dy * (vc - vn[IDX(i - 1, j)]) -
                           kpy * (p[IDX(i + 1, j)] - p[IDX(i - 1, j)]) +
                           nu * (kdx * (vn[IDX(i, j + 1)] - 2 * vc + vn[IDX(i, j - 1)]) +
                                 kdy * (vn[IDX(i + 1, j)] - 2 * vc + vn[IDX(i - 1, j)]));
        }
    orc_bc_apply_list(u, nx, ny, ubc, P->n_ubc, dx, dy);
    orc_bc_apply_list(v, nx, ny, vbc, P->n_vbc, dx, dy);
    free(un); free(vn); free(b);
}

/* ---- src/direct_fd/simulate.py:129-144 (simulate): state advanced IN PLACE (the
 * reference aliases the caller's IC arrays, :132), no BC pass before step 0.          */
void orc_direct_simulate(const orc_direct_params *P, const orc_bc *ubc, const orc_bc *vbc, const orc_bc *pbc,
                         double *u, double *v, double *p, int nt, double *traj_u, double *traj_v,
                         double *traj_p) {
    const size_t N = (size_t)P->nx * P->ny, B = sizeof(double) * N;
    for (int n = 0; n < nt; ++n) {
        orc_direct_step(P, ubc, vbc, pbc, u, v, p);
        if (traj_u) memcpy(traj_u + (size_t)n * N, u, B);
        if (traj_v) memcpy(traj_v + (size_t)n * N, v, B);
        if (traj_p) memcpy(traj_p + (size_t)n * N, p, B);
    }
}

int orc_direct_ensemble_run(const orc_direct_params *P, const double *nu, const orc_bc *ubc,
                            const orc_bc *vbc, const orc_bc *pbc, int batch, int nt, double *u, double *v,
                            double *p) {
    const size_t N = (size_t)P->nx * P->ny;
    int used = 1;
#ifdef _OPENMP
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < batch; ++b) {
        orc_direct_params Q = *P;
        if (nu) Q.nu = nu[b];
        orc_direct_simulate(&Q, ubc + (size_t)b * P->n_ubc, vbc + (size_t)b * P->n_vbc,
                            pbc + (size_t)b * P->n_pbc, u + b * N, v + b * N, p + b * N, nt, NULL, NULL, NULL);
    }
    return used;
}

/* Set the OpenMP team size explicitly (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers). */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
