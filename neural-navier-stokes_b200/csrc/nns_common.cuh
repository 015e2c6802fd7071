// nns_common.cuh -- shared host/device declarations of libnns_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nns_b200.h"

namespace nns {

// Ordered BC list of one field, passed to kernels by value.
// Semantics: src/boundary.py:34-48 (Dirichlet) and :56-86 (Neumann) of the reference.
struct BcList {
    int n;
    int side[NNS_MAX_BC];
    int type[NNS_MAX_BC];
    int slot[NNS_MAX_BC];   // index into the per-member value table
    double value[NNS_MAX_BC];
};

struct Geometry {
    int nx, ny, batch, nit, method;
    double dt, rho, nu, beta, tol, dx, dy;
};

void set_error(const char *fmt, ...);

}  // namespace nns

struct nns_handle {
    nns::Geometry g;
    nns_params params;
    nns::BcList bc[3];        // u, v, p
    int n_bcs;                // total entries (per-member value table width)
    double *d_nu;             // [batch] or nullptr
    double *d_bcval;          // [batch][n_bcs] or nullptr
    double *h_nu, *h_bcval;   // host copies of the two (the tiled path steps the members of a batch one after the other)
    double *d_scratch[4];     // rotation / ui,vi workspace, [batch][nx][ny] each, lazily allocated
    double *d_stage[7];       // device staging of the host-buffer entry points
    void *d_pool[10];         // cached device buffers of the *_run_host entry points (fields, trajectories, sweeps)
    size_t pool_bytes[10];
    int scratch_slot;         // chorin_fd stream path: which of its 4 per-CTA scratch sets the next launch uses (one per internal stream)
    cudaStream_t streams[4];  // copy/compute pipelining of the host-buffer entry points
    void *chip_plan;          // chorin_fd chip path: cached ChipPlan (host) and block table (device)
    void *d_blockdesc;
    double *d_cprime;         // SOR right-hand side when it does not fit in shared memory
    void *stream_plan;        // chorin_fd persistent stream path: block tables + C' images
    void *slab;               // chorin_fd tiled / row-slab path: SlabState (partition, NCCL communicator, scratch)
    void *spectral;           // chorin_spectral: SpectralPlan (operators + workspace)
    double *d_b;              // direct_fd rhs / second p buffer
    double *d_p2;
    int32_t *d_sweeps;        // [batch] scratch
    unsigned long long *d_nonfinite;
    int device;
    int sm_count;
    int max_smem_optin;
    int64_t launches;
};

#define NNS_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            nns::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                           cudaGetErrorString(e__));                                           \
            return NNS_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

namespace nns {

// ---- device helpers -------------------------------------------------------------------

// Sequential application of one BC list to a row-major field in GLOBAL memory by the whole
// CTA (one member per CTA).  Barriers between entries keep the reference's list order:
// a Neumann edge reads the neighbouring line including corner cells that an earlier entry
// may already have written (boundary.py:73-84).
__device__ __forceinline__ void cta_apply_bc_global(double *A, int nx, int ny, const BcList &L,
                                                    const double *bcval, double dx, double dy) {
    for (int k = 0; k < L.n; ++k) {
        const double g = bcval ? bcval[L.slot[k]] : L.value[k];
        const int side = L.side[k];
        const bool neu = L.type[k] == NNS_BC_NEUMANN;
        if (side == NNS_SIDE_LEFT || side == NNS_SIDE_RIGHT) {
            const int i = side == NNS_SIDE_LEFT ? 0 : nx - 1;
            const int in = side == NNS_SIDE_LEFT ? 1 : nx - 2;
            const double sgn = side == NNS_SIDE_LEFT ? -dx : dx;
            for (int j = threadIdx.x; j < ny; j += blockDim.x)
                A[(size_t)i * ny + j] = neu ? A[(size_t)in * ny + j] + sgn * g : g;
        } else {
            const int j = side == NNS_SIDE_BOTTOM ? 0 : ny - 1;
            const int jn = side == NNS_SIDE_BOTTOM ? 1 : ny - 2;
            const double sgn = side == NNS_SIDE_BOTTOM ? -dy : dy;
            for (int i = threadIdx.x; i < nx; i += blockDim.x)
                A[(size_t)i * ny + j] = neu ? A[(size_t)i * ny + jn] + sgn * g : g;
        }
        __syncthreads();
    }
}

}  // namespace nns
