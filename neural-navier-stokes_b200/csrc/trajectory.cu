// trajectory.cu -- the sink of the time-step path: turns the device trajectories the solvers write
// ([members][nt][nx][ny] float64 u, v, p) into the two formats the rest of the reference consumes, without a
// round trip through the host:
//
//   * spatial_coarsen (src/utils.py:13-60): block means over agg_x x agg_y cells, float64 [nt][nx/agg_x][ny/agg_y]
//     per field.  The mean of a block is computed in NumPy's own order -- np.mean over the flattened block =
//     pairwise_sum with eight running accumulators over the row-major block, then one division by the count -- so
//     the result is bit-identical to the reference's.  The reference's inner loop runs over ny // agg_x columns
//     (utils.py:49): output columns beyond that stay zero, like there.
//   * the observation tensor of the neural scripts (src/neural_spectral/rnn.py:77-82, spectral_ode.py:158-163):
//     float32 [nt][3][nx'][ny'] = stack([u, v, p]).permute(1, 0, 2, 3) after .float(), optionally of the coarsened
//     fields.
//
// HBM-bound: 24 B read per fine cell, 24 / (agg_x agg_y) B (float64) or 12 / (agg_x agg_y) B (float32) written.
// One thread per output cell and field triple; the threads of a warp cover adjacent blocks of one block row, so a
// warp reads agg_y * 32 consecutive doubles per fine row.
#include <algorithm>

#include "nns_common.cuh"

namespace nns {

namespace {

constexpr int TRAJ_MAX_BLOCK = 128;      // NumPy sums up to 128 elements in one unrolled pass (PW_BLOCKSIZE)

// np.add.reduce over n contiguous doubles (numpy/core/src/umath/loops_utils.h.src, pairwise_sum), n <= 128
template <typename F>
__device__ __forceinline__ double numpy_pairwise_sum(int n, F at) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, at(i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = at(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], at(i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, at(i));
    return res;
}

struct TrajArgs {
    const double *u, *v, *p;      // [members][nt][nx][ny]
    void *ou, *ov, *op;           // float64: three arrays [members][nt][cx][cy]; float32: ou = [members][nt][3][cx][cy]
    long long frames;             // members * nt
    int nx, ny, ax, ay, cx, cy, jlim;
};

template <bool F32>
__global__ void __launch_bounds__(256) traj_pack_kernel(const TrajArgs a) {
    const long long cells = (long long)a.cx * a.cy;
    const long long total = a.frames * cells;
    const int n = a.ax * a.ay;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const long long f = q / cells;
        const int c = (int)(q - f * cells), ci = c / a.cy, cj = c - ci * a.cy;
        const size_t base = (size_t)f * a.nx * a.ny + (size_t)ci * a.ax * a.ny + (size_t)cj * a.ay;
        double m[3] = {0.0, 0.0, 0.0};
        if (cj < a.jlim) {
            const double *src[3] = {a.u, a.v, a.p};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double *s = src[k] + base;
                const int ay = a.ay, ny = a.ny;
                const double sum = numpy_pairwise_sum(n, [&](int e) { return s[(size_t)(e / ay) * ny + (e % ay)]; });
                m[k] = n == 1 ? sum : __ddiv_rn(sum, (double)n);
            }
        }
        if (F32) {
            float *o = static_cast<float *>(a.ou) + (size_t)f * 3 * cells + c;
            o[0] = (float)m[0]; o[cells] = (float)m[1]; o[2 * cells] = (float)m[2];
        } else {
            static_cast<double *>(a.ou)[q] = m[0];
            static_cast<double *>(a.ov)[q] = m[1];
            static_cast<double *>(a.op)[q] = m[2];
        }
    }
}

int traj_launch(const double *u, const double *v, const double *p, long long frames, int nx, int ny, int ax, int ay,
                void *ou, void *ov, void *op, bool f32, cudaStream_t st) {
    if (!u || !v || !p || !ou || (!f32 && (!ov || !op)) || frames < 0 || nx <= 0 || ny <= 0 || ax <= 0 || ay <= 0) {
        set_error("nns_traj: bad argument");
        return NNS_ERR_INVALID;
    }
    if (nx % ax || ny % ay) { set_error("nns_traj: the grid is not a multiple of the coarsening factors (utils.py:39-40)"); return NNS_ERR_INVALID; }
    if (ax * ay > TRAJ_MAX_BLOCK) { set_error("nns_traj: blocks of more than %d cells are not supported", TRAJ_MAX_BLOCK); return NNS_ERR_UNSUPPORTED; }
    if (ny / ax > ny / ay) { set_error("nns_traj: ny // agg_x > ny // agg_y indexes past the output (IndexError in utils.py:55)"); return NNS_ERR_INVALID; }
    if (frames == 0) return NNS_OK;
    TrajArgs a{u, v, p, ou, ov, op, frames, nx, ny, ax, ay, nx / ax, ny / ay, ny / ax};
    const long long total = frames * a.cx * a.cy;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (total + 255) / 256;
    const int grid = (int)std::min<long long>(want, (long long)sms * 8);        // 8 resident CTAs of 256 threads per SM
    if (f32) traj_pack_kernel<true><<<grid, 256, 0, st>>>(a);
    else traj_pack_kernel<false><<<grid, 256, 0, st>>>(a);
    NNS_CUDA(cudaGetLastError());
    return NNS_OK;
}

}  // namespace

}  // namespace nns

extern "C" {

int32_t nns_traj_coarsen(const double *u, const double *v, const double *p, int64_t frames, int32_t nx, int32_t ny,
                         int32_t agg_x, int32_t agg_y, double *u_out, double *v_out, double *p_out, void *stream) {
    return nns::traj_launch(u, v, p, frames, nx, ny, agg_x, agg_y, u_out, v_out, p_out, false, (cudaStream_t)stream);
}

int32_t nns_traj_observations(const double *u, const double *v, const double *p, int64_t frames, int32_t nx, int32_t ny,
                              int32_t agg_x, int32_t agg_y, float *obs_out, void *stream) {
    return nns::traj_launch(u, v, p, frames, nx, ny, agg_x, agg_y, obs_out, nullptr, nullptr, true, (cudaStream_t)stream);
}

}  // extern "C"
