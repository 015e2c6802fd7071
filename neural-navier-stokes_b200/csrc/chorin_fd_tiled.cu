// chorin_fd_tiled.cu -- chorin_fd for grids whose pressure field does not fit in one SM
// (placeholder until the skewed-tile path lands; fails loudly, there is no CPU fallback).
#include "nns_common.cuh"

namespace nns {

int chorin_tiled_run(nns_handle *h, double *bufU[3], double *bufV[3], double *p, int nsteps, int nsteps_total,
                     int step0, int phases, int fixup, double *tu, double *tv, double *tp, int32_t *sweeps,
                     cudaStream_t st, int m0, int count) {
    (void)m0; (void)count;
    (void)bufU; (void)bufV; (void)p; (void)nsteps; (void)nsteps_total; (void)step0; (void)phases; (void)fixup;
    (void)tu; (void)tv; (void)tp; (void)sweeps; (void)st;
    set_error("chorin_fd: grid %dx%d does not fit the on-chip path and the tiled path is not built yet", h->g.nx,
              h->g.ny);
    return NNS_ERR_UNSUPPORTED;
}

}  // namespace nns
