// optimize_kernels.cuh -- device half of the marker optimisation: particle_compute_dist_pertb_abs_v
// (/root/reference/src/pic1dp_particle.F90:356-403), the O(N) reduction  dist[iv] = sum |w| * linear weight in v
// that particle_merge / particle_remove / particle_split read.  16 B/marker (v, w), HBM-bound; the nv-point grid
// (input_nv = 128) lives in one private shared-memory copy per warp, lanes that hit the same v cell are ordered
// with MATCH.ANY (Depositor<DEP_WARP_PRIVATE>), warp grids are summed in warp order and CTA grids in CTA order:
// no atomics, bitwise run-to-run deterministic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "particle_kernels.cuh"

namespace pic1dp {

struct DistArgs {
  const double *v, *w;
  int64_t np;
  int nv;
  double v_max;
  double *partial;  // [gridDim.x][nv]
};

__global__ void __launch_bounds__(512) k_dist_pertb_abs_v(const DistArgs a) {
  extern __shared__ __align__(16) double sh[];
  const int nw = blockDim.x >> 5, nvp = a.nv + 1;  // one spare cell: v just below v_max can round to iv + 1 == nv
  for (int j = threadIdx.x; j < nw * nvp; j += blockDim.x) sh[j] = 0.0;
  __syncthreads();
  Depositor<DEP_WARP_PRIVATE> dep;
  dep.g = sh + (size_t)(threadIdx.x >> 5) * nvp;
  const double two_vmax = dmul(a.v_max, 2.0), rnv = (double)(a.nv - 1);
  // warp-uniform trip count: every lane of a warp calls dep.add together
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.np; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const bool in = i < a.np;
    const double v = in ? __ldcs(a.v + i) : 0.0, w = in ? __ldcs(a.w + i) : 0.0;
    const bool ok = in && fabs(v) < a.v_max;  // "ignore too fast particle" (:380)
    double sv = dmul(ddiv(dadd(v, a.v_max), two_vmax), rnv);  // :382-383
    const int iv = ok ? __double2int_rd(sv) : 0;
    sv = dsub(1.0, dsub(sv, (double)iv));  // :385
    const double aw = fabs(w);
    dep.add(iv, iv + 1, dmul(sv, aw), dmul(dsub(1.0, sv), aw), ok);  // :387-389
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.nv; j += blockDim.x) {
    double t = sh[j];
    for (int q = 1; q < nw; q++) t = dadd(t, sh[(size_t)q * nvp + j]);
    a.partial[(size_t)blockIdx.x * a.nv + j] = t;
  }
}

// dist[j] = sum over CTA grids in CTA order
__global__ void k_dist_final(const double *partial, int ngrids, int nv, double *dist) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nv) return;
  double t = partial[j];
  for (int k = 1; k < ngrids; k++) t = dadd(t, partial[(size_t)k * nv + j]);
  dist[j] = t;
}

}  // namespace pic1dp
