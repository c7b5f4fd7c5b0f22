// diag_kernels.cuh -- on-device diagnostics of pic1dp_output, so that an output step moves a few KB to the host
// instead of every marker (the reference reads x, v, p, w on the host at every output,
// /root/reference/src/pic1dp_output.F90:128-150, :228-237; at 1e8 markers that is 3.2 GB per output).
//
//  k_diag<SUMS, HIST>:
//    SUMS: sum v^2, sum v^2 p, sum v^2 w per species (output_field, :126-172; VecPointwiseMult + VecSum)
//    HIST: bilinear x-v histograms of g (markers), f (p-weighted) and delta f (w-weighted) on an nx_opd x nv_opd
//          grid (output_ptcldist, :239-313), skipping |v| >= v_max (:241).  The v-only histograms of the reference
//          (:297-312) are the x-sums of these (sx + (1-sx) = 1) and are formed from them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "particle_kernels.cuh"

namespace pic1dp {

struct DiagArgs {
  const double *x, *v, *p, *w;
  int64_t np;
  int deltaf;
  // sums
  double *sum_partial;  // [gridDim.x][3]
  // histogram
  int nx_opd, nv_opd, ncopies;
  double lx, v_max;
  double *hist;         // [ncopies][3][nv_opd*nx_opd], accumulated with RED.ADD.F64 (L2 resident)
};

template <bool SUMS, bool HIST>
__global__ void __launch_bounds__(512) k_diag(const DiagArgs a) {
  __shared__ double s_red[3][16];
  double s_vv = 0.0, s_vvp = 0.0, s_vvw = 0.0;
  double *hm = nullptr, *ht = nullptr, *hp = nullptr;
  const int ncell = a.nx_opd * a.nv_opd;
  if (HIST) {
    hm = a.hist + (size_t)(blockIdx.x % a.ncopies) * 3 * ncell;
    ht = hm + ncell;
    hp = ht + ncell;
  }
  const double rnx = (double)a.nx_opd, rnv = (double)(a.nv_opd - 1), two_vmax = dmul(a.v_max, 2.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ldcs(a.v + i), p = __ldcs(a.p + i);
    const double w = a.deltaf ? __ldcs(a.w + i) : 0.0;
    if (SUMS) {
      const double vv = dmul(v, v);  // VecPointwiseMult(tmp1, v, v)  :128
      s_vv = dadd(s_vv, vv);
      s_vvp = dadd(s_vvp, dmul(vv, p));  // :138
      if (a.deltaf) s_vvw = dadd(s_vvw, dmul(vv, w));  // :147
    }
    if (HIST) {
      if (fabs(v) >= a.v_max) continue;  // :241
      const double x = __ldcs(a.x + i);
      double sx = dmul(ddiv(x, a.lx), rnx);  // :243
      int ix = __double2int_rd(sx);
      sx = dsub(1.0, dsub(sx, (double)ix));  // :245
      double sv = dmul(ddiv(dadd(v, a.v_max), two_vmax), rnv);  // :247-248
      const int iv = __double2int_rd(sv);
      sv = dsub(1.0, dsub(sv, (double)iv));  // :250
      if ((unsigned)ix >= (unsigned)a.nx_opd) {  // x == lx exactly (the reference would index out of bounds)
        ix = 0;
        sx = 1.0;
      }
      int ix2 = ix + 1;
      if (ix2 > a.nx_opd - 1) ix2 = 0;  // :272
      const double sx2 = dsub(1.0, sx), sv2 = dsub(1.0, sv);  // :273
      const int c00 = iv * a.nx_opd + ix, c10 = (iv + 1) * a.nx_opd + ix;
      const int c01 = iv * a.nx_opd + ix2, c11 = (iv + 1) * a.nx_opd + ix2;
      const double w00 = dmul(sx, sv), w10 = dmul(sx, sv2), w01 = dmul(sx2, sv), w11 = dmul(sx2, sv2);
      // iv + 1 == nv_opd happens when (v + v_max) / (2 v_max) rounds to 1 (v within an ulp of v_max): the reference
      // writes out of bounds there; those two contributions (weight sv2 ~ 0) are dropped
      const bool up = iv + 1 < a.nv_opd;
      atomicAdd(hm + c00, w00);  // results unused: RED.E.ADD.F64
      atomicAdd(hm + c01, w01);
      atomicAdd(ht + c00, dmul(w00, p));
      atomicAdd(ht + c01, dmul(w01, p));
      if (up) {
        atomicAdd(hm + c10, w10);
        atomicAdd(hm + c11, w11);
        atomicAdd(ht + c10, dmul(w10, p));
        atomicAdd(ht + c11, dmul(w11, p));
      }
      if (a.deltaf) {
        atomicAdd(hp + c00, dmul(w00, w));
        atomicAdd(hp + c01, dmul(w01, w));
        if (up) {
          atomicAdd(hp + c10, dmul(w10, w));
          atomicAdd(hp + c11, dmul(w11, w));
        }
      }
    }
  }
  if (SUMS) {  // fixed-shape block reduction, one partial per CTA, summed in CTA order by k_diag_sums_final
    double r[3] = {s_vv, s_vvp, s_vvw};
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r[q] = dadd(r[q], __shfl_xor_sync(0xffffffffu, r[q], o));
      if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = r[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = s_red[threadIdx.x][0];
      for (int k = 1; k < (int)(blockDim.x >> 5); k++) t = dadd(t, s_red[threadIdx.x][k]);
      a.sum_partial[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
  }
}

// Shared-memory form of the histogram, fused with the output_field sums: ONE pass over x, v, p, w per output step.
// The CTA keeps a private grid of (nv_opd + 1) rows (the spare row takes the iv + 1 contributions of a velocity within
// an ulp of v_max, where (v + v_max) / (2 v_max) rounds to 1 -- the reference indexes out of bounds there) and
// accumulates with 128-bit CAS loops (sm_100a has no native 64-bit shared atomic add; the unit retires ~1 lane-atomic
// per clock per SM, which is what bounds this kernel, not HBM):
//   {g, f} pairs per cell            : 4 CAS.128 per marker (one per bilinear corner)
//   delta f as {left, right} pairs   : 2 CAS.128 per marker (one per row; pertb[r][ix] = left[r][ix] + right[r][ix-1])
// = 6 atomics per marker instead of the 8 of the round-1 kernel.  Both divisions use the exact-reciprocal form of the
// push kernels (div_const: bit-identical to IEEE division, IEEE fallback inside).  The grid is flushed to one of the
// ncopies L2-resident grids with RED.ADD.F64.
// CTA size of the fused output kernel.  Measured on B200 at 1e8 markers (profiles/r02_ab_experiments.md): 1024 threads
// with six serial CAS loops 2.83 ms; 512 threads 2.96 ms; issuing the six CAS of a marker as one overlapped batch is
// slower (3.97 ms at 512 threads, 5.2 ms at 1024: the batch needs ~80 registers).
#ifndef PIC1DP_DIAG_THREADS
#define PIC1DP_DIAG_THREADS 1024
#endif
template <bool SUMS>
__global__ void __launch_bounds__(PIC1DP_DIAG_THREADS, 1) k_diag_fused(const DiagArgs a) {
  extern __shared__ __align__(16) double sh[];
  __shared__ double s_red[3][32];
  const int nxo = a.nx_opd, ncell = nxo * a.nv_opd, ncell1 = nxo * (a.nv_opd + 1);
  double *s_mt = sh;                          // {markr, total} pairs, ncell1 slots
  double *s_pp = sh + 2 * (size_t)ncell1;     // pertb {left, right} pairs, ncell1 slots
  for (int j = threadIdx.x; j < 4 * ncell1; j += blockDim.x) sh[j] = 0.0;
  __syncthreads();
  const double rnx = (double)nxo, rnv = (double)(a.nv_opd - 1), two_vmax = dmul(a.v_max, 2.0);
  const double rlx = 1.0 / a.lx, r2v = 1.0 / two_vmax;
  double s_vv = 0.0, s_vvp = 0.0, s_vvw = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __ldcs(a.v + i), p = __ldcs(a.p + i);
    const double w = a.deltaf ? __ldcs(a.w + i) : 0.0;
    if (SUMS) {
      const double vv = dmul(v, v);  // VecPointwiseMult(tmp1, v, v)  :128
      s_vv = dadd(s_vv, vv);
      s_vvp = dadd(s_vvp, dmul(vv, p));  // :138
      if (a.deltaf) s_vvw = dadd(s_vvw, dmul(vv, w));  // :147
    }
    if (a.nx_opd == 0 || fabs(v) >= a.v_max) continue;  // :241 (nx_opd == 0: sums only)
    const double x = __ldcs(a.x + i);
    double sx = dmul(div_const(x, a.lx, rlx), rnx);  // :243
    int ix = __double2int_rd(sx);
    sx = dsub(1.0, dsub(sx, (double)ix));  // :245
    double sv = dmul(div_const(dadd(v, a.v_max), two_vmax, r2v), rnv);  // :247-248
    const int iv = __double2int_rd(sv);
    sv = dsub(1.0, dsub(sv, (double)iv));  // :250
    if ((unsigned)ix >= (unsigned)nxo) {  // x == lx exactly (the reference would index out of bounds)
      ix = 0;
      sx = 1.0;
    }
    int ix2 = ix + 1;
    if (ix2 > nxo - 1) ix2 = 0;  // :272
    const double sx2 = dsub(1.0, sx), sv2 = dsub(1.0, sv);  // :273
    const int r0 = iv * nxo, r1 = r0 + nxo;
    const double w00 = dmul(sx, sv), w10 = dmul(sx, sv2), w01 = dmul(sx2, sv), w11 = dmul(sx2, sv2);
    const int slot[6] = {r0 + ix, r1 + ix, r0 + ix2, r1 + ix2, r0 + ix, r1 + ix};
    const double ax[6] = {w00, w10, w01, w11, dmul(w00, w), dmul(w10, w)};
    const double ay[6] = {dmul(w00, p), dmul(w10, p), dmul(w01, p), dmul(w11, p), dmul(w01, w), dmul(w11, w)};
    Depositor<DEP_SMEM_ATOMIC> gf, pp;
    gf.g = s_mt;
    pp.g = s_pp;
#pragma unroll
    for (int k = 0; k < 4; k++) gf.add(slot[k], 0, ax[k], ay[k], true);
    if (a.deltaf) {
      pp.add(slot[4], 0, ax[4], ay[4], true);
      pp.add(slot[5], 0, ax[5], ay[5], true);
    }
  }
  __syncthreads();
  if (a.nx_opd > 0) {
    double *hm = a.hist + (size_t)(blockIdx.x % a.ncopies) * 3 * ncell;
    const double2 *mt2 = reinterpret_cast<const double2 *>(s_mt), *pp2 = reinterpret_cast<const double2 *>(s_pp);
    for (int j = threadIdx.x; j < ncell; j += blockDim.x) {   // the spare row (j >= ncell) is dropped
      const double2 mt = mt2[j];
      if (mt.x != 0.0) atomicAdd(hm + j, mt.x);
      if (mt.y != 0.0) atomicAdd(hm + ncell + j, mt.y);
      const int ixj = j % nxo, jl = (ixj == 0) ? j + nxo - 1 : j - 1;  // periodic left neighbour in the same row
      const double pv = dadd(pp2[j].x, pp2[jl].y);
      if (pv != 0.0) atomicAdd(hm + 2 * ncell + j, pv);
    }
  }
  if (SUMS) {  // fixed-shape block reduction, one partial per CTA, summed in CTA order by k_diag_sums_final
    double r[3] = {s_vv, s_vvp, s_vvw};
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r[q] = dadd(r[q], __shfl_xor_sync(0xffffffffu, r[q], o));
      if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = r[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double t = s_red[threadIdx.x][0];
      for (int k = 1; k < (int)(blockDim.x >> 5); k++) t = dadd(t, s_red[threadIdx.x][k]);
      a.sum_partial[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
  }
}

// out[3] = sum over CTAs in CTA order
__global__ void k_diag_sums_final(const double *partial, int nparts, double *out) {
  if (threadIdx.x < 3) {
    double t = partial[threadIdx.x];
    for (int k = 1; k < nparts; k++) t = dadd(t, partial[(size_t)k * 3 + threadIdx.x]);
    out[threadIdx.x] = t;
  }
}

// out[3][ncell] = sum over the ncopies private histograms (copy order), and clears them for the next use
__global__ void k_diag_hist_final(double *hist, int ncopies, int ncell3, double *out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncell3) return;
  double t = hist[j];
  hist[j] = 0.0;
  for (int k = 1; k < ncopies; k++) {
    t = dadd(t, hist[(size_t)k * ncell3 + j]);
    hist[(size_t)k * ncell3 + j] = 0.0;
  }
  out[j] = t;
}

}  // namespace pic1dp
