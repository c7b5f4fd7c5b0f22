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
  // limb path (k_diag_limb): exact fixed-point accumulation with native 32-bit shared-memory adds
  __int128 *tab;                  // [gridDim.x][3][ncell] per-CTA integer histograms (L2 resident), zeroed before the launch
  const unsigned *max_p_hi, *max_w_hi;   // high words of (a bound on) max |p| and max |w| over this species
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


// ---- limb path: the same fused pass with the histograms kept as exact fixed-point integers ---------------------------
// What bounds the CAS form above is the shared-memory atomic unit.  Measured on B200 (tools_py3/dev/smem_scatter_bench.cu,
// 1024 threads per SM, a 96 KB table, lane addresses with few bank conflicts): LDS.128 + ATOMS.CAS.128 retires 0.69
// lane-operations per clock per SM, the 64-bit CAS 1.39, the NATIVE 32-bit integer add (ATOMS.ADD) 12-13 when its
// result is unused and 8.4 when it is used (at random banks ncu counts ~3.6 shared-memory wavefronts per ATOMS
// instruction, ~9 lane-operations per clock); sm_100a has no native 64-bit or floating-point shared-memory add.
// So every contribution c (a bilinear weight, or weight * p, or weight * w) is scaled by a power of two to an integer
// I = RN(c * 2^k), |I| <= 2^46, and added to a two-word counter {low 32 bits (wraps), I >> 32 (signed)} with two
// native adds: the first returns the old low word, and the thread whose add wrapped it adds the carry along with its
// high part.  24 native adds per marker replace 6 CAS.128 loops.  Per output step at 1e8 markers (device time of
// pic1dp_gpu_output_all, profiles/r02_ab_experiments.md): CAS.128 kernel 2.72 ms; three 16-bit limbs with 36
// fire-and-forget adds 1.87-1.93 ms; two words with carry 1.52 ms; with the register double buffer of the loads 1.37 ms.
// The high word absorbs 2^16 additions (2^16 * (2^14 + 1) < 2^31), so every PIC1DP_LIMB_FLUSH = 64 tile steps of 1024
// markers the CTA folds its counters into its own 128-bit integer histogram in global memory (L2 resident, no
// atomics: one table per CTA; 128 bits cannot overflow for any marker count), and k_diag_limb_final adds the tables.
// Integer addition is associative: the histograms are bitwise reproducible from run to run and independent of the
// order in which markers, warps and CTAs arrive.  Accuracy: each contribution is rounded once to 2^-47 of the power of
// two above the largest |p| (|w|) of the species (<= 7.2e-15 relative to it); the sums are exact.
// The scales come from device-resident maxima (p: k_absmax_hi once per marker set; w: the running maximum that the
// fixed-point deposit keeps anyway, else k_absmax_hi before the output), so no host synchronisation is needed between
// the passes.
#ifndef PIC1DP_LIMB_FLUSH
#define PIC1DP_LIMB_FLUSH 64
#endif
#define PIC1DP_LIMB_BITS 46

// 2^k such that |x| < 2^(E - 1022) (E = biased exponent of the maximum) maps to |x| 2^k < 2^ibits (ibits = PIC1DP_LIMB_BITS)
__device__ __forceinline__ double limb_scale(unsigned max_hi, int ibits) {
  if (max_hi < 0x00100000u) return 1.0;  // zero or denormal maximum: nothing to resolve
  int k = ibits - ((int)(max_hi >> 20) - 1022);
  k = min(max(k, -1022), 1023);
  return __hiloint2double((k + 1023) << 20, 0);
}

// A counter is two words {low 32 bits, I >> 32}: the add to the low word returns the old value, and the thread whose add
// wrapped it carries one into the high word.  7 words per bin (3 counters + one pad word: 7 is odd, so consecutive
// bins start in different banks and all 32 banks are reachable).
#define PIC1DP_LIMB_W 7
__device__ __forceinline__ void limb_split(double val, double scale, unsigned &lo, unsigned &hi) {
  // 1.5 * 2^52 + I: the low mantissa bits hold I = RN(val * scale) in two's complement (|I| < 2^51)
  const double t = fma(val, scale, 6755399441055744.0);
  lo = (unsigned)__double2loint(t);
  hi = (unsigned)(__double2hiint(t) - 0x43380000);   // floor(I / 2^32)
}

// The (up to) three contributions of one marker to one bin: the returning adds, then the carries that need their results.
// (Issuing the returning adds of the next corner before the carries of this one measured the same: 1.524 vs 1.527 ms
// before the register double buffer of the loads, 1.383 vs 1.375 ms with it.)
struct Limb3 {
  unsigned *c;
  unsigned lo[3], hi[3], old[3];
};
__device__ __forceinline__ void limb_issue(Limb3 &L, unsigned *c, bool third, double v0, double s0, double v1, double s1,
                                           double v2, double s2) {
  L.c = c;
  limb_split(v0, s0, L.lo[0], L.hi[0]);
  limb_split(v1, s1, L.lo[1], L.hi[1]);
  L.lo[2] = L.hi[2] = L.old[2] = 0u;
  if (third) limb_split(v2, s2, L.lo[2], L.hi[2]);
  L.old[0] = atomicAdd(c, L.lo[0]);
  L.old[1] = atomicAdd(c + 2, L.lo[1]);
  if (third) L.old[2] = atomicAdd(c + 4, L.lo[2]);
}
__device__ __forceinline__ void limb_carry(const Limb3 &L, bool third) {
  // results unused: fire-and-forget ATOMS.ADD.  (Skipping the add when high part + carry is zero -- about a quarter of
  // them with Maxwellian weights -- measured slower, 1.64 vs 1.52 ms: the branches cost more than the lanes save.)
  atomicAdd(L.c + 1, L.hi[0] + ((L.old[0] + L.lo[0]) < L.lo[0] ? 1u : 0u));
  atomicAdd(L.c + 3, L.hi[1] + ((L.old[1] + L.lo[1]) < L.lo[1] ? 1u : 0u));
  if (third) atomicAdd(L.c + 5, L.hi[2] + ((L.old[2] + L.lo[2]) < L.lo[2] ? 1u : 0u));
}

// fold the CTA's counters into its 128-bit table and clear them; the spare row (bin >= ncell) is dropped
__device__ __forceinline__ void limb_flush(unsigned *s_l, int ncell, int ncell1, __int128 *tab) {
  for (int j = threadIdx.x; j < 3 * ncell1; j += blockDim.x) {   // j = bin * 3 + quantity
    const int bin = j / 3, q = j - 3 * bin;
    unsigned *c = s_l + PIC1DP_LIMB_W * bin + 2 * q;
    const unsigned a0 = c[0], a1 = c[1];
    if ((a0 | a1) == 0u) continue;
    c[0] = 0u;
    c[1] = 0u;
    if (bin < ncell) tab[(size_t)q * ncell + bin] += (__int128)((long long)a0 + (long long)(int)a1 * 4294967296LL);
  }
}

template <bool SUMS>
__global__ void __launch_bounds__(1024, 1) k_diag_limb(const DiagArgs a) {
  extern __shared__ __align__(16) unsigned s_l[];   // [ncell1][PIC1DP_LIMB_W]: {low, high} words of g, f, delta f + pad
  __shared__ double s_red[3][32];
  const int nxo = a.nx_opd, ncell = nxo * a.nv_opd, ncell1 = nxo * (a.nv_opd + 1);
  for (int j = threadIdx.x; j < PIC1DP_LIMB_W * ncell1; j += blockDim.x) s_l[j] = 0u;
  __syncthreads();
  __int128 *tab = a.tab + (size_t)blockIdx.x * 3 * ncell;
  const double sg = __hiloint2double((PIC1DP_LIMB_BITS + 1023) << 20, 0);   // weights <= 1
  const double sf = limb_scale(*a.max_p_hi, PIC1DP_LIMB_BITS), sw = limb_scale(*a.max_w_hi, PIC1DP_LIMB_BITS);
  const double rnx = (double)nxo, rnv = (double)(a.nv_opd - 1), two_vmax = dmul(a.v_max, 2.0);
  const double rlx = 1.0 / a.lx, r2v = 1.0 / two_vmax;
  double s_vv = 0.0, s_vvp = 0.0, s_vvw = 0.0;
  int since_flush = 0;
  // CTA-uniform trip count: the flush below needs every thread at its barriers.
  // Register double buffer: the four loads of the next tile step are in flight while this one's atomics are issued
  // (ncu without it: 55 % of the stall samples at the first use of the loaded v and x; 1.51 -> 1.37 ms per output step
  // at 1e8 markers; an L2 prefetch instead 1.44 ms, both together 1.44 ms).
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double v_n = 0.0, p_n = 0.0, w_n = 0.0, x_n = 0.0;
  {
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 < a.np) {
      v_n = __ldcs(a.v + i0);
      p_n = __ldcs(a.p + i0);
      if (a.deltaf) w_n = __ldcs(a.w + i0);
      x_n = __ldcs(a.x + i0);   // also for |v| >= v_max markers, which do not need it
    }
  }
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.np; base += stride) {
    const int64_t i = base + threadIdx.x, nxt = i + stride;
    const double v = v_n, p = p_n, w = w_n, x = x_n;
    if (nxt < a.np) {
      v_n = __ldcs(a.v + nxt);
      p_n = __ldcs(a.p + nxt);
      if (a.deltaf) w_n = __ldcs(a.w + nxt);
      x_n = __ldcs(a.x + nxt);
    }
    if (i < a.np) {
      if (SUMS) {
        const double vv = dmul(v, v);  // VecPointwiseMult(tmp1, v, v)  :128
        s_vv = dadd(s_vv, vv);
        s_vvp = dadd(s_vvp, dmul(vv, p));  // :138
        if (a.deltaf) s_vvw = dadd(s_vvw, dmul(vv, w));  // :147
      }
      if (fabs(v) < a.v_max) {  // :241
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
        const int r0 = iv * nxo, r1 = r0 + nxo;   // row iv + 1 == nv_opd is the spare row
        const double wc[4] = {dmul(sx, sv), dmul(sx, sv2), dmul(sx2, sv), dmul(sx2, sv2)};
        const int slot[4] = {r0 + ix, r1 + ix, r0 + ix2, r1 + ix2};
#pragma unroll
        const bool third = a.deltaf != 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          Limb3 L;
          limb_issue(L, s_l + PIC1DP_LIMB_W * slot[k], third, wc[k], sg, dmul(wc[k], p), sf, dmul(wc[k], w), sw);
          limb_carry(L, third);
        }
      }
    }
    if (++since_flush == PIC1DP_LIMB_FLUSH) {
      since_flush = 0;
      __syncthreads();
      limb_flush(s_l, ncell, ncell1, tab);
      __syncthreads();
    }
  }
  __syncthreads();
  limb_flush(s_l, ncell, ncell1, tab);
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

// out[3][ncell] = (sum over the CTA tables, exact in 128-bit integers) / scale.  CTA = 64 entries x 4 groups of tables.
__global__ void __launch_bounds__(256) k_diag_limb_final(const __int128 *tab, int ntab, int ncell, const unsigned *max_p_hi,
                                                         const unsigned *max_w_hi, double *out) {
  __shared__ __int128 s_part[4][64];
  const int jl = threadIdx.x & 63, grp = threadIdx.x >> 6, j = blockIdx.x * 64 + jl;
  __int128 t = 0;
  if (j < 3 * ncell)
    for (int k = grp; k < ntab; k += 4) t += tab[(size_t)k * 3 * ncell + j];
  s_part[grp][jl] = t;
  __syncthreads();
  if (grp != 0 || j >= 3 * ncell) return;
  t = s_part[0][jl] + s_part[1][jl] + s_part[2][jl] + s_part[3][jl];
  const int q = j / ncell;
  const double scale = q == 0 ? __hiloint2double((PIC1DP_LIMB_BITS + 1023) << 20, 0)
                              : limb_scale(q == 1 ? *max_p_hi : *max_w_hi, PIC1DP_LIMB_BITS);
  // two roundings when |t| >= 2^64 (more than 2^18 markers of the largest magnitude in one cell), one otherwise
  const bool neg = t < 0;
  const unsigned __int128 u = neg ? (unsigned __int128)(-t) : (unsigned __int128)t;
  double val = fma((double)(unsigned long long)(u >> 64), 18446744073709551616.0, (double)(unsigned long long)u);
  if (neg) val = -val;
  out[j] = ddiv(val, scale);   // power of two: exact
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
