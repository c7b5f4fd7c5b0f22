// field_kernels.cuh -- grid-side kernels of the PIC1D hot path for sm_100a.
//
//  k_reduce_charge : fixed-order sum of the per-CTA private grids + species charge
//                    (/root/reference/src/pic1dp_interaction.F90:81, :126-127; matrix path :52-59)
//  k_finalize_rho  : scale to charge density + full-f offset after the all-reduce (:140-148; matrix path :64-78)
//  k_field_solve   : mode-filtered partial DFT, 1/k, inverse (/root/reference/src/pic1dp_field.F90:231-256).
//                    There is no KSP / tridiagonal solve in the reference: only the kept modes survive.
//  k_field_energy  : |E|_2^2 * lx / nx (/root/reference/src/pic1dp_output.F90:120-123)
//
// The grid is tiny (nx <= a few thousand): these are single- or few-CTA kernels, latency- not bandwidth-bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "particle_kernels.cuh"

namespace pic1dp {

struct GridArgs {
  int nx, nmode, nspecies, ngrids;  // ngrids = CTAs of the particle kernel (private grids per species)
  int deltaf, matrix_path, zero_partials;
  double lx, rnx;
  double Z[4], n[4];
  double *partial;        // [nspecies][ngrids][nx]
  double *red;            // [nred][nx], nred = matrix_path ? nspecies : 1
  double *rho, *E, *mode_re, *mode_im;
  const double *F_re, *F_im, *ginv;  // [nx*nmode] row-major, [nmode]
  double a_im, a_re;      // -1.0/nx, 1.0/nx  (src/pic1dp_field.F90:234, :239)
  double nx_over_lx;      // input_nx / input_lx (src/pic1dp_interaction.F90:77)
  double *energy;
  // peer-memory all-reduce of `red` fused into the reduce kernel (scatter) and the finalize / solve kernels (gather);
  // p2p_nranks == 0: off (single rank, or ncclAllReduce between the kernels)
  int p2p_nranks, p2p_rank, p2p_parity;
  unsigned long long p2p_epoch;
  unsigned long long *p2p_peer[8];  // exchange buffer of every rank: flags[2][nranks] (padded to 256 B) | data[2][nranks][count]
  unsigned int *p2p_counter;        // local: CTAs of the reduce kernel that have stored their part
  unsigned long long *p2p_timeouts; // local error counter
};

__device__ __forceinline__ double *p2p_data(unsigned long long *base, int nranks, int count, int parity, int r) {
  const size_t flag_bytes = ((size_t)2 * nranks * 8 + 255) & ~(size_t)255;
  return reinterpret_cast<double *>(reinterpret_cast<char *>(base) + flag_bytes) + ((size_t)parity * nranks + r) * count;
}

// scatter side: store one reduced value into slot [parity][my rank] of every rank's exchange buffer (NVLink stores)
__device__ __forceinline__ void p2p_store(const GridArgs &g, int idx, double v) {
  const int count = (g.matrix_path ? g.nspecies : 1) * g.nx;
  for (int r = 0; r < g.p2p_nranks; r++) p2p_data(g.p2p_peer[r], g.p2p_nranks, count, g.p2p_parity, g.p2p_rank)[idx] = v;
}

// scatter side, end of the kernel: the last CTA publishes this rank's epoch flag in every rank's buffer
__device__ __forceinline__ void p2p_publish(const GridArgs &g) {
  __threadfence_system();  // my stores are visible system-wide before the counter / flag
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(g.p2p_counter, 1u);
    if (prev == gridDim.x - 1) {
      *g.p2p_counter = 0;
      __threadfence_system();
      for (int r = 0; r < g.p2p_nranks; r++) {
        unsigned long long *flag = g.p2p_peer[r] + (size_t)g.p2p_parity * g.p2p_nranks + g.p2p_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(g.p2p_epoch) : "memory");
      }
    }
  }
}

// gather side, start of the kernel: wait (bounded) until every rank's flag shows this epoch; returns false on timeout
__device__ __forceinline__ bool p2p_wait(const GridArgs &g) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < g.p2p_nranks) {
    const unsigned long long *flag = g.p2p_peer[g.p2p_rank] + (size_t)g.p2p_parity * g.p2p_nranks + threadIdx.x;
    unsigned long long seen = 0;
    long long spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
      if (seen >= g.p2p_epoch) break;
      if (++spins > 200000000LL) {  // seconds: a peer died; report instead of hanging the GPU
        s_ok = 0;
        if (blockIdx.x == 0) atomicAdd(g.p2p_timeouts, 1ULL);
        break;
      }
      __nanosleep(20);
    }
  }
  __syncthreads();
  return s_ok != 0;
}

// gather side: the all-reduced value = sum over ranks in rank order (bitwise identical on every rank)
__device__ __forceinline__ double p2p_sum(const GridArgs &g, int idx) {
  const int count = (g.matrix_path ? g.nspecies : 1) * g.nx;
  double t = p2p_data(g.p2p_peer[g.p2p_rank], g.p2p_nranks, count, g.p2p_parity, 0)[idx];
  for (int r = 1; r < g.p2p_nranks; r++)
    t = dadd(t, p2p_data(g.p2p_peer[g.p2p_rank], g.p2p_nranks, count, g.p2p_parity, r)[idx]);
  return t;
}

// CTA = 32 cells x 8 warps.  Warp q sums the private grids q, q+8, q+16, ... of its 32 cells in that order (4 loads
// in flight), then warp 0 adds the 8 partial sums in warp order: a fixed summation tree, hence deterministic.
__global__ void __launch_bounds__(256) k_reduce_charge(const GridArgs g) {
  __shared__ double s_part[8][33];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  double c2 = 0.0;
  for (int s = 0; s < g.nspecies; s++) {
    double c1 = 0.0;
    if (j < g.nx) {
      double *ps = g.partial + (size_t)s * g.ngrids * g.nx + j;
      int k = wq;
      for (; k + 24 < g.ngrids; k += 32) {
        const double t0 = ps[(size_t)k * g.nx], t1 = ps[(size_t)(k + 8) * g.nx];
        const double t2 = ps[(size_t)(k + 16) * g.nx], t3 = ps[(size_t)(k + 24) * g.nx];
        c1 = dadd(dadd(dadd(dadd(c1, t0), t1), t2), t3);
      }
      for (; k < g.ngrids; k += 8) c1 = dadd(c1, ps[(size_t)k * g.nx]);
      if (g.zero_partials)
        for (k = wq; k < g.ngrids; k += 8) ps[(size_t)k * g.nx] = 0.0;
    }
    s_part[wq][lane] = c1;
    __syncthreads();
    if (wq == 0 && j < g.nx) {
      double t = s_part[0][lane];
#pragma unroll
      for (int q = 1; q < 8; q++) t = dadd(t, s_part[q][lane]);
      if (g.matrix_path) {  // field_tmp = S^T w per species (:52-59)
        if (g.p2p_nranks) p2p_store(g, s * g.nx + j, t);
        else g.red[(size_t)s * g.nx + j] = t;
      } else {
        c2 = dadd(c2, dmul(t, g.Z[s]));   // charge2 += charge1 * Z (:126-127)
      }
    }
    __syncthreads();
  }
  if (wq == 0 && j < g.nx && !g.matrix_path) {
    if (g.p2p_nranks) p2p_store(g, j, c2);
    else g.red[j] = c2;
  }
  if (g.p2p_nranks) p2p_publish(g);  // MPI_Allreduce (:132-133), scatter half
}

// rho from the (all-)reduced grid: charge1 * nx / lx and the full-f offset (:140-148); matrix path :64-78
__device__ __forceinline__ double red_value(const GridArgs &g, int idx) {
  if (!g.p2p_nranks) return g.red[idx];
  const double t = p2p_sum(g, idx);  // MPI_Allreduce (:132-133), gather half
  g.red[idx] = t;
  return t;
}

__device__ __forceinline__ double finalize_rho(const GridArgs &g, int j) {
  double rho;
  if (!g.matrix_path) {
    rho = ddiv(dmul(red_value(g, j), g.rnx), g.lx);  // :140-141
    if (!g.deltaf)
      for (int s = 0; s < g.nspecies; s++) rho = dsub(rho, dmul(g.Z[s], g.n[s]));  // :142-148
  } else {
    rho = 0.0;  // :47
    for (int s = 0; s < g.nspecies; s++) {
      double t = red_value(g, s * g.nx + j);
      if (!g.deltaf) t = dsub(t, ddiv(dmul(g.n[s], g.lx), g.rnx));  // :67
      rho = dadd(rho, dmul(g.Z[s], t));                              // VecAXPY :71
    }
    rho = dmul(rho, g.nx_over_lx);  // VecScale :77
  }
  return rho;
}

__global__ void __launch_bounds__(128) k_finalize_rho(const GridArgs g) {
  const bool ok = g.p2p_nranks ? p2p_wait(g) : true;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < g.nx) g.rho[j] = ok ? finalize_rho(g, j) : __longlong_as_double(0x7ff8000000000000LL);
}

// Single CTA, 1024 threads.  smem: rho[nx] + 2*nmode mode values.
// SEQ=true : thread m walks j = 0..nx-1 in order (bit-identical to sequential-AIJ MatMultTranspose).
// SEQ=false: one or more warps per mode, lanes stride j, fixed shuffle tree + warp-ordered sum (deterministic).
// FINALIZE: also does k_finalize_rho's work first (one launch less per substep inside step()).
template <bool SEQ, bool FINALIZE>
__global__ void __launch_bounds__(1024) k_field_solve(const GridArgs g) {
  extern __shared__ __align__(16) double smem[];
  double *s_rho = smem;
  double *s_re = smem + g.nx;
  double *s_im = s_re + g.nmode;
  const int M = g.nmode, nx = g.nx;
  const bool p2p_ok = (FINALIZE && g.p2p_nranks) ? p2p_wait(g) : true;
  for (int j = threadIdx.x; j < nx; j += blockDim.x) {
    double r;
    if (FINALIZE) {
      r = p2p_ok ? finalize_rho(g, j) : __longlong_as_double(0x7ff8000000000000LL);
      g.rho[j] = r;
    } else {
      r = g.rho[j];
    }
    s_rho[j] = r;
  }
  __syncthreads();
  if (SEQ) {
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      double sim = 0.0, sre = 0.0;
      for (int j = 0; j < nx; j++) {
        sim = dadd(sim, dmul(g.F_re[(size_t)j * M + m], s_rho[j]));  // MatMultTranspose(F_re, rho) :231
        sre = dadd(sre, dmul(g.F_im[(size_t)j * M + m], s_rho[j]));  // MatMultTranspose(F_im, rho) :236
      }
      s_im[m] = dmul(dmul(sim, g.a_im), g.ginv[m]);  // :234, :246
      s_re[m] = dmul(dmul(sre, g.a_re), g.ginv[m]);  // :239, :243
    }
  } else {
    // G warps share one mode (G = largest power of two <= warps / modes, 1 when there are more modes than warps): with
    // the reference's single kept mode the whole CTA projects it instead of one warp walking all nx cells.  Lanes
    // stride j, fixed shuffle tree per warp, then the G warp sums are added in warp order: a fixed summation tree.
    __shared__ double s_part[32][2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int G = 1;
    while (2 * G * M <= nw) G *= 2;
    const int per_pass = nw / G, grp = wid / G, sub = wid % G;
    for (int mb = 0; mb < M; mb += per_pass) {
      const int m = mb + grp;
      double sim = 0.0, sre = 0.0;
      if (m < M && grp < per_pass) {
        for (int j = sub * 32 + lane; j < nx; j += 32 * G) {
          sim = dadd(sim, dmul(g.F_re[(size_t)j * M + m], s_rho[j]));
          sre = dadd(sre, dmul(g.F_im[(size_t)j * M + m], s_rho[j]));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sim = dadd(sim, __shfl_xor_sync(0xffffffffu, sim, o));
        sre = dadd(sre, __shfl_xor_sync(0xffffffffu, sre, o));
      }
      if (lane == 0) {
        s_part[wid][0] = sim;
        s_part[wid][1] = sre;
      }
      __syncthreads();
      if (lane == 0 && sub == 0 && m < M && grp < per_pass) {
        for (int q = 1; q < G; q++) {
          sim = dadd(sim, s_part[wid + q][0]);
          sre = dadd(sre, s_part[wid + q][1]);
        }
        s_im[m] = dmul(dmul(sim, g.a_im), g.ginv[m]);
        s_re[m] = dmul(dmul(sre, g.a_re), g.ginv[m]);
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    g.mode_re[m] = s_re[m];
    g.mode_im[m] = s_im[m];
  }
  // E = 2 * (F_re . mode_re + F_im . mode_im): MatMult then MatMultAdd, row sums left to right (:251-256)
  for (int j = threadIdx.x; j < nx; j += blockDim.x) {
    double sum = 0.0;
    for (int m = 0; m < M; m++) sum = dadd(sum, dmul(g.F_re[(size_t)j * M + m], s_re[m]));
    for (int m = 0; m < M; m++) sum = dadd(sum, dmul(g.F_im[(size_t)j * M + m], s_im[m]));
    g.E[j] = dmul(sum, 2.0);
  }
}

// single CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_field_energy(const GridArgs g) {
  __shared__ double s_part[32];
  double s = 0.0;
  for (int j = threadIdx.x; j < g.nx; j += blockDim.x) s = dadd(s, dmul(g.E[j], g.E[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = dadd(s, __shfl_xor_sync(0xffffffffu, s, o));
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? s_part[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = dadd(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (threadIdx.x == 0) {
      const double nrm = sqrt(s);  // VecNorm(NORM_2)
      *g.energy = ddiv(dmul(dmul(nrm, nrm), g.lx), g.rnx);
    }
  }
}

}  // namespace pic1dp
