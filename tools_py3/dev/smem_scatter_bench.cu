// smem_scatter_bench.cu -- what a B200 SM can do for a scattered accumulation into a 96 KB shared-memory table
// (the x-v histograms of output_ptcldist: 3 x 64 x 64 doubles, 12 accumulations per marker at pseudo-random cells).
// Measures lane-operations per clock per SM for the candidate primitives, so that DESIGN.md can state the bound of
// k_diag_fused from numbers instead of guesses:
//   cas128   : LDS.128 + ATOMS.CAS.128 loop on {a, b} pairs (what k_diag_fused does, 6 per marker)
//   cas64    : LDS.64 + ATOMS.CAS.64 loop on one double
//   add32    : native ATOMS.ADD (32-bit integer, result unused) -- the only native shared-memory add
//   rmw64    : LDS.64 + DADD + STS.64 without atomicity (racy: what an ownership scheme would pay per accumulation
//              if routing were free and conflicts impossible)
//   rmw128   : LDS.128 + 2 DADD + STS.128, same
//   add32ret : ATOMS.ADD with the old value used (carry detection of a two-word counter)
//   add32cf  : ATOMS.ADD with conflict-free addresses (lane l always hits bank l): the unit's peak
//   add32h   : ATOMS.ADD with every other lane predicated off (bank-conflict scaling)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/smem_scatter_bench tools_py3/dev/smem_scatter_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define NCELL 6144   // 128-bit slots (= 12288 doubles = 96 KB)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) {
  s = s * 1664525u + 1013904223u;
  return s >> 8;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_bench(int iters, double *out, unsigned long long *cycles) {
  extern __shared__ __align__(16) double sh[];
  for (int j = threadIdx.x; j < 2 * NCELL; j += blockDim.x) sh[j] = 0.0;
  __syncthreads();
  uint32_t s = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
  const double inc = 1.0 + threadIdx.x * 1e-3;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const uint32_t r = lcg(s);
    if (MODE == 0) {  // cas128
      double2 *slot = reinterpret_cast<double2 *>(sh) + (r % NCELL);
      double2 old = *slot;
      for (;;) {
        const double2 nw = make_double2(old.x + inc, old.y + inc);
        unsigned long long o0 = __double_as_longlong(old.x), o1 = __double_as_longlong(old.y);
        unsigned long long n0 = __double_as_longlong(nw.x), n1 = __double_as_longlong(nw.y), r0, r1;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(slot);
        asm volatile(
            "{ .reg .b128 c, n, d;\n mov.b128 c, {%3, %4};\n mov.b128 n, {%5, %6};\n"
            " atom.shared.cas.b128 d, [%2], c, n;\n mov.b128 {%0, %1}, d; }"
            : "=l"(r0), "=l"(r1)
            : "r"(addr), "l"(o0), "l"(o1), "l"(n0), "l"(n1)
            : "memory");
        if (r0 == o0 && r1 == o1) break;
        old.x = __longlong_as_double(r0);
        old.y = __longlong_as_double(r1);
      }
    } else if (MODE == 1) {  // cas64
      unsigned long long *slot = reinterpret_cast<unsigned long long *>(sh) + (r % (2 * NCELL));
      unsigned long long old = *slot;
      for (;;) {
        const unsigned long long nw = __double_as_longlong(__longlong_as_double(old) + inc);
        const unsigned long long got = atomicCAS(slot, old, nw);
        if (got == old) break;
        old = got;
      }
    } else if (MODE == 2) {  // native 32-bit add, no result
      unsigned *slot = reinterpret_cast<unsigned *>(sh) + (r & 16383u);   // power-of-two range: no modulo sequence
      atomicAdd(slot, r | 1u);
    } else if (MODE == 3) {  // rmw64, racy
      volatile double *slot = sh + (r % (2 * NCELL));
      *slot = *slot + inc;
    } else if (MODE == 5) {  // returning add: carry into a second word when the low word wraps
      unsigned *slot = reinterpret_cast<unsigned *>(sh) + 2 * (r & 8191u);
      const unsigned x = r * 2654435761u;
      const unsigned old = atomicAdd(slot, x);
      atomicAdd(slot + 1, (r & 0xffu) + ((old + x) < x ? 1u : 0u));
      it++;   // two atomics per trip
    } else if (MODE == 6) {  // conflict-free: bank = lane
      unsigned *slot = reinterpret_cast<unsigned *>(sh) + ((r & 511u) * 32 + (threadIdx.x & 31));
      atomicAdd(slot, r | 1u);
    } else if (MODE == 7) {  // half of the lanes
      unsigned *slot = reinterpret_cast<unsigned *>(sh) + (r & 16383u);
      if (threadIdx.x & 1) atomicAdd(slot, r | 1u);
    } else if (MODE == 8) {  // nine adds at consecutive words of one random slot (the limb kernel's pattern)
      unsigned *slot = reinterpret_cast<unsigned *>(sh) + 9 * (r & 2047u);
#pragma unroll
      for (int k = 0; k < 9; k++) atomicAdd(slot + k, (r >> k) | 1u);
      it += 8;
    } else {  // rmw128, racy
      const unsigned addr = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double2 *>(sh) + (r % NCELL));
      double a, b;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
      a += inc;
      b += inc;
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  double acc = 0.0;
  for (int j = threadIdx.x; j < 2 * NCELL; j += blockDim.x) acc += sh[j];
  out[blockIdx.x * 1024 + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
static void run(const char *name, int nsm, int iters) {
  double *out;
  unsigned long long *cyc, h[1024];
  cudaMalloc(&out, (size_t)nsm * 1024 * 8);
  cudaMalloc(&cyc, (size_t)nsm * 8);
  const int smem = 2 * NCELL * 8;
  cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_bench<MODE><<<nsm, 1024, smem>>>(iters / 10, out, cyc);
  cudaEventRecord(e0);
  k_bench<MODE><<<nsm, 1024, smem>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(h, cyc, (size_t)nsm * 8, cudaMemcpyDeviceToHost);
  double mean = 0.0;
  for (int i = 0; i < nsm; i++) mean += (double)h[i];
  mean /= nsm;
  const double ops = 1024.0 * iters;
  printf("%-8s %8.3f ms  %10.0f cycles/CTA  %6.3f lane-ops/clk/SM  (%s)\n", name, ms, mean, ops / mean,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount, iters = 20000;
  printf("%s, %d SMs, 1024 threads per SM, %d scattered operations per thread into a 96 KB table\n", p.name, nsm, iters);
  run<0>("cas128", nsm, iters);
  run<1>("cas64", nsm, iters);
  run<2>("add32", nsm, iters);
  run<3>("rmw64", nsm, iters);
  run<4>("rmw128", nsm, iters);
  run<5>("add32ret", nsm, iters);
  run<6>("add32cf", nsm, iters);
  run<7>("add32h", nsm, iters);   // lane-ops counted for all 32 lanes: halve the figure
  run<8>("add32x9", nsm, iters);
  return 0;
}
