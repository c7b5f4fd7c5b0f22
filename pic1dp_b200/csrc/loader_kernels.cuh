// loader_kernels.cuh -- non-template marker kernels used only by the C ABI unit: the device side of particle_load
// (/root/reference/src/pic1dp_particle.F90:172-264) and the weight export of get_shape_x.
#pragma once
#include "particle_kernels.cuh"

namespace pic1dp {

// ---- device-side particle_load for uniform-v markers (src/pic1dp_particle.F90:179-264) ----
struct LoadArgs {
  double *x, *v, *p, *w;  // in: x holds rand_x, v holds rand_v (uploaded in place); out: the loaded markers
  int64_t np;
  double lx, v_max, ninit;
  SpeciesConst c;
  double T2;              // temperature2 (SpeciesConst keeps only T2/m)
  int dist, linear, init_nmode;
  int imarker;            // 2: uniform v in [-v_max, v_max] (v holds uniforms); 1: physical Maxwellian (v holds Gaussians)
  int init_mode[8];
  double init_cos[8], init_sin[8];
};

__global__ void __launch_bounds__(256) k_load_markers(const LoadArgs a) {
  const double PI = 3.14159265358979323846264338327950288419716939937510582;  // PETSC_PI
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += (int64_t)gridDim.x * blockDim.x) {
    const double n = a.c.n, v0 = a.c.v0, T = a.c.T, m = a.c.m, T2 = a.T2;
    // input_imarker == 1 (:172-178): markers loaded like the physical (shifted) Maxwellian from the Gaussian stream,
    // all with the same p; input_imarker == 2 (:179-219): uniform in velocity space
    const double pv = (a.imarker == 1) ? dadd(dmul(a.v[i], __dsqrt_rn(ddiv(T, m))), v0)
                                       : dmul(dmul(dsub(a.v[i], 0.5), 2.0), a.v_max);  // :181
    // lx * 2 * v_max / nparticle_init, left to right as written (:183-185, :198-200)
    double pp;
    if (a.imarker == 1) {  // :176-177
      pp = ddiv(dmul(n, a.lx), a.ninit);
    } else if (a.dist == 1) {  // two-stream1 :183-186
      const double pre = ddiv(dmul(dmul(dmul(n, a.lx), 2.0), a.v_max), a.ninit);
      pp = ddiv(dmul(dmul(pre, dmul(pv, pv)), exp(ddiv(-dmul(pv, pv), 2.0))), __dsqrt_rn(dmul(2.0, PI)));
    } else if (a.dist == 2) {  // two-stream2 :188-196
      const double pre = ddiv(dmul(dmul(dmul(n, a.lx), 2.0), a.v_max), a.ninit);
      const double vp = dadd(pv, v0), vm = dsub(pv, v0);
      const double twoTm = ddiv(dmul(2.0, T), m);
      const double e = dadd(exp(ddiv(-dmul(vp, vp), twoTm)), exp(ddiv(-dmul(vm, vm), twoTm)));
      pp = ddiv(dmul(pre, e), __dsqrt_rn(ddiv(dmul(dmul(8.0, PI), T), m)));
    } else if (a.dist == 3) {  // bump-on-tail :198-209
      const double pre = ddiv(dmul(dmul(dmul(1.0, a.lx), 2.0), a.v_max), a.ninit);
      const double vm = dsub(pv, v0);
      const double t1 = ddiv(dmul(n, exp(ddiv(-dmul(pv, pv), ddiv(dmul(2.0, T), m)))),
                             __dsqrt_rn(ddiv(dmul(dmul(2.0, PI), T), m)));
      const double t2 = ddiv(dmul(dsub(1.0, n), exp(ddiv(-dmul(vm, vm), ddiv(dmul(2.0, T2), m)))),
                             __dsqrt_rn(ddiv(dmul(dmul(2.0, PI), T2), m)));
      pp = dmul(pre, dadd(t1, t2));
    } else {  // (shifted) Maxwellian :211-217
      const double pre = ddiv(dmul(dmul(dmul(n, a.lx), 2.0), a.v_max), a.ninit);
      const double vm = dsub(pv, v0);
      pp = ddiv(dmul(pre, exp(ddiv(-dmul(vm, vm), ddiv(dmul(2.0, T), m)))), __dsqrt_rn(ddiv(dmul(dmul(2.0, PI), T), m)));
    }
    const double px = dmul(a.x[i], a.lx);  // :223
    double pw = 0.0;                       // :225
    for (int im = 0; im < a.init_nmode; im++) {  // :226-232
      const double k = dmul(ddiv(dmul(2.0, PI), a.lx), (double)a.init_mode[im]);
      const double arg = dmul(k, px);
      double sn, cs;
      sincos(arg, &sn, &cs);   // one argument reduction for both (same results as sin() and cos())
      pw = dadd(dadd(pw, dmul(a.init_cos[im], cs)), dmul(a.init_sin[im], sn));
    }
    pw = dmul(dmul(pw, pp), 1.0);  // * p * input_pertb_shape (= 1.0), :235-236
    if (!a.linear) pp = dadd(pp, dmul(1.0, pw));  // VecAXPY(p, 1.0, w), :260-263
    a.x[i] = px;
    a.v[i] = pv;
    a.p[i] = pp;
    a.w[i] = pw;
  }
}

// high word of max |src[i]| (the scale source of the fixed-point deposit), raised into *out with atomicMax
__global__ void __launch_bounds__(256) k_absmax_hi(const double *__restrict__ src, const int64_t np, unsigned *out) {
  unsigned m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (unsigned)__double2hiint(src[i]) & 0x7fffffffu);
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// weights of the current x, exported for parity checks (particle_shape_x_indexes / _values of iptclshape 3,
// src/pic1dp_particle.F90:331-332)
__global__ void __launch_bounds__(256) k_shape_x(const ParticleArgs a, int *ix, double *sl, double *sr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.np; i += (int64_t)gridDim.x * blockDim.x) {
    bool oob = false;
    const Shape s = shape_of(a.x_cur[i], a.lx, a.rlx, a.rnx, a.nx, a.right_frac, oob);
    ix[i] = s.ix;
    sl[i] = s.sl;
    sr[i] = s.sr;
  }
}

}  // namespace pic1dp
