// particle_kernels.cuh -- marker-side kernels of the PIC1D hot path for sm_100a.
//
// One fused kernel per RK substep does, per marker: weights (index + linear weights, never stored), field
// gather S.E from a shared-memory copy of E, RK2 push of x, w, v, periodic wrap, and charge deposition S^T w
// into an on-chip grid.  Replaces interaction_push_particle (/root/reference/src/pic1dp_interaction.F90:161-370)
// followed by the marker loop of interaction_collect_charge (:96-114).  Unfused variants (push only, deposit
// only) keep the reference's call-by-call side effects.
//
// Arithmetic is "strict": every fp64 operation is an explicit round-to-nearest intrinsic in the reference's
// left-to-right order, so nvcc cannot contract to FMA.  Given identical inputs the cell index, weights, x and v
// equal the x86-64 reference build's bit for bit -- proven for every division the fast path does not hand to the IEEE
// routine except one corner (quotient an exact power of two AND within 2^-104 of the midpoint below it: probability
// below 2^-150 per division, see div_suspect); w differs only through exp() (<= 1 ulp vs glibc).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fp_strict.cuh"

// Resident threads per SM the particle kernels are compiled for: 1024 -> 64 registers/thread, 768 -> 85, 512 -> 128.
#ifndef PIC1DP_MAXTHREADS
#define PIC1DP_MAXTHREADS 1024
#endif
// which fused-kernel variants issue L2 prefetches: bit 0/1 = atomic deposits irk 1/2, bit 2/3 = warp-private irk 1/2
#ifndef PIC1DP_PF_MASK
#define PIC1DP_PF_MASK 13
#endif
// which fused-kernel variants stage the next tile step's v in registers (same bit layout as PIC1DP_PF_MASK).
// Measured on B200 at 1e8 markers, nx = 1024: warp-private irk=1 1.264 -> 1.158 ms; atomic deposits +0.5 % slower.
// how many tile steps ahead the L2 prefetch runs (2 measured 6 % slower than 1 in both arithmetic modes, call r02n)
#ifndef PIC1DP_PF_DIST
#define PIC1DP_PF_DIST 1
#endif
#ifndef PIC1DP_PV_MASK
#define PIC1DP_PV_MASK 4
#endif
// bit 0: the fp64 pair-grid depositor deposits the two markers of a thread with overlapped round trips (Depositor::add2).
// Measured on B200 (profiles/r02_ab_experiments.md): the overlap hides the CAS latency but the extra live registers
// spill at 64 registers per thread and the kernels get 2-4 % slower.
// bit 1: the same for the fixed-point depositor with native adds (four returning adds, then the four carries): no spill,
// 2.161-2.185 -> 2.134-2.136 ms per step at 1e8 markers (strict arithmetic).
#ifndef PIC1DP_ADD2
#define PIC1DP_ADD2 2
#endif
// fixed-point conversion: 1 = F2I.S64.F64, 0 = magic-number add + integer subtract
#ifndef PIC1DP_FIXED_F2I
#define PIC1DP_FIXED_F2I 1
#endif
// experiment switches (profiles/r02_ab_experiments.md): CAS loops that re-read the slot on retry; second Newton step of
// the reciprocal in the fast divisions (not needed for correctness: the residual test certifies the quotient)
#ifndef PIC1DP_CAS_RELOAD
#define PIC1DP_CAS_RELOAD 0
#endif
// fixed-point deposit: 1 = 64-bit integer adds as two native 32-bit ATOMS.ADD (low word returning, high word + carry),
// 0 = the round-2 ATOMS.CAS.128 loop on {left, right} int64 pairs
#ifndef PIC1DP_FIXED_NATIVE
#define PIC1DP_FIXED_NATIVE 1
#endif
#ifndef PIC1DP_NEWTON2
#define PIC1DP_NEWTON2 0
#endif
// CAS retries written as straight-line code before the spin loop (whose body carries ptxas's YIELD): with ~2 retrying
// warp-instructions per tile step at nx = 1024 the loop's YIELD is still executed about as often as a deposit
#ifndef PIC1DP_CAS_UNROLL
#define PIC1DP_CAS_UNROLL 4
#endif

namespace pic1dp {


// ---- exp() with its polynomial in the constant bank ------------------------------------------------------
// exp(a) = 2^n * (1 + r + r^2 q(r)), n = rint(a log2 e), r = a - n ln2 (two-term Cody-Waite with FMA),
// q = degree-9 near-minimax polynomial (tools_py3/gen_exp_coeffs.py, approximation error 2^-55.8).  Total error
// <= 1.05 ulp against 80-bit expl() over [-700, 0] (host model tests/csrc/exp_model.c); the coefficients are read as constant-bank operands of DFMA instead of being
// rebuilt in registers with 2 moves each, which is what makes libdevice's exp cost ~60 issue slots here.
// |a| > 700 (never reached by physical velocities) falls back to libdevice.
static __constant__ double c_exp_poly[10] = {
    0x1.af38a9b0ec855p-26, 0x1.289185613a3d6p-22, 0x1.71de0dae63bb3p-19, 0x1.a019b90d2ae7ap-16,
    0x1.a01a01a7c41d5p-13, 0x1.6c16c1788bd90p-10, 0x1.11111111109b3p-7,  0x1.5555555553d63p-5,
    0x1.5555555555556p-3,  0x1.0000000000001p-1};
static __constant__ double c_exp_red[3] = {0x1.71547652b82fep+0 /* log2 e */, 0x1.62e42fefa39efp-1 /* ln2 hi */,
                                    0x1.abc9e3b39803fp-56 /* ln2 lo */};

__device__ double exp_slow(double a);

// ---- correctly rounded division with a straight-line fast path ---------------------------------------------
// q0 = a*y, r = a - q0*b (exact, FMA), q1 = q0 + r*y is RN(a/b + d) with |d| <= 2^-104 |a/b| when y is 1/b to
// ~1 ulp, so q1 can only differ from RN(a/b) when a/b lies within 2^-104 of a rounding midpoint.  The exact
// residual r1 = a - q1*b detects that: RN(a/b) != q1 implies |r1| > b * ulp(q1) / 2; such operands (and anything
// outside 2^+-500) are re-divided with the IEEE routine on a rare noinline path.  (Not covered: q1 an exact power
// of two AND a/b within 2^-104 of the midpoint just below it, where the lower half-ulp is smaller.)
// rare paths: IEEE division, libdevice exp, fmod.  They are reached only from the scalar functions (push_one, wrap_x,
// shape_of), i.e. from push_pair_slow; the 2-wide hot loop only raises a flag.
#define PIC1DP_RARE __forceinline__
__device__ PIC1DP_RARE double div_slow(double a, double b) { return __ddiv_rn(a, b); }
__device__ PIC1DP_RARE double exp_slow(double a) { return exp(a); }
__device__ PIC1DP_RARE double wrap_slow(double x, double lx) {
  double r = fmod(x, lx);
  if (r < 0.0) r = __dadd_rn(r, lx);
  return r;
}

// true when q1 (within an ulp of a/b) is NOT provably RN(a/b): |a - q1*b| > b*ulp(q1)/2
__device__ __forceinline__ bool div_suspect(double a, double b, double q1) {
  const double r1 = fma(-q1, b, a);
  const int e = __double2hiint(q1) & 0x7ff00000;
  const double h = __hiloint2double(__double2hiint(b) + e - (1076 << 20), __double2loint(b));  // b * ulp(q1) / 2
  return fabs(r1) > h;
}

// a / b for a constant divisor b > 0 with y = RN(1/b) precomputed on the host; a >= 0 (a marker coordinate).
// a == 0, a < 2^-800 (residuals would underflow), a < 0, NaN and the near-midpoint case take the IEEE division.
__device__ __forceinline__ double div_const(double a, double b, double y) {
  const double q0 = a * y;
  const double r = fma(-q0, b, a);
  double q1 = fma(r, y, q0);
  if (__builtin_expect(!(a >= 0x1p-800) || div_suspect(a, b, q1), 0)) q1 = div_slow(a, b);
  return q1;
}

// general a / b, b > 0
__device__ __forceinline__ double div_pos(double a, double b) {
  const unsigned eb = (unsigned)(__double2hiint(b) & 0x7ff00000) - (523u << 20);  // exponent of b in [-500, 500)
  const unsigned ea = (unsigned)(__double2hiint(a) & 0x7ff00000) - (523u << 20);
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));  // MUFU.RCP64H: ~20 good bits
  double e = fma(-b, y, 1.0);   // 2^-20; y (1 + e + e^2) is 1/b to ~2^-60: one cubic step reaches working precision,
  e = fma(e, e, e);             // and q1 below only needs y to an ulp (error of q1 ~ eps(q0) * eps(y) = 2^-104)
  y = fma(y, e, y);
  const double q0 = a * y;
  const double r = fma(-q0, b, a);
  double q1 = fma(r, y, q0);
  if (__builtin_expect((eb | ea) >= (1000u << 20) || !(b > 0.0) || div_suspect(a, b, q1), 0)) q1 = div_slow(a, b);
  return q1;
}

// two independent exponentials evaluated as interleaved straight-line code with a single rare-path test, so the
// two 12-deep FMA chains overlap in the fp64 pipe
__device__ __forceinline__ void exp_fast2(double a0, double a1, double &e0, double &e1) {
  const double magic = 6755399441055744.0;
  const double t0 = fma(a0, c_exp_red[0], magic), t1 = fma(a1, c_exp_red[0], magic);
  const int n0 = __double2loint(t0), n1 = __double2loint(t1);
  const double f0 = t0 - magic, f1 = t1 - magic;
  double r0 = fma(f0, -c_exp_red[1], a0), r1 = fma(f1, -c_exp_red[1], a1);
  r0 = fma(f0, -c_exp_red[2], r0);
  r1 = fma(f1, -c_exp_red[2], r1);
  double q0 = c_exp_poly[0], q1 = c_exp_poly[0];
#pragma unroll
  for (int k = 1; k < 10; k++) {
    q0 = fma(q0, r0, c_exp_poly[k]);
    q1 = fma(q1, r1, c_exp_poly[k]);
  }
  const double p0 = fma(r0 * r0, q0, r0) + 1.0, p1 = fma(r1 * r1, q1, r1) + 1.0;
  e0 = __hiloint2double(__double2hiint(p0) + (n0 << 20), __double2loint(p0));
  e1 = __hiloint2double(__double2hiint(p1) + (n1 << 20), __double2loint(p1));
  if (__builtin_expect(!(fabs(a0) <= 700.0 && fabs(a1) <= 700.0), 0)) {  // never reached by physical velocities
    e0 = exp_slow(a0);
    e1 = exp_slow(a1);
  }
}

// Compile-time configuration of the model switches.  CFG < 0: read them from the kernel arguments (generic
// instantiation); CFG >= 0: bit 0 deltaf, bit 1 linear, bit 2 right_frac (iptclshape 1,2), bit 3 all constant
// divisors are powers of two.  The specialised instantiations drop every select / uniform branch on them.
template <int CFG> struct Cfg {
  static __device__ __forceinline__ bool deltaf(const int a) { return CFG < 0 ? a != 0 : (CFG & 1) != 0; }
  static __device__ __forceinline__ bool linear(const int a) { return CFG < 0 ? a != 0 : (CFG & 2) != 0; }
  static __device__ __forceinline__ bool right_frac(const int a) { return CFG < 0 ? a != 0 : (CFG & 4) != 0; }
  // (the TOLERANCE configurations keep this switch at run time: only the strict v update still divides by m)
  static __device__ __forceinline__ bool pow2(const int a) { return (CFG < 0 || (CFG & 32)) ? a != 0 : (CFG & 8) != 0; }
  // bit 4 (with bit 3): T/m = T2/m = m = T = 1, so x / c == x for those divisors and only 2T/m = 2 remains
  static constexpr bool unit = CFG >= 0 && (CFG & 16) != 0;
  // bit 5: PIC1DP_ARITH_TOLERANCE -- the w path (tmp2 and the w update) in the algebraically reduced form below;
  // index, weights, gather, x and v keep the reference's operation order
  static constexpr bool tol = CFG >= 0 && (CFG & 32) != 0;
};

// Per-species constants, all evaluated on the host in IEEE double exactly as the Fortran compiler folds them.
struct SpeciesConst {
  double Z, m, T;           // charge, mass, temperature
  double n, omn, v0;        // density, 1-density, v0
  double Tm, T2m;           // T/m, T2/m
  double twoTm, twoT2m;     // 2*T/m, 2*T2/m
  double sqTm, sqT2m;       // sqrt(T/m), sqrt(T2/m)
  // exact reciprocals, used only when every divisor above is a power of two (x/c == x*(1/c) bit for bit)
  double i_m, i_T, i_Tm, i_T2m, i_twoTm, i_twoT2m, i_sqTm, i_sqT2m;
  int pow2;  // every divisor above is a power of two
  int unit;  // m = T = T/m = T2/m = 1 (so 2T/m = 2T2/m = 2)
  // PIC1DP_ARITH_TOLERANCE constants (host-evaluated): bump-on-tail tmp2 = (A v + B (v - v0) r) / (C + D r) with
  // r = exp(v^2 h1 - (v - v0)^2 h2); two-stream2 tmp2 = (vp + vm r) / (1 + r) * mT with r = exp(v k2)
  double tolA, tolB, tolC, tolD, tolh1, tolh2, tolk2, tolmT;
  double Zm;  // Z / m
};

struct ParticleArgs {
  // current state (midpoint state at irk == 2)
  const double *x_cur, *v_cur, *w_cur, *p;
  // start-of-step state (x_bak ...; equal to *_cur at irk == 1 and then not read)
  const double *x_bak, *v_bak, *w_bak;
  // outputs (at irk == 2 these alias *_bak: each thread reads its own element before writing it)
  double *x_out, *v_out, *w_out;
  const double *dep_src;   // deposit-only kernel: w (delta-f) or p (full-f)
  const double *E;         // field_electric, nx
  double *partial;         // per-CTA private grids of this species, [gridDim.x][nx]
  unsigned long long *noob;
  int64_t np;
  int nx;
  double lx, rlx, rnx, dt; // rlx = RN(1/lx); dt is already halved at irk == 1 (src/pic1dp_interaction.F90:179)
  SpeciesConst c;
  int deltaf, linear, right_frac;
  // fixed-point deposit (DEP_FIXED): high word of max |deposit source| seen so far for this species (device-resident so
  // that a captured step graph can be replayed; raised with atomicMax by the kernels), and the overflow counter
  unsigned *dep_wmax_hi;
  unsigned long long *dep_overflow;
};

// ---- periodic wrap: px = mod(px, lx); if (px < 0) px = px + lx  (src/pic1dp_interaction.F90:102-104) ----
// fmod is exact; for lx <= x < 2 lx it equals x - lx (exact by Sterbenz), for -lx < x < 0 it returns x.
__device__ __forceinline__ double wrap_x(double x, double lx) {
  double xw = (x >= lx) ? dsub(x, lx) : x;
  xw = (x < 0.0) ? dadd(x, lx) : xw;
  // more than one box length away (or NaN): the general definition
  if (__builtin_expect(!(x > -lx && x < dadd(lx, lx)), 0)) xw = wrap_slow(x, lx);
  return xw;
}

struct Shape {
  int ix, ixr;
  double sl, sr;
};

// ---- weights: sx = x/lx*nx; ix = floor(sx); s = 1-(sx-ix)  (src/pic1dp_interaction.F90:106-108, :250-252;
// matrix modes src/pic1dp_particle.F90:312-323 use `frac` as the right weight) ----
__device__ __forceinline__ Shape shape_of(double x, double lx, double rlx, double rnx, int nx, int right_frac,
                                          bool &oob) {
  Shape s;
  // x >= 0 here (wrapped); anything else takes div_const's generic path and is caught by the range check below
  const double sx = dmul(div_const(x, lx, rlx), rnx);
  int ix = __double2int_rd(sx);
  double frac = dsub(sx, (double)ix);
  double sl = dsub(1.0, frac);
  if ((unsigned)ix >= (unsigned)nx) {  // x wrapped to exactly lx (reference writes out of bounds here)
    oob = true;
    ix = 0;
    sl = 1.0;
    frac = 0.0;
  }
  s.ix = ix;
  s.sl = sl;
  s.sr = right_frac ? frac : dsub(1.0, sl);
  int ixr = ix + 1;
  if (ixr > nx - 1) ixr = 0;
  s.ixr = ixr;
  return s;
}

// which constant divisors equal 2 (rather than 1) in the UNIT specialisation
struct DivIsTwo {
  static constexpr bool m = false, T = false, Tm = false, T2m = false, sqTm = false, sqT2m = false;
  static constexpr bool twoTm = true, twoT2m = true;
};

// ---- -d f0/dv / f0  (src/pic1dp_interaction.F90:275-326) ----
template <int DIST, bool POW2, bool UNIT>
__device__ __forceinline__ double dlnf0_impl(const SpeciesConst &c, double v) {
#define DIVC(x, name) (UNIT && !DivIsTwo::name ? (x) : UNIT ? dmul((x), 0.5) : POW2 ? dmul((x), c.i_##name) : ddiv((x), c.name))
  if (DIST == 1) {  // two-stream1 :276
    return dsub(v, ddiv(2.0, v));
  } else if (DIST == 2) {  // two-stream2 :278-292
    const double vp = dadd(v, c.v0), vm = dsub(v, c.v0);
    double ep, em;
    exp_fast2(-DIVC(dmul(vp, vp), twoTm), -DIVC(dmul(vm, vm), twoTm), ep, em);
    const double num = dadd(dmul(vp, ep), dmul(vm, em));
    const double den = dadd(ep, em);
    double r = div_pos(num, den);
    if (!UNIT) r = dmul(r, c.m);
    return DIVC(r, T);
  } else if (DIST == 3) {  // bump-on-tail :294-321
    const double vm = dsub(v, c.v0);
    double e1, e2;
    exp_fast2(-DIVC(dmul(v, v), twoTm), -DIVC(dmul(vm, vm), twoT2m), e1, e2);
    const double a = DIVC(dmul(DIVC(dmul(c.n, v), Tm), e1), sqTm);
    const double b = DIVC(dmul(DIVC(dmul(c.omn, vm), T2m), e2), sqT2m);
    const double num = dadd(a, b);
    const double den = dadd(DIVC(dmul(c.n, e1), sqTm), DIVC(dmul(c.omn, e2), sqT2m));
    return div_pos(num, den);
  } else {  // (shifted) Maxwellian :323-325
    return DIVC(dsub(v, c.v0), Tm);
  }
#undef DIVC
}

template <int DIST, int CFG>
__device__ __forceinline__ double dlnf0(const SpeciesConst &c, double v) {
  if (Cfg<CFG>::unit) return dlnf0_impl<DIST, true, true>(c, v);
  return Cfg<CFG>::pow2(c.pow2) ? dlnf0_impl<DIST, true, false>(c, v) : dlnf0_impl<DIST, false, false>(c, v);
}

// ---- gather + push of one marker (src/pic1dp_interaction.F90:250-338) ----
template <int DIST, int CFG>
__device__ __forceinline__ void push_one(const ParticleArgs &a, const double *sE, double x, double v, double w,
                                         double p, double xb, double vb, double wb, double &xo, double &vo,
                                         double &wo) {
  typedef Cfg<CFG> F;
  bool oob = false;  // x == lx exactly was already counted by the deposit that produced this x
  const Shape s = shape_of(x, a.lx, a.rlx, a.rnx, a.nx, F::right_frac(a.right_frac), oob);
  const double electric = dadd(dmul(sE[s.ix], s.sl), dmul(sE[s.ixr], s.sr));  // :254-257
  xo = dadd(xb, dmul(a.dt, v));                                              // :261
  wo = w;
  if (F::deltaf(a.deltaf)) {
    const double tmp1 = F::linear(a.linear) ? dmul(p, electric) : dmul(dsub(p, w), electric);  // :268-272
    const double tmp2 = dlnf0<DIST, CFG>(a.c, v);
    double t = dmul(dmul(dmul(a.dt, tmp1), tmp2), a.c.Z);              // :329
    if (!F::unit) t = F::pow2(a.c.pow2) ? dmul(t, a.c.i_m) : ddiv(t, a.c.m);  // :330
    wo = dadd(wb, t);
  }
  vo = v;
  if (!F::linear(a.linear)) {
    double t = dmul(dmul(a.dt, electric), a.c.Z);  // :336
    if (!F::unit) t = F::pow2(a.c.pow2) ? dmul(t, a.c.i_m) : ddiv(t, a.c.m);
    vo = dadd(vb, t);
  }
}

// ------------------------------------------------------------------------------------------------------------
// N-wide fast forms of the functions above.  The per-marker code is a chain of dependent fp64 operations (exact
// division, two exponentials, another division); processing the N = 2 markers of a thread as explicitly interleaved
// straight-line code lets the scheduler overlap the independent chains instead of stalling on fixed-latency
// dependencies.  These forms contain NO rare path at all: every operation that could need the IEEE routine (division
// next to a rounding midpoint or out of range, |exp argument| > 700, x more than a box length away) only ORs a flag
// into `rare`, and the caller redoes a flagged marker pair with the scalar functions above (push_one, wrap_x,
// shape_of), which hold the rare paths.  Without call sites in the hot loop ptxas keeps the polynomial and kernel
// constants in uniform registers and the body spill-free (measured: 420 -> ~360 issued instructions per tile step).
// When no flag is raised every lane has performed exactly the scalar sequence of operations, so the results are
// bit-identical to the scalar functions.
// ------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void div_const_n(const double (&a)[N], double b, double y, double (&q)[N], bool &rare) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double q0 = a[k] * y;
    const double r = fma(-q0, b, a[k]);
    q[k] = fma(r, y, q0);
    rare = rare | !(a[k] >= 0x1p-800) | div_suspect(a[k], b, q[k]);  // no short-circuit: predicate logic, no branches
  }
}

template <int N>
__device__ __forceinline__ void div_pos_n(const double (&a)[N], const double (&b)[N], double (&q)[N], bool &rare) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const unsigned eb = (unsigned)(__double2hiint(b[k]) & 0x7ff00000) - (523u << 20);
    const unsigned ea = (unsigned)(__double2hiint(a[k]) & 0x7ff00000) - (523u << 20);
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b[k]));
    double e = fma(-b[k], y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
#if PIC1DP_NEWTON2
    e = fma(-b[k], y, 1.0);
    y = fma(y, e, y);
#endif
    const double q0 = a[k] * y;
    const double r = fma(-q0, b[k], a[k]);
    q[k] = fma(r, y, q0);
    rare = rare | ((eb | ea) >= (1000u << 20)) | !(b[k] > 0.0) | div_suspect(a[k], b[k], q[k]);
  }
}

// e[k] = exp(-na[k]): callers hold the positive quantity v^2/(2T/m); taking it un-negated keeps the negation inside the
// FMA operand modifiers
template <int M>
__device__ __forceinline__ void exp_fast_neg_n(const double (&na)[M], double (&e)[M], bool &rare) {
  const double magic = 6755399441055744.0;
  double r[M], q[M];
  int n[M];
#pragma unroll
  for (int k = 0; k < M; k++) {
    const double t = fma(-na[k], c_exp_red[0], magic);
    n[k] = __double2loint(t);
    const double f = t - magic;
    r[k] = fma(f, -c_exp_red[1], -na[k]);
    r[k] = fma(f, -c_exp_red[2], r[k]);
    q[k] = c_exp_poly[0];
    rare = rare | !(fabs(na[k]) <= 700.0);
  }
#pragma unroll
  for (int j = 1; j < 10; j++) {
#pragma unroll
    for (int k = 0; k < M; k++) q[k] = fma(q[k], r[k], c_exp_poly[j]);
  }
#pragma unroll
  for (int k = 0; k < M; k++) {
    const double pr = fma(r[k] * r[k], q[k], r[k]) + 1.0;
    e[k] = __hiloint2double(__double2hiint(pr) + (n[k] << 20), __double2loint(pr));
  }
}

template <int N>
__device__ __forceinline__ void wrap_n(double (&x)[N], double lx, bool &rare) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    double xw = (x[k] >= lx) ? dsub(x[k], lx) : x[k];
    xw = (x[k] < 0.0) ? dadd(x[k], lx) : xw;
    rare = rare | !((x[k] > -lx) & (x[k] < dadd(lx, lx)));  // more than a box length away (or NaN): fmod
    x[k] = xw;
  }
}

template <int N>
__device__ __forceinline__ void shape_n(const double (&x)[N], double lx, double rlx, double rnx, int nx, int right_frac,
                                        Shape (&s)[N], bool (&oob)[N], bool &rare) {
  double q[N];
  div_const_n<N>(x, lx, rlx, q, rare);
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double sx = dmul(q[k], rnx);
    int ix = __double2int_rd(sx);
    double frac = dsub(sx, (double)ix);
    double sl = dsub(1.0, frac);
    oob[k] = (unsigned)ix >= (unsigned)nx;
    if (oob[k]) {
      ix = 0;
      sl = 1.0;
      frac = 0.0;
    }
    s[k].ix = ix;
    s[k].sl = sl;
    s[k].sr = right_frac ? frac : dsub(1.0, sl);
    int ixr = ix + 1;
    if (ixr > nx - 1) ixr = 0;
    s[k].ixr = ixr;
  }
}

template <int DIST, bool POW2, bool UNIT, int N>
__device__ __forceinline__ void dlnf0_impl_n(const SpeciesConst &c, const double (&v)[N], double (&out)[N], bool &rare) {
#define DIVC(x, name) (UNIT && !DivIsTwo::name ? (x) : UNIT ? dmul((x), 0.5) : POW2 ? dmul((x), c.i_##name) : ddiv((x), c.name))
  if (DIST == 1) {
#pragma unroll
    for (int k = 0; k < N; k++) out[k] = dsub(v[k], ddiv(2.0, v[k]));
  } else if (DIST == 2) {
    double arg[2 * N], e[2 * N], num[N], den[N], r[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double vp = dadd(v[k], c.v0), vm = dsub(v[k], c.v0);
      arg[2 * k] = DIVC(dmul(vp, vp), twoTm);
      arg[2 * k + 1] = DIVC(dmul(vm, vm), twoTm);
    }
    exp_fast_neg_n<2 * N>(arg, e, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double vp = dadd(v[k], c.v0), vm = dsub(v[k], c.v0);
      num[k] = dadd(dmul(vp, e[2 * k]), dmul(vm, e[2 * k + 1]));
      den[k] = dadd(e[2 * k], e[2 * k + 1]);
    }
    div_pos_n<N>(num, den, r, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      double t = r[k];
      if (!UNIT) t = dmul(t, c.m);
      out[k] = DIVC(t, T);
    }
  } else if (DIST == 3) {
    double arg[2 * N], e[2 * N], num[N], den[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double vm = dsub(v[k], c.v0);
      arg[2 * k] = DIVC(dmul(v[k], v[k]), twoTm);
      arg[2 * k + 1] = DIVC(dmul(vm, vm), twoT2m);
    }
    exp_fast_neg_n<2 * N>(arg, e, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double vm = dsub(v[k], c.v0), e1 = e[2 * k], e2 = e[2 * k + 1];
      const double a = DIVC(dmul(DIVC(dmul(c.n, v[k]), Tm), e1), sqTm);
      const double b = DIVC(dmul(DIVC(dmul(c.omn, vm), T2m), e2), sqT2m);
      num[k] = dadd(a, b);
      den[k] = dadd(DIVC(dmul(c.n, e1), sqTm), DIVC(dmul(c.omn, e2), sqT2m));
    }
    div_pos_n<N>(num, den, out, rare);
  } else {
#pragma unroll
    for (int k = 0; k < N; k++) out[k] = DIVC(dsub(v[k], c.v0), Tm);
  }
#undef DIVC
}

// a / b without the correctly-rounded guarantee (<= ~1 ulp): reciprocal seed, one cubic Newton step, one residual correction.
// TOLERANCE arithmetic only; operands far outside the normal range still raise the flag.
template <int N>
__device__ __forceinline__ void div_fast_n(const double (&a)[N], const double (&b)[N], double (&q)[N], bool &rare) {
#pragma unroll
  for (int k = 0; k < N; k++) {
    const unsigned eb = (unsigned)(__double2hiint(b[k]) & 0x7ff00000) - (523u << 20);
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b[k]));
    double e = fma(-b[k], y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
#if PIC1DP_NEWTON2
    e = fma(-b[k], y, 1.0);
    y = fma(y, e, y);
#endif
    const double q0 = a[k] * y;
    q[k] = fma(fma(-q0, b[k], a[k]), y, q0);
    rare = rare | (eb >= (1000u << 20)) | !(b[k] > 0.0);
  }
}

// -d ln f0 / dv in PIC1DP_ARITH_TOLERANCE: numerator and denominator of src/pic1dp_interaction.F90:294-321 (bump-on-
// tail) / :278-292 (two-stream2) divided through by the first exponential, so that ONE exponential of the difference
// of the two arguments remains; all constant divisors are folded into host-evaluated coefficients and the sums are
// fused multiply-adds.  Algebraically identical to the reference expression; differs from it by a few ulp.
template <int DIST, int N>
__device__ __forceinline__ void dlnf0_tol_n(const SpeciesConst &c, const double (&v)[N], double (&out)[N], bool &rare) {
  double nd[N], r[N], num[N], den[N];
  if (DIST == 3) {
    double vm[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
      vm[k] = v[k] - c.v0;
      nd[k] = fma(vm[k] * vm[k], c.tolh2, -(v[k] * v[k]) * c.tolh1);  // -(v^2 h1 - vm^2 h2)
    }
    exp_fast_neg_n<N>(nd, r, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      num[k] = fma(c.tolB * vm[k], r[k], c.tolA * v[k]);
      den[k] = fma(c.tolD, r[k], c.tolC);
    }
    div_fast_n<N>(num, den, out, rare);
  } else {  // DIST == 2
#pragma unroll
    for (int k = 0; k < N; k++) nd[k] = -(v[k] * c.tolk2);
    exp_fast_neg_n<N>(nd, r, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      num[k] = fma(v[k] - c.v0, r[k], v[k] + c.v0);
      den[k] = r[k] + 1.0;
    }
    div_fast_n<N>(num, den, out, rare);
#pragma unroll
    for (int k = 0; k < N; k++) out[k] *= c.tolmT;
  }
}

// gather + push of N markers (src/pic1dp_interaction.F90:250-338), lane by lane the same operations as push_one
template <int DIST, int CFG, int N>
__device__ __forceinline__ void push_n(const ParticleArgs &a, const double *sE, const double (&x)[N],
                                       const double (&v)[N], const double (&w)[N], const double (&p)[N],
                                       const double (&xb)[N], const double (&vb)[N], const double (&wb)[N],
                                       double (&xo)[N], double (&vo)[N], double (&wo)[N], bool &rare) {
  typedef Cfg<CFG> F;
  Shape s[N];
  bool oob[N];
  shape_n<N>(x, a.lx, a.rlx, a.rnx, a.nx, F::right_frac(a.right_frac), s, oob, rare);
  double electric[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    electric[k] = dadd(dmul(sE[s[k].ix], s[k].sl), dmul(sE[s[k].ixr], s[k].sr));  // :254-257
    xo[k] = dadd(xb[k], dmul(a.dt, v[k]));                                          // :261
    wo[k] = w[k];
    vo[k] = v[k];
  }
  if (F::tol && (DIST == 2 || DIST == 3)) {   // delta-f nonlinear by construction of the tolerance configurations
    double tmp2[N];
    dlnf0_tol_n<DIST, N>(a.c, v, tmp2, rare);
    const double cz = a.dt * a.c.Zm;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double tmp1 = dmul(dsub(p[k], w[k]), electric[k]);   // :271, as written
      wo[k] = fma(tmp1 * cz, tmp2[k], wb[k]);                    // :329-330 with dt Z / m folded and one rounding less
    }
  } else if (F::deltaf(a.deltaf)) {
    double tmp2[N];
    if (F::unit)
      dlnf0_impl_n<DIST, true, true, N>(a.c, v, tmp2, rare);
    else if (F::pow2(a.c.pow2))
      dlnf0_impl_n<DIST, true, false, N>(a.c, v, tmp2, rare);
    else
      dlnf0_impl_n<DIST, false, false, N>(a.c, v, tmp2, rare);
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double tmp1 = F::linear(a.linear) ? dmul(p[k], electric[k]) : dmul(dsub(p[k], w[k]), electric[k]);  // :268-272
      double t = dmul(dmul(dmul(a.dt, tmp1), tmp2[k]), a.c.Z);                          // :329
      if (!F::unit) t = F::pow2(a.c.pow2) ? dmul(t, a.c.i_m) : ddiv(t, a.c.m);           // :330
      wo[k] = dadd(wb[k], t);
    }
  }
  if (!F::linear(a.linear)) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      double t = dmul(dmul(a.dt, electric[k]), a.c.Z);  // :336
      if (!F::unit) t = F::pow2(a.c.pow2) ? dmul(t, a.c.i_m) : ddiv(t, a.c.m);
      vo[k] = dadd(vb[k], t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Deposit strategies.  Each adds (sl*w) to cell ix and (sr*w) to cell ixr of an on-chip grid.
// ------------------------------------------------------------------------------------------------------------
enum { DEP_SMEM_ATOMIC = 1, DEP_GLOBAL_RED = 2, DEP_WARP_PRIVATE = 3, DEP_FIXED = 4 };

template <int DEP>
struct Depositor;

// every depositor: prescale(q) is applied to the deposit source before the weights (identity except DEP_FIXED)
#define PIC1DP_DEP_NO_PRESCALE \
  __device__ __forceinline__ double prescale(double q) { return q; }
// add2: the two markers of a thread in one call (same validity); the CAS depositors overlap the two round trips
#define PIC1DP_DEP_ADD2_SERIAL                                                                                     \
  __device__ __forceinline__ void add2(int ix0, int ixr0, double a0, double b0, int ix1, int ixr1, double a1, double b1, \
                                       bool valid) {                                                               \
    add(ix0, ixr0, a0, b0, valid);                                                                                 \
    add(ix1, ixr1, a1, b1, valid);                                                                                 \
  }

// 128-bit compare-and-swap on a shared-memory slot: returns the previous content in (f0, f1)
__device__ __forceinline__ void cas128(unsigned addr, unsigned long long e0, unsigned long long e1, unsigned long long d0,
                                       unsigned long long d1, unsigned long long &f0, unsigned long long &f1) {
  asm volatile(
      "{\n\t.reg .b128 c, d, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 d, {%4, %5};\n\t"
      "atom.shared.cas.b128 o, [%6], c, d;\n\tmov.b128 {%0, %1}, o;\n\t}"
      : "=l"(f0), "=l"(f1)
      : "l"(e0), "l"(e1), "l"(d0), "l"(d1), "r"(addr)
      : "memory");
}

// per-CTA shared grid of pairs {sum of left weights landing in cell j, sum of right weights of markers whose LEFT
// cell is j}: both contributions of a marker go to one 16-byte slot, so one ATOMS.CAS.128 loop replaces two 64-bit
// CAS loops (sm_100a has no native fp64 shared atomic add).  rho[j] = pair[j].x + pair[j-1].y at flush time.
template <>
struct Depositor<DEP_SMEM_ATOMIC> {
  double *g;  // pair grid, 2*nx doubles
  PIC1DP_DEP_NO_PRESCALE
  __device__ __forceinline__ void add(int ix, int ixr, double a, double b, bool valid) {
    (void)ixr;
    if (!valid) return;
    double2 *slot = reinterpret_cast<double2 *>(g) + ix;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(slot);
#if PIC1DP_CAS_RELOAD
    // a failed attempt re-reads the slot instead of recycling the value the CAS returned (saves four register moves per
    // attempt, but the slot load can no longer be hoisted above the weight arithmetic: measured 35 % SLOWER)
    for (;;) {
      const double2 old = *slot;
      const unsigned long long e0 = __double_as_longlong(old.x), e1 = __double_as_longlong(old.y);
      unsigned long long f0, f1;
      cas128(addr, e0, e1, __double_as_longlong(dadd(old.x, a)), __double_as_longlong(dadd(old.y, b)), f0, f1);
      if (f0 == e0 && f1 == e1) break;
    }
#else
    double2 old = *slot;   // outside the loop: ptxas hoists this load above the weight arithmetic
    unsigned long long e0 = __double_as_longlong(old.x), e1 = __double_as_longlong(old.y), f0, f1;
    // first attempt peeled by hand (see Depositor<DEP_FIXED>::add: no YIELD on the success path)
    cas128(addr, e0, e1, __double_as_longlong(dadd(old.x, a)), __double_as_longlong(dadd(old.y, b)), f0, f1);
#pragma unroll
    for (int r = 0; r < PIC1DP_CAS_UNROLL; r++) {   // straight-line retries (no YIELD)
      if (f0 == e0 && f1 == e1) return;
      e0 = f0;
      e1 = f1;
      cas128(addr, e0, e1, __double_as_longlong(dadd(__longlong_as_double(e0), a)),
             __double_as_longlong(dadd(__longlong_as_double(e1), b)), f0, f1);
    }
    if (__builtin_expect(!(f0 == e0 && f1 == e1), 0)) {
      do {
        e0 = f0;
        e1 = f1;
        cas128(addr, e0, e1, __double_as_longlong(dadd(__longlong_as_double(e0), a)),
               __double_as_longlong(dadd(__longlong_as_double(e1), b)), f0, f1);
      } while (!(f0 == e0 && f1 == e1));
    }
#endif
  }
  // Both markers of a thread: the two slot loads, then the two CAS, are issued back to back so that their shared-memory
  // round trips overlap (the serial form exposes LDS -> DADD -> CAS -> compare twice: 22 % of the stall samples of the
  // irk = 1 kernel); a failed first attempt falls into the ordinary retry loop.  Equal slots are fine: the second CAS
  // then fails against the first one's update and retries.
  __device__ __forceinline__ void add2(int ix0, int ixr0, double a0, double b0, int ix1, int ixr1, double a1, double b1,
                                       bool valid) {
    (void)ixr0;
    (void)ixr1;
    if (!valid) return;
    double2 *s0 = reinterpret_cast<double2 *>(g) + ix0, *s1 = reinterpret_cast<double2 *>(g) + ix1;
    const unsigned ad0 = (unsigned)__cvta_generic_to_shared(s0), ad1 = (unsigned)__cvta_generic_to_shared(s1);
    const double2 o0 = *s0, o1 = *s1;
    unsigned long long e00 = __double_as_longlong(o0.x), e01 = __double_as_longlong(o0.y);
    unsigned long long e10 = __double_as_longlong(o1.x), e11 = __double_as_longlong(o1.y);
    unsigned long long f00, f01, f10, f11;
    cas128(ad0, e00, e01, __double_as_longlong(dadd(o0.x, a0)), __double_as_longlong(dadd(o0.y, b0)), f00, f01);
    cas128(ad1, e10, e11, __double_as_longlong(dadd(o1.x, a1)), __double_as_longlong(dadd(o1.y, b1)), f10, f11);
    while (!(f00 == e00 && f01 == e01)) {
      e00 = f00;
      e01 = f01;
      cas128(ad0, e00, e01, __double_as_longlong(dadd(__longlong_as_double(e00), a0)),
             __double_as_longlong(dadd(__longlong_as_double(e01), b0)), f00, f01);
    }
    while (!(f10 == e10 && f11 == e11)) {
      e10 = f10;
      e11 = f11;
      cas128(ad1, e10, e11, __double_as_longlong(dadd(__longlong_as_double(e10), a1)),
             __double_as_longlong(dadd(__longlong_as_double(e11), b1)), f10, f11);
    }
  }
};

// fire-and-forget RED.ADD.F64 into this CTA's private global (L2-resident) grid
template <>
struct Depositor<DEP_GLOBAL_RED> {
  double *g;
  PIC1DP_DEP_NO_PRESCALE
  PIC1DP_DEP_ADD2_SERIAL
  __device__ __forceinline__ void add(int ix, int ixr, double a, double b, bool valid) {
    if (valid) {
      atomicAdd(&g[ix], a);  // result unused -> REDG.E.ADD.F64
      atomicAdd(&g[ixr], b);
    }
  }
};

// per-warp private shared grid, no atomics.  Lanes of one instruction that hit the same cell are found with
// MATCH.ANY and take turns in ascending lane order (round r: the r-th lane of every group does its
// read-modify-write), so each cell receives its contributions in marker order.  Fixed marker->lane mapping + fixed
// order => bitwise run-to-run deterministic.  Must be called by all 32 lanes (invalid lanes pass valid=false).
template <>
struct Depositor<DEP_WARP_PRIVATE> {
  double *g;  // this warp's grid
  PIC1DP_DEP_NO_PRESCALE
  PIC1DP_DEP_ADD2_SERIAL
  __device__ __forceinline__ void add(int ix, int ixr, double a, double b, bool valid) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    if (!valid) {  // joins the ix == 0 group with zero contributions; ixr must agree with that group's (nx >= 2)
      ix = 0;
      ixr = 1;
      a = 0.0;
      b = 0.0;
    }
    const unsigned peers = __match_any_sync(full, ix);
    const int myturn = __popc(peers & ((1u << lane) - 1u));
    const int rounds = (int)__reduce_max_sync(full, (unsigned)__popc(peers));  // warp-uniform, 1 when no lane collides
    for (int r = 0; r < rounds; r++) {
      if (myturn == r) g[ix] = dadd(g[ix], a);
      __syncwarp();
    }
    for (int r = 0; r < rounds; r++) {  // same groups: equal ix means equal ixr
      if (myturn == r) g[ixr] = dadd(g[ixr], b);
      __syncwarp();
    }
  }
};


// ------------------------------------------------------------------------------------------------------------
// Fixed-point deposit: bitwise reproducible for every grid the pair layout fits (nx <= ~9600), at the speed of the
// 128-bit CAS deposit.  Same {left, right} pair slots, but the contributions are converted to 64-bit integers
// (value * 2^e, round to nearest) and added as integers inside the CAS loop: integer addition is associative, so the
// slot sums do not depend on the order in which the warps arrive.  The reference's own deposit is sequential, hence
// deterministic (src/pic1dp_interaction.F90:96-114); this is the order-independent counterpart.
//   scale   : 2^e, chosen per launch so that NO slot can overflow whatever the marker distribution: a contribution is
//             below B = 2^(62 - ceil(log2(n_cta))) (n_cta = markers this CTA handles; capped at 2^50), hence a slot sum
//             stays below 2^62 even if every marker of the CTA hit the same cell; and sources up to 2^(E+3) map below
//             B, E = exponent of the running maximum of |source| (max |w| so far, kept on the device), i.e. >= 4x
//             headroom for growth within the substep.  (Two's-complement partial sums may even wrap: only the final
//             sum has to be representable.)  No overflow checks, spills or barriers in the persistent loop -- an earlier
//             version flushed the integer grid every 32 tile steps and lost 30 % to the CTA-wide barriers (the warps
//             drift many iterations apart and the laggards then run alone).
//   quantum : 2^-e <= 8 max|w| / B: 2^-39 max|w| at 1e8 markers per GPU (6.8e5 per CTA), 2^-47 max|w| below 4096
//             markers per CTA.  The rounding error of a cell sum of n contributions is ~sqrt(n / 12) quanta: for the
//             bench state (2e5 contributions per cell) 2e-10 max|w| against cell sums of ~3e4 max|w|, i.e. 1e-14 of
//             max|rho|; it stays below 1e-12 of max|rho| unless the density is pure noise at > 2e8 markers per GPU.
//   overflow: |source| >= 2^(E+4) (growth by more than 8x .. 16x inside one substep) is outside the planned headroom; it
//             is counted in dep_overflow and reported by the next synchronising call.
// ------------------------------------------------------------------------------------------------------------
template <>
struct Depositor<DEP_FIXED> {
  double *g;      // pair grid, 2*nx int64 viewed as doubles by dep_setup
  double scale;   // 2^e
  double inv;     // 2^-e
  unsigned bound_hi;  // high word of the largest admissible |source|: at or above it the headroom is exhausted
  unsigned seen_hi;   // high word of the largest |source| this thread has deposited
  __device__ __forceinline__ double prescale(double q) {
    seen_hi = max(seen_hi, (unsigned)__double2hiint(q) & 0x7fffffffu);
    return dmul(q, scale);  // exact (power of two)
  }
  // double -> int64, round to nearest
  static __device__ __forceinline__ long long to_fixed(double x) {
#if PIC1DP_FIXED_F2I
    return __double2ll_rn(x);
#else
    const double magic = 6755399441055744.0;   // the integer lands in the low mantissa bits of x + 1.5 * 2^52
    return __double_as_longlong(dadd(x, magic)) - __double_as_longlong(magic);
#endif
  }
  // a, b = weight * prescaled source: RN(weight * source) * 2^e exactly, then rounded to the nearest integer
#if PIC1DP_FIXED_NATIVE
  // The grid is four planes of nx 32-bit words {low left, high left, low right, high right}: a 64-bit integer add is the
  // native ATOMS.ADD of the low word (returning the old value) plus the add of (high word + carry), the carry being
  // known to exactly the thread whose add wrapped the low word.  Four native adds, no compare-and-swap loop, no retry
  // path: on B200 the native 32-bit shared-memory add retires ~9 lane-operations per clock per SM at random banks, the
  // LDS.128 + CAS.128 pair 0.7 (tools_py3/dev/smem_scatter_bench.cu).  Word planes instead of 16-byte slots so that the
  // lanes of one ATOMS spread over all 32 banks (slot-major words would use 8).
  int nx;
  __device__ __forceinline__ void add(int ix, int ixr, double a, double b, bool valid) {
    (void)ixr;
    if (!valid) return;
    const double magic = 6755399441055744.0;   // the integer lands in the low mantissa bits of x + 1.5 * 2^52
    const double ta = dadd(a, magic), tb = dadd(b, magic);
    const unsigned la = (unsigned)__double2loint(ta), lb = (unsigned)__double2loint(tb);
    const unsigned ha = (unsigned)(__double2hiint(ta) - 0x43380000), hb = (unsigned)(__double2hiint(tb) - 0x43380000);
    unsigned *w = reinterpret_cast<unsigned *>(g) + ix;
    const unsigned oa = atomicAdd(w, la), ob = atomicAdd(w + 2 * nx, lb);
    atomicAdd(w + nx, ha + ((oa + la) < la ? 1u : 0u));        // results unused: fire-and-forget
    atomicAdd(w + 3 * nx, hb + ((ob + lb) < lb ? 1u : 0u));
  }
  // both markers of a thread: the four returning adds are issued before the four carries wait for them
  __device__ __forceinline__ void add2(int ix0, int ixr0, double a0, double b0, int ix1, int ixr1, double a1, double b1,
                                       bool valid) {
    (void)ixr0;
    (void)ixr1;
    if (!valid) return;
    const double magic = 6755399441055744.0;
    const double t[4] = {dadd(a0, magic), dadd(b0, magic), dadd(a1, magic), dadd(b1, magic)};
    unsigned *w0 = reinterpret_cast<unsigned *>(g) + ix0, *w1 = reinterpret_cast<unsigned *>(g) + ix1;
    unsigned *lo[4] = {w0, w0 + 2 * nx, w1, w1 + 2 * nx};
    unsigned l[4], o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      l[k] = (unsigned)__double2loint(t[k]);
      o[k] = atomicAdd(lo[k], l[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
      atomicAdd(lo[k] + nx, (unsigned)(__double2hiint(t[k]) - 0x43380000) + ((o[k] + l[k]) < l[k] ? 1u : 0u));
  }
#define PIC1DP_FIXED_HAS_ADD2 1
#else
  __device__ __forceinline__ void add(int ix, int ixr, double a, double b, bool valid) {
    (void)ixr;
    if (!valid) return;
    const unsigned long long ia = (unsigned long long)to_fixed(a), ib = (unsigned long long)to_fixed(b);
    longlong2 *slot = reinterpret_cast<longlong2 *>(g) + ix;
    const unsigned addr = (unsigned)__cvta_generic_to_shared(slot);
    const longlong2 old = *slot;   // outside the loop: hoisted above the weight arithmetic
    unsigned long long e0 = (unsigned long long)old.x, e1 = (unsigned long long)old.y, f0, f1;
    // first attempt peeled by hand: ptxas puts a YIELD into every spin-loop body, and with the attempt inside the loop
    // the warp yields its issue slot on EVERY deposit -- measured: 30 % of the kernel spent at the end-of-loop barrier
    // (warps drift apart), 4.0 instead of 2.3 ms per step.  Only genuine retries may spin.
    cas128(addr, e0, e1, e0 + ia, e1 + ib, f0, f1);
#pragma unroll
    for (int r = 0; r < PIC1DP_CAS_UNROLL; r++) {   // straight-line retries (no YIELD)
      if (f0 == e0 && f1 == e1) return;
      e0 = f0;
      e1 = f1;
      cas128(addr, e0, e1, e0 + ia, e1 + ib, f0, f1);
    }
    if (__builtin_expect(!(f0 == e0 && f1 == e1), 0)) {
      do {
        e0 = f0;
        e1 = f1;
        cas128(addr, e0, e1, e0 + ia, e1 + ib, f0, f1);
      } while (!(f0 == e0 && f1 == e1));
    }
  }
#endif
#ifndef PIC1DP_FIXED_HAS_ADD2
  PIC1DP_DEP_ADD2_SERIAL
#endif
};

// scale of this launch: wmax_hi = high word of the running max |source|, ncap = markers this CTA handles
__device__ __forceinline__ void fixed_scale(Depositor<DEP_FIXED> &dep, const unsigned wmax_hi, const long long ncap, const int nx) {
#if PIC1DP_FIXED_NATIVE
  dep.nx = nx;
#else
  (void)nx;
#endif
  int ex = (int)((wmax_hi >> 20) & 0x7ff);   // biased exponent of the maximum
  if (ex < 100) ex = 100;                     // all-zero / denormal sources: any scale works, keep 2^e finite
  int lb = 62 - (64 - __clzll(ncap > 1 ? ncap - 1 : 1));   // log2 B = 62 - ceil(log2 ncap)
  if (lb > 50) lb = 50;
  const int se = 1023 + lb - (ex - 1023 + 3);   // biased exponent of 2^e: sources below 2^(E+3) map below B
  dep.seen_hi = 0;
  dep.scale = __hiloint2double(se << 20, 0);
  dep.inv = __hiloint2double((2046 - se) << 20, 0);
  dep.bound_hi = (unsigned)((ex + 4) << 20);        // sources at or above 2 * 2^(E+3): beyond the planned bound
}

// end of the kernel: publish the largest source seen (the next launch scales by it) and count headroom violations
__device__ __forceinline__ void fixed_finish(const Depositor<DEP_FIXED> &dep, unsigned *wmax_hi, unsigned long long *overflow) {
  const unsigned m = __reduce_max_sync(0xffffffffu, dep.seen_hi);
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicMax(wmax_hi, m);
    if (m >= dep.bound_hi) atomicAdd(overflow, 1ULL);
  }
}

// flush of the integer pair grid into the CTA's global grid; called by every thread of the CTA after the marker loop.
// rho[j] = left[j] + right[j-1], summed as integers (exact) and converted once.
__device__ __forceinline__ void fixed_flush(double *pairs, int nx, double *my_partial, const double inv) {
  __syncthreads();
#if PIC1DP_FIXED_NATIVE
  const unsigned *w = reinterpret_cast<const unsigned *>(pairs);   // planes: low left, high left, low right, high right
  for (int j = threadIdx.x; j < nx; j += blockDim.x) {
    const int jl = (j == 0) ? nx - 1 : j - 1;  // right weights of the cell to the left (periodic, :111-112)
    const long long left = (long long)(((unsigned long long)w[nx + j] << 32) | w[j]);
    const long long right = (long long)(((unsigned long long)w[3 * nx + jl] << 32) | w[2 * nx + jl]);
    my_partial[j] = dmul((double)(left + right), inv);
  }
#else
  const longlong2 *sl = reinterpret_cast<const longlong2 *>(pairs);
  for (int j = threadIdx.x; j < nx; j += blockDim.x) {
    const int jl = (j == 0) ? nx - 1 : j - 1;  // right weights of the cell to the left (periodic, :111-112)
    my_partial[j] = dmul((double)(sl[j].x + sl[jl].y), inv);
  }
#endif
}

// Marker arrays are touched once per substep.  Measured on B200 (profiles/r01_ab_experiments.md): the default cache
// policy beats the streaming hints (.cs evict-first loads/stores cost ~3% of the step; .cg is on par with default).
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double ld1(const double *p) { return *p; }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
__device__ __forceinline__ void st1(double *p, double v) { *p = v; }
// pull the line a later tile step will read into L2 (no destination register, no scoreboard wait)
__device__ __forceinline__ void prefetch_l2(const double *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Shared-memory layout of the particle kernels: [E : nx rounded up to even] [deposit grid(s) : nx * ngrids]
// (16-byte alignment of the pair grid for the 128-bit CAS)
template <int DEP>
__device__ __forceinline__ double *dep_setup(double *smem_after_E, int nx, double *my_partial) {
  if (DEP == DEP_SMEM_ATOMIC || DEP == DEP_FIXED) {
    for (int j = threadIdx.x; j < 2 * nx; j += blockDim.x) smem_after_E[j] = 0.0;   // +0.0 is integer 0 as well
    return smem_after_E;
  } else if (DEP == DEP_WARP_PRIVATE) {
    const int nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < nx * nw; j += blockDim.x) smem_after_E[j] = 0.0;
    return smem_after_E + (size_t)(threadIdx.x >> 5) * nx;
  } else {
    return my_partial;  // zeroed by the reduce kernel of the previous substep
  }
}

template <int DEP>
__device__ __forceinline__ void dep_flush(double *smem_after_E, int nx, double *my_partial) {
  if (DEP == DEP_SMEM_ATOMIC) {
    __syncthreads();
    for (int j = threadIdx.x; j < nx; j += blockDim.x) {
      const int jl = (j == 0) ? nx - 1 : j - 1;  // right weights of the cell to the left (periodic, :111-112)
      my_partial[j] = dadd(smem_after_E[2 * j], smem_after_E[2 * jl + 1]);
    }
  } else if (DEP == DEP_WARP_PRIVATE) {
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < nx; j += blockDim.x) {
      double t = smem_after_E[j];
      for (int wq = 1; wq < nw; wq++) t = dadd(t, smem_after_E[(size_t)wq * nx + j]);  // fixed warp order
      my_partial[j] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fused substep kernel: gather + push + wrap + deposit.  Persistent grid, 2 markers per thread per iteration
// through 128-bit loads/stores.  IRK2 selects the second RK substep (reads midpoint + start-of-step state).
// FUSED=false gives the reference's push-only side effects (x left unwrapped, no deposit).
// ------------------------------------------------------------------------------------------------------------
// One tile step of one thread (markers i, i+1) comes in two forms.  push_pair_fast: both markers valid, 2-wide
// interleaved code without any rare path; when one of its operations raises the `rare` flag nothing is stored or
// deposited and the caller redoes the pair with push_pair_slow.  push_pair_slow: scalar code with the IEEE fallbacks,
// per-marker validity (tail of the arrays, or `ok` = this thread's pair needs the redo); loads its own inputs.
template <int DIST, bool IRK2, int DEP, bool FUSED, int CFG>
__device__ __forceinline__ bool push_pair_fast(const ParticleArgs &a, const double *sE, Depositor<DEP> &dep,
                                               const int64_t i, unsigned long long &noob, double2 x, double2 v, double2 w,
                                               double2 p, double2 xb, double2 vb, double2 wb) {
  typedef Cfg<CFG> F;
  const bool deltaf = F::deltaf(a.deltaf), linear = F::linear(a.linear), right_frac = F::right_frac(a.right_frac);
  if (!IRK2) {  // xb, vb, wb are ignored at irk == 1
    xb = x;
    vb = v;
    wb = w;
  } else if (!deltaf) {
    wb = w;
  }
  bool rare = false;
  Shape sd[2];
  bool od[2] = {false, false};
  const double ax[2] = {x.x, x.y}, av[2] = {v.x, v.y}, aw[2] = {w.x, w.y}, ap[2] = {p.x, p.y};
  const double axb[2] = {xb.x, xb.y}, avb[2] = {vb.x, vb.y}, awb[2] = {wb.x, wb.y};
  double axo[2], avo[2], awo[2];
  push_n<DIST, CFG, 2>(a, sE, ax, av, aw, ap, axb, avb, awb, axo, avo, awo, rare);
  if (FUSED) {
    wrap_n<2>(axo, a.lx, rare);
    shape_n<2>(axo, a.lx, a.rlx, a.rnx, a.nx, right_frac, sd, od, rare);
  }
  if (!rare) {
    st2(a.x_out + i, make_double2(axo[0], axo[1]));
    if (!linear) st2(a.v_out + i, make_double2(avo[0], avo[1]));
    if (deltaf) st2(a.w_out + i, make_double2(awo[0], awo[1]));
  }
  if (FUSED) {
    // deposit source: w (delta-f) or p (full-f)  (src/pic1dp_interaction.F90:84-91)
    const double q0 = dep.prescale(deltaf ? awo[0] : p.x), q1 = dep.prescale(deltaf ? awo[1] : p.y);
    constexpr bool OVERLAP = (DEP == DEP_SMEM_ATOMIC && (PIC1DP_ADD2 & 1)) || (DEP == DEP_FIXED && (PIC1DP_ADD2 & 2));
    if constexpr (OVERLAP) {
      dep.add2(sd[0].ix, sd[0].ixr, dmul(sd[0].sl, q0), dmul(sd[0].sr, q0),       // :110, :113
               sd[1].ix, sd[1].ixr, dmul(sd[1].sl, q1), dmul(sd[1].sr, q1), !rare);
    } else {
      dep.add(sd[0].ix, sd[0].ixr, dmul(sd[0].sl, q0), dmul(sd[0].sr, q0), !rare);  // :110, :113
      dep.add(sd[1].ix, sd[1].ixr, dmul(sd[1].sl, q1), dmul(sd[1].sr, q1), !rare);
    }
    noob += (!rare && od[0]) + (!rare && od[1]);
  }
  return rare;
}

template <int DIST, bool IRK2, int DEP, bool FUSED, int CFG>
__device__ __forceinline__ void push_pair_slow(const ParticleArgs &a, const double *sE, Depositor<DEP> &dep, const int64_t i,
                                            unsigned long long &noob, const bool ok, const int64_t end) {
  typedef Cfg<CFG> F;
  const bool deltaf = F::deltaf(a.deltaf), linear = F::linear(a.linear), right_frac = F::right_frac(a.right_frac);
  const bool need_p = deltaf || FUSED;  // full-f deposits p (src/pic1dp_interaction.F90:88-90)
  const bool v0ok = ok && i < end, v1ok = ok && i + 1 < end;   // end: one past this CTA's last marker (<= np)
  double2 x = {0.0, 0.0}, v = {0.0, 0.0}, w = {0.0, 0.0}, p = {0.0, 0.0};
  double2 xb = {0.0, 0.0}, vb = {0.0, 0.0}, wb = {0.0, 0.0};
  if (v1ok) {
    x = ld2(a.x_cur + i);
    v = ld2(a.v_cur + i);
    if (deltaf) w = ld2(a.w_cur + i);
    if (need_p) p = ld2(a.p + i);
    if (IRK2) {
      xb = ld2(a.x_bak + i);
      vb = ld2(a.v_bak + i);
      if (deltaf) wb = ld2(a.w_bak + i);
    }
  } else if (v0ok) {
    x.x = ld1(a.x_cur + i);
    v.x = ld1(a.v_cur + i);
    if (deltaf) w.x = ld1(a.w_cur + i);
    if (need_p) p.x = ld1(a.p + i);
    if (IRK2) {
      xb.x = ld1(a.x_bak + i);
      vb.x = ld1(a.v_bak + i);
      if (deltaf) wb.x = ld1(a.w_bak + i);
    }
  }
  if (!IRK2) {
    xb = x;
    vb = v;
    wb = w;
  } else if (!deltaf) {
    wb = w;
  }
  double2 xo = {0.0, 0.0}, vo = {0.0, 0.0}, wo = {0.0, 0.0};
  Shape sd[2];
  bool od[2] = {false, false};
  if (v0ok) push_one<DIST, CFG>(a, sE, x.x, v.x, w.x, p.x, xb.x, vb.x, wb.x, xo.x, vo.x, wo.x);
  if (v1ok) push_one<DIST, CFG>(a, sE, x.y, v.y, w.y, p.y, xb.y, vb.y, wb.y, xo.y, vo.y, wo.y);
  if (FUSED) {
    if (v0ok) xo.x = wrap_x(xo.x, a.lx);
    if (v1ok) xo.y = wrap_x(xo.y, a.lx);
    sd[0] = shape_of(xo.x, a.lx, a.rlx, a.rnx, a.nx, right_frac, od[0]);
    sd[1] = shape_of(xo.y, a.lx, a.rlx, a.rnx, a.nx, right_frac, od[1]);
  }
  if (v1ok) {
    st2(a.x_out + i, xo);
    if (!linear) st2(a.v_out + i, vo);
    if (deltaf) st2(a.w_out + i, wo);
  } else if (v0ok) {
    st1(a.x_out + i, xo.x);
    if (!linear) st1(a.v_out + i, vo.x);
    if (deltaf) st1(a.w_out + i, wo.x);
  }
  if (FUSED) {
    const double q0 = dep.prescale(deltaf ? wo.x : p.x), q1 = dep.prescale(deltaf ? wo.y : p.y);
    dep.add(sd[0].ix, sd[0].ixr, dmul(sd[0].sl, q0), dmul(sd[0].sr, q0), v0ok);  // :110, :113
    dep.add(sd[1].ix, sd[1].ixr, dmul(sd[1].sl, q1), dmul(sd[1].sr, q1), v1ok);
    noob += (v0ok && od[0]) + (v1ok && od[1]);
  }
}

// direct 128-bit loads of a full tile step's markers
template <bool IRK2, bool FUSED, int CFG>
__device__ __forceinline__ void load_pair(const ParticleArgs &a, const int64_t i, double2 &x, double2 &v, double2 &w,
                                          double2 &p, double2 &xb, double2 &vb, double2 &wb) {
  const bool deltaf = Cfg<CFG>::deltaf(a.deltaf);
  x = ld2(a.x_cur + i);
  v = ld2(a.v_cur + i);
  w = p = xb = vb = wb = make_double2(0.0, 0.0);
  if (deltaf) w = ld2(a.w_cur + i);
  if (deltaf || FUSED) p = ld2(a.p + i);
  if (IRK2) {
    xb = ld2(a.x_bak + i);
    vb = ld2(a.v_bak + i);
    if (deltaf) wb = ld2(a.w_bak + i);
  }
}

// the redo / tail step shared by all fused kernels: threads whose pair was flagged (or every thread of a partial tile)
// run the scalar code; the warp-private depositor needs all 32 lanes present
template <int DIST, bool IRK2, int DEP, bool FUSED, int CFG>
__device__ __forceinline__ void push_pair_redo(const ParticleArgs &a, const double *sE, Depositor<DEP> &dep,
                                               const int64_t i, unsigned long long &noob, const bool redo, const int64_t end) {
  const bool enter = (FUSED && DEP == DEP_WARP_PRIVATE) ? __any_sync(0xffffffffu, redo) : redo;
  if (__builtin_expect(enter, 0)) push_pair_slow<DIST, IRK2, DEP, FUSED, CFG>(a, sE, dep, i, noob, redo, end);
}

template <int DIST, bool IRK2, int DEP, bool FUSED, int CFG>
__global__ void __launch_bounds__(PIC1DP_MAXTHREADS, 1) k_push(const ParticleArgs a) {
  extern __shared__ __align__(16) double smem[];
  double *sE = smem;
  for (int j = threadIdx.x; j < a.nx; j += blockDim.x) sE[j] = a.E[j];
  double *my_partial = FUSED ? a.partial + (size_t)blockIdx.x * a.nx : nullptr;
  Depositor<DEP> dep;
  dep.g = FUSED ? dep_setup<DEP>(smem + ((a.nx + 1) & ~1), a.nx, my_partial) : nullptr;
  __syncthreads();

  const int64_t tile = (int64_t)blockDim.x * 2;
  unsigned long long noob = 0;
  const bool deltaf_pf = Cfg<CFG>::deltaf(a.deltaf);
  // register-staged v of the next tile step: the fast body starts with the v-only work (both exponentials), so v is
  // the one load whose latency nothing hides; it is issued one iteration ahead (4 registers)
  constexpr bool PV = (PIC1DP_PV_MASK & ((DEP == DEP_WARP_PRIVATE ? 4 : 1) << (IRK2 ? 1 : 0))) != 0;
  double2 v_next = make_double2(0.0, 0.0);
  // tiles of 2 * blockDim markers are dealt round-robin over the persistent CTAs.  (Contiguous per-CTA ranges with a
  // constant-offset prefetch were measured slower on B200 -- 2.31 vs 2.25 ms per step in TOLERANCE mode, 2.84 vs 2.41 ms
  // with the warp-private deposit: 148 separate streams per array instead of one front moving through HBM.)
  const int64_t start = (int64_t)blockIdx.x * tile, end = a.np, stride = (int64_t)gridDim.x * tile;
  if constexpr (DEP == DEP_FIXED) fixed_scale(dep, FUSED ? *a.dep_wmax_hi : 0u, (end - start) / stride * tile + tile, a.nx);
  if (PV) {
    if (start + tile <= end) v_next = ld2(a.v_cur + start + (int64_t)threadIdx.x * 2);
  }
  for (int64_t base = start; base < end; base += stride) {
    const int64_t i = base + (int64_t)threadIdx.x * 2;
    const double *px = a.x_cur + i, *pv = a.v_cur + i, *pw = a.w_cur + i, *pp = a.p + i;   // shared by prefetch and loads
    // L2 prefetch of this thread's markers of the next tile step: the loads then hit L2 instead of HBM.
    // Measured on B200 at 1e8 markers (profiles/r01_ab_experiments.md), enabled per variant: it helps the warp-private
    // deposit in both substeps and the atomic deposit at irk=1 (1.060 -> 1.040 ms: with the rare-path-free body the
    // first use of the streamed v is the top stall) and hurts the HBM-bound atomic irk=2 kernel (1.24 -> 1.54 ms).
    if (PIC1DP_PF_MASK & ((DEP == DEP_WARP_PRIVATE ? 4 : 1) << (IRK2 ? 1 : 0))) {
      const int64_t ahead = PIC1DP_PF_DIST * stride;
      if (i + ahead + 1 < end) {
        prefetch_l2(px + ahead);
        prefetch_l2(pv + ahead);
        if (deltaf_pf) prefetch_l2(pw + ahead);
        if (deltaf_pf || FUSED) prefetch_l2(pp + ahead);
        if (IRK2) {
          prefetch_l2(a.x_bak + i + ahead);
          prefetch_l2(a.v_bak + i + ahead);
          if (deltaf_pf) prefetch_l2(a.w_bak + i + ahead);
        }
      }
    }
    const bool full = base + tile <= end;
    bool redo = true;  // partial tile: every thread takes the scalar path (validity per marker)
    if (full) {
      double2 x, v, w, p, xb, vb, wb;
      x = ld2(px);
      v = ld2(pv);
      w = p = xb = vb = wb = make_double2(0.0, 0.0);
      if (deltaf_pf) w = ld2(pw);
      if (deltaf_pf || FUSED) p = ld2(pp);
      if (IRK2) {
        xb = ld2(a.x_bak + i);
        vb = ld2(a.v_bak + i);
        if (deltaf_pf) wb = ld2(a.w_bak + i);
      }
      if (PV) {
        v = v_next;
        const int64_t nb = base + stride;
        if (nb + tile <= end) v_next = ld2(a.v_cur + nb + (int64_t)threadIdx.x * 2);
      }
      redo = push_pair_fast<DIST, IRK2, DEP, FUSED, CFG>(a, sE, dep, i, noob, x, v, w, p, xb, vb, wb);
    }
    push_pair_redo<DIST, IRK2, DEP, FUSED, CFG>(a, sE, dep, i, noob, redo, end);
  }
  if constexpr (DEP == DEP_FIXED && FUSED) {
    fixed_flush(smem + ((a.nx + 1) & ~1), a.nx, my_partial, dep.inv);
    fixed_finish(dep, a.dep_wmax_hi, a.dep_overflow);
  } else if (FUSED) {
    dep_flush<DEP>(smem + ((a.nx + 1) & ~1), a.nx, my_partial);
  }
  if (FUSED && noob) atomicAdd(a.noob, noob);
}

// ------------------------------------------------------------------------------------------------------------
// cp.async-staged variant of the fused substep kernel (delta-f nonlinear, fused; opt-in / AUTO where it measured
// faster).  The direct kernel's loads are only outstanding at the top of an iteration and with 64 registers per
// thread there is no room to hold the next tile in registers, so ~25 % of its stall samples are the first use of
// the streamed x.  Here every thread copies ITS OWN 2 markers of the next tile step into a private 16-byte slot per
// array of a 2-stage shared-memory ring with cp.async (LDGSTS.128, no registers, no L1), computes the current tile
// step from the other stage, and waits with cp.async.wait_group: slots are thread-private, so there is no mbarrier,
// no elected producer and no CTA-wide synchronisation in the loop (unlike the TMA ring below).  Tail tiles fall back
// to direct loads.  Ring: STAGES x NARR x blockDim x 16 B (128 KB at irk = 1 with 1024 threads).
// ------------------------------------------------------------------------------------------------------------
namespace cpa {
__device__ __forceinline__ void copy16(unsigned dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
}  // namespace cpa

template <int DIST, bool IRK2, int DEP, int CFG>
__global__ void __launch_bounds__(PIC1DP_MAXTHREADS, 1) k_push_cpa(const ParticleArgs a) {
  constexpr int NARR = IRK2 ? 7 : 4;  // x v w p [xb vb wb]
  extern __shared__ __align__(16) double smem[];
  double *sE = smem;
  for (int j = threadIdx.x; j < a.nx; j += blockDim.x) sE[j] = a.E[j];
  double *my_partial = a.partial + (size_t)blockIdx.x * a.nx;
  Depositor<DEP> dep;
  double *dep_base = smem + ((a.nx + 1) & ~1);
  dep.g = dep_setup<DEP>(dep_base, a.nx, my_partial);
  const int ngr = (DEP == DEP_WARP_PRIVATE) ? (int)(blockDim.x >> 5) : (DEP == DEP_SMEM_ATOMIC ? 2 : 0);
  double2 *ring = reinterpret_cast<double2 *>(dep_base + (((size_t)a.nx * ngr + 1) & ~(size_t)1));  // [2][NARR][blockDim]
  __syncthreads();

  const double *src[7] = {a.x_cur, a.v_cur, a.w_cur, a.p, a.x_bak, a.v_bak, a.w_bak};
  const int64_t tile = (int64_t)blockDim.x * 2, stride = (int64_t)gridDim.x * tile;
  const unsigned slot0 = (unsigned)__cvta_generic_to_shared(ring + threadIdx.x);
  const unsigned arr_bytes = blockDim.x * 16u, stage_bytes = NARR * arr_bytes;
  auto stage_in = [&](int64_t base, int st) {  // this thread's markers base + 2 tid, base + 2 tid + 1 of a full tile
    const int64_t i = base + (int64_t)threadIdx.x * 2;
#pragma unroll
    for (int q = 0; q < NARR; q++) cpa::copy16(slot0 + st * stage_bytes + q * arr_bytes, src[q] + i);
    cpa::commit();
  };
  unsigned long long noob = 0;
  int64_t base = (int64_t)blockIdx.x * tile;
  bool cur_staged = base + tile <= a.np;
  if (cur_staged) stage_in(base, 0);
  for (int st = 0; base < a.np; base += stride, st ^= 1) {
    const int64_t next = base + stride;
    const bool next_staged = next + tile <= a.np;
    if (next_staged) stage_in(next, st ^ 1);
    const int64_t i = base + (int64_t)threadIdx.x * 2;
    bool redo = true;
    if (cur_staged) {
      if (next_staged)
        cpa::wait<1>();
      else
        cpa::wait<0>();
      const double2 *rs = ring + (size_t)st * NARR * blockDim.x + threadIdx.x;
      const int B = blockDim.x;
      const double2 x = rs[0], v = rs[B], w = rs[2 * B], p = rs[3 * B];
      double2 xb = x, vb = v, wb = w;
      if (IRK2) {
        xb = rs[4 * B];
        vb = rs[5 * B];
        wb = rs[6 * B];
      }
      redo = push_pair_fast<DIST, IRK2, DEP, true, CFG>(a, sE, dep, i, noob, x, v, w, p, xb, vb, wb);
    }
    push_pair_redo<DIST, IRK2, DEP, true, CFG>(a, sE, dep, i, noob, redo, a.np);
    cur_staged = next_staged;
  }
  dep_flush<DEP>(dep_base, a.nx, my_partial);
  if (noob) atomicAdd(a.noob, noob);
}

// ------------------------------------------------------------------------------------------------------------
// TMA-pipelined variant of the fused substep kernel (delta-f, nonlinear; the flagship path).
// Marker tiles of 2*blockDim markers are streamed into a 2-stage shared-memory ring with 1-D bulk copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) issued by one thread one tile ahead; consumers wait on the
// stage's "full" mbarrier, pull their 2 markers into registers with 128-bit shared loads, release the stage ("empty"
// mbarrier, one arrival per warp) and run the same 2-wide code as the direct kernel.  The next tile is in flight
// while the current one is computed, which supplies the memory-level parallelism the direct kernel lacks (its
// loads are only outstanding at the top of an iteration).  Outputs go straight to global memory.
// ------------------------------------------------------------------------------------------------------------
struct PairIn2 {
  double2 x, v, w, p, xb, vb, wb;
};

namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
}  // namespace tma

template <int DIST, bool IRK2, int DEP, int CFG>
__global__ void __launch_bounds__(1024, 1) k_push_tma(const ParticleArgs a) {
  constexpr int NARR = IRK2 ? 7 : 4;        // x v w p [xb vb wb]
  constexpr int STAGES = 2;
  const int TILE = 2 * blockDim.x;          // 2 markers per thread, one 128-bit shared load per array
  extern __shared__ __align__(16) double smem[];
  double *sE = smem;
  for (int j = threadIdx.x; j < a.nx; j += blockDim.x) sE[j] = a.E[j];
  double *my_partial = a.partial + (size_t)blockIdx.x * a.nx;
  Depositor<DEP> dep;
  double *dep_base = smem + ((a.nx + 1) & ~1);
  dep.g = dep_setup<DEP>(dep_base, a.nx, my_partial);
  const int ngr = (DEP == DEP_WARP_PRIVATE) ? (int)(blockDim.x >> 5) : (DEP == DEP_SMEM_ATOMIC ? 2 : 0);
  double *ring = dep_base + (((size_t)a.nx * ngr + 1) & ~(size_t)1);  // [STAGES][NARR][TILE]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + (size_t)STAGES * NARR * TILE);
  const unsigned full0 = tma::smem_u32(bars), empty0 = tma::smem_u32(bars + STAGES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      tma::mbar_init(full0 + 8 * s, 1);
      tma::mbar_init(empty0 + 8 * s, blockDim.x >> 5);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t ntiles = (a.np + TILE - 1) / TILE;
  const int64_t my_tiles = (ntiles > (int64_t)blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const double *src[7] = {a.x_cur, a.v_cur, a.w_cur, a.p, a.x_bak, a.v_bak, a.w_bak};
  auto issue = [&](int64_t k) {  // thread 0: start the bulk copies of this CTA's k-th tile
    const int st = (int)(k % STAGES);
    if (k >= STAGES) tma::mbar_wait(empty0 + 8 * st, (unsigned)(((k / STAGES) - 1) & 1));
    const int64_t m0 = ((int64_t)blockIdx.x + k * gridDim.x) * TILE;
    const int64_t cnt = (a.np - m0 < TILE) ? a.np - m0 : TILE;
    const unsigned bytes = (unsigned)(((cnt + 1) & ~(int64_t)1) * 8);  // arrays are padded to an even count
    tma::mbar_expect_tx(full0 + 8 * st, bytes * NARR);
#pragma unroll
    for (int q = 0; q < NARR; q++)
      tma::bulk_g2s(tma::smem_u32(ring + ((size_t)st * NARR + q) * TILE), src[q] + m0, bytes, full0 + 8 * st);
  };
  if (threadIdx.x == 0)
    for (int64_t k = 0; k < STAGES - 1 && k < my_tiles; k++) issue(k);

  const int right_frac = Cfg<CFG>::right_frac(a.right_frac);
  unsigned long long noob = 0;
  for (int64_t k = 0; k < my_tiles; k++) {
    if (threadIdx.x == 0 && k + STAGES - 1 < my_tiles) issue(k + STAGES - 1);
    const int st = (int)(k % STAGES);
    tma::mbar_wait(full0 + 8 * st, (unsigned)((k / STAGES) & 1));
    const double2 *rs = reinterpret_cast<const double2 *>(ring + (size_t)st * NARR * TILE) + threadIdx.x;
    const int T2 = TILE / 2;
    PairIn2 in;
    in.x = rs[0];
    in.v = rs[T2];
    in.w = rs[2 * T2];
    in.p = rs[3 * T2];
    if (IRK2) {
      in.xb = rs[4 * T2];
      in.vb = rs[5 * T2];
      in.wb = rs[6 * T2];
    } else {
      in.xb = in.x;
      in.vb = in.v;
      in.wb = in.w;
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tma::mbar_arrive(empty0 + 8 * st);
    const int64_t m0 = ((int64_t)blockIdx.x + k * gridDim.x) * TILE, i = m0 + (int64_t)threadIdx.x * 2;
    bool redo = true;  // partial tile: scalar path with per-marker validity (reads global memory itself)
    if (m0 + TILE <= a.np)
      redo = push_pair_fast<DIST, IRK2, DEP, true, CFG>(a, sE, dep, i, noob, in.x, in.v, in.w, in.p, in.xb, in.vb, in.wb);
    push_pair_redo<DIST, IRK2, DEP, true, CFG>(a, sE, dep, i, noob, redo, a.np);
  }
  dep_flush<DEP>(dep_base, a.nx, my_partial);
  if (noob) atomicAdd(a.noob, noob);
}

// ------------------------------------------------------------------------------------------------------------
// Deposit-only kernel: the marker loop of interaction_collect_charge (src/pic1dp_interaction.F90:96-114) /
// particle_compute_shape_x (src/pic1dp_particle.F90:306-334): wrap x in place, deposit dep_src.
// DEPOSIT=false: wrap only (compute_shape_x for iptclshape 1-3).
// ------------------------------------------------------------------------------------------------------------
template <int DEP, bool DEPOSIT>
__global__ void __launch_bounds__(1024, 1) k_deposit(const ParticleArgs a) {
  extern __shared__ __align__(16) double smem[];
  double *my_partial = DEPOSIT ? a.partial + (size_t)blockIdx.x * a.nx : nullptr;
  Depositor<DEP> dep;
  dep.g = DEPOSIT ? dep_setup<DEP>(smem, a.nx, my_partial) : nullptr;
  const int64_t tile = (int64_t)blockDim.x * 2;
  if constexpr (DEP == DEP_FIXED)   // wmax set by k_absmax_hi before this launch; this CTA handles <= ceil(tiles / grid) tiles
    fixed_scale(dep, DEPOSIT ? *a.dep_wmax_hi : 0u, ((a.np + tile - 1) / tile + gridDim.x - 1) / gridDim.x * tile, a.nx);
  __syncthreads();
  unsigned long long noob = 0;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < a.np; base += (int64_t)gridDim.x * tile) {
    const int64_t i = base + (int64_t)threadIdx.x * 2;
    const bool v0ok = i < a.np, v1ok = i + 1 < a.np;
    double2 x = {0.0, 0.0}, q = {0.0, 0.0};
    if (v1ok) {
      x = ld2(a.x_cur + i);
      if (DEPOSIT) q = ld2(a.dep_src + i);
    } else if (v0ok) {
      x.x = ld1(a.x_cur + i);
      if (DEPOSIT) q.x = ld1(a.dep_src + i);
    }
    const double2 xw = {wrap_x(x.x, a.lx), wrap_x(x.y, a.lx)};
    // the reference always stores the wrapped value (:102); skip the store when it is bit-identical
    if (v1ok) {
      if (xw.x != x.x || xw.y != x.y) st2(a.x_out + i, xw);
    } else if (v0ok) {
      if (xw.x != x.x) st1(a.x_out + i, xw.x);
    }
    if (DEPOSIT) {
      bool o0 = false, o1 = false;
      const double q0 = dep.prescale(q.x), q1 = dep.prescale(q.y);
      const Shape s0 = shape_of(xw.x, a.lx, a.rlx, a.rnx, a.nx, a.right_frac, o0);
      dep.add(s0.ix, s0.ixr, dmul(s0.sl, q0), dmul(s0.sr, q0), v0ok);
      const Shape s1 = shape_of(xw.y, a.lx, a.rlx, a.rnx, a.nx, a.right_frac, o1);
      dep.add(s1.ix, s1.ixr, dmul(s1.sl, q1), dmul(s1.sr, q1), v1ok);
      noob += (v0ok && o0) + (v1ok && o1);
    }
  }
  if constexpr (DEP == DEP_FIXED && DEPOSIT) {
    fixed_flush(smem, a.nx, my_partial, dep.inv);
    fixed_finish(dep, a.dep_wmax_hi, a.dep_overflow);
  } else if (DEPOSIT) {
    dep_flush<DEP>(smem, a.nx, my_partial);
  }
  if (DEPOSIT && noob) atomicAdd(a.noob, noob);
}

}  // namespace pic1dp
