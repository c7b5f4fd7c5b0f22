// optimize_host.hpp -- host halves of the marker optimisation of pic1dp_particle
// (/root/reference/src/pic1dp_particle.F90:411-746: particle_merge, particle_remove, particle_split).
//
// These three routines are sequential by definition: a merged or removed marker is overwritten by the LAST marker
// and the same slot is examined again, bins pair markers in visiting order, and remove / split consume one RNG stream
// (multirand, itself a sequential generator) in visiting order.  A result that is identical to the reference's
// therefore needs the reference's visiting order; they run on the host over a staged copy of the marker arrays
// (a few events per run -- input_nmerge / input_nremove / input_nsplit -- not per step), while the O(N) reduction
// they depend on, particle_compute_dist_pertb_abs_v, runs on the device (optimize_kernels.cuh).
//
// Arithmetic follows the reference expression by expression (compile with -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

namespace pic1dp {
namespace hostopt {

struct Markers {
  double *x, *v, *p, *w;
  void move(int64_t dst, int64_t src) const {
    x[dst] = x[src];
    v[dst] = v[src];
    p[dst] = p[src];
    w[dst] = w[src];
  }
};

// importance of a marker = int |delta f| dx interpolated at its velocity (:452-466, :567-581, :679-693);
// outside the v grid the end values are used
struct Importance {
  const double *dist;
  int nv;
  double v_max, peak;  // peak = maxval(particle_dist_pertb_abs_v(ispecies, :))
  Importance(const double *d, int nv_, double v_max_) : dist(d), nv(nv_), v_max(v_max_) {
    peak = d[0];
    for (int i = 1; i < nv_; i++) peak = d[i] > peak ? d[i] : peak;
  }
  double at(double vel, int &iv) const {
    const double sv = (vel + v_max) / (v_max * 2.0) * (double)(nv - 1);
    const double cell = floor(sv);
    if (cell < 0.0) return dist[iv = 0];
    if (cell >= (double)(nv - 1)) return dist[iv = nv - 1];
    iv = (int)cell;
    const double s = 1.0 - (sv - (double)iv);
    return dist[iv] * s + dist[iv + 1] * (1.0 - s);
  }
};

// particle_merge (:411-522): unimportant markers (importance < thsh * peak) that fall into the same
// (x cell, v cell, sign of w) bin are merged pairwise in visiting order.  Returns the new marker count.
inline int64_t merge(const Markers &m, int64_t n, const double *dist, int nv, double v_max, double thsh, int nx,
                     double lx) {
  const Importance imp(dist, nv, v_max);
  const double limit = imp.peak * thsh;
  std::vector<int64_t> waiting((size_t)nx * (size_t)nv * 2, -1);  // marker waiting in each bin, -1: none
  int64_t i = 0;
  while (i < n) {
    int iv;
    if (imp.at(m.v[i], iv) >= limit) {  // important marker: keep (:468)
      i++;
      continue;
    }
    double xi = fmod(m.x[i], lx);  // periodic boundary (:471-473)
    if (xi < 0.0) xi = xi + lx;
    m.x[i] = xi;
    int ix = (int)floor(xi / lx * (double)nx);
    if (ix >= nx) ix = 0;  // x == lx exactly: the reference indexes out of bounds; cell 0 like the deposit does
    const size_t bin = ((size_t)(m.w[i] > 0.0 ? 1 : 0) * (size_t)nv + (size_t)iv) * (size_t)nx + (size_t)ix;
    const int64_t j = waiting[bin];
    if (j < 0) {  // first of a pair
      waiting[bin] = i++;
      continue;
    }
    const double wsum = m.w[j] + m.w[i];  // (:488-493) x and v are w-weighted means, p and w add
    m.x[j] = (m.w[j] * m.x[j] + m.w[i] * m.x[i]) / wsum;
    m.v[j] = (m.w[j] * m.v[j] + m.w[i] * m.v[i]) / wsum;
    m.p[j] = m.p[j] + m.p[i];
    m.w[j] = wsum;
    waiting[bin] = -1;
    n--;
    if (i < n) m.move(i, n);  // the last marker takes the freed slot and is examined next (:495-503)
  }
  return n;
}

typedef double (*real64_fn)(void *);
typedef void (*gaussian_array_fn)(void *, double *, int32_t);

// particle_remove (:530-627).  typeremove 1: unimportant markers are removed with probability remove_frac and the
// survivors scaled by 1/(1-remove_frac); typeremove 2: a marker survives with probability importance/peak and is
// scaled by the inverse.  One dice per examined marker, in visiting order.
inline int64_t remove(const Markers &m, int64_t n, const double *dist, int nv, double v_max, double thsh, int typeremove,
                      double remove_frac, real64_fn dice, void *ctx) {
  const Importance imp(dist, nv, v_max);
  const double limit = imp.peak * thsh;
  int64_t i = 0;
  while (i < n) {
    int iv;
    double df = imp.at(m.v[i], iv);
    if (typeremove == 1 && df >= limit) {  // (:582-585)
      i++;
      continue;
    }
    df = df / imp.peak;  // (:587)
    const double d = dice(ctx);
    const bool drop = (typeremove == 1 && d < remove_frac) || (typeremove == 2 && d > df);
    if (drop) {
      n--;
      if (i < n) m.move(i, n);  // examined again at the same index (:597-604)
      continue;
    }
    const double scale = (typeremove == 1) ? 1.0 - remove_frac : df;  // (:607-613)
    m.p[i] = m.p[i] / scale;
    m.w[i] = m.w[i] / scale;
    i++;
  }
  return n;
}

// particle_split (:635-746): every important marker (importance > thsh * peak) becomes 2*ngroup markers at
// v +- Gaussian offsets with 1/(2*ngroup) of its weights; new markers are appended.  capacity = length of the arrays.
inline int64_t split(const Markers &m, int64_t n, int64_t capacity, const double *dist, int nv, double v_max, double thsh,
                     int ngroup, double dv_sig_frac, int deltaf, gaussian_array_fn gauss, void *ctx) {
  const int64_t need = 2 * (int64_t)ngroup - 1;
  if (capacity - n < need) return n;  // (:656-657)
  const Importance imp(dist, nv, v_max);
  const double limit = imp.peak * thsh, share = (double)ngroup * 2.0;
  std::vector<double> dv((size_t)ngroup);
  int64_t tail = n;  // next free slot
  for (int64_t i = 0; i < n && capacity - tail >= need; i++) {  // (:672-675)
    int iv;
    if (imp.at(m.v[i], iv) <= limit) continue;  // (:695)
    gauss(ctx, dv.data(), ngroup);
    for (int g = 0; g < ngroup; g++) dv[g] = dv[g] * 2.0 * v_max / (double)nv * dv_sig_frac;  // (:699-700)
    for (int g = 0; g < ngroup; g++) {
      const int64_t up = tail++;                           // v + dv  (:707-711)
      const int64_t down = (g == ngroup - 1) ? i : tail++;  // v - dv; the last one replaces the parent (:716-724)
      m.x[up] = m.x[i];
      m.v[up] = m.v[i] + dv[g];
      m.p[up] = m.p[i] / share;
      if (deltaf == 1) m.w[up] = m.w[i] / share;
      m.x[down] = m.x[i];
      m.v[down] = m.v[i] - dv[g];
      m.p[down] = m.p[i] / share;
      if (deltaf == 1) m.w[down] = m.w[i] / share;
    }
  }
  return tail;
}

}  // namespace hostopt
}  // namespace pic1dp
