/*
 * oracle/optimize_oracle.c -- CPU restatement of the marker-optimisation routines of pic1dp_particle
 * (/root/reference/src/pic1dp_particle.F90:356-746): particle_compute_dist_pertb_abs_v, particle_merge,
 * particle_remove, particle_split.
 *
 * TEST INFRASTRUCTURE ONLY (see pic1dp_oracle.h).  PARITY STATUS: "parity unpinned" -- the reference stores no
 * vectors for these routines and cannot be built here.  The restatement keeps the reference's 1-based marker
 * indices and visits markers in the reference's order (the routines are order dependent: a removed or merged
 * marker is overwritten by the last one and the same slot is visited again), expressions left to right, no FMA
 * contraction.  Random numbers come from the KAT-pinned multirand restatement.
 */
#include <math.h>
#include <stdlib.h>

#include "pic1dp_oracle.h"

/* marker arrays as the Fortran pointer arrays px(1:), pv(1:), ... */
#define PX(i) x[(i) - 1]
#define PV(i) v[(i) - 1]
#define PP(i) pp[(i) - 1]
#define PW(i) w[(i) - 1]

/* src/pic1dp_particle.F90:378-390: one rank, one species; dist[0 .. nv-1] is zeroed first (:371) */
void orc_dist_pertb_abs_v_rank(int64_t np, const double *v, const double *w, int nv, double v_max, double *dist) {
  for (int i = 0; i < nv; i++) dist[i] = 0.0;
  for (int64_t ip = 1; ip <= np; ip++) {
    if (fabs(PV(ip)) >= v_max) continue; /* :380 */
    double sv = (PV(ip) + v_max) / (v_max * 2.0) * (double)(nv - 1); /* :382-383 */
    const int iv = (int)floor(sv);
    sv = 1.0 - (sv - (double)iv); /* :385 */
    dist[iv] = dist[iv] + sv * fabs(PW(ip));                     /* :387 */
    /* v just below v_max can round to sv == nv-1: the reference then adds 0 * |w| one past the end of the grid */
    if (iv + 1 < nv) dist[iv + 1] = dist[iv + 1] + (1.0 - sv) * fabs(PW(ip)); /* :388-389 */
  }
}

/* MPI_Allreduce(..., MPI_SUM) of the per-rank grids (:392-395), emulated ranks added in rank order */
void orc_dist_pertb_abs_v(int nranks, const int64_t *np, double **v, double **w, int nv, double v_max, double *dist) {
  double *loc = (double *)malloc(sizeof(double) * (size_t)nv);
  for (int r = 0; r < nranks; r++) {
    orc_dist_pertb_abs_v_rank(np[r], v[r], w[r], nv, v_max, loc);
    for (int i = 0; i < nv; i++) dist[i] = (r == 0) ? loc[i] : dist[i] + loc[i];
  }
  free(loc);
}

static double maxval(const double *a, int n) {
  double m = a[0];
  for (int i = 1; i < n; i++)
    if (a[i] > m) m = a[i];
  return m;
}

/* the importance lookup shared by merge / remove / split (:452-466, :567-581, :679-693) */
static double lookup_df(const double *dist, int nv, double v_max, double vel, int *iv_out) {
  double sv = (vel + v_max) / (v_max * 2.0) * (double)(nv - 1);
  const double fl = floor(sv);
  int iv;
  double df;
  if (fl < 0.0) {
    iv = 0;
    df = dist[iv];
  } else if (fl >= (double)(nv - 1)) {
    iv = nv - 1;
    df = dist[iv];
  } else {
    iv = (int)fl;
    sv = 1.0 - (sv - (double)iv);
    df = dist[iv] * sv + dist[iv + 1] * (1.0 - sv);
  }
  *iv_out = iv;
  return df;
}

/* src/pic1dp_particle.F90:411-522, one species on one rank; returns the new particle_np */
int64_t orc_particle_merge(const orc_params *p, int64_t np, double *x, double *v, double *pp, double *w,
                           const double *dist, int nv, double v_max, double thsh) {
  const int nx = p->nx;
  const size_t nbin = (size_t)nx * (size_t)nv * 2;
  /* ipbin(ix, iv, iw, 1) and ipbin_top(ix, iv, iw): :427-431 */
  int64_t *ipbin = (int64_t *)malloc(sizeof(int64_t) * nbin);
  int *ipbin_top = (int *)malloc(sizeof(int) * nbin);
  for (size_t b = 0; b < nbin; b++) ipbin_top[b] = 1;
  const double df_thsh = maxval(dist, nv) * thsh; /* :432-433 */
  int64_t ip = 0;
  for (;;) {
    ip = ip + 1;
    if (ip > np) break;
    int iv;
    const double df = lookup_df(dist, nv, v_max, PV(ip), &iv);
    if (df >= df_thsh) continue; /* :468 */
    PX(ip) = fmod(PX(ip), p->lx); /* :471 */
    if (PX(ip) < 0.0) PX(ip) = PX(ip) + p->lx;
    const double sx = PX(ip) / p->lx * (double)nx; /* :475 */
    int ix = (int)floor(sx);
    if (ix >= nx) ix = 0; /* x == lx exactly: out of bounds in the reference, defined as cell 0 like the deposit */
    const int iw = (PW(ip) > 0.0) ? 2 : 1;
    const size_t b = ((size_t)(iw - 1) * (size_t)nv + (size_t)iv) * (size_t)nx + (size_t)ix;
    if (ipbin_top[b] < 2) { /* :482-484 */
      ipbin[b] = ip;
      ipbin_top[b] = ipbin_top[b] + 1;
    } else { /* :485-507 */
      const int64_t ip1 = ipbin[b];
      PX(ip1) = (PW(ip1) * PX(ip1) + PW(ip) * PX(ip)) / (PW(ip1) + PW(ip));
      PV(ip1) = (PW(ip1) * PV(ip1) + PW(ip) * PV(ip)) / (PW(ip1) + PW(ip));
      PP(ip1) = PP(ip1) + PP(ip);
      PW(ip1) = PW(ip1) + PW(ip);
      if (ip < np) {
        PX(ip) = PX(np);
        PV(ip) = PV(np);
        PP(ip) = PP(np);
        PW(ip) = PW(np);
        ip = ip - 1;
      }
      np = np - 1;
      ipbin_top[b] = 1;
    }
  }
  free(ipbin);
  free(ipbin_top);
  return np;
}

/* src/pic1dp_particle.F90:530-627, one species on one rank; returns the new particle_np */
int64_t orc_particle_remove(int64_t np, double *x, double *v, double *pp, double *w, const double *dist, int nv,
                            double v_max, double thsh, int typeremove, double remove_frac, orc_multirand *rng) {
  const double df_thsh = maxval(dist, nv) * thsh; /* :547-548 */
  int64_t ip = 0;
  for (;;) {
    ip = ip + 1;
    if (ip > np) break;
    int iv;
    double df = lookup_df(dist, nv, v_max, PV(ip), &iv);
    if (typeremove == 1) {
      if (df >= df_thsh) continue; /* :582-585 */
    }
    df = df / maxval(dist, nv);                    /* :587 */
    const double dice = orc_multirand_real64(rng); /* :591 */
    if ((typeremove == 1 && dice < remove_frac) || (typeremove == 2 && dice > df)) { /* :594-604 */
      if (ip < np) {
        PX(ip) = PX(np);
        PV(ip) = PV(np);
        PP(ip) = PP(np);
        PW(ip) = PW(np);
        ip = ip - 1;
      }
      np = np - 1;
    } else { /* :605-614 */
      if (typeremove == 1) {
        PP(ip) = PP(ip) / (1.0 - remove_frac);
        PW(ip) = PW(ip) / (1.0 - remove_frac);
      } else {
        PP(ip) = PP(ip) / df;
        PW(ip) = PW(ip) / df;
      }
    }
  }
  return np;
}

/* src/pic1dp_particle.F90:635-746, one species on one rank; capacity = particle_ip_high - particle_ip_low;
 * returns the new particle_np */
int64_t orc_particle_split(const orc_params *p, int64_t np, int64_t capacity, double *x, double *v, double *pp,
                           double *w, const double *dist, int nv, double v_max, double thsh, int ngroup,
                           double dv_sig_frac, orc_multirand *rng) {
  if (capacity - np < 2 * ngroup - 1) return np; /* :656-657 */
  int64_t np_inc = 0;
  const double df_thsh = maxval(dist, nv) * thsh; /* :660-661 */
  double *grand = (double *)malloc(sizeof(double) * (size_t)ngroup);
  for (int64_t ip = 1; ip <= np; ip++) {
    if (capacity - (np + np_inc) < 2 * ngroup - 1) break; /* :674-675 */
    int iv;
    const double df = lookup_df(dist, nv, v_max, PV(ip), &iv);
    if (df <= df_thsh) continue; /* :695 */
    orc_multirand_gaussian_array(rng, grand, ngroup); /* :697 */
    for (int g = 0; g < ngroup; g++) grand[g] = grand[g] * 2.0 * v_max / (double)nv * dv_sig_frac; /* :699-700 */
    for (int igroup = 1; igroup <= ngroup; igroup++) {
      int64_t ip1 = np + np_inc + igroup * 2 - 1; /* :707-711 */
      PX(ip1) = PX(ip);
      PV(ip1) = PV(ip) + grand[igroup - 1];
      PP(ip1) = PP(ip) / ((double)ngroup * 2.0);
      if (p->deltaf == 1) PW(ip1) = PW(ip) / ((double)ngroup * 2.0);
      if (igroup == ngroup) /* :716-724 */
        ip1 = ip;
      else
        ip1 = np + np_inc + igroup * 2;
      PX(ip1) = PX(ip);
      PV(ip1) = PV(ip) - grand[igroup - 1];
      PP(ip1) = PP(ip) / ((double)ngroup * 2.0);
      if (p->deltaf == 1) PW(ip1) = PW(ip) / ((double)ngroup * 2.0);
    }
    np_inc = np_inc + (2 * ngroup - 1); /* :730 */
  }
  free(grand);
  return np + np_inc; /* :743 */
}
