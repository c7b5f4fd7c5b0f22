/*
 * oracle/pic1dp_oracle.c -- CPU restatement of the PIC1D-PETSc per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see the header of pic1dp_oracle.h.  "Parity unpinned": the reference holds no
 * golden vectors for this path and cannot be built here; this file restates it statement by statement.
 * Build: gcc -O3 -ffp-contract=off -pthread (mirrors FFLAGS := -O3 on x86-64 without FMA contraction,
 * /root/reference/Makefile:26).  Every expression is evaluated left to right exactly as written in the
 * Fortran source, honouring its parentheses.
 */
#include "pic1dp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#define ORC_PI 3.14159265358979323846264338327950288419716939937510582 /* PETSC_PI */

void orc_params_default(orc_params *p) {
  memset(p, 0, sizeof(*p));
  p->nx = 192;                                        /* src/pic1dp_input.F90:128 */
  p->nmode = 1;                                       /* :75 */
  p->modes[0] = 1;                                    /* :80 */
  p->lx = 2.0 * 3.1415926535897932384626 / 0.36;      /* :47 */
  p->dt = 0.05;                                       /* :109 */
  p->nspecies = 1;                                    /* :57 */
  p->charge[0] = -1.0;                                /* :67-72 */
  p->mass[0] = 1.0;
  p->temperature[0] = 1.0;
  p->temperature2[0] = 1.0;
  p->density[0] = 0.9;
  p->v0[0] = 5.0;
  p->iptcldist = 3;                                   /* :54 */
  p->deltaf = 1;                                      /* :106 */
  p->linear = 0;                                      /* :43 */
  p->iptclshape = 4;                                  /* :138 */
  p->v_max = 8.0;                                     /* :125 */
  p->imarker = 2;                                     /* :122 */
  p->init_nmode = 1;                                  /* :87-98 */
  p->init_mode[0] = 1;
  p->init_mode_cos[0] = 0.00;
  p->init_mode_sin[0] = 1e-5;
}

/* src/pic1dp_field.F90:163-167 and :184-202 */
void orc_field_init(const orc_params *p, double *F_re, double *F_im, double *grad_inv) {
  const int nx = p->nx, M = p->nmode;
  for (int m = 0; m < M; m++) grad_inv[m] = 1.0 / (2.0 * ORC_PI / p->lx * (double)p->modes[m]);
  for (int j = 0; j < nx; j++) {
    for (int m = 0; m < M; m++) {
      /* 2.0_kpr * PETSC_PI / input_nx * real(input_modes(imode), kpr) * ix */
      double arg = 2.0 * ORC_PI / (double)nx * (double)p->modes[m] * (double)j;
      F_re[(size_t)j * M + m] = cos(arg);
      F_im[(size_t)j * M + m] = -sin(arg);
    }
  }
}

/* src/pic1dp_field.F90:231-256; PETSc SeqAIJ kernels: MatMultTranspose accumulates y[col] += a*x[row] in row
 * order, MatMult sums a row left to right starting from 0, MatMultAdd starts from the addend. */
void orc_field_solve(const orc_params *p, const double *F_re, const double *F_im, const double *grad_inv,
                     const double *rho, double *E, double *mode_re, double *mode_im) {
  const int nx = p->nx, M = p->nmode;
  for (int m = 0; m < M; m++) {
    mode_im[m] = 0.0;
    mode_re[m] = 0.0;
  }
  for (int j = 0; j < nx; j++)
    for (int m = 0; m < M; m++) mode_im[m] = mode_im[m] + F_re[(size_t)j * M + m] * rho[j]; /* :231 */
  {
    const double a = -1.0 / (double)nx; /* :234 */
    for (int m = 0; m < M; m++) mode_im[m] = mode_im[m] * a;
  }
  for (int j = 0; j < nx; j++)
    for (int m = 0; m < M; m++) mode_re[m] = mode_re[m] + F_im[(size_t)j * M + m] * rho[j]; /* :236 */
  {
    const double a = 1.0 / (double)nx; /* :239 */
    for (int m = 0; m < M; m++) mode_re[m] = mode_re[m] * a;
  }
  for (int m = 0; m < M; m++) mode_re[m] = mode_re[m] * grad_inv[m]; /* :243 */
  for (int m = 0; m < M; m++) mode_im[m] = mode_im[m] * grad_inv[m]; /* :246 */
  for (int j = 0; j < nx; j++) {
    double sum = 0.0;
    for (int m = 0; m < M; m++) sum = sum + F_re[(size_t)j * M + m] * mode_re[m]; /* :251 */
    for (int m = 0; m < M; m++) sum = sum + F_im[(size_t)j * M + m] * mode_im[m]; /* :253 */
    E[j] = sum * 2.0;                                                             /* :256 */
  }
}

/* wrap of one coordinate: src/pic1dp_interaction.F90:102-104 == src/pic1dp_particle.F90:308-310 */
static inline double wrap_x(double x, double lx) {
  x = fmod(x, lx);
  if (x < 0.0) x = x + lx;
  return x;
}

void orc_shape(const orc_params *p, int64_t np, double *x, int32_t *ix, double *s_left, double *s_right,
               int right_frac) {
  const double lx = p->lx, rnx = (double)p->nx;
  for (int64_t i = 0; i < np; i++) {
    x[i] = wrap_x(x[i], lx);
    double sx = x[i] / lx * rnx; /* :106 / particle.F90:312 */
    int32_t k = (int32_t)floor(sx);
    double frac = sx - (double)k;
    double sl = 1.0 - frac;
    ix[i] = k;
    s_left[i] = sl;
    s_right[i] = right_frac ? frac /* particle.F90:323 */ : (1.0 - sl) /* interaction.F90:113 */;
  }
}

/* src/pic1dp_interaction.F90:83-114, array path (iptclshape 3/4).  right_frac selects the matrix-mode weight. */
static int64_t deposit_species(const orc_params *p, int64_t np, double *x, const double *w, double *c1,
                               int right_frac) {
  const int nx = p->nx;
  const double lx = p->lx, rnx = (double)nx;
  int64_t noob = 0;
  for (int j = 0; j < nx; j++) c1[j] = 0.0; /* :83 */
  for (int64_t ip = 0; ip < np; ip++) {
    double px = fmod(x[ip], lx);   /* :102 */
    if (px < 0.0) px = px + lx;    /* :104 */
    x[ip] = px;
    double sx = px / lx * rnx;     /* :106 */
    int32_t ix = (int32_t)floor(sx); /* :107 */
    double frac = sx - (double)ix;
    sx = 1.0 - frac;               /* :108 */
    if (ix >= nx) {                /* reference OOB (x wrapped to exactly lx); defined here as cell 0, s = 1 */
      noob++;
      ix = 0;
      sx = 1.0;
      frac = 0.0;
    }
    c1[ix] = c1[ix] + sx * w[ip];  /* :110 */
    ix = ix + 1;                   /* :111 */
    if (ix > nx - 1) ix = 0;       /* :112 */
    if (right_frac)
      c1[ix] = c1[ix] + frac * w[ip];        /* S^T w with values(1) = sx, particle.F90:323 */
    else
      c1[ix] = c1[ix] + (1.0 - sx) * w[ip];  /* :113 */
  }
  return noob;
}

int64_t orc_deposit_species(const orc_params *p, int64_t np, double *x, const double *w, double *charge1) {
  return deposit_species(p, np, x, w, charge1, p->iptclshape <= 2);
}

/* Combination of per-rank, per-species local grids c1_all[s][r][nx] into rho.
 * Array path (iptclshape 3,4): src/pic1dp_interaction.F90:81, :126-127, :132-133 (rank-ordered sum stands in
 * for MPI_Allreduce), :140-148.  Matrix path (iptclshape 1,2): :47-77 (field_tmp = S^T w as a rank-ordered
 * sum of per-rank products, full-f offset :67, VecAXPY :71, VecScale :77). */
static void combine_charge(const orc_params *p, int nranks, const double *c1_all, double *rho) {
  const int nx = p->nx, S = p->nspecies;
  double *acc = (double *)malloc(sizeof(double) * nx);
  if (p->iptclshape >= 3) {
    double *tot = (double *)malloc(sizeof(double) * nx);
    for (int r = 0; r < nranks; r++) {
      for (int j = 0; j < nx; j++) acc[j] = 0.0; /* :81 */
      for (int s = 0; s < S; s++) {
        const double *c1 = c1_all + ((size_t)s * nranks + r) * nx;
        for (int j = 0; j < nx; j++) acc[j] = acc[j] + c1[j] * p->charge[s]; /* :126-127 */
      }
      if (r == 0)
        memcpy(tot, acc, sizeof(double) * nx);
      else
        for (int j = 0; j < nx; j++) tot[j] = tot[j] + acc[j];
    }
    for (int j = 0; j < nx; j++) rho[j] = tot[j] * (double)nx / p->lx; /* :140-141 */
    if (p->deltaf == 0)                                                /* :142-148 */
      for (int s = 0; s < S; s++)
        for (int j = 0; j < nx; j++) rho[j] = rho[j] - p->charge[s] * p->density[s];
    free(tot);
  } else {
    for (int j = 0; j < nx; j++) rho[j] = 0.0; /* :47 */
    for (int s = 0; s < S; s++) {
      for (int r = 0; r < nranks; r++) {
        const double *c1 = c1_all + ((size_t)s * nranks + r) * nx;
        if (r == 0)
          memcpy(acc, c1, sizeof(double) * nx);
        else
          for (int j = 0; j < nx; j++) acc[j] = acc[j] + c1[j];
      }
      if (p->deltaf == 0) /* :67 */
        for (int j = 0; j < nx; j++) acc[j] = acc[j] - p->density[s] * p->lx / (double)nx;
      for (int j = 0; j < nx; j++) rho[j] = rho[j] + p->charge[s] * acc[j]; /* VecAXPY :71 */
    }
    {
      const double a = (double)nx / p->lx; /* :77 */
      for (int j = 0; j < nx; j++) rho[j] = rho[j] * a;
    }
  }
  free(acc);
}

int64_t orc_collect_charge(const orc_params *p, int nranks, const int64_t *np, double **x, double **wsrc,
                           double *rho) {
  const int nx = p->nx, S = p->nspecies;
  int64_t noob = 0;
  double *c1_all = (double *)malloc(sizeof(double) * (size_t)S * nranks * nx);
  for (int s = 0; s < S; s++)
    for (int r = 0; r < nranks; r++)
      noob += deposit_species(p, np[s * nranks + r], x[s * nranks + r], wsrc[s * nranks + r],
                              c1_all + ((size_t)s * nranks + r) * nx, p->iptclshape <= 2);
  combine_charge(p, nranks, c1_all, rho);
  free(c1_all);
  return noob;
}

/* src/pic1dp_interaction.F90:275-326 */
double orc_dlnf0(const orc_params *p, int isp, double v) {
  const double T = p->temperature[isp], T2 = p->temperature2[isp], m = p->mass[isp];
  const double n = p->density[isp], v0 = p->v0[isp];
  double tmp2;
  if (p->iptcldist == 1) { /* two-stream1 :276 */
    tmp2 = v - 2.0 / v;
  } else if (p->iptcldist == 2) { /* two-stream2 :278-292 */
    double ep = exp(-((v + v0) * (v + v0)) / (2.0 * T / m));
    double em = exp(-((v - v0) * (v - v0)) / (2.0 * T / m));
    tmp2 = ((v + v0) * ep + (v - v0) * em) / (ep + em) * m / T;
  } else if (p->iptcldist == 3) { /* bump-on-tail :294-321 */
    double e1 = exp(-(v * v) / (2.0 * T / m));
    double e2 = exp(-((v - v0) * (v - v0)) / (2.0 * T2 / m));
    double num = n * v / (T / m) * e1 / sqrt(T / m) + (1.0 - n) * (v - v0) / (T2 / m) * e2 / sqrt(T2 / m);
    double den = n * e1 / sqrt(T / m) + (1.0 - n) * e2 / sqrt(T2 / m);
    tmp2 = num / den;
  } else { /* (shifted) Maxwellian :323-325 */
    tmp2 = (v - v0) / (T / m);
  }
  return tmp2;
}

/* src/pic1dp_interaction.F90:178-193 (dt, backup) and :238-339 (loop body) */
void orc_push_species(const orc_params *p, int isp, int irk, int64_t np, double *x, double *v,
                      const double *pp, double *w, double *xb, double *vb, double *wb, const double *E) {
  const int nx = p->nx;
  const double lx = p->lx, rnx = (double)nx;
  const double Z = p->charge[isp], m = p->mass[isp];
  const int right_frac = p->iptclshape <= 2;
  double dt;
  if (irk == 1) {
    dt = 0.5 * p->dt; /* :179 */
    memcpy(xb, x, sizeof(double) * (size_t)np); /* :181 */
    memcpy(vb, v, sizeof(double) * (size_t)np); /* :183 */
    if (p->deltaf == 1) memcpy(wb, w, sizeof(double) * (size_t)np); /* :186 */
  } else {
    dt = p->dt; /* :192 */
  }
  for (int64_t ip = 0; ip < np; ip++) {
    double sx = x[ip] / lx * rnx;     /* :250 */
    int32_t ix = (int32_t)floor(sx);  /* :251 */
    double frac = sx - (double)ix;
    sx = 1.0 - frac;                  /* :252 */
    if (ix >= nx) {                   /* x == lx exactly: same definition as the deposit */
      ix = 0;
      sx = 1.0;
      frac = 0.0;
    }
    double electric = E[ix] * sx;     /* :254 */
    ix = ix + 1;
    if (ix > nx - 1) ix = 0;          /* :256 */
    if (right_frac)
      electric = electric + E[ix] * frac;        /* MatMult(S,E) row sum, interaction.F90:215 */
    else
      electric = electric + E[ix] * (1.0 - sx);  /* :257 */

    const double vcur = v[ip];
    x[ip] = xb[ip] + dt * vcur;       /* :261 */
    if (p->deltaf == 1) {
      double tmp1;
      if (p->linear == 1)
        tmp1 = pp[ip] * electric;     /* :269 */
      else
        tmp1 = (pp[ip] - w[ip]) * electric; /* :271 */
      double tmp2 = orc_dlnf0(p, isp, vcur);
      w[ip] = wb[ip] + dt * tmp1 * tmp2 * Z / m; /* :329-330 */
    }
    if (p->linear == 0) v[ip] = vb[ip] + dt * electric * Z / m; /* :336-337 */
  }
}

/* src/pic1dp_output.F90:120-123: VecNorm(NORM_2)^2 * lx / nx */
double orc_field_energy(const orc_params *p, const double *E) {
  double s = 0.0;
  for (int j = 0; j < p->nx; j++) s += E[j] * E[j];
  double nrm = sqrt(s);
  return nrm * nrm * p->lx / (double)p->nx;
}

/* src/pic1dp_output.F90:117-172 */
void orc_output_field(const orc_params *p, int nranks, const int64_t *np, double **v, double **pp, double **w,
                      const double *E, double *out) {
  out[0] = orc_field_energy(p, E);
  for (int s = 0; s < p->nspecies; s++) {
    double vv = 0.0, vvp = 0.0, vvw = 0.0;
    for (int r = 0; r < nranks; r++) {
      const int k = s * nranks + r;
      double a = 0.0, b = 0.0, c = 0.0;
      for (int64_t i = 0; i < np[k]; i++) {
        const double t1 = v[k][i] * v[k][i]; /* :128 */
        a += t1;                             /* :133 */
        b += t1 * pp[k][i];                  /* :138-141 */
        if (p->deltaf == 1) c += t1 * w[k][i]; /* :147-150 */
      }
      vv = (r == 0) ? a : vv + a;
      vvp = (r == 0) ? b : vvp + b;
      vvw = (r == 0) ? c : vvw + c;
    }
    out[1 + 3 * s] = vv;
    out[2 + 3 * s] = vvp;
    double energy;
    if (p->deltaf == 1) {
      energy = vvw;
      if (p->linear == 1) out[2 + 3 * s] = out[2 + 3 * s] + energy; /* :154 */
    } else {
      energy = vvp;
      if (p->iptcldist == 1)
        energy = energy - 3.0 * p->density[s] * p->lx; /* :160 */
      else if (p->iptcldist == 0)
        energy = energy - p->temperature[s] / p->mass[s] * p->density[s] * p->lx; /* :166-168 */
    }
    out[3 + 3 * s] = energy; /* :171 */
  }
}

/* src/pic1dp_output.F90:196-477 */
void orc_output_ptcldist(const orc_params *p, int isp, int nranks, const int64_t *np, double **x, double **v,
                         double **pp, double **w, int nx_opd, int nv_opd, double v_max, double *markr_xv,
                         double *total_xv, double *pertb_xv, double *markr_v, double *total_v, double *pertb_v) {
  const int nc = nx_opd * nv_opd;
  const double delv_inv = (double)(nv_opd - 1) / (2.0 * v_max); /* :207-208 */
  const double delx_inv = (double)nx_opd / p->lx;               /* :209 */
  double *mxv = (double *)malloc(sizeof(double) * nc), *txv = (double *)malloc(sizeof(double) * nc);
  double *pxv = (double *)malloc(sizeof(double) * nc);
  double *mv = (double *)malloc(sizeof(double) * nv_opd), *tv = (double *)malloc(sizeof(double) * nv_opd);
  double *pv_ = (double *)malloc(sizeof(double) * nv_opd);
  for (int r = 0; r < nranks; r++) {
    const int k = isp * nranks + r;
    for (int c = 0; c < nc; c++) mxv[c] = txv[c] = pxv[c] = 0.0;
    for (int c = 0; c < nv_opd; c++) mv[c] = tv[c] = pv_[c] = 0.0;
    for (int64_t ip = 0; ip < np[k]; ip++) {
      const double vi = v[k][ip];
      if (fabs(vi) >= v_max) continue; /* :241 */
      double sx = x[k][ip] / p->lx * (double)nx_opd; /* :243 */
      int ix = (int)floor(sx);
      sx = 1.0 - (sx - (double)ix);
      double sv = (vi + v_max) / (v_max * 2.0) * (double)(nv_opd - 1); /* :247-248 */
      const int iv = (int)floor(sv);
      sv = 1.0 - (sv - (double)iv);
      if (ix >= nx_opd) { /* x == lx exactly: out of bounds in the reference; defined as cell 0, weight 1 */
        ix = 0;
        sx = 1.0;
      }
      const double P = pp[k][ip], W = (p->deltaf == 1) ? w[k][ip] : 0.0;
      mxv[iv * nx_opd + ix] += sx * sv;
      txv[iv * nx_opd + ix] += sx * sv * P;
      if (p->deltaf == 1) pxv[iv * nx_opd + ix] += sx * sv * W;
      mxv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv);
      txv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv) * P;
      if (p->deltaf == 1) pxv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv) * W;
      ix = ix + 1;
      if (ix > nx_opd - 1) ix = 0; /* :272 */
      sx = 1.0 - sx;               /* :273 */
      mxv[iv * nx_opd + ix] += sx * sv;
      txv[iv * nx_opd + ix] += sx * sv * P;
      if (p->deltaf == 1) pxv[iv * nx_opd + ix] += sx * sv * W;
      mxv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv);
      txv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv) * P;
      if (p->deltaf == 1) pxv[(iv + 1) * nx_opd + ix] += sx * (1.0 - sv) * W;
      mv[iv] += sv; /* :297-312 */
      tv[iv] += sv * P;
      if (p->deltaf == 1) pv_[iv] += sv * W;
      mv[iv + 1] += (1.0 - sv);
      tv[iv + 1] += (1.0 - sv) * P;
      if (p->deltaf == 1) pv_[iv + 1] += (1.0 - sv) * W;
    }
    if (p->linear == 1) { /* :326-329 */
      for (int c = 0; c < nc; c++) txv[c] = txv[c] + pxv[c];
      for (int c = 0; c < nv_opd; c++) tv[c] = tv[c] + pv_[c];
    }
    for (int c = 0; c < nc; c++) { /* MPI_Reduce, rank order */
      markr_xv[c] = (r == 0) ? mxv[c] : markr_xv[c] + mxv[c];
      total_xv[c] = (r == 0) ? txv[c] : total_xv[c] + txv[c];
      pertb_xv[c] = (r == 0) ? pxv[c] : pertb_xv[c] + pxv[c];
    }
    for (int c = 0; c < nv_opd; c++) {
      markr_v[c] = (r == 0) ? mv[c] : markr_v[c] + mv[c];
      total_v[c] = (r == 0) ? tv[c] : total_v[c] + tv[c];
      pertb_v[c] = (r == 0) ? pv_[c] : pertb_v[c] + pv_[c];
    }
  }
  for (int c = 0; c < nc; c++) { /* :362-363 */
    markr_xv[c] = markr_xv[c] * delx_inv * delv_inv;
    total_xv[c] = total_xv[c] * delx_inv * delv_inv;
  }
  for (int c = 0; c < nv_opd; c++) {
    markr_v[c] = markr_v[c] * delv_inv;
    total_v[c] = total_v[c] * delv_inv;
  }
  if (p->deltaf == 1) {
    for (int c = 0; c < nc; c++) pertb_xv[c] = pertb_xv[c] * delx_inv * delv_inv;
    for (int c = 0; c < nv_opd; c++) pertb_v[c] = pertb_v[c] * delv_inv;
  } else { /* :371-455 */
    const double n = p->density[isp], v0 = p->v0[isp], T = p->temperature[isp], T2 = p->temperature2[isp];
    const double m = p->mass[isp];
    for (int iv = 0; iv < nv_opd; iv++) {
      const double sv = ((double)iv / (double)(nv_opd - 1) * 2.0 - 1.0) * v_max;
      double f0;
      if (p->iptcldist == 1)
        f0 = n * (sv * sv) * exp(-(sv * sv) / 2.0) / sqrt(2.0 * ORC_PI);
      else if (p->iptcldist == 2)
        f0 = n * (exp(-((sv + v0) * (sv + v0)) / (2.0 * T / m)) + exp(-((sv - v0) * (sv - v0)) / (2.0 * T / m))) /
             (sqrt(8.0 * ORC_PI) * T / m);
      else if (p->iptcldist == 3)
        f0 = n * exp(-(sv * sv) / (2.0 * T / m)) / (sqrt(2.0 * ORC_PI) * T / m) +
             (1.0 - n) * exp(-((sv - v0) * (sv - v0)) / (2.0 * T2 / m)) / (sqrt(2.0 * ORC_PI) * T2 / m);
      else
        f0 = n * exp(-((sv - v0) * (sv - v0)) / (2.0 * T / m)) / (sqrt(2.0 * ORC_PI) * T / m);
      for (int ix = 0; ix < nx_opd; ix++) pertb_xv[iv * nx_opd + ix] = total_xv[iv * nx_opd + ix] - f0;
      pertb_v[iv] = total_v[iv] - p->lx * f0;
    }
  }
  free(mxv); free(txv); free(pxv); free(mv); free(tv); free(pv_);
}

void orc_petsc_decide(int64_t n, int npe, int rank, int64_t *low, int64_t *high) {
  int64_t base = n / npe, rem = n % npe;
  int64_t lo = base * rank + (rank < rem ? rank : rem);
  *low = lo;
  *high = lo + base + (rank < rem ? 1 : 0);
}

/* src/pic1dp_particle.F90:172-264.  input_pertb_shape == 1.0 (src/pic1dp_input.F90:271). */
void orc_particle_load(const orc_params *p, int isp, int al_int, int mype, int warmup, int64_t nlocal,
                       int64_t ninit, double *x, double *v, double *pp, double *w) {
  const double T = p->temperature[isp], T2 = p->temperature2[isp], m = p->mass[isp];
  const double n = p->density[isp], v0 = p->v0[isp], lx = p->lx, vmax = p->v_max;
  const double rn = (double)ninit;
  orc_multirand *g = orc_multirand_new();
  /* the generator is initialised once per particle_load (:159) and species are drawn in sequence from it;
   * this restatement is exact for species 0 and re-seeds for later species only if called per species. */
  orc_multirand_init_const(g, al_int, mype, warmup);
  if (p->imarker == 1) { /* :172-178 */
    orc_multirand_gaussian_array(g, v, nlocal);
    for (int64_t i = 0; i < nlocal; i++) {
      v[i] = v[i] * sqrt(T / m) + v0;
      pp[i] = n * lx / rn;
    }
  } else { /* :179-219 */
    orc_multirand_real_array(g, v, nlocal);
    for (int64_t i = 0; i < nlocal; i++) v[i] = (v[i] - 0.5) * 2.0 * vmax; /* :181 */
    for (int64_t i = 0; i < nlocal; i++) {
      const double pv = v[i];
      if (p->iptcldist == 1) { /* :183-186 */
        pp[i] = n * lx * 2.0 * vmax / rn * (pv * pv) * exp(-(pv * pv) / 2.0) / sqrt(2.0 * ORC_PI);
      } else if (p->iptcldist == 2) { /* :188-196 */
        pp[i] = n * lx * 2.0 * vmax / rn *
                (exp(-((pv + v0) * (pv + v0)) / (2.0 * T / m)) + exp(-((pv - v0) * (pv - v0)) / (2.0 * T / m))) /
                sqrt(8.0 * ORC_PI * T / m);
      } else if (p->iptcldist == 3) { /* :198-209 */
        pp[i] = 1.0 * lx * 2.0 * vmax / rn *
                (n * exp(-(pv * pv) / (2.0 * T / m)) / sqrt(2.0 * ORC_PI * T / m) +
                 (1.0 - n) * exp(-((pv - v0) * (pv - v0)) / (2.0 * T2 / m)) / sqrt(2.0 * ORC_PI * T2 / m));
      } else { /* :211-217 */
        pp[i] = n * lx * 2.0 * vmax / rn * exp(-((pv - v0) * (pv - v0)) / (2.0 * T / m)) /
                sqrt(2.0 * ORC_PI * T / m);
      }
    }
  }
  orc_multirand_real_array(g, x, nlocal); /* :222 */
  for (int64_t i = 0; i < nlocal; i++) x[i] = x[i] * lx; /* :223 */
  for (int64_t i = 0; i < nlocal; i++) w[i] = 0.0;       /* :225 */
  for (int im = 0; im < p->init_nmode; im++) {           /* :226-232 */
    for (int64_t i = 0; i < nlocal; i++) {
      double k = 2.0 * ORC_PI / lx * (double)p->init_mode[im];
      w[i] = w[i] + p->init_mode_cos[im] * cos(k * x[i]) + p->init_mode_sin[im] * sin(k * x[i]);
    }
  }
  for (int64_t i = 0; i < nlocal; i++) w[i] = w[i] * pp[i] * 1.0; /* :235-236 */
  if (p->linear == 0)                                             /* :260-263 VecAXPY(p, 1.0, w) */
    for (int64_t i = 0; i < nlocal; i++) pp[i] = pp[i] + 1.0 * w[i];
  orc_multirand_free(g);
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Time loop of src/pic1dp.F90:78-93 over emulated ranks.  The per-rank passes (push, deposit) run on
 * persistent pthreads (rank r is served by thread r mod nthreads, like one MPI process per core); the
 * rank-ordered grid sum, scaling and field solve are done once by the calling thread (every rank would
 * compute the same values). */
typedef struct run_ctx {
  const orc_params *p;
  int nranks, nthreads, irk, stop;
  orc_rank_state *ranks;
  double *c1_all;
  const double *E;
  pthread_barrier_t go, done;
} run_ctx;

typedef struct run_worker {
  run_ctx *ctx;
  int tid;
} run_worker;

static void rank_substep(run_ctx *c, int r) {
  const orc_params *p = c->p;
  const int nx = p->nx, S = p->nspecies, nranks = c->nranks;
  /* interaction_push_particle */
  for (int s = 0; s < S; s++) {
    orc_rank_state *st = &c->ranks[s * nranks + r];
    orc_push_species(p, s, c->irk, st->np, st->x, st->v, st->p, st->w, st->xb, st->vb, st->wb, c->E);
  }
  /* interaction_collect_charge, local part (:83-114) */
  for (int s = 0; s < S; s++) {
    orc_rank_state *st = &c->ranks[s * nranks + r];
    deposit_species(p, st->np, st->x, p->deltaf == 1 ? st->w : st->p, c->c1_all + ((size_t)s * nranks + r) * nx,
                    p->iptclshape <= 2);
  }
}

static void *run_worker_main(void *arg) {
  run_worker *w = (run_worker *)arg;
  run_ctx *c = w->ctx;
  for (;;) {
    pthread_barrier_wait(&c->go);
    if (c->stop) break;
    for (int r = w->tid; r < c->nranks; r += c->nthreads) rank_substep(c, r);
    pthread_barrier_wait(&c->done);
  }
  return NULL;
}

double orc_run(const orc_params *p, int nranks, orc_rank_state *ranks, int nsteps, double *rho, double *E,
               double *mode_re, double *mode_im, double *energy_out, int nthreads) {
  const int nx = p->nx, M = p->nmode, S = p->nspecies;
  double *F_re = (double *)malloc(sizeof(double) * (size_t)nx * M);
  double *F_im = (double *)malloc(sizeof(double) * (size_t)nx * M);
  double *ginv = (double *)malloc(sizeof(double) * M);
  double *c1_all = (double *)calloc((size_t)S * nranks * nx, sizeof(double));
  orc_field_init(p, F_re, F_im, ginv);
  if (nthreads <= 0 || nthreads > nranks) nthreads = nranks;
  run_ctx ctx;
  ctx.p = p;
  ctx.nranks = nranks;
  ctx.nthreads = nthreads;
  ctx.irk = 1;
  ctx.stop = 0;
  ctx.ranks = ranks;
  ctx.c1_all = c1_all;
  ctx.E = E;
  /* thread 0 is the caller */
  pthread_barrier_init(&ctx.go, NULL, (unsigned)nthreads);
  pthread_barrier_init(&ctx.done, NULL, (unsigned)nthreads);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  run_worker *wk = (run_worker *)malloc(sizeof(run_worker) * (size_t)nthreads);
  for (int t = 1; t < nthreads; t++) {
    wk[t].ctx = &ctx;
    wk[t].tid = t;
    pthread_create(&th[t], NULL, run_worker_main, &wk[t]);
  }
  const double t0 = now_s();
  for (int it = 0; it < nsteps; it++) {
    for (int irk = 1; irk <= 2; irk++) {
      ctx.irk = irk;
      pthread_barrier_wait(&ctx.go);
      for (int r = 0; r < nranks; r += nthreads) rank_substep(&ctx, r);
      pthread_barrier_wait(&ctx.done);
      combine_charge(p, nranks, c1_all, rho);
      orc_field_solve(p, F_re, F_im, ginv, rho, E, mode_re, mode_im);
    }
    if (energy_out) energy_out[it] = orc_field_energy(p, E);
  }
  const double t1 = now_s();
  ctx.stop = 1;
  pthread_barrier_wait(&ctx.go);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  pthread_barrier_destroy(&ctx.go);
  pthread_barrier_destroy(&ctx.done);
  free(th);
  free(wk);
  free(F_re);
  free(F_im);
  free(ginv);
  free(c1_all);
  return t1 - t0;
}
