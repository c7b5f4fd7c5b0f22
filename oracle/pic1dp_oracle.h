/*
 * oracle/pic1dp_oracle.h -- CPU restatement of the PIC1D-PETSc per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (pic1dp_b200/) never links, imports or executes this code.
 *
 * PARITY STATUS: "parity unpinned" for the hot path.  The reference (Fortran 2003 + PETSc + MPI)
 * cannot be compiled here (no gfortran/mpif90/PETSc in the image) and holds no golden vectors for
 * deposit / push / field solve.  The only stored vectors in the reference are the multirand RNG
 * known-answer sequences (src/multirand.F90:396-425); the RNG restatement in multirand_oracle.c
 * is pinned by them.  The hot-path restatement follows the reference statement by statement,
 * left-to-right, no FMA contraction (build with -O3 -ffp-contract=off), glibc exp/sin/cos/fmod.
 *
 * All file:line citations are into /root/reference/.
 */
#ifndef PIC1DP_ORACLE_H
#define PIC1DP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SPECIES 4
#define ORC_MAX_MODES 64

/* run-time copy of the compile-time parameters of src/pic1dp_input.F90:32-256 that the hot path reads */
typedef struct orc_params {
  int32_t nx;                         /* input_nx            :128 */
  int32_t nmode;                      /* input_nmode         :75  */
  int32_t modes[ORC_MAX_MODES];       /* input_modes         :79-80 */
  double lx;                          /* input_lx            :46-47 */
  double dt;                          /* input_dt            :109 */
  int32_t nspecies;                   /* input_nspecies      :57 */
  double charge[ORC_MAX_SPECIES];     /* input_species_*     :66-72 */
  double mass[ORC_MAX_SPECIES];
  double temperature[ORC_MAX_SPECIES];
  double temperature2[ORC_MAX_SPECIES];
  double density[ORC_MAX_SPECIES];
  double v0[ORC_MAX_SPECIES];
  int32_t iptcldist;                  /* input_iptcldist     :54  (0 Maxwellian,1 two-stream1,2 two-stream2,3 bump-on-tail) */
  int32_t deltaf;                     /* input_deltaf        :106 */
  int32_t linear;                     /* input_linear        :43 */
  int32_t iptclshape;                 /* input_iptclshape    :138 (1,2 matrix; 3 cached arrays; 4 inline) */
  double v_max;                       /* input_v_max         :125 (loader only) */
  int32_t imarker;                    /* input_imarker       :122 (loader only) */
  int32_t init_nmode;                 /* input_init_nmode    :87 */
  int32_t init_mode[ORC_MAX_MODES];   /* input_init_mode     :90-91 */
  double init_mode_cos[ORC_MAX_MODES];/* :96-98 */
  double init_mode_sin[ORC_MAX_MODES];
} orc_params;

/* fill *p with the defaults of src/pic1dp_input.F90 (electron bump-on-tail, PRE 83 056402 V.A.2) */
void orc_params_default(orc_params *p);

/* ---- field operators: src/pic1dp_field.F90:158-210 ---- */
/* F_re, F_im: nx*nmode row-major [j*nmode+m]; grad_inv: nmode */
void orc_field_init(const orc_params *p, double *F_re, double *F_im, double *grad_inv);

/* ---- field solve: src/pic1dp_field.F90:231-256 (sequential-AIJ summation order, one rank) ---- */
void orc_field_solve(const orc_params *p, const double *F_re, const double *F_im, const double *grad_inv,
                     const double *rho, double *E, double *mode_re, double *mode_im);

/* ---- weights: src/pic1dp_interaction.F90:101-108 / src/pic1dp_particle.F90:308-323 ----
 * wraps x in place exactly as the reference does, returns left cell index and both weights.
 * right_frac != 0 selects the matrix-mode (iptclshape 1,2) right weight `frac`; else `1-(1-frac)`. */
void orc_shape(const orc_params *p, int64_t np, double *x, int32_t *ix, double *s_left, double *s_right,
               int right_frac);

/* ---- deposit of one species on one rank: src/pic1dp_interaction.F90:83-114 (array path)
 * charge1[nx] is zeroed first, x is wrapped in place (iptclshape==4 semantics applied for all shapes,
 * which is idempotent for shapes 1-3 whose x was already wrapped by particle_compute_shape_x).
 * Returns the number of markers whose wrapped x landed exactly on lx (ix == nx, reference OOB, SURVEY App.A-7);
 * those are deposited as ix=0,s=1 (defined behaviour of this restatement). */
int64_t orc_deposit_species(const orc_params *p, int64_t np, double *x, const double *w, double *charge1);

/* ---- whole collect_charge over emulated ranks: src/pic1dp_interaction.F90:79-151 (and :46-78 for shapes 1,2)
 * x[isp][rank] pointers are flattened: arrays of nspecies*nranks pointers, np likewise.
 * wsrc is w (deltaf) or p (full-f) as the reference selects at :84-91.  rho[nx] out. */
int64_t orc_collect_charge(const orc_params *p, int nranks, const int64_t *np, double **x, double **wsrc,
                           double *rho);

/* ---- push of one species on one rank: src/pic1dp_interaction.F90:178-193, 238-339 ----
 * irk = 1 or 2 (global_irk).  At irk==1 the backup copy (VecCopy, :181-187) is done here. */
void orc_push_species(const orc_params *p, int isp, int irk, int64_t np,
                      double *x, double *v, const double *pw_p, double *w,
                      double *xb, double *vb, double *wb, const double *E);

/* -d f0/dv / f0 for one velocity: src/pic1dp_interaction.F90:275-326 */
double orc_dlnf0(const orc_params *p, int isp, double v);

/* ---- diagnostics: src/pic1dp_output.F90:117-124 ---- */
double orc_field_energy(const orc_params *p, const double *E);

/* ---- output_field scalars: src/pic1dp_output.F90:117-172.  Per-rank sequential sums (VecSum), added in rank order.
 * out[0] = energy, out[1+3s..3+3s] per species.  x/v/p/w: [nspecies*nranks] pointers, species-major. ---- */
void orc_output_field(const orc_params *p, int nranks, const int64_t *np, double **v, double **pp, double **w,
                      const double *E, double *out);

/* ---- output_ptcldist for one species: src/pic1dp_output.F90:196-477 (per-rank histograms, MPI_Reduce in rank
 * order, scaling, linear / full-f post-processing).  Arrays sized nx_opd*nv_opd and nv_opd. ---- */
void orc_output_ptcldist(const orc_params *p, int isp, int nranks, const int64_t *np, double **x, double **v,
                         double **pp, double **w, int nx_opd, int nv_opd, double v_max, double *markr_xv,
                         double *total_xv, double *pertb_xv, double *markr_v, double *total_v, double *pertb_v);

/* ---- loader: src/pic1dp_particle.F90:172-264 with multirand (seed_type 1, constant seeds) ----
 * nlocal markers of species isp for rank mype; arrays x,v,pp,w length nlocal. */
void orc_particle_load(const orc_params *p, int isp, int al_int, int mype, int warmup, int64_t nlocal,
                       int64_t nparticle_init_total, double *x, double *v, double *pp, double *w);

/* PETSC_DECIDE block split requested at src/pic1dp_particle.F90:91: rank r owns N/npe + (r < N mod npe) */
void orc_petsc_decide(int64_t n, int npe, int rank, int64_t *low, int64_t *high);

/* ---- whole-run driver in the reference's structure (src/pic1dp.F90:63-109), emulated ranks run as
 * OpenMP threads: separate push and deposit passes, explicit backup copy, per-rank private grids summed in
 * rank order (stands in for MPI_Allreduce, :132-133).  State is caller-owned.
 * Runs nsteps timesteps; if energy_out != NULL stores the field energy after each step (nsteps values).
 * Returns wall seconds spent in the time loop. */
typedef struct orc_rank_state {
  int64_t np;
  double *x, *v, *p, *w, *xb, *vb, *wb;
} orc_rank_state;

double orc_run(const orc_params *p, int nranks, orc_rank_state *ranks /* [nspecies*nranks], species-major */,
               int nsteps, double *rho, double *E, double *mode_re, double *mode_im, double *energy_out,
               int nthreads);

/* ---- multirand restatement (multirand_oracle.c): src/multirand.F90 ---- */
typedef struct orc_multirand orc_multirand;
orc_multirand *orc_multirand_new(void);
void orc_multirand_free(orc_multirand *g);
/* al_int: 1 KISS64, 2 MT19937-64, 3 SuperKISS64.  Default seeds of the self test (:481-515). */
void orc_multirand_seed_default(orc_multirand *g, int al_int);
/* constant-seed path (seed_type 1) + warm-up: src/multirand.F90:301-381 */
void orc_multirand_init_const(orc_multirand *g, int al_int, int mype, int warmup);
int64_t orc_multirand_int64(orc_multirand *g);
void orc_multirand_get_seeds4(const orc_multirand *g, uint64_t out[4]); /* multirand_seeds(0:3), KISS64 state */
void orc_multirand_skip(orc_multirand *g, int64_t n);                   /* discard n outputs */
double orc_multirand_real64(orc_multirand *g);            /* INT2REAL64, :49 */
void orc_multirand_real_array(orc_multirand *g, double *a, int64_t n);      /* :664-690 */
void orc_multirand_gaussian_array(orc_multirand *g, double *a, int64_t n);  /* :838-872 */

/* ---- marker optimisation (optimize_oracle.c): src/pic1dp_particle.F90:356-746 ----
 * dist = particle_dist_pertb_abs_v(ispecies, 0:nv-1) of one species; merge / remove / split act on the marker
 * arrays of one species on one rank (1-based visiting order of the reference) and return the new particle_np. */
void orc_dist_pertb_abs_v_rank(int64_t np, const double *v, const double *w, int nv, double v_max, double *dist);
void orc_dist_pertb_abs_v(int nranks, const int64_t *np, double **v, double **w, int nv, double v_max, double *dist);
int64_t orc_particle_merge(const orc_params *p, int64_t np, double *x, double *v, double *pp, double *w,
                           const double *dist, int nv, double v_max, double thsh);
int64_t orc_particle_remove(int64_t np, double *x, double *v, double *pp, double *w, const double *dist, int nv,
                            double v_max, double thsh, int typeremove, double remove_frac, orc_multirand *rng);
int64_t orc_particle_split(const orc_params *p, int64_t np, int64_t capacity, double *x, double *v, double *pp,
                           double *w, const double *dist, int nv, double v_max, double thsh, int ngroup,
                           double dv_sig_frac, orc_multirand *rng);

#ifdef __cplusplus
}
#endif
#endif
