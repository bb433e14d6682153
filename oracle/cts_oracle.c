/* cts_oracle.c -- plain-C CPU restatement of the ClimaTimeSteppers per-step state-update path on the
 * synthetic benchmark models, in the REFERENCE'S PASS STRUCTURE: one loop per broadcast the
 * reference materialises (imex_ark.jl:69-308, fused_increment.jl:220-236, newtons_method.jl:609-665,
 * lsrk.jl:56-67).  Float64 only.
 *
 * *** TEST INFRASTRUCTURE ONLY ***  It is the full-size parity checker for the CUDA path and the CPU
 * baseline timed by bench.py (`cpu_baseline`, `--impl reference`).  Nothing in the product links or
 * calls it.  It is itself validated bit-for-bit against the numpy oracle (oracle/cts_ref.py +
 * oracle/cts_models.py, which are pinned to the reference's golden vectors) in tests/test_oracle_c.py.
 *
 * Build: gcc -O3 -march=native -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps every a*b+c un-fused, as Julia does without @fastmath/muladd.
 * OpenMP only splits independent loop iterations: results do not depend on the thread count.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXS 8

typedef struct {
  int nlev;
  size_t nx, ny_local;
  int halo;
  double dz, nu;
  int nonlinear;
  const double* K;  /* per local column (ghost rows included) */
  const double* S;  /* nlev source profile or NULL */
} col_model;

typedef struct {
  int s;
  double a_exp[MAXS * MAXS], b_exp[MAXS], c_exp[MAXS], a_imp[MAXS * MAXS], b_imp[MAXS], c_imp[MAXS];
} tableau;

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- synthetic inputs: splitmix64(seed ^ idx) -> [0,1) ------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
void oracle_fill_hash(double* out, size_t n, uint64_t seed, uint64_t offset, double scale, double shift) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    double u = (double)(splitmix64(seed ^ (offset + i)) >> 11) * (1.0 / 9007199254740992.0);
    out[i] = scale * (u + shift);
  }
}

/* ---- K1: out = (base + G1) + G2, zero coefficients dropped, left folds (fused_increment.jl:90-91,220-236) ---- */
static void fused_assign(double* out, const double* base, size_t n, int n1, const double* c1, double* const* t1, int n2,
                         const double* c2, double* const* t2) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    double r = base[i];
    if (n1 > 0) {
      double g = c1[0] * t1[0][i];
      for (int k = 1; k < n1; ++k) g = g + c1[k] * t1[k][i];
      r = r + g;
    }
    if (n2 > 0) {
      double g = c2[0] * t2[0][i];
      for (int k = 1; k < n2; ++k) g = g + c2[k] * t2[k][i];
      r = r + g;
    }
    out[i] = r;
  }
}

/* gather the non-zero (dt*coef, tendency) pairs of one row/vector of a tableau */
static int gather(double dt, const double* coefs, int count, double* const* tend, double* c_out, double** t_out) {
  int n = 0;
  for (int j = 0; j < count; ++j) {
    if (coefs[j] == 0.0) continue;
    c_out[n] = dt * coefs[j];
    t_out[n] = tend[j];
    ++n;
  }
  return n;
}

/* ---- column model ------------------------------------------------------------------------------------ */
static inline double kface(double a, double ua, double ub) {
  double ubar = 0.5 * (ua + ub);
  return a * (1.0 + ubar * ubar);
}

void oracle_col_T_imp(const col_model* m, double* out, const double* u) {
  const size_t ncols = (m->ny_local + 2 * (size_t)m->halo) * m->nx;
  const int nl = m->nlev;
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < ncols; ++c) {
    const double a = m->K[c] / (m->dz * m->dz);
    const double* U = u + c * nl;
    double* o = out + c * nl;
    for (int l = 0; l < nl; ++l) {
      double um = l > 0 ? U[l - 1] : 0.0, uc = U[l], up = l < nl - 1 ? U[l + 1] : 0.0;
      if (!m->nonlinear)
        o[l] = a * ((um - 2.0 * uc) + up);
      else
        o[l] = kface(a, uc, up) * (up - uc) - kface(a, um, uc) * (uc - um);
    }
  }
}

void oracle_col_Wfact(const col_model* m, double* dl, double* d, double* du, const double* u, double dtg) {
  const size_t ncols = (m->ny_local + 2 * (size_t)m->halo) * m->nx;
  const int nl = m->nlev;
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < ncols; ++c) {
    const double a = m->K[c] / (m->dz * m->dz);
    const double* U = u + c * nl;
    for (int l = 0; l < nl; ++l) {
      size_t i = c * nl + l;
      if (!m->nonlinear) {
        double o = dtg * a;
        dl[i] = o;
        du[i] = o;
        d[i] = (-2.0 * o) - 1.0;
      } else {
        double um = l > 0 ? U[l - 1] : 0.0, uc = U[l], up = l < nl - 1 ? U[l + 1] : 0.0;
        double lo = dtg * kface(a, um, uc), hi = dtg * kface(a, uc, up);
        dl[i] = lo;
        du[i] = hi;
        d[i] = (-(lo + hi)) - 1.0;
      }
    }
  }
}

void oracle_col_T_exp(const col_model* m, double* out, const double* u, double t) {
  const int nl = m->nlev, h = m->halo;
  const size_t nx = m->nx, ny = m->ny_local, rows = ny + 2 * (size_t)h, re = nx * nl;
  const double g = 1.0 + 0.5 * sin(2.0 * M_PI * t);
#pragma omp parallel for schedule(static)
  for (size_t jr = 0; jr < ny; ++jr) {
    const size_t j = jr + h;
    const size_t js = h ? j - 1 : (j == 0 ? rows - 1 : j - 1), jn = h ? j + 1 : (j == rows - 1 ? 0 : j + 1);
    for (size_t ix = 0; ix < nx; ++ix) {
      const size_t ixw = ix == 0 ? nx - 1 : ix - 1, ixe = ix == nx - 1 ? 0 : ix + 1;
      for (int l = 0; l < nl; ++l) {
        double val = m->S ? m->S[l] * g : 0.0;
        if (m->nu != 0.0) {
          double uc = u[j * re + ix * nl + l];
          double uw = u[j * re + ixw * nl + l], ue = u[j * re + ixe * nl + l];
          double us = u[js * re + ix * nl + l], un = u[jn * re + ix * nl + l];
          double lap = (((uw + ue) + us) + un) - 4.0 * uc;
          val = val + m->nu * lap;
        }
        out[j * re + ix * nl + l] = val;
      }
    }
  }
}

void oracle_col_dss(const col_model* m, double* u) { /* single rank: periodic ghost refresh */
  if (!m->halo) return;
  const size_t re = m->nx * (size_t)m->nlev, ny = m->ny_local;
  memcpy(u, u + ny * re, re * sizeof(double));
  memcpy(u + (ny + 1) * re, u + re, re * sizeof(double));
}

/* two-sided, product-form Thomas elimination: the operation sequence of csrc/tile_device.cuh / cts_models.babe_solve */
static inline double pow2_rescale(double p) {
  uint64_t bits;
  memcpy(&bits, &p, 8);
  int be = (int)((bits >> 52) & 0x7ff);
  if (be == 0 || be >= 2046) return 1.0;
  uint64_t out = (uint64_t)(2046 - be) << 52;
  double sc;
  memcpy(&sc, &out, 8);
  return sc;
}
static inline double rcp_checked(double p) { /* csrc/tile_device.cuh::rcp_checked */
  uint64_t bits;
  memcpy(&bits, &p, 8);
  int be = (int)((bits >> 52) & 0x7ff);
  if (be < 25 || be > 2021) return NAN;
  return 1.0 / p;
}
typedef struct {
  double P1, P2, S1, aprev;
} sweep_t;
static inline void sweep_row(sweep_t* w, double toward, double dd, double away, double rr, int s, double* fac,
                             double* rhs) {
  const double lu = toward * w->aprev;
  const double P = dd * w->P1 - lu * w->P2;
  const double S = rr * w->P1 - toward * w->S1;
  const double inv = rcp_checked(P);
  *fac = (away * w->P1) * inv;
  *rhs = S * inv;
  w->P2 = w->P1;
  w->P1 = P;
  w->S1 = S;
  w->aprev = away;
  if (((s + 1) & 7) == 0) { /* RescaleEvery<double> = 8 */
    const double sc = pow2_rescale(w->P1);
    w->P1 = w->P1 * sc;
    w->P2 = w->P2 * sc;
    w->S1 = w->S1 * sc;
  }
}
static void babe_column(const double* dl, const double* d, const double* du, const double* rhs, double* x, double* fac,
                        int n) {
  if (n == 1) {
    x[0] = rhs[0] * (1.0 / d[0]);
    return;
  }
  const int m = n / 2;
  double fp = 0.0, rp = 0.0;
  sweep_t w = {1.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < m; ++i) {
    sweep_row(&w, i == 0 ? 0.0 : dl[i], d[i], du[i], rhs[i], i, &fp, &rp);
    fac[i] = fp;
    x[i] = rp;
  }
  const double cpm = fp, rpm = rp;
  sweep_t w2 = {1.0, 0.0, 0.0, 0.0};
  for (int j = n - 1, s = 0; j >= m; --j, ++s) {
    sweep_row(&w2, j == n - 1 ? 0.0 : du[j], d[j], dl[j], rhs[j], s, &fp, &rp);
    fac[j] = fp;
    x[j] = rp;
  }
  const double lpm = fp, spm = rp;
  const double det = 1.0 - cpm * lpm;
  const double xm1 = (rpm - cpm * spm) / det;
  const double xm = spm - lpm * xm1;
  x[m - 1] = xm1;
  x[m] = xm;
  for (int i = m - 2; i >= 0; --i) x[i] = x[i] - fac[i] * x[i + 1];
  for (int j = m + 1; j < n; ++j) x[j] = x[j] - fac[j] * x[j - 1];
}

void oracle_tridiag_solve(int nlev, size_t ncols, const double* dl, const double* d, const double* du, const double* rhs,
                          double* x) {
#pragma omp parallel
  {
    double* fac = (double*)malloc(sizeof(double) * (size_t)nlev);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)nlev);
#pragma omp for schedule(static)
    for (size_t c = 0; c < ncols; ++c) {
      size_t o = c * (size_t)nlev;
      babe_column(dl + o, d + o, du + o, rhs + o, tmp, fac, nlev);
      memcpy(x + o, tmp, sizeof(double) * (size_t)nlev);
    }
    free(fac);
    free(tmp);
  }
}

/* ---- IMEX ARK step on the column model (imex_ark.jl:69-272), direct Newton --------------------------- */
typedef struct {
  size_t n;
  tableau tab;
  double *U, *temp, *dx, *f, *dl, *d, *du;
  double* T_exp[MAXS];
  double* T_imp[MAXS];
  int max_iters;
  int update_j; /* 0 NewTimeStep, 1 NewNewtonSolve, 2 NewNewtonIteration (cts_signal numbering) */
  double gamma;
} ark_cache;

static int col_nonzero(const double* a, int s, int j) {
  for (int i = 0; i < s; ++i)
    if (a[i * s + j] != 0.0) return 1;
  return 0;
}

ark_cache* oracle_ark_create(const col_model* m, const tableau* tab, int max_iters, int update_j) {
  ark_cache* c = (ark_cache*)calloc(1, sizeof(ark_cache));
  c->n = (m->ny_local + 2 * (size_t)m->halo) * m->nx * (size_t)m->nlev;
  c->tab = *tab;
  c->max_iters = max_iters;
  c->update_j = update_j;
  const int s = tab->s;
  double** all[] = {&c->U, &c->temp, &c->dx, &c->f, &c->dl, &c->d, &c->du};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) *all[k] = (double*)calloc(c->n, sizeof(double));
  for (int i = 0; i < s; ++i) {
    if (col_nonzero(tab->a_exp, s, i) || tab->b_exp[i] != 0.0) c->T_exp[i] = (double*)calloc(c->n, sizeof(double));
    if (col_nonzero(tab->a_imp, s, i) || tab->b_imp[i] != 0.0) c->T_imp[i] = (double*)calloc(c->n, sizeof(double));
    if (tab->a_imp[i * s + i] != 0.0) c->gamma = tab->a_imp[i * s + i];
  }
  return c;
}

void oracle_ark_destroy(ark_cache* c) {
  if (!c) return;
  free(c->U); free(c->temp); free(c->dx); free(c->f); free(c->dl); free(c->d); free(c->du);
  for (int i = 0; i < MAXS; ++i) {
    free(c->T_exp[i]);
    free(c->T_imp[i]);
  }
  free(c);
}

double* oracle_ark_buffer(ark_cache* c, const char* name, int idx) {
  if (!strcmp(name, "U")) return c->U;
  if (!strcmp(name, "T_exp")) return c->T_exp[idx];
  if (!strcmp(name, "T_imp")) return c->T_imp[idx];
  return NULL;
}

void oracle_ark_step(ark_cache* c, const col_model* m, double* u, double t, double dt) {
  const tableau* tab = &c->tab;
  const int s = tab->s;
  const size_t n = c->n;
  double c1[MAXS], c2[MAXS];
  double *t1[MAXS], *t2[MAXS];
  if (c->update_j == 0) oracle_col_Wfact(m, c->dl, c->d, c->du, u, dt * c->gamma); /* async_utils.jl:11-31 */
  for (int i = 0; i < s; ++i) {
    const double t_exp = t + dt * tab->c_exp[i];
    const double dtg = dt * tab->a_imp[i * s + i];
    double* U = c->U;
    /* compute_stage_value! :140-165 */
    int n1 = gather(dt, &tab->a_exp[i * s], i, c->T_exp, c1, t1);
    int n2 = gather(dt, &tab->a_imp[i * s], i, c->T_imp, c2, t2);
    fused_assign(U, u, n, n1, c1, t1, n2, c2, t2);
    /* solve_stage_implicit! :174-242 (default hooks: initialize_imp! = Returns(nothing) is not `nothing`) */
    if (i != 0) oracle_col_dss(m, U);
    if (dtg != 0.0) {
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < n; ++k) c->temp[k] = U[k]; /* K2 :209 */
      oracle_col_dss(m, U);                             /* :213 */
      /* solve_newton! newtons_method.jl:609-665 */
      if (c->update_j == 1) oracle_col_Wfact(m, c->dl, c->d, c->du, U, dtg);
      oracle_col_T_imp(m, c->f, U);
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < n; ++k) c->f[k] = (c->temp[k] + dtg * c->f[k]) - U[k]; /* K3 */
      for (int it = 1; it <= c->max_iters; ++it) {
        if (c->update_j == 2) oracle_col_Wfact(m, c->dl, c->d, c->du, U, dtg);
        oracle_tridiag_solve(m->nlev, n / (size_t)m->nlev, c->dl, c->d, c->du, c->f, c->dx); /* K4 */
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; ++k) U[k] = U[k] - c->dx[k]; /* K5 */
        if (it < c->max_iters) {
          oracle_col_T_imp(m, c->f, U);
#pragma omp parallel for schedule(static)
          for (size_t k = 0; k < n; ++k) c->f[k] = (c->temp[k] + dtg * c->f[k]) - U[k];
        }
      }
      oracle_col_dss(m, U); /* :234 */
    }
    /* evaluate_stage_tendencies! :252-272 */
    if (c->T_exp[i]) oracle_col_T_exp(m, c->T_exp[i], U, t_exp);
    if (c->T_imp[i]) {
      if (dtg == 0.0)
        oracle_col_T_imp(m, c->T_imp[i], U);
      else {
        double* Ti = c->T_imp[i];
#pragma omp parallel for schedule(static)
        for (size_t k = 0; k < n; ++k) Ti[k] = (U[k] - c->temp[k]) / dtg; /* K6 */
      }
    }
  }
  int n1 = gather(dt, tab->b_exp, s, c->T_exp, c1, t1);
  int n2 = gather(dt, tab->b_imp, s, c->T_imp, c2, t2);
  fused_assign(u, u, n, n1, c1, t1, n2, c2, t2); /* :90 */
  oracle_col_dss(m, u);                          /* :98 */
}

/* ---- LSRK 2N on 1-D periodic upwind advection (lsrk.jl:56-67), two passes per stage as in the reference ---- */
void oracle_lsrk_advection_step(double* u, double* du, size_t n, double c, double dx, double dt, int ns, const double* A,
                                const double* B) {
  const double negc = -c;
  for (int s = 0; s < ns; ++s) {
    const double beta = A[s], dtB = dt * B[s];
    const double wrap = u[n - 1];
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) { /* user f(du,u,p,t,1,A[s]): du = 1*rhs + A*du */
      double ul = i == 0 ? wrap : u[i - 1];
      double rhs = (negc * (u[i] - ul)) / dx;
      du[i] = 1.0 * rhs + beta * du[i];
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) u[i] = u[i] + dtB * du[i]; /* lsrk.jl:65 */
  }
}
