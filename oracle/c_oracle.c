/*
 * c_oracle.c -- plain-C restatement of the reference finite-volume step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as a
 * second, independent checker and by bench.py as the CPU baseline
 * ("cpu_baseline", "--impl reference").  Never linked into the product.
 *
 * Every function follows a reference function (paths relative to the
 * reference checkout) and keeps its operation order; build with
 * -ffp-contract=off so that no multiply-add is fused (NumPy rounds after
 * every ufunc).  The ARZ relaxation uses the closed-form root of the affine
 * system the reference hands to fsolve (arz.py:462-463, 469-492), exactly as
 * oracle/np_oracle.py documents.
 *
 * Parity status: bit-equal to oracle/np_oracle.py (and through it pinned to
 * the golden vectors generated from the reference) for lam in {1, 2, 0.5}.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { BC_LOOP = 0, BC_EXTEND = 1, BC_CONSTANT = 2, BC_HALO = 3 };

typedef struct {
  double dx, dt, rho_max, v_max, tau, lam, rho_crit;
  int bc_l, bc_r;
  double bl0, bl1, br0, br1;
} oracle_params;

/* (1 - x) ** lam with NumPy's exact fast paths (lwr.py:316, arz.py:256) */
static inline double powlam(double t, double lam) {
  if (lam == 1.0) return t;
  if (lam == 2.0) return t * t;
  if (lam == 0.5) return sqrt(t);
  return pow(t, lam);
}

/* Greenshields: v_max * ((1 - rho / rho_max) ** lam) */
static inline double v_eq(const oracle_params* p, double rho) {
  return p->v_max * powlam(1.0 - (rho / p->rho_max), p->lam);
}

/* numpy.minimum: NaN-propagating */
static inline double npmin(double a, double b) { return (a < b || a != a) ? a : b; }

/* ---- LWR (lwr.py:192-316) ------------------------------------------------ */
static void lwr_row(const oracle_params* p, const double* in, double* out, long n,
                    double* f) {
  const double step = p->dt / p->dx;
  const double rho_crit = p->rho_max / 2;                       /* lwr.py:282 */
  /* demand of cell i and supply of cell i+1 (lwr.py:284-297) */
  double s_next, d;
  for (long i = 0; i < n; ++i) {
    const double rho = in[i], ve = v_eq(p, rho), qmax = rho_crit * ve;
    d = rho * ve * (double)(rho < rho_crit) + qmax * (double)(rho >= rho_crit);
    const long j = (i + 1 < n) ? i + 1 : i;                      /* lwr.py:294 */
    const double rj = in[j], vj = v_eq(p, rj), qj = rho_crit * vj;
    s_next = rj * vj * (double)(rj > rho_crit) + qj * (double)(rj <= rho_crit);
    f[i] = npmin(d, s_next);
  }
  for (long i = 0; i < n; ++i) {
    const double fm = (i == 0) ? f[0] : f[i - 1];                /* lwr.py:231 */
    out[i] = in[i] - step * (f[i] - fm);                         /* lwr.py:234 */
  }
  /* boundary cells (lwr.py:237-260); the rule of each end uses UPDATED values */
  const double upd_last = out[n - 1], upd_second_last = out[n - 2];
  if (p->bc_l == BC_LOOP) out[0] = upd_last;
  else if (p->bc_l == BC_CONSTANT) out[0] = p->bl0;
  else if (p->bc_l == BC_HALO) out[0] = in[0];
  if (p->bc_r == BC_LOOP) out[n - 1] = upd_second_last;
  else if (p->bc_r == BC_CONSTANT) out[n - 1] = p->br0;
  else if (p->bc_r == BC_HALO) out[n - 1] = in[n - 1];
  for (long i = 0; i < n; ++i) out[n + i] = v_eq(p, out[i]);    /* lwr.py:201 */
}

/* ---- ARZ (arz.py:205-467) ------------------------------------------------ */
static void arz_row(const oracle_params* p, const double* in, double* out, long n,
                    double* y, double* q, double* pf) {
  const double step = p->dt / p->dx;
  const double rc = p->rho_crit;
  const double qc = rc * v_eq(p, rc);                            /* arz.py:396 */
  const double c = 1 + p->dt / p->tau;
  const double* rho = in;
  const double* v = in + n;
  for (long i = 0; i < n; ++i)                                   /* arz.py:238 */
    y[i] = (v[i] - v_eq(p, rho[i])) * rho[i];
  for (long i = 0; i < n; ++i) {                                 /* arz.py:395-411 */
    const double ve = v_eq(p, rho[i]);
    const double qmax = qc + y[i];
    const double d = (rho[i] * ve + y[i]) * (double)(rho[i] < rc) + qmax * (double)(rho[i] >= rc);
    const long j = (i + 1 < n) ? i + 1 : i;                      /* arz.py:405 */
    const double vj = v_eq(p, rho[j]);
    const double qj = qc + y[j];
    const double s = (rho[j] * vj + y[j]) * (double)(rho[j] > rc) + qj * (double)(rho[j] <= rc);
    q[i] = npmin(d, s);
    pf[i] = q[i] * (y[i] / rho[i]);
  }
  double* rn = out;
  double* vn = out + n;
  /* interior cells (arz.py:456, closed-form root, arz.py:464-465); vn holds y' */
  for (long i = 1; i < n - 1; ++i) {
    rn[i] = rho[i] + (step * (q[i - 1] - q[i]));
    vn[i] = (y[i] + (step * (pf[i - 1] - pf[i]))) / c;
  }
  /* boundary cells (arz.py:316-360) */
  int pass_l = 0, pass_r = 0;
  if (p->bc_l == BC_LOOP) { rn[0] = rho[n - 1]; vn[0] = y[n - 1]; }
  else if (p->bc_l == BC_EXTEND) { rn[0] = rn[1]; vn[0] = vn[1]; }
  else if (p->bc_l == BC_CONSTANT) { rn[0] = p->bl0; vn[0] = p->bl1; }
  else pass_l = 1;
  if (p->bc_r == BC_LOOP) { rn[n - 1] = rho[n - 2]; vn[n - 1] = y[n - 2]; }
  else if (p->bc_r == BC_EXTEND) { rn[n - 1] = rn[n - 2]; vn[n - 1] = vn[n - 2]; }
  else if (p->bc_r == BC_CONSTANT) { rn[n - 1] = p->br0; vn[n - 1] = p->br1; }
  else pass_r = 1;
  /* u(): zero patch, y/rho + V(rho)  (arz.py:258-277) */
  for (long i = pass_l; i < n - pass_r; ++i) {
    if (rn[i] == 0) rn[i] = 0.0000000001;
    vn[i] = (vn[i] / rn[i]) + v_eq(p, rn[i]);
  }
  if (pass_l) { rn[0] = rho[0]; vn[0] = v[0]; }
  if (pass_r) { rn[n - 1] = rho[n - 1]; vn[n - 1] = v[n - 1]; }
}

/* model: 0 = LWR, 1 = ARZ.  obs_in/obs_out: (B, 2N) row-major doubles. */
int mt_oracle_step(int model, const double* obs_in, double* obs_out, long B, long N,
                   const oracle_params* p, int n_threads) {
  if (N < 3 || B < 1) return -1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
  int failed = 0;
#pragma omp parallel
  {
    double* scratch = (double*)malloc(sizeof(double) * 3 * (size_t)N);
    if (!scratch) {
#pragma omp atomic write
      failed = 1;
    } else {
#pragma omp for schedule(static)
      for (long b = 0; b < B; ++b) {
        const double* in = obs_in + (size_t)b * 2 * N;
        double* out = obs_out + (size_t)b * 2 * N;
        if (model == 1) arz_row(p, in, out, N, scratch, scratch + N, scratch + 2 * N);
        else lwr_row(p, in, out, N, scratch);
      }
      free(scratch);
    }
  }
  return failed ? -2 : 0;
}

int mt_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
