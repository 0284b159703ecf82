/* CPU ORACLE in plain C for the abdpymc inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__ (build / smoke) and bench.py's CPU-baseline leg may compile, load or call this
 * file; the product path (abdpymc_b200/) never does and has no CPU fallback.
 *
 * It is a second, independent restatement (beside oracle/abd_oracle.py) of the data log-likelihood of the
 * reference's model and of its gradient w.r.t. the 13 constrained parameters, in the O(G N) recurrence form
 * the reference itself proves equal to its dense (G,G,N) form (abd.py:277-293 `_temp_response_scan`,
 * test_abd.py:987-1011).  All `abd.py:NNN` citations are relative to /root/reference/abdpymc/.  It serves as
 *   - a checker of the NumPy oracle (tests/test_oracle.py: both against the goldens produced by executing
 *     the reference's own code, tests/golden/model_goldens.npz), and
 *   - the algorithmically fair, compiled, multi-threaded CPU implementation timed beside the GPU number
 *     (bench.py cpu_baseline.c_port): what a careful CPU port of this path achieves on the box's host cores.
 * Parity pinning: through the NumPy oracle's pins (see its header) -- this file is compared with the same
 * golden vectors.
 *
 * Layout: matrices over (gap, individual) are (G, N) with the individual axis contiguous, as the reference
 * graph has them after abd.py:413-418.
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -o oracle/_build/libabd_oracle_c.so oracle/abd_oracle_c.c -lm
 * (oracle/c_oracle.py does it).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* theta13 order of the C ABI / oracle.THETA13 */
enum { N_PERM = 0, N_TEMP, N_RHO, N_INIT, S_PERM, S_RHO, S_INIT, N_B, N_D, N_SIGMA, S_B, S_D, S_SIGMA };

/* i_raw column -> constrained infections of one individual (strides between gaps: raw_stride, N for pcr).
 *   no splits: i = mask_three_gaps((i_raw + pcrpos) > 0)                                  abd.py:640-649
 *   splits   : per time chunk keep the first raw 1 (mask_multiple_infections, abd.py:792-862: entries where
 *              the running count exceeds 1 become 0), replace the chunk by the PCR+ chunk if that has any 1
 *              (incorporate_pcrpos abd.py:732-771, per chunk :691-697, :722-729), then mask_three_gaps.
 *   mask_three_gaps (abd.py:560-601): out[t] = 0 if out[t-1] | out[t-2] | out[t-3] else in[t]  (taps on OUTPUTS). */
static void constrain_column(int G, long raw_stride, const int8_t* raw, long N, const int8_t* pcr, int n_splits,
                             const int32_t* splits, int8_t* out) {
  int8_t m[256];
  if (n_splits == 0) {
    for (int t = 0; t < G; ++t) m[t] = (raw[t * raw_stride] + pcr[t * N]) > 0;
  } else {
    int lo = 0;
    for (int k = 0; k <= n_splits; ++k) {
      const int hi = (k < n_splits) ? splits[k] : G;
      int count = 0, any_pcr = 0;
      for (int t = lo; t < hi; ++t) any_pcr |= pcr[t * N] != 0;
      for (int t = lo; t < hi; ++t) {
        count += raw[t * raw_stride] != 0;
        const int8_t kept = (count > 1) ? 0 : (raw[t * raw_stride] != 0);
        m[t] = any_pcr ? (pcr[t * N] != 0) : kept;
      }
      lo = hi;
    }
  }
  for (int t = 0; t < G; ++t) {
    const int prev = (t >= 1 && out[t - 1]) || (t >= 2 && out[t - 2]) || (t >= 3 && out[t - 3]);
    out[t] = prev ? 0 : m[t];
  }
}

/* out[0] = loglik, out[1..13] = d loglik / d theta13.  scratch: 7 * G * N doubles.
 * rows of antigen a (0 = N, 1 = S): x[a], od[a], gap[a], ind[a], R[a] of them, any order.
 * Returns 0, or 1 for invalid sizes. */
int abd_c_loglik_grad(int G, long N, int n_splits, const int32_t* splits, const int8_t* pcr, const int8_t* vac,
                      long Rn, const double* xn, const double* odn, const int32_t* gapn, const int32_t* indn,
                      long Rs, const double* xs, const double* ods, const int32_t* gaps, const int32_t* inds,
                      const double* th, const int8_t* i_raw, const int8_t* waner, double* out, double* scratch,
                      int n_threads) {
  if (G < 1 || G > 256 || N < 1 || n_splits < 0 || n_splits > 2) return 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  const long GN = (long)G * N;
  double* mu_n = scratch;           /* ab_n_mu                      abd.py:341 */
  double* mu_s = scratch + GN;      /* ab_s_mu                      abd.py:389-391 */
  double* P_n = scratch + 2 * GN;   /* 1[any infection so far]      abd.py:296-306, :330 */
  double* P_s = scratch + 3 * GN;   /* 1[any exposure so far]       abd.py:368 */
  double* T_n = scratch + 4 * GN;   /* sum_s rho^(t-s) i_s          abd.py:242-260 */
  double* dT_n = scratch + 5 * GN;  /* its rho-derivative */
  double* dU_s = scratch + 6 * GN;  /* d U / d rho_s (0 for non-waners: rho_ind = 1, abd.py:374) */
  const double rho_n = th[N_RHO], rho_s = th[S_RHO];

  /* ---- per individual: constraints, then the two recurrences over the gaps ---- */
#pragma omp parallel for schedule(static)
  for (long n = 0; n < N; ++n) {
    int8_t inf[256];
    constrain_column(G, N, i_raw + n, N, pcr + n, n_splits, splits, inf);
    const int w = waner[n] != 0;
    const double rho_ind = w ? rho_s : 1.0; /* rho * waner + 1 - waner, abd.py:374 */
    double T = 0.0, dT = 0.0, U = 0.0, dU = 0.0;
    int any_i = 0, any_e = 0;
    for (int t = 0; t < G; ++t) {
      const double it = (double)inf[t];
      const double et = it + (double)vac[t * N + n]; /* exposure = i + v: a month with both counts twice, abd.py:378-386 */
      dT = rho_n * dT + T;                            /* T'_t = rho T'_{t-1} + T_{t-1} */
      T = rho_n * T + it;                             /* T_t = rho T_{t-1} + i_t       (abd.py:277-293) */
      dU = rho_ind * dU + U;
      U = rho_ind * U + et;                           /* no `temp` factor on the S antigen (abd.py:263-274 ignores it) */
      any_i |= inf[t] != 0;
      any_e |= et > 0.0;
      const long k = (long)t * N + n;
      P_n[k] = (double)any_i;
      P_s[k] = (double)any_e;
      T_n[k] = T;
      dT_n[k] = dT;
      dU_s[k] = w ? dU : 0.0;
      mu_n[k] = th[N_PERM] * (double)any_i + th[N_TEMP] * T + th[N_INIT];
      mu_s[k] = th[S_PERM] * (double)any_e + U + th[S_INIT];
    }
  }

  /* ---- per OD row: logistic curve, Normal likelihood, gradient (abd.py:445-469, :556-557) ---- */
  double acc[14];
  for (int k = 0; k < 14; ++k) acc[k] = 0.0;
  const double half_log_2pi = 0.918938533204672741780329736406;
  for (int a = 0; a < 2; ++a) {
    const long R = a ? Rs : Rn;
    const double *x = a ? xs : xn, *od = a ? ods : odn;
    const int32_t *gap = a ? gaps : gapn, *ind = a ? inds : indn;
    const double* mu = a ? mu_s : mu_n;
    const double* P = a ? P_s : P_n;
    const double b = th[a ? S_B : N_B], d = th[a ? S_D : N_D], sg = th[a ? S_SIGMA : N_SIGMA];
    const double log_sg = log(sg), inv_sg = 1.0 / sg;
    double ll = 0.0, g_d = 0.0, g_b = 0.0, g_sg = 0.0, g_init = 0.0, g_perm = 0.0, g_temp = 0.0, g_rho = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ll, g_d, g_b, g_sg, g_init, g_perm, g_temp, g_rho)
    for (long r = 0; r < R; ++r) {
      const long k = (long)gap[r] * N + ind[r];
      const double m = mu[k];
      const double xm = x[r] - m;
      const double s = 1.0 / (1.0 + exp(-b * xm));     /* logistic, abd.py:556-557 */
      const double eps = (od[r] - d * s) * inv_sg;     /* (od - pred) / sigma */
      const double es = eps * inv_sg;
      ll += -0.5 * eps * eps - half_log_2pi - log_sg;  /* Normal(pred, sigma), abd.py:459-469 */
      g_d += es * s;
      const double q = es * d * s * (1.0 - s);
      g_b += q * xm;
      g_sg += (eps * eps - 1.0) * inv_sg;
      const double dm = -q * b;                        /* d l / d mu */
      g_init += dm;
      g_perm += dm * P[k];
      if (a == 0) {
        g_temp += dm * T_n[k];
        g_rho += dm * th[N_TEMP] * dT_n[k];
      } else {
        g_rho += dm * dU_s[k];
      }
    }
    acc[0] += ll;
    if (a == 0) {
      acc[1 + N_D] = g_d, acc[1 + N_B] = g_b, acc[1 + N_SIGMA] = g_sg;
      acc[1 + N_INIT] = g_init, acc[1 + N_PERM] = g_perm, acc[1 + N_TEMP] = g_temp, acc[1 + N_RHO] = g_rho;
    } else {
      acc[1 + S_D] = g_d, acc[1 + S_B] = g_b, acc[1 + S_SIGMA] = g_sg;
      acc[1 + S_INIT] = g_init, acc[1 + S_PERM] = g_perm, acc[1 + S_RHO] = g_rho;
    }
  }
  memcpy(out, acc, sizeof acc);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * The Gibbs sweep over (i_raw, ab_s_waner), restated as the CUDA kernel schedules it (include/abd_b200.h
 * abd_gibbs_sweep; oracle/abd_oracle.py device_gibbs_sweep is the same in NumPy): per individual n the proposals
 * j = 0..G (j < G flips i_raw[j, n], j == G flips waner[n]) are visited in the order of their Philox keys;
 * mode 0 = PyMC's BinaryGibbsMetropolis rule (flip proposed with probability transit_p, accepted iff
 * log U < logp_prop - logp_curr; pm.sample's default step for binary variables, abd.py:922), mode 1 = heat bath.
 * Only the flipped bit's individual enters the difference (every other term of the model cancels exactly).
 * Random numbers: Philox4x32-10, counter (j, GLOBAL individual, chain, sweep low word), key (seed low word,
 * seed high word ^ sweep high word); word 0 = sort key, word 1 = transit uniform, word 2 = accept uniform.
 * ------------------------------------------------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
static double u01(uint32_t r) { return ((double)r + 1.0) * 2.3283064365386963e-10; } /* (0, 1] */

typedef struct {
  int G;
  long N;
  int n_splits;
  const int32_t* splits;
  const int8_t *pcr, *vac;                 /* (G, N) */
  const int64_t* ptr[2];                   /* rows of individual n: [ptr[a][n], ptr[a][n + 1]) (rows sorted by individual) */
  const double *x[2], *od[2];
  const int32_t* gap[2];
  const double* th;
} Model;

/* data log-likelihood of individual n's OD rows for the raw column `col` (contiguous) and waner bit wn */
static double loglik_individual(const Model* M, long n, const int8_t* col, int wn) {
  const int G = M->G;
  const long N = M->N;
  const double* th = M->th;
  int8_t inf[256];
  double mu_n[256], mu_s[256];
  constrain_column(G, 1, col, N, M->pcr + n, M->n_splits, M->splits, inf);
  const double rho_n = th[N_RHO], rho_ind = wn ? th[S_RHO] : 1.0;
  double T = 0.0, U = 0.0;
  int any_i = 0, any_e = 0;
  for (int t = 0; t < G; ++t) {
    const double it = (double)inf[t], et = it + (double)M->vac[t * N + n];
    T = rho_n * T + it;
    U = rho_ind * U + et;
    any_i |= inf[t] != 0;
    any_e |= et > 0.0;
    mu_n[t] = th[N_PERM] * (double)any_i + th[N_TEMP] * T + th[N_INIT];
    mu_s[t] = th[S_PERM] * (double)any_e + U + th[S_INIT];
  }
  const double half_log_2pi = 0.918938533204672741780329736406;
  double total = 0.0;
  for (int a = 0; a < 2; ++a) {
    const double b = th[a ? S_B : N_B], d = th[a ? S_D : N_D], sg = th[a ? S_SIGMA : N_SIGMA];
    const double log_sg = log(sg);
    const double* mu = a ? mu_s : mu_n;
    double ll = 0.0;
    for (int64_t r = M->ptr[a][n]; r < M->ptr[a][n + 1]; ++r) {
      const double pred = d / (1.0 + exp(-b * (M->x[a][r] - mu[M->gap[a][r]])));
      const double z = (M->od[a][r] - pred) / sg;
      ll += -0.5 * z * z - half_log_2pi - log_sg;
    }
    total += ll;
  }
  return total;
}

/* One sweep of one chain, in place.  stats2 = {proposals, flips}.  Returns 0, or 1 for invalid sizes. */
int abd_c_gibbs_sweep(int G, long N, int n_splits, const int32_t* splits, const int8_t* pcr, const int8_t* vac,
                      const int64_t* ptr_n, const double* xn, const double* odn, const int32_t* gapn,
                      const int64_t* ptr_s, const double* xs, const double* ods, const int32_t* gaps,
                      const double* th, double p, double p_w, int8_t* i_raw, int8_t* waner, uint64_t seed,
                      uint64_t sweep, uint32_t chain, int mode, double transit_p, long ind_offset, int64_t* stats2,
                      int n_threads) {
  if (G < 1 || G > 255 || N < 1 || n_splits < 0 || n_splits > 2 || mode < 0 || mode > 1) return 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  const Model M = {G, N, n_splits, splits, pcr, vac, {ptr_n, ptr_s}, {xn, xs}, {odn, ods}, {gapn, gaps}, th};
  const double lo_i = log(p) - log1p(-p), lo_w = log(p_w) - log1p(-p_w);
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(sweep >> 32)};
  long long n_prop = 0, n_flip = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n_prop, n_flip)
  for (long n = 0; n < N; ++n) {
    int8_t col[256], col2[256];
    for (int t = 0; t < G; ++t) col[t] = i_raw[t * N + n] != 0;
    int wn = waner[n] != 0;
    double cur = loglik_individual(&M, n, col, wn);
    uint32_t rnd[256][4];
    int order[256];
    for (int j = 0; j <= G; ++j) {
      const uint32_t ctr[4] = {(uint32_t)j, (uint32_t)(n + ind_offset), chain, (uint32_t)sweep};
      philox4x32_10(ctr, key, rnd[j]);
      order[j] = j;
    }
    for (int a = 1; a <= G; ++a) { /* insertion sort by (key, j): stable on ties */
      const int v = order[a];
      int b = a - 1;
      while (b >= 0 && rnd[order[b]][0] > rnd[v][0]) {
        order[b + 1] = order[b];
        --b;
      }
      order[b + 1] = v;
    }
    for (int q = 0; q <= G; ++q) {
      const int j = order[q];
      const double u_t = u01(rnd[j][1]), u_a = u01(rnd[j][2]);
      if (mode == 0 && !(u_t <= transit_p)) continue;
      memcpy(col2, col, (size_t)G);
      int wn2 = wn, cur_bit;
      double lo;
      if (j == G) {
        wn2 = 1 - wn;
        cur_bit = wn;
        lo = lo_w;
      } else {
        col2[j] = 1 - col[j];
        cur_bit = col[j];
        lo = lo_i;
      }
      const double nw = loglik_individual(&M, n, col2, wn2);
      const double d10 = cur_bit ? (cur - nw + lo) : (nw - cur + lo);
      int flip;
      if (mode == 0) {
        const double delta = cur_bit ? -d10 : d10;
        flip = isfinite(delta) && log(u_a) < delta;
      } else {
        flip = (u_a <= 1.0 / (1.0 + exp(-d10))) != (cur_bit != 0);
      }
      ++n_prop;
      if (flip) {
        ++n_flip;
        memcpy(col, col2, (size_t)G);
        wn = wn2;
        cur = nw;
      }
    }
    for (int t = 0; t < G; ++t) i_raw[t * N + n] = col[t];
    waner[n] = (int8_t)wn;
  }
  stats2[0] = n_prop;
  stats2[1] = n_flip;
  return 0;
}

int abd_c_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
