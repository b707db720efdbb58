/*
 * rlsde_oracle.c -- plain-C CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs through oracle/c_oracle.py; never by
 * the product package.
 *
 * What it restates (citations into /root/reference/src/rl_sde_is/):
 *   - one Euler-Maruyama pass + hit test + running cost        environments.py:139-162, 201-226
 *   - the rollout loops with first-hitting-time bookkeeping     reinforce_deterministic_core.py:30-93,
 *                                                               approximate_methods.py:577-695
 *   - the policy MLP                                            models.py:4-18
 *   - the transition tensor / reward table                      dynamic_programming.py:3-36, environments.py:87-102
 * and, in addition, the counter-based noise stream of the CUDA kernels (Philox4x32-10 + Box-Muller,
 * csrc/common.cuh) so that rollouts with IN-KERNEL random numbers can be checked per trajectory.
 *
 * Pinning: the Python restatement (oracle/reference_semantics.py) is checked bit-for-bit against
 * fixtures recorded from the unmodified reference; this file is checked against that restatement on
 * injected noise (tests/test_oracle_golden.py), and its Philox against the Random123 known-answer
 * vectors.  Arithmetic of a pass uses the reference's association with contraction disabled
 * (-ffp-contract=off); the GEMV uses explicit fmaf chains in the kernels' order.
 *
 * Build: gcc -O2 -std=c11 -fPIC -shared -fopenmp -ffp-contract=off -o librlsde_oracle.so rlsde_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_D 16
#define MAX_H 256

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define PHILOX_TAG 0x52534445u

/* 4 increments of noise block b of trajectory traj: dB = sqrt(dt) N(0,1); uniforms exactly as the kernel
   forms them (float fma), transcendental part in double (the kernel uses MUFU approximations) */
static void noise_block(uint64_t seed, uint64_t traj, uint32_t b, double dt, float z[4]) {
  const uint32_t ctr[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), b, PHILOX_TAG};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  oracle_philox4x32_10(ctr, key, r);
  float u[4];
  for (int i = 0; i < 4; ++i) u[i] = fmaf((float)r[i], 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const double ra = sqrt(-2.0 * dt * log((double)u[0])), rb = sqrt(-2.0 * dt * log((double)u[2]));
  const double ta = 6.283185307179586 * (double)u[1], tb = 6.283185307179586 * (double)u[3];
  z[0] = (float)(ra * cos(ta)); z[1] = (float)(ra * sin(ta));
  z[2] = (float)(rb * cos(tb)); z[3] = (float)(rb * sin(tb));
}

/* increment for (trajectory, pass, coordinate): same block / slot mapping as NoisePlan in rollout_fwd.cuh */
static void noise_pass(uint64_t seed, uint64_t traj, int64_t pass, int d, double dt, float* dB) {
  if (d <= 2) {
    const int spb = 4 / d;
    float z[4];
    noise_block(seed, traj, (uint32_t)(pass / spb), dt, z);
    for (int i = 0; i < d; ++i) dB[i] = z[(pass % spb) * d + i];
  } else {
    const int bpp = (d + 3) / 4;
    for (int q = 0; q < bpp; ++q) {
      float z[4];
      noise_block(seed, traj, (uint32_t)pass * bpp + q, dt, z);
      for (int s = 0; s < 4; ++s)
        if (4 * q + s < d) dB[4 * q + s] = z[s];
    }
  }
}

void oracle_noise_fill(uint64_t seed, int64_t traj_offset, int64_t K, int d, int64_t pass_begin, int64_t n_pass,
                       double dt, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < K; ++k)
    for (int64_t p = 0; p < n_pass; ++p)
      noise_pass(seed, (uint64_t)(traj_offset + k), pass_begin + p, d, dt, out + (p * K + k) * d);
}

/* ------------------------------------------------------------------ policy (state_dict order, float32) */
typedef struct { int d, H; const float *W1, *b1, *W2, *b2, *W3, *b3; } mlp_t;

static mlp_t mlp_view(const float* p, int d, int H) {
  mlp_t m; m.d = d; m.H = H;
  m.W1 = p; m.b1 = m.W1 + H * d; m.W2 = m.b1 + H; m.b2 = m.W2 + H * H; m.W3 = m.b2 + H; m.b3 = m.W3 + d * H;
  return m;
}

static void mlp_forward(const mlp_t* m, const float* x, float* u) {
  float h1[MAX_H], h2[MAX_H];
  const int d = m->d, H = m->H;
  for (int j = 0; j < H; ++j) {
    float a = m->b1[j];
    for (int i = 0; i < d; ++i) a = fmaf(x[i], m->W1[j * d + i], a);
    h1[j] = tanhf(a);
  }
  for (int j = 0; j < H; ++j) {
    float a = m->b2[j];
    for (int i = 0; i < H; ++i) a = fmaf(h1[i], m->W2[j * H + i], a);
    h2[j] = tanhf(a);
  }
  for (int k = 0; k < d; ++k) {
    float a = m->b3[k];
    for (int j = 0; j < H; ++j) a = fmaf(h2[j], m->W3[k * H + j], a);
    u[k] = a;
  }
}

/* ------------------------------------------------------------------ rollouts */
/* flags */
#define OF_NOISE_INJECTED 1
#define OF_STOCH_INT_EXACT 4
#define OF_STATE_F64 8
#define HIT_ALL_GE_LB 0
#define HIT_X0_IN_LB_RB 1

/* One trajectory, float32 torch semantics.  Returns k* or -1. */
static int64_t rollout_one_f32(const mlp_t* m, const double* alpha, double sigma, double dt, double lb, double rb,
                               int hit_rule, const double* x0, uint64_t seed, uint64_t gtraj, const float* noise,
                               int64_t K_global, int64_t lim, int flags, float* G_out, float* S_out, float* logw_out) {
  const int d = m->d;
  const float sig = (float)sigma, dtf = (float)dt, lbf = (float)lb, rbf = (float)rb;
  float x[MAX_D], u[MAX_D], dB[MAX_D], c4a[MAX_D];
  for (int i = 0; i < d; ++i) { x[i] = (float)x0[i]; c4a[i] = (float)(4.0 * alpha[i]); }
  float G = 0.f, S = 0.f;
  for (int64_t k = 0; k < lim; ++k) {
    mlp_forward(m, x, u);
    if (flags & OF_NOISE_INJECTED) for (int i = 0; i < d; ++i) dB[i] = noise[(k * K_global + (int64_t)gtraj) * d + i];
    else noise_pass(seed, gtraj, k, d, dt, dB);
    int hit;
    if (hit_rule == HIT_X0_IN_LB_RB) hit = x[0] >= lbf && x[0] <= rbf;
    else { hit = 1; for (int i = 0; i < d; ++i) hit = hit && (x[i] >= lbf); }
    float su = u[0] * dB[0], n2 = u[0] * u[0];
    for (int i = 1; i < d; ++i) { su = su + u[i] * dB[i]; n2 = n2 + u[i] * u[i]; }
    const float S_prev = S;
    S = S + su;
    if (hit) {
      *G_out = G; *S_out = (flags & OF_STOCH_INT_EXACT) ? S_prev : S; *logw_out = G - S_prev;
      return k;
    }
    const float nn = (d == 1) ? n2 : sqrtf(n2) * sqrtf(n2);
    G = G + (-((1.0f + 0.5f * nn) * dtf));
    for (int i = 0; i < d; ++i) {
      const float g = (c4a[i] * x[i]) * (x[i] * x[i] - 1.0f);
      const float drift = (-g + sig * u[i]) * dtf;
      x[i] = (x[i] + drift) + sig * dB[i];
    }
  }
  *G_out = G; *S_out = S; *logw_out = G - S;
  return -1;
}

/* One trajectory, numpy semantics: float64 state and accumulators, float32 policy. */
static int64_t rollout_one_f64(const mlp_t* m, const double* alpha, double sigma, double dt, double lb, double rb,
                               int hit_rule, const double* x0, uint64_t seed, uint64_t gtraj, const float* noise,
                               int64_t K_global, int64_t lim, int flags, const float* policy_opt, int64_t n_grid,
                               double grid_lo, double grid_hi, double grid_h, double* G_out, double* S_out,
                               double* l2_out, double* logw_out) {
  const int d = m->d;
  double x[MAX_D];
  float xf[MAX_D], u[MAX_D], dB[MAX_D];
  for (int i = 0; i < d; ++i) x[i] = (double)(float)x0[i];
  double G = 0.0, S = 0.0, L2 = 0.0;
  for (int64_t k = 0; k < lim; ++k) {
    for (int i = 0; i < d; ++i) xf[i] = (float)x[i];
    mlp_forward(m, xf, u);
    if (flags & OF_NOISE_INJECTED) for (int i = 0; i < d; ++i) dB[i] = noise[(k * K_global + (int64_t)gtraj) * d + i];
    else noise_pass(seed, gtraj, k, d, dt, dB);
    int hit;
    if (hit_rule == HIT_X0_IN_LB_RB) hit = x[0] >= lb && x[0] <= rb;
    else { hit = 1; for (int i = 0; i < d; ++i) hit = hit && (x[i] >= lb); }
    double su = (double)u[0] * (double)dB[0];
    float n2 = u[0] * u[0];
    for (int i = 1; i < d; ++i) { su = su + (double)u[i] * (double)dB[i]; n2 = n2 + u[i] * u[i]; }
    const double S_prev = S;
    S = S + su;
    if (policy_opt) {
      double xc = x[0] < grid_lo ? grid_lo : (x[0] > grid_hi ? grid_hi : x[0]);
      int64_t gi = (int64_t)floor((xc - grid_lo) / grid_h);
      if (gi < 0) gi = 0;
      if (gi >= n_grid) gi = n_grid - 1;
      const float du = u[0] - policy_opt[gi];
      L2 = L2 + (double)(du * du) * dt;
    }
    if (hit) {
      *G_out = G; *S_out = (flags & OF_STOCH_INT_EXACT) ? S_prev : S; *l2_out = L2; *logw_out = G - S_prev;
      return k;
    }
    const float nn = (d == 1) ? n2 : sqrtf(n2) * sqrtf(n2);
    G = G + (-((1.0 + (double)(0.5f * nn)) * dt));
    for (int i = 0; i < d; ++i) {
      double g;
      if (d == 1 && k == 0) {   /* float32 state + python-float alpha on the first pass of the 1-D env */
        const float xs = (float)x[i], c = (float)(4.0 * alpha[i]);
        g = (double)((c * xs) * (xs * xs - 1.0f));
      } else if (k == 0) {      /* d-D: float64 alpha array, state**2 - 1 still float32 on the first pass */
        const float xs = (float)x[i];
        g = ((4.0 * alpha[i]) * x[i]) * (double)(xs * xs - 1.0f);
      } else {
        g = ((4.0 * alpha[i]) * x[i]) * (x[i] * x[i] - 1.0);
      }
      const double drift = (-g + sigma * (double)u[i]) * dt;
      x[i] = (x[i] + drift) + sigma * (double)dB[i];
    }
  }
  *G_out = G; *S_out = S; *l2_out = L2; *logw_out = G - S;
  return -1;
}

/*
 * K trajectories [traj_offset, traj_offset + K).  Outputs: G, S, l2, logw are float (double with
 * OF_STATE_F64), T is int32 (k* or -1).  Returns the number of useful passes (sum of k*+1, or lim if
 * unfinished) -- the throughput metric of SURVEY 8d.
 */
int64_t oracle_rollout(int d, int H, const float* params, const double* alpha, double sigma, double dt, double lb,
                       double rb, int hit_rule, const double* x0, int64_t K, int64_t traj_offset, int64_t K_global,
                       uint64_t seed, int64_t n_steps_lim, int64_t noise_steps, int flags, const float* noise,
                       const float* policy_opt, int64_t n_grid, double grid_lo, double grid_hi, double grid_h, void* G,
                       void* S, int32_t* T, void* l2, void* logw) {
  if (d < 1 || d > MAX_D || H < 1 || H > MAX_H) return -1;
  const mlp_t m = mlp_view(params, d, H);
  const int64_t lim = (flags & OF_NOISE_INJECTED) && noise_steps < n_steps_lim ? noise_steps : n_steps_lim;
  int64_t useful = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : useful)
  for (int64_t k = 0; k < K; ++k) {
    const uint64_t gt = (uint64_t)(traj_offset + k);
    int64_t t;
    if (flags & OF_STATE_F64) {
      double g, s, e, w;
      t = rollout_one_f64(&m, alpha, sigma, dt, lb, rb, hit_rule, x0, seed, gt, noise, K_global, lim, flags, policy_opt,
                          n_grid, grid_lo, grid_hi, grid_h, &g, &s, &e, &w);
      ((double*)G)[k] = g; ((double*)S)[k] = s;
      if (l2) ((double*)l2)[k] = e;
      if (logw) ((double*)logw)[k] = w;
    } else {
      float g, s, w;
      t = rollout_one_f32(&m, alpha, sigma, dt, lb, rb, hit_rule, x0, seed, gt, noise, K_global, lim, flags, &g, &s, &w);
      ((float*)G)[k] = g; ((float*)S)[k] = s;
      if (logw) ((float*)logw)[k] = w;
    }
    T[k] = (int32_t)t;
    useful += t >= 0 ? t + 1 : lim;
  }
  return useful;
}

/* ------------------------------------------------------------------ tables */
static double ndtr_c(double a) {   /* cephes ndtr.c branch structure (what scipy.stats.norm.cdf evaluates) */
  const double x = a * 0.70710678118654752440, z = fabs(x);
  double y;
  if (z < 0.70710678118654752440) y = 0.5 + 0.5 * erf(x);
  else { y = 0.5 * erfc(z); if (x > 0) y = 1.0 - y; }
  return y;
}

/* P[s', s, a] for s' in [sp_begin, sp_end) and R[s, a]; per-cell CDF differences exactly as
   environments.py:94-101 writes them (no sharing of edges between neighbouring cells) */
void oracle_tables(const double* sgrid, int64_t Ns, const double* agrid, int64_t Na, const uint8_t* in_ts, int64_t n_ts,
                   double alpha, double sigma, double dt, double h, int64_t sp_begin, int64_t sp_end, double* P, double* R) {
  const double sd = sigma * sqrt(dt);
  if (P) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < Ns; ++s) {
      for (int64_t a = 0; a < Na; ++a) {
        const double xs = sgrid[s];
        const double mu = xs + (-(4 * alpha * xs * (xs * xs - 1)) + sigma * agrid[a]) * dt;
        for (int64_t sp = sp_begin; sp < sp_end; ++sp) {
          double p;
          if (in_ts[s]) p = in_ts[sp] ? 1.0 / (double)n_ts : 0.0;
          else {
            const double up = ndtr_c((sgrid[sp] + h - mu) / sd), lo = ndtr_c((sgrid[sp] - h - mu) / sd);
            p = up - lo;
            if (sp == 0) p += ndtr_c((sgrid[0] - h - mu) / sd);
            if (sp == Ns - 1) p += 1 - ndtr_c((sgrid[Ns - 1] + h - mu) / sd);
          }
          P[((sp - sp_begin) * Ns + s) * Na + a] = p;
        }
      }
    }
  }
  if (R)
    for (int64_t s = 0; s < Ns; ++s)
      for (int64_t a = 0; a < Na; ++a) R[s * Na + a] = in_ts[s] ? -0.0 : -((1.0 + 0.5 * (agrid[a] * agrid[a])) * dt);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
