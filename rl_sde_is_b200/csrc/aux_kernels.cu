// Small kernels around the rollout: deterministic statistics reduction, Philox noise export,
// single-pass environment step (API completeness for env.step / env.step_torch).
#include "aux_kernels.cuh"

namespace rlsde {

// ------------------------------------------------------------------ Adam (device-resident REINFORCE loop)
// torch.optim.Adam's update for one flat parameter vector (amsgrad / weight decay / maximize off), the step the
// reference takes after eff_loss.backward() (reinforce_deterministic_core.py:143-146,243):
//   m <- m + (g - m)(1 - beta1);  v <- v beta2 + (1 - beta2) g g;  theta <- theta - (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// The gradient comes either from this GPU alone (n_ranks == 0: grad float[P], stats double[NSTATS]) or as the rows
// `[grad (P doubles) | stats (NSTATS doubles)]` that the ranks of a data-parallel job exchanged with ONE all-gather
// (n_ranks >= 1: packed double[n_ranks][P + NSTATS]); rows are added in rank order, so every rank forms the same bits.
// A batch in which any trajectory ran into the pass budget (stats[N_UNFINISHED] > 0) is dropped from the gradient by
// the reverse pass; applying the remainder scaled by 1/K would be a biased step, so the update is skipped on the
// device (theta, m, v untouched) and the caller sees the count in the statistics record.
__global__ void adam_step_kernel(int P, int n_ranks, const float* __restrict__ grad, const double* __restrict__ stats,
                                 const double* __restrict__ packed, float* __restrict__ theta, float* __restrict__ m,
                                 float* __restrict__ v, float one_minus_b1, float b2, float one_minus_b2, float eps,
                                 float step_size, float bc2_sqrt, float* __restrict__ grad_out, double* __restrict__ stats_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = P + RLSDE_NSTATS;
  double unfinished = 0.0;
  if (n_ranks > 0) {
    for (int r = 0; r < n_ranks; ++r) unfinished += packed[(long long)r * row + P + RLSDE_ST_N_UNFINISHED];
  } else {
    unfinished = stats[RLSDE_ST_N_UNFINISHED];
  }
  if (n_ranks > 0 && p < RLSDE_NSTATS && stats_out != nullptr) {
    double acc = packed[P + p];
    for (int r = 1; r < n_ranks; ++r) {
      const double x = packed[(long long)r * row + P + p];
      acc = (p == RLSDE_ST_MAX_T) ? (x > acc ? x : acc) : acc + x;
    }
    stats_out[p] = acc;
  }
  if (p >= P) return;
  float g;
  if (n_ranks > 0) {
    double acc = 0.0;
    for (int r = 0; r < n_ranks; ++r) acc += packed[(long long)r * row + p];
    g = (float)acc;
    if (grad_out != nullptr) grad_out[p] = g;
  } else {
    g = grad[p];
  }
  if (unfinished > 0.0) return;
  const float mp = __fmaf_rn(__fsub_rn(g, m[p]), one_minus_b1, m[p]);
  const float vp = __fadd_rn(__fmul_rn(v[p], b2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
  m[p] = mp;
  v[p] = vp;
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vp), bc2_sqrt), eps);
  theta[p] = __fsub_rn(theta[p], __fmul_rn(step_size, __fdiv_rn(mp, denom)));
}

int launch_adam_step(int P, int n_ranks, const float* grad, const double* stats, const double* packed, float* theta, float* m,
                     float* v, double lr, double beta1, double beta2, double eps, long long step_t, float* grad_out,
                     double* stats_out, cudaStream_t stream) {
  const double bc1 = 1.0 - pow(beta1, (double)step_t), bc2 = 1.0 - pow(beta2, (double)step_t);
  const int n = P > RLSDE_NSTATS ? P : RLSDE_NSTATS;
  adam_step_kernel<<<(n + 127) / 128, 128, 0, stream>>>(P, n_ranks, grad, stats, packed, theta, m, v, (float)(1.0 - beta1),
                                                        (float)beta2, (float)(1.0 - beta2), (float)eps, (float)(lr / bc1),
                                                        (float)sqrt(bc2), grad_out, stats_out);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

// packed[0 .. P) = grad (as doubles), packed[P .. P + NSTATS) = stats: this rank's row of the iteration's one exchange
__global__ void pack_grad_stats_kernel(int P, const float* __restrict__ grad, const double* __restrict__ stats,
                                       double* __restrict__ packed) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) packed[p] = (double)grad[p];
  else if (p < P + RLSDE_NSTATS) packed[p] = stats[p - P];
}

int launch_pack_grad_stats(int P, const float* grad, const double* stats, double* packed, cudaStream_t stream) {
  pack_grad_stats_kernel<<<(P + RLSDE_NSTATS + 127) / 128, 128, 0, stream>>>(P, grad, stats, packed);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ statistics
// Two-stage fixed-shape reduction: stage 1 = STATS_BLOCKS blocks, each summing a contiguous slice
// with a fixed in-block tree; stage 2 = one block summing the partials in index order.  The result
// does not depend on scheduling (unlike atomics), so repeated runs are bit-identical.
template <bool F64>
__global__ void __launch_bounds__(256) stats_partial_kernel(long long K, long long n_steps_lim, const void* Gp,
                                                            const void* Sp, const int* T, const void* l2p,
                                                            const void* logwp, double* partial) {
  __shared__ double sh[256];
  double acc[RLSDE_NSTATS];
#pragma unroll
  for (int s = 0; s < RLSDE_NSTATS; ++s) acc[s] = 0.0;
  const long long per = (K + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per;
  const long long hi = lo + per < K ? lo + per : K;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int t = T[i];
    const double g = F64 ? ((const double*)Gp)[i] : (double)((const float*)Gp)[i];
    const double s = F64 ? ((const double*)Sp)[i] : (double)((const float*)Sp)[i];
    acc[RLSDE_ST_N] += 1.0;
    if (t < 0) {
      acc[RLSDE_ST_N_UNFINISHED] += 1.0;
      acc[RLSDE_ST_USEFUL_STEPS] += (double)n_steps_lim;
    } else {
      acc[RLSDE_ST_SUM_G] += g;
      acc[RLSDE_ST_SUM_G2] += g * g;
      acc[RLSDE_ST_SUM_T] += (double)t;
      acc[RLSDE_ST_SUM_T2] += (double)t * (double)t;
      acc[RLSDE_ST_SUM_S] += s;
      if (l2p) acc[RLSDE_ST_SUM_L2] += F64 ? ((const double*)l2p)[i] : (double)((const float*)l2p)[i];
      if (logwp) {
        const double w = exp(F64 ? ((const double*)logwp)[i] : (double)((const float*)logwp)[i]);
        acc[RLSDE_ST_SUM_W] += w;
        acc[RLSDE_ST_SUM_W2] += w * w;
      }
      acc[RLSDE_ST_SUM_LOSS] += -g - g * s;
      acc[RLSDE_ST_USEFUL_STEPS] += (double)t + 1.0;
      acc[RLSDE_ST_MAX_T] = fmax(acc[RLSDE_ST_MAX_T], (double)t);
    }
  }
  for (int s = 0; s < RLSDE_NSTATS; ++s) {
    sh[threadIdx.x] = acc[s];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if ((int)threadIdx.x < w)
        sh[threadIdx.x] = (s == RLSDE_ST_MAX_T) ? fmax(sh[threadIdx.x], sh[threadIdx.x + w]) : sh[threadIdx.x] + sh[threadIdx.x + w];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)blockIdx.x * RLSDE_NSTATS + s] = sh[0];
    __syncthreads();
  }
}

__global__ void stats_final_kernel(int nblocks, const double* partial, double* stats) {
  const int s = threadIdx.x;
  if (s >= RLSDE_NSTATS) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    const double v = partial[(long long)b * RLSDE_NSTATS + s];
    a = (s == RLSDE_ST_MAX_T) ? fmax(a, v) : a + v;
  }
  stats[s] = a;
}

int launch_reduce_stats(long long K, long long n_steps_lim, bool f64, const void* G, const void* S, const int* T,
                        const void* l2, const void* logw, double* stats, double* partial, cudaStream_t stream) {
  int nb = (int)((K + 255) / 256);
  if (nb > STATS_BLOCKS) nb = STATS_BLOCKS;
  if (nb < 1) nb = 1;
  if (f64) stats_partial_kernel<true><<<nb, 256, 0, stream>>>(K, n_steps_lim, G, S, T, l2, logw, partial);
  else stats_partial_kernel<false><<<nb, 256, 0, stream>>>(K, n_steps_lim, G, S, T, l2, logw, partial);
  stats_final_kernel<<<1, 32, 0, stream>>>(nb, partial, stats);
  note_kernel_launches(2);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ noise export
// out[p][k][i] for passes [pass_begin, pass_begin + n_pass): the exact increments the rollout kernels
// consume for (seed, traj_offset + k).  Mapping of (pass, coordinate) to Philox block / slot = NoisePlan.
__global__ void noise_fill_kernel(unsigned long long seed, long long traj_offset, long long K, int d, long long pass_begin,
                                  long long n_pass, float scale2, float* out) {
  const long long total = n_pass * K;
  const int spb = (d == 1) ? 4 : (d == 2 ? 2 : 1);
  const int bpp = (d <= 2) ? 1 : (d + 3) / 4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / K, k = e % K;
    const long long pass = pass_begin + p;
    const unsigned long long gt = (unsigned long long)(traj_offset + k);
    float* dst = out + e * d;
    if (d <= 2) {
      float z[4];
      noise_block(seed, gt, (unsigned)(pass / spb), scale2, z);
      const int sub = (int)(pass % spb);
      for (int i = 0; i < d; ++i) dst[i] = z[sub * d + i];
    } else {
      for (int q = 0; q < bpp; ++q) {
        float z[4];
        noise_block(seed, gt, (unsigned)pass * bpp + q, scale2, z);
        for (int s = 0; s < 4; ++s)
          if (4 * q + s < d) dst[4 * q + s] = z[s];
      }
    }
  }
}

int launch_noise_fill(unsigned long long seed, long long traj_offset, long long K, int d, long long pass_begin,
                      long long n_pass, double dt, float* out, cudaStream_t stream) {
  const long long total = n_pass * K;
  if (total <= 0) return 0;
  long long nb = (total + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  const float scale2 = (float)(-2.0 * dt * 0.6931471805599453);
  noise_fill_kernel<<<(unsigned)nb, 256, 0, stream>>>(seed, traj_offset, K, d, pass_begin, n_pass, scale2, out);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ single environment pass
// env.step (numpy, F64: environments.py:139-162, environments_2d.py:134-153) and env.step_torch
// (f32: environments.py:201-226, environments_2d.py:184-205); same association as the rollout kernel.
template <bool F64>
__global__ void env_step_kernel(StepArgs A) {
  typedef typename RealT<F64>::type real;
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.K) return;
  const int d = A.d;
  const real* x = (const real*)A.state + k * d;
  const float* a = A.action + k * d;
  real* xn = (real*)A.next_state + k * d;
  float dB[RLSDE_MAX_D];
  if (A.dbt_in) {
    for (int i = 0; i < d; ++i) dB[i] = A.dbt_in[k * d + i];
  } else {
    const unsigned long long gt = (unsigned long long)(A.traj_offset + k);
    const int spb = (d == 1) ? 4 : (d == 2 ? 2 : 1);
    const int bpp = (d <= 2) ? 1 : (d + 3) / 4;
    if (d <= 2) {
      float z[4];
      noise_block(A.seed, gt, (unsigned)(A.pass_index / spb), A.noise_scale2, z);
      const int sub = (int)(A.pass_index % spb);
      for (int i = 0; i < d; ++i) dB[i] = z[sub * d + i];
    } else {
      for (int q = 0; q < bpp; ++q) {
        float z[4];
        noise_block(A.seed, gt, (unsigned)A.pass_index * bpp + q, A.noise_scale2, z);
        for (int s = 0; s < 4; ++s)
          if (4 * q + s < d) dB[4 * q + s] = z[s];
      }
    }
  }
  float n2 = 0.f;
  for (int i = 0; i < d; ++i) {
    n2 = (i == 0) ? __fmul_rn(a[i], a[i]) : __fadd_rn(n2, __fmul_rn(a[i], a[i]));
    if (F64) {
      const double xi = (double)x[i];
      // a float32 state handed to the 1-D numpy env gives a float32 gradient (python-float alpha is a weak scalar)
      double g;
      if (A.grad_f32 && d == 1) {
        const float xs = (float)xi;
        g = (double)__fmul_rn(__fmul_rn(A.c4a_f[i], xs), __fsub_rn(__fmul_rn(xs, xs), 1.0f));
      } else if (A.grad_f32) {   // d-D env: 4 * alpha (f64 array) * state (f32) * (state**2 - 1) (f32)
        const float xs = (float)xi;
        g = __dmul_rn(__dmul_rn(A.c4a_d[i], xi), (double)__fsub_rn(__fmul_rn(xs, xs), 1.0f));
      } else {
        g = __dmul_rn(__dmul_rn(A.c4a_d[i], xi), __dsub_rn(__dmul_rn(xi, xi), 1.0));
      }
      const double drift = __dmul_rn(__dadd_rn(-g, __dmul_rn(A.sigma_d, (double)a[i])), A.dt_d);
      xn[i] = (real)__dadd_rn(__dadd_rn(xi, drift), __dmul_rn(A.sigma_d, (double)dB[i]));
    } else {
      const float xi = (float)x[i];
      const float g = __fmul_rn(__fmul_rn(A.c4a_f[i], xi), __fsub_rn(__fmul_rn(xi, xi), 1.0f));
      const float drift = __fmul_rn(__fadd_rn(-g, __fmul_rn(A.sigma_f, a[i])), A.dt_f);
      xn[i] = (real)__fadd_rn(__fadd_rn(xi, drift), __fmul_rn(A.sigma_f, dB[i]));
    }
    if (A.dbt_out) A.dbt_out[k * d + i] = dB[i];
  }
  const real* xt = (A.reward_type == RLSDE_REWARD_STATE_ACTION) ? x : xn;
  bool done;
  if (A.hit_rule == RLSDE_HIT_X0_IN_LB_RB) {
    done = F64 ? ((double)xt[0] >= A.lb_d && (double)xt[0] <= A.rb_d) : ((float)xt[0] >= A.lb_f && (float)xt[0] <= A.rb_f);
  } else {
    done = true;
    for (int i = 0; i < d; ++i) done = done && (F64 ? ((double)xt[i] >= A.lb_d) : ((float)xt[i] >= A.lb_f));
  }
  const float nn = (d == 1) ? n2 : __fmul_rn(sqrtf(n2), sqrtf(n2));
  real run;
  if (F64) run = (real)(-__dmul_rn(__dadd_rn(1.0, (double)__fmul_rn(0.5f, nn)), A.dt_d));
  else run = (real)(-__fmul_rn(__fadd_rn(1.0f, __fmul_rn(0.5f, nn)), A.dt_f));
  // 'state-action': done ? -g(x) = -0 : running cost;  'state-action-next-state': running cost (- g(x') = 0) either way
  real r = run;
  if (A.reward_type == RLSDE_REWARD_STATE_ACTION && done) r = (real)(-0.0);
  ((real*)A.reward)[k] = r;
  A.done[k] = done ? 1 : 0;
}

int launch_env_step(const StepArgs& A, bool f64, cudaStream_t stream) {
  if (A.K <= 0) return 0;
  const unsigned nb = (unsigned)((A.K + 127) / 128);
  if (f64) env_step_kernel<true><<<nb, 128, 0, stream>>>(A);
  else env_step_kernel<false><<<nb, 128, 0, stream>>>(A);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

}  // namespace rlsde
