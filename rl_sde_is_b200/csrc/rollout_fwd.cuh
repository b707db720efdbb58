// K1: fused forward rollout (policy MLP + Euler-Maruyama pass + hit bookkeeping + running
// work / stochastic-integral accumulation), one trajectory per thread, state in registers for
// the whole rollout, lane refill from a global work counter.
//
// Replaces the per-pass Python loops of sample_loss_vectorized (reinforce_deterministic_core.py:52-88),
// test_policy_vectorized (approximate_methods.py:597-638) and estimate_fht_vectorized (:667-688),
// i.e. one model.forward + env.step[_torch] (environments.py:139-162,201-226) per pass.
//
// Per-pass semantics:
//   u = policy(X_k);  hit = X_k in target set (CURRENT state);  S += u . dB_{k+1} (also on the hit pass);
//   hit  -> record (G, S, k) and retire;   else G += -(1 + |u|^2/2) dt;  X_{k+1} = X_k + (-gradV + sigma u) dt + sigma dB.
// Arithmetic of the pass is written with the reference's association and without FMA contraction
// so that, given the same action and increment, next state / reward are bit-identical to torch's
// (f32) or numpy's (f64, RLSDE_F_STATE_F64) results.
//
// Scheduling: a warp's 32 lanes run 32 independent trajectories in lock step.  A lane whose
// trajectory retires takes the next global trajectory id (warp-aggregated atomicAdd) at the next
// noise-block boundary, so lanes stay busy until the work runs out; the warp leaves the loop when a
// ballot shows no live lane (first-hitting-time divergence only costs the tail).
#pragma once
#include "common.cuh"
#include "../../include/rlsde.h"

namespace rlsde {

template <bool F64> struct RealT { typedef float type; };
template <> struct RealT<true> { typedef double type; };

struct FwdArgs {
  // environment (both precisions are precomputed on the host the way torch / numpy round them)
  float c4a_f[RLSDE_MAX_D];     // float32(4 * alpha_i)
  double c4a_d[RLSDE_MAX_D];    // 4 * alpha_i
  float x0_f[RLSDE_MAX_D];
  double x0_d[RLSDE_MAX_D];
  float sigma_f, dt_f, lb_f, rb_f;
  double sigma_d, dt_d, lb_d, rb_d;
  float noise_scale2;           // -2 dt ln 2
  int hit_rule;
  // work
  long long K, traj_offset, K_global;
  unsigned long long seed;
  long long n_steps_lim, noise_steps;
  unsigned flags;
  int ckpt_every;
  long long ckpt_stride;
  long long n_grid;
  double grid_lo, grid_hi, grid_h;
  // buffers
  const float* noise;
  const float* policy_opt;
  void* G;
  void* S;
  int* T;
  void* l2;
  void* logw;
  float* path;
  unsigned long long* counter;   // work counter of THIS launch (zeroed by the host wrapper)
  const long long* order;        // reverse pass: processing order of the trajectories (or nullptr = identity)
  // ---- transition stream (replay-buffer sampler, approximate_methods.py:513-545): trajectory `t` writes the tuple of
  //      its pass k at slot tr_base[t] + k of the five arrays below.  nullptr = no stream.
  const long long* tr_base;
  float* tr_state;               // [n][d]  state the pass starts from
  float* tr_action;              // [n][d]
  float* tr_reward;              // [n]
  float* tr_next;                // [n][d]  state after the pass (also computed on the pass that detects the hit)
  unsigned char* tr_done;        // [n]
  // ---- resumable rollouts (tail compaction, see rollout_fwd_inst.cuh).  A launch draws work items from
  //      [continuation records | fresh trajectories]; when its step budget ends, live lanes dump a record.
  const unsigned char* cont_in;        // records to resume, or nullptr
  const unsigned* cont_count_in;       // device-side number of records in cont_in
  unsigned char* cont_out;             // where live lanes dump their state at the deadline, or nullptr (run to completion)
  unsigned* cont_count_out;
  long long K_fresh;                   // fresh trajectories started by this launch (ids [0, K_fresh))
  unsigned round_steps;                // warp iterations after which live lanes dump (0xffffffff: never)
  unsigned drain_steps;                // iterations a warp keeps going once the work counter is seen exhausted
  // scratch handed in by the C ABI (caller's workspace)
  unsigned long long* ws_work_counters;
  unsigned* ws_cont_counts;
  unsigned char* ws_cont_buf[2];
  long long ws_cont_capacity;          // records per buffer
};

// State of an in-flight trajectory: everything a lane needs to continue it (x, accumulators, pass index).
template <int D, bool F64>
struct alignas(8) ContRec {
  long long traj;
  int k;
  int pad;
  typename RealT<F64>::type x[D];
  typename RealT<F64>::type G, S, L2;
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <int D>
struct NoisePlan {
  // passes served by one Philox block (D = 1, 2) or blocks needed per pass (D >= 3)
  static constexpr int SPB = (D == 1) ? 4 : (D == 2 ? 2 : 1);
  static constexpr int BPP = (D <= 2) ? 1 : (D + 3) / 4;
  static constexpr int NZ = 4 * BPP;
};

template <int D, int H, bool F64, bool FAST>
__global__ void __launch_bounds__(128) rollout_fwd_kernel(const __grid_constant__ MlpConst<D, H> W,
                                                          const __grid_constant__ FwdArgs A) {
  typedef typename RealT<F64>::type real;
  constexpr int SPB = NoisePlan<D>::SPB;
  constexpr int BPP = NoisePlan<D>::BPP;
  constexpr int NZ = NoisePlan<D>::NZ;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool inject = (A.flags & RLSDE_F_NOISE_INJECTED) != 0;
  const bool s_exact = (A.flags & RLSDE_F_STOCH_INT_EXACT) != 0;
  const bool store_path = (A.flags & RLSDE_F_STORE_PATH) != 0 && A.path != nullptr;
  const bool want_l2 = (D == 1) && A.policy_opt != nullptr && A.l2 != nullptr;
  const long long lim = inject ? (A.noise_steps < A.n_steps_lim ? A.noise_steps : A.n_steps_lim) : A.n_steps_lim;
  typedef ContRec<D, F64> Rec;
  const long long n_cont = A.cont_count_in ? (long long)*A.cont_count_in : 0;
  const long long n_work = n_cont + A.K_fresh;
  unsigned deadline = A.round_steps;

  bool alive = false, exhausted = false;
  long long traj = 0;
  int k = 0, ck = 0;
  real x[D];
  real G = 0, S = 0, L2 = 0;
  float z[NZ];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = F64 ? (real)A.x0_d[i] : (real)A.x0_f[i];
#pragma unroll
  for (int i = 0; i < NZ; ++i) z[i] = 0.f;

  for (unsigned it = 0;; ++it) {
    if ((it & (SPB - 1)) == 0) {
      if (A.cont_out != nullptr) {
        // every 16 iterations: has the launch run out of work items?  then finish this round soon, so that the
        // survivors can be re-packed into dense warps by the next launch
        if ((it & 15u) == 0 && deadline == A.round_steps) {
          unsigned long long c = 0;
          if (lane == 0) c = *reinterpret_cast<volatile unsigned long long*>(A.counter);
          c = __shfl_sync(FULL, c, 0);
          if ((long long)c >= n_work && A.drain_steps != 0xffffffffu) {
            const unsigned dl = it + A.drain_steps;
            deadline = dl < deadline ? dl : deadline;
          }
        }
        if (it >= deadline) {
          const unsigned live = __ballot_sync(FULL, alive);
          if (live) {
            unsigned base = 0;
            const int leader = __ffs(live) - 1;
            if (lane == leader) base = atomicAdd(A.cont_count_out, (unsigned)__popc(live));
            base = __shfl_sync(FULL, base, leader);
            if (alive) {
              Rec* r = reinterpret_cast<Rec*>(A.cont_out) + (base + __popc(live & ((1u << lane) - 1u)));
              r->traj = traj; r->k = k; r->pad = 0; r->G = G; r->S = S; r->L2 = L2;
#pragma unroll
              for (int i = 0; i < D; ++i) r->x[i] = x[i];
            }
          }
          break;
        }
      }
      const unsigned need = __ballot_sync(FULL, !alive && !exhausted);
      if (need) {
        unsigned long long base = 0;
        const int leader = __ffs(need) - 1;
        if (lane == leader) base = atomicAdd(A.counter, (unsigned long long)__popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!alive && !exhausted) {
          const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
          if (idx < n_cont) {
            const Rec* r = reinterpret_cast<const Rec*>(A.cont_in) + idx;
            traj = r->traj; k = r->k; G = r->G; S = r->S; L2 = r->L2; alive = true;
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] = r->x[i];
            const int rem = k % A.ckpt_every;
            ck = rem ? A.ckpt_every - rem : 0;
          } else if (idx < n_work) {
            traj = idx - n_cont; k = 0; ck = 0; alive = true;
            G = 0; S = 0; L2 = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] = F64 ? (real)A.x0_d[i] : (real)A.x0_f[i];
          } else {
            exhausted = true;
          }
        }
      }
      if (!__any_sync(FULL, alive)) break;
      if (!inject) {
        const unsigned long long gt = (unsigned long long)(A.traj_offset + traj);
#pragma unroll
        for (int q = 0; q < BPP; ++q) {
          float zz[4];
          noise_block(A.seed, gt, (unsigned)(k / SPB) * BPP + q, A.noise_scale2, zz);
#pragma unroll
          for (int s = 0; s < 4; ++s) z[4 * q + s] = zz[s];
        }
      }
    }

    // ---- policy
    float xf[D], u[D];
#pragma unroll
    for (int i = 0; i < D; ++i) xf[i] = (float)x[i];
    mlp_forward<D, H, FAST>(W, xf, u);

    // ---- this pass's Brownian increments
    float dB[D];
    if (inject) {
      const long long row = (long long)k * A.K_global + (A.traj_offset + traj);
#pragma unroll
      for (int i = 0; i < D; ++i) dB[i] = (alive && k < lim) ? __ldg(A.noise + row * D + i) : 0.f;
    } else {
      const int sub = (int)(it & (SPB - 1));
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if constexpr (SPB == 1) {
          dB[i] = z[i];
        } else {
          float v = z[i];
#pragma unroll
          for (int s = 1; s < SPB; ++s) v = (sub == s) ? z[s * D + i] : v;
          dB[i] = v;
        }
      }
    }

    // ---- hit test on the current state
    bool hit;
    if (A.hit_rule == RLSDE_HIT_X0_IN_LB_RB) {
      hit = F64 ? ((double)x[0] >= A.lb_d && (double)x[0] <= A.rb_d) : ((float)x[0] >= A.lb_f && (float)x[0] <= A.rb_f);
    } else {
      hit = true;
#pragma unroll
      for (int i = 0; i < D; ++i) hit = hit && (F64 ? ((double)x[i] >= A.lb_d) : ((float)x[i] >= A.lb_f));
    }

    // ---- stochastic integral, running cost, l2 error
    real su = 0;
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      su = (i == 0) ? mul_rn((real)u[i], (real)dB[i]) : add_rn(su, mul_rn((real)u[i], (real)dB[i]));
      n2 = (i == 0) ? __fmul_rn(u[i], u[i]) : __fadd_rn(n2, __fmul_rn(u[i], u[i]));
    }
    const real S_prev = S;
    S = add_rn(S, su);
    if (want_l2) {
      // idx = floor((clip(x) - lo) / h)   environments.py:318-321;  |u - u_opt|^2 dt  approximate_methods.py:610-615
      double xc = (double)x[0];
      xc = xc < A.grid_lo ? A.grid_lo : (xc > A.grid_hi ? A.grid_hi : xc);
      long long gi = (long long)floor((xc - A.grid_lo) / A.grid_h);
      gi = gi < 0 ? 0 : (gi >= A.n_grid ? A.n_grid - 1 : gi);
      const float uo = alive ? __ldg(A.policy_opt + gi) : 0.f;
      const float du = __fsub_rn(u[0], uo);
      L2 = F64 ? (real)__dadd_rn((double)L2, __dmul_rn((double)__fmul_rn(du, du), A.dt_d))
               : (real)__fadd_rn((float)L2, __fmul_rn(__fmul_rn(du, du), A.dt_f));
    }

    if (alive) {
      if (store_path) {
        if (ck == 0) {
          float* dst = A.path + ((long long)traj * A.ckpt_stride + k / A.ckpt_every) * D;
#pragma unroll
          for (int i = 0; i < D; ++i) dst[i] = (float)x[i];
          ck = A.ckpt_every;
        }
        --ck;
      }
      // running cost  r = -(1 + 0.5 |u|^2) dt   (environments.py:104-110,121-127; f = 1, g = 0)
      real r;
      {
        const float nn = (D == 1) ? n2 : __fmul_rn(sqrtf(n2), sqrtf(n2));
        if (F64) {
          // numpy: f (float64 ones) + float32(0.5 * norm(a)^2)  -> float64, times python-float dt
          r = (real)(-__dmul_rn(__dadd_rn(1.0, (double)__fmul_rn(0.5f, nn)), A.dt_d));
        } else {
          r = (real)(-__fmul_rn(__fadd_rn(1.0f, __fmul_rn(0.5f, nn)), A.dt_f));
        }
      }
      // Euler-Maruyama:  x + (-gradV(x) + sigma u) dt + sigma dB,  gradV = 4 alpha x (x^2 - 1)
      real xn[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (F64) {
          const double xi = (double)x[i];
          double g;
          if (D == 1 && k == 0) {
            // numpy 1-D: the first pass sees a float32 state and a python-float alpha -> float32 gradient
            const float xs = (float)xi;
            g = (double)__fmul_rn(__fmul_rn(A.c4a_f[i], xs), __fsub_rn(__fmul_rn(xs, xs), 1.0f));
          } else if (k == 0) {
            // numpy d-D: float64 alpha array, but state**2 - 1 is still float32 on the first pass
            const float xs = (float)xi;
            g = __dmul_rn(__dmul_rn(A.c4a_d[i], xi), (double)__fsub_rn(__fmul_rn(xs, xs), 1.0f));
          } else {
            g = __dmul_rn(__dmul_rn(A.c4a_d[i], xi), __dsub_rn(__dmul_rn(xi, xi), 1.0));
          }
          const double drift = __dmul_rn(__dadd_rn(-g, __dmul_rn(A.sigma_d, (double)u[i])), A.dt_d);
          xn[i] = (real)__dadd_rn(__dadd_rn(xi, drift), __dmul_rn(A.sigma_d, (double)dB[i]));
        } else {
          const float xi = (float)x[i];
          const float g = __fmul_rn(__fmul_rn(A.c4a_f[i], xi), __fsub_rn(__fmul_rn(xi, xi), 1.0f));
          const float drift = __fmul_rn(__fadd_rn(-g, __fmul_rn(A.sigma_f, u[i])), A.dt_f);
          xn[i] = (real)__fadd_rn(__fadd_rn(xi, drift), __fmul_rn(A.sigma_f, dB[i]));
        }
      }
      if (A.tr_base != nullptr) {
        // (state, action, reward, next_state, done) of this pass, cast to the replay buffer's float32 arrays
        // (replay_buffers.py:23-33,56-68); on the detecting pass the reward is -g(x) = -0 (environments.py:104-110)
        const long long slot = A.tr_base[traj] + k;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          A.tr_state[slot * D + i] = (float)x[i];
          A.tr_action[slot * D + i] = u[i];
          A.tr_next[slot * D + i] = (float)xn[i];
        }
        A.tr_reward[slot] = hit ? -0.0f : (float)r;
        A.tr_done[slot] = hit ? 1 : 0;
      }
      if (hit) {
        const real Sout = s_exact ? S_prev : S;
        if (F64) {
          ((double*)A.G)[traj] = (double)G;
          ((double*)A.S)[traj] = (double)Sout;
          if (A.l2) ((double*)A.l2)[traj] = (double)L2;
          if (A.logw) ((double*)A.logw)[traj] = (double)G - (double)S_prev;
        } else {
          ((float*)A.G)[traj] = (float)G;
          ((float*)A.S)[traj] = (float)Sout;
          if (A.l2) ((float*)A.l2)[traj] = (float)L2;
          if (A.logw) ((float*)A.logw)[traj] = (float)G - (float)S_prev;
        }
        A.T[traj] = k;
        alive = false;
      } else {
        G = add_rn(G, r);
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = xn[i];
        ++k;
        if (k >= lim) {
          // not detected within the pass budget: flagged, never garbage (failure row)
          if (F64) {
            ((double*)A.G)[traj] = (double)G;
            ((double*)A.S)[traj] = (double)S;
            if (A.l2) ((double*)A.l2)[traj] = (double)L2;
            if (A.logw) ((double*)A.logw)[traj] = (double)G - (double)S;
          } else {
            ((float*)A.G)[traj] = (float)G;
            ((float*)A.S)[traj] = (float)S;
            if (A.l2) ((float*)A.l2)[traj] = (float)L2;
            if (A.logw) ((float*)A.logw)[traj] = (float)G - (float)S;
          }
          A.T[traj] = -1;
          alive = false;
        }
      }
    }
  }
}

// host-side launcher for one (D, H) shape; defined in rollout_fwd_inst.cuh
template <int D, int H>
int launch_rollout_fwd(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream);

}  // namespace rlsde
