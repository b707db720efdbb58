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
// trajectory retires takes the next work item (warp-aggregated atomicAdd) at the next noise-block
// boundary, so lanes stay busy until the work runs out.
//
// Time slicing.  Trajectory lengths are unknown in advance and bounded only by n_steps_lim, so with
// run-to-completion scheduling the launch ends with a tail of (longest trajectory) x (time of a pass): the
// last-started long trajectory runs on an otherwise empty GPU (measured: 4.3 ms of a 27.5 ms launch at
// 1e6 trajectories, whatever the batch size).  Instead a trajectory runs for at most `q_quantum` passes
// per slice; a lane whose slice is over appends the trajectory's state (24 B at d = 1) to a FIFO in
// global memory and takes the next item.  Items are served in order -- first all fresh trajectories,
// then their continuations in the order they were queued -- which makes the schedule breadth-first: all
// trajectories advance together and the tail shrinks to a few quanta.  (Warps do not advance at the same pace -- with
// eight warps per scheduler some execute 8 000 passes while others execute 700 in the same launch -- so the last
// slice of a trajectory that sits in a starved warp is what ends the launch: short slices, 8 passes at n_steps_lim =
// 1000, measured best; they cost ~3 % in steady state.)  The FIFO is a ring of records with
// a sequence word per slot (producer: write, fence, publish epoch; consumer: claim an index, poll its
// slot between passes, never blocking the warp's other lanes); the launch ends when every trajectory has
// reported completion.  Per-trajectory arithmetic does not depend on the slicing: results are bit-identical.
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
  int ckpt_log2;                 // log2(ckpt_every) if it is a power of two (checkpoint index = k >> ckpt_log2), else -1
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
  int grad_accumulate;           // reverse pass: add this launch's gradient to grad instead of overwriting it
  int blocks_per_sm_cap;         // > 0: cap on resident blocks per SM of the thread-per-trajectory forward kernel (tuning)
  // ---- transition stream (replay-buffer sampler, approximate_methods.py:513-545): trajectory `t` writes the tuple of
  //      its pass k at slot tr_base[t] + k of the five arrays below.  nullptr = no stream.
  const long long* tr_base;
  float* tr_state;               // [n][d]  state the pass starts from
  float* tr_action;              // [n][d]
  float* tr_reward;              // [n]
  float* tr_next;                // [n][d]  state after the pass (also computed on the pass that detects the hit)
  unsigned char* tr_done;        // [n]
  // ---- time-sliced scheduling (see the header comment).  q_ring == nullptr: every trajectory runs to completion.
  unsigned char* q_ring;               // ring of ContRec<D, F64> records, all-zero at launch
  long long q_cap;                     // records in the ring: a power of two >= K + lanes of the grid
  int q_cap_log2;
  unsigned long long* q_ctrl;          // [0] items claimed (== counter), [1] continuation records queued, [2] trajectories
                                       // completed, [3] records taken by the resume kernel, [4] trajectories that ran into
                                       // the pass budget, [5] tail strategy of an adaptive launch
  int q_quantum;                       // passes per slice (0: run to completion)
  int q_adaptive;                      // 1: start run-to-completion and pick the tail strategy from what the launch sees:
                                       //    q_ctrl[5] = 1 (time slices) once >= 5 % of the completed trajectories ran into the
                                       //    pass budget, = 2 (hand-off) if the work runs out first
  long long q_handoff;                 // > 0 (run-to-completion only): once all trajectories have been started and at most
                                       // this many are still live, they are left in the ring for the warp-per-trajectory
                                       // kernel, which advances a lone trajectory ~5x faster (rollout_warp.cuh, RESUME)
};

// State of an in-flight trajectory: everything a lane needs to continue it (x, accumulators, pass index).
template <int D, bool F64>
struct alignas(8) ContRec {
  long long traj;
  int k;
  int seq;            // 0: never used; e > 0: holds the e-th record of this slot; -e: that record has been taken
  typename RealT<F64>::type x[D];
  typename RealT<F64>::type G, S, L2;
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// Cold path of the FIFO: a producer found its slot still holding an unread record (the ring is larger than the number of
// records in flight, so this needs a reader stalled for a whole lap).  Kept out of line: inlined, the spin loop costs the
// rollout kernel ~15 registers.
static __device__ __noinline__ void wait_for_slot(volatile int* seq, int free_mark) {
  while (*seq != free_mark) __nanosleep(64);
}

// A warp without a live lane parks here until the record one of its lanes waits for has been published, or every
// trajectory has completed.  Out of line and with exponential back-off: towards the end of a launch most warps are in
// this state, and a tight poll loop (the whole refill block every 200 ns) took the issue slots the last working warps
// needed (measured: the end of the launch stretched from ~0.4 ms to 2-4 ms).
static __device__ __noinline__ void park_idle_warp(volatile int* seq, int epoch, bool pending,
                                                   const volatile unsigned long long* completed, unsigned long long K) {
  unsigned ns = 128;
  for (;;) {
    const bool wake = (pending && *seq == epoch) || *completed >= K;
    if (__any_sync(0xffffffffu, wake)) return;
    __nanosleep(ns);
    if (ns < 4096) ns <<= 1;
  }
}

template <int D>
struct NoisePlan {
  // passes served by one Philox block (D = 1, 2) or blocks needed per pass (D >= 3)
  static constexpr int SPB = (D == 1) ? 4 : (D == 2 ? 2 : 1);
  static constexpr int BPP = (D <= 2) ? 1 : (D + 3) / 4;
  static constexpr int NZ = 4 * BPP;
};

// Register budgets.  Left alone, ptxas settles at 67 registers for d = 1 and at ~250 for the wider shapes (2 blocks per
// SM: two warps per scheduler, every FFMA2 waiting for its LDCU.128), although spill-free allocations exist at 64
// (d = 1: 8 blocks per SM), <= 102 (d = 2..4: 5 blocks) and <= 128 (d = 10: 4 blocks) -- it finds them when asked.
template <int D, int H, bool F64, bool FAST>
__global__ void __launch_bounds__(128, (D == 1 ? 8 : (D <= 4 ? 5 : 4))) rollout_fwd_kernel(const __grid_constant__ MlpConst<D, H> W,
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
  // tail strategy: 1 = time slices, 2 = hand-off to the latency kernel, 3 = none, 0 = adaptive launch, not decided yet
  const bool adaptive = A.q_ring != nullptr && A.q_adaptive != 0;
  const bool tracked = A.q_ring != nullptr && (A.q_quantum > 0 || A.q_handoff > 0);
  int mode = adaptive ? 0 : (A.q_ring != nullptr && A.q_quantum > 0 ? 1 : (A.q_ring != nullptr && A.q_handoff > 0 ? 2 : 3));
  Rec* const ring = reinterpret_cast<Rec*>(A.q_ring);
  // live state is kept small (the register budget of 8 blocks per SM is 64): while a lane waits for a record, `traj`
  // holds the record's index; slices end where the pass index is a multiple of the quantum (a power of two)
  const int qmask = A.q_quantum - 1;

  bool alive = false, exhausted = false, pending = false, fin = false, capped = false;
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
      if (tracked) {
        // (a) report the trajectories that completed since the last boundary
        const unsigned fm = __ballot_sync(FULL, fin);
        if (fm) {
          if (lane == __ffs(fm) - 1) atomicAdd(A.q_ctrl + 2, (unsigned long long)__popc(fm));
          fin = false;
        }
      }
      if (mode == 0) {
        // adaptive launch, tail strategy not decided yet: count the trajectories that ran into the pass budget and, every
        // other boundary, look whether somebody has decided / whether it is time to decide for time slices
        const unsigned cm = __ballot_sync(FULL, capped);
        if (cm) {
          if (lane == __ffs(cm) - 1) atomicAdd(A.q_ctrl + 4, (unsigned long long)__popc(cm));
          capped = false;
        }
        if ((it & (2u * SPB - 1u)) == 0) {
          unsigned long long m = 0;
          if (lane == 0) {
            m = *reinterpret_cast<volatile unsigned long long*>(A.q_ctrl + 5);
            if (m == 0) {
              const unsigned long long n_cap = *reinterpret_cast<volatile unsigned long long*>(A.q_ctrl + 4);
              const unsigned long long n_done = *reinterpret_cast<volatile unsigned long long*>(A.q_ctrl + 2);
              if (n_done >= 4096 && n_cap * 20 >= n_done) {
                const unsigned long long old = atomicCAS(A.q_ctrl + 5, 0ull, 1ull);
                m = old == 0 ? 1 : old;
              }
            }
          }
          mode = (int)__shfl_sync(FULL, m, 0);
        }
      }
      if (mode == 2 && (it & (4u * SPB - 1u)) == 0) {
        // every 4th boundary: is this the tail of the launch?  then leave the live trajectories to the latency kernel
        const unsigned long long claimed = *reinterpret_cast<volatile unsigned long long*>(A.q_ctrl);
        const unsigned long long completed = *reinterpret_cast<volatile unsigned long long*>(A.q_ctrl + 2);
        if (claimed >= (unsigned long long)A.K && (unsigned long long)A.K - completed <= (unsigned long long)A.q_handoff) {
          const unsigned live = __ballot_sync(FULL, alive);
          if (live) {
            unsigned long long base = 0;
            const int leader = __ffs(live) - 1;
            if (lane == leader) base = atomicAdd(A.q_ctrl + 1, (unsigned long long)__popc(live));
            base = __shfl_sync(FULL, base, leader);
            if (alive) {
              Rec* r = ring + (base + __popc(live & ((1u << lane) - 1u)));
              r->traj = traj; r->k = k; r->seq = 1; r->G = G; r->S = S; r->L2 = L2;
#pragma unroll
              for (int i = 0; i < D; ++i) r->x[i] = x[i];
            }
          }
          break;
        }
      }
      if (mode == 1) {
        // (b) slice over: append the trajectory to the FIFO
        const bool expire = alive && k != 0 && (k & qmask) == 0;   // (a resumed lane is past this check: it resumes in (d))
        const unsigned em = __ballot_sync(FULL, expire);
        if (em) {
          unsigned long long base = 0;
          const int leader = __ffs(em) - 1;
          if (lane == leader) base = atomicAdd(A.q_ctrl + 1, (unsigned long long)__popc(em));
          base = __shfl_sync(FULL, base, leader);
          if (expire) {
            const unsigned long long j = base + __popc(em & ((1u << lane) - 1u));
            Rec* r = ring + (long long)(j & (unsigned long long)(A.q_cap - 1));
            const int epoch = (int)(j >> A.q_cap_log2) + 1;
            // at most K records are in flight and the ring is larger, so the slot's previous record (item j - cap) was
            // claimed long ago; wait for its reader all the same
            volatile int* seq = &r->seq;
            const int free_mark = epoch == 1 ? 0 : 1 - epoch;
            if (*seq != free_mark) wait_for_slot(seq, free_mark);
            r->traj = traj; r->k = k; r->G = G; r->S = S; r->L2 = L2;
#pragma unroll
            for (int i = 0; i < D; ++i) r->x[i] = x[i];
            __threadfence();
            *seq = epoch;
            alive = false;
          }
        }
      }
      // (c) idle lanes claim the next item: a fresh trajectory, or the index of a continuation record to wait for
      const unsigned need = __ballot_sync(FULL, !alive && !exhausted && !pending);
      if (need) {
        unsigned long long base = 0;
        const int leader = __ffs(need) - 1;
        if (lane == leader) base = atomicAdd(A.counter, (unsigned long long)__popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!alive && !exhausted && !pending) {
          const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
          if (idx < A.K) {
            traj = idx; k = 0; ck = 0; alive = true;
            G = 0; S = 0; L2 = 0;
#pragma unroll
            for (int i = 0; i < D; ++i) x[i] = F64 ? (real)A.x0_d[i] : (real)A.x0_f[i];
          } else {
            // no fresh work left.  An undecided adaptive launch decides now: nothing ran into the budget, so the tail
            // is a few long trajectories -> hand-off (unless another warp has just chosen time slices)
            if (mode == 0) {
              const unsigned long long old = atomicCAS(A.q_ctrl + 5, 0ull, 2ull);
              mode = old == 0 ? 2 : (int)old;
            }
            if (mode == 1) { pending = true; traj = idx - A.K; }
            else exhausted = true;
          }
        }
      }
      // the tail strategy is a property of the launch (one global decision): lanes that have just learnt it in (c) tell
      // the rest of the warp, so that `mode` stays warp-uniform (the blocks above use full-mask collectives)
      if (adaptive) mode = (int)__reduce_max_sync(FULL, (unsigned)mode);
      // (d) lanes waiting for a record look at their slot; the other lanes of the warp are not held up
      if (pending) {
        Rec* r = ring + (traj & (A.q_cap - 1));
        const int epoch = (int)(traj >> A.q_cap_log2) + 1;
        volatile int* seq = &r->seq;
        if (*seq == epoch) {
          __threadfence();
          const volatile Rec* v = r;
          traj = v->traj; k = v->k; G = v->G; S = v->S; L2 = v->L2;
#pragma unroll
          for (int i = 0; i < D; ++i) x[i] = v->x[i];
          *seq = -epoch;
          alive = true; pending = false;
          const int rem = k % A.ckpt_every;
          ck = rem ? A.ckpt_every - rem : 0;
        } else if (*reinterpret_cast<volatile unsigned long long*>(A.q_ctrl + 2) >= (unsigned long long)A.K) {
          pending = false; exhausted = true;      // every trajectory has completed: no record will ever arrive
        }
      }
      if (!__any_sync(FULL, alive)) {
        if (!__any_sync(FULL, pending)) break;
        park_idle_warp(&(ring + (traj & (A.q_cap - 1)))->seq, (int)(traj >> A.q_cap_log2) + 1, pending, A.q_ctrl + 2,
                       (unsigned long long)A.K);
        it |= (unsigned)(SPB - 1);                  // stay on a boundary: the refill block above takes it from here
        continue;
      }
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
    if constexpr (!FAST && RLSDE_FWD_FOLDED) mlp_forward_folded<D, H>(W, xf, u);      // W is a pack_mlp_const_folded image
    else mlp_forward<D, H, FAST>(W, xf, u);

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
          const int ci = A.ckpt_log2 >= 0 ? (k >> A.ckpt_log2) : (k / A.ckpt_every);      // no integer division per pass
          float* dst = A.path + ((long long)traj * A.ckpt_stride + ci) * D;
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
        alive = false; fin = true;
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
          alive = false; fin = true; capped = true;
        }
      }
    }
  }
}

// host-side launcher for one (D, H) shape; defined in rollout_fwd_inst.cuh
template <int D, int H>
int launch_rollout_fwd(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream);

}  // namespace rlsde
