// Launcher + explicit instantiation helper for one policy shape (D, H) of the forward rollout.
#pragma once
#include <cstdlib>
#include "rollout_fwd.cuh"

namespace rlsde {

template <int D, int H, bool F64, bool FAST>
static int launch_fwd_variant(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream) {
  MlpConst<D, H> W;
  pack_mlp_const<D, H>(params_host, FAST, W);
  auto kern = rollout_fwd_kernel<D, H, F64, FAST>;
  // Small batches: one warp per block so the warps spread over SMs (latency-bound regime);
  // large batches: persistent grid of 128-thread blocks, as many as are co-resident.
  int block = 128;
  long long grid;
  if (args.K <= (long long)sm_count * 128) {
    block = 32;
    grid = (args.K + 31) / 32;
  } else {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (const char* e = getenv("RLSDE_FWD_BLOCKS_PER_SM")) {   // tuning knob (bench / profiling only)
      const int v = atoi(e);
      if (v >= 1 && v < per_sm) per_sm = v;
    }
    grid = (long long)sm_count * per_sm;
    const long long need = (args.K + block - 1) / block;
    if (grid > need) grid = need;
  }
  if (grid < 1) grid = 1;
  // Time slicing (rollout_fwd.cuh): large batches only, and only if the caller's workspace holds a ring with one record
  // per trajectory plus one per lane.  RLSDE_FWD_QUANTUM: passes per slice (tuning knob; 0 = run to completion).
  FwdArgs a = args;
  // Passes per slice.  The launch's tail is ~4 slices of the slowest warp; a hand-off every 8 passes costs ~3 % in steady
  // state.  Measured at n_steps_lim = 1000, 1e6 trajectories, 24 % of them running into the limit: 24.3 / 24.5 / 24.9 /
  // 25.4 / 27.8 ms at 4 / 8 / 16 / 32 / 128 passes, 27.8 ms without slicing -> 1/128 of the pass budget, between 8 and 128.
  // Training rollouts (RLSDE_F_STORE_PATH) run to completion: their launches end with a handful of very long
  // trajectories whose sequential length no schedule can hide (K = 4e5, limit 4000: 22.6 ms unsliced, 23.2-24.2 ms sliced).
  const long long lim_eff = (args.flags & RLSDE_F_NOISE_INJECTED) && args.noise_steps < args.n_steps_lim ? args.noise_steps : args.n_steps_lim;
  int quantum = (int)(lim_eff / 128 > 128 ? 128 : (lim_eff / 128 < 8 ? 8 : lim_eff / 128));
  if (args.flags & RLSDE_F_STORE_PATH) quantum = 0;
  // Budgets far above the typical length (the metastable configuration: mean 7e4 passes, a few trajectories near the
  // 1e6-pass limit) also run to completion: their tail is the sequential length of the longest trajectory at the pace
  // of a lone warp, and round-robin slices only delay it (8e6 trajectories on 8 GPUs: 5.96 s unsliced, 6.36 s sliced).
  if (lim_eff > 16384) quantum = 0;
  if (const char* e = getenv("RLSDE_FWD_QUANTUM")) quantum = atoi(e);
  if (quantum > 0) {                                      // a power of two, at least one noise block (4 passes)
    int q2 = 4;
    while (q2 < quantum && q2 < (1 << 30)) q2 <<= 1;
    quantum = q2;
  }
  const long long lanes = grid * block;
  long long need_cap = 1;
  int cap_log2 = 0;
  while (need_cap < args.K + lanes) { need_cap <<= 1; ++cap_log2; }      // power of two: slot = index & (cap - 1)
  const size_t rec = sizeof(ContRec<D, F64>);
  if (block == 128 && quantum > 0 && args.q_ring != nullptr && (size_t)need_cap * rec <= (size_t)args.q_cap) {
    a.q_cap = need_cap;                                  // args.q_cap came in as the BYTES available for the ring
    a.q_cap_log2 = cap_log2;
    a.q_quantum = quantum;
    cudaError_t e = cudaMemsetAsync(a.q_ring, 0, (size_t)need_cap * rec, stream);
    if (e != cudaSuccess) return (int)e;
  } else {
    a.q_ring = nullptr; a.q_cap = 0; a.q_cap_log2 = 0; a.q_quantum = 0;
  }
  kern<<<(unsigned)grid, block, 0, stream>>>(W, a);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

template <int D, int H>
int launch_rollout_fwd(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream) {
  const bool f64 = (args.flags & RLSDE_F_STATE_F64) != 0;
  const bool fast = (args.flags & RLSDE_F_TANH_FAST) != 0;
  if (f64) {
    return fast ? launch_fwd_variant<D, H, true, true>(params_host, args, sm_count, stream)
                : launch_fwd_variant<D, H, true, false>(params_host, args, sm_count, stream);
  }
  return fast ? launch_fwd_variant<D, H, false, true>(params_host, args, sm_count, stream)
              : launch_fwd_variant<D, H, false, false>(params_host, args, sm_count, stream);
}

}  // namespace rlsde

#define RLSDE_INSTANTIATE_FWD(D, H) \
  template int rlsde::launch_rollout_fwd<D, H>(const float*, const rlsde::FwdArgs&, int, cudaStream_t);
