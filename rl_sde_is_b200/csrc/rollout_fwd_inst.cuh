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
  // ---- single launch: small batches, or no scratch for continuation records
  const long long lanes = grid * block;
  const bool compact = block == 128 && args.ws_cont_buf[0] != nullptr && lanes <= args.ws_cont_capacity &&
                       getenv("RLSDE_FWD_NO_COMPACTION") == nullptr;
  FwdArgs a = args;
  a.K_fresh = args.K;
  a.cont_in = nullptr; a.cont_count_in = nullptr; a.cont_out = nullptr; a.cont_count_out = nullptr;
  a.round_steps = 0xffffffffu; a.drain_steps = 0xffffffffu;
  a.counter = args.ws_work_counters;
  if (!compact) {
    kern<<<(unsigned)grid, block, 0, stream>>>(W, a);
    note_kernel_launches(1);
    return (int)cudaGetLastError();
  }
  // ---- tail compaction.  The main launch stops DRAIN iterations after the work counter runs dry: its live
  // lanes dump their trajectories (24 B each at d = 1) and leave.  Resume rounds then re-pack the survivors
  // into dense warps, each round running a bounded number of passes; as trajectories finish, fewer warps stay
  // resident and the remaining ones run faster.  Without this, warps stay resident until the last of their 32
  // lanes has finished and ~15 % of the executed passes of the 1e6-trajectory workload are idle lanes.
  a.cont_out = args.ws_cont_buf[0];
  a.cont_count_out = args.ws_cont_counts + 0;
  a.drain_steps = 32;
  kern<<<(unsigned)grid, block, 0, stream>>>(W, a);
  const long long lim = (args.flags & RLSDE_F_NOISE_INJECTED) && args.noise_steps < args.n_steps_lim ? args.noise_steps : args.n_steps_lim;
  long long done = 0;
  unsigned budget = 96;
  int r = 0;
  for (; r < 24 && done < lim; ++r) {
    FwdArgs b = a;
    b.K_fresh = 0;
    b.counter = args.ws_work_counters + 1 + r;
    b.cont_in = args.ws_cont_buf[r & 1]; b.cont_count_in = args.ws_cont_counts + r;
    b.cont_out = args.ws_cont_buf[(r + 1) & 1]; b.cont_count_out = args.ws_cont_counts + r + 1;
    b.round_steps = budget; b.drain_steps = 0xffffffffu;
    kern<<<(unsigned)grid, block, 0, stream>>>(W, b);
    done += budget;
    if (r & 1) budget = budget < (1u << 20) ? budget + budget / 2 : budget;   // 96 96 144 144 216 216 324 ...
    budget = (budget + 3u) & ~3u;
  }
  FwdArgs f = a;                       // whatever is left runs to completion
  f.K_fresh = 0;
  f.counter = args.ws_work_counters + 1 + r;
  f.cont_in = args.ws_cont_buf[r & 1]; f.cont_count_in = args.ws_cont_counts + r;
  f.cont_out = nullptr; f.cont_count_out = nullptr;
  f.round_steps = 0xffffffffu; f.drain_steps = 0xffffffffu;
  kern<<<(unsigned)grid, block, 0, stream>>>(W, f);
  note_kernel_launches(r + 2);          // main launch + r resume rounds + the run-to-completion launch
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
