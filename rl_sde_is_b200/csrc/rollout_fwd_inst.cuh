// Launcher + explicit instantiation helper for one policy shape (D, H) of the forward rollout.
#pragma once
#include "rollout_fwd.cuh"

namespace rlsde {

template <int D, int H, bool F64, bool FAST>
static int launch_fwd_variant(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream) {
  MlpConst<D, H> W;
  if (!FAST && RLSDE_FWD_FOLDED) pack_mlp_const_folded<D, H>(params_host, W);
  else pack_mlp_const<D, H>(params_host, FAST, W);
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
    if (args.blocks_per_sm_cap >= 1 && args.blocks_per_sm_cap < per_sm) per_sm = args.blocks_per_sm_cap;   // tuning knob
    grid = (long long)sm_count * per_sm;
    const long long need = (args.K + block - 1) / block;
    if (grid > need) grid = need;
  }
  if (grid < 1) grid = 1;
  // Schedule (rollout_fwd.cuh; chosen by the C ABI layer: args.q_quantum passes per slice, or run to completion with
  // the tail handed over to the warp-per-trajectory kernel at args.q_handoff live trajectories): large batches only, and
  // only if the caller's workspace holds a ring with one record per trajectory plus one per lane.
  FwdArgs a = args;
  const long long lanes = grid * block;
  long long need_cap = 1;
  int cap_log2 = 0;
  while (need_cap < args.K + lanes) { need_cap <<= 1; ++cap_log2; }      // power of two: slot = index & (cap - 1)
  const size_t rec = sizeof(ContRec<D, F64>);
  if (block == 128 && (args.q_quantum > 0 || args.q_handoff > 0) && args.q_ring != nullptr &&
      (size_t)need_cap * rec <= (size_t)args.q_cap) {
    a.q_cap = need_cap;                                  // args.q_cap came in as the BYTES available for the ring
    a.q_cap_log2 = cap_log2;
    if (a.q_quantum > 0) {                               // the FIFO protocol needs all-zero sequence words
      cudaError_t e = cudaMemsetAsync(a.q_ring, 0, (size_t)need_cap * rec, stream);
      if (e != cudaSuccess) return (int)e;
    }
  } else {
    a.q_ring = nullptr; a.q_cap = 0; a.q_cap_log2 = 0; a.q_quantum = 0; a.q_handoff = 0;
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
