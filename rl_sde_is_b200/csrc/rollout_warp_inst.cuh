// Launchers + explicit instantiation helper for the warp-per-trajectory kernels of one state dimension D (H = 32).
#pragma once
#include "rollout_bwd_inst.cuh"
#include "rollout_warp.cuh"

namespace rlsde {

static inline long long warp_grid(long long K, int sm_count) {
  long long g = (K + 3) / 4;
  const long long cap = (long long)sm_count * 4;     // 16 warps per SM: every warp still has an SM sub-partition mostly to itself
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

template <int D, bool F64, bool FAST>
static int launch_fwd_warp_variant(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream) {
  MlpConst<D, WARP_H> W;
  pack_mlp_const<D, WARP_H>(params_host, FAST, W);
  rollout_fwd_warp_kernel<D, F64, FAST><<<(unsigned)warp_grid(args.K, sm_count), 128, 0, stream>>>(W, args);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

template <int D>
int launch_rollout_fwd_warp(const float* params_host, const FwdArgs& args, int sm_count, cudaStream_t stream) {
  const bool f64 = (args.flags & RLSDE_F_STATE_F64) != 0, fast = (args.flags & RLSDE_F_TANH_FAST) != 0;
  if (f64) return fast ? launch_fwd_warp_variant<D, true, true>(params_host, args, sm_count, stream)
                       : launch_fwd_warp_variant<D, true, false>(params_host, args, sm_count, stream);
  return fast ? launch_fwd_warp_variant<D, false, true>(params_host, args, sm_count, stream)
              : launch_fwd_warp_variant<D, false, false>(params_host, args, sm_count, stream);
}

template <int D, bool FAST>
static int launch_bwd_warp_variant(const float* params_host, const FwdArgs& args, float scale, float* grad, float* partial,
                                   int sm_count, cudaStream_t stream) {
  constexpr int P = D * WARP_H + WARP_H + WARP_H * WARP_H + WARP_H + WARP_H * D + D;
  MlpConst<D, WARP_H> W;
  pack_mlp_const<D, WARP_H>(params_host, FAST, W);
  const long long grid = warp_grid(args.K, sm_count);
  rollout_bwd_warp_kernel<D, FAST><<<(unsigned)grid, 128, 0, stream>>>(W, args, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  bwd_reduce_kernel<<<(P + 127) / 128, 128, 0, stream>>>(partial, (int)(grid * 4), P, scale, grad);
  note_kernel_launches(2);
  return (int)cudaGetLastError();
}

template <int D>
int launch_rollout_bwd_warp(const float* params_host, const FwdArgs& args, float scale, float* grad, float* partial,
                            int sm_count, cudaStream_t stream) {
  return (args.flags & RLSDE_F_TANH_FAST) ? launch_bwd_warp_variant<D, true>(params_host, args, scale, grad, partial, sm_count, stream)
                                          : launch_bwd_warp_variant<D, false>(params_host, args, scale, grad, partial, sm_count, stream);
}

}  // namespace rlsde

#define RLSDE_INSTANTIATE_WARP(D)                                                                                   \
  template int rlsde::launch_rollout_fwd_warp<D>(const float*, const rlsde::FwdArgs&, int, cudaStream_t);           \
  template int rlsde::launch_rollout_bwd_warp<D>(const float*, const rlsde::FwdArgs&, float, float*, float*, int, cudaStream_t);
