// Launcher + explicit instantiation helper for the tensor-core reverse pass of one state dimension D (H = 32).
#pragma once
#include <cmath>
#include <cstring>
#include "rollout_bwd_mma.cuh"

namespace rlsde {

template <int D, bool FAST>
static int launch_bwd_mma_variant(const float* params_host, const FwdArgs& args, float scale, float* grad, void* partial,
                                  int sm_count, cudaStream_t stream, cudaEvent_t join) {
  constexpr int P = D * MMA_H + MMA_H + MMA_H * MMA_H + MMA_H + MMA_H * D + D;
  static_assert((size_t)P * sizeof(double) <= (size_t)BWD_MAX_PARAMS * sizeof(float) , "fp64 partials must fit the reverse-pass workspace");
  if (args.ckpt_every > BWD_MAX_SEG) return (int)cudaErrorInvalidValue;
  MlpConst<D, MMA_H> W;
  pack_mlp_const<D, MMA_H>(params_host, FAST, W);
  BwdMmaWeights F;
  pack_bwd_mma_weights<D>(W, F);
  typename BwdMmaSmallSel<D>::type FS;
  pack_bwd_mma_small<D>(W, FAST, FS);
  auto kern = rollout_bwd_mma_kernel<D, FAST>;
  int block = 128;
  long long grid;
  // d > 4: fragment table of the d-sized products + per-warp staging rows; accumulator tiles (rollout_bwd_mma.cuh)
  auto dyn_smem = [](int blk) { return bwd_mma_dyn_smem(D, blk); };
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem(128));
  if (args.K <= (long long)sm_count * 128) {
    block = 32;
    grid = (args.K + 31) / 32;
  } else {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, dyn_smem(block)) != cudaSuccess || per_sm < 1) per_sm = 1;
    grid = (long long)sm_count * per_sm;
    const long long need = (args.K + block - 1) / block;
    if (grid > need) grid = need;
  }
  if (grid < 1) grid = 1;
  long long n_warps = grid * (block / 32);
  if (n_warps > BWD_MAX_WARPS) { grid = BWD_MAX_WARPS / (block / 32); n_warps = grid * (block / 32); }
  kern<<<(unsigned)grid, block, dyn_smem(block), stream>>>(W, args, F, FS, reinterpret_cast<double*>(partial));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (join != nullptr && (e = cudaStreamWaitEvent(stream, join, 0)) != cudaSuccess) return (int)e;
  bwd_reduce64_kernel<<<(P + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const double*>(partial), (int)n_warps, P, scale, grad,
                                                           args.grad_accumulate);
  note_kernel_launches(2);
  return (int)cudaGetLastError();
}

template <int D>
int launch_rollout_bwd_mma(const float* params_host, const FwdArgs& args, float scale, float* grad, void* partial,
                           int sm_count, cudaStream_t stream, cudaEvent_t join) {
  return (args.flags & RLSDE_F_TANH_FAST) ? launch_bwd_mma_variant<D, true>(params_host, args, scale, grad, partial, sm_count, stream, join)
                                          : launch_bwd_mma_variant<D, false>(params_host, args, scale, grad, partial, sm_count, stream, join);
}

}  // namespace rlsde

#define RLSDE_INSTANTIATE_BWD_MMA(D) \
  template int rlsde::launch_rollout_bwd_mma<D>(const float*, const rlsde::FwdArgs&, float, float*, void*, int, cudaStream_t, cudaEvent_t);
