// warp-per-trajectory (small-batch, latency-oriented) rollout kernels for d = 3, hidden width = 32
#include "rollout_warp_inst.cuh"
RLSDE_INSTANTIATE_WARP(3)
