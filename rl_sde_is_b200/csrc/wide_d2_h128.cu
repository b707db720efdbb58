// wide-policy rollout kernels (tile contraction through shared memory) for d = 2, hidden width = 128
#include "rollout_wide_inst.cuh"
RLSDE_INSTANTIATE_WIDE(2, 128)
