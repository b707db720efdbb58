// wide-policy rollout kernels (tile contraction through shared memory) for d = 1, hidden width = 256
#include "rollout_wide_inst.cuh"
RLSDE_INSTANTIATE_WIDE(1, 256)
