// tcgen05 reverse pass (producer / consumer CTAs, dW2 in tensor memory) for d = 2, hidden width = 256
#include "rollout_umma_bwd_inst.cuh"
RLSDE_INSTANTIATE_UMMA_BWD(2, 256)
