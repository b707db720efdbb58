"""Time the table build (config 3) with CUDA events: quadrature fast path and erf/erfc path."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
env = DoubleWellStoppingTime1D(); env.set_action_space_bounds(); env.discretize_state_space(0.01); env.discretize_action_space(0.01)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for exact in (False, True):
    ts = []
    for it in range(6):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); P = compute_p_tensor_batch(env, device_out=True, exact_cdf=exact); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)); del P
    ms = min(ts[2:])
    print(json.dumps({"path": "erf/erfc" if exact else "gauss-legendre", "ms": ms, "GBps": 773131208 / ms / 1e6, "frac_hbm_6536.7": 773131208 / ms / 1e6 / 6536.7, "all_ms": ts}))
