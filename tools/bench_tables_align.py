"""Does the alignment of the rows of P matter?  Na = 601 (rows shift by 72 bytes) against Na = 608 (every row 128-byte aligned)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
for na in (601, 608, 600, 592):
    env = DoubleWellStoppingTime1D(); env.set_action_space_bounds(); env.discretize_state_space(0.01)
    env.action_space_h = -3.0 + 0.01 * np.arange(na); env.n_actions = na; env.h_action = 0.01
    bufs = [torch.empty((env.n_states, env.n_states, na), dtype=torch.float64, device="cuda") for _ in range(2)]
    ts = []
    for trial in range(4):
        compute_p_tensor_batch(env, out=bufs[1])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            compute_p_tensor_batch(env, out=bufs[i & 1])
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 10)
    nb = bufs[0].numel() * 8
    print(json.dumps({"Na": na, "row_shift_bytes": (env.n_states * na * 8) % 128, "ms": min(ts[1:]), "GBps": nb / min(ts[1:]) / 1e6}))
    del bufs
