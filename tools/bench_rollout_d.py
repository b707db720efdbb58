"""Test-rollout throughput (no path store) for a d-dimensional double well: useful steps/s and FP32 fraction."""
import argparse, json, os, sys
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L, rollout as R
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D, DoubleWellStoppingTimeND
from rl_sde_is_b200.models import DeterministicPolicy
ap = argparse.ArgumentParser(); ap.add_argument("--d", type=int, default=10); ap.add_argument("--K", type=int, default=1000000)
ap.add_argument("--lim", type=int, default=2000); ap.add_argument("--bias", type=float, default=3.0)
a = ap.parse_args()
d = a.d
env = DoubleWellStoppingTime1D() if d == 1 else DoubleWellStoppingTimeND(d)
torch.manual_seed(1)
m = DeterministicPolicy(d, d, [32, 32], nn.Tanh()); m.policy[4].bias.data.fill_(a.bias)
params = R.flat_parameters(m).detach().numpy()
env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(d, 32)
ts = []
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = R.rollout_forward(env_c, mlp_c, params, a.K, seed=it, n_steps_lim=a.lim, stoch_int="exact"); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
u = float(out.stats[L.ST_USEFUL_STEPS]); ms = float(np.median(ts[2:]))
F = 2 * (2 * d * 32 + 32 * 32) + 14 * d + 4
print(json.dumps({"d": d, "K": a.K, "lim": a.lim, "useful_steps": u, "mean_steps": u / a.K, "ms": ms, "steps_per_s": u / ms * 1e3,
                  "flop_per_step": F, "fp32_frac": u / ms * 1e3 * F / 74.45e12, "n_unfinished": int(out.stats[L.ST_N_UNFINISHED])}))
