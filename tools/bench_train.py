"""Throughput of one REINFORCE loss + gradient (K1 with checkpoints + K2) at a large batch, CUDA-event timed."""
import argparse, json, os, sys
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L, rollout as R
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D, DoubleWellStoppingTimeND
from rl_sde_is_b200.models import DeterministicPolicy
ap = argparse.ArgumentParser(); ap.add_argument("--d", type=int, default=1); ap.add_argument("--K", type=int, default=400000)
ap.add_argument("--ckpt", type=int, default=1); ap.add_argument("--lim", type=int, default=4000); ap.add_argument("--bias", type=float, default=0.5)
a = ap.parse_args()
d = a.d
env = DoubleWellStoppingTime1D() if d == 1 else DoubleWellStoppingTimeND(d)
torch.manual_seed(1)
m = DeterministicPolicy(d, d, [32, 32], nn.Tanh()); m.policy[4].bias.data.fill_(a.bias)
params = R.flat_parameters(m).detach().numpy()
env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(d, 32)
res = []
for it in range(4):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = R.rollout_forward(env_c, mlp_c, params, a.K, seed=it, n_steps_lim=a.lim, store_path=True, ckpt_every=a.ckpt, want_logw=False)
    e[1].record()
    g = R.rollout_backward(env_c, mlp_c, params, out, 1.0 / a.K)
    e[2].record(); torch.cuda.synchronize()
    st = out.stats
    res.append((st[L.ST_USEFUL_STEPS], e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), int(st[L.ST_N_UNFINISHED]), float(g.norm())))
u, f, b, unf, gn = res[-1]
F_train = 3 * 2 * (2 * d * 32 + 32 * 32) + 24 * d + 4
print(json.dumps({"d": d, "K": a.K, "ckpt_every": a.ckpt, "useful_steps": u, "n_unfinished": unf, "fwd_ms": f, "bwd_ms": b,
                  "fwd_steps_per_s": u / f * 1e3, "bwd_steps_per_s": u / b * 1e3, "train_steps_per_s": u / (f + b) * 1e3,
                  "train_fp32_frac": u / (f + b) * 1e3 * F_train / 74.45e12, "grad_norm": gn}))
