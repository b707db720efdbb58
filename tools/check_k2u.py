#!/usr/bin/env python
"""K2u (tcgen05 reverse pass, producer / consumer CTAs) against K2x (CUDA-core tile kernel) on the same forward rollout.
    python tools/check_k2u.py [--H 256] [--d 1] [--K 20000] [--lim 1500] [--time]"""
import argparse, json, os, sys, time
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L, rollout as R
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D, DoubleWellStoppingTimeND
from rl_sde_is_b200.models import DeterministicPolicy
ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=256); ap.add_argument("--d", type=int, default=1); ap.add_argument("--K", type=int, default=20000)
ap.add_argument("--lim", type=int, default=1500); ap.add_argument("--time", action="store_true"); ap.add_argument("--skip-ffma", action="store_true")
a = ap.parse_args()
d, H, K = a.d, a.H, a.K
env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005) if d == 1 else DoubleWellStoppingTimeND(d, beta=1.0, alpha=1.0, dt=0.005)
torch.manual_seed(4)
m = DeterministicPolicy(d, d, [H, H], nn.Tanh())
m.policy[4].bias.data.fill_(1.0 if d == 1 else 3.0)
params = R.flat_parameters(m).detach().numpy()
env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(d, H)
fwd = R.rollout_forward(env_c, mlp_c, params, K, seed=12, n_steps_lim=a.lim, store_path=True, ckpt_every=1, tuning={"wide_kernel": "umma"})
T = fwd.T.cpu().numpy()
print(json.dumps({"K": K, "H": H, "d": d, "hit": int((T >= 0).sum()), "mean_T": float(T[T >= 0].mean()), "max_T": int(T.max())}), flush=True)
res = {}
for name, wk in (("umma", 1), ("ffma", 2), ("umma_again", 1)):
    if name.startswith("ffma") and a.skip_ffma:
        continue
    fwd.cfg.wide_kernel = wk
    ts = []
    for it in range(3 if a.time else 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = R.rollout_backward(env_c, mlp_c, params, fwd, 1.0 / K, balance=not name.endswith('unsorted'))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = g.cpu().numpy().astype(np.float64)
    steps = float((T[T >= 0] + 1).sum())
    print(json.dumps({"kernel": name, "ms": min(ts), "steps_per_s": steps / min(ts) * 1e3, "finite": bool(np.isfinite(res[name]).all()),
                      "norm": float(np.linalg.norm(res[name]))}), flush=True)
if "umma" in res and "ffma" in res:
    lo, hi = d * H + H, d * H + H + H * H
    x, y = res["umma"][lo:hi], res["ffma"][lo:hi]
    print(json.dumps({"W2 shrink coefficient (x-y).y/|y|^2": float(((x - y) * y).sum() / (y * y).sum()), "rms_err_over_rms": float(np.sqrt(((x - y) ** 2).mean() / (y * y).mean())),
                      "corr(err, y)": float(np.corrcoef(x - y, y)[0, 1])}))
if len(res) >= 2 and "ffma" in res:
    lo, hi = d * H + H, d * H + H + H * H
    for k1 in res:
        for k2 in res:
            if k1 < k2:
                print(json.dumps({"W2 pair": [k1, k2], "err": float(np.abs(res[k1][lo:hi] - res[k2][lo:hi]).max() / np.abs(res[k2][lo:hi]).max())}))
if "ffma" in res:
    P = {"W1": (0, d * H), "b1": (d * H, d * H + H), "W2": (d * H + H, d * H + H + H * H), "b2": (d * H + H + H * H, d * H + 2 * H + H * H),
         "W3": (d * H + 2 * H + H * H, 2 * d * H + 2 * H + H * H), "b3": (2 * d * H + 2 * H + H * H, 2 * d * H + 2 * H + H * H + d)}
    ok = True
    for k, (lo, hi) in P.items():
        x, y = res["umma"][lo:hi], res["ffma"][lo:hi]
        err = np.abs(x - y).max() / max(np.abs(y).max(), 1e-30)
        print(json.dumps({"block": k, "max_abs_err_over_max": err, "max_ref": float(np.abs(y).max())}))
        ok = ok and err < 2e-3
    print("K2U_OK" if ok else "K2U_MISMATCH")
