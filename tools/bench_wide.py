#!/usr/bin/env python
"""Wide-policy kernels (hidden width 64 / 128 / 256): forward test-rollout throughput and training evaluation, CUDA events.
    python tools/bench_wide.py [--H 256] [--K 200000] [--Ktrain 20000] [--ref]"""
import argparse, json, os, sys, time
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L, rollout as R
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
from rl_sde_is_b200.models import DeterministicPolicy

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=256); ap.add_argument("--K", type=int, default=200000)
ap.add_argument("--Ktrain", type=int, default=20000); ap.add_argument("--ref", action="store_true")
ap.add_argument("--wide", default="auto", help="forward kernel at H = 128 / 256: auto | umma | ffma"); ap.add_argument("--no-train", action="store_true")
a = ap.parse_args()
H = a.H
env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
torch.manual_seed(1)
m = DeterministicPolicy(1, 1, [H, H], nn.Tanh())
params = R.flat_parameters(m).detach().numpy()
F_fwd = 2 * (2 * H + H * H) + 18
F_train = 3 * 2 * (2 * H + H * H) + 28
env_n, env_t, mlp_c = R.env_struct(env, L.HIT_X0_IN_LB_RB), R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(1, H)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = R.rollout_forward(env_n, mlp_c, params, a.K, seed=it, n_steps_lim=1000, stoch_int="exact", tuning={"wide_kernel": a.wide})
    e1.record(); torch.cuda.synchronize()
u = float(out.stats[L.ST_USEFUL_STEPS]); ms = e0.elapsed_time(e1)
print(json.dumps({"what": "forward test rollout", "kernel": a.wide, "H": H, "K": a.K, "ms": ms, "steps_per_s": u / ms * 1e3, "fp32_frac": u / ms * 1e3 * F_fwd / 74.45e12,
                  "flop_per_step": F_fwd}))
m.policy[4].bias.data.fill_(0.5)
params = R.flat_parameters(m).detach().numpy()
for Kt in (() if a.no_train else (1000, a.Ktrain)):
    for it in range(2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        fo = R.rollout_forward(env_t, mlp_c, params, Kt, seed=it, n_steps_lim=4000, store_path=True, ckpt_every=1, want_logw=False)
        e[1].record()
        g = R.rollout_backward(env_t, mlp_c, params, fo, 1.0 / Kt)
        e[2].record(); torch.cuda.synchronize()
    u = float(fo.stats[L.ST_USEFUL_STEPS]); f_ms, b_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(json.dumps({"what": "loss + gradient", "H": H, "K": Kt, "fwd_ms": f_ms, "bwd_ms": b_ms, "max_T": int(fo.stats[L.ST_MAX_T]),
                      "train_steps_per_s": u / (f_ms + b_ms) * 1e3, "fp32_frac_train": u / (f_ms + b_ms) * 1e3 * F_train / 74.45e12}))
if a.ref:
    from oracle import ref_loader
    ref = ref_loader.load()
    renv = ref.environments.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005); renv.discretize_state_space(0.05)
    torch.manual_seed(1)
    rm = ref.core.DeterministicPolicy(1, 1, [H, H], nn.Tanh())
    cnt = ref_loader.UsefulStepCounter(renv, "step")
    t0 = time.perf_counter()
    ref.approx.test_policy_vectorized(renv, rm, batch_size=2000, k_max=1000, policy_opt=np.zeros((renv.n_states, 1)))
    w = time.perf_counter() - t0
    print(json.dumps({"what": "reference CPU test rollout", "H": H, "K": 2000, "s": w, "steps_per_s": cnt.useful / w, "threads": torch.get_num_threads()}))
    cnt.remove()
    rm.policy[4].bias.data.fill_(0.5)
    torch.manual_seed(0); t0 = time.perf_counter()
    loss, ret, steps = ref.core.sample_loss_vectorized(renv, rm, 1000); loss.backward()
    w = time.perf_counter() - t0
    print(json.dumps({"what": "reference CPU loss + gradient", "H": H, "K": 1000, "s": w, "train_steps_per_s": float(np.sum(steps)) / w}))
