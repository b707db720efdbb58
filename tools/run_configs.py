#!/usr/bin/env python
"""Run the BASELINE.json configs other than the bench headline and print one JSON line per config.

    python tools/run_configs.py                      # single GPU
    torchrun --nproc-per-node N tools/run_configs.py # trajectories sharded over N GPUs (configs 3 and 4)

  config 0  1-D REINFORCE, K = 100, lr 1e-2, seed 1 (CPU-runnable reference): iterations / s
  config 2  tabular tables h_state = h_action = 0.01 (see tools/bench_tables.py for the kernel-only timing)
  config 3  d = 10 double well, MLP 10-32-32-10 with head bias +3 (the uncontrolled d = 10 problem
            never hits), n_steps_lim = 2000, one REINFORCE loss + gradient + all-reduce; K per GPU via --k3
  config 4  metastable 1-D beta = 4, dt = 0.001, seed-1 initial policy, no step cap below 1e6; K per GPU via --k4
Sizes default to a few seconds of GPU time; the full sizes (1e7 / 1e8 trajectories over 8 GPUs) are weak-scaled
versions of the same launches.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L  # noqa: E402
from rl_sde_is_b200 import rollout as R  # noqa: E402
from rl_sde_is_b200.approximate_methods import is_estimate  # noqa: E402
from rl_sde_is_b200.distributed import Shard  # noqa: E402
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D, DoubleWellStoppingTimeND  # noqa: E402
from rl_sde_is_b200.models import DeterministicPolicy  # noqa: E402
from rl_sde_is_b200.reinforce_deterministic_core import reinforce, sample_loss_vectorized  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k3", type=int, default=200_000, help="config 3 trajectories per GPU")
    ap.add_argument("--k4", type=int, default=200_000, help="config 4 trajectories per GPU")
    ap.add_argument("--iters0", type=int, default=100)
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    skip = set(a.skip.split(","))

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    # ---------------------------------------------------------------- config 0
    if "0" not in skip and rank == 0:
        env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
        data = reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=a.iters0, seed=1, verbose=False, save=False, device=dev)
        cts = data["cts"]
        emit({"config": 0, "what": "REINFORCE K=100 lr=1e-2 seed=1", "iterations": a.iters0,
              "iter_per_s_it0": 1 / cts[0], "iter_per_s_it10_29": 1 / np.mean(cts[10:30]), "iter_per_s_last20": 1 / np.mean(cts[-20:]),
              "mean_steps_it0": data["exp_time_steps"][0], "mean_steps_last20": float(np.mean(data["exp_time_steps"][-20:])),
              "mean_return_it0": data["exp_returns"][0], "mean_return_last20": float(np.mean(data["exp_returns"][-20:])),
              "total_s": float(cts.sum()), "reference_cpu": "it.0 0.14 it/s, it.10-20 ~2.4 it/s (BASELINE.md)"})
    if world > 1:
        dist.barrier()

    # ---------------------------------------------------------------- config 3: d = 10, loss + gradient + all-reduce
    if "3" not in skip:
        d = 10
        env = DoubleWellStoppingTimeND(d, beta=1.0, alpha=1.0, dt=0.005)
        torch.manual_seed(1)
        model = DeterministicPolicy(d, d, [32, 32], nn.Tanh())
        model.policy[4].bias.data.fill_(3.0)
        K = a.k3
        shard = Shard(K * world, rank, world)
        lim = 2000
        times = []
        for it in range(3):
            model.zero_grad()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            loss, ret, steps = sample_loss_vectorized(env, model, K, seed=100 + it, n_steps_lim=lim, ckpt_every=16, device=dev,
                                                      dist=shard if world > 1 else None)
            t1 = time.perf_counter()
            loss.backward()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            times.append((t1 - t0, t2 - t1))
        useful = float(steps.sum())
        tt = torch.tensor([useful, times[-1][0], times[-1][1]], dtype=torch.float64, device=dev)
        if world > 1:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt)
            useful, fwd_s, bwd_s = float(tt[0]), float(mx[1]), float(mx[2])
        else:
            fwd_s, bwd_s = times[-1]
        gnorm = float(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).norm())
        F_train = 3 * 2 * (2 * d * 32 + 32 * 32) + 24 * d + 4
        emit({"config": 3, "what": "d=10 double well, MLP 10-32-32-10 (head bias +3), n_steps_lim 2000, loss+gradient+allreduce",
              "n_gpus": world, "K_global": K * world, "useful_steps": useful, "mean_steps": useful / (K * world), "loss": float(loss.detach()),
              "grad_norm": gnorm, "fwd_s": fwd_s, "bwd_s": bwd_s, "train_steps_per_s": useful / (fwd_s + bwd_s),
              "fwd_steps_per_s": useful / fwd_s, "bwd_steps_per_s": useful / bwd_s,
              "train_fp32_frac_per_gpu": useful / world / (fwd_s + bwd_s) * F_train / 74.45e12})

    # ---------------------------------------------------------------- config 4: metastable beta = 4
    if "4" not in skip:
        env = DoubleWellStoppingTime1D(beta=4.0, alpha=1.0, dt=0.001)
        torch.manual_seed(1)
        model = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
        K = a.k4
        shard = Shard(K * world, rank, world)
        for it in range(2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            s = is_estimate(env, model, K, n_steps_lim=10**6, seed=7 + it, device=dev, dist=shard if world > 1 else None)
            torch.cuda.synchronize()
            dt_wall = time.perf_counter() - t0
        tt = torch.tensor([dt_wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall = float(tt[0])
        emit({"config": 4, "what": "metastable 1-D beta=4 dt=0.001, seed-1 initial policy, n_steps_lim 1e6", "n_gpus": world,
              "K_global": s["n"], "useful_steps": s["useful_steps"], "mean_hit_index": s.get("mean_hit_index"),
              "max_hit_index": s["max_hit_index"], "n_unfinished": s["n_unfinished"], "is_mean": s.get("is_mean"),
              "is_rel_error": s.get("is_rel_error"), "wall_s": wall, "steps_per_s": s["useful_steps"] / wall,
              "fp32_frac_per_gpu": s["useful_steps"] / world / wall * 2194 / 74.45e12,
              "expect": "mean hitting pass ~6.95e4, E[exp(-tau)] ~0.0060 (BASELINE.md)"})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
