"""Two-rank check of the data-parallel paths, run under torchrun on 2 GPUs (tests/test_gpu_multi.py launches it):
trajectories sharded by global id, one exchange per call / per training iteration.  Every rank also computes the
unsharded result on its own GPU and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200 import _lib as L  # noqa: E402
from rl_sde_is_b200.approximate_methods import is_estimate, test_policy_vectorized  # noqa: E402
from rl_sde_is_b200.distributed import Shard  # noqa: E402
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D  # noqa: E402
from rl_sde_is_b200.models import DeterministicPolicy  # noqa: E402
from rl_sde_is_b200.reinforce_deterministic_core import _loss_and_grads_fused, reinforce  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(3)
    model = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(0.8)

    # ---- test rollout: statistics of the sharded batch = statistics of the whole batch
    K = 40001                                   # odd: shards of different sizes
    shard = Shard(K, rank, world)
    whole = is_estimate(env, model, K, n_steps_lim=5000, seed=11, device=dev)
    part = is_estimate(env, model, shard.K_local, n_steps_lim=5000, seed=11, device=dev, dist=shard)
    for key in ("n", "n_unfinished", "useful_steps", "max_hit_index", "mean_hit_index"):
        assert whole[key] == part[key], (key, whole[key], part[key])
    for key in ("mean_return", "var_return", "is_mean", "is_rel_error"):
        assert abs(whole[key] - part[key]) <= 1e-12 * abs(whole[key]), (key, whole[key], part[key])
    a = test_policy_vectorized(env, model, K, k_max=5000, seed=12, device=dev)
    b = test_policy_vectorized(env, model, shard.K_local, k_max=5000, seed=12, device=dev, dist=shard)
    assert a[2] == b[2] and abs(a[0] - b[0]) <= 1e-12 * abs(a[0])

    # ---- training evaluation, host route: ONE all-gather of [gradient | statistics] rows; global gradient on every rank
    Kt = 20000
    sh = Shard(Kt, rank, world)
    model.zero_grad()
    loss_w, ret_w, steps_w = _loss_and_grads_fused(env, model, Kt, seed=21, n_steps_lim=5000, device=dev)
    g_whole = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    model.zero_grad()
    loss_p, ret_p, steps_p = _loss_and_grads_fused(env, model, sh.K_local, seed=21, n_steps_lim=5000, device=dev, dist=sh)
    g_part = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
    lo = sh.traj_offset
    assert np.array_equal(ret_p, ret_w[lo:lo + sh.K_local]) and np.array_equal(steps_p, steps_w[lo:lo + sh.K_local])   # bitwise per trajectory
    assert abs(loss_p - loss_w) <= 2e-6 * abs(loss_w)
    scale = float(g_whole.abs().max())
    assert float((g_part - g_whole).abs().max()) <= 2e-5 * scale, float((g_part - g_whole).abs().max()) / scale
    g_dev = g_part.to(dev)
    gathered = [torch.empty_like(g_dev) for _ in range(world)]
    dist.all_gather(gathered, g_dev)
    assert all(torch.equal(gathered[0], t) for t in gathered)                       # every rank formed the same bits

    # ---- device-resident training loop, data parallel: identical parameters on both ranks, close to the single-GPU run
    kw = dict(d_hidden_layer=32, batch_size=128, lr=1e-2, n_iterations=8, seed=5, verbose=False, save=False, device=dev)
    single = reinforce(env, device_loop=True, **kw)
    multi = reinforce(env, device_loop=True, dist=Shard(128, rank, world), **kw)
    th = torch.cat([p.detach().reshape(-1) for p in multi["model"].parameters()]).to(dev)
    both = [torch.empty_like(th) for _ in range(world)]
    dist.all_gather(both, th)
    assert torch.equal(both[0], both[1])
    ts = torch.cat([p.detach().reshape(-1) for p in single["model"].parameters()]).to(dev)
    assert float((th - ts).abs().max()) < 5e-4, float((th - ts).abs().max())
    half = 64
    assert np.array_equal(multi["returns"][:half], single["returns"][rank * half:(rank + 1) * half])   # iteration 0, this rank's shard
    assert abs(multi["losses"][0] - single["losses"][0]) < 1e-5 * abs(single["losses"][0])
    host = reinforce(env, device_loop=False, dist=Shard(128, rank, world), **kw)
    tho = torch.cat([p.detach().reshape(-1) for p in host["model"].parameters()]).to(dev)
    assert float((tho - ts).abs().max()) < 5e-4
    dist.barrier()
    if rank == 0:
        print("MGPU_CHECK_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
