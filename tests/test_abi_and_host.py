"""CPU-only checks: the C-ABI library loads and exports every symbol include/rlsde.h declares, argument
validation works without a GPU, and the host-side logic (grids, policy container, sharding) is right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn as nn

from rl_sde_is_b200 import _lib as L
from rl_sde_is_b200 import distributed as D
from rl_sde_is_b200.models import DeterministicPolicy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rlsde.h")).read()
    declared = sorted(set(re.findall(r"\b(rlsde_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(L.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.load().rlsde_version() == 200


def test_argument_validation_without_gpu():
    lib = L.load()
    assert lib.rlsde_supported(1, 32, 2) == 1 and lib.rlsde_supported(1, 48, 2) == 0 and lib.rlsde_supported(1, 32, 3) == 0
    assert lib.rlsde_param_count(L.make_mlp(1, 32)) == 1153
    assert lib.rlsde_param_count(L.make_mlp(10, 32)) == 1738
    assert lib.rlsde_workspace_bytes(100) > 0
    assert lib.rlsde_strerror(-2).decode().startswith("no fused kernel")
    env = L.make_env(1, 1.0, np.sqrt(2.0), 0.005, 1.0, 2.0, [-1.0], L.HIT_ALL_GE_LB)
    cfg = L.RlsdeRolloutCfg()
    cfg.K, cfg.n_steps_lim = 8, 100
    # unsupported shape and null pointers are rejected before any CUDA call
    assert lib.rlsde_rollout_fwd(env, L.make_mlp(1, 48), None, cfg, *([None] * 9), None, 0, None) == -2
    assert lib.rlsde_rollout_fwd(env, L.make_mlp(1, 32), None, cfg, *([None] * 9), None, 0, None) == -1
    assert lib.rlsde_tables(None, 4, None, 4, None, 1, 1.0, 1.0, 0.1, 0.1, 1.0, 2.0, 0, 4, None, None, 0, 0.0, None) == -1
    assert lib.rlsde_noise_fill(0, 0, 4, 99, 0, 1, 0.1, None, None) == -1


def test_product_path_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    from rl_sde_is_b200.reinforce_deterministic_core import sample_loss_vectorized
    env = DoubleWellStoppingTime1D()
    model = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    with pytest.raises(L.RlsdeError):
        sample_loss_vectorized(env, model, 4)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "rl_sde_is_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_policy_container_layout_and_init():
    torch.manual_seed(3)
    m = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    sd = m.state_dict()
    assert list(sd) == ["policy.0.weight", "policy.0.bias", "policy.2.weight", "policy.2.bias", "policy.4.weight", "policy.4.bias"]
    assert [tuple(v.shape) for v in sd.values()] == [(32, 1), (32,), (32, 32), (32,), (1, 32), (1,)]
    assert sd["policy.4.weight"].abs().max() <= 5e-3 and sd["policy.4.bias"].abs().max() <= 5e-3


def test_policy_init_matches_reference_draws(golden):
    """Same torch seed -> the same parameters as the reference's DeterministicPolicy (fixture a_ of the torch rollouts
    was built with torch.manual_seed(3), head bias then set to 1.0)."""
    g = golden("rollout_torch_1d")
    torch.manual_seed(3)
    m = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    m.policy[4].bias.data.fill_(1.0)
    for k, v in m.state_dict().items():
        assert np.array_equal(v.numpy(), g["a_param." + k]), k


def test_grids_match_reference(golden):
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    g = golden("tables")
    for tag in ("h01", "b4", "h001"):
        alpha, beta, dt, hs, ha = g[tag + "_cfg"]
        env = DoubleWellStoppingTime1D(beta=beta, alpha=alpha, dt=dt)
        env.set_action_space_bounds()
        env.discretize_state_space(hs)
        env.discretize_action_space(ha)
        assert np.array_equal(env.state_space_h, g[tag + "_state_grid"])
        assert np.array_equal(env.action_space_h, g[tag + "_action_grid"])       # raw arange values, not rounded
        assert np.array_equal(env.is_in_ts, g[tag + "_is_in_ts"])
    assert env.n_states == 401 and env.n_actions == 601 and env.ts_idx.size == 101 and env.lb_idx == 300
    assert list(env.state_init_idx) == [100] and list(env.null_action_idx) == [300]


def test_state_index_lookup_is_within_one_cell():
    """The reference's own test of the discretisation (tests/test_environments.py:44-57)."""
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    env = DoubleWellStoppingTime1D()
    env.discretize_state_space(0.1)
    x = np.random.default_rng(0).uniform(-2, 2, (1000, 1))
    idx = env.get_state_idx(x)
    assert (np.abs(x[:, 0] - env.state_space_h[idx]) <= 0.1).all()


def test_shard_bounds_partition():
    for K, W in [(10, 3), (1000000, 8), (7, 8), (100, 1)]:
        edges = [D.shard_bounds(K, r, W) for r in range(W)]
        assert edges[0][0] == 0 and edges[-1][1] == K
        assert all(edges[i][1] == edges[i + 1][0] for i in range(W - 1))
        assert max(e - b for b, e in edges) - min(e - b for b, e in edges) <= 1


def test_ckpt_spacing_choice():
    from rl_sde_is_b200.rollout import choose_ckpt_every
    assert choose_ckpt_every(100, 1, 10**6) == 1
    assert choose_ckpt_every(10**4, 1, 10**6) == 8
    with pytest.raises(L.RlsdeError):
        choose_ckpt_every(10**7, 10, 10**6)


def _gloo_worker(rank, world, port, K_global, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sh = D.Shard.from_env(K_global)
    # each rank contributes the "gradient" and "statistics" of its shard; the packed all-reduce must give the
    # same result as a single rank holding the whole batch
    ids = torch.arange(sh.traj_offset, sh.traj_offset + sh.K_local, dtype=torch.float64)
    grad = torch.stack([ids.sum(), (ids ** 2).sum(), torch.tensor(float(sh.K_local), dtype=torch.float64)]).to(torch.float32)
    stats = torch.zeros(L.RLSDE_NSTATS, dtype=torch.float64)
    stats[L.ST_N] = sh.K_local
    stats[L.ST_SUM_G] = ids.sum()
    buf = sh.all_reduce_sum(D.pack_grad_and_stats(grad, stats))
    g, s = D.unpack_grad_and_stats(buf, 3)
    rec = stats.clone()
    rec[L.ST_MAX_T] = 100.0 + rank                  # the maximum entry must be max-reduced, everything else summed
    rec = sh.all_reduce_stats(rec)
    assert float(rec[L.ST_MAX_T]) == 100.0 + world - 1 and float(rec[L.ST_N]) == K_global
    # the training path's exchange: ONE all-gather of [gradient | statistics] rows, added in rank order by every rank
    stats2 = stats.clone()
    stats2[L.ST_MAX_T] = 7.0 + 3 * rank
    rows = sh.all_gather_rows(D.pack_grad_and_stats(grad, stats2))
    assert rows.shape == (world, 3 + L.RLSDE_NSTATS)
    g2, s2 = D.reduce_gathered(rows, 3)
    assert torch.equal(g2, g) and float(s2[L.ST_N]) == K_global and float(s2[L.ST_MAX_T]) == 7.0 + 3 * (world - 1)
    q.put((rank, sh.traj_offset, sh.K_local, g.tolist(), float(s[L.ST_N]), float(s[L.ST_SUM_G])))
    dist.destroy_process_group()


def test_packed_allreduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    K_global, port = 101, 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, K_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[2] for r in res] == [51, 50] and [r[1] for r in res] == [0, 51]
    ids = np.arange(K_global, dtype=np.float64)
    for _, _, _, g, n, sg in res:
        assert g[2] == K_global and abs(g[0] - ids.sum()) < 1e-3 and n == K_global and sg == ids.sum()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU (it is the CPU arm) and prints one JSON line with the contract's keys."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-trajectories", "200"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_tools_and_bench_compile():
    import glob, py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")] + glob.glob(os.path.join(root, "tools", "*.py")):
        py_compile.compile(f, doraise=True)


def test_table_row_class_indexing_covers_every_entry_exactly_once():
    """Index arithmetic of the quadrature table kernel (csrc/tables.cu: row residue classes, column tiles shifted to 32-byte
    boundaries, chunks capped in steps and in standard deviations), restated in Python: for many shapes, slabs and output
    alignments every (row, column) of the slab is produced exactly once, nothing outside, and every warp's first store address is a
    multiple of 32 bytes."""
    rng = np.random.default_rng(7)
    GL_ANCHOR, GL_SPAN_SD = 64, 21.0
    for trial in range(60):
        Ns, Na = int(rng.integers(2, 70)), int(rng.integers(1, 40))
        sp_begin = int(rng.integers(0, Ns))
        sp_end = int(rng.integers(sp_begin + 1, Ns + 1))
        base = int(rng.integers(0, 4))                     # output pointer in doubles modulo 4
        step_sd0 = float(rng.choice([0.05, 0.1, 0.2]))
        cols = Ns * Na
        Q = 1 if cols % 4 == 0 else (2 if cols % 2 == 0 else 4)
        rows_per_class = (Ns + Q - 1) // Q
        steps_cap = int(GL_SPAN_SD / (Q * step_sd0))
        steps_cap = max(1, min(GL_ANCHOR, steps_cap))
        n_chunks_full = (rows_per_class + steps_cap - 1) // steps_cap
        steps = (rows_per_class + n_chunks_full - 1) // n_chunks_full
        kc_lo = ((sp_begin - (Q - 1)) // Q if sp_begin >= Q else 0) // steps
        kc_hi = ((sp_end - 1) // Q) // steps
        grid_x, grid_y = (cols + 3 + 127) // 128, (kc_hi - kc_lo + 1) * Q
        hits = np.zeros((sp_end - sp_begin, cols), dtype=np.int32)
        for by in range(grid_y):
            j, kc = by % Q, kc_lo + by // Q
            rel = (((j - sp_begin) % 4 + 4) % 4) * (cols & 3)
            shift = (base + rel) & 3
            r0 = Q * kc * steps + j
            if r0 >= sp_end:
                continue
            i_lo = 0 if r0 >= sp_begin else (sp_begin - r0 + Q - 1) // Q
            i_hi = min(steps, (sp_end - r0 + Q - 1) // Q)
            if i_lo >= i_hi:
                continue
            rows = r0 + Q * np.arange(i_lo, i_hi)
            assert rows.min() >= sp_begin and rows.max() < sp_end
            for bx in range(grid_x):
                c = bx * 128 + np.arange(128) - shift
                for w in range(4):                         # first lane of each warp: address multiple of 4 doubles on every row
                    c0 = bx * 128 + 32 * w - shift
                    assert all((base + (r - sp_begin) * cols + c0) % 4 == 0 for r in rows)
                c = c[(c >= 0) & (c < cols)]
                hits[np.ix_(rows - sp_begin, c)] += 1
        assert (hits == 1).all(), (Ns, Na, sp_begin, sp_end, base)


def test_tcgen05_reverse_pass_work_assignment():
    """Static work assignment of the producer / consumer reverse kernel (csrc/rollout_umma_bwd*.cuh), restated: every
    128-trajectory tile goes to exactly one producer (boustrophedon deal), every producer is polled by exactly one consumer,
    and the grid fits the SMs (the launch is cooperative)."""
    UMMA_M, PPC = 128, 2
    for sm in (148, 132, 16, 3):
        for K in (1, 127, 128, 129, 1000, 9472, 60000, 10**6):
            tiles = (K + UMMA_M - 1) // UMMA_M
            n_prod = max(1, min(tiles, sm * PPC // (PPC + 1)))
            n_cons = (n_prod + PPC - 1) // PPC
            assert n_prod + n_cons <= max(sm, 2)
            seen = np.zeros(tiles, dtype=np.int32)
            for p in range(n_prod):
                rnd = 0
                while rnd * n_prod < tiles:
                    t = rnd * n_prod + ((n_prod - 1 - p) if (rnd & 1) else p)
                    if t < tiles:
                        seen[t] += 1
                    rnd += 1
            assert (seen == 1).all(), (sm, K)
            polled = np.zeros(n_prod, dtype=np.int32)
            for c in range(n_cons):
                for i in range(PPC):
                    if c + i * n_cons < n_prod:
                        polled[c + i * n_cons] += 1
            assert (polled == 1).all(), (sm, K)
