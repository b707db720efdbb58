#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by executing the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference and does not
travel to the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (riberaborrell/rl-sde-is v1.0.0) is imported from /root/reference/src with
three import stubs: a two-attribute ``rl_sde_is.config`` module and
MagicMock stand-ins for matplotlib / shapely, which are not installed and are not on the hot
path.  Nothing of the reference is copied: we only call its public functions and record
their inputs (the Brownian increments they draw) and outputs.

Fixtures (all arrays little-endian, names are the keys of the .npz):

  rollout_torch_1d.npz   reference ``sample_loss_vectorized`` (reinforce_deterministic_core.py:30-93)
                         + ``eff_loss.backward()`` on DoubleWellStoppingTime1D, noise recorded from
                         ``env.step_torch`` (environments.py:201-226).  Two cases (``a_*``, ``b_*``).
  rollout_torch_2d.npz   same on DoubleWellStoppingTime2D (environments_2d.py:184-205).
  rollout_numpy_1d.npz   reference ``test_policy_vectorized`` / ``estimate_fht_vectorized``
                         (approximate_methods.py:577-695), noise recorded from ``env.step``.
  rollout_numpy_2d.npz   same on the 2-D env.
  env_step.npz           single ``env.step`` / ``env.step_torch`` calls, both reward types.
  reinforce_iters.npz    three ``zero_grad -> loss -> backward -> Adam.step`` iterations
                         (reinforce_deterministic_core.py:234-243) with recorded noise.
  dp_sweeps.npz          ``q_table_update_vect`` / ``v_table_update_vect`` / ``policy_update_vect`` on the h = 0.1 tables
                         (tabular_dp_{qvalue,value,policy}_iteration.py), 3 sweeps at gamma = 1 and 0.97, and 200 sweeps.
  replay.npz             ``sample_trajectories_buffer_vectorized`` (approximate_methods.py:513-545) into a ``ReplayBuffer``
                         (replay_buffers.py:7-88), noise recorded from ``env.step``: 1-D (finished and n_max-capped) and 2-D.
  formats.json           run-directory names of ``utils_path`` (:100-330) for the argument sets the tests use, and the
                         key/dtype/shape listing of an ``agent.npz`` written by the reference's ``reinforce`` (:102-336).
  rollout_wide.npz       the torch and NumPy rollouts above for hidden widths 64, 128 and 256 (256 is the default of the
                         reference's ``reinforce``, reinforce_deterministic_core.py:102), noise recorded.
  rollout_long.npz       long / large cases whose noise is NOT stored: the reference draws it from ``torch.manual_seed(s)``
                         / ``np.random.seed(s)`` streams (environments.py:145,208), which the tests replay call by call
                         (``replay_torch_noise`` / ``replay_numpy_noise`` below): K = 256 with > 2 000-pass paths, and the
                         metastable beta = 4, dt = 0.001 environment of BASELINE config 5.
  tables.npz             ``compute_r_table`` / ``compute_p_tensor_batch`` (dynamic_programming.py:3-36):
                         full tensors at h=0.1, strided sub-sample + checksums at h=0.01, and a
                         (alpha, beta) = (1, 4) case.
"""
import hashlib
import os
import sys
import tempfile
import time
import types
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn

REF_SRC = os.environ.get("RLSDE_REFERENCE_SRC", "/root/reference/src")
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "shapely", "shapely.geometry"]:
        sys.modules.setdefault(name, MagicMock())
    tmp = tempfile.mkdtemp(prefix="rlsde_ref_")
    cfg = types.ModuleType("rl_sde_is.config")
    cfg.PROJECT_ROOT_DIR = tmp
    cfg.DATA_ROOT_DIR = os.path.join(tmp, "data")
    sys.modules["rl_sde_is.config"] = cfg
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import rl_sde_is.environments as env1
    import rl_sde_is.environments_2d as env2
    import rl_sde_is.reinforce_deterministic_core as core
    import rl_sde_is.approximate_methods as am
    import rl_sde_is.dynamic_programming as dp
    import rl_sde_is.tabular_dp_tables as tdt
    return env1, env2, core, am, dp, tdt


def flat_params(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def put_params(out, prefix, model):
    for k, v in flat_params(model).items():
        out[f"{prefix}param.{k}"] = v


class NoiseRecorder:
    """Wraps env.step / env.step_torch and records the 4th return value (dbt)."""

    def __init__(self, env, which):
        self.env, self.which, self.noise = env, which, []
        self.orig = getattr(env, which)

    def __enter__(self):
        def wrapped(*a, **kw):
            res = self.orig(*a, **kw)
            dbt = res[3]
            self.noise.append(dbt.detach().numpy().copy() if torch.is_tensor(dbt) else np.array(dbt, copy=True))
            return res
        setattr(self.env, self.which, wrapped)
        return self

    def __exit__(self, *exc):
        setattr(self.env, self.which, self.orig)

    def stacked(self):
        return np.stack(self.noise, axis=0).astype(np.float32)


def torch_rollout_case(core, env, model, K, seed, out, prefix):
    torch.manual_seed(seed)
    model.zero_grad()
    with NoiseRecorder(env, "step_torch") as rec:
        loss, return_fht, time_steps = core.sample_loss_vectorized(env, model, K)
    loss.backward()
    out[prefix + "noise"] = rec.stacked()
    out[prefix + "loss"] = np.array(loss.detach().numpy())
    out[prefix + "return_fht"] = return_fht.copy()
    out[prefix + "time_steps"] = time_steps.copy()
    put_params(out, prefix, model)
    for k, p in model.named_parameters():
        out[f"{prefix}grad.{k}"] = p.grad.detach().numpy().copy()
    out[prefix + "env"] = np.array([env.d, float(np.ravel(env.alpha)[0]), env.beta, env.dt], dtype=np.float64)
    print(f"  {prefix}: K={K} passes={len(rec.noise)} loss={float(loss):.10f} T={time_steps[:8]}")


def numpy_rollout_case(am, env, model, K, seed, policy_opt, out, prefix):
    np.random.seed(seed)
    with NoiseRecorder(env, "step") as rec:
        res = am.test_policy_vectorized(env, model, batch_size=K, policy_opt=policy_opt)
    out[prefix + "noise"] = rec.stacked()
    out[prefix + "result"] = np.array(res, dtype=np.float64)
    out[prefix + "policy_opt"] = policy_opt
    put_params(out, prefix, model)
    out[prefix + "env"] = np.array([env.d, float(np.ravel(env.alpha)[0]), env.beta, env.dt, env.h_state], dtype=np.float64)
    print(f"  {prefix}: K={K} passes={len(rec.noise)} result={res}")
    np.random.seed(seed)
    with NoiseRecorder(env, "step") as rec2:
        fht = am.estimate_fht_vectorized(env, model, batch_size=K)
    assert np.array_equal(rec2.stacked(), out[prefix + "noise"])
    out[prefix + "fht"] = np.array(fht, dtype=np.float64)


def replay_torch_noise(seed, n_pass, K, d, dt):
    """The increments ``env.step_torch`` draws (environments.py:208) after ``torch.manual_seed(seed)``: one
    ``torch.randn((K, d))`` per pass, times ``sqrt(dt_tensor)`` in float32.  Shared by the tests (imported from here)."""
    torch.manual_seed(seed)
    sq = torch.sqrt(torch.tensor(dt, dtype=torch.float32))
    return torch.stack([sq * torch.randn((K, d), dtype=torch.float32) for _ in range(n_pass)]).numpy()


def replay_numpy_noise(seed, n_pass, K, d, dt):
    """The increments ``env.step`` draws (environments.py:145) after ``np.random.seed(seed)``."""
    np.random.seed(seed)
    return np.stack([np.array(np.sqrt(dt) * np.random.randn(K, d), dtype=np.float32) for _ in range(n_pass)])


def wide_cases(env1, env2, core, am):
    out = {}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.discretize_state_space(0.05)
    for tag, H, K, seed in (("h64_", 64, 8, 31), ("h128_", 128, 6, 32), ("h256_", 256, 6, 33)):
        torch.manual_seed(100 + H)
        model = core.DeterministicPolicy(1, 1, [H, H], nn.Tanh())
        model.policy[4].bias.data.fill_(1.0)
        torch_rollout_case(core, env, model, K, seed, out, "t" + tag)
        policy_opt = (1.5 * np.cos(env.state_space_h) + 0.25).reshape(-1, 1)
        numpy_rollout_case(am, env, model, K + 3, seed + 50, policy_opt, out, "n" + tag)
    envd2 = env2.DoubleWellStoppingTime2D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(7)
    model = core.DeterministicPolicy(2, 2, [128, 128], nn.Tanh())
    model.policy[4].bias.data.fill_(1.5)
    torch_rollout_case(core, envd2, model, 5, 34, out, "t2d_h128_")
    # the reference's reinforce() default shape at its own initialisation (near-null control: long paths), H = 256
    torch.manual_seed(1)
    model = core.DeterministicPolicy(1, 1, [256, 256], nn.Tanh())
    torch_rollout_case(core, env, model, 4, 35, out, "tinit_h256_")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_wide.npz"), **out)
    print("wrote rollout_wide.npz")


def long_cases(env1, env2, core, am):
    """Cases too long to store their noise: the tests regenerate it from the seed (see replay_*_noise)."""
    out = {}

    def torch_case(env, model, K, seed, prefix):
        torch.manual_seed(seed)
        model.zero_grad()
        with NoiseRecorder(env, "step_torch") as rec:
            loss, return_fht, time_steps = core.sample_loss_vectorized(env, model, K)
        loss.backward()
        noise = rec.stacked()
        assert np.array_equal(noise, replay_torch_noise(seed, noise.shape[0], K, env.d, env.dt)), "torch noise replay differs"
        out[prefix + "seed"] = np.array([seed, K, noise.shape[0]], dtype=np.int64)
        out[prefix + "loss"] = np.array(loss.detach().numpy())
        out[prefix + "return_fht"] = return_fht.copy()
        out[prefix + "time_steps"] = time_steps.copy()
        put_params(out, prefix, model)
        for k, p in model.named_parameters():
            out[f"{prefix}grad.{k}"] = p.grad.detach().numpy().copy()
        out[prefix + "env"] = np.array([env.d, float(np.ravel(env.alpha)[0]), env.beta, env.dt], dtype=np.float64)
        print(f"  {prefix}: K={K} passes={noise.shape[0]} loss={float(loss):.8f} max T={time_steps.max()}")

    def numpy_case(env, model, K, seed, policy_opt, prefix):
        np.random.seed(seed)
        with NoiseRecorder(env, "step") as rec:
            res = am.test_policy_vectorized(env, model, batch_size=K, policy_opt=policy_opt)
        noise = rec.stacked()
        assert np.array_equal(noise, replay_numpy_noise(seed, noise.shape[0], K, env.d, env.dt)), "numpy noise replay differs"
        out[prefix + "seed"] = np.array([seed, K, noise.shape[0]], dtype=np.int64)
        out[prefix + "result"] = np.array(res, dtype=np.float64)
        out[prefix + "policy_opt"] = policy_opt
        put_params(out, prefix, model)
        out[prefix + "env"] = np.array([env.d, float(np.ravel(env.alpha)[0]), env.beta, env.dt, env.h_state], dtype=np.float64)
        print(f"  {prefix}: K={K} passes={noise.shape[0]} result={res}")

    # K = 256, near-null initial policy (seed 1): the longest of 256 uncontrolled paths runs for thousands of passes
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.discretize_state_space(0.05)
    torch.manual_seed(1)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    torch_case(env, model, 256, 41, "t256_")
    numpy_case(env, model, 256, 42, np.zeros((env.n_states, 1)), "n256_")
    # metastable environment of config 5 (beta = 4, dt = 0.001) under a moderate constant push (the uncontrolled mean is
    # 7e4 passes: too long for a CPU fixture)
    envm = env1.DoubleWellStoppingTime1D(beta=4.0, alpha=1.0, dt=0.001)
    envm.discretize_state_space(0.05)
    torch.manual_seed(2)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.2)
    torch_case(envm, model, 48, 43, "tmeta_")
    numpy_case(envm, model, 48, 44, np.zeros((envm.n_states, 1)), "nmeta_")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_long.npz"), **out)
    print("wrote rollout_long.npz")


def dp_sweeps(env1, dp):
    """dp_sweeps.npz: reference Bellman sweeps over the h = 0.1 tables (tabular_dp_qvalue_iteration.py:35-43,
    tabular_dp_value_iteration.py:41-52, tabular_dp_policy_iteration.py:37-49) -- SURVEY 8f-1."""
    import rl_sde_is.tabular_dp_policy_iteration as pit
    import rl_sde_is.tabular_dp_qvalue_iteration as qit
    import rl_sde_is.tabular_dp_value_iteration as vit
    out = {}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.set_action_space_bounds()
    env.discretize_state_space(0.1)
    env.discretize_action_space(0.1)
    R = dp.compute_r_table(env)
    P = dp.compute_p_tensor_batch(env)
    rng = np.random.default_rng(42)
    q0 = -rng.random((env.n_states, env.n_actions))
    v0 = -rng.random(env.n_states)
    out["q0"], out["v0"], out["gamma"] = q0, v0, np.array([1.0, 0.97])
    for gi, gamma in enumerate((1.0, 0.97)):
        q = q0.copy()
        for it in range(3):
            q = qit.q_table_update_vect(env, R, P, q, gamma)
            out[f"g{gi}_q{it + 1}"] = q
        out[f"g{gi}_v1"] = vit.v_table_update_vect(env, R, P, v0, gamma)
        out[f"g{gi}_pi1"] = pit.policy_update_vect(env, R, P, v0, gamma)
    q = q0.copy()
    for it in range(200):
        q = qit.q_table_update_vect(env, R, P, q, 1.0)
    out["q200"] = q
    out["null_action_idx"] = np.array(env.null_action_idx)
    np.savez_compressed(os.path.join(OUT_DIR, "dp_sweeps.npz"), **out)
    print("  dp_sweeps: V(s_init) after 200 sweeps =", np.max(q[env.state_init_idx]))


def replay_cases(env1, env2, core, am):
    """replay.npz: buffer contents after the reference's vectorised sampler, with the increments it drew."""
    from rl_sde_is.replay_buffers import ReplayBuffer
    out = {}

    def case(prefix, env, model, K, n_max, seed):
        buf = ReplayBuffer(size=K * n_max, state_dim=env.d, action_dim=env.d)
        np.random.seed(seed)
        with NoiseRecorder(env, "step") as rec:
            am.sample_trajectories_buffer_vectorized(env, model, buf, K, n_max)
        n = buf.size
        out[prefix + "noise"] = rec.stacked()
        for name in ("states", "actions", "rewards", "next_states", "done"):
            out[prefix + name] = getattr(buf, name)[:n].copy()
        out[prefix + "cfg"] = np.array([K, n_max, n, buf.ptr], dtype=np.int64)
        put_params(out, prefix, model)
        out[prefix + "env"] = np.array([env.d, float(np.ravel(env.alpha)[0]), env.beta, env.dt], dtype=np.float64)
        print(f"  replay {prefix}: K={K} n_max={n_max} passes={len(rec.noise)} transitions={n} done={int(buf.done[:n].sum())}")

    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(3)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.0)
    case("a_", env, model, 12, 5000, 41)             # every episode reaches the target set
    torch.manual_seed(1)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    case("b_", env, model, 7, 150, 42)               # near-null policy, capped at n_max: most episodes unfinished
    env = env2.DoubleWellStoppingTime2D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(4)
    model = core.DeterministicPolicy(2, 2, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.5)
    case("c_", env, model, 5, 5000, 43)
    np.savez_compressed(os.path.join(OUT_DIR, "replay.npz"), **out)


def format_cases(env1, core, dp):
    """formats.json: directory names and the agent.npz listing produced by the reference's own I/O layer."""
    import json
    import rl_sde_is.utils_path as up
    res = {"dirs": [], "agent_npz": {}}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env4 = env1.DoubleWellStoppingTime1D(beta=4.0, alpha=1.0, dt=0.001)
    for e, tag in ((env, "b1"), (env4, "b4")):
        kw = dict(agent="reinforce-deterministic", gamma=1.0, d_hidden_layer=32, batch_size=100, lr=1e-2,
                  n_iterations=1000, seed=1)
        res["dirs"].append(dict(fn="get_reinforce_det_dir_path", env=tag, kwargs=kw, path=up.get_reinforce_det_dir_path(e, **kw)))
        kw = dict(agent="reinforce-deterministic", gamma=0.99, d_hidden_layer=256, batch_size=1000, lr=1e-3,
                  n_iterations=100, seed=None)
        res["dirs"].append(dict(fn="get_reinforce_det_dir_path", env=tag, kwargs=kw, path=up.get_reinforce_det_dir_path(e, **kw)))
        e.set_action_space_bounds()
        e.discretize_state_space(0.01)
        e.discretize_action_space(0.01)
        res["dirs"].append(dict(fn="get_dynamic_programming_tables_dir_path", env=tag, kwargs={},
                                path=up.get_dynamic_programming_tables_dir_path(e)))
        kw = dict(agent="dp-q-value-iteration", n_iterations=2000)
        res["dirs"].append(dict(fn="get_dynamic_programming_dir_path", env=tag, kwargs=kw,
                                path=up.get_dynamic_programming_dir_path(e, **kw)))
    data = core.reinforce(env, d_hidden_layer=32, batch_size=10, lr=1e-2, n_iterations=4, seed=1, backup_freq_iterations=2,
                          policy_opt=np.zeros((env.n_states if hasattr(env, "n_states") else 401, 1)))
    rel = data["rel_dir_path"]
    root = os.path.join(up.get_data_dir(), rel)
    res["agent_npz"]["rel_dir_path"] = rel
    res["agent_npz"]["files"] = sorted(os.listdir(root))
    loaded = up.load_data(rel)
    listing = {}
    for k, v in loaded.items():
        if isinstance(v, np.ndarray):
            listing[k] = ["ndarray", str(v.dtype), list(v.shape)]
        else:
            listing[k] = [type(v).__name__]
    res["agent_npz"]["keys"] = listing
    sd = torch.load(os.path.join(root, "model_n-it2"))
    res["agent_npz"]["state_dict"] = {k: list(v.shape) for k, v in sd.items()}
    with open(os.path.join(OUT_DIR, "formats.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True, default=str)
    print("  formats:", res["agent_npz"]["files"], [d["path"] for d in res["dirs"]][:3])


def main():
    t_start = time.time()
    env1, env2, core, am, dp, tdt = import_reference()
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "replay":  # regenerate only the replay-buffer fixture
        replay_cases(env1, env2, core, am)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "formats":
        format_cases(env1, core, dp)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dp":      # regenerate only the DP-sweep fixture
        dp_sweeps(env1, dp)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wide":    # only the wide-policy fixture
        wide_cases(env1, env2, core, am)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "long":    # only the long-path fixture
        long_cases(env1, env2, core, am)
        return

    # ---------------------------------------------------------------- torch path, 1-D
    out = {}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(3)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.0)
    torch_rollout_case(core, env, model, 8, 7, out, "a_")        # SURVEY Appendix D recipe
    assert abs(float(out["a_loss"]) - 0.7900703549385071) < 1e-12, float(out["a_loss"])
    torch.manual_seed(1)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())  # near-null initial policy: long paths
    torch_rollout_case(core, env, model, 6, 11, out, "b_")
    envb4 = env1.DoubleWellStoppingTime1D(beta=2.0, alpha=0.5, dt=0.01)
    torch.manual_seed(5)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(0.5)
    torch_rollout_case(core, envb4, model, 12, 13, out, "c_")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_torch_1d.npz"), **out)

    # ---------------------------------------------------------------- torch path, 2-D
    out = {}
    env = env2.DoubleWellStoppingTime2D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(4)
    model = core.DeterministicPolicy(2, 2, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.5)
    torch_rollout_case(core, env, model, 6, 9, out, "a_")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_torch_2d.npz"), **out)

    # ---------------------------------------------------------------- numpy path, 1-D / 2-D
    out = {}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.discretize_state_space(0.05)
    policy_opt = (1.5 * np.cos(env.state_space_h) + 0.25).reshape(-1, 1)      # any table: only its lookup is tested
    torch.manual_seed(3)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.0)
    numpy_rollout_case(am, env, model, 16, 21, policy_opt, out, "a_")
    torch.manual_seed(1)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    numpy_rollout_case(am, env, model, 5, 22, np.zeros((env.n_states, 1)), out, "b_")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_numpy_1d.npz"), **out)

    # ---------------------------------------------------------------- single env steps
    out = {}
    rng = np.random.default_rng(0)
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    states = rng.uniform(-2, 2.2, (64, 1)).astype(np.float32)
    actions = rng.uniform(-3, 3, (64, 1)).astype(np.float32)
    out["states1"], out["actions1"] = states, actions
    for rt in ["state-action", "state-action-next-state"]:
        tag = rt.replace("-", "_")
        np.random.seed(5)
        ns, r, done, dbt = env.step(states, actions, reward_type=rt)
        out[f"np1_{tag}_next"], out[f"np1_{tag}_r"], out[f"np1_{tag}_done"], out[f"np1_{tag}_dbt"] = ns, r, done, dbt
        torch.manual_seed(5)
        ns, r, done, dbt = env.step_torch(torch.from_numpy(states), torch.from_numpy(actions), reward_type=rt)
        out[f"th1_{tag}_next"], out[f"th1_{tag}_r"], out[f"th1_{tag}_done"], out[f"th1_{tag}_dbt"] = \
            ns.numpy(), r.numpy(), done.numpy(), dbt.numpy()
    env = env2.DoubleWellStoppingTime2D(beta=1.0, alpha=1.0, dt=0.005)
    states = rng.uniform(-2, 2.2, (64, 2)).astype(np.float32)
    states[:8] = np.abs(states[:8]) + 0.9
    actions = rng.uniform(-3, 3, (64, 2)).astype(np.float32)
    out["states2"], out["actions2"] = states, actions
    np.random.seed(6)
    ns, r, done, dbt = env.step(states, actions)
    out["np2_next"], out["np2_r"], out["np2_done"], out["np2_dbt"] = ns, r, done, dbt
    torch.manual_seed(6)
    ns, r, done, dbt = env.step_torch(torch.from_numpy(states), torch.from_numpy(actions))
    out["th2_next"], out["th2_r"], out["th2_done"], out["th2_dbt"] = ns.numpy(), r.numpy(), done.numpy(), dbt.numpy()
    np.savez_compressed(os.path.join(OUT_DIR, "env_step.npz"), **out)

    # ---------------------------------------------------------------- REINFORCE iterations
    out = {}
    env = env1.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    torch.manual_seed(3)
    model = core.DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.0)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    put_params(out, "it0_", model)
    torch.manual_seed(17)
    for it in range(3):
        opt.zero_grad()
        with NoiseRecorder(env, "step_torch") as rec:
            loss, ret, ts = core.sample_loss_vectorized(env, model, 8)
        loss.backward()
        opt.step()
        out[f"it{it}_noise"] = rec.stacked()
        out[f"it{it}_loss"] = np.array(loss.detach().numpy())
        out[f"it{it}_return_fht"] = ret.copy()
        out[f"it{it}_time_steps"] = ts.copy()
        put_params(out, f"it{it + 1}_", model)
        print(f"  reinforce it{it}: loss={float(loss):.8f} passes={len(rec.noise)}")
    np.savez_compressed(os.path.join(OUT_DIR, "reinforce_iters.npz"), **out)

    # ---------------------------------------------------------------- tables
    out = {}

    def table_case(tag, alpha, beta, h_state, h_action, full, stride=None):
        env = env1.DoubleWellStoppingTime1D(beta=beta, alpha=alpha, dt=0.005)
        env.set_action_space_bounds()
        env.discretize_state_space(h_state)
        env.discretize_action_space(h_action)
        t0 = time.time()
        R = dp.compute_r_table(env)
        P = dp.compute_p_tensor_batch(env)
        ok = tdt.check_p_tensor(env, P)
        print(f"  tables {tag}: P{P.shape} R{R.shape} check={ok} {time.time() - t0:.1f}s")
        out[f"{tag}_cfg"] = np.array([alpha, beta, env.dt, h_state, h_action], dtype=np.float64)
        out[f"{tag}_state_grid"] = env.state_space_h
        out[f"{tag}_action_grid"] = env.action_space_h
        out[f"{tag}_is_in_ts"] = env.is_in_ts
        out[f"{tag}_check"] = np.array(bool(ok))
        out[f"{tag}_R"] = R
        out[f"{tag}_P_sum"] = np.array(P.sum())
        out[f"{tag}_P_sumsq"] = np.array((P * P).sum())
        out[f"{tag}_P_colsum_maxdev"] = np.array(np.abs(P.sum(axis=0) - 1).max())
        out[f"{tag}_P_sha256"] = np.frombuffer(hashlib.sha256(P.tobytes()).digest(), dtype=np.uint8)
        out[f"{tag}_R_sha256"] = np.frombuffer(hashlib.sha256(R.tobytes()).digest(), dtype=np.uint8)
        if full:
            out[f"{tag}_P"] = P
        else:
            out[f"{tag}_P_stride"] = np.array(stride)
            out[f"{tag}_P_sub"] = P[::stride[0], ::stride[1], ::stride[2]].copy()
            out[f"{tag}_P_s100"] = P[:, 100, :].copy()      # all next-states/actions from the initial state x=-1

    table_case("h01", 1.0, 1.0, 0.1, 0.1, full=True)
    assert hashlib.sha256(out["h01_P"].tobytes()).hexdigest()[:16] == "53031080f5d0ce25"
    table_case("b4", 1.0, 4.0, 0.1, 0.5, full=True)
    table_case("h001", 1.0, 1.0, 0.01, 0.01, full=False, stride=(10, 10, 20))
    np.savez_compressed(os.path.join(OUT_DIR, "tables.npz"), **out)

    # 2-D numpy path last (long hitting times uncontrolled; use a drifting policy)
    out = {}
    env = env2.DoubleWellStoppingTime2D(beta=1.0, alpha=1.0, dt=0.005)
    env.h_state = 0.1
    torch.manual_seed(4)
    model = core.DeterministicPolicy(2, 2, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(1.5)
    np.random.seed(31)
    with NoiseRecorder(env, "step") as rec:
        fht = am.estimate_fht_vectorized(env, model, batch_size=6)
    out["a_noise"] = rec.stacked()
    out["a_fht"] = np.array(fht, dtype=np.float64)
    put_params(out, "a_", model)
    out["a_env"] = np.array([env.d, 1.0, env.beta, env.dt], dtype=np.float64)
    print(f"  numpy 2d fht={fht} passes={len(rec.noise)}")
    np.savez_compressed(os.path.join(OUT_DIR, "rollout_numpy_2d.npz"), **out)

    dp_sweeps(env1, dp)
    replay_cases(env1, env2, core, am)
    format_cases(env1, core, dp)

    for f in sorted(os.listdir(OUT_DIR)):
        if f.endswith(".npz"):
            print(f"{f}: {os.path.getsize(os.path.join(OUT_DIR, f)) / 1024:.1f} KiB")
    print(f"done in {time.time() - t_start:.1f}s  (numpy {np.__version__}, torch {torch.__version__})")


if __name__ == "__main__":
    main()
