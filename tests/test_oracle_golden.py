"""Pin the oracle to the reference: every restated function reproduces, bit for bit where the
arithmetic is the same and to stated tolerances otherwise, the fixtures recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import reference_semantics as ref


def _env(npz, p):
    e = npz[p + "env"]
    return int(e[0]), float(e[1]), float(e[2]), float(e[3])


@pytest.mark.parametrize("fixture,prefix", [("rollout_torch_1d", "a_"), ("rollout_torch_1d", "b_"),
                                            ("rollout_torch_1d", "c_"), ("rollout_torch_2d", "a_")])
def test_torch_rollout_matches_reference(golden, fixture, prefix):
    torch.set_num_threads(1)
    g = golden(fixture)
    d, alpha, beta, dt = _env(g, prefix)
    out = ref.rollout_loss_torch(d, alpha, beta, dt, ref.params_from_npz(g, prefix), g[prefix + "noise"])
    assert out["all_hit"]
    assert np.array_equal(out["time_steps"], g[prefix + "time_steps"].astype(np.int64))
    assert np.array_equal(out["return_fht"], g[prefix + "return_fht"])          # bit-exact
    assert out["loss"] == g[prefix + "loss"]                                     # bit-exact
    for k in ref.PARAM_KEYS:                                                     # autograd graph differs in op order only
        np.testing.assert_allclose(out["grads"][k], g[f"{prefix}grad.{k}"], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("prefix,with_l2", [("a_", True), ("b_", True)])
def test_numpy_rollout_matches_reference(golden, prefix, with_l2):
    g = golden("rollout_numpy_1d")
    e = g[prefix + "env"]
    d, alpha, beta, dt, h = int(e[0]), float(e[1]), float(e[2]), float(e[3]), float(e[4])
    st = ref.rollout_stats_numpy(d, alpha, beta, dt, ref.params_from_npz(g, prefix), g[prefix + "noise"],
                                 policy_opt=g[prefix + "policy_opt"], h_state=h)
    res = ref.test_policy_result(st, with_l2)
    assert np.array_equal(np.array(res), g[prefix + "result"])                   # bit-exact
    assert dt * np.mean(st["ep_lens"]) == float(g[prefix + "fht"])


def test_numpy_rollout_2d_matches_reference(golden):
    g = golden("rollout_numpy_2d")
    d, alpha, beta, dt = _env(g, "a_")
    st = ref.rollout_stats_numpy(d, alpha, beta, dt, ref.params_from_npz(g, "a_"), g["a_noise"])
    assert st["all_hit"] and dt * np.mean(st["ep_lens"]) == float(g["a_fht"])


def test_env_step_matches_reference(golden):
    g = golden("env_step")
    for rt in ("state-action", "state-action-next-state"):
        tag = rt.replace("-", "_")
        nxt, r, done = ref.env_step_numpy(1, 1.0, 1.0, 0.005, g["states1"], g["actions1"], g[f"np1_{tag}_dbt"], rt)
        assert np.array_equal(nxt, g[f"np1_{tag}_next"]) and np.array_equal(r, g[f"np1_{tag}_r"])
        assert np.array_equal(done, g[f"np1_{tag}_done"])
        nxt, r, done = ref.env_step_torch(1, 1.0, 1.0, 0.005, torch.from_numpy(g["states1"]), torch.from_numpy(g["actions1"]),
                                          torch.from_numpy(g[f"th1_{tag}_dbt"]), rt)
        assert np.array_equal(nxt.numpy(), g[f"th1_{tag}_next"]) and np.array_equal(r.numpy(), g[f"th1_{tag}_r"])
        assert np.array_equal(done.numpy(), g[f"th1_{tag}_done"])
    nxt, r, done = ref.env_step_numpy(2, 1.0, 1.0, 0.005, g["states2"], g["actions2"], g["np2_dbt"])
    assert np.array_equal(nxt, g["np2_next"]) and np.array_equal(r, g["np2_r"]) and np.array_equal(done, g["np2_done"])
    nxt, r, done = ref.env_step_torch(2, 1.0, 1.0, 0.005, torch.from_numpy(g["states2"]), torch.from_numpy(g["actions2"]),
                                      torch.from_numpy(g["th2_dbt"]))
    assert np.array_equal(nxt.numpy(), g["th2_next"]) and np.array_equal(done.numpy(), g["th2_done"])
    np.testing.assert_allclose(r.numpy(), g["th2_r"], rtol=2e-7)     # sum-of-squares order inside torch.linalg.norm


@pytest.mark.parametrize("tag", ["h01", "b4"])
def test_tables_match_reference_bitwise(golden, tag):
    g = golden("tables")
    alpha, beta, dt, hs, _ = g[tag + "_cfg"]
    P = ref.p_tensor(g[tag + "_state_grid"], g[tag + "_action_grid"], g[tag + "_is_in_ts"], alpha, beta, dt, hs)
    R = ref.r_table(g[tag + "_state_grid"], g[tag + "_action_grid"], g[tag + "_is_in_ts"], dt)
    assert hashlib.sha256(P.tobytes()).digest() == g[tag + "_P_sha256"].tobytes()
    assert hashlib.sha256(R.tobytes()).digest() == g[tag + "_R_sha256"].tobytes()
    assert np.array_equal(P, g[tag + "_P"])


def test_tables_h001_subsample_matches_reference(golden):
    g = golden("tables")
    alpha, beta, dt, hs, _ = g["h001_cfg"]
    st = g["h001_P_stride"]
    sg, ag, ts = g["h001_state_grid"], g["h001_action_grid"], g["h001_is_in_ts"]
    sub = ref.p_tensor(sg, ag, ts, alpha, beta, dt, hs, sprime=np.arange(0, sg.size, st[0]),
                       s_idx=np.arange(0, sg.size, st[1]), a_idx=np.arange(0, ag.size, st[2]))
    assert np.array_equal(sub, g["h001_P_sub"])
    col = ref.p_tensor(sg, ag, ts, alpha, beta, dt, hs, s_idx=[100])[:, 0, :]
    assert np.array_equal(col, g["h001_P_s100"])
    assert hashlib.sha256(ref.r_table(sg, ag, ts, dt).tobytes()).digest() == g["h001_R_sha256"].tobytes()


# ------------------------------------------------------------------ C restatement
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert list(c_oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    ff = 0xffffffff
    assert list(c_oracle.philox([ff] * 4, [ff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(c_oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_c_noise_is_standard_normal_scaled():
    z = c_oracle.noise_fill(5, 4096, 1, 64, 0.005).ravel() / np.sqrt(0.005)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01 and abs((z ** 4).mean() - 3) < 0.1
    z2 = c_oracle.noise_fill(5, 1000, 3, 8, 0.01, traj_offset=7)
    z3 = c_oracle.noise_fill(5, 10, 3, 2, 0.01, traj_offset=17, pass_begin=3)
    assert np.array_equal(z2[3:5, 10:20], z3)                       # pure function of (seed, trajectory, pass)


@pytest.mark.parametrize("fixture,prefix", [("rollout_torch_1d", "a_"), ("rollout_torch_1d", "c_"), ("rollout_torch_2d", "a_")])
def test_c_rollout_f32_matches_python_restatement(golden, fixture, prefix):
    g = golden(fixture)
    d, alpha, beta, dt = _env(g, prefix)
    params = ref.params_from_npz(g, prefix)
    py = ref.rollout_loss_torch(d, alpha, beta, dt, params, g[prefix + "noise"], need_grad=False)
    K = g[prefix + "noise"].shape[1]
    c = c_oracle.rollout(d, 32, ref.flatten_params(params), alpha, beta, dt, K, noise=g[prefix + "noise"])
    # same semantics; the GEMV/tanh rounding differs from torch's (fma chain vs MKL, tanhf vs Sleef)
    assert np.array_equal(c["T"] + 1, py["time_steps"])
    np.testing.assert_allclose(c["G"], py["return_fht"], rtol=1e-5)
    np.testing.assert_allclose(c["S"], py["stoch_int_fht"], rtol=1e-4, atol=1e-5)


def test_c_rollout_f64_matches_python_restatement(golden):
    g = golden("rollout_numpy_1d")
    e = g["a_env"]
    params = ref.params_from_npz(g, "a_")
    st = ref.rollout_stats_numpy(1, float(e[1]), float(e[2]), float(e[3]), params, g["a_noise"], g["a_policy_opt"], float(e[4]))
    c = c_oracle.rollout(1, 32, ref.flatten_params(params), float(e[1]), float(e[2]), float(e[3]), g["a_noise"].shape[1],
                         noise=g["a_noise"], state_f64=True, hit_rule=c_oracle.HIT_X0_IN_LB_RB,
                         policy_opt=g["a_policy_opt"], grid=(-2.0, 2.0, float(e[4])))
    assert np.array_equal(c["T"], st["ep_lens"])
    np.testing.assert_allclose(c["G"], st["ep_rets"], rtol=1e-6)
    np.testing.assert_allclose(c["l2"], st["l2"], rtol=1e-4)


def test_c_tables_match_reference(golden):
    g = golden("tables")
    alpha, beta, dt, hs, _ = g["h01_cfg"]
    P, R = c_oracle.tables(g["h01_state_grid"], g["h01_action_grid"], g["h01_is_in_ts"], alpha, beta, dt, hs)
    np.testing.assert_allclose(P, g["h01_P"], rtol=0, atol=1e-15)   # glibc erf/erfc vs cephes ndtr
    assert np.array_equal(R, g["h01_R"])


def test_dp_sweeps_match_reference(golden):
    g, t = golden("dp_sweeps"), golden("tables")
    R, P, ts = t["h01_R"], t["h01_P"], t["h01_is_in_ts"]
    for gi, gamma in enumerate(g["gamma"]):
        q = g["q0"].copy()
        for it in range(3):
            q = ref.q_sweep(R, P, ts, q, gamma)
            np.testing.assert_allclose(q, g[f"g{gi}_q{it + 1}"], rtol=0, atol=1e-13)      # BLAS vs einsum summation order
        np.testing.assert_allclose(ref.v_sweep(R, P, ts, g["v0"], gamma), g[f"g{gi}_v1"], rtol=0, atol=1e-13)
        assert np.array_equal(ref.greedy_policy_indices(R, P, ts, g["v0"], gamma, int(g["null_action_idx"][0])), g[f"g{gi}_pi1"])


@pytest.mark.parametrize("prefix", ["a_", "b_", "c_"])
def test_replay_transitions_match_reference(golden, prefix):
    """oracle transitions_numpy == buffer contents written by the reference's sample_trajectories_buffer_vectorized."""
    g = golden("replay")
    d, alpha, beta, dt = _env(g, prefix)
    K, n_max, n, ptr = (int(v) for v in g[prefix + "cfg"])
    out = ref.transitions_numpy(d, alpha, beta, dt, ref.params_from_npz(g, prefix), g[prefix + "noise"], n_max)
    assert out["rewards"].shape[0] == n == ptr
    for name in ("states", "actions", "rewards", "next_states", "done"):
        want = g[prefix + name]
        assert out[name].dtype == want.dtype and out[name].shape == want.shape, name
        assert np.array_equal(out[name], want), name
    assert np.array_equal(np.signbit(out["rewards"]), np.signbit(g[prefix + "rewards"]))      # the -0.0 of detecting passes


# ------------------------------------------------------------------------------ round 2 fixtures: wide policies, long replays
@pytest.mark.parametrize("prefix", ["th64_", "th256_", "t2d_h128_"])
def test_wide_torch_rollout_matches_reference(golden, prefix):
    torch.set_num_threads(1)
    g = golden("rollout_wide")
    d, alpha, beta, dt = _env(g, prefix)
    out = ref.rollout_loss_torch(d, alpha, beta, dt, ref.params_from_npz(g, prefix), g[prefix + "noise"])
    assert np.array_equal(out["time_steps"], g[prefix + "time_steps"].astype(np.int64))
    assert np.array_equal(out["return_fht"], g[prefix + "return_fht"]) and out["loss"] == g[prefix + "loss"]
    for k in ref.PARAM_KEYS:
        np.testing.assert_allclose(out["grads"][k], g[f"{prefix}grad.{k}"], rtol=2e-5, atol=2e-7)


@pytest.mark.parametrize("prefix", ["nh64_", "nh256_"])
def test_wide_numpy_rollout_matches_reference(golden, prefix):
    g = golden("rollout_wide")
    e = g[prefix + "env"]
    st = ref.rollout_stats_numpy(int(e[0]), float(e[1]), float(e[2]), float(e[3]), ref.params_from_npz(g, prefix), g[prefix + "noise"],
                                 policy_opt=g[prefix + "policy_opt"], h_state=float(e[4]))
    assert np.array_equal(np.array(ref.test_policy_result(st, True)), g[prefix + "result"])


def test_long_replays_regenerate_their_noise_and_match_reference(golden):
    """rollout_long.npz stores seeds, not noise: the reference's host generators are replayed call by call."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import replay_numpy_noise
    g = golden("rollout_long")
    for prefix, dt in (("n256_", 0.005), ("nmeta_", 0.001)):
        seed, K, n_pass = (int(v) for v in g[prefix + "seed"])
        e = g[prefix + "env"]
        st = ref.rollout_stats_numpy(1, float(e[1]), float(e[2]), float(e[3]), ref.params_from_npz(g, prefix),
                                     replay_numpy_noise(seed, n_pass, K, 1, dt), policy_opt=g[prefix + "policy_opt"], h_state=float(e[4]))
        assert np.array_equal(np.array(ref.test_policy_result(st, True)), g[prefix + "result"])
