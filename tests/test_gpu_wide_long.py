"""GPU parity, second batch: wide policies (hidden width 64 / 128 / 256), long and large replays whose noise is regenerated
from the reference's seeded host generators, the metastable environment of BASELINE config 5, and the drop-in calls fed
with the REFERENCE's own environment / policy objects (when the reference sources are present: /root/reference here,
baseline/_ref/src on the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import reference_semantics as ref

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import replay_numpy_noise, replay_torch_noise  # noqa: E402  (pure torch / numpy helpers)

pytestmark = pytest.mark.gpu


def _make_env(d, alpha, beta, dt):
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D, DoubleWellStoppingTimeND
    return DoubleWellStoppingTime1D(beta=beta, alpha=alpha, dt=dt) if d == 1 else DoubleWellStoppingTimeND(d, beta=beta, alpha=alpha, dt=dt)


def _model_from(npz, prefix, d):
    from rl_sde_is_b200.models import DeterministicPolicy
    H = int(npz[prefix + "param.policy.0.bias"].shape[0])
    m = DeterministicPolicy(d, d, [H, H], nn.Tanh())
    m.load_state_dict({k: torch.from_numpy(np.array(npz[f"{prefix}param.{k}"])) for k in ref.PARAM_KEYS})
    return m


def _check_torch_case(g, prefix, noise, kernel="auto", grad_rtol=2e-4, grad_atol=5e-4):
    from rl_sde_is_b200.reinforce_deterministic_core import sample_loss_vectorized
    e = g[prefix + "env"]
    d = int(e[0])
    env, model = _make_env(d, float(e[1]), float(e[2]), float(e[3])), _model_from(g, prefix, d)
    K = noise.shape[1]
    loss, ret, steps = sample_loss_vectorized(env, model, K, noise=noise, kernel=kernel)
    assert np.array_equal(steps, g[prefix + "time_steps"])                               # hit passes: exact
    np.testing.assert_allclose(ret, g[prefix + "return_fht"], rtol=1e-5)
    np.testing.assert_allclose(float(loss.detach()), float(g[prefix + "loss"]), rtol=3e-5, atol=2e-6)
    loss.backward()
    gscale = max(np.abs(g[f"{prefix}grad.{k}"]).max() for k, _ in model.named_parameters())
    for k, p in model.named_parameters():
        ref_g = g[f"{prefix}grad.{k}"]
        np.testing.assert_allclose(p.grad.numpy(), ref_g, rtol=grad_rtol, atol=grad_atol * max(np.abs(ref_g).max(), gscale), err_msg=k)


def _check_numpy_case(g, prefix, noise, kernel="auto"):
    from rl_sde_is_b200.approximate_methods import test_policy_vectorized
    e = g[prefix + "env"]
    env = _make_env(1, float(e[1]), float(e[2]), float(e[3]))
    env.discretize_state_space(float(e[4]))
    model = _model_from(g, prefix, 1)
    res = test_policy_vectorized(env, model, batch_size=noise.shape[1], policy_opt=g[prefix + "policy_opt"], noise=noise, kernel=kernel)
    want = g[prefix + "result"]
    assert res[2] == want[2]                                                     # mean hit index: exact
    np.testing.assert_allclose(res[0], want[0], rtol=1e-6)
    np.testing.assert_allclose(res[1], want[1], rtol=1e-5)
    np.testing.assert_allclose(res[3], want[3], rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------------------ wide policies (rollout_wide.cuh)
@pytest.mark.parametrize("prefix", ["th64_", "th128_", "th256_", "t2d_h128_", "tinit_h256_"])
def test_wide_policy_loss_and_gradient_match_reference(golden, prefix):
    g = golden("rollout_wide")
    _check_torch_case(g, prefix, g[prefix + "noise"])


@pytest.mark.parametrize("prefix", ["nh64_", "nh128_", "nh256_"])
def test_wide_policy_test_rollout_matches_reference(golden, prefix):
    g = golden("rollout_wide")
    _check_numpy_case(g, prefix, g[prefix + "noise"])


def test_wide_policy_supported_shapes_and_rng_consistency():
    """rlsde_supported advertises the wide shapes; an in-kernel-RNG rollout equals the replay of its own noise, is independent
    of the tile size (16 / 32 trajectories per block is chosen from the batch size) and of sharding."""
    from rl_sde_is_b200 import _lib as L, rollout as R
    from rl_sde_is_b200.models import DeterministicPolicy
    lib = L.load()
    for d in (1, 2):
        for H in (32, 64, 128, 256):
            assert lib.rlsde_supported(d, H, 2) == 1
    assert lib.rlsde_supported(1, 96, 2) == 0 and lib.rlsde_supported(3, 256, 2) == 0
    torch.manual_seed(5)
    m = DeterministicPolicy(1, 1, [256, 256], nn.Tanh())
    m.policy[4].bias.data.fill_(1.0)
    env = _make_env(1, 1.0, 1.0, 0.005)
    params = R.flat_parameters(m).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(1, 256)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    K = 32 * n_sm + 77                                       # 32-row tiles
    ffma = {"wide_kernel": "ffma"}                           # (from 512 trajectories the automatic choice at H = 256 is the tcgen05 kernel)
    big = R.rollout_forward(env_c, mlp_c, params, K, seed=9, n_steps_lim=3000, tuning=ffma)
    assert int(big.stats[L.ST_N_UNFINISHED]) == 0
    k_small = 300                                            # 16-row tiles, shard in the middle of the global batch
    small = R.rollout_forward(env_c, mlp_c, params, k_small, seed=9, n_steps_lim=3000, traj_offset=1000, K_global=K, tuning=ffma)
    for a, b in ((big.G[1000:1300], small.G), (big.S[1000:1300], small.S), (big.T[1000:1300], small.T)):
        assert torch.equal(a, b)
    noise = R.noise_fill(9, 64, 1, int(big.T[:64].max().item()) + 1, env.dt)
    rep = R.rollout_forward(env_c, mlp_c, params, 64, n_steps_lim=3000, noise=noise)
    assert torch.equal(rep.T, big.T[:64]) and torch.equal(rep.G, big.G[:64]) and torch.equal(rep.S, big.S[:64])


def test_reinforce_runs_with_the_reference_defaults(tmp_path):
    """reinforce(env) with the reference's default policy width (256, reinforce_deterministic_core.py:102) trains: a few
    iterations at a reduced batch, losses finite, parameters move, mean hitting time of the trained policy drops."""
    from rl_sde_is_b200 import utils_path as up
    from rl_sde_is_b200.reinforce_deterministic_core import reinforce
    up.set_data_dir(str(tmp_path))
    env = _make_env(1, 1.0, 1.0, 0.005)
    data = reinforce(env, batch_size=200, lr=1e-2, n_iterations=12, seed=1, verbose=False, save=False)
    assert data["d_hidden_layer"] == 256 and np.isfinite(data["losses"]).all()
    assert data["exp_time_steps"][-3:].mean() < 0.6 * data["exp_time_steps"][0]


def test_reinforce_large_batch_wide_policy_takes_the_tensor_core_kernels(tmp_path):
    """reinforce() at the reference's default width with a batch large enough for the tcgen05 kernels (forward K1u, reverse
    K2u through the fused loss-and-gradient call): trains, and the first iteration's loss equals the CUDA-core kernels'."""
    from rl_sde_is_b200 import _lib as L, utils_path as up
    from rl_sde_is_b200.reinforce_deterministic_core import reinforce
    up.set_data_dir(str(tmp_path))
    env = _make_env(1, 1.0, 1.0, 0.005)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    K = 64 * n_sm + 500                                      # >= 64 x SMs: automatic dispatch picks the tensor-core kernels
    before = L.load().rlsde_launch_count()
    data = reinforce(env, batch_size=K, lr=1e-2, n_iterations=4, seed=1, verbose=False, save=False, n_steps_lim=20000)
    assert data["d_hidden_layer"] == 256 and np.isfinite(data["losses"]).all()
    assert data["exp_time_steps"][-1] < 0.9 * data["exp_time_steps"][0]
    assert L.load().rlsde_launch_count() > before
    import os
    os.environ["RLSDE_WIDE_KERNEL"] = "ffma"
    try:
        ref = reinforce(env, batch_size=K, lr=1e-2, n_iterations=1, seed=1, verbose=False, save=False, n_steps_lim=20000)
    finally:
        del os.environ["RLSDE_WIDE_KERNEL"]
    np.testing.assert_allclose(data["losses"][0], ref["losses"][0], rtol=2e-4)


# ------------------------------------------------------------------------------ long / large replays (noise from the seed)
@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_batch_of_256_with_multi_thousand_pass_paths(golden, kernel):
    g = golden("rollout_long")
    seed, K, n_pass = (int(v) for v in g["t256_seed"])
    assert K == 256 and n_pass > 2000
    _check_torch_case(g, "t256_", replay_torch_noise(seed, n_pass, K, 1, 0.005), kernel=kernel)
    seed, K, n_pass = (int(v) for v in g["n256_seed"])
    _check_numpy_case(g, "n256_", replay_numpy_noise(seed, n_pass, K, 1, 0.005), kernel=kernel)


@pytest.mark.parametrize("kernel", ["thread", "warp"])
def test_metastable_environment_replay(golden, kernel):
    """beta = 4, dt = 0.001 (BASELINE config 5's environment), paths of up to 2e4 passes."""
    g = golden("rollout_long")
    seed, K, n_pass = (int(v) for v in g["tmeta_seed"])
    assert n_pass > 10000
    _check_torch_case(g, "tmeta_", replay_torch_noise(seed, n_pass, K, 1, 0.001), kernel=kernel)
    seed, K, n_pass = (int(v) for v in g["nmeta_seed"])
    _check_numpy_case(g, "nmeta_", replay_numpy_noise(seed, n_pass, K, 1, 0.001), kernel=kernel)


# ------------------------------------------------------------------------------ the reference's own objects through the drop-ins
def _reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not present (neither /root/reference nor baseline/_ref)")
    return ref_loader.load()


def test_drop_ins_accept_the_reference_objects():
    """The reference's DoubleWellStoppingTime1D / 2D and DeterministicPolicy instances, unmodified, handed to this package's
    sample_loss_vectorized / test_policy_vectorized / estimate_fht_vectorized / compute_r_table / compute_p_tensor_batch:
    results equal the reference's own functions on the same objects (noise recorded from the reference's env)."""
    R = _reference()
    from rl_sde_is_b200.approximate_methods import estimate_fht_vectorized, test_policy_vectorized
    from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch, compute_r_table
    from rl_sde_is_b200.reinforce_deterministic_core import sample_loss_vectorized

    class Rec:
        def __init__(self, env, which):
            self.env, self.which, self.orig, self.noise = env, which, getattr(env, which), []
            setattr(env, which, self)

        def __call__(self, *a, **k):
            out = self.orig(*a, **k)
            self.noise.append(out[3].numpy().copy() if torch.is_tensor(out[3]) else np.array(out[3], copy=True))
            return out

        def done(self):
            delattr(self.env, self.which)
            return np.stack(self.noise).astype(np.float32)

    for d, Env in ((1, R.environments.DoubleWellStoppingTime1D), (2, R.environments_2d.DoubleWellStoppingTime2D)):
        env = Env(beta=1.0, alpha=1.0, dt=0.005)
        torch.manual_seed(20 + d)
        model = R.core.DeterministicPolicy(d, d, [32, 32], nn.Tanh())
        model.policy[4].bias.data.fill_(1.0 if d == 1 else 1.5)
        # torch path
        torch.manual_seed(3)
        rec = Rec(env, "step_torch")
        loss_ref, ret_ref, steps_ref = R.core.sample_loss_vectorized(env, model, 7)
        noise = rec.done()
        model.zero_grad()
        loss_ref.backward()
        g_ref = [p.grad.clone() for p in model.parameters()]
        model.zero_grad()
        loss, ret, steps = sample_loss_vectorized(env, model, 7, noise=noise)
        loss.backward()
        assert np.array_equal(steps, steps_ref)
        np.testing.assert_allclose(ret, ret_ref, rtol=1e-5)
        np.testing.assert_allclose(float(loss.detach()), float(loss_ref.detach()), rtol=3e-5)
        gs = max(float(gr.abs().max()) for gr in g_ref)
        for p, gr in zip(model.parameters(), g_ref):
            np.testing.assert_allclose(p.grad.numpy(), gr.numpy(), rtol=2e-4, atol=5e-4 * gs)
        # numpy path
        if d == 1:
            env.discretize_state_space(0.05)
            pol = (1.5 * np.cos(env.state_space_h)).reshape(-1, 1)
            np.random.seed(4)
            rec = Rec(env, "step")
            want = R.approx.test_policy_vectorized(env, model, batch_size=9, policy_opt=pol)
            noise = rec.done()
            got = test_policy_vectorized(env, model, batch_size=9, policy_opt=pol, noise=noise)
            assert got[2] == want[2]
            np.testing.assert_allclose(got[0], want[0], rtol=1e-6)
            np.testing.assert_allclose(got[3], want[3], rtol=1e-4)
        np.random.seed(5)
        rec = Rec(env, "step")
        want = R.approx.estimate_fht_vectorized(env, model, batch_size=6)
        noise = rec.done()
        assert estimate_fht_vectorized(env, model, batch_size=6, noise=noise) == want
    # tables from the reference's env object
    env = R.environments.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.set_action_space_bounds()
    env.discretize_state_space(0.1)
    env.discretize_action_space(0.1)
    P_ref, R_ref = R.dp.compute_p_tensor_batch(env), R.dp.compute_r_table(env)
    assert np.abs(compute_p_tensor_batch(env) - P_ref).max() < 1e-13
    assert np.array_equal(compute_r_table(env), R_ref)
    assert R.tables.check_p_tensor(env, compute_p_tensor_batch(env))


# ------------------------------------------------------------------------------ HJB reference solution vs the IS estimator
def test_hjb_psi_lies_inside_the_importance_sampling_estimate():
    """Second pin for hjb_1d: Psi(-1) = E[exp(-tau)] from the finite-difference solve against the GPU importance-sampling
    estimator of the same quantity.  The Euler-Maruyama estimate monitors the target set at discrete times and
    is biased low by O(sqrt(dt)); two step sizes a factor 4 apart are extrapolated to dt -> 0, and the HJB value must lie
    within the extrapolation's confidence interval (plus 0.5 % for the next-order term)."""
    from types import SimpleNamespace
    from rl_sde_is_b200.approximate_methods import is_estimate
    from rl_sde_is_b200.hjb_1d import HJBSolution1D
    from rl_sde_is_b200.models import DeterministicPolicy
    torch.manual_seed(1)
    model = DeterministicPolicy(1, 1, [32, 32], nn.Tanh())
    model.policy[4].bias.data.fill_(0.6)                       # a mild push: smaller relative error than the null control
    K = 2_000_000
    est = {}
    for dt in (0.004, 0.001):
        env = _make_env(1, 1.0, 1.0, dt)
        r = is_estimate(env, model, K, n_steps_lim=400000, seed=123)
        assert r["n_unfinished"] == 0
        est[dt] = (r["is_mean"], r["is_mean"] * r["is_rel_error"] / np.sqrt(K))
    psi0 = 2.0 * est[0.001][0] - est[0.004][0]                 # bias ~ c sqrt(dt): sqrt(0.004) = 2 sqrt(0.001)
    se = np.sqrt(4.0 * est[0.001][1] ** 2 + est[0.004][1] ** 2)
    hjb = float(HJBSolution1D(SimpleNamespace(beta=1.0, alpha=1.0, lb=1.0, sigma=np.sqrt(2.0)), h=5e-4).psi_at(-1.0))
    assert est[0.004][0] < est[0.001][0] < hjb                 # discrete monitoring misses crossings: biased low, less so at small dt
    assert abs(psi0 - hjb) < 4.0 * se + 5e-3 * hjb, (psi0, hjb, se)


# ------------------------------------------------------------------------------ tcgen05 forward kernel (rollout_umma.cuh)
@pytest.mark.parametrize("prefix", ["th64_", "th128_", "th256_", "t2d_h128_", "tinit_h256_"])
def test_tcgen05_forward_matches_reference(golden, monkeypatch, prefix):
    """The tensor-core forward kernel (float16 hi/lo operand split, fp32 accumulation in tensor memory) on the reference's
    recorded noise: hit passes exact, returns / loss within 1e-5, and -- since the reverse pass reads its checkpoints -- the
    reference's gradient."""
    monkeypatch.setenv("RLSDE_WIDE_KERNEL", "umma")
    g = golden("rollout_wide")
    _check_torch_case(g, prefix, g[prefix + "noise"])


@pytest.mark.parametrize("prefix", ["nh64_", "nh128_", "nh256_"])
def test_tcgen05_forward_numpy_path_matches_reference(golden, monkeypatch, prefix):
    monkeypatch.setenv("RLSDE_WIDE_KERNEL", "umma")
    g = golden("rollout_wide")
    _check_numpy_case(g, prefix, g[prefix + "noise"])


@pytest.mark.parametrize("H", [64, 128, 256])
def test_tcgen05_forward_agrees_with_cuda_core_kernel(H):
    """Same Philox stream through both wide forward kernels (tiles of 128 on the tensor cores / tiles of 32 on the CUDA
    cores): the policies are evaluated to fp32-level accuracy by both, so almost every trajectory hits on the same pass with
    the same return; the statistics agree closely.  Lane refill (more trajectories than tile slots) is exercised."""
    from rl_sde_is_b200 import _lib as L, rollout as R
    from rl_sde_is_b200.models import DeterministicPolicy
    torch.manual_seed(6)
    m = DeterministicPolicy(1, 1, [H, H], nn.Tanh())
    m.policy[4].bias.data.fill_(0.9)
    env = _make_env(1, 1.0, 1.0, 0.005)
    params = R.flat_parameters(m).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(1, H)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    K = 128 * n_sm * 2 + 333
    a = R.rollout_forward(env_c, mlp_c, params, K, seed=3, n_steps_lim=4000, tuning={"wide_kernel": "umma"})
    b = R.rollout_forward(env_c, mlp_c, params, K, seed=3, n_steps_lim=4000, tuning={"wide_kernel": "ffma"})
    Ta, Tb = a.T.cpu().numpy(), b.T.cpu().numpy()
    same = Ta == Tb
    assert same.mean() > 0.995 and (Ta >= 0).all()
    np.testing.assert_allclose(a.G.cpu().numpy()[same], b.G.cpu().numpy()[same], rtol=2e-5)
    np.testing.assert_allclose(a.S.cpu().numpy()[same], b.S.cpu().numpy()[same], rtol=2e-3, atol=2e-5)
    sa, sb = a.stats, b.stats
    assert abs(sa[L.ST_SUM_G] - sb[L.ST_SUM_G]) < 2e-4 * abs(sb[L.ST_SUM_G])
    # determinism and independence of the launch shape: a shard of the same global batch reproduces its slice bit for bit
    c = R.rollout_forward(env_c, mlp_c, params, 5000, seed=3, n_steps_lim=4000, traj_offset=700, K_global=K, tuning={"wide_kernel": "umma"})
    assert torch.equal(c.T, a.T[700:5700]) and torch.equal(c.G, a.G[700:5700])


# ------------------------------------------------------------------------------ tcgen05 reverse pass (rollout_umma_bwd.cuh)
@pytest.mark.parametrize("d,H,K,opts", [
    (1, 256, 128 * 3 + 77, {}),                         # partial last tile, several producers, one consumer
    (1, 128, 128 * 5 + 1, {}),
    (2, 128, 128 * 2 + 50, {}),
    (2, 256, 700, {"stoch_int": "exact"}),
    (1, 256, 900, {"tanh": "fast"}),
    (1, 256, 40000, {}),                                # more tiles than producers: second-round tiles, float64 drains
])
def test_tcgen05_reverse_pass_matches_cuda_core_kernel(d, H, K, opts):
    """K2u (producer / consumer CTAs, all three H x H products on the tensor cores) against K2x (FFMA tile kernel) on the
    same forward rollout: two independent implementations of the reverse recursion.  Small blocks agree to fp32 rounding;
    the H x H block carries the tensor cores' truncating accumulation (measured ~1e-5 of max |g|).  Deterministic."""
    from rl_sde_is_b200 import _lib as L, rollout as R
    from rl_sde_is_b200.models import DeterministicPolicy
    torch.manual_seed(4)
    m = DeterministicPolicy(d, d, [H, H], nn.Tanh())
    m.policy[4].bias.data.fill_(1.0 if d == 1 else 3.0)
    env = _make_env(d, 1.0, 1.0, 0.005)
    params = R.flat_parameters(m).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(d, H)
    fwd = R.rollout_forward(env_c, mlp_c, params, K, seed=12, n_steps_lim=600, store_path=True, ckpt_every=1,
                            tuning={"wide_kernel": "umma"}, **opts)
    T = fwd.T.cpu().numpy()
    assert (T >= 0).mean() > 0.5                 # (d = 1: some trajectories run into the pass budget and are skipped by both)
    grads = {}
    for name, wk in (("umma", 1), ("ffma", 2), ("umma2", 1)):
        fwd.cfg.wide_kernel = wk
        grads[name] = R.rollout_backward(env_c, mlp_c, params, fwd, 1.0 / K).cpu().numpy().astype(np.float64)
    assert np.isfinite(grads["umma"]).all()
    assert np.array_equal(grads["umma"], grads["umma2"])
    o = [0, d * H, d * H + H, d * H + H + H * H, d * H + 2 * H + H * H, 2 * d * H + 2 * H + H * H, 2 * d * H + 2 * H + H * H + d]
    for blk, (lo, hi) in zip(("W1", "b1", "W2", "b2", "W3", "b3"), zip(o[:-1], o[1:])):
        x, y = grads["umma"][lo:hi], grads["ffma"][lo:hi]
        tol = 1e-4 if blk == "W2" else 2e-5
        assert np.abs(x - y).max() <= tol * max(np.abs(y).max(), 1e-12), (blk, float(np.abs(x - y).max() / np.abs(y).max()))


def test_tcgen05_reverse_pass_needs_its_workspace():
    """rlsde_workspace_bytes_bwd covers the tensor-core reverse kernel's exchange ring; forced onto that kernel with less,
    the call answers RLSDE_ERR_WORKSPACE instead of writing past the buffer."""
    from rl_sde_is_b200 import _lib as L, rollout as R
    from rl_sde_is_b200.models import DeterministicPolicy
    lib = L.load()
    torch.manual_seed(4)
    m = DeterministicPolicy(1, 1, [256, 256], nn.Tanh())
    m.policy[4].bias.data.fill_(1.0)
    env = _make_env(1, 1.0, 1.0, 0.005)
    params = R.flat_parameters(m).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(1, 256)
    K = 300
    fwd = R.rollout_forward(env_c, mlp_c, params, K, seed=1, n_steps_lim=3000, store_path=True, ckpt_every=1)
    full = int(lib.rlsde_workspace_bytes_bwd(K, 1, 256))
    assert full >= int(lib.rlsde_workspace_bytes(K)) and full >= (100 << 20)          # 98 producers x 4 passes x 256 KB and partials
    ws = torch.empty(full, dtype=torch.uint8, device="cuda")
    g = torch.empty(int(lib.rlsde_param_count(mlp_c)), dtype=torch.float32, device="cuda")
    args = lambda nbytes: (env_c, mlp_c, params.ctypes.data, fwd.cfg, 0, fwd.G.data_ptr(), fwd.T.data_ptr(), fwd.path.data_ptr(), 0,
                           1.0 / K, g.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    fwd.cfg.wide_kernel = 1
    assert lib.rlsde_rollout_bwd(*args(full - (96 << 20))) == -5                 # RLSDE_ERR_WORKSPACE
    L.check(lib.rlsde_rollout_bwd(*args(full)), "rlsde_rollout_bwd")
    torch.cuda.synchronize()
    assert torch.isfinite(g).all()
