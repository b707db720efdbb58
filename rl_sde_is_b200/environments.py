"""Host-side mirror of the reference's environment interface for the hot path.

Reference: rl_sde_is/environments.py:5-423 (DoubleWellStoppingTime1D) and
environments_2d.py:5-368 (DoubleWellStoppingTime2D).  Same attribute names (``d, alpha, beta, sigma,
dt, lb, rb, state_init, state_space_h, action_space_h, h_state, ts_idx, not_ts_idx, is_in_ts, ...``)
and the same method names / argument meaning / return types for the calls on the hot path, so the
drop-in samplers and table builders accept either these objects or the reference's own.

What runs where:
  * ``step`` / ``step_torch`` (environments.py:139-162,201-226): one Euler-Maruyama pass on the GPU
    through ``rlsde_env_step`` -- API completeness only; the fast path is the whole-rollout kernel.
  * grids, index sets, action bounds (environments.py:250-360): NumPy on the host, called once
    ("stay in Python/NumPy; pass resulting arrays to the kernels").
  * ``step_vectorized_stopped`` (environments.py:164-199; environments_2d.py:155-182) has no caller in the reference and
    raises a broadcasting ValueError there for every input (``f(states[idx])`` has shape (n,) against (n, 1) action
    terms; verified in 1-D and 2-D).  It is provided with the semantics its body spells out -- only the lanes in ``idx``
    move, ``done`` is tested on the NEXT state with ``x >= lb``, rewards are the running cost of the moved lanes and 0
    elsewhere -- through the same single-pass kernel as ``step``.
  * ``DoubleWellStoppingTimeND`` generalises the 2-D class to any d <= 16 (the d = 10 config has no
    reference environment; semantics follow environments_2d.py:15,50-61,114-121,184-205).
"""
import numpy as np
import torch

from . import _lib as L
from .rollout import _cuda_device, _ptr


class _DoubleWellBase:
    """Shared constants: dX = (-grad V + sigma u) dt + sigma dW,  V = sum_i alpha_i (x_i^2 - 1)^2."""

    def _init_common(self, d, beta, alpha, dt, is_state_init_sampled, name_fmt):
        self.d = d
        self.beta = beta
        self.sigma = np.sqrt(2.0 / beta)
        self.sigma_tensor = torch.tensor(self.sigma, dtype=torch.float32)
        self.dt = dt
        self.dt_tensor = torch.tensor(dt, dtype=torch.float32)
        self.lb, self.rb = 1.0, 2.0
        self.is_state_init_sampled = is_state_init_sampled
        self.state_init = -np.ones((1, d), dtype=np.float32)
        self.state_space_dim = self.action_space_dim = d
        self.state_space_low, self.state_space_high = -2.0, 2.0
        self.action_space_low, self.action_space_high = 0.0, 3.0
        self.name = name_fmt.format(beta, alpha)
        # Philox key of the increments step()/step_torch() draw when none are passed.  None = not chosen yet: the first
        # such call takes it from the host generator the reference would have drawn from (np.random for step, torch for
        # step_torch), so np.random.seed / torch.manual_seed make a run reproducible as they do for the reference, while
        # two environments (train / test), two runs with different seeds or two processes do not replay the same noise.
        # Assign an integer to pin it.
        self.rng_seed = None
        self._pass_counter = 0

    # -- cheap closed forms, host side (used by tests, grids and plots, not by the kernels)
    def potential(self, state):
        v = self._alpha_vec() * (np.asarray(state) ** 2 - 1.0) ** 2
        return v if self.d == 1 else v.sum(axis=1)

    def gradient(self, state):
        return 4 * self.alpha * state * (state ** 2 - 1)

    def _alpha_vec(self):
        return np.broadcast_to(np.asarray(self.alpha, dtype=np.float64), (self.d,))

    def f(self, state):
        return np.ones(state.shape[0])

    def g(self, state):
        return np.zeros(state.shape[0])

    def reset(self, batch_size=1):
        if self.is_state_init_sampled:
            start = np.random.uniform(self.state_space_low, self.lb, (self.d,))   # one start shared by the batch
            return np.full((batch_size, self.d), start)
        return np.full((batch_size, self.d), self.state_init)

    def sample_state(self, batch_size=1):
        return np.random.uniform(self.state_space_low, self.state_space_high, (batch_size, self.d))

    def sample_action(self, batch_size=1):
        return np.random.uniform(self.action_space_low, self.action_space_high, (batch_size, self.d))

    def get_new_in_ts_idx(self, is_in_target_set, been_in_target_set):
        idx = np.flatnonzero(np.asarray(is_in_target_set) & ~np.asarray(been_in_target_set))
        been_in_target_set[idx] = True
        return idx

    def get_new_in_ts_idx_torch(self, is_in_target_set, been_in_target_set):
        idx = torch.nonzero(is_in_target_set & ~been_in_target_set).flatten()
        been_in_target_set[idx] = True
        return idx

    # -- one Euler-Maruyama pass on the GPU
    def _noise_key(self, source):
        if self.rng_seed is None:
            if source == "torch":
                self.rng_seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
            else:
                self.rng_seed = int(np.random.randint(0, 2**62, dtype=np.int64))
        return int(self.rng_seed)

    @staticmethod
    def _rank_offset():
        """Ranks of a torch.distributed job must not share Brownian increments: the rank goes into the trajectory id."""
        import torch.distributed as dist
        return (dist.get_rank() << 40) if dist.is_available() and dist.is_initialized() else 0

    def _device_step(self, state, action, f64, hit_rule, reward_type, dbt, source="numpy"):
        lib = L.load()
        in_dtype = state.dtype if not torch.is_tensor(state) else None
        grad_f32 = f64 and in_dtype == np.float32      # numpy promotion with a float32 state array
        dev = _cuda_device(None)
        K = int(state.shape[0])
        real = torch.float64 if f64 else torch.float32
        st = torch.as_tensor(np.ascontiguousarray(state), device=dev).to(real).contiguous() if not torch.is_tensor(state) \
            else state.to(device=dev, dtype=real).contiguous()
        ac = torch.as_tensor(np.ascontiguousarray(action), device=dev).to(torch.float32).contiguous() if not torch.is_tensor(action) \
            else action.detach().to(device=dev, dtype=torch.float32).contiguous()
        db_in = None
        if dbt is not None:
            db_in = torch.as_tensor(np.ascontiguousarray(dbt), device=dev).to(torch.float32).contiguous() if not torch.is_tensor(dbt) \
                else dbt.to(device=dev, dtype=torch.float32).contiguous()
        nxt = torch.empty((K, self.d), dtype=real, device=dev)
        rew = torch.empty(K, dtype=real, device=dev)
        done = torch.empty(K, dtype=torch.uint8, device=dev)
        db_out = torch.empty((K, self.d), dtype=torch.float32, device=dev)
        env_c = L.make_env(self.d, self.alpha, self.sigma, self.dt, self.lb, self.rb, self.state_init.reshape(-1), hit_rule)
        rt = {"state-action": L.REWARD_STATE_ACTION, "state-action-next-state": L.REWARD_STATE_ACTION_NEXT_STATE}[reward_type]
        with torch.cuda.device(dev):
            rc = lib.rlsde_env_step(env_c, K, _ptr(st), _ptr(ac), _ptr(db_in), 0 if dbt is not None else self._noise_key(source),
                                    self._rank_offset(), int(self._pass_counter),
                                    (L.F_STATE_F64 if f64 else 0) | (L.F_GRAD_F32 if grad_f32 else 0), rt, _ptr(nxt), _ptr(rew), _ptr(done), _ptr(db_out),
                                    torch.cuda.current_stream(dev).cuda_stream)
        L.check(rc, "rlsde_env_step")
        self._pass_counter += 1
        return nxt, rew, done.bool(), db_out

    def step(self, state, action, reward_type="state-action", dbt=None):
        """NumPy-path pass (environments.py:139-162): float64 results (numpy promotion)."""
        rule = L.HIT_X0_IN_LB_RB if self.d == 1 else L.HIT_ALL_GE_LB
        nxt, rew, done, db = self._device_step(state, action, True, rule, reward_type, dbt)
        return nxt.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), db.cpu().numpy()

    def step_vectorized_stopped(self, states, actions, idx, dbt=None):
        """Masked pass (environments.py:164-199): lanes ``idx`` take one Euler-Maruyama step, the others keep their state.
        Returns ``(next_states (K, d) f64, rewards (K, 1) f64, done (K, d) bool on the next states, dbt (n, d) f32)``."""
        states = np.asarray(states)
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        next_states = states.astype(np.float64, copy=True)
        rewards = np.zeros((states.shape[0], 1))
        db = np.zeros((0, self.d), dtype=np.float32)
        if idx.size:
            nxt, rew, _, db = self._device_step(states[idx], np.asarray(actions)[idx], True, L.HIT_ALL_GE_LB,
                                                "state-action-next-state", dbt)
            next_states[idx] = nxt.cpu().numpy()
            rewards[idx, 0] = rew.cpu().numpy()
            db = db.cpu().numpy()
        return next_states, rewards, next_states >= self.lb, db

    def step_torch(self, state, action, reward_type="state-action", dbt=None):
        """Torch-path pass (environments.py:201-226): float32; results come back on the input's device."""
        out_dev = state.device if torch.is_tensor(state) else torch.device("cpu")
        nxt, rew, done, db = self._device_step(state, action, False, L.HIT_ALL_GE_LB, reward_type, dbt, source="torch")
        return nxt.to(out_dev), rew.to(out_dev), done.to(out_dev), db.to(out_dev)


class DoubleWellStoppingTime1D(_DoubleWellBase):
    """1-D double well, target set [1, 2], x0 = -1  (environments.py:5-40)."""

    def __init__(self, beta=1.0, alpha=1.0, dt=0.005, is_state_init_sampled=False):
        self.alpha = alpha
        self._init_common(1, beta, alpha, dt, is_state_init_sampled, "doublewell-1d-st__beta{:.1f}_alpha{:.1f}")

    def is_done(self, state):
        x = np.asarray(state)[:, 0]
        return (x >= self.lb) & (x <= self.rb)

    def is_done_torch(self, state):
        return state[:, 0] >= self.lb

    # -- closed forms used by the tabular builder's host-side checks (environments.py:87-136)
    def state_action_transition_function(self, next_states, state, action, h):
        """Cell probabilities of one (state, action) column; host closed form used by tests only.
        The table builder computes all columns on the GPU (dynamic_programming.compute_p_tensor_batch)."""
        from scipy.special import ndtr
        mu = state + (-self.gradient(state) + self.sigma * action) * self.dt
        sd = self.sigma * np.sqrt(self.dt)
        upper, lower = ndtr((next_states + h - mu) / sd), ndtr((next_states - h - mu) / sd)
        prob = upper - lower
        prob[0] += lower[0]
        prob[-1] += 1 - upper[-1]
        return prob

    def reward_signal_state_action(self, state, action, done):
        running = -(self.f(state) + 0.5 * np.linalg.norm(action, axis=1) ** 2) * self.dt
        return np.where(done, -self.g(state), running)

    # -- grids and index sets (host, once): environments.py:250-360
    def set_action_space_bounds(self):
        table = {(1.0, 1.0): 3, (5.0, 1.0): 8, (1.0, 4.0): 5, (10.0, 1.0): 20}   # (alpha, beta) -> |a|_max
        a = table.get((float(self.alpha), float(self.beta)))
        if a is not None:
            self.action_space_low, self.action_space_high = -a, a

    def discretize_state_space(self, h_state):
        grid = np.arange(self.state_space_low, self.state_space_high + h_state, h_state)
        self.state_space_h = np.around(grid, decimals=3)
        self.n_states = self.state_space_h.shape[0]
        self.h_state = h_state
        self.get_state_init_idx()
        self.get_target_set_idx()

    def discretize_action_space(self, h_action):
        # not rounded, exactly like the reference: the table entries depend on these raw values
        self.action_space_h = np.arange(self.action_space_low, self.action_space_high + h_action, h_action)
        self.n_actions = self.action_space_h.shape[0]
        self.h_action = h_action
        self.get_null_action_idx()

    @staticmethod
    def _as_batch(v):
        v = np.asarray(v)
        if v.ndim == 0:
            return v[np.newaxis, np.newaxis]
        if v.ndim == 1:
            return v[np.newaxis]
        return v

    def get_state_idx(self, state):
        return self.get_state_idx_truncate(self._as_batch(state))

    def get_state_idx_truncate(self, state):
        clipped = np.clip(state, self.state_space_low, self.state_space_high)
        return np.floor((clipped - self.state_space_low) / self.h_state).astype(int)[:, 0]

    def get_action_idx(self, action):
        return self.get_action_idx_truncate(self._as_batch(action))

    def get_action_idx_truncate(self, action):
        clipped = np.clip(action, self.action_space_low, self.action_space_high)
        return np.floor((clipped - self.action_space_low) / self.h_action).astype(int)[:, 0]

    def get_state_init_idx(self):
        self.state_init_idx = self.get_state_idx(self.state_init)

    def get_target_set_idx(self):
        grid = self.state_space_h
        self.is_in_ts = (grid >= self.lb) & (grid <= self.rb)
        self.lb_idx = self.get_state_idx(np.array([[self.lb]]))[0]
        self.rb_idx = self.get_state_idx(np.array([[self.rb]]))[0]
        self.ts_idx = np.flatnonzero(self.is_in_ts)
        self.not_ts_idx = np.flatnonzero(~self.is_in_ts)

    def get_null_action_idx(self):
        self.null_action_idx = self.get_action_idx(np.zeros((1, self.d)))

    def get_hjb_solver(self, h_hjb=0.001):
        """Reference solution on ``state_space_h`` (environments.py:392-420).  The reference loads it from the external
        ``sde_hjb_solver`` package; here it comes from the finite-difference solve of ``hjb_1d``.  The
        returned object has the attributes the callers read: ``u_opt`` (n_states, 1) and ``value_function``
        (n_states,) = -log Psi, so that ``-value_function`` is the optimal value table (tabular_dp_tables.py:88)."""
        from .hjb_1d import HJBSolution1D
        sol = HJBSolution1D(self, h=h_hjb)
        sol.u_opt = sol.u_opt_at(self.state_space_h).reshape(-1, 1)
        sol.value_function = sol.value_function_at(self.state_space_h)
        return sol


class DoubleWellStoppingTimeND(_DoubleWellBase):
    """d-dimensional double well, target set {x_i >= 1 for all i}; d = 2 is the reference's 2-D class."""

    def __init__(self, d=2, beta=1.0, alpha=1.0, dt=0.005, is_state_init_sampled=False):
        if not 1 <= d <= L.RLSDE_MAX_D:
            raise ValueError(f"d must be in 1..{L.RLSDE_MAX_D}")
        self.alpha = np.full(d, alpha)
        self.alpha_tensor = torch.tensor(self.alpha, dtype=torch.float32)
        self._init_common(d, beta, alpha, dt, is_state_init_sampled, "doublewell-%dd-st__beta{:.1f}_alpha{:.1f}" % d)

    def is_done(self, state):
        return (np.asarray(state) >= self.lb).all(axis=1)

    def is_done_torch(self, state):
        return (state >= self.lb).all(axis=1)

    def step(self, state, action, dbt=None):
        return super().step(state, action, "state-action", dbt)

    def step_torch(self, state, action, dbt=None):
        return super().step_torch(state, action, "state-action", dbt)


class DoubleWellStoppingTime2D(DoubleWellStoppingTimeND):
    """environments_2d.py:5-43"""

    def __init__(self, beta=1.0, alpha=1.0, dt=0.005, is_state_init_sampled=False):
        super().__init__(2, beta, alpha, dt, is_state_init_sampled)
        self.name = "doublewell-2d-st__beta{:.1f}_alpha{:.1f}".format(beta, alpha)
