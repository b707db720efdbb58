"""Replay buffers fed by the fused rollout kernel.

Reference: rl_sde_is/replay_buffers.py:7-88 -- a ring of five preallocated arrays (``states``, ``actions``, ``rewards``,
``next_states``, ``done``; float32 / bool) with a write pointer ``ptr`` and a fill level ``size``.  Both classes here
expose exactly those attributes and the methods ``reset``, ``store``, ``store_vectorized``, ``sample_batch`` and
``estimate_episode_length``:

  * ``ReplayBuffer``        host NumPy arrays, the reference's container.  Like the reference's (:56-68) its
                            ``store_vectorized`` does not wrap around: a batch that does not fit behind ``ptr`` raises
                            NumPy's broadcasting ValueError.
  * ``DeviceReplayBuffer``  the same ring as CUDA tensors, for consumers that train on the GPU; batches wrap around and
                            sampling uses torch's CUDA generator.

They share one implementation parameterised by the array backend.
"""
import numpy as np
import torch

_FIELDS = ("states", "actions", "rewards", "next_states", "done")


class _TransitionRing:
    wraps = False

    # ---- backend hooks
    def _zeros(self, shape, kind):
        raise NotImplementedError

    def _arange(self, n):
        raise NotImplementedError

    def _randint(self, hi, n):
        raise NotImplementedError

    # ---- common logic
    def _allocate(self):
        shapes = {
            "states": ((self.max_size, self.state_dim), "f32"),
            "next_states": ((self.max_size, self.state_dim), "f32"),
            "actions": ((self.max_size, self.action_dim), "f32") if self.is_action_continuous else ((self.max_size,), "i64"),
            "rewards": ((self.max_size,), "f32"),
            "done": ((self.max_size,), "bool"),
        }
        for name in _FIELDS:
            setattr(self, name, self._zeros(*shapes[name]))
        self.ptr, self.size, self.is_full = 0, 0, False

    def reset(self):
        self._allocate()

    def _advance(self, n):
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)

    def store(self, state, action, reward, next_state, done):
        for name, value in zip(_FIELDS, (state, action, reward, next_state, done)):
            getattr(self, name)[self.ptr] = value
        self._advance(1)
        if not self.is_full and self.size == self.max_size:
            self.is_full = True
            print('Replay buffer is full!')

    def store_vectorized(self, states, actions, rewards, next_states, done):
        batch = dict(zip(_FIELDS, (states, actions, rewards, next_states, done)))
        n = int(states.shape[0])
        if self.wraps and n > self.max_size:            # only the newest max_size tuples can survive a wrap
            batch = {k: v[n - self.max_size:] for k, v in batch.items()}
            n = self.max_size
        head = min(n, self.max_size - self.ptr) if self.wraps else n
        for name, src in batch.items():
            dst = getattr(self, name)
            dst[self.ptr:self.ptr + head] = src[:head]
            if head < n:
                dst[:n - head] = src[head:]
        self._advance(n)

    def sample_batch(self, batch_size=None):
        idxs = self._arange(self.size) if batch_size is None else self._randint(self.size, int(batch_size))
        return {name: getattr(self, name)[idxs] for name in _FIELDS}

    def estimate_episode_length(self):
        return self.done.sum() / self.size


class ReplayBuffer(_TransitionRing):

    def __init__(self, size, state_dim, action_dim=None, is_action_continuous=True):
        if is_action_continuous:
            assert action_dim is not None, ''
        self.max_size, self.state_dim, self.action_dim = size, state_dim, action_dim
        self.is_action_continuous = is_action_continuous
        self._allocate()

    def _zeros(self, shape, kind):
        return np.zeros(shape, dtype={"f32": np.float32, "i64": np.int64, "bool": bool}[kind])

    def _arange(self, n):
        return np.arange(n)

    def _randint(self, hi, n):
        return np.random.randint(0, hi, size=n)


class DeviceReplayBuffer(_TransitionRing):
    wraps = True

    def __init__(self, size, state_dim, action_dim, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceReplayBuffer needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_size, self.state_dim, self.action_dim = int(size), int(state_dim), int(action_dim)
        self.is_action_continuous = True
        self._allocate()

    def _zeros(self, shape, kind):
        return torch.zeros(shape, dtype={"f32": torch.float32, "i64": torch.int64, "bool": torch.bool}[kind], device=self.device)

    def _arange(self, n):
        return torch.arange(n, device=self.device)

    def _randint(self, hi, n):
        return torch.randint(0, hi, (n,), device=self.device)

    def estimate_episode_length(self):
        return float(self.done[:self.size].sum().item()) / self.size
