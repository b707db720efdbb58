#!/usr/bin/env python
"""Benchmarks of the rl-sde-is hot path on B200: one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # headline = config 2 (BASELINE.json configs[1])
    python bench.py --config {1,2,3,4,5} ...                       # the other BASELINE.json configs (1-based: C1..C5)
    python bench.py --impl reference [--config c] ...              # the reference's own CPU implementation of that config
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, trajectories sharded (weak scaling)

Configs (BASELINE.json `configs[c-1]`):
  C1  1-D REINFORCE, K = 100, lr 1e-2, seed 1: a step = one training iteration; metric = iterations/s
  C2  1-D policy test rollout, 1e6 trajectories per GPU, n_steps_lim 1000, IS estimator + relative error (HEADLINE):
      a step = one launch of the rollout kernel over the whole batch + the statistics reduction;
      metric = useful Euler-Maruyama trajectory-steps/s (one pass of one trajectory up to and including the pass on
      which its hit is detected, or n_steps_lim if it is not)
  C3  tabular tables h_state = h_action = 0.01: a step = one build of P (401, 401, 601) + R; metric = GB/s written
  C4  d = 10 double well, MLP 10-32-32-10 (head bias +3), 1e7 trajectories over 8 GPUs (1.25e6 per GPU), n_steps_lim 2000:
      a step = loss + gradient of the global batch incl. the one exchange; metric = useful trajectory-steps/s (training)
  C5  metastable 1-D beta = 4, dt = 0.001, 1e8 trajectories over 8 GPUs (1.25e7 per GPU), budget 1e6 passes:
      a step = one rollout of the whole batch; metric = useful trajectory-steps/s
`value` is timed with CUDA events with everything resident on the GPU (max over ranks); `e2e` is the same metric through
the public, reference-named Python call with host inputs and host results (wall clock).  `--impl reference` runs the
UNMODIFIED reference (baseline/_ref/src, imported with the three stubs of SURVEY App. B; oracle/ref_loader.py) on the
host cores, on a bounded sample of the same workload; without the reference sources it falls back to the oracle port.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(alpha=1.0, beta=1.0, dt=0.005, d=1, hidden=32, n_steps_lim=1000, K_per_gpu=1_000_000, policy_seed=1)
UNIT = "trajectory-steps/s"
METRIC = "SDE trajectory-steps/sec"
MEASURED_FFMA_TFLOPS = 69.2        # best FFMA rate measured on this pool (profiles/r01/microbench_pipes.jsonl: ffma_rcr 3.458e13 FMA/s)


def flop_fwd(d, H):                # forward rollout, per useful trajectory-step
    return 2 * (2 * d * H + H * H) + 14 * d + 4


def flop_train(d, H):              # forward + reverse (what autograd executes: 1x forward, 2x backward of the MLP)
    return 3 * 2 * (2 * d * H + H * H) + 24 * d + 4


def fp32_peak_tflops(sm_count, sm_max_mhz):
    return sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def make_policy(seed, d=1, hidden=32, head_bias=None, module=None):
    """Reference-style policy at its seeded initialisation (reinforce_deterministic_core.py:126-139)."""
    import torch
    import torch.nn as nn
    if module is None:
        from rl_sde_is_b200 import models as module
    np.random.seed(seed)
    torch.manual_seed(seed)
    m = module.DeterministicPolicy(d, d, [hidden] * 2, nn.Tanh())
    if head_bias is not None:
        m.policy[4].bias.data.fill_(head_bias)
    return m


def cpu_threads():
    """Host threads the CPU arm uses: all of them, whatever OMP_NUM_THREADS torchrun exported (RLSDE_REF_THREADS overrides)."""
    n = int(os.environ.get("RLSDE_REF_THREADS", "0")) or (os.cpu_count() or 1)
    return max(1, min(n, 64))


# --------------------------------------------------------------------------------------------------
# CPU arms.  kind "reference" = the unmodified reference; kind "port" = oracle/ (when the reference is absent)
# --------------------------------------------------------------------------------------------------
def ref_or_none():
    try:
        from oracle import ref_loader
        return ref_loader.load() if ref_loader.available() else None
    except Exception:
        return None


def cpu_rollout_reference(K, seed=0, beta=1.0, dt=0.005, lim=None, hidden=32):
    """Reference `test_policy_vectorized` (approximate_methods.py:577-648) with k_max = n_steps_lim, policy_opt = zeros
    (BASELINE.md section 4).  Useful steps are counted by a wrapper on the env INSTANCE's step method."""
    import torch
    from oracle import ref_loader
    ref = ref_loader.load()
    lim = lim or WORKLOAD["n_steps_lim"]
    env = ref.environments.DoubleWellStoppingTime1D(beta=beta, alpha=WORKLOAD["alpha"], dt=dt)
    env.discretize_state_space(0.05)
    model = make_policy(WORKLOAD["policy_seed"], hidden=hidden, module=ref.core)
    counter = ref_loader.UsefulStepCounter(env, "step")
    torch.set_num_threads(cpu_threads())
    np.random.seed(seed)
    t0 = time.perf_counter()
    ref.approx.test_policy_vectorized(env, model, batch_size=K, k_max=lim, policy_opt=np.zeros((env.n_states, 1)))
    wall = time.perf_counter() - t0
    return counter.useful / wall, counter.useful, wall, int(torch.get_num_threads())


def cpu_rollout_port(K, seed=0):
    """Oracle port of the same loop (used only when the reference sources are absent)."""
    import torch
    from oracle import reference_semantics as ref
    torch.set_num_threads(cpu_threads())
    model = make_policy(WORKLOAD["policy_seed"])
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    lim = WORKLOAD["n_steps_lim"]
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    noise = (np.sqrt(WORKLOAD["dt"]) * rng.standard_normal((lim, K, 1))).astype(np.float32)
    st = ref.rollout_stats_numpy(1, WORKLOAD["alpha"], WORKLOAD["beta"], WORKLOAD["dt"], params, noise)
    wall = time.perf_counter() - t0
    lens = st["ep_lens"]
    useful = int(np.where(lens >= 0, lens + 1, lim).sum())
    return useful / wall, useful, wall, int(torch.get_num_threads())


def cpu_rollout(K, seed=0):
    if ref_or_none() is not None:
        v, u, w, t = cpu_rollout_reference(K, seed)
        return v, u, w, t, "reference", "reference test_policy_vectorized (approximate_methods.py:577-648), k_max = n_steps_lim, policy_opt = zeros"
    v, u, w, t = cpu_rollout_port(K, seed)
    return v, u, w, t, "port", "oracle port of the reference's per-pass torch forward + NumPy step loop (oracle/reference_semantics.py)"


def cpu_reinforce_reference(n_iterations, threads=1):
    """Reference `reinforce()` (reinforce_deterministic_core.py:102-336), config C1: K = 100, lr 1e-2, seed 1."""
    import contextlib
    import io
    import torch
    from oracle import ref_loader
    ref = ref_loader.load()
    env = ref.environments.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.discretize_state_space(0.05)
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(io.StringIO()):
        data = ref.core.reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=n_iterations, seed=1,
                                  backup_freq_iterations=None, policy_opt=np.zeros((env.n_states, 1)))
    return np.asarray(data["cts"], dtype=np.float64), np.asarray(data["exp_time_steps"], dtype=np.float64)


def cpu_tables_reference(h):
    """Reference `compute_r_table` + `compute_p_tensor_batch` (dynamic_programming.py:3-36); returns (seconds, bytes)."""
    from oracle import ref_loader
    ref = ref_loader.load()
    env = ref.environments.DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.set_action_space_bounds()
    env.discretize_state_space(h)
    env.discretize_action_space(h)
    t0 = time.perf_counter()
    r = ref.dp.compute_r_table(env)
    p = ref.dp.compute_p_tensor_batch(env)
    wall = time.perf_counter() - t0
    return wall, int(p.nbytes + r.nbytes), p, r, env


def cpu_train_reference(d, K, beta, dt, head_bias, n_steps_hint=""):
    """Reference `sample_loss_vectorized` + backward on the reference's own env of dimension d (1 or 2)."""
    import torch
    from oracle import ref_loader
    ref = ref_loader.load()
    env = (ref.environments.DoubleWellStoppingTime1D if d == 1 else ref.environments_2d.DoubleWellStoppingTime2D)(beta=beta, alpha=1.0, dt=dt)
    model = make_policy(1, d=d, head_bias=head_bias, module=ref.core)
    torch.set_num_threads(cpu_threads())
    torch.manual_seed(0)
    t0 = time.perf_counter()
    loss, ret, steps = ref.core.sample_loss_vectorized(env, model, K)
    loss.backward()
    wall = time.perf_counter() - t0
    return float(np.sum(steps)) / wall, float(np.sum(steps)), wall, int(torch.get_num_threads())


def run_reference_arm(args):
    """`--impl reference`: rank 0 alone measures; the other ranks of a torchrun launch exit at once."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    have_ref = ref_or_none() is not None
    c = args.config
    vals, units, walls = [], 0.0, 0.0
    threads, kind, how, unit, metric, dtype = 1, "reference" if have_ref else "port", "", UNIT, METRIC, "f64 state / f32 policy"
    if c == 2:
        K = args.ref_trajectories or 100000
        for _ in range(args.warmup):
            cpu_rollout(2000)
        for s in range(args.steps):
            v, u, w, threads, kind, how = cpu_rollout(K, seed=s)
            units += u; walls += w
        sample = f"{K} trajectories x n_steps_lim {WORKLOAD['n_steps_lim']} per step (of the 1e6-trajectory workload): {how}"
    elif c == 1:
        if not have_ref:
            return print(json.dumps({"impl": "reference", "unavailable": "config 1 needs the reference sources (baseline/_ref/src)"})) or 0
        n = max(args.steps, 1) + args.warmup
        cts, _ = cpu_reinforce_reference(n, threads=1)
        units, walls = float(args.steps), float(cts[args.warmup:].sum())
        unit, metric, dtype = "iterations/s", "REINFORCE iter/s", "f32"
        sample = (f"reference reinforce() K=100 lr=1e-2 seed=1, iterations {args.warmup}..{n - 1} (wall clock `cts`), 1 torch thread "
                  "(measured faster than all cores at this batch size: BASELINE.md)")
    elif c == 3:
        if not have_ref:
            return print(json.dumps({"impl": "reference", "unavailable": "config 3 needs the reference sources (baseline/_ref/src)"})) or 0
        for _ in range(min(args.warmup, 1)):
            cpu_tables_reference(0.1)
        for s in range(args.steps):
            w, nbytes, _, _, _ = cpu_tables_reference(0.05 if args.ref_full else 0.1)
            units += nbytes / 1e9; walls += w
        unit, metric, dtype = "GB/s", "tabular P/R table build, bytes written per second", "f64"
        sample = ("reference compute_r_table + compute_p_tensor_batch at h_state = h_action = %s (the h = 0.01 build takes ~55 s per "
                  "step; its time per (s, a) column is the same)" % ("0.05" if args.ref_full else "0.1"))
    elif c in (4, 5):
        if not have_ref:
            return print(json.dumps({"impl": "reference", "unavailable": f"config {c} needs the reference sources (baseline/_ref/src)"})) or 0
        for s in range(args.steps):
            if c == 4:      # no d = 10 env exists in the reference: its 2-D env, same policy recipe
                v, u, w, threads = cpu_train_reference(2, 1000, 1.0, 0.005, 3.0)
            else:           # beta = 4: 1 000 passes of the reference's test rollout on 1 000 trajectories
                v, u, w, threads = cpu_rollout_reference(1000, seed=s, beta=4.0, dt=0.001, lim=1000)
            units += u; walls += w
        dtype = "f32" if c == 4 else "f64 state / f32 policy"
        sample = ("reference sample_loss_vectorized + backward on the reference's 2-D env (no d = 10 env exists), K = 1000, head bias +3"
                  if c == 4 else "reference test_policy_vectorized, beta = 4, dt = 0.001, K = 1000, first 1000 passes (k_max = 1000)")
    else:
        raise SystemExit("unknown config")
    value = units / max(walls, 1e-12)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * walls / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": workload_config(c, args.gpus, args),
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample, "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(c, n_gpus, args=None):
    K2 = (args.trajectories if args is not None and args.trajectories else WORKLOAD["K_per_gpu"])
    table = {
        1: {"workload": "configs[0] (C1): 1-D double well alpha=1 beta=1 dt=0.005, REINFORCE deterministic, batch-size 100, lr 1e-2, seed 1",
            "policy": "MLP 1-32-32-1 tanh", "step": "one training iteration (rollout, reverse pass, Adam), device-resident loop"},
        2: {"workload": "configs[1] (C2): 1-D double well alpha=1 beta=1 dt=0.005, policy test rollout, 1e6 trajectories per GPU, "
                        "n_steps_lim=1000, IS estimator + relative error",
            "trajectories_per_gpu": K2, "n_steps_lim": WORKLOAD["n_steps_lim"], "policy": "MLP 1-32-32-1 tanh, seed-1 init",
            "parallelism": f"trajectories sharded over {n_gpus} GPU(s), statistics all-reduced",
            "l2": "kernel inputs are by-value parameters only (no reusable HBM inputs); outputs are write-only; "
                  "a 256 MiB flush write runs between steps, outside the timed events"},
        3: {"workload": "configs[2] (C3): tabular_dp_tables 1-D alpha=1 beta=1 dt=0.005 h-state=h-action=0.01: P (401,401,601) f64 + R (401,601)",
            "step": "one build of P and R into alternating 773 MB outputs (larger than L2)", "parallelism": "replicas only (no exchange)"},
        4: {"workload": "configs[3] (C4): d=10 double well alpha=1 beta=1 dt=0.005, MLP 10-32-32-10 (head bias +3), 1e7 trajectories over "
                        "8 GPUs, n_steps_lim 2000, REINFORCE loss + gradient with one exchange",
            "trajectories_per_gpu": (args.trajectories if args is not None and args.trajectories else 1_250_000),
            "parallelism": f"trajectories sharded over {n_gpus} GPU(s); one all-gather of [gradient | statistics] rows per step"},
        5: {"workload": "configs[4] (C5): metastable 1-D beta=4 alpha=1 dt=0.001, seed-1 initial policy, 1e8 trajectories over 8 GPUs, "
                        "pass budget 1e6, IS estimator + relative error",
            "trajectories_per_gpu": (args.trajectories if args is not None and args.trajectories else 12_500_000),
            "parallelism": f"trajectories sharded over {n_gpus} GPU(s), statistics all-reduced",
            "warmup_note": "warm-up steps run 1/64 of the batch (a full-size step takes ~50 s)"},
    }
    return table[c]


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields through NVML)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:        # NVML indices ignore CUDA_VISIBLE_DEVICES: find the device torch calls `index` by its UUID
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv is not None:
            self.thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# GPU arm: shared scaffolding
# --------------------------------------------------------------------------------------------------
class Job:
    """Process-group setup, timing of K steps bracketed by barrier + synchronize, max over ranks, one JSON line on rank 0."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        # stdout carries exactly ONE line, the JSON record: everything libraries print on file descriptor 1 meanwhile
        # (NCCL's version banner comes from C code) is sent to stderr
        sys.stdout.flush()
        self.json_fd = os.dup(1)
        os.dup2(2, 1)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.sm = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.peaks = load_peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, step_fn, steps, warmup, warm_fn=None):
        """W untimed warm-up steps, then exactly K steps timed with CUDA events on the launching stream (an L2 flush
        write between steps, outside the events).  Returns (device ms summed over steps, max over ranks; per-step results;
        this library's kernel launches inside the timed region; clock summary; wall seconds)."""
        torch = self.torch
        from rl_sde_is_b200 import _lib as L
        for w in range(warmup):
            (warm_fn or step_fn)(1000 + w)
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        results = []
        launches0 = L.load().rlsde_launch_count()
        with ClockSampler(self.local_rank) as clk:
            wall0 = time.perf_counter()
            for s in range(steps):
                self.flush.fill_(s & 0xFF)
                ev[s][0].record()
                results.append(step_fn(s))
                ev[s][1].record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - wall0
        launches = int(L.load().rlsde_launch_count() - launches0)
        self.barrier()
        dev_ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
        return dev_ms, results, launches, clk.summary(), wall

    def sm_max_mhz(self, clocks):
        return float(self.peaks.get("sm_max_mhz", clocks.get("sm_max_mhz") or 1965.0))

    def fp32_roofline(self, per_gpu_units_per_s, flop_per_unit, clocks, note, traffic=None):
        peak = fp32_peak_tflops(self.sm, self.sm_max_mhz(clocks))
        achieved = per_gpu_units_per_s * flop_per_unit / 1e12
        return {"bound": "fp32_cuda_core", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "frac_of_measured_ffma": achieved / MEASURED_FFMA_TFLOPS, "peak_measured_ffma": MEASURED_FFMA_TFLOPS,
                "traffic": traffic,
                "note": f"peak = {self.sm} SMs x 128 FMA lanes x 2 x {self.sm_max_mhz(clocks):.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json); "
                        f"peak_measured_ffma = best FFMA microbenchmark rate on this pool (profiles/r01/microbench_pipes.jsonl); "
                        f"algorithmic work {flop_per_unit} FLOP per useful trajectory-step (SURVEY 8d). " + note}

    def emit(self, line):
        if self.rank == 0:
            sys.stdout.flush()
            os.write(self.json_fd, (json.dumps(line) + "\n").encode())
        if self.world > 1:
            self.dist.destroy_process_group()
        return 0


# --------------------------------------------------------------------------------------------------
# C2 (headline): policy test rollout
# --------------------------------------------------------------------------------------------------
def run_config2(args):
    job = Job(args)
    torch = job.torch
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200 import rollout as R
    from rl_sde_is_b200.approximate_methods import is_estimate, test_policy_vectorized
    from rl_sde_is_b200.distributed import Shard
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    dev, world, rank = job.dev, job.world, job.rank
    env = DoubleWellStoppingTime1D(beta=WORKLOAD["beta"], alpha=WORKLOAD["alpha"], dt=WORKLOAD["dt"])
    env.discretize_state_space(0.05)
    model = make_policy(WORKLOAD["policy_seed"])
    params = R.flat_parameters(model).detach().numpy()
    K = args.trajectories or WORKLOAD["K_per_gpu"]
    shard = Shard(K * world, rank, world)
    env_c, mlp_c = R.env_struct(env, L.HIT_X0_IN_LB_RB), L.make_mlp(1, WORKLOAD["hidden"])
    lim = WORKLOAD["n_steps_lim"]

    def one_step(seed):
        out = R.rollout_forward(env_c, mlp_c, params, K, seed=seed, n_steps_lim=lim, stoch_int="exact", want_logw=True,
                                traj_offset=shard.traj_offset, K_global=shard.K_global, tanh=args.tanh, device=dev)
        stats = out.stats_dev
        if world > 1:
            shard.all_reduce_stats(stats)               # the path's only exchange step: 16 doubles
        return stats

    dev_ms, stats_steps, gpu_launches, clocks, wall = job.timed(one_step, args.steps, args.warmup)
    st = [s.cpu().numpy() for s in stats_steps]        # already global sums when world > 1
    useful = float(sum(x[L.ST_USEFUL_STEPS] for x in st))
    value = useful / (dev_ms * 1e-3)
    summ = R.summarize(st[-1])

    # ---- end to end through the public API: host parameters in, statistics out, wall clock
    e2e_wall, e2e_useful = 0.0, 0.0
    job.barrier()
    for s in range(args.steps):
        t0 = time.perf_counter()
        res = is_estimate(env, model, K, n_steps_lim=lim, seed=5000 + s, tanh=args.tanh, device=dev, dist=shard if world > 1 else None)
        e2e_wall += time.perf_counter() - t0
        e2e_useful += res["useful_steps"]
    e2e_value = e2e_useful / job.max_over_ranks(e2e_wall)

    # ---- the drop-in's own arithmetic (reference NumPy path): float64 state and accumulators, float32
    # policy, l2-error lookup against policy_opt -- what test_policy_vectorized(env, model, K, k_max, policy_opt) runs
    pol0 = np.zeros((env.n_states, 1))
    f64 = {}
    if not args.no_f64:
        def f64_step(seed):
            out = R.rollout_forward(env_c, mlp_c, params, K, seed=seed, n_steps_lim=lim, state_f64=True, want_logw=False,
                                    policy_opt=pol0, grid=(env.state_space_low, env.state_space_high, env.h_state),
                                    traj_offset=shard.traj_offset, K_global=shard.K_global, tanh=args.tanh, device=dev)
            stats = out.stats_dev
            if world > 1:
                shard.all_reduce_stats(stats)
            return stats
        n64 = max(2, min(args.steps, 3))
        ms64, st64, _, _, _ = job.timed(f64_step, n64, 1)
        useful64 = float(sum(x.cpu().numpy()[L.ST_USEFUL_STEPS] for x in st64))
        job.barrier()
        t0 = time.perf_counter()
        res64 = test_policy_vectorized(env, model, K, k_max=lim, policy_opt=pol0, seed=7000, tanh=args.tanh, device=dev,
                                       dist=shard if world > 1 else None)
        torch.cuda.synchronize()
        w64 = job.max_over_ranks(time.perf_counter() - t0)
        v64 = useful64 / (ms64 * 1e-3)
        f64 = {"value": v64, "unit": UNIT, "ms_per_step": ms64 / n64, "steps": n64, "dtype": "f64 state and accumulators, f32 policy",
               "fp32_frac": v64 / world * flop_fwd(1, 32) / 1e12 / fp32_peak_tflops(job.sm, job.sm_max_mhz(clocks)),
               "e2e_value": useful64 / n64 / w64, "e2e_call": "test_policy_vectorized(env, model, K, k_max=1000, policy_opt=zeros)",
               "result": [None if (isinstance(x, float) and np.isnan(x)) else float(x) for x in res64],
               "note": "the reference's NumPy-path arithmetic (drop-in default state_f64=True); fraction is of the FP32 roofline with the "
                       "same 2194 FLOP credit (the f64 environment pass and accumulators are uncredited)"}
    if rank != 0:
        return job.emit(None)

    per_gpu = useful / world / (dev_ms * 1e-3)
    roofline = job.fp32_roofline(per_gpu, flop_fwd(1, 32), clocks,
                                 "Timed per launch with CUDA events incl. the statistics reduction; state lives in registers "
                                 "(results: 16 B per trajectory).", traffic=1.852e9)
    roofline["traffic_unit"] = ("bytes per launch (dram read + write, ncu --set full, profiles/r02/ncu_rollout_fwd_d1_session2.txt): the "
                                "continuation records of the time-sliced schedule, 24 B per 8 passes; 74 GB/s")
    extra = {"is_mean": summ.get("is_mean"), "is_rel_error": summ.get("is_rel_error"), "mean_return": summ.get("mean_return"),
             "frac_unfinished": summ["n_unfinished"] / max(summ["n"], 1), "useful_steps_per_step": useful / args.steps,
             "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps, "tanh": args.tanh, "f64_state": f64}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_rollout(2000)                                 # warm-up (thread pools, allocator)
        ck = args.ref_trajectories or 200000           # ~15 s of host work (the spec's 10-30 s sample)
        v, u, w, thr, kind, how = cpu_rollout(ck)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": thr, "kind": kind, "host_cores": os.cpu_count(),
                        "sample": f"{ck} trajectories x n_steps_lim {lim} ({u} useful steps, {w:.1f} s): {how}"}
        if not args.no_extras:
            extra.update(secondary_measurements(dev))
            extra.update(reference_cpu_measurements())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(2, world, args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes + 1400),
                "d2h_bytes_per_step": int(L.RLSDE_NSTATS * 8),
                "note": "is_estimate(env, model, K): host policy parameters travel as kernel arguments, the 16-double statistics "
                        "record comes back; wall clock around the call"},
        "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra,
    }
    return job.emit(line)


# --------------------------------------------------------------------------------------------------
# C1: REINFORCE iterations/s at K = 100
# --------------------------------------------------------------------------------------------------
def run_config1(args):
    job = Job(args)
    torch = job.torch
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200.distributed import Shard
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    from rl_sde_is_b200.reinforce_deterministic_core import reinforce
    world, rank, dev = job.world, job.rank, job.dev
    K = (args.trajectories or 100) * world               # weak scaling: 100 trajectories per GPU
    n = args.warmup + args.steps
    env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    shard = Shard(K, rank, world) if world > 1 else None
    launches0 = L.load().rlsde_launch_count()
    with ClockSampler(job.local_rank) as clk:
        job.barrier()
        t0 = time.perf_counter()
        data = reinforce(env, d_hidden_layer=32, batch_size=K, lr=1e-2, n_iterations=n, seed=1, verbose=False, save=False,
                         device=dev, **({"dist": shard} if shard else {}))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    launches = int(L.load().rlsde_launch_count() - launches0)
    cts = np.asarray(data["cts"], dtype=np.float64)
    dev_s = job.max_over_ranks(float(cts[args.warmup:].sum()))
    value = args.steps / dev_s
    steps_timed = float(np.sum(data["time_steps"][args.warmup * (K // world):]))        # this rank's useful passes in the timed iterations
    useful = job.sum_over_ranks(steps_timed)
    # same loop through the host route (torch.optim.Adam on the CPU module after every fused step) = the reference's call pattern
    e2e = reinforce(env, d_hidden_layer=32, batch_size=K, lr=1e-2, n_iterations=n, seed=1, verbose=False, save=False, device=dev,
                    device_loop=False, **({"dist": shard} if shard else {}))
    e2e_s = job.max_over_ranks(float(np.sum(e2e["cts"][args.warmup:])))
    if rank != 0:
        return job.emit(None)
    clocks = clk.summary()
    per_gpu = useful / world / dev_s
    roofline = job.fp32_roofline(per_gpu, flop_train(1, 32), clocks,
                                 "Latency-bound regime by construction: 100 trajectories per GPU occupy 100 warps; an iteration "
                                 "costs (longest trajectory) x (dependent latency of a forward + a reverse pass).")
    line = {"metric": "REINFORCE iter/s", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(1, world, args), "clocks": clocks,
            "e2e": {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": 4612, "d2h_bytes_per_step": 128 + 4612 + 8 * (K // world),
                    "note": "reinforce(device_loop=False): every iteration the host packs the parameters into the kernel arguments, "
                            "reads loss, gradient, returns and hit passes back in one copy and steps torch.optim.Adam on the CPU module"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None,
            "extra": {"iterations_total": n, "mean_steps_timed": float(np.mean(data["exp_time_steps"][args.warmup:])),
                      "mean_steps_it0": float(data["exp_time_steps"][0]), "wall_s_total": wall,
                      "useful_train_steps_per_s": useful / dev_s, "batch_global": K}}
    return job.emit(line)


# --------------------------------------------------------------------------------------------------
# C3: tables
# --------------------------------------------------------------------------------------------------
def run_config3(args):
    job = Job(args)
    torch = job.torch
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch, compute_r_table
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    dev, world, rank = job.dev, job.world, job.rank
    env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
    env.set_action_space_bounds(); env.discretize_state_space(0.01); env.discretize_action_space(0.01)
    bufs = [torch.empty((env.n_states, env.n_states, env.n_actions), dtype=torch.float64, device=dev) for _ in range(2)]
    nbytes = bufs[0].numel() * 8 + env.n_states * env.n_actions * 8
    reps = 8

    def step(s):                                       # alternating 773 MB outputs: nothing stays in L2
        for i in range(reps):
            compute_p_tensor_batch(env, out=bufs[i & 1], exact_cdf=args.exact_cdf)
            compute_r_table(env, device=dev, device_out=True)
    compute_p_tensor_batch(env, out=bufs[1])           # keeps the stream busy at the first start event
    dev_ms, _, launches, clocks, wall = job.timed(step, args.steps, args.warmup)
    per_build_ms = dev_ms / (args.steps * reps)
    gbps = nbytes / per_build_ms / 1e6
    job.barrier()
    t0 = time.perf_counter()
    P = compute_p_tensor_batch(env)                    # the reference-named call: NumPy array out (773 MB device-to-host)
    Rt = compute_r_table(env)
    e2e_s = job.max_over_ranks(time.perf_counter() - t0)
    if rank != 0:
        return job.emit(None)
    hbm = float(job.peaks.get("hbm_gbs", 6536.7))
    line = {"metric": "tabular P/R table build, bytes written per second", "value": gbps * world, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(3, world, args), "clocks": clocks,
            "e2e": {"value": nbytes / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(8 * (env.n_states + env.n_actions) + env.n_states),
                    "d2h_bytes_per_step": int(P.nbytes + Rt.nbytes),
                    "note": "compute_p_tensor_batch(env) + compute_r_table(env) returning NumPy arrays like the reference: the 773 MB "
                            "device-to-host copy dominates"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": gbps, "peak": hbm, "unit": "GB/s", "frac": gbps / hbm,
                         "frac_of_8TBs_nominal": gbps / 8000.0, "traffic": None,
                         "note": f"algorithmic bytes {nbytes} per build (8 Ns^2 Na + 8 Ns Na), {reps} builds per step, CUDA events; peak = "
                                 "measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)"},
            "cpu_baseline": None, "extra": {"ms_per_build": per_build_ms, "cdf_path": "erf/erfc" if args.exact_cdf else "gauss-legendre",
                                            "colsum_ok": bool(np.isclose(P.sum(axis=0), 1).all())}}
    return job.emit(line)


# --------------------------------------------------------------------------------------------------
# C4: d = 10 training step with the gradient exchange
# --------------------------------------------------------------------------------------------------
def run_config4(args):
    job = Job(args)
    torch = job.torch
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200 import rollout as R
    from rl_sde_is_b200.distributed import Shard
    from rl_sde_is_b200.environments import DoubleWellStoppingTimeND
    from rl_sde_is_b200.reinforce_deterministic_core import _loss_and_grads_fused
    dev, world, rank = job.dev, job.world, job.rank
    d, lim = 10, 2000
    K = args.trajectories or 1_250_000
    env = DoubleWellStoppingTimeND(d, beta=1.0, alpha=1.0, dt=0.005)
    model = make_policy(1, d=d, head_bias=3.0)
    params = R.flat_parameters(model).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_ALL_GE_LB), L.make_mlp(d, 32)
    shard = Shard(K * world, rank, world)
    ck = args.ckpt_every or 16
    holder = {}

    def step(seed):                                    # device-resident form of the training evaluation: K1 + statistics + K2 + exchange
        fo = R.rollout_forward(env_c, mlp_c, params, K, seed=seed, n_steps_lim=lim, store_path=True, ckpt_every=ck, want_logw=False,
                               traj_offset=shard.traj_offset, K_global=shard.K_global, device=dev)
        g = R.rollout_backward(env_c, mlp_c, params, fo, 1.0 / shard.K_global, device=dev)
        if world > 1:
            from rl_sde_is_b200.distributed import pack_grad_and_stats, reduce_gathered
            g, st = reduce_gathered(shard.all_gather_rows(pack_grad_and_stats(g, fo.stats_dev)), g.numel())
        else:
            st = fo.stats_dev
        holder["g"] = g
        return st

    dev_ms, sts, launches, clocks, wall = job.timed(step, args.steps, args.warmup)
    st = [s.cpu().numpy() for s in sts]
    useful = float(sum(x[L.ST_USEFUL_STEPS] for x in st))
    value = useful / (dev_ms * 1e-3)
    job.barrier()
    e2e_wall, e2e_useful = 0.0, 0.0
    for s in range(min(args.steps, 2)):
        model.zero_grad()
        t0 = time.perf_counter()
        loss, ret, steps = _loss_and_grads_fused(env, model, K, seed=9000 + s, n_steps_lim=lim, ckpt_every=ck, device=dev,
                                                 dist=shard if world > 1 else None)
        e2e_wall += time.perf_counter() - t0
        e2e_useful += float(np.sum(steps))
    e2e_useful = job.sum_over_ranks(e2e_useful)
    e2e_value = e2e_useful / job.max_over_ranks(e2e_wall)
    if rank != 0:
        return job.emit(None)
    P = int(params.size)
    roofline = job.fp32_roofline(useful / world / (dev_ms * 1e-3), flop_train(d, 32), clocks,
                                 f"forward (state checkpoints every {ck} passes) + reverse pass; recomputation is not credited.")
    line = {"metric": METRIC + " (REINFORCE loss + gradient)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(4, world, args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * P + 1400), "d2h_bytes_per_step": int(128 + 4 * P + 8 * K),
                    "note": "the fused training evaluation reinforce() uses (host parameters in; loss, gradient, returns and hit passes "
                            "of this rank's shard out in one copy)"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None,
            "extra": {"K_global": K * world, "mean_steps": useful / args.steps / (K * world), "n_unfinished_last": float(st[-1][L.ST_N_UNFINISHED]),
                      "grad_norm": float(holder["g"].norm().item()), "exchange": "one all-gather of (P + 16) doubles per step" if world > 1 else "none (1 GPU)",
                      "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps}}
    return job.emit(line)


# --------------------------------------------------------------------------------------------------
# C5: metastable rollout
# --------------------------------------------------------------------------------------------------
def run_config5(args):
    job = Job(args)
    torch = job.torch
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200 import rollout as R
    from rl_sde_is_b200.approximate_methods import is_estimate
    from rl_sde_is_b200.distributed import Shard
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    dev, world, rank = job.dev, job.world, job.rank
    K = args.trajectories or 12_500_000
    lim = 10**6
    env = DoubleWellStoppingTime1D(beta=4.0, alpha=1.0, dt=0.001)
    model = make_policy(1)
    params = R.flat_parameters(model).detach().numpy()
    env_c, mlp_c = R.env_struct(env, L.HIT_X0_IN_LB_RB), L.make_mlp(1, 32)
    shard = Shard(K * world, rank, world)

    def run(seed, k):
        sh = Shard(k * world, rank, world)
        out = R.rollout_forward(env_c, mlp_c, params, k, seed=seed, n_steps_lim=lim, stoch_int="exact", want_logw=True,
                                traj_offset=sh.traj_offset, K_global=sh.K_global, tanh=args.tanh, device=dev)
        stats = out.stats_dev
        if world > 1:
            sh.all_reduce_stats(stats)
        return stats

    dev_ms, sts, launches, clocks, wall = job.timed(lambda s: run(s, K), args.steps, args.warmup, warm_fn=lambda s: run(s, max(K // 64, 1)))
    st = [s.cpu().numpy() for s in sts]
    useful = float(sum(x[L.ST_USEFUL_STEPS] for x in st))
    value = useful / (dev_ms * 1e-3)
    summ = R.summarize(st[-1])
    job.barrier()
    e2e_value = None
    if not args.no_e2e:
        t0 = time.perf_counter()
        res = is_estimate(env, model, K, n_steps_lim=lim, seed=5000, tanh=args.tanh, device=dev, dist=shard if world > 1 else None)
        e2e_value = res["useful_steps"] / job.max_over_ranks(time.perf_counter() - t0)
    if rank != 0:
        return job.emit(None)
    roofline = job.fp32_roofline(useful / world / (dev_ms * 1e-3), flop_fwd(1, 32), clocks,
                                 "One launch of the thread-per-trajectory kernel (run to completion) + the warp-per-trajectory kernel for the tail.")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(5, world, args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes + 1400), "d2h_bytes_per_step": 128,
                    "note": "is_estimate(env, model, K, n_steps_lim=1e6), wall clock"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None,
            "extra": {"K_global": K * world, "mean_hit_index": summ.get("mean_hit_index"), "max_hit_index": summ.get("max_hit_index"),
                      "n_unfinished": summ.get("n_unfinished"), "is_mean": summ.get("is_mean"), "is_rel_error": summ.get("is_rel_error"),
                      "expect": "mean hitting pass ~6.95e4, E[exp(-tau)] ~0.0060 (BASELINE.md)", "wall_s_per_step": wall / args.steps}}
    return job.emit(line)


# --------------------------------------------------------------------------------------------------
# secondary measurements attached to the headline line (N = 1 only)
# --------------------------------------------------------------------------------------------------
def reference_cpu_measurements():
    """The reference's own CPU implementation of the other two subsystems, timed in THIS run on the box's host cores."""
    out = {}
    if ref_or_none() is None:
        return {"reference_cpu": {"unavailable": "baseline/_ref/src not present"}}
    try:
        t0 = time.perf_counter()
        cts, steps = cpu_reinforce_reference(20, threads=1)
        w01, nb01, _, _, _ = cpu_tables_reference(0.1)
        w005, nb005, _, _, _ = cpu_tables_reference(0.05)
        out["reference_cpu"] = {
            "host_cores": os.cpu_count(),
            "reinforce_config1": {"iterations": 20, "threads": 1, "iter_per_s_it0": float(1.0 / cts[0]), "iter_per_s_it10_19": float(1.0 / np.mean(cts[10:])),
                                  "seconds_total": float(cts.sum()), "mean_steps_it0": float(steps[0]), "mean_steps_it10_19": float(np.mean(steps[10:])),
                                  "what": "reference reinforce() K=100 lr=1e-2 seed=1 (reinforce_deterministic_core.py:226-265), wall clock cts"},
            "tables_h0.1": {"seconds": w01, "bytes": nb01, "GBps": nb01 / w01 / 1e9},
            "tables_h0.05": {"seconds": w005, "bytes": nb005, "GBps": nb005 / w005 / 1e9,
                             "what": "reference compute_r_table + compute_p_tensor_batch (dynamic_programming.py:3-36), 1 thread; the h = 0.01 build "
                                     "(55 s in the build container) has the same cost per (s, a) column"},
            "seconds_spent": time.perf_counter() - t0}
    except Exception as exc:
        out["reference_cpu"] = {"error": repr(exc)}
    return out


def secondary_measurements(dev):
    """Short timings of the other rows of the hot path (REINFORCE iterations, table build, large-batch training)."""
    import torch
    from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    from rl_sde_is_b200.reinforce_deterministic_core import reinforce
    out = {}
    peaks = load_peaks()
    try:
        env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
        data = reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=30, seed=1, verbose=False, save=False, device=dev)
        cts = data["cts"]
        out["reinforce_config1"] = {"iter_per_s_first": 1.0 / cts[0], "iter_per_s_it10_29": float(1.0 / np.mean(cts[10:])),
                                    "mean_steps_it0": float(data["exp_time_steps"][0]),
                                    "mean_steps_it10_29": float(np.mean(data["exp_time_steps"][10:]))}
        env.set_action_space_bounds(); env.discretize_state_space(0.01); env.discretize_action_space(0.01)
        tab = {}
        bufs = [torch.empty((env.n_states, env.n_states, env.n_actions), dtype=torch.float64, device=dev) for _ in range(2)]
        nbytes = bufs[0].numel() * 8
        reps = 8
        for name, exact in (("gauss_legendre", False), ("erf_erfc", True)):
            ts = []
            for trial in range(3):
                compute_p_tensor_batch(env, out=bufs[1], exact_cdf=exact)     # untimed: keeps the stream busy at the start event
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(reps):                                         # alternating 773 MB outputs: nothing stays in L2
                    compute_p_tensor_batch(env, out=bufs[i & 1], exact_cdf=exact)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / reps)
            ms = float(min(ts[1:]))
            tab[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / float(peaks.get("hbm_gbs", 6536.7))}
        # one Bellman sweep over the resident tensor: HBM-read bound, same timing method
        from rl_sde_is_b200.dynamic_programming import compute_r_table
        from rl_sde_is_b200.tabular_dp_sweeps import DeviceTables
        T = DeviceTables(env, compute_r_table(env, device=dev, device_out=True), bufs[0])
        v = torch.full((env.n_states,), -1.0, dtype=torch.float64, device=dev)
        ts = []
        for trial in range(3):
            T.sweep(v, 1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(16):
                T.sweep(v, 1.0)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 16)
        ms = float(min(ts[1:]))
        out["dp_sweep_config3"] = {"bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6,
                                   "frac_of_hbm_peak": nbytes / ms / 1e6 / float(peaks.get("hbm_gbs", 6536.7)),
                                   "note": "values = R + (1 - d) gamma P^T v over the 773 MB tensor (q/v/policy updates of tabular_dp_*_iteration.py)"}
        del T, bufs
        out["tables_config3"] = {"bytes": nbytes, **tab,
                                 "note": "P (401, 401, 601) float64 device-resident; roofline = HBM write, peak = MEASURED_PEAKS hbm_gbs; "
                                         "CUDA events around 8 back-to-back builds into alternating outputs"}
        # large-batch REINFORCE loss + gradient (K1 with state checkpoints + K2), CUDA-event timed
        from rl_sde_is_b200 import _lib as L2
        from rl_sde_is_b200 import rollout as R2
        env1 = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
        m = make_policy(1, head_bias=0.5)
        params = R2.flat_parameters(m).detach().numpy()
        env_c, mlp_c = R2.env_struct(env1, L2.HIT_ALL_GE_LB), L2.make_mlp(1, 32)
        Kt = 400000
        for it in range(3):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            fo = R2.rollout_forward(env_c, mlp_c, params, Kt, seed=it, n_steps_lim=4000, store_path=True, ckpt_every=1, want_logw=False, device=dev)
            e[1].record()
            R2.rollout_backward(env_c, mlp_c, params, fo, 1.0 / Kt, device=dev)
            e[2].record()
            torch.cuda.synchronize()
        u = float(fo.stats[L2.ST_USEFUL_STEPS])
        f_ms, b_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        out["train_large_batch"] = {"K": Kt, "useful_steps": u, "fwd_ms": f_ms, "bwd_ms": b_ms,
                                    "train_steps_per_s": u / (f_ms + b_ms) * 1e3, "bwd_steps_per_s": u / b_ms * 1e3,
                                    "fp32_frac_train": u / (f_ms + b_ms) * 1e3 * flop_train(1, 32) / 74.45e12,
                                    "note": "6556 FLOP per useful step (forward + reverse, SURVEY 8d)"}
        # wide policy (hidden width 256, the reference's reinforce() default): forward test rollout on the tcgen05 kernel
        # (rollout_umma.cuh) and on the CUDA-core tile kernel, K = 2e5.  Tensor roofline: executed MMA work = 3 products
        # (float16 hi / lo split) x 2 H^2 FLOP per trajectory-step, against the measured bf16 GEMM peak.
        H = 256
        mw = make_policy(1, hidden=H)
        pw = R2.flat_parameters(mw).detach().numpy()
        env_w, mlp_w = R2.env_struct(env1, L2.HIT_X0_IN_LB_RB), L2.make_mlp(1, H)
        wide = {}
        for name in ("umma", "ffma"):
            ts = []
            for it in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ow = R2.rollout_forward(env_w, mlp_w, pw, 200000, seed=it, n_steps_lim=1000, stoch_int="exact", tuning={"wide_kernel": name}, device=dev)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            uw = float(ow.stats[L2.ST_USEFUL_STEPS])
            ms = float(min(ts[1:]))
            wide[name] = {"ms": ms, "steps_per_s": uw / ms * 1e3, "x_fp32_roofline": uw / ms * 1e3 * flop_fwd(1, H) / 74.45e12}
        exec_tflops = wide["umma"]["steps_per_s"] * 3 * 2 * H * H / 1e12
        peak_bf16 = float(peaks.get("bf16_tflops", 1599.4))
        out["wide_policy_h256"] = {"K": 200000, "flop_per_step": flop_fwd(1, H), **wide,
                                   "tensor_roofline": {"bound": "tensor", "achieved": exec_tflops, "peak": peak_bf16, "unit": "TFLOP/s",
                                                       "frac": exec_tflops / peak_bf16,
                                                       "note": "executed tcgen05 work (three f16 MMAs per product) / measured bf16 GEMM peak (MEASURED_PEAKS.json, burst)"},
                                   "note": "K1u (tcgen05, tensor-memory accumulator, streamed float16 hi/lo weights) against K1x (FFMA tile kernel)"}
        # the same policy in training: forward with stored states (K1u) + reverse pass on the tensor cores (K2u: producer /
        # consumer CTAs, rollout_umma_bwd.cuh) against the CUDA-core tile kernel (K2x), K = 6e4, n_steps_lim = 1000
        Kw = 60000
        tw = {}
        fo = None
        for it in range(2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            fo = R2.rollout_forward(env_w, mlp_w, pw, Kw, seed=it, n_steps_lim=1000, store_path=True, ckpt_every=1, want_logw=False,
                                    tuning={"wide_kernel": "umma"}, device=dev)
            e[1].record()
            R2.rollout_backward(env_w, mlp_w, pw, fo, 1.0 / Kw, device=dev)
            e[2].record()
            torch.cuda.synchronize()
            tw = {"fwd_ms": e[0].elapsed_time(e[1]), "bwd_umma_ms": e[1].elapsed_time(e[2])}
        Tw = fo.T
        uw_tr = float((Tw[Tw >= 0] + 1).sum())            # passes the reverse recursion walks (trajectories that hit)
        fo.cfg.wide_kernel = 2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        R2.rollout_backward(env_w, mlp_w, pw, fo, 1.0 / Kw, device=dev)
        b.record()
        torch.cuda.synchronize()
        tw["bwd_ffma_ms"] = a.elapsed_time(b)
        f_rev = 2 * flop_fwd(1, H)                          # two more H x H products per pass than the forward
        out["wide_policy_h256_training"] = {
            "K": Kw, "reverse_steps": uw_tr, **tw,
            "bwd_umma_steps_per_s": uw_tr / tw["bwd_umma_ms"] * 1e3, "bwd_ffma_steps_per_s": uw_tr / tw["bwd_ffma_ms"] * 1e3,
            "bwd_umma_x_fp32_roofline": uw_tr / tw["bwd_umma_ms"] * 1e3 * f_rev / 74.45e12,
            "bwd_ffma_x_fp32_roofline": uw_tr / tw["bwd_ffma_ms"] * 1e3 * f_rev / 74.45e12,
            "note": "reverse pass credit = 2 x the forward's FLOP per step (dH1 = dZ2 W2 and dW2 += dZ2^T H1; the recomputed "
                    "forward product is not credited); K2u runs all three on tcgen05 with float16 hi/lo operands"}
        del fo
    except Exception as exc:
        out["secondary_error"] = repr(exc)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[c-1] (SURVEY C1..C5); 2 = headline")
    ap.add_argument("--tanh", default="precise", choices=["precise", "fast"])
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU (default: the config's own size)")
    ap.add_argument("--ref-trajectories", type=int, default=0, help="trajectories per step of the CPU arm's bounded sample")
    ap.add_argument("--ref-full", action="store_true", help="config 3 reference arm at h = 0.05 instead of 0.1")
    ap.add_argument("--ckpt-every", type=int, default=0, help="config 4: state checkpoint spacing (default 16)")
    ap.add_argument("--exact-cdf", action="store_true", help="config 3: erf/erfc path instead of the quadrature recurrence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-f64", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return {1: run_config1, 2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args)


if __name__ == "__main__":
    sys.exit(main())
