#!/usr/bin/env python
"""Headline benchmark: useful Euler-Maruyama trajectory-steps per second of the fused rollout.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, weak scaling

Workload = BASELINE.json configs[1]: 1-D double well (alpha = 1, beta = 1, dt = 0.005, x0 = -1), policy
DeterministicPolicy(1, 1, [32, 32], Tanh) at its seed-1 initialisation, 1e6 trajectories per GPU,
n_steps_lim = 1000, importance-sampling estimator + relative error from the same rollout.  A "step" is
one pass of the hot path over one batch: one launch of the rollout kernel for all K trajectories plus
the statistics reduction.  Unit of work: one useful trajectory-step = one pass of one
trajectory up to and including the pass on which its hit is detected (or n_steps_lim if it is not).

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events with everything resident on the GPU;
`e2e` is the same metric through the public Python call (host parameters in, statistics out, wall clock).
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
FLOP_PER_STEP = 2 * (2 * 1 * 32 + 32 * 32) + 14 * 1 + 4          # 2194 (forward rollout, d = 1, H = 32)
METRIC = "SDE trajectory-steps/sec"
UNIT = "trajectory-steps/s"


def make_policy(seed):
    import torch
    import torch.nn as nn
    from rl_sde_is_b200.models import DeterministicPolicy
    np.random.seed(seed)
    torch.manual_seed(seed)
    return DeterministicPolicy(1, 1, [WORKLOAD["hidden"]] * 2, nn.Tanh())


# --------------------------------------------------------------------------------------------------
# CPU arms (oracle port).  Bounded samples of the same workload.
# --------------------------------------------------------------------------------------------------
def cpu_port_python(K, seed=0):
    """The reference's algorithm as the reference implements it: one torch forward + one NumPy
    Euler-Maruyama pass per time step over the whole batch (oracle/reference_semantics.py)."""
    import torch
    from oracle import reference_semantics as ref
    model = make_policy(WORKLOAD["policy_seed"])
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    lim = WORKLOAD["n_steps_lim"]
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    noise = (np.sqrt(WORKLOAD["dt"]) * rng.standard_normal((lim, K, 1))).astype(np.float32)
    st = ref.rollout_stats_numpy(1, WORKLOAD["alpha"], WORKLOAD["beta"], WORKLOAD["dt"], params, noise)
    dt_wall = time.perf_counter() - t0
    lens = st["ep_lens"]
    useful = int(np.where(lens >= 0, lens + 1, lim).sum())
    return useful / dt_wall, useful, dt_wall, int(torch.get_num_threads())


def cpu_port_c(K, seed=0):
    """Plain-C restatement, OpenMP over trajectories (a stronger CPU arm than the reference's Python)."""
    from oracle import c_oracle
    from rl_sde_is_b200 import rollout as R
    params = R.flat_parameters(make_policy(WORKLOAD["policy_seed"])).detach().numpy()
    t0 = time.perf_counter()
    out = c_oracle.rollout(1, 32, params, WORKLOAD["alpha"], WORKLOAD["beta"], WORKLOAD["dt"], K, seed=seed,
                           n_steps_lim=WORKLOAD["n_steps_lim"], hit_rule=c_oracle.HIT_X0_IN_LB_RB, stoch_int_exact=True)
    dt_wall = time.perf_counter() - t0
    return out["useful_steps"] / dt_wall, out["useful_steps"], dt_wall, c_oracle.num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    K = 50000                      # ~3 s of host work per step: larger batches vectorise better, to the reference's advantage
    for _ in range(args.warmup):
        cpu_port_python(2000)
    vals, useful_tot, wall_tot, threads = [], 0, 0.0, 1
    for s in range(args.steps):
        v, useful, wall, threads = cpu_port_python(K, seed=s)
        vals.append(v); useful_tot += useful; wall_tot += wall
    value = useful_tot / wall_tot
    sample = f"{K} trajectories x n_steps_lim {WORKLOAD['n_steps_lim']} per step (of the 1e6-trajectory workload), numpy/torch per-pass loop"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall_tot / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 state / f32 policy", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {"workload": "configs[1]: 1-D double well alpha=1 beta=1 dt=0.005, policy test rollout, 1e6 trajectories per GPU, "
                        "n_steps_lim=1000, IS estimator + relative error",
            "trajectories_per_gpu": WORKLOAD["K_per_gpu"], "n_steps_lim": WORKLOAD["n_steps_lim"], "policy": "MLP 1-32-32-1 tanh, seed-1 init",
            "parallelism": f"trajectories sharded over {n_gpus} GPU(s), statistics all-reduced",
            "l2": "kernel inputs are by-value parameters only (no reusable HBM inputs); outputs are write-only; "
                  "a 256 MiB flush write runs between steps, outside the timed events"}


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
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from rl_sde_is_b200 import _lib as L
    from rl_sde_is_b200 import rollout as R
    from rl_sde_is_b200.approximate_methods import is_estimate
    from rl_sde_is_b200.distributed import Shard
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE line, the JSON record: everything libraries print on file descriptor 1 meanwhile (NCCL's
    # version banner comes from C code, whatever NCCL_DEBUG or /etc/nccl.conf say) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    env = DoubleWellStoppingTime1D(beta=WORKLOAD["beta"], alpha=WORKLOAD["alpha"], dt=WORKLOAD["dt"])
    model = make_policy(WORKLOAD["policy_seed"])
    params = R.flat_parameters(model).detach().numpy()
    K = args.trajectories or WORKLOAD["K_per_gpu"]
    shard = Shard(K * world, rank, world)
    env_c, mlp_c = R.env_struct(env, L.HIT_X0_IN_LB_RB), L.make_mlp(1, WORKLOAD["hidden"])
    lim = WORKLOAD["n_steps_lim"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one_step(seed):
        out = R.rollout_forward(env_c, mlp_c, params, K, seed=seed, n_steps_lim=lim, stoch_int="exact", want_logw=True,
                                traj_offset=shard.traj_offset, K_global=shard.K_global, tanh=args.tanh, device=dev)
        stats = out.stats_dev
        if world > 1:
            shard.all_reduce_stats(stats)               # the path's only exchange step: 16 doubles
        return stats

    for w in range(args.warmup):
        one_step(1000 + w)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stats_steps = []
    launches0 = L.load().rlsde_launch_count()
    with ClockSampler(local_rank) as clk:
        wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)                       # L2 flush, outside the timed events
            ev[s][0].record()
            stats_steps.append(one_step(s))
            ev[s][1].record()
        torch.cuda.synchronize()
        wall1 = time.perf_counter()
    gpu_launches = int(L.load().rlsde_launch_count() - launches0)    # this library's kernels inside the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    st = [s.cpu().numpy() for s in stats_steps]        # already global sums when world > 1
    useful = float(sum(x[L.ST_USEFUL_STEPS] for x in st))
    value = useful / (dev_ms * 1e-3)
    summ = R.summarize(st[-1])

    # ---- end to end through the public API: host parameters in, statistics out, wall clock
    e2e_wall, e2e_useful = 0.0, 0.0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for s in range(args.steps):
        t0 = time.perf_counter()
        res = is_estimate(env, model, K, n_steps_lim=lim, seed=5000 + s, tanh=args.tanh, device=dev,
                          dist=shard if world > 1 else None)
        e2e_wall += time.perf_counter() - t0
        e2e_useful += res["useful_steps"]
    tt = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = e2e_useful / float(tt.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (rollout_fwd_kernel): FP32 CUDA-core FMA throughput
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", clk.summary()["sm_max_mhz"] or 1965.0))
    peak_tflops = sm * 128 * 2 * sm_max * 1e6 / 1e12
    per_gpu_steps = useful / n_gpus / (dev_ms * 1e-3)
    achieved = per_gpu_steps * FLOP_PER_STEP / 1e12
    roofline = {"bound": "fp32_cuda_core", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": achieved / peak_tflops, "traffic": 2.199e9, "traffic_unit": "bytes per launch (dram read + write, ncu --set full, profiles/r01/ncu_rollout_fwd_v2.txt): "
                                                   "the continuation records of the time-sliced schedule, 24 B per 8 passes; 86 GB/s",
                "note": f"peak = {sm} SMs x 128 FMA lanes x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; FFMA "
                        f"microbenchmark in profiles/ confirms 128 lanes/clk/SM); algorithmic work {FLOP_PER_STEP} FLOP per useful "
                        "trajectory-step (SURVEY 8d), timed per launch with CUDA events incl. the statistics reduction; the "
                        "kernel keeps its state in registers (results: 16 B per trajectory)"}

    extra = {"is_mean": summ.get("is_mean"), "is_rel_error": summ.get("is_rel_error"), "mean_return": summ.get("mean_return"),
             "frac_unfinished": summ["n_unfinished"] / max(summ["n"], 1), "useful_steps_per_step": useful / args.steps,
             "wall_ms_per_step_incl_flush": 1e3 * (wall1 - wall0) / args.steps, "tanh": args.tanh}
    cpu_baseline = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        cpu_port_python(2000)                             # warm-up (thread pools, allocator)
        ck = 100000
        v, u, w, thr = cpu_port_python(ck)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "host_cores": os.cpu_count(),
                        "sample": f"{ck} trajectories x n_steps_lim {lim} ({u} useful steps, {w:.1f} s): the reference's per-pass "
                                  "torch forward + NumPy step loop (oracle/reference_semantics.py)"}
        try:
            vc, uc, wc, thc = cpu_port_c(200000)
            extra["cpu_c_port"] = {"value": vc, "unit": UNIT, "cores": thc, "sample": f"200000 trajectories ({uc} useful steps, {wc:.1f} s), "
                                   "plain C + OpenMP (oracle/rlsde_oracle.c)"}
        except Exception as exc:                         # the C oracle is optional test infrastructure
            extra["cpu_c_port"] = {"error": str(exc)}
        if not args.no_extras:
            extra.update(secondary_measurements(dev))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus),
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes + 1400),
                "d2h_bytes_per_step": int(L.RLSDE_NSTATS * 8),
                "note": "is_estimate(env, model, K): host policy parameters travel as kernel arguments, the 16-double statistics "
                        "record comes back; wall clock around the call"},
        "gpu_launches": gpu_launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extra,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def secondary_measurements(dev):
    """Short timings of the other rows of the hot path (REINFORCE iterations, table build): reported, not the headline."""
    import torch
    from rl_sde_is_b200.dynamic_programming import compute_p_tensor_batch
    from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
    from rl_sde_is_b200.reinforce_deterministic_core import reinforce
    out = {}
    try:
        env = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
        data = reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=30, seed=1, verbose=False, save=False, device=dev)
        cts = data["cts"]
        out["reinforce_config0"] = {"iter_per_s_first": 1.0 / cts[0], "iter_per_s_it10_29": float(1.0 / np.mean(cts[10:])),
                                    "mean_steps_it0": float(data["exp_time_steps"][0]),
                                    "mean_steps_it10_29": float(np.mean(data["exp_time_steps"][10:])),
                                    "reference_cpu": "0.14 it/s at it.0, ~2.4 it/s at it.10-20 (BASELINE.md, 8-core Xeon)"}
        env.set_action_space_bounds(); env.discretize_state_space(0.01); env.discretize_action_space(0.01)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
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
        out["dp_sweep_config2"] = {"bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6,
                                   "frac_of_hbm_peak": nbytes / ms / 1e6 / float(peaks.get("hbm_gbs", 6536.7)),
                                   "note": "values = R + (1 - d) gamma P^T v over the 773 MB tensor (q/v/policy updates of "
                                           "tabular_dp_*_iteration.py); reference ~0.3 s per sweep in NumPy"}
        del T, bufs
        out["tables_config2"] = {"bytes": nbytes, **tab, "reference_cpu_s": 55.3,
                                 "note": "P (401, 401, 601) float64 device-resident; roofline = HBM write, peak = MEASURED_PEAKS hbm_gbs; "
                                         "CUDA events around 8 back-to-back builds into alternating outputs"}
        # large-batch REINFORCE loss + gradient (K1 with state checkpoints + K2), CUDA-event timed
        from rl_sde_is_b200 import _lib as L2
        from rl_sde_is_b200 import rollout as R2
        env1 = DoubleWellStoppingTime1D(beta=1.0, alpha=1.0, dt=0.005)
        m = make_policy(1)
        m.policy[4].bias.data.fill_(0.5)
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
                                    "fp32_frac_train": u / (f_ms + b_ms) * 1e3 * 6556 / 74.45e12,
                                    "note": "6556 FLOP per useful step (forward + reverse, SURVEY 8d)"}
    except Exception as exc:
        out["secondary_error"] = repr(exc)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tanh", default="precise", choices=["precise", "fast"])
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories per GPU (default: the workload's 1e6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
