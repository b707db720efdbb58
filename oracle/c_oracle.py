"""ctypes loader for oracle/librlsde_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from rl_sde_is_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "librlsde_oracle.so")
SRC_PATH = os.path.join(_DIR, "rlsde_oracle.c")

OF_NOISE_INJECTED, OF_STOCH_INT_EXACT, OF_STATE_F64 = 1, 4, 8
HIT_ALL_GE_LB, HIT_X0_IN_LB_RB = 0, 1

_lib = None


def build(force=False):
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC_PATH):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", LIB_PATH, SRC_PATH, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        vp, i64, u64, dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_double
        lib.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        lib.oracle_noise_fill.argtypes = [u64, i64, i64, C.c_int, i64, i64, dbl, vp]
        lib.oracle_rollout.restype = i64
        lib.oracle_rollout.argtypes = [C.c_int, C.c_int, vp, vp, dbl, dbl, dbl, dbl, C.c_int, vp, i64, i64, i64, u64, i64, i64,
                                       C.c_int, vp, vp, i64, dbl, dbl, dbl, vp, vp, vp, vp, vp]
        lib.oracle_tables.argtypes = [vp, i64, vp, i64, vp, i64, dbl, dbl, dbl, dbl, i64, i64, vp, vp]
        lib.oracle_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    load().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def noise_fill(seed, K, d, n_pass, dt, traj_offset=0, pass_begin=0):
    out = np.empty((n_pass, K, d), dtype=np.float32)
    load().oracle_noise_fill(int(seed), int(traj_offset), int(K), int(d), int(pass_begin), int(n_pass), float(dt), out.ctypes.data)
    return out


def rollout(d, H, params, alpha, beta, dt, K, *, seed=0, n_steps_lim=10**6, noise=None, traj_offset=0, K_global=None,
            hit_rule=HIT_ALL_GE_LB, state_f64=False, stoch_int_exact=False, policy_opt=None, grid=None, lb=1.0, rb=2.0,
            x0=-1.0):
    """Per-trajectory (G, S, T, l2, logw) and the useful-pass count, from the C restatement."""
    params = np.ascontiguousarray(params, dtype=np.float32)
    al = np.full(d, float(alpha), dtype=np.float64)
    x0v = np.full(d, float(x0), dtype=np.float64)
    flags = 0
    noise_steps = 0
    nz = None
    if noise is not None:
        nz = np.ascontiguousarray(noise, dtype=np.float32)
        flags |= OF_NOISE_INJECTED
        noise_steps = nz.shape[0]
    if stoch_int_exact:
        flags |= OF_STOCH_INT_EXACT
    real = np.float64 if state_f64 else np.float32
    if state_f64:
        flags |= OF_STATE_F64
    G, S, l2, logw = (np.zeros(K, dtype=real) for _ in range(4))
    T = np.zeros(K, dtype=np.int32)
    pol, n_grid, lo, hi, h = None, 0, 0.0, 0.0, 1.0
    if policy_opt is not None:
        pol = np.ascontiguousarray(np.asarray(policy_opt, dtype=np.float32).reshape(-1))
        n_grid = pol.size
        lo, hi, h = grid
    useful = load().oracle_rollout(d, H, params.ctypes.data, al.ctypes.data, float(np.sqrt(2.0 / beta)), float(dt), lb, rb,
                                   hit_rule, x0v.ctypes.data, K, traj_offset, K_global if K_global is not None else K,
                                   int(seed), int(n_steps_lim), noise_steps, flags, nz.ctypes.data if nz is not None else None,
                                   pol.ctypes.data if pol is not None else None, n_grid, lo, hi, h, G.ctypes.data, S.ctypes.data,
                                   T.ctypes.data, l2.ctypes.data, logw.ctypes.data)
    return dict(G=G, S=S, T=T, l2=l2, logw=logw, useful_steps=int(useful))


def tables(state_grid, action_grid, is_in_ts, alpha, beta, dt, h_state, sprime_range=None, want_p=True, want_r=True):
    sg = np.ascontiguousarray(state_grid, dtype=np.float64)
    ag = np.ascontiguousarray(action_grid, dtype=np.float64)
    ts = np.ascontiguousarray(is_in_ts, dtype=np.uint8)
    Ns, Na = sg.size, ag.size
    lo, hi = (0, Ns) if sprime_range is None else sprime_range
    P = np.empty((hi - lo, Ns, Na), dtype=np.float64) if want_p else None
    R = np.empty((Ns, Na), dtype=np.float64) if want_r else None
    load().oracle_tables(sg.ctypes.data, Ns, ag.ctypes.data, Na, ts.ctypes.data, int(ts.sum()), float(alpha),
                         float(np.sqrt(2.0 / beta)), float(dt), float(h_state) / 2.0, lo, hi,
                         P.ctypes.data if P is not None else None, R.ctypes.data if R is not None else None)
    return P, R


def num_threads():
    return load().oracle_num_threads()
