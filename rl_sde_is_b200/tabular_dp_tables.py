"""Table build + check, the hot part of rl_sde_is/tabular_dp_tables.py (:15-17, :44-58).

``compute_optimal_v_table`` / ``compute_optimal_q_table`` (:19-42) need the external HJB reference
solution (sde_hjb_solver) and are a "next" row of SURVEY section 8f, not built here.
"""
import numpy as np
import torch

from .dynamic_programming import compute_p_tensor_batch, compute_r_table, p_tensor_column_sums


def check_p_tensor(env, p_tensor):
    """Every (state, action) column sums to 1 over the next-state axis (np.isclose tolerances)."""
    if torch.is_tensor(p_tensor) and p_tensor.is_cuda:
        sums = p_tensor_column_sums(p_tensor).cpu().numpy()
    else:
        sums = np.sum(np.asarray(p_tensor), axis=0)
    return bool(np.isclose(sums, 1).all())


def dynamic_programming_tables(env, value_function_opt=None, policy_opt=None, load=False, *, device=None, device_out=False):
    """Build ``r_table`` and ``p_tensor`` and assert the column-sum check, like the reference (:54-58).
    The Bellman-expectation tables under the HJB policy (:60-61) are left to the caller."""
    if load:
        raise NotImplementedError("loading reference run directories is outside the hot-path scope (SURVEY 8f-2)")
    r_table = compute_r_table(env, device=device, device_out=device_out)
    p_tensor = compute_p_tensor_batch(env, device=device, device_out=device_out)
    assert check_p_tensor(env, p_tensor)
    return {"r_table": r_table, "p_tensor": p_tensor}
