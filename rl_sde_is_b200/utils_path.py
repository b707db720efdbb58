"""On-disk formats of the reference: run directories, ``agent.npz`` and model backups.

Reference: rl_sde_is/utils_path.py
  * ``save_data`` / ``load_data`` (:53-68): ``np.savez`` of the result dictionary into ``<data>/<rel_dir>/agent.npz``
    (objects such as the policy module are pickled); loading unwraps 0-d entries with ``.item()`` and adds ``rel_dir_path``.
  * ``save_model`` / ``load_model`` (:70-77): ``torch.save(model.state_dict())`` as ``<rel_dir>/<file_name>``
    (``model_n-it{i}``, reinforce_deterministic_core.py:95-99,169,293).
  * directory names (:79-98, :154-162, :300-325): ``<env.name>/<algorithm>/<parameter string>``.
Files written here are read by the reference's ``--load`` / ``--plot`` code and vice versa (state_dict keys
``policy.{0,2,4}.{weight,bias}`` are identical; a pickled module needs the writing package importable on the reader).

The reference takes its data root from an untracked ``rl_sde_is.config`` module (config_template.py:3-7); here it is
``set_data_dir(path)``, else ``$RLSDE_DATA_DIR``, else ``./data``.
"""
import os
import shutil

import numpy as np
import torch

_data_dir = None


def set_data_dir(path):
    global _data_dir
    _data_dir = os.fspath(path)


def get_data_dir():
    return _data_dir or os.environ.get("RLSDE_DATA_DIR") or os.path.join(os.getcwd(), "data")


def make_dir_path(dir_path):
    os.makedirs(dir_path, exist_ok=True)


def empty_dir(dir_path):
    if not os.path.isdir(dir_path):
        return
    for entry in os.scandir(dir_path):
        if entry.is_dir(follow_symlinks=False):
            shutil.rmtree(entry.path)
        else:
            os.unlink(entry.path)


def _abs(rel_dir_path, file_name):
    return os.path.join(get_data_dir(), rel_dir_path, file_name)


def save_data(data_dict, rel_dir_path):
    make_dir_path(os.path.join(get_data_dir(), rel_dir_path))
    np.savez(_abs(rel_dir_path, "agent.npz"), **data_dict)


def load_data(rel_dir_path):
    """Raises FileNotFoundError when there is no ``agent.npz`` (the reference prints the error and ``sys.exit()``s)."""
    with np.load(_abs(rel_dir_path, "agent.npz"), allow_pickle=True) as npz:
        data = {k: (npz[k].item() if npz[k].ndim == 0 else npz[k]) for k in npz.files}
    data["rel_dir_path"] = rel_dir_path
    return data


def save_model(model, rel_dir_path, file_name):
    make_dir_path(os.path.join(get_data_dir(), rel_dir_path))
    torch.save(model.state_dict(), _abs(rel_dir_path, file_name))


def load_model(model, rel_dir_path, file_name):
    model.load_state_dict(torch.load(_abs(rel_dir_path, file_name)))


# ---------------------------------------------------------------------------------------------- directory names
def get_rel_dir_path(env, algorithm_name, param_str):
    rel = os.path.join(env.name, algorithm_name, param_str)
    make_dir_path(os.path.join(get_data_dir(), rel))
    return rel


def get_initial_point_str(env):
    if env.is_state_init_sampled:
        return "explorable-starts_"
    return "init-state{:2.1f}_".format(env.state_init[0, 0].item())


def get_iter_str(**kwargs):
    if "n_episodes" in kwargs:
        return "n-episodes{:.0e}_".format(kwargs["n_episodes"])
    if "n_total_steps" in kwargs:
        return "n-total-steps{:.0e}_".format(kwargs["n_total_steps"])
    if "n_iterations" in kwargs:
        return "n-iter{:.0e}_".format(kwargs["n_iterations"])
    return ""


def get_seed_str(**kwargs):
    seed = kwargs.get("seed")
    return "seed{:1d}".format(seed) if seed else "seedNone"       # seed 0 also prints as None, like the reference (:135-140)


def get_dynamic_programming_tables_dir_path(env):
    param_str = "h-state{:.0e}_h-action{:.0e}_dt{:.0e}".format(env.h_state, env.h_action, env.dt)
    return get_rel_dir_path(env, "dp-tables", param_str)


def get_dynamic_programming_dir_path(env, **kwargs):
    param_str = "h-state{:.0e}_h-action{:.0e}_dt{:.0e}_n-it{:.0e}".format(env.h_state, env.h_action, env.dt, kwargs["n_iterations"])
    return get_rel_dir_path(env, kwargs["agent"], param_str)


def get_reinforce_det_dir_path(env, **kwargs):
    param_str = get_initial_point_str(env) \
        + "gamma{:.3f}_hidden-size{:d}_".format(kwargs["gamma"], kwargs["d_hidden_layer"]) \
        + "K{:.0e}_lr{:.1e}_".format(kwargs["batch_size"], kwargs["lr"]) \
        + get_iter_str(**kwargs) + get_seed_str(**kwargs)
    return get_rel_dir_path(env, kwargs["agent"], param_str)
