"""Import of the UNMODIFIED reference (riberaborrell/rl-sde-is v1.0.0) -- TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python.  It is importable with three stubs: ``rl_sde_is.config`` (a
git-ignored file the reference expects the user to create from ``config_template.py:3-7``), ``matplotlib*`` and
``shapely*`` (plotting only; not installed in this image).  No source line of the reference is changed or copied into
the repository: the sources are read from

  * ``/root/reference/src``              in the build container, or
  * ``<repo>/baseline/_ref/src``         a git-ignored copy made by ``__graft_entry__.build()`` (``make_ref_copy``);
                                          it travels to the GPU box with the gpurun snapshot, where
                                          ``/root/reference`` does not exist.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU arms may import this module; nothing under
``rl_sde_is_b200/`` does.
"""
import importlib
import os
import shutil
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref", "src")
REF_ORIGINAL = "/root/reference/src"

_loaded = None


def ref_src():
    """Directory holding the reference's ``rl_sde_is`` package, or None."""
    for cand in (REF_ORIGINAL, REF_COPY):
        if os.path.isfile(os.path.join(cand, "rl_sde_is", "environments.py")):
            return cand
    return None


def available():
    return ref_src() is not None


def make_ref_copy():
    """Copy the reference's Python package next to the repo's build products (git-ignored) so that the CPU arm of the
    bench and the drop-in tests can run it on the GPU box.  No-op without ``/root/reference``."""
    src = os.path.join(REF_ORIGINAL, "rl_sde_is")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(REF_COPY, "rl_sde_is")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(REF_COPY, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return dst


def load(data_dir=None):
    """Import the reference's modules; returns a namespace with ``environments, environments_2d, core
    (reinforce_deterministic_core), approx (approximate_methods), dp (dynamic_programming), tables (tabular_dp_tables),
    models, replay_buffers, utils_path, src``.  Raises RuntimeError if the reference is not present."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = ref_src()
    if src is None:
        raise RuntimeError("reference sources not found (neither /root/reference/src nor baseline/_ref/src)")
    from unittest.mock import MagicMock
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "shapely", "shapely.geometry"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    tmp = data_dir or tempfile.mkdtemp(prefix="rlsde_ref_")
    cfg = types.ModuleType("rl_sde_is.config")
    cfg.PROJECT_ROOT_DIR = tmp
    cfg.DATA_ROOT_DIR = os.path.join(tmp, "data")
    sys.modules["rl_sde_is.config"] = cfg
    if src not in sys.path:
        sys.path.insert(0, src)
    ns = types.SimpleNamespace(src=src, data_dir=cfg.DATA_ROOT_DIR)
    ns.environments = importlib.import_module("rl_sde_is.environments")
    ns.environments_2d = importlib.import_module("rl_sde_is.environments_2d")
    ns.models = importlib.import_module("rl_sde_is.models")
    ns.dp = importlib.import_module("rl_sde_is.dynamic_programming")
    ns.replay_buffers = importlib.import_module("rl_sde_is.replay_buffers")
    ns.core = importlib.import_module("rl_sde_is.reinforce_deterministic_core")
    ns.approx = importlib.import_module("rl_sde_is.approximate_methods")
    ns.tables = importlib.import_module("rl_sde_is.tabular_dp_tables")
    ns.utils_path = importlib.import_module("rl_sde_is.utils_path")
    _loaded = ns
    return ns


class UsefulStepCounter:
    """Wraps ``env.step`` / ``env.step_torch`` of a reference environment INSTANCE (the class and its source stay
    untouched) and counts useful trajectory-steps the way SURVEY 8d defines them: every pass a trajectory executes up to
    and including the one on which its hit is detected.  Three small boolean operations per pass."""

    def __init__(self, env, attr="step"):
        self.env, self.attr, self.inner = env, attr, getattr(env, attr)
        self.useful, self.seen, self.passes = 0, None, 0
        setattr(env, attr, self)

    def __call__(self, *args, **kwargs):
        out = self.inner(*args, **kwargs)
        done = out[2]
        done = done.numpy() if hasattr(done, "numpy") else done
        if self.seen is None:
            self.seen = done.copy()
            self.seen[:] = False
        self.useful += int((~self.seen).sum())
        self.seen |= done
        self.passes += 1
        return out

    def reset(self):
        self.useful, self.seen, self.passes = 0, None, 0

    def remove(self):
        try:
            delattr(self.env, self.attr)       # the instance attribute shadows the class method
        except AttributeError:
            pass
