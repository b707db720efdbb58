"""Drop-in names of rl_sde_is/tabular_dp_policy_iteration.py:37-49 on the GPU sweep kernel (see tabular_dp_sweeps.py)."""
from .tabular_dp_sweeps import policy_update_vect  # noqa: F401
