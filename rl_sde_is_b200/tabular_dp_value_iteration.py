"""Drop-in names of rl_sde_is/tabular_dp_value_iteration.py:41-52 on the GPU sweep kernel (see tabular_dp_sweeps.py)."""
from .tabular_dp_sweeps import v_table_update_vect  # noqa: F401
