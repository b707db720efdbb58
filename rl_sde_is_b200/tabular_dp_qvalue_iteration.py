"""Drop-in names of rl_sde_is/tabular_dp_qvalue_iteration.py:35-119 on the GPU sweep kernel (see tabular_dp_sweeps.py)."""
from .tabular_dp_sweeps import q_table_update_vect, qvalue_iteration  # noqa: F401
