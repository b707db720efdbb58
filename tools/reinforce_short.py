import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
from rl_sde_is_b200.reinforce_deterministic_core import reinforce
d = reinforce(DoubleWellStoppingTime1D(), d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=40, seed=1, verbose=False, save=False)
print("ok", d["cts"][20:].mean(), d["time_steps"].reshape(40, 100)[20:].max(axis=1).mean())
