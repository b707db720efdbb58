"""cProfile of the REINFORCE iteration loop (config 0) to find host-side overhead."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_sde_is_b200.environments import DoubleWellStoppingTime1D
from rl_sde_is_b200.reinforce_deterministic_core import reinforce
env = DoubleWellStoppingTime1D()
reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=30, seed=1, verbose=False, save=False)   # warm-up
pr = cProfile.Profile(); pr.enable()
data = reinforce(env, d_hidden_layer=32, batch_size=100, lr=1e-2, n_iterations=200, seed=2, verbose=False, save=False)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
print("mean ms/iter (it 20+):", 1e3 * data["cts"][20:].mean(), "mean steps:", data["exp_time_steps"][20:].mean())
