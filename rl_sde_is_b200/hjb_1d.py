"""1-D reference solution of the first-hitting-time control problem.

Host-side stand-in for the external ``sde_hjb_solver`` package the reference loads in
``DoubleWellStoppingTime1D.get_hjb_solver`` (environments.py:392-420): it only provides the reference solution
(``u_opt``, ``value_function``) used for the policy l2 error of ``test_policy_vectorized``
(approximate_methods.py:610-615), for plots and for ``compute_optimal_{v,q}_table``; no hot-path arithmetic.

With running cost f = 1, terminal cost g = 0 and generator L = -V'(x) d/dx + (1/beta) d2/dx2, the quantity of
interest Psi(x) = E_x[exp(-tau)] solves the linear boundary value problem

    (1/beta) Psi'' - V'(x) Psi' - Psi = 0   on the complement of the target set,   Psi = 1 on the target set,

and the optimal control / value function of the stochastic optimal control problem are
u*(x) = sigma Psi'(x) / Psi(x),  V*(x) = -log Psi(x).  Central finite differences on a uniform grid, Neumann
condition at the left end (placed well inside the steep wall of the potential), banded solve.
"""
import numpy as np
from scipy.linalg import solve_banded


class HJBSolution1D:
    def __init__(self, env, h=1e-3, x_min=-3.0):
        beta, alpha, lb = float(env.beta), float(env.alpha), float(env.lb)
        n = int(round((lb - x_min) / h))
        x = x_min + h * np.arange(n + 1)                      # x[n] = lb
        grad_v = 4.0 * alpha * x * (x * x - 1.0)
        eps = 1.0 / beta
        # unknowns Psi_0 .. Psi_{n-1}; Psi_n = 1
        lower = eps / h**2 + grad_v / (2 * h)                 # coefficient of Psi_{i-1}
        diag = -2 * eps / h**2 - 1.0 + 0 * x
        upper = eps / h**2 - grad_v / (2 * h)                 # coefficient of Psi_{i+1}
        ab = np.zeros((3, n))
        ab[0, 1:] = upper[:n - 1]
        ab[1, :] = diag[:n]
        ab[2, :-1] = lower[1:n]
        ab[0, 1] += lower[0]                                  # Neumann at x_min: ghost Psi_{-1} = Psi_{1}
        rhs = np.zeros(n)
        rhs[-1] = -upper[n - 1]                               # known Psi_n = 1
        psi = np.append(solve_banded((1, 1), ab, rhs), 1.0)
        self.x, self.psi = x, psi
        self.sigma = float(env.sigma)
        dpsi = np.gradient(psi, h)
        self.u_opt_fine = self.sigma * dpsi / psi
        self.value_fine = -np.log(psi)
        self.lb = lb

    def psi_at(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x >= self.lb, 1.0, np.interp(x, self.x, self.psi))

    def u_opt_at(self, x):
        """Optimal control on arbitrary points (0 on the target set, like the reference's solution)."""
        x = np.asarray(x, dtype=np.float64)
        return np.where(x >= self.lb, 0.0, np.interp(x, self.x, self.u_opt_fine))

    def value_function_at(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x >= self.lb, 0.0, np.interp(x, self.x, self.value_fine))

    def policy_opt_table(self, env):
        """``policy_opt`` argument of test_policy_vectorized / reinforce: shape (n_states, 1) on env.state_space_h."""
        return self.u_opt_at(env.state_space_h).reshape(-1, 1)
