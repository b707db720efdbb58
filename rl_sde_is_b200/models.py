"""Policy containers with the reference's parameter layout.

Reference: rl_sde_is/models.py:4-18 (``mlp``) and reinforce_deterministic_core.py:13-28
(``DeterministicPolicy``).  The module is only a parameter container + a plain torch forward for
host-side use (plots, tests); the rollouts evaluate the policy inside the fused CUDA kernels, reading
the parameters in state_dict order: ``policy.0.weight (H,d)``, ``policy.0.bias (H)``,
``policy.2.weight (H,H)``, ``policy.2.bias (H)``, ``policy.4.weight (d,H)``, ``policy.4.bias (d)``.
"""
import torch.nn as nn


def mlp(sizes, activation, output_activation=nn.Identity()):
    """Linear / activation stack; the last Linear is followed by ``output_activation``."""
    n = len(sizes) - 1
    stack = []
    for j, (fan_in, fan_out) in enumerate(zip(sizes[:-1], sizes[1:])):
        stack.append(nn.Linear(fan_in, fan_out))
        stack.append(output_activation if j == n - 1 else activation)
    return nn.Sequential(*stack)


class DeterministicPolicy(nn.Module):
    """a = mu_theta(x); head initialised U(-5e-3, 5e-3) so the initial control is ~0."""

    HEAD_INIT = 5e-3

    def __init__(self, state_dim, action_dim, hidden_sizes, activation):
        super().__init__()
        self.sizes = [state_dim, *hidden_sizes, action_dim]
        self.policy = mlp(sizes=self.sizes, activation=activation)
        # the reference re-initialises every Linear whose out_features equals the action dimension
        for layer in self.policy:
            if isinstance(layer, nn.Linear) and layer.out_features == self.sizes[-1]:
                nn.init.uniform_(layer.weight, -self.HEAD_INIT, self.HEAD_INIT)
                nn.init.uniform_(layer.bias, -self.HEAD_INIT, self.HEAD_INIT)

    def forward(self, state):
        return self.policy(state)
