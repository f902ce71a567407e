"""Drop-in for `torchcfm.conditional_flow_matching.ConditionalFlowMatcher` as the reference uses it
(configs/model/conditional_flow_matching.yaml:28-30; src/models/conditional_flow_matching.py:66).

`sample_location_and_conditional_flow(x0, x1, t=None, return_noise=False)` keeps torchcfm's semantics:
t ~ U(0,1)^B drawn from the CPU default generator and moved, eps ~ N(0, I) drawn on x0's device even when sigma = 0
(THIS method advances the device RNG exactly like the reference; the LitModules' fused sigma == 0 training path does not
call it and draws no eps -- see lit.py), xt = t x1 + (1-t) x0 + sigma eps, ut = x1 - x0.
The LitModule's training path does not materialise xt/ut at all (they are fused into the stem operand packing and the
loss kernel); this class exists for API compatibility and for user code that calls it directly.
"""
from __future__ import annotations

import torch


def pad_t_like_x(t, x):
    if isinstance(t, (float, int)):
        return t
    return t.reshape(-1, *([1] * (x.dim() - 1)))


class ConditionalFlowMatcher:
    def __init__(self, sigma: float = 0.0):
        self.sigma = sigma

    def compute_mu_t(self, x0, x1, t):
        t = pad_t_like_x(t, x0)
        return t * x1 + (1 - t) * x0

    def compute_sigma_t(self, t):
        return self.sigma

    def sample_xt(self, x0, x1, t, epsilon):
        return self.compute_mu_t(x0, x1, t) + self.compute_sigma_t(t) * epsilon

    def compute_conditional_flow(self, x0, x1, t, xt):
        return x1 - x0

    def sample_noise_like(self, x):
        return torch.randn_like(x)

    def sample_time(self, x0):
        return torch.rand(x0.shape[0]).type_as(x0)

    def sample_location_and_conditional_flow(self, x0, x1, t=None, return_noise=False):
        if t is None:
            t = self.sample_time(x0)
        assert len(t) == x0.shape[0], "t has to have batch size dimension"
        eps = self.sample_noise_like(x0)
        xt = self.sample_xt(x0, x1, t, eps)
        ut = self.compute_conditional_flow(x0, x1, t, xt)
        if return_noise:
            return t, xt, ut, eps
        return t, xt, ut
