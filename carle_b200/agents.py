"""Input generators for rollouts (reference carle/agents.py:15-42).

``RandomAgent`` draws Bernoulli(toggle_rate) toggles shaped ``[N, 1, action_width,
action_height]`` exactly like the reference (CPU RNG by default, so a seeded rollout
feeds both implementations identical actions); pass ``device="cuda"`` to generate where
the environment lives and skip the per-step host->device copy."""
import torch
import torch.nn as nn


class RandomAgent(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.action_width = kwargs.get("action_width", 64)
        self.action_height = kwargs.get("action_height", 64)
        self.observation_width = kwargs.get("observation_width", 256)
        self.observation_height = kwargs.get("observation_height", 256)
        self.device = kwargs.get("device", None)
        self.toggle_rate = 0.100

    def forward(self, obs):
        instances = obs.shape[0]
        noise = torch.rand(instances, 1, self.action_width, self.action_height,
                           device=self.device)
        return 1.0 * (noise <= self.toggle_rate)
