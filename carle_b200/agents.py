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


class DeviceRandomAgent(nn.Module):
    """The same Bernoulli(toggle_rate) policy generated where the environment lives, directly
    in the packed action layout (``CARLE.random_action``): no float tensor, no host->device
    copy.  ``forward(obs)`` returns a ``PackedAction`` that ``CARLE.step`` accepts; call
    ``.to_float()`` on it for code that wants the reference's float32 ``[N,1,aw,ah]``."""

    def __init__(self, env, toggle_rate=0.100, seed=0, lazy=True):
        super().__init__()
        self.env = env.inner_env if getattr(env, "inner_env", None) is not None else env
        self.toggle_rate, self.seed, self.calls, self.lazy = toggle_rate, seed, 0, lazy

    def forward(self, obs=None):
        """``lazy`` (default): return the recipe ``RandomAction(seed, call#, rate)`` — the
        environment draws the toggles inside its step kernel; otherwise materialise the
        (identical) packed toggles now."""
        from .env import RandomAction
        if self.lazy:
            self.env._ensure_handle()
            action = RandomAction(self.env, self.seed, self.calls, self.toggle_rate,
                                  self.env.instances)
        else:
            action = self.env.random_action(self.seed, self.calls, self.toggle_rate)
        self.calls += 1
        return action
