"""torch-CPU port of the reference step — the TIMED CPU BASELINE (test infrastructure).

The reference's hot path is a sequence of torch ops on a float32 ``[N,1,H,W]`` tensor
(carle/env.py:150-242).  ``/root/reference`` does not exist on the GPU box, so
``bench.py`` (``cpu_baseline`` and ``--impl reference``) times this port instead: it
issues the same torch operators in the same order as the reference — ZeroPad2d,
logical_xor, sum / mean host checks, a 3x3 circular ``nn.Conv2d`` with the Moore kernel,
one ``==`` per rule digit folded with ``1.0 * (a + b)``, and the final blend — so its cost
on the host cores is the reference's cost (kind: "port").  ``tests/test_oracle.py`` pins
it bit-for-bit to the golden vectors recorded from the real reference.

Only ``tests/`` and ``bench.py`` import this module; the product never does.
"""
from functools import reduce

import torch
import torch.nn as nn


class TorchPortCARLE:
    def __init__(self, width=256, height=256, action_width=64, action_height=64,
                 instances=1):
        self.width, self.height, self.instances = width, height, instances
        # carle/env.py:119-132
        asym_w = (width - action_width) % 2
        asym_h = (height - action_height) % 2
        self.action_width = action_width - (width % 2)
        self.action_height = action_height - (height % 2)
        wpad = (width - self.action_width) // 2
        hpad = (height - self.action_height) // 2
        self.action_padding = nn.ZeroPad2d((hpad, hpad + asym_h, wpad, wpad + asym_w))
        # carle/env.py:87-116
        self.neighborhood = nn.Conv2d(1, 1, 3, padding=1, padding_mode="circular",
                                      bias=False)
        with torch.no_grad():
            self.neighborhood.weight.copy_(torch.tensor(
                [[[[1., 1., 1.], [1., 0., 1.], [1., 1., 1.]]]]))
        self.neighborhood.weight.requires_grad_(False)
        self.birth, self.survive = [3], [2, 3]
        self.step_number = 0
        self.steps_since_action = 0
        self.universe = None

    def reset(self):                                     # carle/env.py:134-148
        self.universe = torch.zeros(self.instances, 1, self.height, self.width)
        self.step_number = 0
        self.steps_since_action = 0
        return self.universe

    def apply_action(self, action):                      # carle/env.py:150-182
        while action.dim() < 4:
            action = action.unsqueeze(0)
        padded = self.action_padding(action)
        self.universe = 1.0 * torch.logical_xor(self.universe, padded.detach())

    @torch.no_grad()
    def step(self, action):                              # carle/env.py:188-242
        if not torch.sum(action):
            self.steps_since_action += 1
        self.apply_action(action)
        if torch.mean(action) == 1.0:
            obs = self.reset()
        else:
            counts = self.neighborhood(self.universe)
            born = reduce(lambda a, b: 1.0 * (a + b), [d == counts for d in self.birth])
            stay = reduce(lambda a, b: 1.0 * (a + b), [d == counts for d in self.survive])
            self.universe = (1 - self.universe) * born + self.universe * stay
            self.step_number += 1
            obs = self.universe
        reward = torch.zeros(self.instances, 1)
        done = torch.zeros(self.instances, 1)
        return obs, reward, done, [{}] * self.instances


class TorchPortSpeedDetector:
    """torch-CPU port of ``SpeedDetector`` (carle/mcl.py:730-799), the wrapper of the headline
    workload: the same operators in the same order -- ``arange`` weight planes masked by the
    zero-padded action window (:741-758), ``torch.sum(universe)`` / two ``sum(obs * weight)``
    full-grid passes and the divides per step (:773-779), ``cat``, velocity, ``sqrt(sum(pow))``
    and ``reward += speed`` (:781-795)."""

    def __init__(self, env):
        self.env = self.inner_env = env
        self.center_of_mass = None
        self.speed = None
        self.mass_weight_w = torch.arange(env.height).reshape(1, -1)
        self.mass_weight_h = torch.arange(env.width).reshape(-1, 1)
        action_mask = torch.ones(1, 1, env.action_height, env.action_width)
        action_mask = env.action_padding(action_mask)
        action_mask = torch.ones_like(action_mask) - action_mask
        self.mass_weight_h = self.mass_weight_h * action_mask
        self.mass_weight_w = self.mass_weight_w * action_mask
        self.live_cells = None

    def reset(self):
        return self.env.reset()

    @torch.no_grad()
    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        live_cells = torch.sum(self.inner_env.universe, dim=[1, 2, 3])
        center_of_mass_h = torch.sum(obs * self.mass_weight_h, dim=[1, 2, 3]) / (live_cells + 1e-7)
        center_of_mass_w = torch.sum(obs * self.mass_weight_w, dim=[1, 2, 3]) / (live_cells + 1e-7)
        center_of_mass = torch.cat([center_of_mass_h.unsqueeze(0), center_of_mass_w.unsqueeze(0)])
        if self.center_of_mass is None:
            self.center_of_mass = center_of_mass
        else:
            velocity = self.center_of_mass - center_of_mass
            speed = torch.sqrt(torch.sum(torch.pow(velocity, 2)))
            self.speed = speed
            self.velocity = velocity
            self.center_of_mass = center_of_mass
            reward += speed
        self.live_cells = live_cells
        return obs, reward, done, info
