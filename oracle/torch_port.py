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
