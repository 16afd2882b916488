#!/usr/bin/env python
"""Strict float32-observation mode through the public API (the default obs_mode): 4096 x 128x128 and
4096 x 256x256, eager steps; run under ncu to capture the step kernel that also writes the observation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import carle_b200

dev = torch.device("cuda", 0)
for n, size, win in ((4096, 128, 32), (4096, 256, 64)):
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                           device="cuda", obs_mode="float32")
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device=dev) < 0.5).float()
    acts = [(torch.rand(n, 1, win, win, device=dev) <= 0.1).float() for _ in range(4)]
    for i in range(10):
        env.step(acts[i & 3])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(50):
        env.step(acts[i & 3])
    b.record()
    torch.cuda.synchronize()
    us = 1e3 * a.elapsed_time(b) / 50
    print(f"f32 obs {n} x {size}x{size}: {us:.1f} us/step, {n * size * size / us / 1e6:.3e} cell-updates/s, "
          f"{n * size * size * 4.25 / us / 1e3:.0f} GB/s of 4.25 B/cell")
    del env
    torch.cuda.empty_cache()
