#!/usr/bin/env python
"""Quick A/B of the headline kernels for the library selected by CARLE_B200_LIB (and whatever
CARLE_* switches are set): us per launch of K-step CUDA graphs, same measurement as bench.py.
    python tools/ab_headline.py [label]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("CARLE_B200_LIB", "default")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {"label": label}
    for name, n, size, win, rule, sums, tail, k, reps in (
            ("cfg3_morley_sums", 16384, 256, 64, "B368/S245", True, False, 20, 40),
            ("cfg3_morley_sums_tail", 16384, 256, 64, "B368/S245", True, True, 20, 40),
            ("cfg3_life", 16384, 256, 64, "B3/S23", False, False, 20, 40),
            ("cfg2_life", 4096, 128, 32, "B3/S23", False, False, 200, 40),
            ("cfg2_life_sums_tail", 4096, 128, 32, "B3/S23", True, True, 200, 40),
            ("cfg4shard_life", 131072, 64, 32, "B3/S23", False, False, 20, 20)):
        try:
            wl = bench.StepWorkload(torch, dev, n, size, win, rule, sums, 512, speed_tail=tail)
            out[name] = round(1e3 * bench.graph_rate(torch, wl, dev, False, k, repeats=reps, tail=tail), 3)
            del wl
        except Exception as exc:
            out[name] = repr(exc)[:120]
        torch.cuda.empty_cache()
    # strict float32-observation mode through the public API (GPU-bound: 4 B per cell written)
    import carle_b200
    for name, n, size, win in (("f32obs_cfg2", 4096, 128, 32), ("f32obs_cfg3_4096", 4096, 256, 64)):
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
        for i in range(100):
            env.step(acts[i & 3])
        b.record()
        torch.cuda.synchronize()
        out[name] = round(1e3 * a.elapsed_time(b) / 100, 3)
        del env
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
