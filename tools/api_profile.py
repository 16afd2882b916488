#!/usr/bin/env python
"""Where the host time of the public API goes: cProfile over CARLE.step / SpeedDetector.step loops
(device actions, no synchronisation inside the loop).  python tools/api_profile.py > gpurun_out/api_profile.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import carle_b200


def loop(env, acts, k):
    for i in range(k):
        env.step(acts[i & 3])


def main():
    dev = torch.device("cuda", 0)
    for label, n, size, win, wrap, mode in (("cfg1 1x64x64 packed", 1, 64, 32, False, "packed"),
                                            ("cfg2 4096x128x128 packed", 4096, 128, 32, False, "packed"),
                                            ("cfg2 4096x128x128 float32 obs", 4096, 128, 32, False, "float32"),
                                            ("cfg3 16384x256x256 SpeedDetector packed", 16384, 256, 64, True, "packed")):
        env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                               device="cuda", obs_mode=mode)
        if wrap:
            env = carle_b200.SpeedDetector(env)
        env.reset()
        acts = [(torch.rand(n, 1, win, win, device=dev) <= 0.1).float() for _ in range(4)]
        loop(env, acts, 50)
        torch.cuda.synchronize()
        k = 2000 if n <= 4096 else 300
        t0 = time.perf_counter()
        loop(env, acts, k)
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f"=== {label}: host {1e6 * t_host / k:.2f} us/step to enqueue, {1e6 * t_all / k:.2f} us/step with the final sync")
        prof = cProfile.Profile()
        prof.enable()
        loop(env, acts, k)
        prof.disable()
        torch.cuda.synchronize()
        s = io.StringIO()
        pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(14)
        print("\n".join(s.getvalue().splitlines()[4:26]))
        del env, acts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
