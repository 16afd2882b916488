#!/usr/bin/env python
"""Where does the host->device bandwidth of the staged (pipelined) action path go?
Times 256 MiB pinned->device copies on the current stream, on a side stream with events, and the
full host_rollout, with and without the step kernel running beside them."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import carle_b200


def timed(fn, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(reps)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    n, size, win = 16384, 256, 64
    host = [(torch.rand(n, 1, win, win) <= 0.1).to(torch.float32).pin_memory() for _ in range(2)]
    nbytes = host[0].numel() * 4
    dst = [torch.empty(n, 1, win, win, device=dev) for _ in range(2)]
    side = torch.cuda.Stream(device=dev)
    out = {}

    def plain(reps):
        for i in range(reps):
            dst[i % 2].copy_(host[i % 2], non_blocking=True)
    out["copy_current_stream_ms"] = timed(plain)

    def on_side(reps):
        for i in range(reps):
            with torch.cuda.stream(side):
                dst[i % 2].copy_(host[i % 2], non_blocking=True)
    out["copy_side_stream_ms"] = timed(on_side)

    def fresh_dst(reps):
        for i in range(reps):
            host[i % 2].to(dev, non_blocking=True)
    out["copy_to_fresh_tensor_ms"] = timed(fresh_dst)

    env = carle_b200.SpeedDetector(carle_b200.CARLE(instances=n, height=size, width=size,
                                                    action_width=win, action_height=win,
                                                    device="cuda:0", obs_mode="packed"))
    env.rules_from_string("B368/S245")
    env.reset()
    env.inner_env.packed_universe.random_(-2**31, 2**31 - 1)

    def strict(reps):
        for i in range(reps):
            env.step(host[i % 2])[1].cpu()
    out["strict_step_ms"] = timed(strict)

    def staged_no_readback(reps):
        inner = env.inner_env
        nxt = inner.stage_action(host[0])
        for i in range(reps):
            cur = nxt
            nxt = inner.stage_action(host[(i + 1) % 2])
            env.step(cur)
    out["staged_no_readback_ms"] = timed(staged_no_readback)
    out["staged_no_readback_again_ms"] = timed(staged_no_readback)

    def rollout(reps):
        carle_b200.host_rollout(env, [host[i % 2] for i in range(reps)])
    out["host_rollout_ms"] = timed(rollout)
    out["host_rollout_again_ms"] = timed(rollout)
    pinned_rewards = torch.empty((20, n, 1)).pin_memory()

    def rollout_prealloc(reps):
        carle_b200.host_rollout(env, [host[i % 2] for i in range(reps)], rewards_out=pinned_rewards)
    out["host_rollout_prealloc_ms"] = timed(rollout_prealloc)
    out["gbs_at_5ms"] = nbytes / 5e-3 / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
