"""Stress the one-launch step variants against the plain one-warp kernel on identical inputs
(GPU box).  Prints the first mismatch per variant with its location."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carle_b200  # noqa: E402


def rollout(env_vars, size, win, n, steps, seed, sums, u8):
    for k in ("CARLE_FUSED_IMPL", "CARLE_STRIP_R", "CARLE_PDL", "CARLE_STRIP128"):
        os.environ.pop(k, None)
    os.environ.update(env_vars)
    g = torch.Generator(device="cuda").manual_seed(seed)
    soup = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.4).float()
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                           action_height=win, obs_mode="packed", fused_reductions=sums)
    env.rules_from_string("B368/S245")
    env.reset()
    env.universe = soup
    states, reds = [], []
    for t in range(steps):
        a = (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1)
        a = a.to(torch.uint8) if u8 else a.float()
        env.step(a)
        states.append(env.packed_universe.clone())
        if sums:
            reds.append(env.last_reductions.clone())
    torch.cuda.synchronize()
    return states, reds


def main():
    variants = {
        "strip2": {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "2"},
        "strip4": {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "4"},
        "strip2-nopdl": {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "2", "CARLE_PDL": "0"},
        "strip4-nopdl": {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "4", "CARLE_PDL": "0"},
        "quad": {"CARLE_FUSED_IMPL": "quad"},
        "tma": {"CARLE_FUSED_IMPL": "tma"},
    }
    for size, win, ns in ((256, 64, (13, 200, 3000)), (128, 32, (29, 1200, 9000))):
        for n in ns:
            for sums in (True, False):
                for u8 in (False, True):
                    want, want_red = rollout({"CARLE_FUSED_IMPL": "direct"}, size, win, n, 12, n, sums, u8)
                    for name, ev in variants.items():
                        if size == 128 and name in ("strip4", "strip4-nopdl", "quad"):
                            continue
                        for rep in range(3):
                            got, got_red = rollout(ev, size, win, n, 12, n, sums, u8)
                            bad = None
                            for t in range(len(want)):
                                if not torch.equal(got[t], want[t]):
                                    d = (got[t] != want[t]).nonzero()
                                    bad = (t, d.shape[0], d[:6].tolist())
                                    break
                                if sums and not torch.equal(got_red[t], want_red[t]):
                                    d = (got_red[t] != want_red[t]).nonzero()
                                    bad = ("red", t, d.shape[0], d[:6].tolist())
                                    break
                            tag = "OK " if bad is None else "BAD"
                            if bad is not None or rep == 0:
                                print(tag, size, n, "sums" if sums else "nosums", "u8" if u8 else "f32",
                                      name, rep, bad, flush=True)


if __name__ == "__main__":
    main()
