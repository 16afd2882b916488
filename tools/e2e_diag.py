"""Where an end-to-end step of the headline shape spends its time on the host side: per-call duration of
SpeedDetector(CARLE).step(pinned float32 action) with a synchronisation per step (strict) and without
(pipelined, rollout.host_rollout), for the three ways of shipping the packed words
(CARLE(host_pack_copy=True / False / "auto")).   python tools/e2e_diag.py > gpurun_out/e2e_diag.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import carle_b200  # noqa: E402

N, SIZE, WIN, STEPS = 16384, 256, 64, 20
dev = torch.device("cuda:0")
feed = [(torch.rand(N, 1, WIN, WIN) <= 0.1).to(torch.float32).pin_memory() for _ in range(2)]
out = {}
for mode in (True, False, "auto"):
    env = carle_b200.SpeedDetector(carle_b200.CARLE(instances=N, height=SIZE, width=SIZE, action_width=WIN,
                                                     action_height=WIN, device="cuda:0", obs_mode="packed",
                                                     host_pack_copy=mode))
    env.rules_from_string("B368/S245")
    env.reset()
    env.inner_env.packed_universe.random_(-2**31, 2**31 - 1)
    inner = env.inner_env
    calls = []
    orig = inner._pack_on_host

    def timed_pack(a, orig=orig, calls=calls):
        t0 = time.perf_counter()
        r = orig(a)
        calls.append(time.perf_counter() - t0)
        return r
    inner.__dict__["_pack_on_host"] = timed_pack
    row = {}
    for name in ("strict", "pipelined", "strict_again"):
        for rep in range(2):                       # the first pass warms up
            del calls[:]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name.startswith("strict"):
                for i in range(STEPS):
                    env.step(feed[i & 1])[1].cpu()
            else:
                carle_b200.host_rollout(env, [feed[i & 1] for i in range(STEPS)])
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        c = sorted(calls)
        row[name] = {"ms_per_step": round(1e3 * wall / STEPS, 4), "pack_call_ms_median": round(1e3 * c[len(c) // 2], 4),
                     "pack_call_ms_max": round(1e3 * c[-1], 4), "pack_calls": len(c)}
    out[str(mode)] = row
    print(mode, json.dumps(row), file=sys.stderr)
    del env, inner
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
