#!/usr/bin/env python
"""Test infrastructure (it imports oracle/): cost of oracle/torch_port.py (what bench.py times as the CPU arm) next to the REAL reference on the
same host cores -- run once on the GPU box with the reference pushed as scratch (CARLE_REFERENCE_PATH,
tools/gpu_reference_visit.sh).  Prints one JSON object; the outputs of the two are also compared."""
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ["CARLE_REFERENCE_PATH"]


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.path.insert(0, REF)
    import carle.env as ref_env
    import carle.mcl as ref_mcl
    orig = ref_env.CARLE.set_neighborhood

    def patched(self):
        with torch.no_grad():
            orig(self)
    ref_env.CARLE.set_neighborhood = patched
    return ref_env, ref_mcl


def timed(env, actions, seconds):
    for a in actions[:3]:
        env.step(a)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.step(actions[done % len(actions)])
        done += 1
    return done, time.perf_counter() - t0


def main():
    from oracle.torch_port import TorchPortCARLE, TorchPortSpeedDetector
    from oracle import carle_oracle as oc
    ref_env, ref_mcl = import_reference()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    out = {"threads": threads, "torch": torch.__version__, "cases": []}
    for label, n, size, win, rule, wrapper, seconds in (
            ("configs[2] sample: Morley + SpeedDetector, 256 x 256x256", 256, 256, 64, "B368/S245", True, 12.0),
            ("configs[1] sample: Life, 256 x 128x128", 256, 128, 32, "B3/S23", False, 8.0),
            ("configs[0]: Life, 1 x 64x64", 1, 64, 32, "B3/S23", False, 5.0)):
        torch.manual_seed(3)
        soup = (torch.rand(n, 1, size, size) < 0.5).float()
        actions = [1.0 * (torch.rand(n, 1, win, win) <= 0.1) for _ in range(4)]
        ref = ref_env.CARLE(instances=n, height=size, width=size, action_height=win, action_width=win, device="cpu")
        ref.rules_from_string(rule)
        port = TorchPortCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
        port.birth, port.survive = oc.rules_from_string(rule)
        ref_w = ref_mcl.SpeedDetector(ref) if wrapper else ref
        port_w = TorchPortSpeedDetector(port) if wrapper else port
        ref_w.reset()
        port_w.reset()
        ref.universe = soup.clone()
        port.universe = soup.clone()
        # same outputs first (4 steps)
        same = True
        for a in actions:
            o1, r1 = ref_w.step(a)[:2]
            o2, r2 = port_w.step(a)[:2]
            same &= bool(torch.equal(o1, o2)) and bool(torch.allclose(r1, r2, rtol=1e-6, atol=1e-6))
        k_ref, t_ref = timed(ref_w, actions, seconds)
        k_port, t_port = timed(port_w, actions, seconds)
        cells = n * size * size
        out["cases"].append({
            "case": label, "outputs_equal": same,
            "reference_cell_updates_per_sec": k_ref * cells / t_ref,
            "port_cell_updates_per_sec": k_port * cells / t_port,
            "port_over_reference": (k_port / t_port) / (k_ref / t_ref),
            "reference_ms_per_step": 1e3 * t_ref / k_ref, "port_ms_per_step": 1e3 * t_port / k_port})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
