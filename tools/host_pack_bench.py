"""Host-side action packing (carle_pack_action_host) on this box's cores: GB/s of float32 / uint8 input
per thread count, for the flat multi-stream walk (default), its tuning knobs and the entry-by-entry walk
it replaced.  Needs no GPU (pinned memory is used when there is one, as the e2e path does).

    python tools/host_pack_bench.py [--instances 16384] [--window 64] > gpurun_out/host_pack.json
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carle_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=16384)
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--repeats", type=int, default=7)
    args = ap.parse_args()
    lib = _lib.load()
    n, w = args.instances, args.window
    awpr = w // 32
    pin = torch.cuda.is_available()
    feeds = {}
    for name, dt in (("float32", torch.float32), ("uint8", torch.uint8)):
        ts = [(torch.rand(n, 1, w, w) <= 0.1).to(dt) for _ in range(2)]      # two buffers: > the last-level cache
        feeds[name] = [t.pin_memory() for t in ts] if pin else ts
    out = torch.empty((n, w, awpr), dtype=torch.int32)
    out = out.pin_memory() if pin else out
    flags = (ctypes.c_int32 * 3)()
    cores = len(os.sched_getaffinity(0))

    def rate(name, threads, env):
        for k in ("CARLE_HOST_PACK_FLAT", "CARLE_HOST_PACK_STREAMS", "CARLE_HOST_PACK_PREFETCH", "CARLE_HOST_PACK_ISA"):
            os.environ.pop(k, None)
        os.environ.update(env)
        feed = feeds[name]
        code = _lib.U8 if name == "uint8" else _lib.F32
        times = []
        for i in range(args.repeats + 2):
            t = feed[i & 1]
            t0 = time.perf_counter()
            rc = lib.carle_pack_action_host(w, w, awpr, 0, t.data_ptr(), code, n, out.data_ptr(), flags, threads)
            times.append(time.perf_counter() - t0)
            assert rc == 0
        times = sorted(times[2:])
        return feed[0].numel() * feed[0].element_size() / times[len(times) // 2] / 1e9

    rows = []
    sweep = sorted({t for t in (1, 2, 4, 8, 12, 16, 24, 32) if t <= cores} | {cores})
    variants = [("flat (default: 4 streams, prefetch 4096 B)", {}),
                ("entry by entry (round-2g walk)", {"CARLE_HOST_PACK_FLAT": "0"})]
    for label, env in variants:
        for name in ("float32", "uint8"):
            rows.append({"variant": label, "dtype": name,
                         "gbs_by_threads": {str(t): round(rate(name, t, env), 1) for t in sweep}})
            print(json.dumps(rows[-1]), file=sys.stderr)
    top = [t for t in (8, cores) if t <= cores]
    for label, env in (("flat, 1 stream", {"CARLE_HOST_PACK_STREAMS": "1"}),
                       ("flat, 2 streams", {"CARLE_HOST_PACK_STREAMS": "2"}),
                       ("flat, 8 streams", {"CARLE_HOST_PACK_STREAMS": "8"}),
                       ("flat, no prefetch", {"CARLE_HOST_PACK_PREFETCH": "0"}),
                       ("flat, prefetch 2048 B", {"CARLE_HOST_PACK_PREFETCH": "2048"}),
                       ("flat, prefetch 8192 B", {"CARLE_HOST_PACK_PREFETCH": "8192"}),
                       ("flat, AVX2 only", {"CARLE_HOST_PACK_ISA": "avx2"})):
        rows.append({"variant": label, "dtype": "float32",
                     "gbs_by_threads": {str(t): round(rate("float32", t, env), 1) for t in top}})
        print(json.dumps(rows[-1]), file=sys.stderr)
    # what a plain read of the same bytes reaches (torch.sum over the float32 buffer, all cores)
    torch.set_num_threads(cores)
    best = 1e9
    for i in range(5):
        t0 = time.perf_counter()
        feeds["float32"][i & 1].sum()
        best = min(best, time.perf_counter() - t0)
    cpu = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    cpu = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    print(json.dumps({"cores": cores, "cpu": cpu, "pinned": pin, "instances": n, "window": w,
                      "float32_bytes": feeds["float32"][0].numel() * 4,
                      "torch_sum_all_cores_gbs": round(feeds["float32"][0].numel() * 4 / best / 1e9, 1),
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
