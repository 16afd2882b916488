#!/usr/bin/env python
"""Benchmark of the CARLE environment-step hot path (BASELINE.json metric: cell-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one environment step over the whole batch: pack the step's float32 actions
(flags for the master reset), XOR them into the action window, advance one generation.
Workload at N=1 = BASELINE.json configs[1]: B3/S23, 4096 instances of 128x128, 32x32
window, fresh Bernoulli(0.1) actions every step.  N>1 (torchrun, one rank per GPU):
every rank runs that batch on its own GPU — independent instances, no collective on the
data path ("scaling": "weak").

value      whole-job cell-updates/s, inputs (float32 actions, the reference's action
           format) already in HBM, K steps replayed as one CUDA graph, timed with CUDA
           events, max over ranks, median of `repeats` back-to-back K-step regions.
e2e        same metric through the public API (carle_b200.CARLE.step) with the actions
           in pinned HOST memory: H2D copy of each step's action and D2H read of its
           reward inside the timed region.
roofline   the dominant kernel (longest per-launch) timed alone: algorithmic bytes per
           launch / average launch duration vs. MEASURED_PEAKS.json hbm_gbs.
cpu_baseline  oracle/torch_port.py (torch-CPU port of the reference's op sequence) timed
           on this box's host cores on a bounded sample of the same workload.

--impl reference times that CPU port (the reference itself is pure Python/torch and is
not present on the GPU box) on the same config and prints the same JSON line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(name="cfg2", instances=4096, size=128, window=32, rule="B3/S23")
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=CFG2["instances"])
    ap.add_argument("--size", type=int, default=CFG2["size"])
    ap.add_argument("--window", type=int, default=CFG2["window"])
    ap.add_argument("--rule", default=CFG2["rule"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fused-reductions", action="store_true",
                    help="also produce the SpeedDetector sums in the step kernel")
    ap.add_argument("--repeats", type=int, default=0,
                    help="timed K-step regions (0 = auto: ~1.5 s of GPU time)")
    ap.add_argument("--pool-mib", type=int, default=384,
                    help="size of the rotating action pool (must exceed the 126 MB L2)")
    return ap.parse_args()


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------
# CPU arm: the torch port of the reference's step
# ----------------------------------------------------------------------------------------
def cpu_rollout(instances, size, window, rule, steps, warmup, seconds=None, threads=None):
    """Time `steps` env steps of the torch-CPU port (or as many as fit in `seconds`)."""
    import torch
    from oracle.torch_port import TorchPortCARLE
    from oracle import carle_oracle as oc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    env = TorchPortCARLE(width=size, height=size, action_width=window, action_height=window,
                         instances=instances)
    env.birth, env.survive = oc.rules_from_string(rule)
    env.reset()
    env.universe = (torch.rand(instances, 1, size, size) < 0.5).float()
    pool = [1.0 * (torch.rand(instances, 1, window, window) <= 0.1) for _ in range(8)]
    for i in range(warmup):
        env.step(pool[i % len(pool)])
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        env.step(pool[done % len(pool)])
        done += 1
        if seconds is not None and time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return dict(steps=done, seconds=dt, cells=done * instances * size * size,
                threads=threads)


def run_reference(args):
    """--impl reference: the reference's CPU path (torch port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_instances = min(args.instances, 256)
    r = cpu_rollout(sample_instances, args.size, args.window, args.rule, args.steps,
                    max(args.warmup, 3))
    value = r["cells"] / r["seconds"]
    sample = (f"{r['steps']} steps x {sample_instances} instances of {args.size}x{args.size} "
              f"(the {args.instances}-instance batch is sampled: the CPU path's rate is "
              f"independent of N beyond ~64 instances)")
    line = {
        "impl": "reference", "metric": "cell_updates_per_sec", "value": value,
        "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, extra={"sampled_instances": sample_instances}),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": r["threads"],
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, extra=None):
    cfg = {
        "workload": (f"BASELINE configs[1]: {args.rule}, {args.instances} instances "
                     f"{args.size}x{args.size}, {args.window}x{args.window} action window, "
                     f"fresh Bernoulli(0.1) actions every step, Bernoulli(0.5) initial soup"),
        "instances_per_gpu": args.instances, "grid": [args.size, args.size],
        "window": [args.window, args.window], "rule": args.rule,
        "action_format": "float32 [N,1,aw,ah] (reference format)",
        "obs_format": "bit-packed int32 [N,H,W/32] (obs_mode='packed')",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


class GpuWorkload:
    """One rank's batch, driven through the C ABI with pre-allocated buffers."""

    def __init__(self, args, device, obs_mode="packed", fused_reductions=False, rule=None,
                 instances=None, size=None, window=None, pool_mib=None):
        import torch
        import carle_b200
        from carle_b200 import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.load()
        self.n = instances or args.instances
        self.size = size or args.size
        self.win = window or args.window
        self.device = device
        env = carle_b200.CARLE(instances=self.n, height=self.size, width=self.size,
                               action_width=self.win, action_height=self.win,
                               device=str(device), obs_mode=obs_mode,
                               fused_reductions=fused_reductions)
        env.rules_from_string(rule or args.rule)
        env.reset()
        g = torch.Generator(device=device).manual_seed(1 + device.index)
        env.universe = (torch.rand(self.n, 1, self.size, self.size, device=device,
                                   generator=g) < 0.5).float()
        self.env = env
        # rotating pool of distinct action batches, larger than L2 so every step's
        # actions come from HBM
        bytes_per = self.n * self.win * self.win * 4
        mib = pool_mib or args.pool_mib
        self.pool_len = max(2, -(-mib * 2**20 // bytes_per))
        self.pool = [1.0 * (torch.rand(self.n, 1, self.win, self.win, device=device,
                                       generator=g) <= 0.1) for _ in range(self.pool_len)]
        self.action_bytes = bytes_per
        self.cells_per_step = self.n * self.size * self.size
        self.kernels_per_step = 1           # one fused launch (step_stream / step_strip kernel) per step
        env._sync_rule()

    # raw ABI step: what CARLE.step does minus the python-side allocations
    def abi_step(self, i):
        env, lib, _lib = self.env, self.lib, self._lib
        act = self.pool[i % self.pool_len]
        red = env._red_buf.data_ptr() if env.fused_reductions else None
        rc = lib.carle_step_action(env._handle, env._packed.data_ptr(), env._spare.data_ptr(),
                                   act.data_ptr(), _lib.F32, self.n, env._counters.data_ptr(),
                                   red, env._stream())
        if rc:
            raise RuntimeError("C ABI call failed: " + _lib.last_error())
        env._packed, env._spare = env._spare, env._packed

    def capture(self, steps, start=0):
        torch = self.torch
        graph = torch.cuda.CUDAGraph()
        if steps % 2:
            raise ValueError("graph capture needs an even step count (ping-pong buffers)")
        with torch.cuda.graph(graph):
            for i in range(steps):
                self.abi_step(start + i)
        return graph


def time_graph(torch, graph, device, dist_on, repeats):
    """Median (and list) of `repeats` timed replays; each bracketed by barrier + sync."""
    times = []
    for _ in range(repeats):
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize(device)
        if dist_on:
            torch.distributed.barrier()
        ms = a.elapsed_time(b)
        if dist_on:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
    return statistics.median(times), times


def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: carle_b200 has no CPU path")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    import __graft_entry__
    if local_rank == 0:
        __graft_entry__.build()
    if dist_on:
        torch.distributed.barrier()

    steps, warmup = args.steps, max(args.warmup, 3)
    gsteps = steps + (steps % 2)                     # graphs need an even count (ping-pong)
    wl = GpuWorkload(args, device, fused_reductions=args.fused_reductions)
    for i in range(warmup):
        wl.abi_step(i)
    if warmup % 2:
        wl.abi_step(warmup)
    torch.cuda.synchronize(device)
    graph = wl.capture(gsteps, start=warmup + 1)
    graph.replay()                                   # untimed: graph upload / first-run cost
    torch.cuda.synchronize(device)

    # repeat the K-step region until ~1.5 s of GPU time so nvidia-smi sees the load
    probe_ms, _ = time_graph(torch, graph, device, dist_on, 3)
    repeats = args.repeats or int(min(2000, max(5, 1500.0 / max(probe_ms, 1e-3))))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_region, all_ms = time_graph(torch, graph, device, dist_on, repeats)
    ms_per_step = ms_region / gsteps
    total_cells_per_step = wl.cells_per_step * world
    value = total_cells_per_step / (ms_per_step * 1e-3)

    # ---- roofline: the step is ONE kernel (fused action ingestion + generation); its
    # average launch duration is the graph's time per step, measured above --------------
    peak, peak_src = measured_hbm_peak()
    clocks = sampler.stop() if sampler else None
    env = wl.env
    words_state = wl.n * wl.size * ((wl.size + 31) // 32)
    step_bytes = 2 * 4 * words_state + wl.action_bytes      # state read + write, f32 action read
    step_us = 1e3 * ms_per_step
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
    # `ncu --set full` capture of this exact command (profiles/r1f_step_stream_kernel_cfg2.summary.txt, same figure in r1d):
    # the 16.8 MB of actions plus the part of the 8 MiB state the cold-cache replay re-reads;
    # the freshly written state stays in L2 (0 B written back within the launch).
    default_cfg = (args.instances, args.size, args.window, args.rule) == (4096, 128, 32, "B3/S23")
    traffic = 25186560 if default_cfg else None
    roofline = {
        "bound": "hbm", "kernel": "step_stream_kernel (one launch per env step: TMA-staged "
                                  "state + float32-action ingestion + generation; launches "
                                  "chained with programmatic dependent launch)",
        "achieved": step_bytes / step_us / 1e3, "peak": peak, "unit": "GB/s",
        "frac": step_bytes / step_us / 1e3 / peak, "traffic": traffic,
        "peak_source": peak_src, "us_per_launch": step_us,
        "algorithmic_bytes_per_launch": step_bytes,
        "launches_timed": gsteps * repeats,
        "note": ("one launch per env step; duration = CUDA-event time of the K-launch graph / K, "
                 "so it includes the launch gap between consecutive kernels.  The packed state "
                 "(8 MiB per GPU at configs[1]) is L2-resident between steps by the nature of "
                 "the workload; the float32 actions rotate through a pool larger than L2 and "
                 "come from HBM every step (half of the algorithmic bytes)"),
    }

    # ---- e2e: public API, actions in pinned host memory ----------------------------------
    e2e = None if args.no_e2e else run_e2e(args, torch, device, dist_on, world, steps, warmup)

    extras = None
    if not args.no_extras and not dist_on:
        extras = run_extras(args, torch, device)

    cpu_baseline = None
    if rank == 0 and not dist_on and not args.no_cpu_baseline:
        sample_instances = min(args.instances, 256)
        r = cpu_rollout(sample_instances, args.size, args.window, args.rule, 10**9, 3,
                        seconds=args.cpu_seconds)
        cpu_baseline = {
            "value": r["cells"] / r["seconds"], "unit": "cell-updates/s",
            "cores": r["threads"], "kind": "port",
            "sample": (f"{r['steps']} steps x {sample_instances} instances of "
                       f"{args.size}x{args.size} in {r['seconds']:.1f} s, torch "
                       f"{torch.__version__} CPU, {r['threads']} threads "
                       f"(oracle/torch_port.py)")}

    if rank == 0:
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": gsteps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": workload_config(args, extra={
                "l2": (f"actions rotate through a {wl.pool_len}-batch pool "
                       f"({wl.pool_len * wl.action_bytes / 2**20:.0f} MiB > 126 MB L2); "
                       "packed state is the rollout's own 8 MiB working set"),
                "timing": f"median of {repeats} CUDA-graph replays of {gsteps} steps",
                "env_steps_per_sec": 1e3 / ms_per_step * world,
                "instance_steps_per_sec": 1e3 / ms_per_step * wl.n * world}),
            "gcups": value / 1e9,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": wl.kernels_per_step * gsteps,
            "clocks": clocks,
        }
        if extras:
            line["extras"] = extras
        print(json.dumps(line))
    if dist_on:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def run_e2e(args, torch, device, dist_on, world, steps, warmup):
    """K calls of the public CARLE.step with pinned-host float32 actions; each step's
    reward is read back to the host (as carle/train_mcl.py:68-69 does)."""
    import carle_b200
    n, size, win = args.instances, args.size, args.window
    env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                           action_height=win, device=str(device), obs_mode="packed")
    env.rules_from_string(args.rule)
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device=device) < 0.5).float()
    host_pool = [(1.0 * (torch.rand(n, 1, win, win) <= 0.1)).pin_memory() for _ in range(8)]
    for i in range(warmup):
        env.step(host_pool[i % 8])[1].cpu()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    d2h = 0
    for i in range(steps):
        obs, reward, done, info = env.step(host_pool[i % 8])
        r = reward.cpu()                              # D2H + sync, every step
        d2h = r.numel() * r.element_size()
    b.record()
    torch.cuda.synchronize(device)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    if dist_on:
        torch.distributed.barrier()
    ms = max(a.elapsed_time(b), 0.0)
    if dist_on:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    cells = n * size * size * steps * world
    return {"value": cells / (ms * 1e-3), "unit": "cell-updates/s",
            "h2d_bytes_per_step": n * win * win * 4, "d2h_bytes_per_step": d2h,
            "ms_per_step": ms / steps, "wall_ms_per_step": wall_ms / steps,
            "h2d_gbs_lower_bound": n * win * win * 4 / (ms / steps * 1e-3) / 1e9,
            "api": "carle_b200.CARLE.step(pinned host float32 action) + reward.cpu()"}


def run_extras(args, torch, device):
    """Secondary measurements (not the headline): strict float32-obs mode, fused K-step
    rollout, and the BASELINE configs[2] shape with the fused SpeedDetector sums."""
    out = {}
    try:
        # (a) strict drop-in mode: float32 obs materialised every step
        import carle_b200
        n, size, win = args.instances, args.size, args.window
        env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                               action_height=win, device=str(device), obs_mode="float32")
        env.reset()
        env.universe = (torch.rand(n, 1, size, size, device=device) < 0.5).float()
        pool = [1.0 * (torch.rand(n, 1, win, win, device=device) <= 0.1) for _ in range(24)]
        for i in range(5):
            env.step(pool[i])
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 100
        a.record()
        for i in range(k):
            env.step(pool[i % 24])
        b.record()
        torch.cuda.synchronize(device)
        out["float32_obs_api"] = {"cell_updates_per_sec": n * size * size * k / (a.elapsed_time(b) * 1e-3),
                                  "note": "CARLE.step, device float32 actions, float32 obs each step"}
        # (b) fused rollout: one launch for K generations with K action slabs
        env2 = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                                action_height=win, device=str(device), obs_mode="packed")
        env2.reset()
        env2.universe = (torch.rand(n, 1, size, size, device=device) < 0.5).float()
        k = 16
        acts = torch.stack(pool[:k])
        env2.step_many(acts)
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            env2.step_many(acts)
        b.record()
        torch.cuda.synchronize(device)
        out["fused_rollout_k16"] = {"cell_updates_per_sec": n * size * size * k * reps / (a.elapsed_time(b) * 1e-3),
                                    "note": "CARLE.step_many: 16 generations per launch, 16 float32 action slabs packed by one launch"}
        env2.step_many(64)
        torch.cuda.synchronize(device)
        a.record()
        for _ in range(reps):
            env2.step_many(64)
        b.record()
        torch.cuda.synchronize(device)
        out["free_run_k64"] = {"cell_updates_per_sec": n * size * size * 64 * reps / (a.elapsed_time(b) * 1e-3),
                               "note": "zero-action free run, 64 generations per launch (register-resident)"}
        # (b2) the random agent fused into the step kernel: Bernoulli(0.1) toggles drawn with
        #      Philox inside carle_step_random -- one launch per env step, no action tensor
        from carle_b200 import _lib as _l
        lib = _l.load()
        env3 = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                                action_height=win, device=str(device), obs_mode="packed")
        env3.reset()
        env3.universe = (torch.rand(n, 1, size, size, device=device) < 0.5).float()
        env3._sync_rule()
        words = env3._action_buf

        def agent_steps(k0, k):
            for i in range(k):
                rc = lib.carle_step_random(env3._handle, env3._packed.data_ptr(),
                                           env3._spare.data_ptr(), 7, k0 + i, 0.1, n,
                                           words.data_ptr(), env3._counters.data_ptr(), None,
                                           env3._stream())
                assert rc == 0, _l.last_error()
                env3._packed, env3._spare = env3._spare, env3._packed
        agent_steps(0, 4)
        torch.cuda.synchronize(device)
        gk = 100
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            agent_steps(4, gk)
        gr.replay()
        torch.cuda.synchronize(device)
        ms, _ = time_graph(torch, gr, device, False, 20)
        out["device_random_agent_fused"] = {
            "cell_updates_per_sec": n * size * size * gk / (ms * 1e-3), "us_per_step": ms * 1e3 / gk,
            "note": "random-agent rollout entirely on the GPU: Bernoulli(0.1) toggles drawn with "
                    "Philox inside the step kernel (carle_step_random), one launch per env step"}
        # (b3) end-to-end with uint8 host actions (the API accepts them; 4x fewer PCIe bytes)
        env4 = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                                action_height=win, device=str(device), obs_mode="packed")
        env4.reset()
        env4.universe = (torch.rand(n, 1, size, size, device=device) < 0.5).float()
        host8 = [(torch.rand(n, 1, win, win) <= 0.1).to(torch.uint8).pin_memory() for _ in range(8)]
        for i in range(5):
            env4.step(host8[i])[1].cpu()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 100
        a.record()
        for i in range(k):
            env4.step(host8[i % 8])[1].cpu()
        b.record()
        torch.cuda.synchronize(device)
        out["e2e_uint8_host_actions"] = {
            "cell_updates_per_sec": n * size * size * k / (a.elapsed_time(b) * 1e-3),
            "us_per_step": a.elapsed_time(b) * 1e3 / k,
            "note": "CARLE.step(pinned host uint8 action) + reward.cpu(): 4 MiB H2D per step"}
        del env, env2, env3, env4, host8, pool, acts, words, gr
        # (c) the other BASELINE shapes on one GPU, same measurement as the headline (K steps
        #     in one CUDA graph, float32 actions from a pool larger than L2), each with its
        #     algorithmic bytes per step against the measured HBM peak
        peak, _ = measured_hbm_peak()

        def shape(label, instances, size, window, rule, sums, note):
            wl = GpuWorkload(args, device, fused_reductions=sums, rule=rule,
                             instances=instances, size=size, window=window, pool_mib=512)
            for i in range(4):
                wl.abi_step(i)
            torch.cuda.synchronize(device)
            g = wl.capture(20, start=4)
            g.replay()
            torch.cuda.synchronize(device)
            ms, _ = time_graph(torch, g, device, False, 7)
            nbytes = 2 * 4 * instances * size * ((size + 31) // 32) + wl.action_bytes
            gbs = nbytes * 20 / (ms * 1e-3) / 1e9
            out[label] = {"cell_updates_per_sec": wl.cells_per_step * 20 / (ms * 1e-3),
                          "ms_per_step": ms / 20, "algorithmic_gbs": gbs,
                          "frac_of_hbm_peak": gbs / peak, "note": note}
            del wl, g
            torch.cuda.empty_cache()

        # (c0) BASELINE configs[0] shape (the reference's own CPU-runnable case): one 64x64
        #      instance; per-call latency of the public API, wall clock, 2000 calls
        env1 = carle_b200.CARLE(instances=1, height=64, width=64, action_width=32, action_height=32,
                                device=str(device))
        env1.reset()
        env1.universe = (torch.rand(1, 1, 64, 64, device=device) < 0.5).float()
        dev_acts = [1.0 * (torch.rand(1, 1, 32, 32, device=device) <= 0.1) for _ in range(16)]
        host_acts = [a.cpu() for a in dev_acts]
        lat = {}
        for label, acts, read in (("device_action", dev_acts, False), ("host_action_reward_read", host_acts, True)):
            for i in range(50):
                env1.step(acts[i % 16])
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            for i in range(2000):
                r = env1.step(acts[i % 16])[1]
                if read:
                    r.cpu()
            torch.cuda.synchronize(device)
            lat[label + "_us_per_call"] = (time.perf_counter() - t0) / 2000 * 1e6
        lat["note"] = ("BASELINE configs[0] shape (1 x 64x64, 32x32 window, strict float32 obs): wall-clock "
                       "latency of carle_b200.CARLE.step; the reference's torch CPU step takes ~850 us "
                       "per call (BASELINE.md section 2)")
        out["cfg1_api_latency"] = lat
        del env1, dev_acts, host_acts
        shape("cfg3_morley_speed", 16384, 256, 64, "B368/S245", True,
              "BASELINE configs[2]: B368/S245, 16384 x 256x256, 64x64 window, fused live/Sh/Sw "
              "sums, float32 actions (step_strip_kernel: four independent 64-row strips per instance)")
        # (c1) the same config through the PUBLIC wrapper API: SpeedDetector(CARLE).step with device
        #      float32 actions, reward tensor produced every step (fused sums + one tail launch)
        env5 = carle_b200.SpeedDetector(carle_b200.CARLE(
            instances=16384, height=256, width=256, action_width=64, action_height=64,
            device=str(device), obs_mode="packed"))
        env5.rules_from_string("B368/S245")
        env5.reset()
        env5.inner_env.packed_universe.random_(-2**31, 2**31 - 1)
        acts5 = [1.0 * (torch.rand(16384, 1, 64, 64, device=device) <= 0.1) for _ in range(3)]
        for i in range(3):
            env5.step(acts5[i])
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 30
        a.record()
        for i in range(k):
            reward5 = env5.step(acts5[i % 3])[1]
        b.record()
        torch.cuda.synchronize(device)
        out["cfg3_speeddetector_api"] = {
            "cell_updates_per_sec": 16384 * 256 * 256 * k / (a.elapsed_time(b) * 1e-3),
            "ms_per_step": a.elapsed_time(b) / k,
            "note": "carle_b200.SpeedDetector(CARLE(obs_mode='packed')).step(device float32 action): "
                    "step kernel with fused sums + carle_speed_tail, reward [N,1] every step"}
        del env5, acts5, reward5
        torch.cuda.empty_cache()
        shape("cfg3_shape_life_no_sums", 16384, 256, 64, "B3/S23", False,
              "same shape, B3/S23, no reward sums")
        shape("cfg4_shard_131072x64x64", 131072, 64, 32, "B3/S23", False,
              "BASELINE configs[3] per-GPU shard at 8 GPUs: 131072 instances of 64x64, 32x32 window")
        # (d) configs[4] on ONE GPU: a single 65536 x 65536 Life torus, tiled family with
        #     16-generation temporal blocks (the 8-GPU row-band version: tools/bigrid_check.py)
        big = carle_b200.CARLE(instances=1, height=65536, width=65536, device=str(device),
                               obs_mode="packed")
        big.reset()
        big.packed_universe.random_(-2**31, 2**31 - 1)
        big.step_many(16)
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        big.step_many(64)
        b.record()
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b)
        out["giant_grid_65536_1gpu"] = {
            "cell_updates_per_sec": 65536.0 * 65536.0 * 64 / (ms * 1e-3),
            "us_per_generation": ms * 1e3 / 64,
            "algorithmic_gbs": 65536.0 * 65536.0 * 64 * 0.25 / (ms * 1e-3) / 1e9,
            "note": "single 65536x65536 B3/S23 torus, free run, 16 generations per launch in "
                    "256x256 register tiles (224x224 written), one B200"}
        del big
    except Exception as exc:  # extras never break the headline line
        out["error"] = repr(exc)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
