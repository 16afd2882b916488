#!/usr/bin/env python
"""Benchmark of the CARLE environment-step hot path (BASELINE.json metric: cell-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

HEADLINE WORKLOAD = BASELINE.json configs[2], the largest single-GPU configuration and the
shape the >= 1e12 target is quoted on: Morley B368/S245 with the mcl.py SpeedDetector reward
wrapper, 16384 instances of 256x256, 64x64 action window, fresh Bernoulli(0.1) float32
actions every step.  A "step" is one wrapped environment step over the whole batch: action
ingestion + XOR + one generation + the per-instance SpeedDetector sums + the wrapper's tail
(centre of mass, velocity, batch-wide speed, reward) -- ONE kernel, step_strip_kernel.  With --gpus N (torchrun, one rank per GPU) every rank steps a batch of
that shape -- independent instances, no collective on the data path ("scaling": "weak").

value      whole-job cell-updates/s, float32 actions (the reference's format) already in HBM
           and rotating through a pool larger than L2, K steps replayed as one CUDA graph,
           CUDA events, max over ranks, median of the repeated K-step regions.
e2e        the same metric through the public API, carle_b200.SpeedDetector(CARLE).step, with
           every step's float32 action in pinned HOST memory and the step's reward read back
           to the host (`reward.cpu()`) -- host-side packing, H2D and D2H inside the timed region.
           `e2e_variants` lists the same call with the floats shipped unpacked (host_pack=False),
           the pipelined form and uint8 / pre-packed host actions.
roofline   the dominant kernel (step_strip_kernel) timed alone as a K-launch graph:
           algorithmic bytes per launch / average launch duration vs MEASURED_PEAKS.json.
cpu_baseline  oracle/torch_port.py (torch-CPU port of the reference's op sequence, wrapper
           included) timed on this box's host cores on a bounded sample of the same workload.
extras     per N: BASELINE configs[3] AS STATED (2^20 instances of 64x64 split over the N ranks,
           strong scaling, plus the exact whole-batch-semantics variant with its two 8-byte
           all-reduces) and configs[4] (one 65536^2 torus in N row bands, NVLink halo stores,
           bit-exact self-check at 16384^2 first); at N = 1 also configs[0] / configs[1] shapes,
           the strict float32-observation mode, fused K-step rollouts and the device agent.

--impl reference times the CPU port on the same config (all host threads; every step covers
the whole batch in chunks unless the time budget forces a sample) and prints the same line.
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

HEADLINE = dict(instances=16384, size=256, window=64, rule="B368/S245")
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=HEADLINE["instances"])
    ap.add_argument("--size", type=int, default=HEADLINE["size"])
    ap.add_argument("--window", type=int, default=HEADLINE["window"])
    ap.add_argument("--rule", default=HEADLINE["rule"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=150.0,
                    help="budget of the timed region of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--repeats", type=int, default=0,
                    help="timed K-step regions (0 = auto: ~1.5 s of GPU time)")
    ap.add_argument("--pool-mib", type=int, default=768,
                    help="size of the rotating action pool (must exceed the 126 MB L2)")
    return ap.parse_args()


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload`, from the
    committed `ncu --set full` capture (profiles/r2_kernel_traffic.json, written by
    tools/ncu_traffic.py from the .ncu-rep of this command); None when there is no capture."""
    try:
        with open(TRAFFIC_FILE) as f:
            table = json.load(f)
        row = table["kernels"][kernel][workload]
        return int(row["dram_bytes_read"]) + int(row["dram_bytes_write"]), row.get("source")
    except Exception:
        return None, None


def workload_config(args):
    """Identical in both arms (the driver compares the two lines' configs)."""
    return {
        "workload": (f"BASELINE configs[2]: Morley {args.rule} + mcl.py SpeedDetector reward wrapper, "
                     f"{args.instances} instances {args.size}x{args.size}, {args.window}x{args.window} "
                     f"action window, fresh Bernoulli(0.1) float32 actions every step, "
                     f"Bernoulli(0.5) initial soup"),
        "instances_per_gpu": args.instances, "grid": [args.size, args.size],
        "window": [args.window, args.window], "rule": args.rule, "wrapper": "SpeedDetector",
        "action_format": "float32 [N,1,aw,ah] (reference format)",
    }


# ----------------------------------------------------------------------------------------
# CPU arm: the torch port of the reference's wrapped step
# ----------------------------------------------------------------------------------------
def _cpu_env(instances, size, window, rule, wrapper):
    import torch
    from oracle.torch_port import TorchPortCARLE, TorchPortSpeedDetector
    from oracle import carle_oracle as oc
    env = TorchPortCARLE(width=size, height=size, action_width=window, action_height=window,
                         instances=instances)
    env.birth, env.survive = oc.rules_from_string(rule)
    env.reset()
    env.universe = (torch.rand(instances, 1, size, size) < 0.5).float()
    return TorchPortSpeedDetector(env) if wrapper else env


def cpu_sample(instances, size, window, rule, wrapper, seconds, warmup=3, threads=None):
    """As many steps of a `instances`-sized sample as fit in `seconds` (cpu_baseline)."""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    env = _cpu_env(instances, size, window, rule, wrapper)
    pool = [1.0 * (torch.rand(instances, 1, window, window) <= 0.1) for _ in range(4)]
    for i in range(warmup):
        env.step(pool[i % 4])
    done, t0 = 0, time.perf_counter()
    while True:
        env.step(pool[done % 4])
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return dict(steps=done, seconds=dt, cells=done * instances * size * size, threads=threads)


def run_reference(args):
    """--impl reference: the reference's CPU path (torch port incl. the SpeedDetector wrapper,
    all host threads) on the headline config.  Every timed step walks the batch in chunks of
    512 instances; when the whole batch does not fit the time budget the step covers the first
    chunks only and says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    n, size, win = args.instances, args.size, args.window
    chunk = min(n, 512)
    steps, warmup = args.steps, max(args.warmup, 3)
    # one chunk env, warmed up and timed once, decides how many chunks a step can cover
    first = _cpu_env(chunk, size, win, args.rule, True)
    acts = [1.0 * (torch.rand(chunk, 1, win, win) <= 0.1) for _ in range(2)]
    for i in range(warmup):
        first.step(acts[i % 2])
    t0 = time.perf_counter()
    first.step(acts[0])
    t_chunk = time.perf_counter() - t0
    all_chunks = -(-n // chunk)
    chunks = int(max(1, min(all_chunks, args.ref_seconds / max(steps * t_chunk, 1e-9))))
    envs = [first] + [_cpu_env(chunk, size, win, args.rule, True) for _ in range(chunks - 1)]
    for e in envs[1:]:
        e.step(acts[1])                       # (touch every chunk's state once, untimed)
    t0 = time.perf_counter()
    for s in range(steps):
        for e in envs:
            e.step(acts[s % 2])
    dt = time.perf_counter() - t0
    covered = min(n, chunks * chunk)
    value = steps * covered * size * size / dt
    full = covered == n
    sample = (f"{steps} steps, each over {covered} of the {n} instances in chunks of {chunk} "
              f"({'the whole batch' if full else 'time-bounded sample; ms_per_step is scaled to the whole batch'}), "
              f"torch {torch.__version__} CPU, {threads} threads (oracle/torch_port.py: TorchPortCARLE + "
              f"TorchPortSpeedDetector; bit-pinned to the reference's outputs, 0.95-1.09x its cost: "
              f"profiles/r2_port_vs_reference_cpu.json)")
    line = {
        "impl": "reference", "metric": "cell_updates_per_sec", "value": value,
        "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * dt / steps * (n / covered), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads,
                         "kind": "port", "sample": sample, "whole_batch": full},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


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


def bind_to_gpu_numa_node(torch, index):
    """Pin this rank's host threads to the cores next to its GPU (so the pinned action pools are
    first-touched on that NUMA node and the H2D copies do not cross the socket link).  Best
    effort: returns a description, or None when the topology cannot be read."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


class StepWorkload:
    """One rank's batch, driven through the C ABI with pre-allocated buffers: what
    SpeedDetector(CARLE).step does minus the python."""

    def __init__(self, torch, device, instances, size, window, rule, sums, pool_mib, seed=1,
                 speed_tail=False):
        import carle_b200
        from carle_b200 import _lib
        self.torch, self._lib, self.lib = torch, _lib, _lib.load()
        self.n, self.size, self.win, self.device = instances, size, window, device
        env = carle_b200.CARLE(instances=instances, height=size, width=size, action_width=window,
                               action_height=window, device=str(device), obs_mode="packed",
                               fused_reductions=sums)
        env.rules_from_string(rule)
        env.reset()
        g = torch.Generator(device=device).manual_seed(seed + device.index)
        env.packed_universe.random_(-2**31, 2**31 - 1, generator=g)     # Bernoulli(0.5) soup
        self.env = env
        # rotating pool of distinct action batches, larger than L2: every step's actions come
        # from HBM
        self.action_bytes = instances * window * window * 4
        self.pool_len = max(2, -(-pool_mib * 2**20 // self.action_bytes))
        self.pool = []
        for _ in range(self.pool_len):
            a = torch.rand(instances, 1, window, window, device=device, generator=g)
            self.pool.append((a <= 0.1).to(torch.float32))
            del a
        self.cells_per_step = instances * size * size
        self.state_bytes = instances * size * ((size + 31) // 32) * 4
        self.speed_tail = speed_tail
        if speed_tail:
            self.com = [torch.zeros(2, instances, device=device) for _ in range(2)]
            self.com_cur = 0
            self.vel = torch.zeros(2, instances, device=device)
            self.speed = torch.zeros(1, device=device)
            self.reward = torch.zeros(instances, 1, device=device)
            self.primed = torch.zeros(1, dtype=torch.int32, device=device)
        self.args = _lib.StepArgs()
        self.args.struct_size = ctypes.sizeof(_lib.StepArgs)
        env._sync_rule()

    def abi_step(self, i, tail=True):
        env, lib, a = self.env, self.lib, self.args
        a.state_in, a.state_out = env._packed.data_ptr(), env._spare.data_ptr()
        a.action, a.action_dtype, a.action_batch = \
            self.pool[i % self.pool_len].data_ptr(), self._lib.F32, self.n
        a.counters = env._counters.data_ptr()
        a.reductions = env._red_buf.data_ptr() if env.fused_reductions else None
        a.reward_zero = self.reward.data_ptr() if self.speed_tail else None
        if self.speed_tail and tail:
            # the wrapper's tail (centre of mass, velocity, batch-wide speed, reward) rides in the
            # step kernel: two centre-of-mass buffers, read one / write the other
            a.speed_com_prev = self.com[self.com_cur].data_ptr()
            a.speed_com_next = self.com[self.com_cur ^ 1].data_ptr()
            a.speed_velocity, a.speed_out = self.vel.data_ptr(), self.speed.data_ptr()
            a.speed_primed = self.primed.data_ptr()
            self.com_cur ^= 1
        else:
            a.speed_com_next = None
        if lib.carle_step_ex(env._handle, ctypes.byref(a), env._stream()):
            raise RuntimeError("C ABI call failed: " + self._lib.last_error())
        env._packed, env._spare = env._spare, env._packed

    def capture(self, steps, start=0, tail=True):
        torch = self.torch
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(steps):
                self.abi_step(start + i, tail=tail)
        if steps % 2:                      # odd K: keep the ping-pongs where the replay starts
            self.env._packed, self.env._spare = self.env._spare, self.env._packed
            if self.speed_tail and tail:
                self.com_cur ^= 1
        return graph


def time_graph(torch, graph, device, dist_on, repeats):
    """Median (and list) of `repeats` timed replays; each bracketed by barrier + sync, the time
    of a replay is the max over ranks."""
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


def max_ms(torch, ms, device, dist_on):
    if not dist_on:
        return ms
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


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
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    import __graft_entry__
    if local_rank == 0:
        __graft_entry__.build()
    if dist_on:
        torch.distributed.barrier()

    steps, warmup = args.steps, max(args.warmup, 3)
    wl = StepWorkload(torch, device, args.instances, args.size, args.window, args.rule, True,
                      args.pool_mib, speed_tail=True)
    for i in range(warmup + (warmup % 2)):
        wl.abi_step(i)
    torch.cuda.synchronize(device)
    graph = wl.capture(steps, start=warmup + 2)
    graph.replay()                                   # untimed: graph upload / first-run cost
    torch.cuda.synchronize(device)

    # repeat the K-step region until ~1.5 s of GPU time so nvidia-smi sees the load
    probe_ms, _ = time_graph(torch, graph, device, dist_on, 3)
    repeats = args.repeats or int(min(2000, max(5, 1500.0 / max(probe_ms, 1e-3))))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_region, all_ms = time_graph(torch, graph, device, dist_on, repeats)
    ms_per_step = ms_region / steps
    value = wl.cells_per_step * world / (ms_per_step * 1e-3)

    # ---- roofline: the same kernel without the wrapper's tail (the bytes the tail moves are
    # negligible; what it costs in time shows as share_of_step), K launches as one graph ----
    peak, peak_src = measured_hbm_peak()
    kgraph = wl.capture(steps, start=warmup + 2, tail=False)
    kgraph.replay()
    torch.cuda.synchronize(device)
    k_ms, _ = time_graph(torch, kgraph, device, dist_on, max(5, repeats // 4))
    clocks = sampler.stop() if sampler else None
    step_bytes = 2 * wl.state_bytes + wl.action_bytes + wl.n * 32   # state r+w, f32 action, sums
    kernel_us = 1e3 * k_ms / steps
    default_cfg = (args.instances, args.size, args.window, args.rule) == tuple(HEADLINE.values())
    traffic, traffic_src = kernel_traffic("step_strip_kernel", "cfg3") if default_cfg else (None, None)
    roofline = {
        "bound": "hbm",
        "kernel": ("step_strip_kernel (one launch per env step: TMA-staged 128-row strips + float32-"
                   "action ingestion + one Morley generation + fused SpeedDetector sums; launches "
                   "chained with programmatic dependent launch)"),
        "achieved": step_bytes / kernel_us / 1e3, "peak": peak, "unit": "GB/s",
        "frac": step_bytes / kernel_us / 1e3 / peak, "traffic": traffic,
        "traffic_source": traffic_src, "peak_source": peak_src, "us_per_launch": kernel_us,
        "algorithmic_bytes_per_launch": step_bytes, "launches_timed": steps * max(5, repeats // 4),
        "share_of_step": kernel_us / (1e3 * ms_per_step),
        "note": ("duration = CUDA-event time of a K-launch graph of this kernel alone / K (launch gaps "
                 "included); algorithmic bytes = packed state read + written (2 x 8 KiB per instance), "
                 "float32 action read (16 KiB), sums written (32 B).  The float32 actions rotate "
                 "through a pool larger than L2; the 128 MiB packed state does not fit L2 either."),
    }
    del kgraph

    # ---- e2e: public API, actions in pinned host memory ----------------------------------
    e2e, e2e_variants = (None, None) if args.no_e2e else run_e2e(args, torch, device, dist_on, world,
                                                                 steps, warmup)

    extras = None
    if not args.no_extras:
        del graph
        wl.pool = None
        torch.cuda.empty_cache()
        extras = run_extras(args, torch, device, dist_on, world, rank)

    cpu_baseline = None
    if rank == 0 and not dist_on and not args.no_cpu_baseline:
        sample_instances = min(args.instances, 256)
        r = cpu_sample(sample_instances, args.size, args.window, args.rule, True, args.cpu_seconds)
        cpu_baseline = {
            "value": r["cells"] / r["seconds"], "unit": "cell-updates/s",
            "cores": r["threads"], "kind": "port",
            "sample": (f"{r['steps']} steps x {sample_instances} instances of "
                       f"{args.size}x{args.size} in {r['seconds']:.1f} s, torch "
                       f"{torch.__version__} CPU, {r['threads']} threads (oracle/torch_port.py: "
                       f"TorchPortCARLE + TorchPortSpeedDetector; the port costs 0.95-1.09x the real "
                       f"reference on a GPU box's host cores, profiles/r2_port_vs_reference_cpu.json)")}

    if rank == 0:
        cfg = workload_config(args)
        line = {
            "metric": "cell_updates_per_sec", "value": value, "unit": "cell-updates/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": cfg,
            "measurement": {
                "obs_format": "bit-packed int32 [N,H,W/32] (obs_mode='packed'); reward [N,1] every step",
                "l2": (f"actions rotate through a {wl.pool_len}-batch pool "
                       f"({wl.pool_len * wl.action_bytes / 2**20:.0f} MiB > 126 MB L2); the packed "
                       f"state ({2 * wl.state_bytes / 2**20:.0f} MiB ping-pong) exceeds L2 as well"),
                "timing": (f"median of {repeats} CUDA-graph replays of exactly {steps} steps "
                           f"({ms_region:.3f} ms per region), each bracketed by barrier + synchronize, "
                           f"max over ranks"),
                "launches_per_step": ("ONE: step_strip_kernel with the SpeedDetector tail fused in (centre of "
                                      "mass / velocity by whoever completes an instance's sums, speed and "
                                      "the reward column by the grid's last warp)"),
                "env_steps_per_sec": 1e3 / ms_per_step * world,
                "instance_steps_per_sec": 1e3 / ms_per_step * wl.n * world,
                "numa_binding": numa},
            "gcups": value / 1e9,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "e2e_variants": e2e_variants,
            "gpu_launches": steps,
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
    """K calls of the public SpeedDetector(CARLE).step fed from pinned HOST memory; every step's
    reward crosses back to the host.  The headline `e2e` is the reference drivers' pattern
    (carle/train_mcl.py:62-69): float32 actions, `reward.cpu()` -- a synchronisation -- per step.
    The variants keep the call and change what a caller may change: the copy of action t+1
    overlapping step t with asynchronous reward read-back (rollout.host_rollout), uint8 actions,
    bit-packed actions (CARLE.pack_host_action: 1 bit per toggle)."""
    import carle_b200
    n, size, win = args.instances, args.size, args.window
    cells = n * size * size * steps * world

    def fresh_env():
        env = carle_b200.SpeedDetector(carle_b200.CARLE(
            instances=n, height=size, width=size, action_width=win, action_height=win,
            device=str(device), obs_mode="packed"))
        env.rules_from_string(args.rule)
        env.reset()
        env.inner_env.packed_universe.random_(-2**31, 2**31 - 1)
        return env

    torch.manual_seed(7 + device.index)
    host_f32 = [(torch.rand(n, 1, win, win) <= 0.1).to(torch.float32).pin_memory() for _ in range(2)]

    def as_action(env, t):
        # float32 / uint8 host tensors go to step() as they are (it copies them to the device);
        # packed int32 words are wrapped once they are on the device
        if t.dtype == torch.int32:
            return carle_b200.PackedAction(t.to(device, non_blocking=True), env.inner_env)
        return t

    def timed(fn, env, feed):
        fn(env, feed, 3)          # warm-up through the SAME call path (staging buffers, pinned pools, streams)
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        d2h = fn(env, feed, steps)
        b.record()
        torch.cuda.synchronize(device)
        wall_ms = 1e3 * (time.perf_counter() - t0)
        if dist_on:
            torch.distributed.barrier()
        return max_ms(torch, a.elapsed_time(b), device, dist_on), wall_ms, d2h

    def strict(env, feed, k):
        d2h = 0
        for i in range(k):
            r = env.step(as_action(env, feed[i % len(feed)]))[1].cpu()   # H2D in step, D2H + sync here
            d2h = r.numel() * r.element_size()
        return d2h

    def pipelined(env, feed, k):
        _, rewards = carle_b200.host_rollout(env, [feed[i % len(feed)] for i in range(k)])
        return rewards[0].numel() * rewards.element_size()

    out = {}
    env = fresh_env()
    ms, wall, d2h = timed(strict, env, host_f32)
    raw = n * win * win * 4
    hp = env.inner_env._hp                       # host-side packing state (None: the floats were shipped)
    h2d = hp["host"][0].numel() * 4 if hp else raw
    e2e = {"value": cells / (ms * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps,
           "host_input_bytes_per_step": raw,
           "host_input_gbs": raw / (ms / steps * 1e-3) / 1e9,
           "api": ("carle_b200.SpeedDetector(CARLE).step(pinned host float32 action) + reward.cpu() "
                   "every step (one synchronisation per step, as carle/train_mcl.py:62-69).  The step "
                   "bit-packs the host tensor with the library's host threads (carle_pack_action_host_copy, "
                   f"{hp['threads'] if hp else 0} threads) which also enqueue the copy of each slice of packed words as they finish it: h2d_bytes_per_step is "
                   "what crossed the bus, host_input_bytes_per_step what the caller handed over")}
    # the same call with the floats shipped as they are (host_pack=False): PCIe-bound
    env_raw = fresh_env()
    env_raw.inner_env.host_pack = False
    ms, wall, d2h = timed(strict, env_raw, host_f32)
    out["float32_unpacked_strict_sync"] = {
        "value": cells / (ms * 1e-3), "ms_per_step": ms / steps, "h2d_bytes_per_step": raw,
        "d2h_bytes_per_step": d2h, "h2d_gbs_lower_bound": raw / (ms / steps * 1e-3) / 1e9,
        "api": "CARLE(host_pack=False): the float32 action crosses the bus unpacked (4 bytes per toggle)"}
    del env_raw
    torch.cuda.empty_cache()
    ms, wall, d2h = timed(pipelined, env, host_f32)
    out["float32_pipelined"] = {
        "value": cells / (ms * 1e-3), "ms_per_step": ms / steps, "h2d_bytes_per_step": raw,
        "d2h_bytes_per_step": d2h, "h2d_gbs_lower_bound": raw / (ms / steps * 1e-3) / 1e9,
        "api": "carle_b200.host_rollout: host-side packing of action t+1 while the device runs step t, "
               "rewards copied to pinned memory asynchronously, one synchronisation per K steps"}
    host_u8 = [a.to(torch.uint8).pin_memory() for a in host_f32]
    ms, wall, d2h = timed(pipelined, env, host_u8)
    out["uint8_pipelined"] = {"value": cells / (ms * 1e-3), "ms_per_step": ms / steps,
                              "h2d_bytes_per_step": n * win * win, "d2h_bytes_per_step": d2h}
    host_packed = [env.inner_env.pack_host_action(a) for a in host_f32]
    pbytes = host_packed[0].numel() * 4
    ms, wall, d2h = timed(pipelined, env, host_packed)
    out["packed_pipelined"] = {
        "value": cells / (ms * 1e-3), "ms_per_step": ms / steps, "h2d_bytes_per_step": pbytes,
        "d2h_bytes_per_step": d2h,
        "api": "host_rollout with CARLE.pack_host_action (1 bit per toggle, packed on the host)"}
    ms, wall, d2h = timed(strict, env, host_packed)
    out["packed_strict_sync"] = {"value": cells / (ms * 1e-3), "ms_per_step": ms / steps,
                                 "h2d_bytes_per_step": pbytes, "d2h_bytes_per_step": d2h}
    del env, host_f32, host_u8, host_packed
    torch.cuda.empty_cache()
    return e2e, out


def graph_rate(torch, wl, device, dist_on, steps, repeats=7, tail=True):
    """cells/s and ms/step of `steps` ABI steps of `wl` captured into one graph."""
    for i in range(4):
        wl.abi_step(i, tail=tail)
    torch.cuda.synchronize(device)
    g = wl.capture(steps, start=4, tail=tail)
    g.replay()
    torch.cuda.synchronize(device)
    ms, _ = time_graph(torch, g, device, dist_on, repeats)
    del g
    return ms / steps


def run_extras(args, torch, device, dist_on, world, rank):
    """Secondary measurements.  Every N: configs[3] and configs[4] of BASELINE.json as stated.
    N = 1 only: the other shapes and modes."""
    import carle_b200
    from carle_b200 import _lib
    out = {}
    peak, _ = measured_hbm_peak()

    def guarded(label, fn):
        try:
            fn()
        except Exception as exc:  # extras never break the headline line
            out[label] = {"error": repr(exc)}
        torch.cuda.empty_cache()

    # ---- BASELINE configs[3]: 2^20 instances of 64x64 split over the ranks (strong scaling) ----
    def cfg4():
        total = 1 << 20
        lo, hi = carle_b200.shard_range(total, world, rank)
        wl = StepWorkload(torch, device, hi - lo, 64, 32, "B3/S23", False,
                          pool_mib=256, seed=40)
        k = 20
        ms_step = max_ms(torch, graph_rate(torch, wl, device, dist_on, k), device, dist_on)
        nbytes = 2 * wl.state_bytes + wl.action_bytes
        res = {"cell_updates_per_sec": total * 64 * 64 / (ms_step * 1e-3), "ms_per_step": ms_step,
               "scaling": "strong", "instances_total": total, "instances_per_gpu": hi - lo,
               "algorithmic_gbs_per_gpu": nbytes / (ms_step * 1e-3) / 1e9,
               "frac_of_hbm_peak": nbytes / (ms_step * 1e-3) / 1e9 / peak,
               "note": ("BASELINE configs[3] as stated: B3/S23, 2^20 instances of 64x64, 32x32 window, "
                        "float32 actions, contiguous shards (carle_b200.shard_range), one "
                        "step_stream_kernel launch per step and rank, no collective; max over ranks")}
        del wl
        torch.cuda.empty_cache()
        # the same batch with the reference's whole-batch semantics kept exact across the shards
        env = carle_b200.ShardedCARLE(instances=total, height=64, width=64, action_width=32,
                                      action_height=32, device=str(device), obs_mode="packed")
        env.reset()
        env.packed_universe.random_(-2**31, 2**31 - 1)
        acts = [(torch.rand(hi - lo, 1, 32, 32, device=device) <= 0.1).to(torch.float32)
                for _ in range(2)]
        for i in range(3):
            env.step(acts[i % 2])
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(k):
            env.step(acts[i % 2])
        b.record()
        torch.cuda.synchronize(device)
        ms = max_ms(torch, a.elapsed_time(b), device, dist_on) / k
        res["exact_whole_batch_semantics"] = {
            "cell_updates_per_sec": total * 64 * 64 / (ms * 1e-3), "ms_per_step": ms,
            "note": ("carle_b200.ShardedCARLE.step (public API, eager): step kernel with the reset "
                     "deferred + one 8-byte all-reduce of the shards' reset condition + "
                     "carle_apply_reset per step")}
        out["cfg4_sharded_2^20x64x64"] = res
    guarded("cfg4_sharded_2^20x64x64", cfg4)

    # ---- BASELINE configs[4]: one 65536^2 torus in `world` row bands ----
    def cfg5():
        from carle_b200.bigrid import BandedCARLE
        res = {}
        # bit-exact self-check at 16384^2 (16 generations with actions) against the single-GPU
        # tiled path on rank 0 -- itself pinned to the oracle by tests/
        small = 16384
        grid = BandedCARLE(small, small, halo=16, device=device)
        g = torch.Generator(device=device).manual_seed(1234 + grid.rank)
        band = torch.randint(-2**31, 2**31 - 1, (grid.band_rows, small // 32), dtype=torch.int32,
                             device=device, generator=g)
        grid.set_band(band)
        torch.manual_seed(7)
        actions = 1.0 * (torch.rand(16, 1, 1, 64, 64) <= 0.1)
        grid.step_many(16, actions)
        mine = grid.band.clone()
        if dist_on:
            parts = [torch.empty_like(mine) for _ in range(world)]
            torch.distributed.all_gather(parts, mine)
            start = [torch.empty_like(band) for _ in range(world)]
            torch.distributed.all_gather(start, band)
        else:
            parts, start = [mine], [band]
        ok = None
        if rank == 0:
            env = carle_b200.CARLE(instances=1, height=small, width=small, obs_mode="packed",
                                   device=str(device))
            env.reset()
            env.packed_universe[0].copy_(torch.cat(start))
            env.step_many(actions.to(device))
            ok = bool(torch.equal(env.packed_universe[0], torch.cat(parts)))
            del env
        grid.close()
        del grid, band, mine, parts, start
        torch.cuda.empty_cache()
        res["check_16384"] = ("bit-exact vs the single-GPU tiled path" if ok else
                              ("MISMATCH" if ok is not None else None))
        size = 65536
        grid = BandedCARLE(size, size, halo=16, device=device)
        g = torch.Generator(device=device).manual_seed(99 + grid.rank)
        grid.set_band(torch.randint(-2**31, 2**31 - 1, (grid.band_rows, size // 32),
                                    dtype=torch.int32, device=device, generator=g))
        grid.step_many(16)
        gens = 64
        times = []
        for _ in range(5):
            if dist_on:
                torch.distributed.barrier()
            torch.cuda.synchronize(device)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            grid.step_many(gens)
            b.record()
            torch.cuda.synchronize(device)
            times.append(max_ms(torch, a.elapsed_time(b), device, dist_on))
        ms = statistics.median(times)
        grid.close()
        res.update({"cell_updates_per_sec": float(size) * size * gens / (ms * 1e-3),
                    "us_per_generation": ms * 1e3 / gens, "scaling": "strong", "row_bands": world,
                    "algorithmic_gbs_per_gpu": float(size) * size * gens * 0.25 / (ms * 1e-3) / 1e9 / world,
                    "halo_bytes_per_block_per_gpu": 2 * 16 * (size // 32) * 4,
                    "note": ("BASELINE configs[4]: single 65536x65536 B3/S23 torus, free run, 16 generations "
                             "per launch in 256x256 register tiles; row bands with the edge rows stored "
                             "into the neighbours' buffers over NVLink from inside the kernel "
                             "(carle_b200.bigrid.BandedCARLE); median of 5 x 64 generations, max over ranks")})
        out["cfg5_torus_65536"] = res
    guarded("cfg5_torus_65536", cfg5)
    if dist_on:
        return out

    # ------------------------------------------------------------------ N = 1 only ----
    def shapes():
        for label, n, size, win, rule, sums, note in (
                ("cfg2_4096x128x128", 4096, 128, 32, "B3/S23", False,
                 "BASELINE configs[1]: B3/S23, 4096 x 128x128, 32x32 window, float32 actions "
                 "(step_stream_kernel; the 8 MiB packed state is L2 resident)"),
                ("cfg3_shape_life_no_sums", 16384, 256, 64, "B3/S23", False,
                 "headline shape, B3/S23, no reward sums"),
                ("cfg4_shard_131072x64x64", 131072, 64, 32, "B3/S23", False,
                 "one 8-GPU shard of configs[3] on one GPU")):
            wl = StepWorkload(torch, device, n, size, win, rule, sums, pool_mib=512)
            k = 200 if n == 4096 else 20
            ms_step = graph_rate(torch, wl, device, False, k)
            nbytes = 2 * wl.state_bytes + wl.action_bytes
            out[label] = {"cell_updates_per_sec": wl.cells_per_step / (ms_step * 1e-3),
                          "ms_per_step": ms_step, "algorithmic_gbs": nbytes / (ms_step * 1e-3) / 1e9,
                          "frac_of_hbm_peak": nbytes / (ms_step * 1e-3) / 1e9 / peak, "note": note}
            del wl
            torch.cuda.empty_cache()
    guarded("shapes", shapes)

    def api_modes():
        # public API with device-resident float32 actions: headline shape, wrapped
        n, size, win = args.instances, args.size, args.window
        env = carle_b200.SpeedDetector(carle_b200.CARLE(
            instances=n, height=size, width=size, action_width=win, action_height=win,
            device=str(device), obs_mode="packed"))
        env.rules_from_string(args.rule)
        env.reset()
        env.inner_env.packed_universe.random_(-2**31, 2**31 - 1)
        acts = [(torch.rand(n, 1, win, win, device=device) <= 0.1).to(torch.float32) for _ in range(3)]

        def eager(e, k):
            for i in range(3):
                e.step(acts[i % 3])
            torch.cuda.synchronize(device)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(k):
                e.step(acts[i % 3])
            b.record()
            torch.cuda.synchronize(device)
            return a.elapsed_time(b) / k
        ms = eager(env, 40)
        out["cfg3_speeddetector_api"] = {
            "cell_updates_per_sec": n * size * size / (ms * 1e-3), "ms_per_step": ms,
            "note": "carle_b200.SpeedDetector(CARLE(obs_mode='packed')).step(device float32 action), "
                    "eager: ONE carle_step_ex launch per step (the wrapper's tail is fused into the step kernel), "
                    "fresh reward [N,1] every step"}
        k = 16
        plan = carle_b200.RolloutPlan(env, torch.stack([acts[i % 3] for i in range(k)]))
        plan.run()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            plan.run()
        b.record()
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b) / (5 * k)
        out["cfg3_speeddetector_rollout_plan"] = {
            "cell_updates_per_sec": n * size * size / (ms * 1e-3), "ms_per_step": ms,
            "note": "the same wrapped env, 16 steps captured by carle_b200.RolloutPlan and replayed as "
                    "one CUDA graph (sync-free rollout driver)"}
        del plan, env
        torch.cuda.empty_cache()
        # strict drop-in mode: float32 observation materialised by the step kernel
        for label, n2, s2, w2 in (("float32_obs_api_cfg2", 4096, 128, 32),
                                  ("float32_obs_api_cfg3", 4096, 256, 64)):
            e = carle_b200.CARLE(instances=n2, height=s2, width=s2, action_width=w2, action_height=w2,
                                 device=str(device), obs_mode="float32")
            e.reset()
            e.packed_universe.random_(-2**31, 2**31 - 1)
            acts2 = [(torch.rand(n2, 1, w2, w2, device=device) <= 0.1).to(torch.float32) for _ in range(8)]
            for i in range(5):
                e.step(acts2[i])
            torch.cuda.synchronize(device)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k2 = 60
            a.record()
            for i in range(k2):
                e.step(acts2[i % 8])
            b.record()
            torch.cuda.synchronize(device)
            ms = a.elapsed_time(b) / k2
            cells = n2 * s2 * s2
            out[label] = {"cell_updates_per_sec": cells / (ms * 1e-3), "ms_per_step": ms,
                          "frac_of_4.25B_per_cell_ceiling": cells * 4.25 / (ms * 1e-3) / 1e9 / peak,
                          "note": f"CARLE.step, {n2} x {s2}x{s2}, device float32 actions, float32 obs written "
                                  "by the step kernel itself (no unpack launch), eager"}
            del e, acts2
            torch.cuda.empty_cache()
        # configs[1] through the public API with packed obs + per-call latency at configs[0]
        e = carle_b200.CARLE(instances=4096, height=128, width=128, action_width=32, action_height=32,
                             device=str(device), obs_mode="packed")
        e.reset()
        e.packed_universe.random_(-2**31, 2**31 - 1)
        acts2 = [(torch.rand(4096, 1, 32, 32, device=device) <= 0.1).to(torch.float32) for _ in range(24)]
        for i in range(10):
            e.step(acts2[i])
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(400):
            e.step(acts2[i % 24])
        b.record()
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b) / 400
        out["cfg2_public_api_packed_obs"] = {
            "cell_updates_per_sec": 4096 * 128 * 128 / (ms * 1e-3), "us_per_step": ms * 1e3,
            "note": "CARLE(obs_mode='packed').step(device float32 action), eager, 400 calls"}
        plan = carle_b200.RolloutPlan(e, torch.stack(acts2))
        plan.run()
        torch.cuda.synchronize(device)
        a.record()
        for _ in range(10):
            plan.run()
        b.record()
        torch.cuda.synchronize(device)
        ms = a.elapsed_time(b) / 240
        out["cfg2_rollout_plan"] = {"cell_updates_per_sec": 4096 * 128 * 128 / (ms * 1e-3),
                                    "us_per_step": ms * 1e3,
                                    "note": "CARLE.step x 24 captured by RolloutPlan, replayed"}
        del plan, e, acts2
        env1 = carle_b200.CARLE(instances=1, height=64, width=64, action_width=32, action_height=32,
                                device=str(device))
        env1.reset()
        env1.universe = (torch.rand(1, 1, 64, 64, device=device) < 0.5).float()
        dev_acts = [1.0 * (torch.rand(1, 1, 32, 32, device=device) <= 0.1) for _ in range(16)]
        host_acts = [x.cpu() for x in dev_acts]
        lat = {}
        for label, acts3, read in (("device_action", dev_acts, False),
                                   ("host_action_reward_read", host_acts, True)):
            for i in range(50):
                env1.step(acts3[i % 16])
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            for i in range(2000):
                r = env1.step(acts3[i % 16])[1]
                if read:
                    r.cpu()
            torch.cuda.synchronize(device)
            lat[label + "_us_per_call"] = (time.perf_counter() - t0) / 2000 * 1e6
        lat["note"] = ("BASELINE configs[0] shape (1 x 64x64, 32x32 window, strict float32 obs): wall-clock "
                       "latency of carle_b200.CARLE.step; the reference's torch CPU step takes ~850 us "
                       "per call (BASELINE.md section 2)")
        out["cfg1_api_latency"] = lat
    guarded("api_modes", api_modes)

    def fused_rollouts():
        n, size, win = 4096, 128, 32
        env2 = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                                action_height=win, device=str(device), obs_mode="packed")
        env2.reset()
        env2.packed_universe.random_(-2**31, 2**31 - 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        env2.step_many(64)
        torch.cuda.synchronize(device)
        a.record()
        for _ in range(reps):
            env2.step_many(64)
        b.record()
        torch.cuda.synchronize(device)
        out["free_run_k64"] = {"cell_updates_per_sec": n * size * size * 64 * reps / (a.elapsed_time(b) * 1e-3),
                               "note": "4096 x 128x128 zero-action free run, 64 generations per launch "
                                       "(register-resident, integer-pipe bound)"}
        lib = _lib.load()
        env2._sync_rule()
        words = env2._action_buf

        def agent_steps(k0, k):
            for i in range(k):
                rc = lib.carle_step_random(env2._handle, env2._packed.data_ptr(),
                                           env2._spare.data_ptr(), 7, k0 + i, 0.1, n,
                                           words.data_ptr(), env2._counters.data_ptr(), None,
                                           env2._stream())
                assert rc == 0, _lib.last_error()
                env2._packed, env2._spare = env2._spare, env2._packed
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
            "note": "4096 x 128x128 random-agent rollout entirely on the GPU: Bernoulli(0.1) toggles drawn "
                    "with Philox inside the step kernel (carle_step_random), one launch per env step"}
    guarded("fused_rollouts", fused_rollouts)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
