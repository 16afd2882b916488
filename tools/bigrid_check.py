"""Multi-GPU row-band giant grid: correctness check and throughput (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tools/bigrid_check.py [--size 65536] [--gens 64] [--halo 16] [--check]

Every rank builds the same pseudo-random band deterministically; with --size <= 16384 the
result is compared bit-for-bit with the single-GPU tiled path on rank 0 ("BIGRID CHECK OK").
Timing: CUDA events on every rank, max over ranks, printed as one JSON line by rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--gens", type=int, default=64)
    ap.add_argument("--halo", type=int, default=16)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import carle_b200
    from carle_b200.bigrid import BandedCARLE

    size, wpr = args.size, args.size // 32
    grid = BandedCARLE(size, size, halo=args.halo, device=device)
    g = torch.Generator(device=device).manual_seed(1234 + grid.rank)
    band = torch.randint(-2**31, 2**31 - 1, (grid.band_rows, wpr), dtype=torch.int32,
                         device=device, generator=g)
    grid.set_band(band)
    win = 64
    torch.manual_seed(7)
    actions = 1.0 * (torch.rand(args.gens, 1, 1, win, win) <= 0.1)

    check = (not args.no_check) and size <= 16384
    if check:
        grid.step_many(args.gens, actions)
        mine = grid.band.clone()
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            start = [torch.empty_like(band) for _ in range(world)]
            dist.all_gather(start, band)
        else:
            parts, start = [mine], [band]
        if rank == 0:
            env = carle_b200.CARLE(instances=1, height=size, width=size, obs_mode="packed",
                                   device=str(device))
            env.reset()
            env.packed_universe[0].copy_(torch.cat(start))
            env.step_many(actions.to(device))
            ok = torch.equal(env.packed_universe[0], torch.cat(parts))
            print("BIGRID CHECK", "OK" if ok else "FAILED", f"size={size} ranks={world} "
                  f"gens={args.gens} halo={args.halo}", flush=True)
            if not ok:
                sys.exit(1)
        grid.set_band(band)

    # ---- timing: free run, CUDA events, max over ranks ----
    grid.step_many(args.halo)                       # warm-up block
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    grid.step_many(args.gens)
    b.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        cells = size * size * args.gens
        print(json.dumps({"workload": f"single {size}x{size} Life torus, {world} row bands, "
                                      f"halo/temporal block {args.halo}",
                          "n_gpus": world, "generations": args.gens,
                          "us_per_generation": float(ms) * 1e3 / args.gens,
                          "cell_updates_per_sec": cells / (float(ms) * 1e-3),
                          "algorithmic_gbs_per_gpu": cells * 0.25 / (float(ms) * 1e-3) / 1e9 / world,
                          "halo_bytes_per_block_per_gpu": 2 * args.halo * wpr * 4}), flush=True)
    grid.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
