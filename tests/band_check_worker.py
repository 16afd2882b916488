"""Row-band giant grid against the ORACLE (run under torchrun by test_parity_gpu.py / by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tests/band_check_worker.py --size 4096 --gens 40 --mode whole
    ... --size 65536 --gens 16 --mode stripes

mode whole   : every rank's band after `gens` generations WITH actions (window straddling a band
               boundary) == the numpy oracle on the whole torus (sizes the oracle handles: <= 8192).
mode stripes : sizes the oracle cannot hold.  Around EVERY band boundary (where the NVLink halo
               stores and the neighbour flags act) and around two in-band tile seams, a stripe of
               64 + 2*gens rows of the initial state is evolved `gens` free-run generations by the
               oracle (full width, so the horizontal torus wrap is exact; the stripe's own top and
               bottom `gens` rows are its halo) and its 64 middle rows are compared with what the
               GPUs produced.
Prints "BAND ORACLE CHECK OK ..." on rank 0 or exits 1.  Test infrastructure: imports oracle/."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from oracle import carle_oracle as oc


def unpack(t, width):
    return np.unpackbits(t.cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[:, :width]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--gens", type=int, default=40)
    ap.add_argument("--halo", type=int, default=16)
    ap.add_argument("--mode", default="whole", choices=["whole", "stripes"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from carle_b200.bigrid import BandedCARLE
    size, wpr, gens = args.size, args.size // 32, args.gens
    grid = BandedCARLE(size, size, halo=args.halo, device=device)
    g = torch.Generator(device=device).manual_seed(4321 + grid.rank)
    band = torch.randint(-2**31, 2**31 - 1, (grid.band_rows, wpr), dtype=torch.int32, device=device,
                         generator=g)
    grid.set_band(band)

    def gather(t):
        if world == 1:
            return [t]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return parts

    ok = True
    if args.mode == "whole":
        torch.manual_seed(7)
        actions = 1.0 * (torch.rand(gens, 1, 1, 64, 64) <= 0.1)
        grid.step_many(gens, actions)
        start, final = gather(band), gather(grid.band.clone())
        if rank == 0:
            ref = oc.OracleCARLE(width=size, height=size, action_width=64, action_height=64, instances=1)
            ref.reset()
            ref.universe = unpack(torch.cat(start), size)[None].copy()
            for t in range(gens):
                want = ref.step(actions[t, 0].numpy())[0]
            ok = bool(np.array_equal(unpack(torch.cat(final), size), want[0]))
    else:
        k = gens
        edge = 32 + k                                   # rows each rank contributes per boundary
        assert grid.band_rows >= 4 * edge
        pitch = 256 - 2 * args.halo                     # rows a tile produces (256-row tiles, halo discarded)
        mid = (grid.band_rows // 2) // pitch * pitch - 32   # 64 rows straddling a seam between two tile rows
        first = {"top": band[:edge], "bot": band[-edge:], "mid": band[mid - k:mid + 64 + k]}
        grid.step_many(k)
        out = grid.band
        last = {"top": out[:32].clone(), "bot": out[-32:].clone(), "mid": out[mid:mid + 64].clone()}
        first = {key: gather(v) for key, v in first.items()}
        last = {key: gather(v) for key, v in last.items()}
        if rank == 0:
            def evolve(rows):                           # free run on a stripe, full width
                u = rows[None].copy()
                for _ in range(k):
                    u = oc.life_like_update(u, [3], [2, 3])
                return u[0]
            checked = 0
            for b in range(world):                      # boundary between rank b-1 (above) and rank b
                above = (b - 1) % world
                stripe = np.concatenate([unpack(first["bot"][above], size), unpack(first["top"][b], size)])
                want = evolve(stripe)[k:k + 64]
                got = np.concatenate([unpack(last["bot"][above], size), unpack(last["top"][b], size)])
                ok &= bool(np.array_equal(got, want))
                checked += 1
            for r in range(world):                      # a seam between two tile rows inside every band
                want = evolve(unpack(first["mid"][r], size))[k:k + 64]
                ok &= bool(np.array_equal(unpack(last["mid"][r], size), want))
                checked += 1
            print(f"stripes checked: {checked} x 64 rows x {size} columns, {k} generations", flush=True)
    if rank == 0:
        print("BAND ORACLE CHECK", "OK" if ok else "FAILED",
              f"mode={args.mode} size={size} ranks={world} gens={gens} halo={args.halo}", flush=True)
    grid.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
