"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: instance sharding, the
max-over-ranks timing reduction, and the row-band halo protocol of bigrid.py (band layout,
T generations per exchange, halo depth, edge-row routing, action row shift).  The compute
inside each simulated rank is the numpy ORACLE — the product's kernels need a GPU — so what
is verified here is the protocol: that bands + halos exchanged every T generations reproduce
the whole-torus evolution exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import carle_oracle as oc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, size, halo, gens, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from carle_b200.bigrid import band_layout
    from carle_b200.sharding import shard_range, max_over_ranks, sum_over_ranks

    # ---- sharding / timing plumbing ----
    lo, hi = shard_range(1000003, world, rank)
    assert sum_over_ranks(hi - lo) == 1000003
    assert max_over_ranks(10.0 + rank) == 10.0 + world - 1

    # ---- band protocol ----
    rng = np.random.default_rng(5)                       # same soup on every rank
    soup = (rng.random((size, size)) < 0.4).astype(np.uint8)
    win = 8
    actions = (rng.random((gens, 1, 1, win, win)) <= 0.2).astype(np.float32)
    row0, rows, up, dn = band_layout(size, world, rank, halo)
    geo = oc.window_geometry(size, size, win, win)
    act_row0, act_col0 = geo[2], geo[3]
    local = np.zeros((rows + 2 * halo, size), dtype=np.uint8)
    local[halo:halo + rows] = soup[row0:row0 + rows]

    def exchange():
        """my first/last `halo` band rows -> neighbours' bottom/top halos"""
        top = torch.from_numpy(local[halo:2 * halo].copy())
        bot = torch.from_numpy(local[rows:rows + halo].copy())
        from_dn = torch.empty_like(top)      # lower neighbour's top rows -> my bottom halo
        from_up = torch.empty_like(bot)      # upper neighbour's bottom rows -> my top halo
        reqs = [dist.isend(top, up, tag=1), dist.isend(bot, dn, tag=2),
                dist.irecv(from_dn, dn, tag=1), dist.irecv(from_up, up, tag=2)]
        for r in reqs:
            r.wait()
        local[rows + halo:] = from_dn.numpy()
        local[:halo] = from_up.numpy()

    exchange()
    done = 0
    while done < gens:
        t = min(halo, gens - done)
        for g in range(t):
            a = actions[done + g, 0, 0] != 0
            for r in range(win):                         # action rows that fall in my buffer
                lr = act_row0 + r - (row0 - halo)        # grid row -> local row (act_row_shift)
                for cand in (lr, lr + size, lr - size):  # torus: halo rows may be wrapped
                    if 0 <= cand < local.shape[0]:
                        local[cand, act_col0:act_col0 + win] ^= a[r].astype(np.uint8)
            # local torus update: wrong only within g+1 rows of the buffer edge (halo absorbs)
            local[:] = oc.life_like_update(local[None], [3], [2, 3])[0]
        exchange()
        done += t
    np.save(os.path.join(out_dir, f"band{rank}.npy"), local[halo:halo + rows])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("size,halo,gens", [(64, 8, 19), (96, 16, 40)])
def test_band_protocol_two_ranks(tmp_path, size, halo, gens):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, size, halo, gens, str(tmp_path)), nprocs=world,
             join=True)
    got = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)])
    rng = np.random.default_rng(5)
    soup = (rng.random((size, size)) < 0.4).astype(np.uint8)
    actions = (rng.random((gens, 1, 1, 8, 8)) <= 0.2).astype(np.float32)
    ref = oc.OracleCARLE(width=size, height=size, action_width=8, action_height=8)
    ref.reset()
    ref.universe = soup[None].copy()
    for t in range(gens):
        want = ref.step(actions[t, 0])[0]
    assert np.array_equal(got, want[0])


def test_band_layout_and_shard_range():
    from carle_b200.bigrid import band_layout
    from carle_b200.sharding import shard_range
    assert band_layout(65536, 8, 3, 16) == (24576, 8192, 2, 4)
    assert band_layout(65536, 8, 0, 16)[2:] == (7, 1)
    assert band_layout(65536, 8, 7, 16)[2:] == (6, 0)
    with pytest.raises(ValueError):
        band_layout(100, 8, 0, 16)
    with pytest.raises(ValueError):
        band_layout(64, 8, 0, 16)             # 8-row bands are shallower than the halo
    spans = [shard_range(10, 4, r) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_range(1 << 20, 8, r)[1] - shard_range(1 << 20, 8, r)[0] for r in range(8)] == [131072] * 8
