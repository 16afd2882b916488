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


def _shard_worker(rank, world, port, n, size, win, out_dir):
    """One rank's shard of a batch, stepped with the numpy oracle; the whole-batch couplings
    (master reset, SpeedDetector.speed) go through the product's combine_shards exactly as
    sharding.ShardedCARLE / ShardedSpeedDetector use it."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from carle_b200.sharding import shard_range, combine_shards
    lo, hi = shard_range(n, world, rank)
    rng = np.random.default_rng(11)                      # the same stream on every rank
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    env = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=hi - lo)
    env.rules_from_string("B368/S245")
    env.reset()
    env.universe = soup[lo:hi].copy()
    mask = oc.outside_window_mask(env)
    com_prev, rewards, states = None, [], []
    for t in range(8):
        a = (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)
        if t == 3:
            a[:] = 1.0                                   # whole batch all ones: reset everywhere
        if t == 5:
            a[:] = 1.0
            a[n - 1, 0, 2, 3] = 0.0                      # only the LAST shard misses: no reset anywhere
        mine = a[lo:hi]
        # the shard's own step with the reset deferred (what the kernel does with defer_reset)
        env.apply_action(mine)
        cond = bool(np.all(mine == 1.0))
        env.universe = oc.life_like_update(env.universe, env.birth, env.survive)
        partial = torch.tensor([0.0 if cond else 1.0], dtype=torch.float64)
        combine_shards(partial)
        if float(partial[0]) == 0.0:                     # carle_apply_reset
            env.universe[:] = 0
        live, sh, sw = oc.speed_sums(env.universe, mask)
        denom = live.astype(np.float32) + np.float32(1e-7)
        com = np.stack([sh.astype(np.float32) / denom, sw.astype(np.float32) / denom])
        reward = np.zeros((hi - lo, 1), dtype=np.float32)
        if com_prev is not None:
            v = (com_prev - com).astype(np.float64)
            sumsq = torch.tensor([float(np.sum(v * v))], dtype=torch.float64)
            combine_shards(sumsq)
            reward += np.float32(np.sqrt(float(sumsq[0])))
        com_prev = com
        rewards.append(reward)
        states.append(env.universe.copy())
    np.save(os.path.join(out_dir, f"states{rank}.npy"), np.stack(states))
    np.save(os.path.join(out_dir, f"rewards{rank}.npy"), np.stack(rewards))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_semantics_two_ranks(tmp_path):
    """Whole-batch master reset and the batch-wide SpeedDetector speed over two shards ==
    the oracle on the unsharded batch (reference semantics: env.py:208, mcl.py:787-795)."""
    world, n, size, win = 2, 7, 64, 32
    port = _free_port()
    mp.spawn(_shard_worker, args=(world, port, n, size, win, str(tmp_path)), nprocs=world, join=True)
    states = np.concatenate([np.load(tmp_path / f"states{r}.npy") for r in range(world)], axis=1)
    rewards = np.concatenate([np.load(tmp_path / f"rewards{r}.npy") for r in range(world)], axis=1)
    rng = np.random.default_rng(11)
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    ref = oc.OracleSpeedDetector(oc.OracleCARLE(width=size, height=size, action_width=win,
                                                action_height=win, instances=n))
    ref.env.rules_from_string("B368/S245")
    ref.reset()
    ref.env.universe = soup.copy()
    for t in range(8):
        a = (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)
        if t == 3:
            a[:] = 1.0
        if t == 5:
            a[:] = 1.0
            a[n - 1, 0, 2, 3] = 0.0
        obs, reward, _, _ = ref.step(a)
        assert np.array_equal(states[t], obs), t
        np.testing.assert_allclose(rewards[t], np.broadcast_to(reward, (n, 1)), rtol=2e-6, atol=1e-6)


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
