"""One batch of universes sharded over the GPUs of a node (BASELINE config 4: 2^20 instances
of 64 x 64 over 2 / 4 / 8 B200).

Batched instances are independent (the reference's only parallelism is the batch dimension
of one tensor, carle/env.py:46-48), so they shard embarrassingly: rank ``r`` of ``G`` owns the
contiguous range ``shard_range(N, G, r)`` and steps it with its own ``CARLE`` -- no collective
on the data path.  Two things in the reference nevertheless span the WHOLE batch:

* the master reset fires when the mean of the entire action tensor is 1.0 (env.py:208), and
* ``SpeedDetector`` adds ONE scalar, the norm of the velocity over all instances, to every
  instance's reward (mcl.py:787-795).

``ShardedCARLE`` / ``ShardedSpeedDetector`` keep those semantics exact: the step kernel runs
with ``defer_reset`` (it never clears on its own; whether the shard's condition held is left
in the device counters), every rank contributes "my shard's condition failed" to an 8-byte
all-reduce, ``carle_apply_reset`` consumes the result, and the wrapper sums the shards'
squared velocities with a second 8-byte all-reduce before the reward update -- all
stream-ordered, no host synchronisation (NCCL on GPUs; the combination logic itself is
backend-agnostic and runs under gloo in the CPU tests).  The data path -- the step kernel --
has no collective; a rollout that does not need the two couplings uses a plain ``CARLE`` per
rank and pays for none at all.
"""
import torch
import torch.distributed as dist

from . import _lib
from .env import CARLE
from .mcl import SpeedDetector


def shard_range(total, world_size, rank):
    """Contiguous, balanced ``[start, stop)`` of ``total`` instances owned by ``rank``."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def max_over_ranks(value, device="cpu", group=None):
    """Max of a python float over all ranks (the time that bounds a multi-GPU step)."""
    if _world(group)[0] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device="cpu", group=None):
    if _world(group)[0] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def combine_shards(partial, group=None):
    """The collective of a sharded step: a float64 tensor of per-shard partials (``1.0 if this
    shard's reset condition FAILED else 0.0``: sum == 0 <=> the reference's whole-batch test
    holds, env.py:208; or the shard's sum of v^2: ``sqrt`` of the sum is the reference's
    batch-wide ``speed``, mcl.py:789), summed in place over the ranks.  Stream-ordered: no host
    synchronisation."""
    if _world(group)[0] > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


class ShardedCARLE(CARLE):
    """This rank's shard of ONE batch of ``instances`` universes.

    Same constructor as ``CARLE`` (``instances`` is the size of the WHOLE batch) plus
    ``group``; ``step(action)`` takes the shard's rows of the action tensor (``local_slice``
    tells which) or a batch-1 action, and behaves exactly as the reference would on the whole
    batch: the master reset fires on every shard, or on none.  (The condition is evaluated per
    shard -- "every toggle of the shard is 1.0", or the reference's mean test when the shard's
    action holds values other than 0 and 1 -- and the shards are AND-ed.)"""

    def __init__(self, group=None, **kwargs):
        self.group = group
        self.world, self.rank = _world(group)
        total = int(kwargs.get("instances", 1))
        start, stop = shard_range(total, self.world, self.rank)
        if stop == start:
            raise ValueError(f"rank {self.rank} of {self.world} would own no instance of {total}")
        kwargs = dict(kwargs, instances=stop - start)
        super().__init__(**kwargs)
        self.total_instances = total
        self.local_slice = slice(start, stop)
        self.defer_reset = True
        self._partial = None            # float64 [1]: shards whose reset condition failed
        self._decision = None           # int32 [1]: non-zero <=> the whole batch resets

    def step(self, action):
        obs, reward, done, info = super().step(action)
        dev = self.my_device
        if self._partial is None:
            self._partial = torch.zeros(1, dtype=torch.float64, device=dev)
            self._decision = torch.zeros(1, dtype=torch.int32, device=dev)
        # [this shard's condition failed] summed over the ranks: 0 <=> the whole batch is all ones
        torch.sub(1.0, self._counters[_lib.CNT_LAST_RESET_COND:_lib.CNT_LAST_RESET_COND + 1],
                  out=self._partial)
        combine_shards(self._partial, self.group)
        self._decision.copy_(self._partial == 0.0)
        # clears state, observation and sums on every shard if the batch asked for the master
        # reset; a launch that exits at once otherwise
        obs_ptr, obs_code = None, _lib.F32
        if self.obs_mode != "packed":
            obs_ptr = obs.data_ptr()
            obs_code = _lib.F32 if self.obs_mode == "float32" else _lib.U8
        red = self.last_reductions
        _lib.check(self._lib.carle_apply_reset(
            self._handle, self._decision.data_ptr(), self._packed.data_ptr(),
            obs_ptr, obs_code, red.data_ptr() if red is not None else None,
            self._counters.data_ptr(), self._stream()), "carle_apply_reset")
        return obs, reward, done, info


class ShardedSpeedDetector(SpeedDetector):
    """``SpeedDetector`` over a ``ShardedCARLE``: the centre-of-mass sums are the shard's own
    (fused in its step kernel), the squared velocities are summed over the shards (a second
    8-byte all-reduce) and every instance's reward gets the batch-wide ``speed``
    (mcl.py:787-795) -- the value the reference computes on the unsharded batch, up to the
    order of the float additions inside the norm."""

    def step(self, action):
        inner = self.inner_env
        if not isinstance(inner, ShardedCARLE):
            raise TypeError("ShardedSpeedDetector wraps a ShardedCARLE")
        obs, reward, done, info = self.env.step(action)
        red = inner.last_reductions
        self._speed_buffers(red.shape[0])
        if getattr(self, "_sumsq_buf", None) is None or self._sumsq_buf.device != red.device:
            self._sumsq_buf = torch.zeros(1, dtype=torch.float64, device=red.device)
        # this shard's part of mcl.py:777-789 (centre of mass, velocity, sum of v^2; on the
        # wrapper's first step the sum stays 0) -- behind carle_apply_reset, on the final sums
        inner._speed_tail(red, self._com[self._com_cur], False, self._velocity_buf, self._speed_buf,
                          None, sumsq=self._sumsq_buf, primed=self._primed)
        self._live_src = red
        total = self._sumsq_buf.clone()
        combine_shards(total, inner.group)
        speed = torch.sqrt(total[0]).to(torch.float32)                     # mcl.py:789
        if self._steps_seen:
            self.velocity = self._velocity_buf
            self.speed = speed
        self._steps_seen += 1
        reward += speed                                                    # mcl.py:795
        return obs, reward, done, info
