"""One giant Life-like torus split into row bands over the GPUs of a node (BASELINE
config 5: a single 65536 x 65536 grid over 8 x B200).

The reference cannot express this (its universe is one dense tensor on one device,
carle/env.py:136, wrap done by ``padding_mode="circular"``, env.py:98-104); the semantics
here are exactly those of ``CARLE.step`` on the whole ``H x W`` torus: the action window in
the middle of the grid, the batch-wide master reset, one generation per step.

Layout: rank ``r`` of ``G`` owns rows ``[r*H/G, (r+1)*H/G)`` in a local packed buffer
``[halo | band | halo]``.  ``step_many(K)`` runs temporal blocks of ``T <= halo``
generations: one ``carle_band_step`` launch per block advances the band T generations in
register tiles and stores the freshly computed edge rows straight into the two neighbours'
buffers over NVLink (peer-mapped pointers, CUDA IPC).  Consecutive blocks are ordered by the
kernels themselves: every rank owns four peer-mapped int32 words; a launch spins until BOTH
NEIGHBOURS (nobody else) have finished as many blocks as it has, and its last CTA publishes
its new count in the neighbours' words with a system-scope release store -- no NCCL and no
host on the hot path, ``step_many`` is a plain chain of launches (``sync="nccl"`` keeps the
earlier 4-byte all-reduce per block for comparison).  Per block each GPU sends 2 x T rows x
W/8 bytes — 256 KiB at T = 16 on the 65536-wide grid — so the exchange is latency-, not
bandwidth-bound, and is hidden inside the compute kernel.

``world_size == 1`` degenerates to a band whose neighbours are itself (no IPC), which is how
the single-GPU tests cover this code path.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .env import _rule_mask


def band_layout(height, world_size, rank, halo):
    """(band_row0, band_rows, up_rank, down_rank) of ``rank``; rows split evenly."""
    if height % world_size:
        raise ValueError(f"height {height} is not divisible by {world_size} ranks")
    rows = height // world_size
    if rows % 8 or rows < halo:
        raise ValueError(f"band of {rows} rows must be a multiple of 8 and >= halo {halo}")
    return rank * rows, rows, (rank - 1) % world_size, (rank + 1) % world_size


class _DeviceBuffer:
    """cudaMalloc'ed int32 matrix exposed to torch through __cuda_array_interface__ (CUDA IPC
    exports whole allocations, so these must not come from torch's caching allocator)."""

    def __init__(self, lib, device_index, rows, cols):
        self._lib, self._device = lib, device_index
        ptr = ctypes.c_void_p()
        _lib.check(lib.carle_dev_alloc(device_index, rows * cols * 4, ctypes.byref(ptr)),
                   "carle_dev_alloc")
        self.ptr = ptr.value
        self.__cuda_array_interface__ = {
            "shape": (rows, cols), "typestr": "<i4", "data": (self.ptr, False), "version": 2}

    def free(self):
        if self.ptr:
            self._lib.carle_dev_free(self._device, ctypes.c_void_p(self.ptr))
            self.ptr = 0


class BandedCARLE:
    """Row-band sharded giant grid; one instance per process (rank)."""

    def __init__(self, height, width, rule="B3/S23", halo=16, action_height=64,
                 action_width=64, device=None, group=None, sync="device"):
        if sync not in ("device", "nccl"):
            raise ValueError("sync must be 'device' (neighbour flags) or 'nccl' (all-reduce per block)")
        self._lib = _lib.load()
        self.sync = sync
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device if device is not None else
                                   f"cuda:{torch.cuda.current_device()}")
        self.height, self.width, self.halo = height, width, halo
        self.wpr = width // 32
        (self.band_row0, self.band_rows, self.up_rank, self.dn_rank) = band_layout(
            height, self.world, self.rank, halo)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.carle_band_create(
            ctypes.byref(handle), self.device.index, height, width, action_height,
            action_width, self.band_row0, self.band_rows, halo), "carle_band_create")
        self._handle = handle
        geo = (ctypes.c_int32 * 8)()
        _lib.check(self._lib.carle_geometry(handle, ctypes.byref(geo)))
        self.row0, self.col0, self._aw, self._ah, _, self._awpr, _, self._aw0 = list(geo)
        self.set_rule(rule)
        rows = self.band_rows + 2 * halo
        self._bufs = [_DeviceBuffer(self._lib, self.device.index, rows, self.wpr)
                      for _ in range(2)]
        # (third allocation: the four synchronisation words, see carle_band_step)
        self._bufs.append(_DeviceBuffer(self._lib, self.device.index, 1, 4))
        self._tensors = [torch.as_tensor(b, device=self.device) for b in self._bufs[:2]]
        self._cur = 0
        self._counters = torch.zeros(8, dtype=torch.int64, device=self.device)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._peer_up, self._peer_dn, self._opened = self._map_peers()

    # -------------------------------------------------------------- plumbing ----
    def _map_peers(self):
        """Exchange CUDA IPC handles of both buffers and map the two neighbours'."""
        if self.world == 1:
            mine = [b.ptr for b in self._bufs]
            return mine, mine, []
        handles = torch.zeros(len(self._bufs), 64, dtype=torch.uint8)
        for i, b in enumerate(self._bufs):
            raw = (ctypes.c_ubyte * 64)()
            _lib.check(self._lib.carle_ipc_export(ctypes.c_void_p(b.ptr), raw),
                       "carle_ipc_export")
            handles[i] = torch.tensor(list(raw), dtype=torch.uint8)
        gathered = [torch.zeros_like(handles, device=self.device) for _ in range(self.world)]
        dist.all_gather(gathered, handles.to(self.device), group=self.group)
        opened, maps = [], {}
        for peer in {self.up_rank, self.dn_rank}:
            ptrs = []
            for i in range(len(self._bufs)):
                raw = (ctypes.c_ubyte * 64)(*gathered[peer][i].cpu().tolist())
                out = ctypes.c_void_p()
                _lib.check(self._lib.carle_ipc_open(raw, ctypes.byref(out)), "carle_ipc_open")
                ptrs.append(out.value)
                opened.append(out.value)
            maps[peer] = ptrs
        return maps[self.up_rank], maps[self.dn_rank], opened

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _barrier(self):
        """Stream-ordered cross-rank barrier: a 4-byte all-reduce (no host sync)."""
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)

    def set_rule(self, rule):
        parts = rule.split("/")
        self.birth = sorted({int(c) for c in parts[0] if c in "012345678"})
        self.survive = sorted({int(c) for c in parts[1] if c in "012345678"})
        _lib.check(self._lib.carle_set_rule(self._handle, _rule_mask(self.birth),
                                            _rule_mask(self.survive)), "carle_set_rule")

    # ------------------------------------------------------------------ state ----
    @property
    def band(self):
        """int32 view ``[band_rows, W/32]`` of this rank's rows (packed, LSB = lowest column)."""
        return self._tensors[self._cur][self.halo:self.halo + self.band_rows]

    def set_band(self, packed_rows):
        """Overwrite this rank's band and refresh the neighbours' halos (collective)."""
        self.band.copy_(packed_rows.to(self.device, torch.int32))
        self.sync_halos()

    def sync_halos(self):
        self._barrier()                      # every rank has finished writing its band
        cur = self._cur
        _lib.check(self._lib.carle_band_push_halos(
            self._handle, ctypes.c_void_p(self._bufs[cur].ptr),
            ctypes.c_void_p(self._peer_up[cur]), ctypes.c_void_p(self._peer_dn[cur]),
            self._stream()), "carle_band_push_halos")
        self._barrier()                      # halos have landed everywhere

    # ------------------------------------------------------------------- step ----
    def step_many(self, generations, actions=None):
        """Advance the whole grid ``generations`` steps.  ``actions``: float32/uint8
        ``[K, 1, 1, aw, ah]`` (the same tensor on every rank) or None for a free run."""
        packed = flags = None
        if actions is not None:
            if actions.shape[0] != generations:
                raise ValueError("one action per generation")
            flat = actions.to(self.device).contiguous()
            flat = flat.to(torch.uint8) if flat.dtype == torch.bool else flat
            if flat.dtype not in (torch.float32, torch.uint8):
                flat = flat.to(torch.float32)
            packed = torch.empty((generations, max(self._aw, 1), self._awpr),
                                 dtype=torch.int32, device=self.device)
            flags = torch.zeros((generations, 2), dtype=torch.int32, device=self.device)
            code = _lib.U8 if flat.dtype == torch.uint8 else _lib.F32
            _lib.check(self._lib.carle_pack_action(
                self._handle, flat.data_ptr(), code, 1, generations, packed.data_ptr(),
                flags.data_ptr(), self._stream()), "carle_pack_action")
        device_sync = self.sync == "device"
        done = 0
        while done < generations:
            t = min(self.halo, generations - done)
            cur, nxt = self._cur, self._cur ^ 1
            _lib.check(self._lib.carle_band_step(
                self._handle, ctypes.c_void_p(self._bufs[cur].ptr),
                ctypes.c_void_p(self._bufs[nxt].ptr), ctypes.c_void_p(self._peer_up[nxt]),
                ctypes.c_void_p(self._peer_dn[nxt]), t,
                ctypes.c_void_p(packed[done].data_ptr()) if packed is not None else None,
                ctypes.c_void_p(flags[done].data_ptr()) if flags is not None else None,
                ctypes.c_void_p(self._counters.data_ptr()),
                ctypes.c_void_p(self._bufs[2].ptr) if device_sync else None,
                ctypes.c_void_p(self._peer_up[2]) if device_sync else None,
                ctypes.c_void_p(self._peer_dn[2]) if device_sync else None,
                self._stream()), "carle_band_step")
            if not device_sync:
                self._barrier()              # neighbours' edge rows have landed in my halos
            self._cur = nxt
            done += t

    def close(self):
        for ptr in self._opened:
            self._lib.carle_ipc_close(ctypes.c_void_p(ptr))
        self._opened = []
        self._tensors = []
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)       # nobody still stores into a buffer that is freed
        for b in self._bufs:
            b.free()
        if self._handle is not None:
            self._lib.carle_destroy(self._handle)
            self._handle = None
