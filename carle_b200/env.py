"""Drop-in ``CARLE`` environment on the B200-native packed-bit kernels.

Mirrors the public surface of the reference ``carle/env.py:15-242`` (same ctor
kwargs, ``reset`` / ``step`` / ``apply_action`` / ``get_observation`` /
``rules_from_string`` and the attributes wrappers read or write), but the grid is
held bit-packed on the GPU and advanced by ``libcarle_b200.so`` (C ABI in
``include/carle_b200.h``).  There is no CPU path and no torch-op fallback.

Differences from the reference, all deliberate and documented in DESIGN.md:

* device: only CUDA devices exist here.  ``CARLE()`` (reference default: CPU)
  runs on the current CUDA device; ``device="cpu"`` raises; any ``"cuda:N"`` is
  accepted (the reference whitelists cuda:0..3 only, env.py:31-33).
* the master reset follows the reference's ``mean(action) == 1.0`` (env.py:208): the kernels
  decide it from the ballot masks for 0/1 actions and evaluate the mean itself (float64 sum)
  when some element is neither 0 nor 1.
* ``obs_mode`` (extra kwarg): ``"float32"`` (default, strict drop-in: a fresh
  float32 ``[N,1,H,W]`` tensor each step that aliases ``env.universe``),
  ``"uint8"``, or ``"packed"`` (int32 ``[N,H,ceil(W/32)]``, no unpack pass).
* host bookkeeping (``step_number``, ``steps_since_action``) lives in a device
  counter block and is read lazily, so ``step`` never synchronises.
"""
from __future__ import annotations

import ctypes
import time

import torch
import torch.nn as nn

from . import _lib

_ALLOWED = [str(d) for d in range(9)]          # reference env.py:57


def _rule_mask(values):
    mask = 0
    for v in values:
        v = int(v)
        if 0 <= v <= 8:
            mask |= 1 << v
    return mask


_raw_stream = torch._C._cuda_getCurrentRawStream      # (device index) -> cudaStream_t as int
_parked_handles = []        # handles dropped during a stream capture (CARLE._free_handle)


class PackedAction:
    """A toggle action already in the library's packed layout: int32 ``[B, AW, AWPR]`` on the
    env's device (see include/carle_b200.h).  Produced by ``agents.DeviceRandomAgent`` /
    ``CARLE.random_action``; accepted by ``CARLE.step`` in place of the float tensor, which
    removes the dominant byte stream of a step (4 bytes per toggle -> 1 bit)."""

    def __init__(self, words, env):
        self.words, self._env = words, env

    @property
    def batch(self):
        return self.words.shape[0]

    def to_float(self):
        """float32 ``[B, 1, action_width, action_height]`` (the reference's format)."""
        env = self._env
        out = torch.empty((self.batch, 1, env.action_width, env.action_height),
                          dtype=torch.float32, device=env.my_device)
        _lib.check(env._lib.carle_unpack_action(env._handle, self.words.data_ptr(), self.batch,
                                                out.data_ptr(), env._stream()),
                   "carle_unpack_action")
        return out


class RandomAction(PackedAction):
    """The device-side random agent's action as a *recipe* ``(seed, step, toggle_rate)``:
    ``CARLE.step`` draws the toggles inside the step kernel (one launch, no action tensor).
    ``.words`` / ``.to_float()`` materialise exactly the same toggles on demand."""

    def __init__(self, env, seed, step, toggle_rate, batch):
        self._env, self.seed, self.step, self.toggle_rate = env, seed, step, toggle_rate
        self._batch, self._words = batch, None

    @property
    def batch(self):
        return self._batch

    @property
    def words(self):
        if self._words is None:
            self._words = self._env.random_action(self.seed, self.step, self.toggle_rate,
                                                  self._batch).words
        return self._words


class StagedAction:
    """A host action on its way to the device (``CARLE.stage_action``): the copy was enqueued on
    the environment's copy stream into one of its rotating device buffers; ``CARLE.step`` makes
    the compute stream wait for ``ready`` and releases the buffer behind the step.  ``payload``
    is what ``step`` consumes: a float32 / uint8 tensor or a ``PackedAction``."""

    __slots__ = ("payload", "ready", "slot")

    def __init__(self, payload, ready, slot):
        self.payload, self.ready, self.slot = payload, ready, slot


class CARLE(nn.Module):
    """Batched Life-like cellular-automaton environment (reference: carle/env.py:15)."""

    def __init__(self, **kwargs):
        super().__init__()
        self._lib = _lib.load()                 # raises if the CUDA library is missing
        self.inner_env = None                   # env.py:20 (wrappers test this)
        self.width = kwargs.get("width", 256)
        self.height = kwargs.get("height", 256)
        self.action_width = kwargs.get("action_width", 64)
        self.action_height = kwargs.get("action_height", 64)
        self.use_cuda = kwargs.get("use_cuda", False)
        self.my_device = self._resolve_device(kwargs)
        self.use_grad = kwargs.get("use_grad", False)       # unused, as upstream
        self.alive_rate = kwargs.get("alive_rate", 0.0)     # unused, as upstream
        self.instances = kwargs.get("instances", 1)
        self.logging = kwargs.get("logging", False)
        self.obs_mode = kwargs.get("obs_mode", "float32")
        if self.obs_mode not in ("float32", "uint8", "packed"):
            raise ValueError("obs_mode must be 'float32', 'uint8' or 'packed'")
        self.fused_reductions = bool(kwargs.get("fused_reductions", False))
        # actions handed over in HOST memory are bit-packed on the host (library thread pool) and cross
        # the bus as 1 bit per toggle; `host_pack_threads` (default: this process's share of the cores)
        self.host_pack = bool(kwargs.get("host_pack", True))
        self.host_pack_threads = int(kwargs.get("host_pack_threads", 0))
        # True / False / "auto": ship each slice of packed words as soon as it is packed (see _pack_on_host)
        self.host_pack_copy = kwargs.get("host_pack_copy", "auto")

        self._ctor_action = (self.action_height, self.action_width)
        self.set_neighborhood()
        self.set_action_padding()

        self.allowed_rules = list(_ALLOWED)
        self.birth = [3]                        # env.py:58-59, Conway's Life
        self.survive = [2, 3]

        self._handle = None
        self._handle_key = None
        self._rule_key = None
        self._packed = None                     # authoritative state, int32 [N,H,WPR]
        self._spare = None                      # ping-pong partner
        self._view = None                       # float32/uint8 [N,1,H,W] handed to callers
        self._view_version = -1
        self._view_stale = True                 # view does not reflect _packed
        self._counters = None
        self.last_reductions = None
        self._last_action = None
        self._rule_lists = None                 # (tuple(birth), tuple(survive)) last sent to the library
        self._args = _lib.StepArgs()            # argument block of carle_step_ex, reused every step
        self._args.struct_size = ctypes.sizeof(_lib.StepArgs)
        self._args_slow = None                  # the rarely changing fields last written into _args
        self._hp = None                         # host-pack staging (pinned buffers, events)
        self._args_sd = None
        self._done = None                       # the step's constant outputs (env.py:239-240)
        self._info = None
        self._stage = None                      # host-action staging (stage_action)
        self.defer_reset = False                # one shard of a larger batch (sharding.ShardedCARLE)
        self._speed_args = None                 # one-shot: SpeedDetector's tail rides in the next step

    # ------------------------------------------------------------------ set-up --
    def _resolve_device(self, kwargs):
        """env.py:28-37, restricted to CUDA (there is no CPU path here)."""
        if not torch.cuda.is_available():
            raise _lib.CarleLibraryError(
                "carle_b200 needs a CUDA device (sm_100a); it has no CPU fallback")
        if "device" in kwargs:
            name = str(kwargs["device"])
            if name == "cpu":
                raise RuntimeError("carle_b200.CARLE has no CPU path; use device='cuda'")
            if name == "cuda" or (name.startswith("cuda:") and name[5:].isdigit()):
                dev = torch.device(name)
            else:
                raise AttributeError(f"unknown device string {name!r} (the reference "
                                     "leaves my_device unset for it)")
        else:
            dev = torch.device("cuda")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        return torch.device("cuda", index)

    def set_neighborhood(self):
        """Kept for ``state_dict`` compatibility only (checkpoints of wrapped envs hold
        ``inner_env.neighborhood.weight``, evaluation/eval.py:47-48).  The kernels count
        neighbours with bit-sliced adders, not with this convolution (env.py:87-116)."""
        self.neighborhood = nn.Conv2d(1, 1, 3, padding=1, padding_mode="circular",
                                      bias=False)
        with torch.no_grad():
            self.neighborhood.weight.copy_(torch.tensor(
                [[[[1., 1., 1.], [1., 0., 1.], [1., 1., 1.]]]]))
        self.neighborhood.weight.requires_grad_(False)
        self.to(self.my_device)

    def set_action_padding(self):
        """env.py:119-132, same arithmetic (including the axis swap in ZeroPad2d)."""
        asym_w = (self.width - self.action_width) % 2
        asym_h = (self.height - self.action_height) % 2
        self.action_width -= (self.width % 2)
        self.action_height -= (self.height % 2)
        wpad = (self.width - self.action_width) // 2
        hpad = (self.height - self.action_height) // 2
        self.action_padding = nn.ZeroPad2d(
            padding=(hpad, hpad + asym_h, wpad, wpad + asym_w))

    # ------------------------------------------------------------------- rules --
    def birth_rule_from_string(self, my_string="B3"):
        """env.py:62-69."""
        self.birth = sorted({int(ch) for ch in my_string if ch in self.allowed_rules})

    def survive_rule_from_string(self, my_string="S23"):
        """env.py:71-78."""
        self.survive = sorted({int(ch) for ch in my_string if ch in self.allowed_rules})

    def rules_from_string(self, my_string="B3/S23"):
        """env.py:80-85 (IndexError without a '/', as upstream)."""
        parts = my_string.split("/")
        self.birth_rule_from_string(parts[0])
        self.survive_rule_from_string(parts[1])

    def _sync_rule(self):
        # callers assign env.birth / env.survive directly (train_mcl.py:56-57), so the lists are
        # looked at on every step; the masks are re-derived only when they changed
        seen = self._rule_lists
        if seen is not None and self.birth == seen[0] and self.survive == seen[1] \
                and self._rule_key is not None:
            return
        key = (_rule_mask(self.birth), _rule_mask(self.survive))
        if key != self._rule_key:
            _lib.check(self._lib.carle_set_rule(self._handle, key[0], key[1]),
                       "carle_set_rule")
            self._rule_key = key
        self._rule_lists = (list(self.birth), list(self.survive))

    # ------------------------------------------------------------------ handle --
    def _stream(self):
        return ctypes.c_void_p(_raw_stream(self.my_device.index))

    def _ensure_handle(self):
        key = (int(self.instances), int(self.height), int(self.width),
               self._ctor_action, self.my_device.index)
        if self._handle is not None and key == self._handle_key:
            return
        self._free_handle()
        while _parked_handles and not torch.cuda.is_current_stream_capturing():
            lib, parked = _parked_handles.pop()
            lib.carle_destroy(parked)
        handle = ctypes.c_void_p()
        _lib.check(self._lib.carle_create(
            ctypes.byref(handle), self.my_device.index, key[0], key[1], key[2],
            self._ctor_action[0], self._ctor_action[1]), "carle_create")
        self._handle, self._handle_key, self._rule_key = handle, key, None
        self._done = self._info = self._stage = None
        geo = (ctypes.c_int32 * 8)()
        _lib.check(self._lib.carle_geometry(handle, ctypes.byref(geo)))
        (self.row0, self.col0, self._aw, self._ah, self._wpr, self._awpr,
         self.kernel_family, self._aw0) = list(geo)
        dev = self.my_device
        n = key[0]
        self._action_buf = torch.zeros((n, max(self._aw, 1), self._awpr),
                                       dtype=torch.int32, device=dev)
        self._flags = torch.zeros(2, dtype=torch.int32, device=dev)
        self._counters = torch.zeros(8, dtype=torch.int64, device=dev)
        self._red_buf = torch.zeros((n, 4), dtype=torch.int64, device=dev)

    def _free_handle(self):
        handle = getattr(self, "_handle", None)
        if handle is None:
            return
        self._handle = None
        # carle_destroy frees device memory, which CUDA forbids while a stream is being captured (it
        # would invalidate the capture): an environment the garbage collector drops in the middle of
        # somebody's graph capture is parked and destroyed on the next call outside a capture
        if torch.cuda.is_current_stream_capturing():
            _parked_handles.append((self._lib, handle))
            return
        self._lib.carle_destroy(handle)

    def __del__(self):
        try:
            self._free_handle()
        except Exception:
            pass

    # ----------------------------------------------------------- state views ----
    def _unpack(self, dtype):
        out = torch.empty((self.instances, 1, self.height, self.width),
                          dtype=dtype, device=self.my_device)
        code = _lib.U8 if dtype == torch.uint8 else _lib.F32
        _lib.check(self._lib.carle_unpack_state(
            self._handle, self._packed.data_ptr(), out.data_ptr(), code,
            self._stream()), "carle_unpack_state")
        return out

    def _materialize_view(self):
        """Unpack the packed state into a fresh float32 [N,1,H,W] tensor and remember it
        (with its version counter) so in-place edits by callers can be detected."""
        view = self._unpack(torch.float32)
        self._view, self._view_version, self._view_stale = view, view._version, False
        return view

    def _absorb_view(self):
        """If the caller wrote into the tensor we handed out (obs/universe alias the
        state upstream: env.py:184-186, 406; mcl.py:190), re-pack it."""
        view = self._view
        if view is None or self._view_stale:
            return
        if view._version == self._view_version:
            return
        self._pack_from(view)
        self._view_version = view._version

    def _pack_from(self, cells):
        cells = cells.detach()
        if cells.device != self.my_device:
            cells = cells.to(self.my_device)
        if cells.dtype == torch.bool:
            cells = cells.to(torch.uint8)
        elif cells.dtype not in (torch.float32, torch.uint8):
            cells = cells.to(torch.float32)
        cells = cells.contiguous()
        expect = self.instances * self.height * self.width
        if cells.numel() != expect:
            raise RuntimeError(f"universe has {cells.numel()} cells, expected "
                               f"{self.instances}x1x{self.height}x{self.width}")
        code = _lib.U8 if cells.dtype == torch.uint8 else _lib.F32
        _lib.check(self._lib.carle_pack_state(
            self._handle, cells.data_ptr(), code, self._packed.data_ptr(),
            self._stream()), "carle_pack_state")
        self._keepalive = cells

    @property
    def universe(self):
        """float32 ``[N,1,H,W]`` view of the state (reference attribute, env.py:136).
        Callers may edit it in place; the edit is picked up at the next step."""
        if self._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        self._absorb_view()
        if self._view is None or self._view_stale:
            self._materialize_view()
        return self._view

    @universe.setter
    def universe(self, value):
        if self._packed is None:
            self.reset()
        if not torch.is_tensor(value):
            value = torch.as_tensor(value)
        self._pack_from(value)
        self._view, self._view_stale = None, True

    def instance_cells(self, index=0):
        """uint8 numpy ``[H, W]`` copy of ONE universe, decoded on the host from its packed
        words (H*ceil(W/32)*4 bytes cross the bus, not the whole float batch) — what the
        logging / RLE / frame helpers use."""
        import numpy as np
        self._absorb_view()
        words = self._packed[index].cpu().numpy().view(np.uint32)
        bits = np.unpackbits(words.view(np.uint8), axis=-1, bitorder="little")
        return bits.reshape(self.height, -1)[:, :self.width]

    @property
    def packed_universe(self):
        """int32 ``[N, H, ceil(W/32)]`` packed state (bit b of word w = column 32w+b)."""
        self._absorb_view()
        return self._packed

    # host bookkeeping lives on the device; reading it synchronises, stepping does not
    @property
    def step_number(self):
        return int(self._counters[_lib.CNT_STEP_NUMBER].item()) if self._counters is not None else 0

    @step_number.setter
    def step_number(self, value):
        if self._counters is not None:
            self._counters[_lib.CNT_STEP_NUMBER] = int(value)

    @property
    def steps_since_action(self):
        return int(self._counters[_lib.CNT_STEPS_SINCE_ACTION].item()) \
            if self._counters is not None else 0

    @steps_since_action.setter
    def steps_since_action(self, value):
        if self._counters is not None:
            self._counters[_lib.CNT_STEPS_SINCE_ACTION] = int(value)

    # ------------------------------------------------------------------- reset --
    def reset(self):
        """env.py:134-148 — all-dead universe; rules are not reset."""
        self._ensure_handle()
        shape = (int(self.instances), int(self.height), self._wpr)
        if self._packed is not None and tuple(self._packed.shape) == shape \
                and self._packed.device == self.my_device:
            self._packed.zero_()            # same buffers: captured rollout plans stay valid
        else:
            self._packed = torch.zeros(shape, dtype=torch.int32, device=self.my_device)
            self._spare = torch.empty_like(self._packed)
            self.__dict__.pop("_plans", None)
        self._counters.zero_()
        self.instance_id = str(int(time.time()))
        self.log = []
        return self._observation()

    def _observation(self):
        self._view, self._view_stale = None, True
        if self.obs_mode == "packed":
            return self._packed
        if self.obs_mode == "uint8":
            return self._unpack(torch.uint8)
        return self._materialize_view()

    def get_observation(self):
        """env.py:184-186."""
        if self.obs_mode == "packed":
            return self.packed_universe
        return self.universe

    # ------------------------------------------------------------------ action --
    def _coerce_action(self, action):
        """env.py:152-177: to tensor, 4-D, on device, optional centre crop, asserts."""
        if type(action) is torch.Tensor and action.dim() == 4 and action.device == self.my_device \
                and action.dtype in (torch.float32, torch.uint8) and action.is_contiguous() \
                and not action.requires_grad:
            shape = action.shape               # the common case: nothing to convert
            if shape[2] == self.action_width and shape[3] == self.action_height and shape[1] == 1 \
                    and shape[0] in (1, self.instances):
                return action
        if not torch.is_tensor(action):
            action = torch.Tensor(action)
        while action.dim() < 4:
            action = action.unsqueeze(0)
        if action.device != self.my_device:
            action = action.to(self.my_device, non_blocking=True)
        if action.shape[3] > self.action_width and action.shape[1] < self.width:
            off_y = (self.width - self.action_width) // 2
            off_x = (self.height - self.action_height) // 2
            action = action[:, :, off_y:-off_y, off_x:-off_x]
        assert action.shape[2] == self.action_width, \
            f"action width is wrong {action.shape[2]} not {self.action_width}, {action.shape}"
        assert action.shape[3] == self.action_height, \
            f"action height is wrong {action.shape[3]} not {self.action_height}, {action.shape}"
        if action.shape[1] != 1 or action.shape[0] not in (1, self.instances):
            raise RuntimeError(
                f"action batch {tuple(action.shape[:2])} does not broadcast to "
                f"({self.instances}, 1)")
        action = action.detach()
        if action.dtype == torch.bool:
            action = action.to(torch.uint8)
        elif action.dtype not in (torch.float32, torch.uint8):
            action = action.to(torch.float32)
        return action.contiguous()

    def _pack_action(self, action, steps=1, out=None, flags=None):
        code = _lib.U8 if action.dtype == torch.uint8 else _lib.F32
        batch = action.shape[-4]
        if self._aw == 0 or self._ah == 0:
            return batch                 # zero-sized window: nothing to toggle, flags stay 0
        out = self._action_buf if out is None else out
        flags = self._flags if flags is None else flags
        _lib.check(self._lib.carle_pack_action(
            self._handle, action.data_ptr(), code, batch, steps, out.data_ptr(),
            flags.data_ptr(), self._stream()), "carle_pack_action")
        return batch

    def apply_action(self, action):
        """env.py:150-182 — toggle the window cells, no generation."""
        if self._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        action = self._coerce_action(action)
        self._absorb_view()
        batch = self._pack_action(action)
        _lib.check(self._lib.carle_apply_action(
            self._handle, self._packed.data_ptr(), self._action_buf.data_ptr(), batch,
            self._stream()), "carle_apply_action")
        self._flags.zero_()              # not consumed by a step: re-arm for the next pack
        self._view, self._view_stale = None, True

    # -------------------------------------------------------------------- step --
    def step(self, action):
        """env.py:188-242: toggle, master reset if ``mean(action) == 1.0``, else one
        generation; returns ``(obs, reward, done, info)`` like the reference.

        ONE kernel for the batched shapes: the action is ingested, the generation advanced, the
        fresh all-zero ``reward`` written and -- in the float32 / uint8 observation modes -- the
        observation materialised from the registers that hold the new rows (``carle_step_ex``).
        Nothing here synchronises with the device.  ``done`` (CPU zeros) and ``info`` are
        constants upstream (env.py:239-240) and are handed out as the same two objects every
        step."""
        if self._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        self._sync_rule()
        # (plain attributes are written through the instance dict: nn.Module.__setattr__ costs ~2 us
        #  per assignment -- isinstance checks against Parameter / Module / buffers -- and a step makes
        #  eight of them, more than the launch itself)
        d = self.__dict__
        d["action"] = action
        if self.logging:
            self.log_universe()
        if self._view is not None and not self._view_stale:
            self._absorb_view()
        if self.host_pack and type(action) is torch.Tensor and action.device.type == "cpu":
            action = self._pack_on_host(action)
        if isinstance(action, StagedAction):
            # the copy stream has the action in flight: order the step behind it
            torch.cuda.current_stream(self.my_device).wait_event(action.ready)
            staged, action = action, action.payload
        else:
            staged = None
        dev, n = self.my_device, self.instances
        red = self._red_buf if self.fused_reductions else None
        if isinstance(action, RandomAction):
            # the random agent fused into the step kernel: toggles drawn in-kernel (Philox)
            d["_last_action"] = action
            _lib.check(self._lib.carle_step_random(
                self._handle, self._packed.data_ptr(), self._spare.data_ptr(),
                int(action.seed) & (2**64 - 1), int(action.step) & 0xFFFFFFFF,
                float(action.toggle_rate), action.batch, self._action_buf.data_ptr(),
                self._counters.data_ptr(), red.data_ptr() if red is not None else None,
                self._stream()), "carle_step_random")
            d["_packed"], d["_spare"] = self._spare, self._packed
            d["last_reductions"] = red
            observation = self._observation()
            reward = torch.zeros(n, 1, device=dev)                             # env.py:238
        else:
            args = self._args
            if isinstance(action, PackedAction):
                # already in the library's packed layout (device-generated or packed on the host)
                d["_last_action"] = action
                args.action, args.action_dtype = action.words.data_ptr(), _lib.PACKED
                args.action_batch = action.batch
            else:
                act = self._coerce_action(action)
                d["_last_action"] = act
                args.action = act.data_ptr()
                args.action_dtype = _lib.U8 if act.dtype == torch.uint8 else _lib.F32
                args.action_batch = act.shape[0]
            if self._aw == 0 or self._ah == 0:
                args.action = None                      # empty window: nothing to toggle
            args_state_in = self._packed
            args.state_in, args.state_out = args_state_in.data_ptr(), self._spare.data_ptr()
            reward = torch.empty((n, 1), dtype=torch.float32, device=dev)      # zero-filled in-kernel
            args.reward_zero = reward.data_ptr()
            mode = self.obs_mode
            if mode == "packed":
                view = None
            else:
                view = torch.empty((n, 1, self.height, self.width), device=dev,
                                   dtype=torch.float32 if mode == "float32" else torch.uint8)
                args.obs = view.data_ptr()
            sd = self._speed_args
            # the fields that rarely change between two steps are written only when they do (a ctypes
            # structure field costs ~0.12 us per assignment)
            slow = (red is not None, mode, self.defer_reset, sd is not None, self._counters.data_ptr())
            if slow != self._args_slow:
                d["_args_slow"] = slow
                args.counters = self._counters.data_ptr()
                args.reductions = red.data_ptr() if red is not None else None
                if mode == "packed":
                    args.obs = None
                else:
                    args.obs_dtype = _lib.F32 if mode == "float32" else _lib.U8
                args.defer_reset = 1 if self.defer_reset else 0
                if sd is None:
                    args.speed_com_next = None
            if sd is not None:
                # mcl.SpeedDetector wrapped directly around this env: its tail is part of the step
                d["_speed_args"] = None
                args.speed_com_prev, args.speed_com_next = sd[0].data_ptr(), sd[1].data_ptr()
                if sd[2] is not self._args_sd:
                    d["_args_sd"] = sd[2]
                    args.speed_velocity, args.speed_out = sd[2].data_ptr(), sd[3].data_ptr()
                    args.speed_primed, args.speed_sumsq = sd[4].data_ptr(), None
            rc = self._lib.carle_step_ex(self._handle, ctypes.byref(args), self._stream())
            if rc:
                _lib.check(rc, "carle_step_ex")
            packed = d["_packed"] = self._spare
            d["_spare"] = args_state_in
            d["last_reductions"] = red
            if mode == "packed":
                d["_view"], d["_view_stale"] = None, True
                observation = packed
            elif mode == "float32":
                d["_view"], d["_view_version"], d["_view_stale"] = view, view._version, False
                observation = view
            else:
                d["_view"], d["_view_stale"] = None, True
                observation = view
        if staged is not None:
            self._stage["free"][staged.slot].record(torch.cuda.current_stream(dev))
        if self.logging and int(self._counters[_lib.CNT_LAST_NOT_ALL_ONES].item()) == 0:
            # the step was a master reset: upstream's reset() also starts a new log (env.py:142-148)
            self.instance_id = str(int(time.time()))
            self.log = []
        if self._done is None or self._done.shape[0] != n:
            self._done = torch.zeros(n, 1)                                     # env.py:239 (CPU)
            self._info = [{}] * n                                              # env.py:240
        return observation, reward, self._done, self._info

    # ---------------------------------------------------- host actions, staged copies ---
    def pack_host_action(self, action, out=None):
        """Host-side bit-packing of an action ``[B, 1, aw, ah]`` (any dtype, non-zero = toggle) into
        the library's grid-aligned packed layout: int32 ``[B, aw, AWPR]``, pinned.  4 bytes per
        toggle -> 1 bit before the action crosses the bus (8 MiB instead of 256 MiB per step at
        16384 x 256 x 256).  Feed the result to ``stage_action`` / ``PackedAction``.  Note that a
        packed action can only say toggle / no toggle: the master reset then fires when every
        toggle is set."""
        import numpy as np
        self._ensure_handle()
        a = action.detach().cpu().numpy() if torch.is_tensor(action) else np.asarray(action)
        a = a.reshape(-1, self._aw, self._ah) != 0
        bit0 = self.col0 - 32 * self._aw0
        bits = np.zeros((a.shape[0], self._aw, 32 * self._awpr), dtype=bool)
        bits[:, :, bit0:bit0 + self._ah] = a
        words = np.packbits(bits, axis=-1, bitorder="little").view("<u4").view(np.int32)
        if out is None:
            out = torch.empty(words.shape, dtype=torch.int32).pin_memory()
        out.numpy()[...] = words
        return out

    def _pack_on_host(self, action):
        """A float32 / uint8 action in HOST memory -> ``PackedAction`` on the device: packed by the
        library's host threads into a pinned buffer (``carle_pack_action_host``), then one small
        asynchronous copy -- 1 bit per toggle crosses the bus instead of 4 bytes (env.py:158-160
        ships the floats).  Returns the tensor itself when it cannot be packed without changing the
        step's meaning: wrong shape (the usual path raises the reference's errors), dtypes other than
        float32 / uint8 / bool, or a float element that is neither 0 nor 1 (the reference's
        ``mean(action) == 1.0`` is then evaluated on the device from the floats)."""
        if action.dim() != 4 or action.shape[1] != 1 or action.shape[2] != self.action_width \
                or action.shape[3] != self.action_height or action.shape[0] not in (1, self.instances) \
                or self._aw == 0 or self._ah == 0 or action.requires_grad:
            return action
        if action.dtype == torch.bool:
            action = action.view(torch.uint8)
        elif action.dtype not in (torch.float32, torch.uint8):
            return action
        if not action.is_contiguous():
            action = action.contiguous()
        hp = self._hp
        batch = action.shape[0]
        if hp is None or hp["batch"] != batch:
            import os
            threads = self.host_pack_threads
            if threads <= 0:
                cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
                threads = max(1, min(32, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
                # a host thread packs ~5 GB/s of float32 entry by entry and ~1.7x that on the flat
                # multi-stream walk of a word-aligned window (host_pack.cu; measured, memory bound);
                # below ~50 GB/s in total the plain DMA of the floats is faster: leave the action as it is
                flat = (self.col0 - 32 * self._aw0) == 0 and self._ah % 32 == 0
                if threads < (6 if flat else 10):
                    self.host_pack = False
                    return action
            shape = (batch, self._aw, self._awpr)
            hp = self.__dict__["_hp"] = {
                "batch": batch, "threads": threads, "next": 0,
                "host": [torch.empty(shape, dtype=torch.int32).pin_memory() for _ in range(2)],
                "dev": [torch.empty(shape, dtype=torch.int32, device=self.my_device) for _ in range(2)],
                "copied": [torch.cuda.Event() for _ in range(2)],
                "flags": (ctypes.c_int32 * 3)()}
        k = hp["next"] & 1
        hp["next"] += 1
        hp["copied"][k].synchronize()            # (the copy that last read this pinned buffer is done)
        host, flags = hp["host"][k], hp["flags"]
        dev = hp["dev"][k]
        stream = torch.cuda.current_stream(self.my_device)
        sliced = self.host_pack_copy
        if sliced == "auto":
            # a caller that waits for every step (the reference drivers' reward.cpu()) finds the stream
            # idle: the copy then hides behind the packing.  With steps still queued the device is not
            # what the caller waits for, and the words go over in one piece behind the packing
            sliced = stream.query()
        if sliced:
            # packed and shipped by one call: each slice of the packed words is on its way to the
            # device as soon as the host threads are through with it
            rc = self._lib.carle_pack_action_host_copy(
                self._aw, self._ah, self._awpr, self.col0 - 32 * self._aw0, action.data_ptr(),
                _lib.U8 if action.dtype == torch.uint8 else _lib.F32, batch, host.data_ptr(), flags,
                hp["threads"], dev.data_ptr(), dev.device.index, self._stream())
            if rc:
                _lib.check(rc, "carle_pack_action_host_copy")
            hp["copied"][k].record(stream)
            if flags[2]:
                return action                    # neither 0 nor 1 somewhere: the floats decide on the device
            return PackedAction(dev, self)
        rc = self._lib.carle_pack_action_host(
            self._aw, self._ah, self._awpr, self.col0 - 32 * self._aw0, action.data_ptr(),
            _lib.U8 if action.dtype == torch.uint8 else _lib.F32, batch, host.data_ptr(), flags,
            hp["threads"])
        if rc:
            _lib.check(rc, "carle_pack_action_host")
        if flags[2]:
            return action
        dev.copy_(host, non_blocking=True)
        hp["copied"][k].record(stream)
        return PackedAction(dev, self)

    def stage_action(self, host_action, slots=2):
        """Enqueue the host->device copy of ``host_action`` (pinned memory for a truly asynchronous
        copy) on the environment's copy stream and return a ``StagedAction`` for ``step``: while
        step t runs, the copy of action t+1 is in flight.  ``host_action``: float32 / uint8
        ``[B, 1, aw, ah]`` or packed int32 ``[B, aw, AWPR]`` (``pack_host_action``).  ``slots``
        rotating device buffers per (shape, dtype); a slot is reused only after the step that
        consumed it has been enqueued and has run."""
        self._ensure_handle()
        dev = self.my_device
        if self._stage is None:
            self._stage = {"stream": torch.cuda.Stream(device=dev), "bufs": {}, "free": [],
                           "next": 0}
        st = self._stage
        key = (tuple(host_action.shape), host_action.dtype)
        if key not in st["bufs"]:
            base = len(st["free"])
            st["bufs"][key] = [(torch.empty(key[0], dtype=key[1], device=dev), base + i)
                               for i in range(slots)]
            st["free"].extend(torch.cuda.Event() for _ in range(slots))
            for ev in st["free"][base:]:
                ev.record(torch.cuda.current_stream(dev))
        ring = st["bufs"][key]
        buf, slot = ring[st["next"] % len(ring)]
        st["next"] += 1
        copy_stream = st["stream"]
        copy_stream.wait_event(st["free"][slot])           # the step that read this slot is done
        with torch.cuda.stream(copy_stream):
            buf.copy_(host_action, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        packed = host_action.dtype == torch.int32 and host_action.dim() == 3
        return StagedAction(PackedAction(buf, self) if packed else buf, ready, slot)

    def _speed_tail(self, red, com, have_prev, velocity, speed, reward, sumsq=None, primed=None):
        """SpeedDetector tail on the device (carle_speed_tail); tensors as in the C header."""
        rc = self._lib.carle_speed_tail(
            self._handle, red.data_ptr(), com.data_ptr(), 1 if have_prev else 0,
            velocity.data_ptr(), speed.data_ptr(),
            reward.data_ptr() if reward is not None else None,
            sumsq.data_ptr() if sumsq is not None else None,
            primed.data_ptr() if primed is not None else None, self._stream())
        if rc:
            _lib.check(rc, "carle_speed_tail")

    def step_many(self, actions, reductions=False):
        """K generations in one launch: ``actions`` is ``[K, B, 1, aw, ah]`` (B = 1 or N)
        or ``None``/int K for a free run.  Equivalent to K calls of :meth:`step`; the
        warp-resident kernels keep the universes in registers for all K generations.
        Returns ``(obs, per_step_reductions or None)``."""
        if self._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        self._sync_rule()
        self._absorb_view()
        dev = self.my_device
        if actions is None or isinstance(actions, int):
            steps, packed, flags, batch = int(actions or 1), None, None, 1
        else:
            if not torch.is_tensor(actions) or actions.dim() != 5:
                raise ValueError("actions must be a tensor [K, B, 1, aw, ah]")
            steps, batch = actions.shape[0], actions.shape[1]
            assert actions.shape[3] == self.action_width, "action width is wrong"
            assert actions.shape[4] == self.action_height, "action height is wrong"
            if actions.shape[2] != 1 or batch not in (1, self.instances):
                raise RuntimeError("action batch must be 1 or N, with one channel")
            flat = actions.detach()
            if flat.device != dev:
                flat = flat.to(dev, non_blocking=True)
            if flat.dtype == torch.bool:
                flat = flat.to(torch.uint8)
            elif flat.dtype not in (torch.float32, torch.uint8):
                flat = flat.to(torch.float32)
            flat = flat.contiguous()
            if self._aw == 0 or self._ah == 0:
                # empty window: nothing to toggle and no reset (the mean of an empty tensor is NaN
                # upstream); zeroed flags would read as "every toggle is 1.0"
                packed = flags = None
            else:
                packed = torch.empty((steps, batch, self._aw, self._awpr),
                                     dtype=torch.int32, device=dev)
                flags = torch.zeros((steps, 2), dtype=torch.int32, device=dev)
                self._pack_action(flat, steps=steps, out=packed, flags=flags)
        red = torch.empty((steps, self.instances, 4), dtype=torch.int64, device=dev) \
            if reductions else None
        scratch = torch.empty_like(self._packed) if (self.kernel_family != 1 and steps > 1) \
            else None
        _lib.check(self._lib.carle_step_many(
            self._handle, self._packed.data_ptr(), self._spare.data_ptr(),
            scratch.data_ptr() if scratch is not None else None,
            packed.data_ptr() if packed is not None else None, batch, steps,
            flags.data_ptr() if flags is not None else None,
            self._counters.data_ptr(), red.data_ptr() if red is not None else None,
            self._stream()), "carle_step_many")
        self._packed, self._spare = self._spare, self._packed
        return self._observation(), red

    # ------------------------------------------------------------- reductions ---
    def reduce(self):
        """Per-instance ``[live, sum i*m*u, sum j*m*u, live inside window]`` (int64
        ``[N,4]``) of the current state — the SpeedDetector sums, mcl.py:773-779."""
        self._absorb_view()
        out = torch.empty((self.instances, 4), dtype=torch.int64, device=self.my_device)
        _lib.check(self._lib.carle_reduce(self._handle, self._packed.data_ptr(),
                                          out.data_ptr(), self._stream()), "carle_reduce")
        return out

    def masked_count(self, plus_mask=None, minus_mask=None):
        """``popcount(u & plus) - popcount(u & minus)`` per instance (int64 ``[N]``);
        masks are packed int32 ``[H, WPR]`` (``carle_b200.mcl.pack_mask``)."""
        self._absorb_view()
        out = torch.empty(self.instances, dtype=torch.int64, device=self.my_device)
        _lib.check(self._lib.carle_masked_count(
            self._handle, self._packed.data_ptr(),
            plus_mask.data_ptr() if plus_mask is not None else None,
            minus_mask.data_ptr() if minus_mask is not None else None,
            out.data_ptr(), self._stream()), "carle_masked_count")
        return out

    def action_count(self):
        """Toggles per action entry of the LAST step (int64 ``[B]``), mcl.py:102-103."""
        act = self._last_action
        if act is None:
            raise RuntimeError("action_count() needs a previous step()")
        if isinstance(act, PackedAction):
            words, batch = act.words, act.batch
        else:
            batch = self._pack_action(act)
            self._flags.zero_()          # not consumed by a step: re-arm for the next pack
            words = self._action_buf
        out = torch.empty(batch, dtype=torch.int64, device=self.my_device)
        _lib.check(self._lib.carle_action_count(
            self._handle, words.data_ptr(), batch, out.data_ptr(), self._stream()),
            "carle_action_count")
        return out

    def random_action(self, seed, step, toggle_rate=0.1, batch=None):
        """Bernoulli(toggle_rate) toggles for every window cell, generated on the device in
        the packed layout (stateless Philox keyed by seed/step; carle/agents.py:35-42)."""
        self._ensure_handle()
        batch = self.instances if batch is None else batch
        words = torch.empty((batch, max(self._aw, 1), self._awpr), dtype=torch.int32,
                            device=self.my_device)
        _lib.check(self._lib.carle_random_action(
            self._handle, int(seed) & (2**64 - 1), int(step) & 0xFFFFFFFF, float(toggle_rate),
            batch, words.data_ptr(), self._stream()), "carle_random_action")
        return PackedAction(words, self)

    # device-side state a rollout plan saves around its warm-up steps (rollout.RolloutPlan)
    def _snapshot(self):
        self._absorb_view()
        return (self._packed.clone(), self._counters.clone())

    def _restore(self, state):
        self._packed.copy_(state[0])
        self._counters.copy_(state[1])
        self._view, self._view_stale = None, True

    def rollout(self, actions):
        """K = ``actions.shape[0]`` environment steps as ONE CUDA-graph replay (``rollout.RolloutPlan``,
        cached per action shape / dtype): ``actions`` (device tensor ``[K, B, 1, aw, ah]`` or packed
        int32 ``[K, B, aw, AWPR]``) is copied into the plan's static buffer.  Returns ``(obs,
        rewards [K, N, 1])``."""
        from .rollout import RolloutPlan
        plans = self.__dict__.setdefault("_plans", {})
        key = (tuple(actions.shape), actions.dtype, tuple(self.birth), tuple(self.survive))
        plan = plans.get(key)
        if plan is None:
            plan = plans[key] = RolloutPlan(self, actions.clone())
        else:
            plan.actions.copy_(actions)
        obs, rewards = plan.run()
        return obs, torch.stack(rewards)

    # ------------------------------------------------ I/O helpers (side layer) ---
    from .rle import (render, rle_to_grid, rle_to_packed, rle_body, read_rle, read_csv,  # noqa: E402
                      load_universe, get_rle, log_universe, save_log, save_rle, save_frame)
