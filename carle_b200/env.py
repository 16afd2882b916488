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
* the master reset fires when every action element equals 1.0 exactly; the
  reference tests ``mean(action) == 1.0`` (env.py:208), identical for 0/1
  actions.
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
        # callers assign env.birth / env.survive directly (train_mcl.py:56-57)
        key = (_rule_mask(self.birth), _rule_mask(self.survive))
        if key != self._rule_key:
            _lib.check(self._lib.carle_set_rule(self._handle, key[0], key[1]),
                       "carle_set_rule")
            self._rule_key = key

    # ------------------------------------------------------------------ handle --
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.my_device).cuda_stream)

    def _ensure_handle(self):
        key = (int(self.instances), int(self.height), int(self.width),
               self._ctor_action, self.my_device.index)
        if self._handle is not None and key == self._handle_key:
            return
        self._free_handle()
        handle = ctypes.c_void_p()
        _lib.check(self._lib.carle_create(
            ctypes.byref(handle), self.my_device.index, key[0], key[1], key[2],
            self._ctor_action[0], self._ctor_action[1]), "carle_create")
        self._handle, self._handle_key, self._rule_key = handle, key, None
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
        if getattr(self, "_handle", None) is not None:
            self._lib.carle_destroy(self._handle)
            self._handle = None

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
        self._packed = torch.zeros(shape, dtype=torch.int32, device=self.my_device)
        self._spare = torch.empty_like(self._packed)
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
        """env.py:188-242: toggle, master reset if every toggle is 1.0, else one
        generation; returns ``(obs, reward, done, info)`` like the reference."""
        if self._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        self._sync_rule()
        self.action = action
        if self.logging:
            self.log_universe()
        red = self._red_buf if self.fused_reductions else None
        red_ptr = red.data_ptr() if red is not None else None
        if isinstance(action, RandomAction):
            # the random agent fused into the step kernel: toggles drawn in-kernel (Philox)
            self._absorb_view()
            self._last_action = action
            _lib.check(self._lib.carle_step_random(
                self._handle, self._packed.data_ptr(), self._spare.data_ptr(),
                int(action.seed) & (2**64 - 1), int(action.step) & 0xFFFFFFFF,
                float(action.toggle_rate), action.batch, self._action_buf.data_ptr(),
                self._counters.data_ptr(), red_ptr, self._stream()), "carle_step_random")
        elif isinstance(action, PackedAction):
            # device-generated, already packed: flags from the packed words, then the step
            self._absorb_view()
            self._last_action = action
            words, batch = action.words, action.batch
            _lib.check(self._lib.carle_pack_action(
                self._handle, words.data_ptr(), _lib.PACKED, batch, 1, None,
                self._flags.data_ptr(), self._stream()), "carle_pack_action")
            _lib.check(self._lib.carle_step(
                self._handle, self._packed.data_ptr(), self._spare.data_ptr(),
                words.data_ptr(), batch, self._flags.data_ptr(), self._counters.data_ptr(),
                red_ptr, self._stream()), "carle_step")
        else:
            act = self._coerce_action(action)
            self._absorb_view()
            self._last_action = act
            code = _lib.U8 if act.dtype == torch.uint8 else _lib.F32
            _lib.check(self._lib.carle_step_action(
                self._handle, self._packed.data_ptr(), self._spare.data_ptr(), act.data_ptr(),
                code, act.shape[0], self._counters.data_ptr(), red_ptr, self._stream()),
                "carle_step_action")
        self._packed, self._spare = self._spare, self._packed
        self.last_reductions = red
        observation = self._observation()
        reward = torch.zeros(self.instances, 1, device=self.my_device)     # env.py:238
        done = torch.zeros(self.instances, 1)                              # env.py:239 (CPU)
        info = [{}] * self.instances                                       # env.py:240
        return observation, reward, done, info

    def _speed_tail(self, red, com, have_prev, velocity, speed, reward):
        """SpeedDetector tail on the device (carle_speed_tail); tensors as in the C header."""
        _lib.check(self._lib.carle_speed_tail(
            self._handle, red.data_ptr(), com.data_ptr(), 1 if have_prev else 0,
            velocity.data_ptr(), speed.data_ptr(), reward.data_ptr(), self._stream()),
            "carle_speed_tail")

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
            packed = torch.empty((steps, batch, max(self._aw, 1), self._awpr),
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

    # ------------------------------------------------ I/O helpers (side layer) ---
    from .rle import (render, rle_to_grid, read_rle, read_csv, load_universe,  # noqa: E402
                      get_rle, log_universe, save_log, save_rle, save_frame)
