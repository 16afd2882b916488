"""Reward wrappers whose grid reductions run on the packed state, device side.

Same classes, constructor signature and ``reset`` / ``step`` protocol as the reference
``carle/mcl.py`` (``Motivator`` :29-84, ``ParsimonyBonus`` :86-105, ``MorphoBonus`` :107-195,
``CornerBonus`` :197-231, ``SpeedDetector`` :730-799, ``PufferDetector`` :804-853), so they stack the
same way (``env = SpeedDetector(CARLE(...))``; ``env.inner_env`` is the ``CARLE``).
Where the reference makes extra full-grid float32 passes with torch
(``sum(obs * weight)``), these read the per-instance integer sums the step kernel
already produced (``CARLE.last_reductions``) or launch one popcount kernel over the
bit-packed grid.  The neural curiosity wrappers (RND2D, AE2D, ...) are out of scope;
the reference's own versions keep working on top of ``carle_b200.CARLE`` because they
only consume the float32 observation.
"""
import numpy as np
import torch
import torch.nn as nn

from .env import RandomAction as _RandomAction


def pack_mask(mask_2d, device):
    """[H, W] 0/1 mask -> packed int32 [H, ceil(W/32)] in the library's state layout
    (bit b of word w = column 32*w + b).  Set-up time helper, host side."""
    m = (np.asarray(torch.as_tensor(mask_2d).detach().cpu()) != 0).astype(np.uint8)
    h, w = m.shape
    wpr = (w + 31) // 32
    padded = np.zeros((h, wpr * 32), dtype=np.uint8)
    padded[:, :w] = m
    words = np.packbits(padded, axis=-1, bitorder="little").view("<u4").reshape(h, wpr)
    return torch.from_numpy(words.astype(np.int64).astype(np.uint32).view(np.int32)
                            .copy()).to(device)


class Motivator(nn.Module):
    """Wrapper base (reference carle/mcl.py:29-84)."""

    def __init__(self, env, **kwargs):
        super().__init__()
        self.inner_env = env if env.inner_env is None else env.inner_env
        self.env = env
        self.height = self.inner_env.height
        self.width = self.inner_env.height          # sic, mcl.py:42
        self.action_height = self.inner_env.action_height
        self.action_width = self.inner_env.action_width
        self.birth = self.inner_env.birth
        self.survive = self.inner_env.survive
        self.my_device = self.inner_env.my_device

    def rules_from_string(self, my_string="B3/S23"):
        self.inner_env.rules_from_string(my_string)
        self.birth, self.survive = self.inner_env.birth, self.inner_env.survive

    def birth_rule_from_string(self, my_string="b3"):
        self.inner_env.birth_rule_from_string(my_string)
        self.birth = self.inner_env.birth

    def survive_rule_from_string(self, my_string="s23"):
        self.inner_env.survive_rule_from_string(my_string)
        self.survive = self.inner_env.survive

    def reset(self):
        return self.env.reset()

    def step(self, action):
        return self.env.step(action)

    def set_no_grad(self):
        pass

    def set_grad(self):
        pass

    def rollout(self, actions):
        """K wrapped steps as one CUDA-graph replay (see ``CARLE.rollout``); returns ``(obs,
        rewards [K, ...])`` with this wrapper's rewards."""
        from .rollout import RolloutPlan
        plans = self.__dict__.setdefault("_plans", {})
        key = (tuple(actions.shape), actions.dtype, tuple(self.inner_env.birth),
               tuple(self.inner_env.survive))
        plan = plans.get(key)
        if plan is None:
            plan = plans[key] = RolloutPlan(self, actions.clone())
        else:
            plan.actions.copy_(actions)
        obs, rewards = plan.run()
        return obs, torch.stack(rewards)

    # device-side state a rollout plan saves around its warm-up steps (rollout.RolloutPlan);
    # wrappers with state of their own extend these
    def _snapshot(self):
        return None

    def _restore(self, state):
        pass

    # called inside a rollout plan's capture, before the first and after the last captured step
    def _plan_begin(self):
        return None

    def _plan_end(self, token):
        pass


class ParsimonyBonus(Motivator):
    """reward <- 100 * reward / max(#toggles, 100)   (reference mcl.py:86-105).

    The toggle count is a popcount of the packed action the step already produced —
    equal to the reference's ``action.sum(axis=[1,2,3])`` for 0/1-valued actions.  The
    reference's (N,1)/(N,) -> (N,N) broadcast for N > 1 is kept."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        self.parsimony_threshold = 128

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        toggles = self.inner_env.action_count().to(torch.float32)
        floor = torch.tensor([100.0], device=toggles.device)
        reward = 100.0 * reward / torch.max(toggles, floor)
        return obs, reward, done, info


class CornerBonus(Motivator):
    """Fixed-mask reward/punish sums (reference mcl.py:197-231) as one masked popcount."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        self.my_name = "CornerBonus"
        self.reward_scale = 1.0
        h, w = self.inner_env.height, self.inner_env.width
        self.reward_mask = torch.zeros(1, 1, h, w)
        self.punish_mask = torch.zeros(1, 1, h, w)
        self.reward_mask[:, :, :16, :16] = 1.0
        for ii in range(96):                                    # mcl.py:213-214
            self.reward_mask[:, :, ii - 4:ii + 4, ii - 4:ii + 4] = 1.0
        self.punish_mask[:, :, -64:, -64:] = -1.0
        self.punish_mask[:, :, :64, -64:] = -1.0
        self._plus = pack_mask(self.reward_mask[0, 0], self.my_device)
        self._minus = pack_mask(self.punish_mask[0, 0], self.my_device)
        self.reward_mask = self.reward_mask.to(self.my_device)
        self.punish_mask = self.punish_mask.to(self.my_device)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        bonus = self.inner_env.masked_count(self._plus, self._minus)      # int64 [N]
        reward += self.reward_scale * bonus.to(torch.float32).unsqueeze(1)
        return obs, reward, done, info


class SpeedDetector(Motivator):
    """Centre-of-mass motion bonus (reference mcl.py:730-799).

    live = sum u, Sh = sum i*m*u, Sw = sum j*m*u (m = 0 inside the action window) are exact
    integers computed in the step kernel's epilogue while the new rows are still in registers.
    Wrapped directly around a ``CARLE`` the O(N) tail (two divides, a subtract, one norm,
    reward += speed; mcl.py:777-795) is part of THE SAME kernel: whoever completes an instance's
    sums turns them into its centre of mass and velocity, and the grid's last warp writes the
    batch-wide speed and the reward column (``carle_step_ex``, ``speed_*`` arguments).  Stacked on
    another wrapper, or for action types without a one-launch kernel, the tail is one more small
    launch (``carle_speed_tail``)."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        self.reward_scale = 1.0
        self.speed_modulator = 32.0
        self.growing_steps = 0
        self.smooth_velocity = None
        self.speed = None
        self.velocity = torch.tensor([self.inner_env.instances, 1, 0., 0.]).to(self.my_device)
        self._live_src = None
        self._com = None                # two [2, N] buffers: the step reads one and writes the other
        self._com_cur = 0               # index of the latest
        self._steps_seen = 0
        self.inner_env.fused_reductions = True

    @property
    def center_of_mass(self):
        """float32 ``[2, N]`` (mcl.py:781-785): None until the first step, as upstream."""
        if self._com is None or not self._steps_seen:
            return None
        return self._com[self._com_cur]

    @center_of_mass.setter
    def center_of_mass(self, value):
        if value is None:
            self._steps_seen = 0
            if self._com is not None:
                self._primed.zero_()
            return
        value = torch.as_tensor(value, dtype=torch.float32, device=self.my_device)
        self._speed_buffers(value.shape[1])
        self._com[self._com_cur].copy_(value)
        self._primed.fill_(1)
        self._steps_seen = max(self._steps_seen, 1)

    @property
    def live_cells(self):
        """float32 ``[N]`` live-cell counts of the last step (mcl.py:773), converted from the
        step kernel's integer sums when somebody looks (no extra launch per step)."""
        src = self._live_src
        return None if src is None else src[:, 0].to(torch.float32)

    @live_cells.setter
    def live_cells(self, value):
        self._live_src = None if value is None else \
            torch.as_tensor(value, dtype=torch.float32).reshape(-1, 1)

    def _speed_buffers(self, n):
        """Allocated on the first step.  ``_primed`` is the device-side "a previous centre of mass
        exists" flag (mcl.py:784): the kernels read and set it, so one and the same launch serves
        the first step and all later ones (a captured rollout can be replayed from any state)."""
        if self._com is None or tuple(self._com[0].shape) != (2, n):
            dev = self.my_device
            self._com = [torch.zeros((2, n), dtype=torch.float32, device=dev) for _ in range(2)]
            self._com_cur = 0
            self._velocity_buf = torch.zeros((2, n), dtype=torch.float32, device=dev)
            self._speed_buf = torch.zeros(1, dtype=torch.float32, device=dev)
            self._speed_view = self._speed_buf[0]           # 0-d, what `speed` shows (mcl.py:789)
            self._primed = torch.zeros(1, dtype=torch.int32, device=dev)
            self._steps_seen = 0

    def step(self, action):
        inner = self.inner_env
        d = self.__dict__        # (plain attributes: nn.Module.__setattr__ costs ~2 us per assignment)
        fused = self.env is inner and not inner.defer_reset and not isinstance(action, _RandomAction)
        if fused:
            self._speed_buffers(int(inner.instances))
            cur = self._com_cur
            inner.__dict__["_speed_args"] = (self._com[cur], self._com[cur ^ 1], self._velocity_buf,
                                             self._speed_buf, self._primed)
            obs, reward, done, info = inner.step(action)
            d["_com_cur"] = cur ^ 1
            red = inner.last_reductions
        else:
            obs, reward, done, info = self.env.step(action)
            red = inner.last_reductions
            if red is None:
                red = inner.reduce()
            self._speed_buffers(red.shape[0])
            # one launch: centre of mass, velocity, batch-wide speed, reward += speed (mcl.py:777-795)
            if not reward.is_contiguous():
                reward = reward.contiguous()
            inner._speed_tail(red, self._com[self._com_cur], False, self._velocity_buf,
                              self._speed_buf, reward, primed=self._primed)
        if self._steps_seen:
            d["velocity"] = self._velocity_buf
            d["speed"] = self._speed_view
        d["_steps_seen"] = self._steps_seen + 1
        # (the env's sum buffer: like upstream's attribute it always shows the latest step)
        d["_live_src"] = red
        return obs, reward, done, info

    def _snapshot(self):
        if self._com is None:
            return None
        return (self._com[0].clone(), self._com[1].clone(), self._com_cur, self._velocity_buf.clone(),
                self._speed_buf.clone(), self._primed.clone(), self._steps_seen)

    def _restore(self, state):
        if state is None:
            if self._com is not None:        # allocated by the warm-up: back to "unprimed"
                self._primed.zero_()
                self._steps_seen = 0
            return
        com0, com1, cur, vel, speed, primed, seen = state
        self._com[0].copy_(com0)
        self._com[1].copy_(com1)
        self._com_cur = cur
        self._velocity_buf.copy_(vel)
        self._speed_buf.copy_(speed)
        self._primed.copy_(primed)
        self._steps_seen = seen

    def _plan_begin(self):
        return self._com_cur

    def _plan_end(self, token):
        # an odd number of captured steps leaves the latest centres in the other buffer: land
        # where a replay starts reading
        if self._com is not None and self._com_cur != token:
            self._com[token].copy_(self._com[self._com_cur])
            self._com_cur = token


class PufferDetector(Motivator):
    """Growth bonus (reference mcl.py:804-853): +1 for every instance when the total live-cell
    count over the last ``growth_threshold`` action-free steps grew.  The count is the exact
    integer sum of the step kernel's per-instance popcounts, and the sliding window lives ON THE
    DEVICE (``carle_puffer_tail``: a ring of ``growth_threshold + 1`` totals, appended / emptied
    by the "no toggle this step" counter, the bonus added to ``reward`` by the same launch), so a
    step never synchronises with the host -- upstream reads the total back every step
    (mcl.py:830).  ``cells`` / ``live_cells`` are read from the device when somebody looks."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        self.my_name = "PufferDetector"
        self.reward_scale = 1.0
        self.growth_threshold = 512
        self.growing_steps = 0
        self._ring = None
        self._state = None
        self.inner_env.fused_reductions = True

    def _window(self):
        cap = int(self.growth_threshold) + 1
        if self._ring is None or self._ring.shape[0] != cap:
            dev = self.my_device
            self._ring = torch.zeros(cap, dtype=torch.int64, device=dev)
            self._state = torch.zeros(8, dtype=torch.int64, device=dev)
        return self._ring, self._state

    @property
    def cells(self):
        """The window's totals, oldest first (mcl.py:817, 835-847) -- a host copy."""
        if self._ring is None:
            return []
        state, ring = self._state.cpu().tolist(), self._ring.cpu().tolist()
        count, head = state[0], state[1]
        return [float(ring[(head + k) % len(ring)]) for k in range(count)]

    @cells.setter
    def cells(self, value):
        ring, state = self._window()
        value = [int(v) for v in value][-ring.shape[0]:]
        state[:2] = torch.tensor([len(value), 0], dtype=torch.int64)
        if value:
            ring[:len(value)] = torch.tensor(value, dtype=torch.int64)

    @property
    def live_cells(self):
        return 0.0 if self._state is None else float(self._state[4].item())

    @live_cells.setter
    def live_cells(self, value):
        pass                               # (mcl.py:818 initialises it; the device owns it here)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        inner = self.inner_env
        red = inner.last_reductions
        if red is None:
            red = inner.reduce()
        ring, state = self._window()
        if not reward.is_contiguous():
            reward = reward.contiguous()
        rc = inner._lib.carle_puffer_tail(inner._handle, red.data_ptr(), inner._counters.data_ptr(),
                                          ring.data_ptr(), state.data_ptr(), int(self.growth_threshold),
                                          reward.data_ptr(), inner._stream())
        if rc:
            from . import _lib
            _lib.check(rc, "carle_puffer_tail")
        return obs, reward, done, info

    def _snapshot(self):
        if self._ring is None:
            return None
        return (self._ring.clone(), self._state.clone())

    def _restore(self, state):
        if self._ring is None:
            return
        if state is None:
            self._ring.zero_()
            self._state.zero_()
        else:
            self._ring.copy_(state[0])
            self._state.copy_(state[1])


#: the two glider phases upstream expects in carle/glider_1.rle / glider_2.rle (mcl.py:143-144) but
#: does not ship
GLIDER_PHASES = ("bob$2bo$3o!", "obo$b2o$bo!")


class MorphoBonus(Motivator):
    """Bonus for matching a body plan (reference mcl.py:107-195): ``F.conv2d`` of the toggled
    universe with the +w / -1 templates of the target patterns in six orientations, reward +=
    max + min of the response over templates and positions, per instance.

    Here the response is never materialised: ``carle_morpho_match`` slides the 8x8 templates over
    the PACKED rows (a window is one 64-bit word, a template score two population counts) and
    returns the two extrema.  ``target_patterns`` keeps the reference's float ``[P, 1, 8, 8]``
    form (appending to it by hand works: the packed form is re-derived when it changes).

    ``abs(universe - action)`` (mcl.py:176) needs an action that broadcasts against the universe:
    grid-sized (the env crops it to the window for the step itself, env.py:164-169) or a window
    that is the whole grid.  A window-sized action is taken here as the zero-padded window --
    upstream raises on it.  Actions are read as toggles (non-zero = 1)."""

    def __init__(self, env, **kwargs):
        super().__init__(env, **kwargs)
        self.use_grad = False
        self.my_name = "MorphoBonus"
        self.reward_scale = 1.0
        self.target_patterns = torch.Tensor().to(self.my_device)
        self._packed_for = None
        self.add_default_patterns()

    def add_default_patterns(self):
        for body in GLIDER_PHASES:
            self.add_rle_text(body)

    def add_rle_pattern(self, rle_path, dim=8):
        """mcl.py:146-172 (``read_rle`` also sets the env's rule from the file, as upstream)."""
        self.add_rle_text(self.inner_env.read_rle(rle_path), dim)

    def add_rle_text(self, body, dim=8):
        grid = self.inner_env.rle_to_grid(body)
        pattern = nn.functional.pad(grid, (1, 1, 2, 1))[:dim, :dim].clone()
        pattern[pattern == 0] = -1
        pattern = pattern.unsqueeze(0).unsqueeze(0).to(self.my_device)
        pattern[pattern == 1] *= 15. / pattern[pattern == 1].sum()
        turned = pattern.transpose(2, 3)
        self.target_patterns = torch.cat([self.target_patterns, pattern, pattern.flip(2), pattern.flip(3),
                                          turned.flip(2), turned.flip(3), turned])

    def _templates(self):
        """uint64 bit masks (bit 8r + c = cell [r][c] live) and the live weight per template."""
        pats = self.target_patterns
        key = (pats.data_ptr(), pats._version, tuple(pats.shape))
        if self._packed_for != key:
            host = pats.detach().cpu().numpy()[:, 0]
            if host.shape[1:] != (8, 8) or not 1 <= host.shape[0] <= 64:
                raise ValueError("MorphoBonus: 1..64 templates of 8x8")
            live = host > 0
            if (host[~live] != -1).any():
                raise ValueError("MorphoBonus: dead template cells must weigh -1 (mcl.py:153)")
            weights = np.array([h[m].max() if m.any() else 0.0 for h, m in zip(host, live)], dtype=np.float32)
            if any((h[m] != w).any() for h, m, w in zip(host, live, weights)):
                raise ValueError("MorphoBonus: one weight per template (mcl.py:157)")
            bits = np.packbits(live.reshape(-1, 64), axis=-1, bitorder="little").view("<u8")[:, 0]
            self._words = torch.from_numpy(bits.astype(np.uint64).view(np.int64).copy()).to(self.my_device)
            self._weights = torch.from_numpy(weights).to(self.my_device)
            self._packed_for = key
        return self._words, self._weights

    def match(self, action=None):
        """``(max, min)`` float32 ``[N]`` of the template response on the universe toggled by
        ``action`` (``None``: as it stands)."""
        from . import _lib
        from .env import PackedAction
        inner = self.inner_env
        if inner._packed is None:
            raise AttributeError("universe is undefined before reset() (as upstream)")
        inner._absorb_view()
        n, h, w = int(inner.instances), inner.height, inner.width
        state, toggles, batch = inner._packed, None, 1
        if action is not None:
            if not isinstance(action, PackedAction) and not torch.is_tensor(action):
                action = torch.Tensor(action)
            if not isinstance(action, PackedAction) and tuple(action.shape[-2:]) == (h, w) \
                    and (inner.action_width, inner.action_height) != (h, w):
                # grid-sized: the whole plane toggles (mcl.py:176 subtracts the uncropped action)
                plane = (action.reshape(-1, 1, h, w) != 0).to(device=inner.my_device, dtype=torch.uint8)
                plane = plane.expand(n, 1, h, w).contiguous()
                toggles = torch.empty_like(inner._packed)
                _lib.check(inner._lib.carle_pack_state(inner._handle, plane.data_ptr(), _lib.U8,
                                                       toggles.data_ptr(), inner._stream()), "carle_pack_state")
                batch = n
            else:
                state = inner._packed.clone()
                if isinstance(action, PackedAction):
                    words, b = action.words, action.batch
                else:
                    b = inner._pack_action(inner._coerce_action(action))
                    inner._flags.zero_()
                    words = inner._action_buf
                if inner._aw and inner._ah:
                    _lib.check(inner._lib.carle_apply_action(inner._handle, state.data_ptr(), words.data_ptr(),
                                                             b, inner._stream()), "carle_apply_action")
        words, weights = self._templates()
        mx = torch.empty(n, dtype=torch.float32, device=inner.my_device)
        mn = torch.empty_like(mx)
        _lib.check(inner._lib.carle_morpho_match(
            inner._handle, state.data_ptr(), toggles.data_ptr() if toggles is not None else None, batch,
            words.data_ptr(), weights.data_ptr(), int(words.shape[0]), mx.data_ptr(), mn.data_ptr(),
            inner._stream()), "carle_morpho_match")
        return mx, mn

    def step(self, action):
        mx, mn = self.match(action)                                        # mcl.py:176-177
        obs, reward, done, info = self.env.step(action)
        reward += self.reward_scale * (mx + mn).unsqueeze(-1)              # mcl.py:181-185
        return obs, reward, done, info

    def reset(self):
        """mcl.py:187-195: a sprinkle of seed cells (0.5 %) on the fresh universe."""
        self.env.reset()
        inner = self.inner_env
        seeds = torch.rand(int(inner.instances), 1, inner.height, inner.width) > 0.995
        inner.universe = seeds.to(torch.uint8)
        return inner.get_observation()
