"""Rollouts that stay on the device (the reference's drivers,
carle/train_mcl.py:52-69 and evaluation/eval.py:56-72, pay a ``.cpu()`` -- a full device
synchronisation -- per environment step).

* ``RolloutPlan(env, actions)`` captures ``K = actions.shape[0]`` calls of ``env.step`` -- the
  bare ``CARLE`` or any stack of the ``carle_b200.mcl`` wrappers, whose ``step`` never
  synchronises -- into ONE CUDA graph.  ``plan.run()`` replays it: one host call for K
  environment steps, K reward tensors produced on the device.  ``plan.actions`` is the graph's
  static input: copy the next K actions into it and run again.
* ``host_rollout(env, host_actions)`` drives ``env.step`` from actions that live in (pinned) HOST
  memory with the copy of action t+1 in flight while step t runs (``CARLE.stage_action``) and
  the rewards read back asynchronously into pinned memory -- one synchronisation per rollout
  instead of one per step.
* ``train_loop`` is the shape of the reference's ``train()`` on top of those two.
"""
import gc

import torch

from .env import PackedAction


def _inner(env):
    return env.inner_env if getattr(env, "inner_env", None) is not None else env


class RolloutPlan:
    """K environment steps as one CUDA graph.

    ``actions``: device tensor ``[K, B, 1, aw, ah]`` (float32 / uint8; B = 1 or N) or packed
    int32 ``[K, B, aw, AWPR]``; it becomes the plan's static input buffer (``plan.actions``).
    The environment's state is advanced by the capture's warm-up and then restored, so building
    a plan leaves ``env`` where it was."""

    def __init__(self, env, actions, warmup=1):
        self.env, self.inner = env, _inner(env)
        inner = self.inner
        if inner._packed is None:
            raise AttributeError("reset() the environment before planning a rollout")
        if not torch.is_tensor(actions) or actions.device != inner.my_device:
            raise ValueError("actions must be a tensor on the environment's device")
        self.actions = actions.contiguous()
        self.steps = int(self.actions.shape[0])
        self.packed = self.actions.dtype == torch.int32 and self.actions.dim() == 4
        if not self.packed and self.actions.dim() != 5:
            raise ValueError("actions must be [K, B, 1, aw, ah] or packed int32 [K, B, aw, AWPR]")
        dev = inner.my_device
        saved = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # lazy set-up (kernel attributes, NVRTC rules) outside the capture
            for k in range(min(warmup, self.steps)):
                env.step(self._action(k))
        torch.cuda.current_stream(dev).wait_stream(side)
        self._restore(saved)
        start = inner._packed
        self.graph = torch.cuda.CUDAGraph()
        self.rewards = []
        # (no cyclic garbage collection inside the capture: a finaliser that frees device memory
        #  -- any tensor, any environment -- would invalidate it)
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            self._capture(env, inner, saved, start)
        finally:
            if gc_was_on:
                gc.enable()
        # (the capture recorded the launches without running them: the device state is still
        #  `saved`; whatever float view the last captured step handed out holds nothing yet)
        inner._view, inner._view_stale = None, True

    def _capture(self, env, inner, saved, start):
        with torch.cuda.graph(self.graph):
            obs = None
            tokens = [(e, e._plan_begin()) for e, _ in saved if hasattr(e, "_plan_begin")]
            for k in range(self.steps):
                obs, reward, _, _ = env.step(self._action(k))
                self.rewards.append(reward)
            if inner._packed is not start:
                # odd K: the ping-pong ended in the other buffer; land where a replay starts
                start.copy_(inner._packed)
                inner._packed, inner._spare = inner._spare, inner._packed
                if inner.obs_mode == "packed":
                    obs = inner._packed
            for e, token in tokens:
                e._plan_end(token)
            self.obs = obs

    def _action(self, k):
        return PackedAction(self.actions[k], self.inner) if self.packed else self.actions[k]

    def _snapshot(self):
        env, chain = self.env, []
        while env is not None:
            chain.append((env, env._snapshot()))
            env = getattr(env, "env", None) if getattr(env, "inner_env", None) is not None else None
        return chain

    @staticmethod
    def _restore(chain):
        for env, state in chain:
            env._restore(state)

    def run(self):
        """Replay the K steps.  Returns ``(obs, rewards)``: the observation after the last step and
        the list of the K per-step reward tensors -- the same tensors on every replay (they live
        in the graph's memory pool), so copy what has to survive the next ``run``."""
        inner = self.inner
        if inner._view is not None and not inner._view_stale:
            inner._absorb_view()                 # the caller edited env.universe in place
        self.graph.replay()
        if inner.obs_mode == "float32":
            inner._view, inner._view_version, inner._view_stale = self.obs, self.obs._version, False
        else:
            inner._view, inner._view_stale = None, True
        return self.obs, self.rewards


def host_rollout(env, host_actions, rewards_out=None, sync_every=None):
    """Step ``env`` through ``host_actions`` (a sequence of pinned host tensors: float32 / uint8
    ``[B, 1, aw, ah]`` or packed int32 ``[B, aw, AWPR]`` from ``CARLE.pack_host_action``).  The copy
    of action t+1 overlaps step t; every step's reward is copied into ``rewards_out[t]`` (pinned
    float32 ``[K, N, 1]``, allocated if None) without blocking.  ``sync_every``: synchronise with
    the device every that many steps (None: once, at the end) -- ``sync_every=1`` is the
    reference drivers' behaviour.  Returns ``(obs, rewards_out)``."""
    inner = _inner(env)
    dev = inner.my_device
    steps = len(host_actions)
    obs = None
    # unpacked float32 / uint8 actions: with host-side packing (CARLE(host_pack=True), the default)
    # step() packs action t+1 on the host threads while the device runs step t and copies 1 bit per
    # toggle; otherwise the tensors are staged as they are on the copy stream
    def direct(a):
        return inner.host_pack and a.dim() == 4 and a.dtype in (torch.float32, torch.uint8, torch.bool)
    nxt = host_actions[0] if direct(host_actions[0]) else inner.stage_action(host_actions[0])
    for t in range(steps):
        cur = nxt
        if t + 1 < steps:
            a = host_actions[t + 1]
            nxt = a if direct(a) else inner.stage_action(a)
        obs, reward, _, _ = env.step(cur)
        if rewards_out is None:
            rewards_out = torch.empty((steps,) + tuple(reward.shape), dtype=reward.dtype).pin_memory()
        rewards_out[t].copy_(reward, non_blocking=True)
        if sync_every and (t + 1) % sync_every == 0:
            torch.cuda.current_stream(dev).synchronize()
    torch.cuda.current_stream(dev).synchronize()
    return obs, rewards_out


def train_loop(env, agent, max_steps, block=64, on_block=None):
    """The reference's ``train()`` loop (carle/train_mcl.py:52-69: ``action = agent(obs)``,
    ``obs, reward, done, info = env.step(action)``, ``reward.cpu()`` every step) without the
    per-step synchronisation: rewards accumulate on the device and are read once per ``block``
    steps.  ``agent(obs)`` may return a float tensor or a ``PackedAction`` (``DeviceRandomAgent``).
    ``on_block(step, rewards[block, N, 1] on the host)`` is the consumer hook; returns the sum of
    all rewards like the reference's running total."""
    obs = env.reset()
    total, pending = 0.0, []
    for t in range(max_steps):
        action = agent(obs)
        obs, reward, _, _ = env.step(action)
        pending.append(reward)
        if len(pending) == block or t + 1 == max_steps:
            host = torch.stack(pending).cpu()            # the only synchronisation of the block
            total += float(host.sum())
            if on_block is not None:
                on_block(t + 1, host)
            pending = []
    return total
