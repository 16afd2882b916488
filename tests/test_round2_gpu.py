"""Round-2 parity tests of the CUDA path (through carle_b200 -> C ABI -> sm_100a kernels):
the reference's `mean(action) == 1.0` reset predicate on non-binary actions, the outputs fused
into the step kernel (float32 / uint8 observation, zero reward), deferred resets of
instance-sharded batches, CUDA-graph rollouts, staged host actions.  Bit-exact everywhere."""
import ctypes

import numpy as np
import pytest
import torch

import _cases as cs
from _golden import by_kind
from oracle import carle_oracle as oc
from test_parity_gpu import ADAPTERS, CudaAdapter, VARIANTS_128, VARIANTS_256

pytestmark = pytest.mark.gpu


def _carle():
    import carle_b200
    return carle_b200


class _Adapter(CudaAdapter):
    @property
    def steps_since_action(self):
        return self.inner.steps_since_action


def _adapter(base):
    return type("A_" + base.__name__, (_Adapter,), {"obs_mode": base.obs_mode})


# ------------------------------------------------ reset predicate (env.py:191, 208) ----
@pytest.mark.parametrize("name", by_kind("master_reset_mean"))
@pytest.mark.parametrize("make", ADAPTERS)
def test_master_reset_on_mean_of_nonbinary_actions(name, make):
    """Golden sequences recorded from the reference: 0/2 and 0.5/1.5 checkerboards (mean 1.0)
    reset, +1/-1 (sum 0) counts as "no action" and toggles every cell, near misses do not reset."""
    cs.check_master_reset_mean(name, _adapter(make))


def _nonbinary_sequence(rng, n, win):
    ii, jj = np.meshgrid(np.arange(win), np.arange(win), indexing="ij")
    checker = ((ii + jj) % 2).astype(np.float32)[None, None].repeat(n, 0)
    rand = lambda: (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)   # noqa: E731
    mixed = np.ones((n, 1, win, win), dtype=np.float32)
    mixed[n // 2:] = 2.0 * checker[n // 2:]
    near = 2.0 * checker
    near[n - 1, 0, win - 1, win - 2] = 0.0
    scaled = rand() * 3.5                                   # toggles with value 3.5: no reset
    one_inst = rand()
    one_inst[n // 3] = 0.5 + checker[0]                     # a single non-binary instance
    return [rand(), 2.0 * checker, rand(), mixed, 2.0 * checker - 1.0, near, scaled, one_inst,
            0.5 + checker, rand()]


@pytest.mark.parametrize("size,win,n,variant",
                         [(256, 64, 9, v) for v in VARIANTS_256] + [(256, 64, 600, VARIANTS_256[1])] +
                         [(128, 32, 21, v) for v in VARIANTS_128] + [(128, 32, 4800, VARIANTS_128[2])] +
                         [(64, 32, 33, {"CARLE_FUSED_IMPL": "tma"}), (64, 32, 33, {"CARLE_FUSED_IMPL": "direct"}),
                          (96, 32, 5, {}), (320, 64, 2, {}), (100, 50, 3, {})])
def test_reset_predicate_in_every_step_kernel(size, win, n, variant, monkeypatch):
    """Every one-launch kernel variant, and the pack + step path of the other geometries, follows
    the reference's predicates on actions with values other than 0 and 1 (oracle = numpy
    restatement, pinned to the reference by the golden sequences above)."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    cb = _carle()
    rng = np.random.default_rng(17 * size + n)
    big = n > 100
    check = np.arange(n) if not big else np.unique(rng.integers(0, n, size=16))
    soup = (rng.random((n, size, size)) < 0.37).astype(np.uint8)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   fused_reductions=True, obs_mode="packed")
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    env.reset()
    ref.reset()
    env.universe = torch.from_numpy(soup).float()[:, None]
    ref.universe = soup.copy()
    for t, a in enumerate(_nonbinary_sequence(rng, n, win)):
        env.step(torch.from_numpy(a))
        want = ref.step(a)[0]
        got = env.universe[:, 0].cpu().numpy().astype(np.uint8)
        assert np.array_equal(got[check], want[check]), t
        assert env.step_number == ref.step_number, t
        assert env.steps_since_action == ref.steps_since_action, t
        red = env.last_reductions.cpu().numpy()
        assert np.array_equal(red[:, 0], got.reshape(n, -1).sum(1)), t


def test_reset_predicate_in_step_many():
    cb = _carle()
    rng = np.random.default_rng(5)
    n, size, win = 6, 64, 32
    seq = _nonbinary_sequence(rng, n, win)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed")
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
    env.reset()
    ref.reset()
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    env.universe = torch.from_numpy(soup).float()[:, None]
    ref.universe = soup.copy()
    env.step_many(torch.from_numpy(np.stack(seq)).cuda())
    for a in seq:
        ref.step(a)
    assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), ref.universe)
    assert env.step_number == ref.step_number
    assert env.steps_since_action == ref.steps_since_action


# ------------------------------------------------------- outputs fused into the step ----
@pytest.mark.parametrize("size,win,n,variant",
                         [(64, 32, 37, {}), (64, 32, 9000, {}), (128, 32, 41, {}), (128, 32, 5000, {}),
                          (128, 32, 41, {"CARLE_FUSED_IMPL": "strip"}), (128, 32, 41, {"CARLE_FUSED_IMPL": "direct"}),
                          (256, 64, 11, {}), (256, 64, 300, {}), (256, 64, 11, {"CARLE_STRIP_R": "2"}),
                          (256, 64, 11, {"CARLE_FUSED_IMPL": "direct"}), (256, 64, 11, {"CARLE_FUSED_IMPL": "tma"}),
                          (96, 32, 4, {}), (320, 64, 2, {}), (100, 50, 3, {})])
@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.uint8])
def test_fused_observation_and_reward(size, win, n, variant, obs_dtype, monkeypatch):
    """carle_step_ex writes the unpacked observation and the zero reward from inside the step
    kernel: identical to unpacking the packed state afterwards, for every kernel family, also
    when the step is a master reset (observation all zero)."""
    from carle_b200 import _lib
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    cb = _carle()
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(size * 7 + n)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed", fused_reductions=True)
    env.rules_from_string("B368/S245")
    env.reset()
    env.universe = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.45).float()
    env._sync_rule()
    code = _lib.F32 if obs_dtype == torch.float32 else _lib.U8
    for t in range(4):
        a = 1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1)
        if t == 2:
            a.fill_(1.0)                                        # master reset
        obs = torch.full((n, 1, size, size), 7, dtype=obs_dtype, device="cuda")
        reward = torch.full((n, 1), 3.0, device="cuda")
        args = _lib.StepArgs()
        args.struct_size = ctypes.sizeof(_lib.StepArgs)
        args.state_in, args.state_out = env._packed.data_ptr(), env._spare.data_ptr()
        args.action, args.action_dtype, args.action_batch = a.data_ptr(), _lib.F32, n
        args.counters, args.reductions = env._counters.data_ptr(), env._red_buf.data_ptr()
        args.reward_zero, args.obs, args.obs_dtype = reward.data_ptr(), obs.data_ptr(), code
        assert lib.carle_step_ex(env._handle, ctypes.byref(args), env._stream()) == 0, _lib.last_error()
        env._packed, env._spare = env._spare, env._packed
        want = torch.empty_like(obs)
        assert lib.carle_unpack_state(env._handle, env._packed.data_ptr(), want.data_ptr(), code,
                                      env._stream()) == 0
        assert torch.equal(obs, want), t
        assert torch.count_nonzero(reward).item() == 0, t
        if t == 2:
            assert torch.count_nonzero(obs).item() == 0
            assert torch.count_nonzero(env._red_buf).item() == 0
        else:
            assert torch.count_nonzero(obs).item() > 0


def test_public_step_outputs_are_fresh_tensors():
    """obs and reward are new tensors every step (callers keep them in replay buffers); done /
    info are the reference's constants."""
    cb = _carle()
    env = cb.CARLE(instances=3, height=64, width=64, action_width=32, action_height=32)
    env.reset()
    a = 1.0 * (torch.rand(3, 1, 32, 32, device="cuda") <= 0.3)
    o1, r1, d1, i1 = env.step(a)
    keep = o1.clone()
    o2, r2, d2, i2 = env.step(a)
    assert o1.data_ptr() != o2.data_ptr() and r1.data_ptr() != r2.data_ptr()
    assert torch.equal(o1, keep)
    assert r1.shape == (3, 1) and float(r1.abs().sum()) == 0.0 and r1.device.type == "cuda"
    assert d1.device.type == "cpu" and d1.shape == (3, 1) and len(i1) == 3 and i1[0] == {}


def test_step_many_with_empty_window_does_not_reset():
    """A zero-sized action window: no toggles and no reset (the mean of an empty tensor is NaN
    upstream), in step() and in step_many()."""
    cb = _carle()
    rng = np.random.default_rng(3)
    soup = (rng.random((2, 64, 64)) < 0.4).astype(np.uint8)
    ref = oc.life_like_update(oc.life_like_update(soup, [3], [2, 3]), [3], [2, 3])
    for many in (False, True):
        env = cb.CARLE(instances=2, height=64, width=64, action_width=0, action_height=0)
        env.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        if many:
            env.step_many(torch.zeros(2, 2, 1, 0, 0, device="cuda"))
        else:
            env.step(torch.zeros(2, 1, 0, 0))
            env.step(torch.zeros(2, 1, 0, 0))
        assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), ref), many
        assert env.step_number == 2


# --------------------------------------------------------- instance-sharded batches ----
@pytest.mark.parametrize("size,win,n", [(64, 32, 10), (128, 32, 12), (256, 64, 6), (96, 32, 4)])
def test_deferred_reset_of_two_shards_equals_the_unsharded_batch(size, win, n):
    """Two shards of one batch on one GPU, stepped with defer_reset and combined by hand the way
    sharding.ShardedCARLE does (AND of the shards' conditions -> carle_apply_reset), against the
    oracle on the whole batch: the reset fires on both shards or on none."""
    cb = _carle()
    from carle_b200 import _lib
    rng = np.random.default_rng(size + n)
    half = n // 2
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
    ref.reset()
    ref.universe = soup.copy()
    shards = []
    for lo, hi in ((0, half), (half, n)):
        e = cb.CARLE(instances=hi - lo, height=size, width=size, action_width=win,
                     action_height=win, obs_mode="float32", fused_reductions=True)
        e.defer_reset = True
        e.reset()
        e.universe = torch.from_numpy(soup[lo:hi]).float()[:, None]
        shards.append((e, lo, hi))
    ones = np.ones((n, 1, win, win), dtype=np.float32)
    part = ones.copy()
    part[half:] = (rng.random((n - half, 1, win, win)) <= 0.1)       # only shard 0 is all ones
    rand = lambda: (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)   # noqa: E731
    for t, a in enumerate([rand(), part, rand(), ones, rand()]):
        outs, conds = [], []
        for e, lo, hi in shards:
            outs.append(e.step(torch.from_numpy(a[lo:hi]))[0])
            conds.append(e._counters[_lib.CNT_LAST_RESET_COND].clone())
        decision = (conds[0] * conds[1]).to(torch.int32).reshape(1)
        for (e, lo, hi), obs in zip(shards, outs):
            _lib.check(e._lib.carle_apply_reset(
                e._handle, decision.data_ptr(), e._packed.data_ptr(), obs.data_ptr(), _lib.F32,
                e.last_reductions.data_ptr(), e._counters.data_ptr(), e._stream()))
        want = ref.step(a)[0]
        got = torch.cat(outs)[:, 0].cpu().numpy().astype(np.uint8)
        assert np.array_equal(got, want), t
        for e, lo, hi in shards:
            assert np.array_equal(e.universe[:, 0].cpu().numpy().astype(np.uint8), want[lo:hi]), t
            assert e.step_number == ref.step_number, t
            assert np.array_equal(e.last_reductions[:, 0].cpu().numpy(),
                                  want[lo:hi].reshape(hi - lo, -1).sum(1)), t


def test_sharded_env_single_rank_equals_oracle():
    """ShardedCARLE / ShardedSpeedDetector with one rank (no process group): same states and
    rewards as the oracle's SpeedDetector, including a master reset."""
    cb = _carle()
    rng = np.random.default_rng(9)
    n, size, win = 5, 128, 32
    env = cb.ShardedSpeedDetector(cb.ShardedCARLE(instances=n, height=size, width=size,
                                                  action_width=win, action_height=win))
    env.rules_from_string("B368/S245")
    ref = oc.OracleSpeedDetector(oc.OracleCARLE(width=size, height=size, action_width=win,
                                                action_height=win, instances=n))
    ref.env.rules_from_string("B368/S245")
    env.reset()
    ref.reset()
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    env.inner_env.universe = torch.from_numpy(soup).float()[:, None]
    ref.env.universe = soup.copy()
    for t in range(7):
        a = (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)
        if t == 4:
            a[:] = 1.0
        obs, reward, _, _ = env.step(torch.from_numpy(a))
        want_obs, want_r, _, _ = ref.step(a)
        assert np.array_equal(obs[:, 0].cpu().numpy().astype(np.uint8), want_obs), t
        np.testing.assert_allclose(reward.cpu().numpy(), np.broadcast_to(want_r, (n, 1)),
                                   rtol=2e-6, atol=1e-6, err_msg=str(t))


# ------------------------------------------------------------- rollouts on the device ----
@pytest.mark.parametrize("steps", [4, 5])
@pytest.mark.parametrize("wrapper", [None, "SpeedDetector", "CornerBonus"])
@pytest.mark.parametrize("obs_mode", ["packed", "float32"])
def test_rollout_plan_equals_eager_steps(steps, wrapper, obs_mode):
    """K steps captured into one CUDA graph (rollout.RolloutPlan) and replayed -- twice, with new
    actions copied into the plan's static buffer -- give the states and rewards of K eager steps;
    building the plan leaves the environment where it was."""
    cb = _carle()
    n, size, win = 7, 128, 32
    g = torch.Generator(device="cuda").manual_seed(steps)

    def make():
        inner = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                         action_height=win, obs_mode=obs_mode)
        env = inner if wrapper is None else getattr(cb, wrapper)(inner)
        inner.rules_from_string("B368/S245")
        env.reset()
        return env, inner

    soup = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.4).float()
    acts = [1.0 * (torch.rand(steps, n, 1, win, win, device="cuda", generator=g) <= 0.1)
            for _ in range(2)]
    eager, e_inner = make()
    e_inner.universe = soup
    want_rewards, want_states = [], []
    for block in acts:
        for k in range(steps):
            want_rewards.append(eager.step(block[k])[1].clone())
        want_states.append(e_inner.packed_universe.clone())
    env, inner = make()
    inner.universe = soup
    before = inner.packed_universe.clone()
    plan = cb.RolloutPlan(env, acts[0].clone())
    assert torch.equal(inner.packed_universe, before)
    got_rewards = []
    for i, block in enumerate(acts):
        plan.actions.copy_(block)
        obs, rewards = plan.run()
        got_rewards += [r.clone() for r in rewards]
        assert torch.equal(inner.packed_universe, want_states[i]), i
        if obs_mode == "float32":
            assert torch.equal(obs, inner.universe)
    for k, (a, b) in enumerate(zip(got_rewards, want_rewards)):
        torch.testing.assert_close(a, b, rtol=2e-6, atol=1e-6, msg=str(k))
    assert inner.step_number == 2 * steps
    # the convenience entry point (plan cached per shape)
    env2, inner2 = make()
    inner2.universe = soup
    _, r = env2.rollout(acts[0])
    assert torch.equal(inner2.packed_universe, want_states[0]) and r.shape[0] == steps


@pytest.mark.parametrize("fmt", ["float32", "uint8", "packed"])
@pytest.mark.parametrize("size,win", [(128, 32), (256, 64), (64, 32)])
def test_host_rollout_with_staged_copies(fmt, size, win):
    """Actions in pinned host memory, copy of action t+1 in flight while step t runs, rewards
    read back asynchronously: same result as feeding device tensors one by one.  Packed host
    actions (1 bit per toggle, packed on the host) give the same states as float32 ones."""
    cb = _carle()
    n, steps = 9, 6
    torch.manual_seed(size)
    soup = (torch.rand(n, 1, size, size) < 0.4).float()
    host = [1.0 * (torch.rand(n, 1, win, win) <= 0.1) for _ in range(steps)]

    def make():
        env = cb.SpeedDetector(cb.CARLE(instances=n, height=size, width=size, action_width=win,
                                        action_height=win, obs_mode="packed"))
        env.reset()
        env.inner_env.universe = soup
        return env

    ref = make()
    want = [ref.step(a.cuda())[1].cpu() for a in host]
    env = make()
    if fmt == "float32":
        feed = [a.pin_memory() for a in host]
    elif fmt == "uint8":
        feed = [a.to(torch.uint8).pin_memory() for a in host]
    else:
        feed = [env.inner_env.pack_host_action(a) for a in host]
    obs, rewards = cb.host_rollout(env, feed)
    assert torch.equal(env.inner_env.packed_universe, ref.inner_env.packed_universe)
    for t in range(steps):
        torch.testing.assert_close(rewards[t], want[t], rtol=2e-6, atol=1e-6)
    assert rewards.is_pinned()


def test_speed_detector_first_step_is_decided_on_the_device():
    """The wrapper's "first step records the centre of mass only" (mcl.py:784) is a device-side
    flag: no reward on the first step, rewards afterwards, and the golden rewards still match."""
    cs.check_wrapper("speed_128", CudaAdapter)
    cb = _carle()
    env = cb.SpeedDetector(cb.CARLE(instances=4, height=64, width=64, action_width=32,
                                    action_height=32))
    env.reset()
    a = 1.0 * (torch.rand(4, 1, 32, 32, device="cuda") <= 0.2)
    r0 = env.step(a)[1]
    assert env.speed is None and float(r0.abs().sum()) == 0.0
    r1 = env.step(a)[1]
    assert env.speed is not None and float(r1.min()) > 0.0
    assert torch.equal(r1, env.speed.expand(4, 1))


@pytest.mark.parametrize("size,win,n,variant",
                         [(64, 32, 40, {}), (64, 32, 7000, {}), (128, 32, 33, {}), (128, 32, 5000, {}),
                          (128, 32, 33, {"CARLE_FUSED_IMPL": "strip"}), (256, 64, 9, {}), (256, 64, 500, {}),
                          (256, 64, 9, {"CARLE_STRIP_R": "2"}), (256, 64, 9, {"CARLE_FUSED_IMPL": "direct"}),
                          (256, 64, 9, {"CARLE_FUSED_IMPL": "quad"}), (96, 32, 4, {}), (100, 50, 3, {})])
def test_speed_detector_tail_inside_the_step_kernel(size, win, n, variant, monkeypatch):
    """SpeedDetector wrapped directly around CARLE: centre of mass, velocity, batch-wide speed and
    the reward column come out of the step kernel itself (or of carle_speed_tail behind kernels
    that cannot carry them): exact centres of mass, rewards within the float32 norm's summation
    order, through a master reset (velocities relative to the centres BEFORE the reset) and a
    near miss, float32 and uint8 actions."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    cb = _carle()
    rng = np.random.default_rng(size * 3 + n)
    env = cb.SpeedDetector(cb.CARLE(instances=n, height=size, width=size, action_width=win,
                                    action_height=win, obs_mode="packed"))
    env.rules_from_string("B368/S245")
    ref = oc.OracleSpeedDetector(oc.OracleCARLE(width=size, height=size, action_width=win,
                                                action_height=win, instances=n))
    ref.env.rules_from_string("B368/S245")
    env.reset()
    ref.reset()
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    env.inner_env.universe = torch.from_numpy(soup).float()[:, None]
    ref.env.universe = soup.copy()
    assert env.center_of_mass is None
    for t in range(8):
        a = (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)
        if t == 3:
            a[:] = 1.0                                       # master reset
        if t == 5:
            a[:] = 1.0
            a[n // 2, 0, win - 1, 0] = 0.0                   # near miss
        ta = torch.from_numpy(a)
        obs, reward, _, _ = env.step(ta.to(torch.uint8) if t % 2 else ta)
        _, want_r, _, _ = ref.step(a)
        got = env.inner_env.universe[:, 0].cpu().numpy().astype(np.uint8)
        assert np.array_equal(got, ref.env.universe), t
        assert np.array_equal(env.center_of_mass.cpu().numpy(), ref.center_of_mass), t
        np.testing.assert_allclose(reward.cpu().numpy(), np.broadcast_to(want_r, (n, 1)),
                                   rtol=2e-6, atol=1e-6, err_msg=str(t))
        assert np.array_equal(env.live_cells.cpu().numpy(), ref.live_cells.astype(np.float32)), t
        if t:
            assert abs(float(env.speed) - float(ref.speed)) <= 2e-6 * float(ref.speed) + 1e-6


@pytest.mark.parametrize("size,win,n,variant",
                         [(64, 32, 50, {}), (64, 32, 9000, {}), (128, 32, 45, {}), (128, 32, 5000, {}),
                          (128, 32, 45, {"CARLE_STRIP128": "1"}), (256, 64, 10, {}), (256, 64, 600, {}),
                          (256, 64, 10, {"CARLE_STRIP_R": "2"}), (256, 64, 10, {"CARLE_FUSED_IMPL": "tma"}),
                          (96, 32, 4, {}), (320, 64, 2, {})])
def test_packed_actions_in_the_step_kernel(size, win, n, variant, monkeypatch):
    """Actions handed over already bit-packed (PackedAction: 1 bit per toggle, grid-aligned words)
    are ingested by the one-launch kernels directly: same states, sums and rewards as the oracle
    fed the float actions, through a master reset (every valid bit set), a near miss and a batch-1
    broadcast action; run-time rules included."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    cb = _carle()
    rng = np.random.default_rng(size * 5 + n)
    big = n > 100
    check = np.arange(n) if not big else np.unique(rng.integers(0, n, size=16))
    for rule in ("B368/S245", "B2/S0123"):
        env = cb.SpeedDetector(cb.CARLE(instances=n, height=size, width=size, action_width=win,
                                        action_height=win, obs_mode="packed"))
        env.rules_from_string(rule)
        inner = env.inner_env
        ref = oc.OracleSpeedDetector(oc.OracleCARLE(width=size, height=size, action_width=win,
                                                    action_height=win, instances=n))
        ref.env.rules_from_string(rule)
        env.reset()
        ref.reset()
        soup = (rng.random((n, size, size)) < 0.35).astype(np.uint8)
        inner.universe = torch.from_numpy(soup).float()[:, None]
        ref.env.universe = soup.copy()
        for t in range(6):
            batch = 1 if t == 1 else n
            a = (rng.random((batch, 1, win, win)) <= 0.12).astype(np.float32)
            if t == 2:
                a[:] = 1.0                                   # master reset
            if t == 4:
                a[:] = 1.0
                a[n - 1, 0, win // 2, win - 1] = 0.0         # near miss
            words = inner.pack_host_action(torch.from_numpy(a)).cuda()
            obs, reward, _, _ = env.step(cb.PackedAction(words, inner))
            _, want_r, _, _ = ref.step(a)
            got = inner.universe[:, 0].cpu().numpy().astype(np.uint8)
            assert np.array_equal(got[check], ref.env.universe[check]), (rule, t)
            assert inner.step_number == ref.env.step_number, (rule, t)
            np.testing.assert_allclose(reward.cpu().numpy(), np.broadcast_to(want_r, (n, 1)),
                                       rtol=2e-6, atol=1e-6, err_msg=f"{rule} {t}")
            assert np.array_equal(inner.action_count().cpu().numpy(),
                                  (a != 0).reshape(batch, -1).sum(1)), (rule, t)


# ------------------------------------------- host actions packed on the host (env.py:158-160) ----
@pytest.mark.parametrize("walk", ["flat", "entrywise"])
@pytest.mark.parametrize("win,bit0,n,slices", [(64, 0, 200, 4), (64, 0, 9, 8), (32, 16, 300, 3), (30, 3, 50, 0)])
def test_host_packing_ships_its_slices(walk, win, bit0, n, slices, monkeypatch):
    """carle_pack_action_host_copy: the packed words reach the device in slices, each enqueued by the
    host thread that finished it -- the device buffer must hold exactly numpy's packing, whatever the
    number of slices, the walk (flat multi-stream / entry by entry), the dtype and the thread count."""
    from carle_b200 import _lib
    lib = _lib.load()
    if walk == "entrywise":
        monkeypatch.setenv("CARLE_HOST_PACK_FLAT", "0")
    if slices:
        monkeypatch.setenv("CARLE_HOST_PACK_SLICES", str(slices))
    rng = np.random.default_rng(win + n)
    awpr = (bit0 + win + 31) // 32
    stream = torch.cuda.Stream()
    for dtype, threads in ((np.float32, 5), (np.uint8, 2), (np.float32, 1)):
        a = torch.from_numpy((rng.random((n, win, win)) < 0.1).astype(dtype)).pin_memory()
        host = torch.full((n, win, awpr), -1, dtype=torch.int32).pin_memory()
        dev = torch.full((n, win, awpr), -1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        flags = (ctypes.c_int32 * 3)()
        rc = lib.carle_pack_action_host_copy(
            win, win, awpr, bit0, a.data_ptr(), _lib.U8 if dtype == np.uint8 else _lib.F32, n,
            host.data_ptr(), flags, threads, dev.data_ptr(), dev.device.index,
            ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0, _lib.last_error()
        stream.synchronize()
        bits = np.zeros((n, win, awpr * 32), dtype=np.uint8)
        bits[:, :, bit0:bit0 + win] = a.numpy() != 0
        want = np.packbits(bits, axis=-1, bitorder="little").view("<u4").reshape(n, win, awpr)
        assert np.array_equal(host.numpy().view(np.uint32), want)
        assert np.array_equal(dev.cpu().numpy().view(np.uint32), want), (walk, dtype, threads)
        assert (flags[0], flags[1], flags[2]) == (1, 1, 0)
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("size,win,n", [(256, 64, 40), (128, 32, 33), (64, 32, 70), (100, 30, 5)])
def test_host_actions_are_packed_on_the_host(size, win, n):
    """An action tensor in HOST memory is bit-packed by the library's host threads and crosses the
    bus as packed words (carle_pack_action_host); the step must be the one the same tensor on the
    device gives -- states, counters, master reset on all ones, batch-1 broadcast, uint8 / bool --
    and a non-binary float action (the 0/2 checkerboard whose mean is 1.0) must still reset."""
    cb = _carle()
    rng = np.random.default_rng(size + n)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)

    def make(host_pack):
        env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                       device="cuda", obs_mode="packed", host_pack=host_pack, host_pack_threads=3)
        env.rules_from_string("B36/S23")
        env.reset()
        env.universe = torch.from_numpy(soup)[:, None]
        return env
    host, dev = make(True), make(False)
    ii, jj = np.meshgrid(np.arange(win), np.arange(win), indexing="ij")
    checker = (2.0 * ((ii + jj) % 2)).astype(np.float32)[None, None]
    seq = [(rng.random((n, 1, win, win)) <= 0.1).astype(np.float32),
           np.zeros((n, 1, win, win), dtype=np.float32),
           (rng.random((1, 1, win, win)) <= 0.3).astype(np.float32),              # batch-1 broadcast
           (rng.random((n, 1, win, win)) <= 0.1).astype(np.uint8),
           (rng.random((n, 1, win, win)) <= 0.1),                                 # bool
           np.ones((n, 1, win, win), dtype=np.float32),                           # master reset
           (rng.random((n, 1, win, win)) <= 0.5).astype(np.float32),
           np.broadcast_to(checker, (n, 1, win, win)).copy(),                     # mean == 1.0, not binary
           (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)]
    for t, a in enumerate(seq):
        a = torch.from_numpy(np.ascontiguousarray(a))
        o_host = host.step(a)[0]
        o_dev = dev.step(a.cuda())[0]
        assert torch.equal(o_host, o_dev), t
        assert host.step_number == dev.step_number and host.steps_since_action == dev.steps_since_action, t
        if t in (5, 7):
            assert int(o_host.abs().sum().item()) == 0 and host.step_number == 0
    # what crossed the bus was packed (except for the non-binary action)
    assert isinstance(host._last_action, cb.PackedAction)
    assert host._hp is not None and host._hp["host"][0].is_pinned()
