"""Parity of the CUDA path (through carle_b200.CARLE -> C ABI -> sm_100a kernels) with
the numpy oracle and with the golden vectors recorded from the reference.  Bit-exact
for every grid; float tolerance only where stated (SpeedDetector's final norm)."""
import numpy as np
import pytest
import torch

import _cases as cs
from _golden import by_kind
from oracle import carle_oracle as oc

pytestmark = pytest.mark.gpu


def _carle():
    import carle_b200
    return carle_b200


class CudaAdapter:
    """Adapter over carle_b200.CARLE (float32 drop-in mode) for tests/_cases.py."""

    obs_mode = "float32"

    def __init__(self, n, size, aw, ah, rule, wrapper=None, tweak=None):
        cb = _carle()
        self.inner = cb.CARLE(instances=n, height=size, width=size, action_width=aw,
                              action_height=ah, device="cuda", obs_mode=self.obs_mode)
        self.env = self.inner
        if wrapper == "SpeedDetector":
            self.env = cb.SpeedDetector(self.inner)
        elif wrapper == "CornerBonus":
            self.env = cb.CornerBonus(self.inner)
        elif wrapper == "PufferDetector":
            self.env = cb.PufferDetector(self.inner)
            if tweak:
                self.env.growth_threshold = tweak["growth_threshold"]
        elif wrapper == "Parsimony(Corner)":
            self.env = cb.ParsimonyBonus(cb.CornerBonus(self.inner))
        elif wrapper == "MorphoBonus":
            self.env = cb.MorphoBonus(self.inner)
        elif wrapper is not None:
            raise KeyError(wrapper)
        self.inner.rules_from_string(rule)

    def reset(self):
        self.env.reset()

    def set_universe(self, u):
        self.inner.universe = torch.from_numpy(np.ascontiguousarray(u)).float()[:, None]

    def _grid(self, obs):
        if self.obs_mode == "packed":
            return self.inner.universe[:, 0].cpu().numpy().astype(np.uint8)
        return (obs[:, 0] != 0).to(torch.uint8).cpu().numpy()

    def get_universe(self):
        return self.inner.universe[:, 0].cpu().numpy().astype(np.uint8)

    def step(self, action):
        obs, reward, done, info = self.env.step(torch.from_numpy(np.asarray(action)))
        assert done.device.type == "cpu" and tuple(done.shape) == (self.inner.instances, 1)
        assert len(info) == self.inner.instances
        return self._grid(obs), reward.detach().cpu().numpy()

    def apply_action(self, action):
        self.inner.apply_action(torch.from_numpy(np.asarray(action)))

    @property
    def step_number(self):
        return self.inner.step_number


class PackedAdapter(CudaAdapter):
    obs_mode = "packed"


class Uint8Adapter(CudaAdapter):
    obs_mode = "uint8"


ADAPTERS = [CudaAdapter, PackedAdapter, Uint8Adapter]


# ---------------------------------------------------------------- golden vectors ----
@pytest.mark.parametrize("name", by_kind("rollout"))
def test_rollout_digests(name):
    cs.check_rollout(name, CudaAdapter)


@pytest.mark.parametrize("make", ADAPTERS)
def test_freerun_g5(make):
    cs.check_freerun("g5", make)


@pytest.mark.parametrize("name", by_kind("sweep"))
@pytest.mark.parametrize("make", [CudaAdapter, PackedAdapter])
def test_sweep(name, make):
    cs.check_sweep(name, make)


def test_action_values_and_broadcast():
    cs.check_action_values(CudaAdapter)


@pytest.mark.parametrize("make", ADAPTERS)
def test_master_reset_sequence(make):
    cs.check_master_reset(make)


def test_grid_sized_action_crop():
    cs.check_grid_sized_action(CudaAdapter)


def test_nonsquare_window():
    cs.check_nonsquare_window(CudaAdapter)


def test_action_placement():
    cs.check_placement(CudaAdapter)


def test_spaceship_known_answer():
    cs.check_spaceship(CudaAdapter)


@pytest.mark.parametrize("name", by_kind("wrapper"))
def test_wrappers(name):
    cs.check_wrapper(name, CudaAdapter)


@pytest.mark.parametrize("name", by_kind("parsimony"))
def test_parsimony(name):
    cs.check_parsimony(name, CudaAdapter)


# ------------------------------------------------- differential tests vs the oracle --
def _random_rule(rng):
    b = [k for k in range(9) if rng.random() < 0.4] or [3]
    s = [k for k in range(9) if rng.random() < 0.4] or [2]
    return b, s


@pytest.mark.parametrize("size,win", [(32, 16), (64, 32), (96, 32), (128, 32), (160, 64),
                                      (192, 64), (224, 64), (256, 64), (20, 10), (100, 36),
                                      (288, 64), (8, 4), (2, 2), (40, 40)])
def test_random_rules_against_oracle(size, win):
    rng = np.random.default_rng(size * 1000 + win)
    n = 3
    cb = _carle()
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                   action_height=win, device="cuda")
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    for trial in range(4):
        b, s = _random_rule(rng)
        env.birth, env.survive = list(b), list(s)          # attribute assignment path
        ref.birth, ref.survive = list(b), list(s)
        env.reset()
        ref.reset()
        soup = (rng.random((n, size, size)) < 0.45).astype(np.uint8)
        env.universe = torch.from_numpy(soup).float()[:, None]
        ref.universe = soup.copy()
        for t in range(5):
            batch = n if t % 2 == 0 else 1
            a = (rng.random((batch, 1, win, win)) <= 0.1).astype(np.float32)
            obs = env.step(torch.from_numpy(a))[0]
            want = ref.step(a)[0]
            got = obs[:, 0].cpu().numpy().astype(np.uint8)
            assert np.array_equal(got, want), (size, win, b, s, trial, t)


@pytest.mark.parametrize("size,win,n", [(64, 32, 7), (128, 32, 5), (256, 64, 3), (100, 36, 2)])
def test_step_many_equals_repeated_step(size, win, n):
    rng = np.random.default_rng(size + n)
    cb = _carle()
    k = 6
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    actions = (rng.random((k, n, 1, win, win)) <= 0.1).astype(np.float32)
    actions[3] = 1.0                                        # master reset in the middle
    envs = []
    for mode in ("step", "many"):
        env = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                       action_height=win, device="cuda", obs_mode="packed")
        env.rules_from_string("B368/S245")
        env.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        envs.append(env)
    reds = []
    envs[0].fused_reductions = True
    for t in range(k):
        envs[0].step(torch.from_numpy(actions[t]))
        reds.append(envs[0].last_reductions.clone())
    obs, red_many = envs[1].step_many(torch.from_numpy(actions), reductions=True)
    assert torch.equal(envs[0].packed_universe, envs[1].packed_universe)
    assert torch.equal(torch.stack(reds), red_many)
    assert envs[0].step_number == envs[1].step_number == 2
    # and both agree with the oracle
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    ref.rules_from_string("B368/S245")
    ref.reset()
    ref.universe = soup.copy()
    for t in range(k):
        want = ref.step(actions[t])[0]
    assert np.array_equal(envs[1].universe[:, 0].cpu().numpy().astype(np.uint8), want)
    live, sh, sw = oc.speed_sums(want, oc.outside_window_mask(ref))
    got = red_many[-1].cpu().numpy()
    assert np.array_equal(got[:, 0], live) and np.array_equal(got[:, 1], sh)
    assert np.array_equal(got[:, 2], sw)


def test_free_run_step_many_int():
    cb = _carle()
    rng = np.random.default_rng(5)
    soup = (rng.random((4, 128, 128)) < 0.35).astype(np.uint8)
    env = cb.CARLE(instances=4, height=128, width=128, action_width=32, action_height=32,
                   device="cuda")
    env.reset()
    env.universe = torch.from_numpy(soup).float()[:, None]
    obs, _ = env.step_many(16)
    u = soup
    for _ in range(16):
        u = oc.life_like_update(u, [3], [2, 3])
    assert np.array_equal(obs[:, 0].cpu().numpy().astype(np.uint8), u)
    assert env.step_number == 16 and env.steps_since_action == 16


# ------------------------------------------------------------------ API semantics ----
def test_reference_test_env_reset_semantics():
    """reference tests/test_env.py:42-67, on the default 256x256 / 64x64 env."""
    env = _carle().CARLE()
    reset_observation = env.reset()
    action = torch.ones(env.instances, 1, env.action_height, env.action_width)
    toggle_observation = env.step(action)[0]
    action[:, :, 0:10, 0:10] = 0.0
    normal_observation = env.step(action)[0]
    assert toggle_observation.mean().item() == 0.0
    assert reset_observation.mean().item() == 0.0
    assert 1.0 == (1.0 * (reset_observation == toggle_observation)).mean().item()
    assert 1.0 != (1.0 * (toggle_observation == normal_observation)).mean().item()


def test_reference_test_env_rule_setting():
    """reference tests/test_env.py:17-39."""
    env = _carle().CARLE()
    env.birth_rule_from_string("asdfasdfB0357*!@#!@$%")
    env.survive_rule_from_string("S2468")
    assert env.birth == [0, 3, 5, 7] and env.survive == [2, 4, 6, 8]
    env.rules_from_string("B0357/S2468")
    assert env.birth == [0, 3, 5, 7] and env.survive == [2, 4, 6, 8]
    env.rules_from_string("23/3")
    assert env.birth == [2, 3] and env.survive == [3]
    with pytest.raises(IndexError):
        env.rules_from_string("B3S23")


def test_obs_aliases_universe_and_inplace_edits_are_seen():
    cb = _carle()
    env = cb.CARLE(instances=2, height=64, width=64, action_width=32, action_height=32)
    obs = env.reset()
    assert obs is env.universe and obs.dtype == torch.float32
    assert tuple(obs.shape) == (2, 1, 64, 64)
    obs[1, 0, 10, 10:13] = 1.0                              # blinker written in place
    zero = torch.zeros(2, 1, 32, 32)
    obs2, reward, done, info = env.step(zero)
    assert obs2 is env.universe and obs2 is not obs
    want = np.zeros((2, 64, 64), dtype=np.uint8)
    want[1, 9:12, 11] = 1
    assert np.array_equal(obs2[:, 0].cpu().numpy().astype(np.uint8), want)
    assert reward.device.type == "cuda" and tuple(reward.shape) == (2, 1)
    assert float(reward.abs().sum()) == 0.0
    # universe[idx, 0] = grid   (reference load_universe path, env.py:406)
    env.universe[0, 0, :, :] = torch.from_numpy(want[1]).float()
    obs3 = env.step(zero)[0]
    assert int(obs3[0].sum()) == 3 and int(obs3[1].sum()) == 3


def test_errors_match_reference_types():
    cb = _carle()
    env = cb.CARLE(instances=2, height=64, width=64, action_width=32, action_height=32)
    with pytest.raises(AttributeError):
        env.step(torch.zeros(2, 1, 32, 32))                 # before reset()
    env.reset()
    with pytest.raises(AssertionError):
        env.step(torch.zeros(2, 1, 16, 32))
    env.survive = []
    with pytest.raises(TypeError):
        env.step(torch.zeros(2, 1, 32, 32))
    env.survive = [2, 3]
    with pytest.raises(ValueError):
        cb.CARLE(height=65, width=65).reset()               # odd grid
    with pytest.raises(ValueError):
        cb.CARLE(height=64, width=128).reset()              # non-square grid
    with pytest.raises(RuntimeError):
        cb.CARLE(device="cpu")
    with pytest.raises(AttributeError):
        cb.CARLE(device="tpu")


def test_counters_are_lazy_device_side():
    cb = _carle()
    env = cb.CARLE(instances=3, height=64, width=64, action_width=32, action_height=32)
    env.reset()
    zero = torch.zeros(3, 1, 32, 32)
    some = torch.zeros(3, 1, 32, 32)
    some[1, 0, 4, 4] = 1.0
    for a in (zero, some, zero, zero):
        env.step(a)
    assert env.step_number == 4 and env.steps_since_action == 3
    env.step(torch.ones(1, 1, 32, 32))
    assert env.step_number == 0 and env.steps_since_action == 0


def test_state_dict_has_reference_key_and_module_api():
    cb = _carle()
    env = cb.SpeedDetector(cb.CARLE(instances=1, height=64, width=64, action_width=32,
                                    action_height=32))
    keys = set(env.state_dict().keys())
    assert {"inner_env.neighborhood.weight", "env.neighborhood.weight"} <= keys
    w = env.state_dict()["inner_env.neighborhood.weight"]
    assert tuple(w.shape) == (1, 1, 3, 3) and float(w.sum()) == 8.0
    env.eval()
    env.to(env.my_device)
    pad = env.inner_env.action_padding(torch.ones(1, 1, 32, 32))
    assert tuple(pad.shape) == (1, 1, 64, 64) and float(pad.sum()) == 1024.0


def test_uint8_and_bool_actions():
    cb = _carle()
    rng = np.random.default_rng(11)
    soup = (rng.random((3, 128, 128)) < 0.4).astype(np.uint8)
    a = (rng.random((3, 1, 32, 32)) <= 0.2)
    outs = []
    for cast in (lambda t: t.float(), lambda t: t.to(torch.uint8), lambda t: t,
                 lambda t: t.double(), lambda t: t.float().cuda()):
        env = cb.CARLE(instances=3, height=128, width=128, action_width=32, action_height=32)
        env.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        outs.append(env.step(cast(torch.from_numpy(a)))[0].cpu())
    for o in outs[1:]:
        assert torch.equal(outs[0], o)


def test_reference_style_wrappers_run_on_float_obs():
    """A reference-style torch reduction over obs/universe (what carle/mcl.py does) gives
    the same numbers as the fused device reductions."""
    cb = _carle()
    rng = np.random.default_rng(3)
    env = cb.CARLE(instances=4, height=128, width=128, action_width=32, action_height=32,
                   fused_reductions=True)
    env.reset()
    env.universe = torch.from_numpy((rng.random((4, 128, 128)) < 0.3).astype(np.float32))[:, None]
    obs = env.step(torch.zeros(4, 1, 32, 32))[0]
    padded = env.action_padding(torch.ones(1, 1, 32, 32, device=obs.device))
    mask = torch.ones_like(padded) - padded
    rows = torch.arange(128, device=obs.device).reshape(-1, 1) * mask
    cols = torch.arange(128, device=obs.device).reshape(1, -1) * mask
    red = env.last_reductions
    assert torch.equal(red[:, 0], obs.sum(dim=[1, 2, 3]).long())
    assert torch.equal(red[:, 1], (obs * rows).sum(dim=[1, 2, 3]).long())
    assert torch.equal(red[:, 2], (obs * cols).sum(dim=[1, 2, 3]).long())
    assert torch.equal(env.reduce(), red)


# ----------------------------------------------- full-size, size-independent checks ----
def test_config2_shape_properties():
    """BASELINE config 2 (4096 x 128x128): oracle on a random subset of instances +
    conservation properties that need no oracle."""
    cb = _carle()
    n, size, win = 4096, 128, 32
    g = torch.Generator(device="cuda").manual_seed(2)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                   action_height=win, obs_mode="packed", fused_reductions=True)
    env.reset()
    soup = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.5).float()
    env.universe = soup
    pick = [0, 1, 777, 2048, 4095]
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=len(pick))
    ref.reset()
    ref.universe = soup[pick, 0].cpu().numpy().astype(np.uint8)
    for t in range(4):
        a = 1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1)
        env.step(a)
        ref.step(a[pick].cpu().numpy())
    got = env.universe[pick, 0].cpu().numpy().astype(np.uint8)
    assert np.array_equal(got, ref.universe)
    # popcount of the float view == fused live count, for all 4096 instances
    assert torch.equal(env.universe.sum(dim=[1, 2, 3]).long(), env.last_reductions[:, 0])
    # translation equivariance on the torus: shifting the soup shifts the result
    env2 = cb.CARLE(instances=8, height=size, width=size, action_width=win,
                    action_height=win)
    env2.reset()
    base = soup[:8]
    env2.universe = base
    o1 = env2.step_many(5)[0].clone()
    env2.universe = torch.roll(base, shifts=(37, -53), dims=(2, 3))
    o2 = env2.step_many(5)[0]
    assert torch.equal(torch.roll(o1, shifts=(37, -53), dims=(2, 3)), o2)


@pytest.mark.parametrize("n,size,win,rule", [(16384, 256, 64, "B368/S245"),
                                             (131072, 64, 32, "B3/S23")])
def test_full_size_batches_properties(n, size, win, rule):
    """BASELINE config 3 (16384 x 256x256, Morley, fused sums) and the config-4 per-GPU shard
    (131072 x 64x64) at full size: oracle on a random subset of instances, the fused live count
    against a popcount of the packed state for EVERY instance, linearity of the action
    (XOR-ing the same action twice is the identity before the generation), and the window-live
    count after an all-ones-but-one action."""
    cb = _carle()
    g = torch.Generator(device="cuda").manual_seed(n)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed", fused_reductions=True)
    env.rules_from_string(rule)
    env.reset()
    words = torch.randint(-2**31, 2**31 - 1, (n, size, size // 32), dtype=torch.int32,
                          device="cuda", generator=g)
    env.packed_universe.copy_(words)
    pick = sorted(set(int(v) for v in torch.randint(0, n, (6,), generator=torch.Generator().manual_seed(n))) | {0, n - 1})
    bits = ((words[pick].unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1)
    soup = bits.reshape(len(pick), size, size).to(torch.uint8).cpu().numpy()
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=len(pick))
    ref.rules_from_string(rule)
    ref.reset()
    ref.universe = soup.copy()
    for t in range(3):
        a = 1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1)
        env.step(a)
        ref.step(a[pick].cpu().numpy())
    packed = env.packed_universe
    gbits = ((packed[pick].unsqueeze(-1) >> torch.arange(32, device="cuda", dtype=torch.int32)) & 1)
    got = gbits.reshape(len(pick), size, size).to(torch.uint8).cpu().numpy()
    assert np.array_equal(got, ref.universe)
    live, sh, sw = oc.speed_sums(ref.universe, oc.outside_window_mask(ref))
    red = env.last_reductions
    assert np.array_equal(red[pick, 0].cpu().numpy(), live)
    assert np.array_equal(red[pick, 1].cpu().numpy(), sh)
    assert np.array_equal(red[pick, 2].cpu().numpy(), sw)
    # fused live count == popcount of the packed words, all instances
    pop = torch.zeros(n, dtype=torch.int64, device="cuda")
    chunk = 4096
    for i in range(0, n, chunk):
        w = packed[i:i + chunk].to(torch.int64) & 0xFFFFFFFF
        c = torch.zeros(w.shape[0], dtype=torch.int64, device="cuda")
        for b in range(32):
            c += ((w >> b) & 1).sum(dim=(1, 2))
        pop[i:i + chunk] = c
    assert torch.equal(pop, red[:, 0])
    # an all-ones action with a single zero toggles the window but must NOT reset
    before = env.packed_universe.clone()
    a = torch.ones(n, 1, win, win, device="cuda")
    a[n // 3, 0, 1, 2] = 0.0
    env.apply_action(a)
    env.apply_action(a)                                 # XOR twice: identity
    assert torch.equal(env.packed_universe, before)
    env.step(a)
    assert env.step_number == 4                         # no master reset
    assert int(env.last_reductions[:, 0].sum()) > 0


def test_large_generic_grid_properties():
    """1 x 2048x2048 through the generic family: a glider returns to itself after
    4 * size generations on the torus ... checked at a cheaper scale by shift
    equivariance and still-life / blinker invariants."""
    cb = _carle()
    size = 2048
    env = cb.CARLE(instances=1, height=size, width=size)
    obs = env.reset()
    u = torch.zeros(1, 1, size, size)
    u[0, 0, 0, 0:2] = 1.0                                    # block across the corner seam
    u[0, 0, size - 1, 0:2] = 1.0
    u[0, 0, 100, size - 1] = 1.0                             # blinker across the right seam
    u[0, 0, 100, 0] = 1.0
    u[0, 0, 100, 1] = 1.0
    env.universe = u
    o = env.step_many(2)[0]
    assert torch.equal(o.cpu(), u)                           # period-2 / still life
    o1 = env.step_many(1)[0].cpu()
    assert int(o1.sum()) == 7
    assert o1[0, 0, 99, 0] == 1 and o1[0, 0, 101, 0] == 1 and o1[0, 0, 100, 0] == 1


# ------------------------------------------------ fused vs two-kernel step paths ------
@pytest.mark.parametrize("size,win,n", [(64, 32, 37), (128, 32, 21), (256, 64, 9),
                                        (64, 31, 5), (128, 30, 5), (96, 32, 6)])
def test_fused_step_matches_packed_path_and_oracle(size, win, n):
    """carle_step_action (one fused kernel when the window is lane-aligned, otherwise
    pack + step) == step_many with pre-packed actions == oracle, including a batch-wide
    master reset spread over several blocks and uint8 / batch-1 actions."""
    cb = _carle()
    rng = np.random.default_rng(size * 7 + win)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    seq = []
    for t in range(7):
        batch = 1 if t == 2 else n
        a = (rng.random((batch, 1, win, win)) <= 0.15).astype(np.float32)
        if t == 4:
            a[:] = 1.0                               # master reset
        if t == 5:
            a[:] = 1.0
            a[n // 2, 0, win - 1, win - 1] = 0.0     # one zero in one instance: no reset
        seq.append(a)
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   fused_reductions=True)
    env.rules_from_string("B368/S245")
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    ref.rules_from_string("B368/S245")
    env.reset()
    ref.reset()
    env.universe = torch.from_numpy(soup).float()[:, None]
    ref.universe = soup.copy()
    for t, a in enumerate(seq):
        ta = torch.from_numpy(a)
        if t % 2:
            ta = ta.to(torch.uint8)
        obs = env.step(ta)[0]
        want = ref.step(a)[0]
        assert np.array_equal(obs[:, 0].cpu().numpy().astype(np.uint8), want), t
        live, sh, sw = oc.speed_sums(want, oc.outside_window_mask(ref))
        red = env.last_reductions.cpu().numpy()
        assert np.array_equal(red[:, 0], live) and np.array_equal(red[:, 1], sh), t
        assert np.array_equal(red[:, 2], sw), t
        assert env.step_number == ref.step_number, t
        assert env.steps_since_action == ref.steps_since_action, t
    # the pre-packed multi-step path gives the same final state
    env2 = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                    action_height=win, obs_mode="packed")
    env2.rules_from_string("B368/S245")
    env2.reset()
    env2.universe = torch.from_numpy(soup).float()[:, None]
    for a in seq:
        full = np.broadcast_to(a, (n, 1, win, win)).copy()
        env2.step_many(torch.from_numpy(full)[None])
    assert torch.equal(env2.packed_universe, env.packed_universe)


VARIANTS_256 = [{"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "2"},
                {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "4"},
                {"CARLE_FUSED_IMPL": "strip", "CARLE_STRIP_R": "2", "CARLE_PDL": "0"},
                {"CARLE_FUSED_IMPL": "quad"}, {"CARLE_FUSED_IMPL": "direct"},
                {"CARLE_FUSED_IMPL": "tma"}]
VARIANTS_128 = [{"CARLE_FUSED_IMPL": "strip"}, {"CARLE_FUSED_IMPL": "strip", "CARLE_PDL": "0"},
                {"CARLE_FUSED_IMPL": "tma"}, {"CARLE_FUSED_IMPL": "tma", "CARLE_PDL": "0"},
                {"CARLE_FUSED_IMPL": "direct"}]


@pytest.mark.parametrize("size,win,n,variant",
                         [(256, 64, 13, v) for v in VARIANTS_256] +
                         [(256, 64, 700, VARIANTS_256[0]), (256, 64, 650, VARIANTS_256[1])] +
                         [(128, 32, 29, v) for v in VARIANTS_128] +
                         [(128, 32, 5000, VARIANTS_128[0]), (128, 32, 5000, VARIANTS_128[2]),
                          (64, 32, 9000, {"CARLE_FUSED_IMPL": "tma"})])
def test_every_step_kernel_variant_matches_oracle(size, win, n, variant, monkeypatch):
    """Every implementation of the one-launch env step (independent strips with 2 or 4 rows per
    lane, four warps per instance, one warp per instance with plain loads or the TMA pipeline;
    with and without programmatic dependent launch) is bit-exact against the oracle: float32 /
    uint8 / batch-1 actions, master reset, fused sums, static and run-time rules, and batches
    large enough for several trips of the persistent warps."""
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    cb = _carle()
    rng = np.random.default_rng(size + n)
    big = n > 100
    check = np.arange(n) if not big else np.unique(rng.integers(0, n, size=24))
    for rule in ("B3/S23", "B368/S245", "B2/S0123"):
        soup = (rng.random((n, size, size)) < 0.37).astype(np.uint8)
        env = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                       action_height=win, fused_reductions=True, obs_mode="packed")
        env.rules_from_string(rule)
        ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                             instances=len(check))
        ref.rules_from_string(rule)
        env.reset()
        ref.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        ref.universe = soup[check].copy()
        for t in range(6 if not big else 3):
            batch = 1 if t == 2 else n
            a = (rng.random((batch, 1, win, win)) <= 0.12).astype(np.float32)
            if t == (1 if big else 4):
                a[:] = 1.0                               # master reset
            if t == 5:
                a[:] = 1.0
                a[n // 2, 0, 0, win - 1] = 0.0           # one zero in one instance: no reset
            ta = torch.from_numpy(a)
            if t % 2:
                ta = ta.to(torch.uint8)
            env.step(ta)
            want = ref.step(a if batch == 1 else a[check])[0]
            got = env.universe[:, 0].cpu().numpy().astype(np.uint8)
            assert np.array_equal(got[check], want), (rule, t)
            live, sh, sw = oc.speed_sums(got, oc.outside_window_mask(ref))
            red = env.last_reductions.cpu().numpy()
            assert np.array_equal(red[:, 0], live) and np.array_equal(red[:, 1], sh), (rule, t)
            assert np.array_equal(red[:, 2], sw), (rule, t)
            assert env.step_number == ref.step_number, (rule, t)


@pytest.mark.parametrize("size,win,n,variant", [(256, 64, 40, {"CARLE_STRIP_R": "2"}),
                                                (256, 64, 40, {"CARLE_STRIP_R": "4"}),
                                                (128, 32, 333, {"CARLE_FUSED_IMPL": "strip"}),
                                                (128, 32, 333, {"CARLE_FUSED_IMPL": "tma"})])
def test_programmatic_dependent_launch_chain_in_a_graph(size, win, n, variant, monkeypatch):
    """Back-to-back steps launched with programmatic stream serialization (each kernel may start
    while its predecessor drains, then waits on griddepcontrol) give the same states as the
    serialised launches, eagerly and as a replayed CUDA graph."""
    import carle_b200
    from carle_b200 import _lib
    lib = _lib.load()
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    g = torch.Generator(device="cuda").manual_seed(size + n)
    soup = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.5).float()
    acts = [1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1) for _ in range(9)]
    red = torch.zeros(n, 4, dtype=torch.int64, device="cuda")

    def rollout(pdl, graph):
        monkeypatch.setenv("CARLE_PDL", pdl)
        env = carle_b200.CARLE(instances=n, height=size, width=size, action_width=win,
                               action_height=win, obs_mode="packed")
        env.reset()
        env.universe = soup
        env._sync_rule()

        def steps():
            for a in acts:
                rc = lib.carle_step_action(env._handle, env._packed.data_ptr(),
                                           env._spare.data_ptr(), a.data_ptr(), _lib.F32, n,
                                           env._counters.data_ptr(), red.data_ptr(), env._stream())
                assert rc == 0
                env._packed, env._spare = env._spare, env._packed
        if graph:
            start = env.packed_universe.clone()
            steps()                                   # warm-up outside the capture (odd count: swaps buffers)
            torch.cuda.synchronize()
            first_in = env._packed                    # the buffer the captured chain starts from
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                steps()
            first_in.copy_(start)
            cg.replay()
        else:
            steps()
        torch.cuda.synchronize()
        return env.packed_universe.clone(), red.clone()

    want, want_red = rollout("0", False)
    for pdl, graph in (("1", False), ("1", True), ("0", True)):
        got, got_red = rollout(pdl, graph)
        assert torch.equal(got, want), (pdl, graph)
        assert torch.equal(got_red, want_red), (pdl, graph)


@pytest.mark.parametrize("size,win,n", [(64, 32, 50), (128, 32, 40), (256, 64, 12)])
def test_jit_specialised_rules_match_runtime_rule_kernels_and_oracle(size, win, n, monkeypatch):
    """A rule without a built-in instantiation is compiled with NVRTC into a StaticRule kernel
    on first use (jit.cu): same states as the run-time-rule kernels (CARLE_JIT=0) and as the
    oracle, also when the rule changes between steps; the specialised kernel really is loaded."""
    from carle_b200 import _lib
    cb = _carle()
    lib = _lib.load()
    rng = np.random.default_rng(size)
    rules = ["B36/S125", "B2/S", "B345/S4567", "B1357/S1357"]
    rules[1] = "B2/S0"                                        # (empty survive set is a TypeError)
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    acts = [(rng.random((n, 1, win, win)) <= 0.1).astype(np.float32) for _ in range(8)]

    def rollout(jit):
        monkeypatch.setenv("CARLE_JIT", jit)
        env = cb.CARLE(instances=n, height=size, width=size, action_width=win,
                       action_height=win, obs_mode="packed", fused_reductions=True)
        env.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        out = []
        for t, a in enumerate(acts):
            env.rules_from_string(rules[t % len(rules)])
            env.step(torch.from_numpy(a))
            out.append((env.universe[:, 0].cpu().numpy().astype(np.uint8),
                        env.last_reductions.cpu().numpy().copy()))
        return out

    before = lib.carle_jit_loaded()
    plain = rollout("0")
    assert lib.carle_jit_loaded() == before
    fast = rollout("1")
    assert lib.carle_jit_loaded() >= before + len(rules)
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
    ref.reset()
    ref.universe = soup.copy()
    for t, a in enumerate(acts):
        ref.rules_from_string(rules[t % len(rules)])
        want = ref.step(a)[0]
        assert np.array_equal(fast[t][0], want), t
        assert np.array_equal(plain[t][0], want), t
        assert np.array_equal(fast[t][1], plain[t][1]), t


@pytest.mark.parametrize("size,win,n,k", [(256, 64, 6, 9), (96, 32, 5, 7), (320, 64, 2, 19),
                                          (100, 36, 3, 4)])
def test_jit_specialised_multi_generation_kernels(size, win, n, k, monkeypatch):
    """NVRTC specialisation of the multi-generation (register-resident), tiled and any-shape
    kernels behind step_many: K zero-action generations of an arbitrary rule == run-time-rule
    kernels == oracle."""
    from carle_b200 import _lib
    cb = _carle()
    lib = _lib.load()
    rng = np.random.default_rng(size + k)
    rule = "B3578/S24678"
    soup = (rng.random((n, size, size)) < 0.45).astype(np.uint8)

    def run(jit):
        monkeypatch.setenv("CARLE_JIT", jit)
        env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win)
        env.rules_from_string(rule)
        env.reset()
        env.universe = torch.from_numpy(soup).float()[:, None]
        env.step_many(k)
        return env.universe[:, 0].cpu().numpy().astype(np.uint8)

    before = lib.carle_jit_loaded()
    plain = run("0")
    fast = run("1")
    assert lib.carle_jit_loaded() > before
    b, sv = oc.rules_from_string(rule)
    want = soup
    for _ in range(k):
        want = oc.life_like_update(want, b, sv)
    assert np.array_equal(fast, want)
    assert np.array_equal(plain, want)


def test_new_rule_first_stepped_inside_graph_capture():
    """A rule whose specialised kernel does not exist yet may be stepped for the first time while
    a CUDA graph is being captured: whether or not the driver accepts the module load there, the
    captured steps are exact (the run-time-rule kernel is the fallback) and replay correctly."""
    from carle_b200 import _lib
    cb = _carle()
    lib = _lib.load()
    n, size, win = 24, 64, 32
    rng = np.random.default_rng(99)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    acts = [(rng.random((n, 1, win, win)) <= 0.1).astype(np.float32) for _ in range(4)]
    dacts = [torch.from_numpy(a).cuda() for a in acts]
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed")
    env.reset()
    env.universe = torch.from_numpy(soup).float()[:, None]
    env.rules_from_string("B3567/S01458")                    # used nowhere else in the suite
    env._sync_rule()
    start = env.packed_universe.clone()

    def abi_steps():
        for a in dacts:
            rc = lib.carle_step_action(env._handle, env._packed.data_ptr(), env._spare.data_ptr(),
                                       a.data_ptr(), _lib.F32, n, env._counters.data_ptr(), None,
                                       env._stream())
            assert rc == 0, _lib.last_error()
            env._packed, env._spare = env._spare, env._packed
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        abi_steps()
    env._packed.copy_(start)
    graph.replay()
    torch.cuda.synchronize()
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
    ref.rules_from_string("B3567/S01458")
    ref.reset()
    ref.universe = soup.copy()
    for a in acts:
        want = ref.step(a)[0]
    assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), want)
    # outside the capture the same rule steps (specialised by now or on this call) and agrees
    a = (rng.random((n, 1, win, win)) <= 0.1).astype(np.float32)
    env.step(torch.from_numpy(a))
    want = ref.step(a)[0]
    assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), want)


@pytest.mark.parametrize("n", [5000, 5003])
def test_speed_detector_tail_kernel_matches_reference_arithmetic(n):
    """carle_speed_tail (one launch for mcl.py:777-795) on a multi-block batch: centre of mass
    bit-exact against float32 numpy, speed within float32 rounding of the reference's
    sqrt(sum(v^2)), reward += speed for every instance (16-byte and scalar tails of the reward
    column), first call leaves reward untouched; live_cells is the step's live count."""
    cb = _carle()
    size, win = 64, 32
    env = cb.SpeedDetector(cb.CARLE(instances=n, height=size, width=size, action_width=win,
                                    action_height=win, obs_mode="packed"))
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    env.inner_env.universe = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.2).float()
    ref = oc.OracleSpeedDetector(oc.OracleCARLE(width=size, height=size, action_width=win,
                                                action_height=win, instances=n))
    ref.reset()
    ref.env.universe = env.inner_env.universe[:, 0].cpu().numpy().astype(np.uint8)
    for t in range(4):
        a = 1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1)
        _, reward, _, _ = env.step(a)
        _, want, _, _ = ref.step(a.cpu().numpy())
        got = reward.cpu().numpy()
        assert got.shape == (n, 1)
        np.testing.assert_allclose(got, np.broadcast_to(np.asarray(want, dtype=np.float32), got.shape),
                                   rtol=3e-6, atol=1e-6, err_msg=f"step {t}")
        assert np.array_equal(env.center_of_mass.cpu().numpy(), ref.center_of_mass), t
        assert np.array_equal(env.live_cells.cpu().numpy(), ref.live_cells.astype(np.float32)), t
        if t == 0:
            assert float(np.abs(got).max()) == 0.0
        else:
            assert float(env.speed) > 0.0


# ------------------------------------------------------- tiled family (large grids) ----
@pytest.mark.parametrize("size,win,n,k", [(288, 64, 2, 5), (320, 64, 2, 21), (512, 64, 1, 37),
                                          (1024, 64, 1, 40), (480, 32, 3, 16)])
def test_tiled_temporal_blocking_matches_oracle(size, win, n, k):
    """step_many on the tiled family runs blocks of 16 generations inside register tiles
    (overlapped 256x256 tiles, halo discarded); it must equal k single generations of the
    oracle, with actions every generation and a master reset inside a block."""
    cb = _carle()
    rng = np.random.default_rng(size + k)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    actions = (rng.random((k, n, 1, win, win)) <= 0.1).astype(np.float32)
    if k > 8:
        actions[7] = 1.0                                      # master reset inside block 0
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed")
    assert env.reset() is not None and env.kernel_family == 2
    env.universe = torch.from_numpy(soup).float()[:, None]
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    ref.reset()
    ref.universe = soup.copy()
    env.step_many(torch.from_numpy(actions))
    for t in range(k):
        want = ref.step(actions[t])[0]
    assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), want)
    assert env.step_number == ref.step_number
    # a further free run, not a multiple of the block length, and per-step API on top
    env.step_many(19)
    obs = env.step(torch.from_numpy(actions[0]))[0] if False else None
    for _ in range(19):
        want = ref.step(np.zeros((n, 1, win, win), dtype=np.float32))[0]
    assert np.array_equal(env.universe[:, 0].cpu().numpy().astype(np.uint8), want)


def test_tiled_glider_crosses_all_seams():
    """A glider on a 576 x 576 torus (not a multiple of the 224 x 192 tile interior) walks
    across tile seams and both torus seams; after 4*576 generations it is back home."""
    cb = _carle()
    size = 576
    env = cb.CARLE(instances=1, height=size, width=size)
    env.reset()
    u = torch.zeros(1, 1, size, size)
    for r, c in ((0, 1), (1, 2), (2, 0), (2, 1), (2, 2)):
        u[0, 0, r, c] = 1.0
    env.universe = u
    env.step_many(4 * size)
    assert torch.equal(env.universe.cpu(), u)
    env.step_many(2 * size)
    moved = torch.roll(u, shifts=(size // 2, size // 2), dims=(2, 3))
    assert torch.equal(env.universe.cpu(), moved)


# --------------------------------------------- row-band giant grid (config 5 code path) ----
def _pack_rows(u):
    """uint8 [H, W] -> int32 [H, W/32] in the library's layout."""
    words = np.packbits(u, axis=-1, bitorder="little").view("<u4")
    return torch.from_numpy(words.view(np.int32).copy())


def _unpack_rows(t, w):
    b = np.unpackbits(t.cpu().numpy().view(np.uint8), axis=-1, bitorder="little")
    return b[:, :w]


@pytest.mark.parametrize("size,halo,k", [(512, 16, 40), (512, 8, 13), (1024, 32, 70)])
def test_banded_single_rank_equals_oracle(size, halo, k):
    """BandedCARLE with one rank (its neighbours are itself) == oracle on the torus, with an
    action every generation in the central window and a master reset."""
    from carle_b200.bigrid import BandedCARLE
    rng = np.random.default_rng(size + halo)
    soup = (rng.random((size, size)) < 0.4).astype(np.uint8)
    win = 64
    actions = (rng.random((k, 1, 1, win, win)) <= 0.1).astype(np.float32)
    actions[5] = 1.0
    grid = BandedCARLE(size, size, rule="B36/S23", halo=halo, action_height=win,
                       action_width=win)
    try:
        grid.set_band(_pack_rows(soup))
        grid.step_many(k, torch.from_numpy(actions))
        ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                             instances=1)
        ref.rules_from_string("B36/S23")
        ref.reset()
        ref.universe = soup[None].copy()
        for t in range(k):
            want = ref.step(actions[t, 0])[0]
        assert np.array_equal(_unpack_rows(grid.band, size), want[0])
        grid.step_many(2 * halo + 3)                      # free run, partial last block
        for _ in range(2 * halo + 3):
            want = ref.step(np.zeros((1, 1, win, win), dtype=np.float32))[0]
        assert np.array_equal(_unpack_rows(grid.band, size), want[0])
    finally:
        grid.close()


def _run_band_worker(script, extra, ranks):
    import subprocess
    import sys
    root = __import__("os").path.dirname(__import__("os").path.dirname(__file__))
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}",
         "--master-addr", "127.0.0.1", "--master-port", "29611", script] + extra,
        cwd=root, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    return proc.stdout


def test_banded_multi_gpu_if_available():
    """Two (or more) ranks, NVLink peer stores + neighbour flags, checked against the single-GPU
    tiled path inside tools/bigrid_check.py.  Skipped on a one-GPU box."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run_band_worker("tools/bigrid_check.py", ["--size", "4096", "--gens", "50"], min(n, 8))
    assert "BIGRID CHECK OK" in out


@pytest.mark.parametrize("mode,size,gens", [("whole", 4096, 40), ("stripes", 65536, 16)])
def test_banded_grid_against_the_oracle(mode, size, gens):
    """Every visible GPU (one is enough: a single band is its own neighbour) runs the row-band
    path; 4096^2 with actions is compared with the oracle on the whole torus, 65536^2 -- the
    BASELINE configs[4] size -- on stripes around every band boundary and tile seam
    (tests/band_check_worker.py)."""
    ranks = min(torch.cuda.device_count(), 8)
    out = _run_band_worker("tests/band_check_worker.py",
                           ["--mode", mode, "--size", str(size), "--gens", str(gens)], ranks)
    assert "BAND ORACLE CHECK OK" in out, out[-1000:]


# ----------------------------------------------------------- more API-surface checks ----
def test_default_env_and_instances_changed_before_reset():
    """CARLE() defaults (256x256, 64x64 window, env.py:21-24); `instances` may be changed
    before reset() (env.py:563); use_cuda kwarg; get_observation aliases universe."""
    cb = _carle()
    env = cb.CARLE(use_cuda=True)
    assert (env.width, env.height, env.action_width, env.action_height) == (256, 256, 64, 64)
    assert env.birth == [3] and env.survive == [2, 3] and env.inner_env is None
    env.instances = 3
    obs = env.reset()
    assert tuple(obs.shape) == (3, 1, 256, 256)
    glider = torch.zeros(1, 1, 64, 64)
    glider[:, :, 32, 32] = 1.0
    glider[:, :, 33, 32:34] = 1.0
    glider[:, :, 34, 31] = 1.0
    glider[:, :, 34, 33] = 1.0                       # carle/mcl.py:872-879 get_glider()
    obs = env.step(glider)[0]
    assert env.get_observation() is obs
    ref = oc.OracleCARLE(instances=3)
    ref.reset()
    want = ref.step(glider.numpy())[0]
    assert np.array_equal(obs[:, 0].cpu().numpy().astype(np.uint8), want)
    for _ in range(7):
        obs = env.step(torch.zeros(1, 1, 64, 64))[0]
        want = ref.step(np.zeros((1, 1, 64, 64), dtype=np.float32))[0]
    assert np.array_equal(obs[:, 0].cpu().numpy().astype(np.uint8), want)
    assert int(obs[0].sum()) == 5                    # still a glider


def test_apply_action_then_zero_step_equals_step():
    cb = _carle()
    rng = np.random.default_rng(21)
    a = torch.from_numpy((rng.random((4, 1, 32, 32)) < 0.3).astype(np.float32))
    soup = torch.from_numpy((rng.random((4, 1, 128, 128)) < 0.4).astype(np.float32))
    e1 = cb.CARLE(instances=4, height=128, width=128, action_width=32, action_height=32)
    e2 = cb.CARLE(instances=4, height=128, width=128, action_width=32, action_height=32)
    for e in (e1, e2):
        e.reset()
        e.universe = soup
    e1.apply_action(a)
    e1.apply_action(a)                               # toggling twice is the identity
    assert torch.equal(e1.universe.cpu(), soup)
    e1.apply_action(a)
    o1 = e1.step(torch.zeros(4, 1, 32, 32))[0]
    o2 = e2.step(a)[0]
    assert torch.equal(o1, o2)


def test_logging_writes_reference_style_rle(tmp_path, monkeypatch):
    """logging=True: step() appends [action_rle, universe_rle] before applying the action
    (env.py:194-195, 466-476); the universe RLE round-trips through load_universe."""
    cb = _carle()
    monkeypatch.chdir(tmp_path)
    (tmp_path / "logs").mkdir()
    env = cb.CARLE(instances=2, height=64, width=64, action_width=32, action_height=32,
                   logging=True)
    env.reset()
    a = torch.zeros(2, 1, 32, 32)
    a[0, 0, 3, 4:9] = 1.0
    env.step(a)
    env.step(torch.zeros(2, 1, 32, 32))
    assert len(env.log) == 2
    action_rle, universe_rle = env.log[1]
    assert "(action)" in action_rle and "(universe)" in universe_rle
    assert "rule = B3/S23:T64, 64" in universe_rle and universe_rle.endswith("!")
    env.save_log()
    rle = env.get_rle(env.universe[0, 0])
    env.save_rle(rle)
    path = next((tmp_path / "logs").glob("universe*.rle"))
    env2 = cb.CARLE(instances=1, height=64, width=64, action_width=32, action_height=32)
    env2.reset()
    env2.load_universe(str(path))
    assert torch.equal(env2.universe[0, 0].cpu(), env.universe[0, 0].cpu())
    (tmp_path / "frames").mkdir()
    env.save_frame()
    png = next((tmp_path / "frames").glob("frame*.png")).read_bytes()
    assert png[:8] == b"\x89PNG\r\n\x1a\n"


def test_cuda_graph_capture_of_step_action():
    """The C ABI launches no hidden memsets/allocations, so K steps capture into one CUDA
    graph and replay deterministically (what bench.py times)."""
    import ctypes
    cb = _carle()
    from carle_b200 import _lib
    lib = _lib.load()
    n, size, win = 64, 128, 32
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                   obs_mode="packed")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(9)
    soup = (torch.rand(n, 1, size, size, device="cuda", generator=g) < 0.5).float()
    acts = [1.0 * (torch.rand(n, 1, win, win, device="cuda", generator=g) <= 0.1) for _ in range(4)]
    env.universe = soup
    env._sync_rule()
    start = env.packed_universe.clone()

    def abi_steps():
        for a in acts:
            rc = lib.carle_step_action(env._handle, env._packed.data_ptr(), env._spare.data_ptr(),
                                       a.data_ptr(), _lib.F32, n, env._counters.data_ptr(), None,
                                       env._stream())
            assert rc == 0
            env._packed, env._spare = env._spare, env._packed
    abi_steps()
    torch.cuda.synchronize()
    eager = env.packed_universe.clone()
    env._packed.copy_(start)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        abi_steps()
    env._packed.copy_(start)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.packed_universe, eager)
    env._packed.copy_(start)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.packed_universe, eager)
    assert env.step_number == 12


# -------------------------------------------------- device-side random agent / packed actions --
def test_device_random_agent_distribution_and_parity():
    """DeviceRandomAgent draws Bernoulli(0.1) toggles on the device in the packed layout
    (carle/agents.py:35-42 gives the distribution; the bit stream is Philox, not torch's).
    Stepping with the PackedAction == stepping with its float32 expansion == oracle."""
    cb = _carle()
    n, size, win = 512, 128, 32
    env = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win)
    env.reset()
    agent = cb.DeviceRandomAgent(env, toggle_rate=0.1, seed=123, lazy=False)
    a0, a1 = agent(None), agent(None)
    f0, f1 = a0.to_float(), a1.to_float()
    assert tuple(f0.shape) == (n, 1, win, win) and f0.dtype == torch.float32
    cells = n * win * win
    for f in (f0, f1):
        vals = torch.unique(f).tolist()
        assert vals == [0.0, 1.0]
        rate = float(f.mean())
        assert abs(rate - 0.1) < 5 * (0.1 * 0.9 / cells) ** 0.5 + 1e-4      # 5 sigma
    # consecutive steps and different instances are independent draws
    both = float((f0 * f1).mean())
    assert abs(both - 0.01) < 5 * (0.01 * 0.99 / cells) ** 0.5 + 1e-4
    rows = f0[:, 0].reshape(n, -1)
    assert float((rows[0] * rows[1]).mean()) < 0.03 and not torch.equal(rows[0], rows[1])
    # stateless: same (seed, step) -> same action; different seed -> different action
    again = env.random_action(123, 0, 0.1)
    assert torch.equal(again.words, a0.words)
    assert not torch.equal(env.random_action(124, 0, 0.1).words, a0.words)
    # per-column / per-row rates are flat (no structure from the packed layout)
    assert float(f0.mean(dim=(0, 1, 2)).min()) > 0.08 and float(f0.mean(dim=(0, 1, 3)).max()) < 0.12
    # parity: packed step == float step == oracle
    rng = np.random.default_rng(0)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    env2 = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                    fused_reductions=True)
    env2.reset()
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=8)
    ref.reset()
    for e in (env, env2):
        e.universe = torch.from_numpy(soup).float()[:, None]
    ref.universe = soup[:8].copy()
    for a in (a0, a1):
        o1 = env.step(a)[0]
        o2 = env2.step(a.to_float())[0]
        want = ref.step(a.to_float()[:8].cpu().numpy())[0]
        assert torch.equal(o1, o2)
        assert np.array_equal(o1[:8, 0].cpu().numpy().astype(np.uint8), want)
    assert int(env.action_count().sum()) == int(f1.sum())
    # all-ones packed action is the master reset, too
    ones = cb.PackedAction(env.random_action(1, 0, 1.0).words, env)
    assert float(ones.to_float().mean()) == 1.0
    assert float(env.step(ones)[0].abs().sum()) == 0.0 and env.step_number == 0


@pytest.mark.parametrize("size,win,n,rule", [(64, 32, 300, "B3/S23"), (128, 32, 70, "B368/S245"),
                                             (256, 64, 9, "B3/S23"), (96, 32, 5, "B36/S23"),
                                             (320, 64, 2, "B3/S23")])
def test_fused_random_agent_step_equals_materialised_action(size, win, n, rule):
    """env.step(RandomAction) draws the toggles INSIDE the step kernel (one launch for the
    batched shapes, generate + flags + step elsewhere); it must equal stepping with the same
    toggles materialised (random_action -> float32) and with the oracle."""
    cb = _carle()
    rng = np.random.default_rng(size + n)
    soup = (rng.random((n, size, size)) < 0.4).astype(np.uint8)
    envs = [cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                     fused_reductions=True) for _ in range(2)]
    ref = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win,
                         instances=n)
    ref.rules_from_string(rule)
    ref.reset()
    ref.universe = soup.copy()
    for e in envs:
        e.rules_from_string(rule)
        e.reset()
        e.universe = torch.from_numpy(soup).float()[:, None]
    agent = cb.DeviceRandomAgent(envs[0], toggle_rate=0.1, seed=99)        # lazy recipe
    for t in range(5):
        a = agent(None)
        if t == 3:                                    # batch-1 and a forced master reset
            a = cb.RandomAction(envs[0], 5, t, 1.0, 1)
        o0 = envs[0].step(a)[0]
        f = a.to_float()
        o1 = envs[1].step(f)[0]
        want = ref.step(f.cpu().numpy())[0]
        assert torch.equal(o0, o1), t
        assert np.array_equal(o0[:, 0].cpu().numpy().astype(np.uint8), want), t
        assert torch.equal(envs[0].last_reductions, envs[1].last_reductions), t
        assert envs[0].step_number == ref.step_number
