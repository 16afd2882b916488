"""Pins the numpy oracle (oracle/carle_oracle.py) to the reference.

Every fixture under tests/golden/ was produced by running the unmodified
reference (tests/golden/make_golden.py); this file replays them through the
restatement.  CPU only."""
import numpy as np
import pytest

import _cases as cs
from _golden import MANIFEST, by_kind, load, unbits
from oracle import carle_oracle as oc

make = cs.OracleAdapter


def test_rule_parser_matches_reference_test_env():
    # reference tests/test_env.py:17-39
    assert oc.parse_rule_digits("asdfasdfB0357*!@#!@$%") == [0, 3, 5, 7]
    assert oc.parse_rule_digits("S2468") == [2, 4, 6, 8]
    assert oc.rules_from_string("B0357/S2468") == ([0, 3, 5, 7], [2, 4, 6, 8])
    assert MANIFEST["rule_parser"] == {"birth": [0, 3, 5, 7],
                                       "survive": [2, 4, 6, 8]}
    b, s = oc.rules_from_string("23/3")
    assert {"birth": b, "survive": s} == MANIFEST["rule_parser_23_3"]
    with pytest.raises(IndexError):
        oc.rules_from_string("B3S23")


def test_reset_semantics_match_reference_test_env():
    # reference tests/test_env.py:42-67 on the default 256x256 / 64x64 env
    env = oc.OracleCARLE()
    reset_obs = env.reset().copy()
    action = np.ones((1, 1, 64, 64), dtype=np.float32)
    toggle_obs = env.step(action)[0].copy()
    action[:, :, 0:10, 0:10] = 0.0
    normal_obs = env.step(action)[0]
    assert toggle_obs.mean() == 0.0 and reset_obs.mean() == 0.0
    assert np.array_equal(reset_obs, toggle_obs)
    assert not np.array_equal(toggle_obs, normal_obs)


def test_geometry_rejections():
    with pytest.raises(ValueError):
        oc.window_geometry(65, 65, 64, 64)      # odd grid
    with pytest.raises(ValueError):
        oc.window_geometry(64, 128, 32, 32)     # non-square grid
    assert oc.window_geometry(64, 64, 31, 31) == (31, 31, 16, 16)
    assert oc.window_geometry(64, 64, 16, 32) == (32, 16, 16, 24)


def test_empty_rule_raises_typeerror():
    env = oc.OracleCARLE(width=16, height=16, action_width=8, action_height=8)
    env.reset()
    env.survive = []
    with pytest.raises(TypeError):
        env.step(np.zeros((1, 1, 8, 8), dtype=np.float32))


@pytest.mark.parametrize("name", by_kind("rollout"))
def test_rollout_digests(name):
    cs.check_rollout(name, make)


def test_freerun_g5():
    cs.check_freerun("g5", make)


@pytest.mark.parametrize("name", by_kind("sweep"))
def test_sweep(name):
    cs.check_sweep(name, make)


def test_action_values_and_broadcast():
    cs.check_action_values(make)


def test_master_reset_sequence():
    cs.check_master_reset(make)


@pytest.mark.parametrize("name", by_kind("master_reset_mean"))
def test_master_reset_on_mean_of_nonbinary_actions(name):
    cs.check_master_reset_mean(name, make)


def test_grid_sized_action_crop():
    cs.check_grid_sized_action(make)


def test_nonsquare_window():
    cs.check_nonsquare_window(make)


def test_action_placement():
    cs.check_placement(make)


def test_spaceship_known_answer():
    cs.check_spaceship(make)


@pytest.mark.parametrize("name", by_kind("wrapper"))
def test_wrappers(name):
    cs.check_wrapper(name, make)


@pytest.mark.parametrize("name", by_kind("parsimony"))
def test_parsimony(name):
    cs.check_parsimony(name, make)


@pytest.mark.parametrize("name", by_kind("morpho"))
def test_morpho_bonus(name):
    cs.check_morpho(name, make)


@pytest.mark.parametrize("name", by_kind("rle"))
def test_rle_text_restatement(name):
    """The oracle's plain-Python restatement reproduces the reference's text byte for byte
    (header and dropped tail included) and its decoded grid."""
    meta, z = load(name)
    size = meta["size"]
    cells = unbits(z["cells"], size)
    b, s = oc.rules_from_string(meta["rule"])
    text = oc.get_rle(cells, b, s, size, size, meta["instance_id"], meta["step_number"])
    assert text == bytes(z["text"]).decode("ascii")
    assert np.array_equal(oc.rle_to_grid(text.split("\n", 3)[3], size, size), unbits(z["decoded"], size))


# ---- the torch-CPU port timed as the CPU baseline is pinned to the same vectors ----
class _PortAdapter:
    def __init__(self, n, size, aw, ah, rule, wrapper=None, tweak=None):
        import torch
        from oracle.torch_port import TorchPortCARLE, TorchPortSpeedDetector
        assert wrapper in (None, "SpeedDetector")
        self.torch = torch
        self.env = TorchPortCARLE(width=size, height=size, action_width=aw,
                                  action_height=ah, instances=n)
        self.env.birth, self.env.survive = oc.rules_from_string(rule)
        self.outer = TorchPortSpeedDetector(self.env) if wrapper else self.env

    def reset(self):
        self.outer.reset()

    def set_universe(self, u):
        self.env.universe = self.torch.from_numpy(np.array(u, dtype=np.float32))[:, None]

    def get_universe(self):
        return self.env.universe[:, 0].numpy().astype(np.uint8)

    def step(self, action):
        obs, reward, done, info = self.outer.step(self.torch.from_numpy(
            np.asarray(action, dtype=np.float32)))
        return obs[:, 0].numpy().astype(np.uint8), reward.numpy()

    def apply_action(self, action):
        self.env.apply_action(self.torch.from_numpy(np.asarray(action, dtype=np.float32)))

    @property
    def step_number(self):
        return self.env.step_number

    @property
    def steps_since_action(self):
        return self.env.steps_since_action


@pytest.mark.parametrize("name", ["g1", "g2", "g4"])
def test_torch_port_rollouts(name):
    cs.check_rollout(name, _PortAdapter)


@pytest.mark.parametrize("name", ["g3", "speed_64", "speed_128", "speed_256"])
def test_torch_port_speed_detector(name):
    # the wrapper of the headline workload, as timed by bench.py's CPU arm
    if name == "g3":
        cs.check_rollout(name, _PortAdapter)
    else:
        cs.check_wrapper(name, _PortAdapter)


@pytest.mark.parametrize("name", by_kind("sweep")[:8])
def test_torch_port_sweep(name):
    cs.check_sweep(name, _PortAdapter)


def test_torch_port_master_reset_and_values():
    cs.check_master_reset(_PortAdapter)
    for name in by_kind("master_reset_mean"):
        cs.check_master_reset_mean(name, _PortAdapter)
    cs.check_action_values(_PortAdapter)
    cs.check_freerun("g5", _PortAdapter)
