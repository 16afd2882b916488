"""Golden-case runners, written against a tiny adapter interface so the same
checks pin the numpy oracle (CPU tests) and the CUDA path (GPU tests).

Adapter protocol (see OracleAdapter below / CudaAdapter in test_parity_gpu.py):
    make(n, size, aw, ah, rule, wrapper=None, tweak=None) -> adapter
    adapter.reset()
    adapter.set_universe(uint8 [N,H,W])
    adapter.get_universe() -> uint8 [N,H,W]
    adapter.step(action float32 [B,1,aw,ah]) -> (obs uint8 [N,H,W], reward ndarray)
    adapter.apply_action(action)       (no generation)
    adapter.step_number -> int
"""
import numpy as np

from _golden import MANIFEST, load, unbits, action_from_bits
from oracle import carle_oracle as oc


class OracleAdapter:
    def __init__(self, n, size, aw, ah, rule, wrapper=None, tweak=None):
        self.inner = oc.OracleCARLE(width=size, height=size, action_width=aw,
                                    action_height=ah, instances=n)
        self.inner.rules_from_string(rule)
        self.env = self.inner
        if wrapper == "SpeedDetector":
            self.env = oc.OracleSpeedDetector(self.inner)
        elif wrapper == "CornerBonus":
            self.env = oc.OracleCornerBonus(self.inner)
        elif wrapper == "PufferDetector":
            self.env = oc.OraclePufferDetector(self.inner)
            if tweak:
                self.env.growth_threshold = tweak["growth_threshold"]
        elif wrapper == "Parsimony(Corner)":
            self.env = _OracleParsimonyCorner(self.inner)
        elif wrapper == "MorphoBonus":
            self.env = oc.OracleMorphoBonus(self.inner, glider_grids(size))
        elif wrapper is not None:
            raise KeyError(wrapper)

    def reset(self):
        self.env.reset()

    def set_universe(self, u):
        self.inner.universe = np.array(u, dtype=np.uint8)

    def get_universe(self):
        return np.array(self.inner.universe)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return np.array(obs), np.asarray(reward)

    def apply_action(self, action):
        self.inner.apply_action(action)

    @property
    def step_number(self):
        return self.inner.step_number

    @property
    def steps_since_action(self):
        return self.inner.steps_since_action


GLIDER_PHASES = (np.array([[0, 1, 0], [0, 0, 1], [1, 1, 1]]), np.array([[1, 0, 1], [0, 1, 1], [0, 1, 0]]))


def glider_grids(size):
    """The two glider phases as ``rle_to_grid`` returns them (top-left of a size x size grid)."""
    out = []
    for g in GLIDER_PHASES:
        full = np.zeros((size, size), dtype=np.uint8)
        full[:3, :3] = g
        out.append(full)
    return out


class _OracleParsimonyCorner:
    def __init__(self, inner):
        self.corner = oc.OracleCornerBonus(inner)

    def reset(self):
        return self.corner.reset()

    def step(self, action):
        obs, reward, done, info = self.corner.step(action)
        return obs, oc.parsimony(reward, action), done, info


def check_rollout(name, make):
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    env = make(n, size, win, win, meta["rule"], wrapper=meta.get("wrapper"))
    env.reset()
    if "init" in z.files:
        env.set_universe(unbits(z["init"], size))
    rewards = []
    for t in range(meta["steps"]):
        obs, r = env.step(action_from_bits(z["actions"][t], win))
        assert int(obs.sum()) == meta["pops"][t], (name, t)
        rewards.append(r)
    assert np.array_equal(obs, unbits(z["final"], size)), name
    assert oc.digest(obs) == meta["digest"], name
    if "rewards" in z.files:
        got = np.stack([np.broadcast_to(np.asarray(r, dtype=np.float32),
                                        z["rewards"][0].shape) for r in rewards])
        # live / Sh / Sw and the per-instance divide are exact; only the final
        # sqrt(sum(v^2)) over 2N values is summation-order dependent (f32).
        np.testing.assert_allclose(got, z["rewards"], rtol=2e-6, atol=1e-6)


def check_freerun(name, make):
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    env = make(n, size, win, win, meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], size))
    zero = np.zeros((n, 1, win, win), dtype=np.float32)
    for gen in range(1, 17):
        obs, _ = env.step(zero)
        if str(gen) in meta["gens"]:
            assert np.array_equal(obs, unbits(z[f"gen{gen}"], size)), (name, gen)
            assert oc.digest(obs) == meta["gens"][str(gen)]["digest"]
            assert int(obs.sum()) == meta["gens"][str(gen)]["pop"]


def check_sweep(name, make):
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    env = make(n, size, win, win, meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], size))
    for t in range(meta["steps"]):
        obs, _ = env.step(action_from_bits(z["actions"][t], win))
        assert np.array_equal(obs, unbits(z["states"][t], size)), (name, t)
    assert oc.digest(obs) == meta["digest"]


def check_action_values(make):
    meta, z = load("action_values")
    env = make(meta["n"], meta["size"], meta["win"], meta["win"], meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], meta["size"]))
    obs, _ = env.step(z["action"])          # batch-1, values 0.5/-2/7/1e-30
    assert np.array_equal(obs, unbits(z["final"], meta["size"]))


def check_master_reset(make):
    meta, z = load("master_reset")
    env = make(meta["n"], meta["size"], meta["win"], meta["win"], meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], meta["size"]))
    for t in range(z["actions"].shape[0]):
        obs, _ = env.step(z["actions"][t])
        assert np.array_equal(obs, unbits(z["states"][t], meta["size"])), t
        assert env.step_number == meta["step_numbers"][t], t


def check_master_reset_mean(name, make):
    """Actions whose elements are neither 0 nor 1: the reference resets on ``mean == 1.0`` and
    counts a step as action-free on ``sum == 0`` (env.py:191, 208)."""
    meta, z = load(name)
    env = make(meta["n"], meta["size"], meta["win"], meta["win"], meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], meta["size"]))
    for t in range(z["actions"].shape[0]):
        obs, _ = env.step(z["actions"][t])
        assert np.array_equal(obs, unbits(z["states"][t], meta["size"])), (name, t)
        assert env.step_number == meta["step_numbers"][t], (name, t)
        assert env.steps_since_action == meta["steps_since_action"][t], (name, t)


def check_grid_sized_action(make):
    meta, z = load("grid_sized_action")
    env = make(meta["n"], meta["size"], meta["win"], meta["win"], meta["rule"])
    env.reset()
    obs, _ = env.step(z["action"])
    assert np.array_equal(obs, unbits(z["final"], meta["size"]))


def check_nonsquare_window(make):
    meta, z = load("nonsquare_window")
    env = make(meta["n"], meta["size"], meta["aw"], meta["ah"], meta["rule"])
    env.reset()
    env.set_universe(unbits(z["init"], meta["size"]))
    obs, _ = env.step(z["action"])
    assert np.array_equal(obs, unbits(z["final"], meta["size"]))


def check_placement(make):
    for p in MANIFEST["placement_probes"]:
        env = make(1, p["size"], p["win"], p["win"], "B3/S23")
        env.reset()
        a = np.zeros((1, 1, p["win"], p["win"]), dtype=np.float32)
        a[0, 0, p["r"], p["c"]] = 1.0
        env.apply_action(a)
        u = env.get_universe()
        assert np.argwhere(u[0]).tolist() == [p["cell"]], p


def check_spaceship(make):
    meta, z = load("spaceship")
    env = make(1, 16, 8, 8, "B3/S23")
    env.reset()
    env.set_universe(unbits(z["duck"], 16)[None] if z["duck"].ndim == 2
                     else unbits(z["duck"], 16))
    obs, _ = env.step(np.zeros((1, 1, 8, 8), dtype=np.float32))
    assert np.array_equal(obs, unbits(z["step"], 16))


def check_wrapper(name, make):
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    tweak = {"growth_threshold": 4} if meta["wrapper"] == "PufferDetector" else None
    env = make(n, size, win, win, meta["rule"], wrapper=meta["wrapper"],
               tweak=tweak)
    env.reset()
    env.set_universe(unbits(z["init"], size))
    for t in range(meta["steps"]):
        obs, r = env.step(action_from_bits(z["actions"][t], win))
        want = z["rewards"][t]
        got = np.broadcast_to(np.asarray(r, dtype=np.float32), want.shape)
        if meta["wrapper"] == "SpeedDetector":
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6,
                                       err_msg=f"{name} step {t}")
        else:                       # integer-valued sums: exact
            assert np.array_equal(got, want), (name, t, got, want)
    assert np.array_equal(obs, unbits(z["final"], size))


def check_parsimony(name, make):
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    env = make(n, size, win, win, meta["rule"], wrapper="Parsimony(Corner)")
    env.reset()
    env.set_universe(unbits(z["init"], size))
    for t in range(meta["steps"]):
        obs, r = env.step(action_from_bits(z["actions"][t], win))
        want = z["rewards"][t]
        assert list(np.asarray(r).shape) == meta["reward_shape"]
        np.testing.assert_allclose(np.asarray(r, dtype=np.float32), want,
                                   rtol=1e-6, atol=0)
    assert np.array_equal(obs, unbits(z["final"], size))


def check_morpho(name, make):
    """MorphoBonus rewards (reference mcl.py:174-185) -- integers for the glider templates, so
    the comparison is exact -- and the template tensor itself."""
    meta, z = load(name)
    n, size, win = meta["n"], meta["size"], meta["win"]
    env = make(n, size, win, win, meta["rule"], wrapper="MorphoBonus")
    patterns = np.asarray(env.env.target_patterns if not hasattr(env.env.target_patterns, "cpu")
                          else env.env.target_patterns.cpu().numpy(), dtype=np.float32)
    assert np.array_equal(patterns.reshape(-1, 8, 8), z["patterns"].reshape(-1, 8, 8))
    env.reset()
    env.set_universe(unbits(z["init"], size))
    a_size = size if meta["grid_sized"] else win
    for t in range(meta["steps"]):
        obs, r = env.step(action_from_bits(z["actions"][t], a_size))
        assert np.array_equal(np.asarray(r, dtype=np.float32), z["rewards"][t]), (name, t, r, z["rewards"][t])
    assert np.array_equal(obs, unbits(z["final"], size))


def check_rle_codec(name, encode, decode):
    """``encode(cells uint8 [H,W], keep_tail) -> body text``, ``decode(text, H, W) -> uint8 [H,W]``
    against the reference's own get_rle text and rle_to_grid grid (env.py:408-464, 260-328)."""
    meta, z = load(name)
    size = meta["size"]
    cells = unbits(z["cells"], size)
    ref_text = bytes(z["text"]).decode("ascii")
    ref_body = ref_text.split("\n", 3)[3]
    # byte-identical to what the reference emits (it drops the last partial line) ...
    assert encode(cells, False) == ref_body, name
    ref_action_body = bytes(z["action_text"]).decode("ascii").split("\n", 3)[3]
    assert encode(cells[:size // 2, :size // 2], False) == ref_action_body, name
    # ... what the reference decodes from its own text is what we decode from it ...
    assert np.array_equal(decode(ref_body, size, size), unbits(z["decoded"], size)), name
    # ... and with the tail kept the text round-trips to the full grid and extends the reference's
    full = encode(cells, True)
    assert full.startswith(ref_body[:-1]) and full.endswith("!")
    assert np.array_equal(decode(full, size, size), cells), name
