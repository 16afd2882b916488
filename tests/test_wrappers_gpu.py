"""SURVEY §8(f) rows on the GPU: MorphoBonus as a packed template match, PufferDetector's window
kept on the device, the RLE codec behind the CARLE methods -- against the fixtures generated from
the reference (tests/golden/make_golden.py) and against the oracle on longer runs."""
import numpy as np
import pytest
import torch

import _cases as cs
from _golden import by_kind, load, unbits
from oracle import carle_oracle as oc
from test_parity_gpu import ADAPTERS, CudaAdapter

pytestmark = pytest.mark.gpu


def _carle():
    import carle_b200
    return carle_b200


# ------------------------------------------------------------- MorphoBonus (mcl.py:107-195) ----
@pytest.mark.parametrize("make", ADAPTERS)
@pytest.mark.parametrize("name", by_kind("morpho"))
def test_morpho_bonus_golden(name, make):
    cs.check_morpho(name, make)


@pytest.mark.parametrize("size,win,n", [(64, 32, 5), (96, 32, 3), (256, 64, 2), (40, 16, 4)])
def test_morpho_match_equals_oracle_convolution(size, win, n):
    """Extrema of the template response on random soups, with and without a toggle plane, window-
    sized actions (zero-padded) included; widths that are no multiple of 32 and multi-word rows."""
    cb = _carle()
    rng = np.random.default_rng(size + n)
    inner = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                     device="cuda", obs_mode="packed")
    env = cb.MorphoBonus(inner)
    env.reset()
    soup = (rng.random((n, size, size)) < 0.3).astype(np.uint8)
    inner.universe = torch.from_numpy(soup)[:, None]
    patterns = env.target_patterns.cpu().numpy()[:, 0]
    assert np.array_equal(patterns, oc.morpho_patterns(cs.glider_grids(size)))
    mx, mn = env.match(None)
    want_mx, want_mn = oc.morpho_scores(soup, patterns)
    assert np.array_equal(mx.cpu().numpy(), want_mx) and np.array_equal(mn.cpu().numpy(), want_mn)
    # grid-sized action: the whole plane toggles (mcl.py:176)
    plane = (rng.random((n, 1, size, size)) < 0.1).astype(np.float32)
    mx, mn = env.match(torch.from_numpy(plane))
    want_mx, want_mn = oc.morpho_scores(soup ^ plane[:, 0].astype(np.uint8), patterns)
    assert np.array_equal(mx.cpu().numpy(), want_mx) and np.array_equal(mn.cpu().numpy(), want_mn)
    # one plane for the whole batch
    mx, mn = env.match(torch.from_numpy(plane[:1]))
    want_mx, want_mn = oc.morpho_scores(soup ^ plane[:1, 0].astype(np.uint8), patterns)
    assert np.array_equal(mx.cpu().numpy(), want_mx) and np.array_equal(mn.cpu().numpy(), want_mn)
    # window-sized action = the zero-padded window
    a = (rng.random((n, 1, win, win)) < 0.2).astype(np.float32)
    padded = np.zeros((n, size, size), dtype=np.uint8)
    r0 = (size - win) // 2
    padded[:, r0:r0 + win, r0:r0 + win] = a[:, 0].astype(np.uint8)
    mx, mn = env.match(torch.from_numpy(a))
    want_mx, want_mn = oc.morpho_scores(soup ^ padded, patterns)
    assert np.array_equal(mx.cpu().numpy(), want_mx) and np.array_equal(mn.cpu().numpy(), want_mn)
    # a non-integer weight (7 live cells: 15 / 7): float32 rounding of the conv2d sum only
    env.add_rle_text("3o$obo$3o!")
    patterns = env.target_patterns.cpu().numpy()[:, 0]
    mx, mn = env.match(None)
    want_mx, want_mn = oc.morpho_scores(soup, patterns)
    np.testing.assert_allclose(mx.cpu().numpy(), want_mx, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(mn.cpu().numpy(), want_mn, rtol=1e-5, atol=1e-5)


def test_morpho_reset_seeds_the_universe():
    cb = _carle()
    env = cb.MorphoBonus(cb.CARLE(instances=3, height=64, width=64, action_width=32, action_height=32,
                                  device="cuda"))
    torch.manual_seed(0)
    obs = env.reset()
    frac = float(obs.mean().item())
    assert 0.001 < frac < 0.012                       # ~0.5 % seed cells (mcl.py:191)
    assert torch.equal(obs, env.inner_env.universe)


# --------------------------------------------------------- PufferDetector (mcl.py:804-853) ----
def test_puffer_window_on_the_device_equals_oracle_over_a_long_run():
    """Window wrap-around, an action that empties the window, growth and decay phases; rewards,
    the window's contents and the live total after every step -- and no host read in step()."""
    cb = _carle()
    n, size, win, threshold = 3, 64, 32, 5
    rng = np.random.default_rng(11)
    inner = cb.CARLE(instances=n, height=size, width=size, action_width=win, action_height=win,
                     device="cuda", obs_mode="packed")
    env = cb.PufferDetector(inner)
    env.growth_threshold = threshold
    inner.rules_from_string("B3/S012345678")           # growing rule (mcl.py:806-808)
    o_inner = oc.OracleCARLE(width=size, height=size, action_width=win, action_height=win, instances=n)
    o_inner.rules_from_string("B3/S012345678")
    o_env = oc.OraclePufferDetector(o_inner, growth_threshold=threshold)
    env.reset()
    o_env.reset()
    soup = (rng.random((n, size, size)) < 0.04).astype(np.uint8)
    inner.universe = torch.from_numpy(soup)[:, None]
    o_inner.universe = soup.copy()
    rewards = []
    for t in range(40):
        if t in (0, 17, 18):
            a = (rng.random((n, 1, win, win)) < 0.05).astype(np.float32)
        else:
            a = np.zeros((n, 1, win, win), dtype=np.float32)
        if t == 30:                                     # shrink: Life thins the soup out again
            inner.rules_from_string("B3/S23")
            o_inner.rules_from_string("B3/S23")
        obs, r, _, _ = env.step(torch.from_numpy(a).cuda())
        _, want, _, _ = o_env.step(a)
        rewards.append((r, np.asarray(want, dtype=np.float32)))
    for t, (r, want) in enumerate(rewards):
        assert np.array_equal(r.cpu().numpy(), np.broadcast_to(want, (n, 1))), t
    assert env.cells == [float(c) for c in o_env.cells]
    assert env.live_cells == float(o_inner.universe.sum())
    assert sum(float(w.max()) for _, w in rewards) > 0          # the bonus did fire


@pytest.mark.parametrize("make", ADAPTERS)
def test_puffer_golden(make):
    cs.check_wrapper("puffer_64", make)


# ------------------------------------------------------------------- RLE behind the env ----
@pytest.mark.parametrize("name", by_kind("rle"))
def test_env_rle_methods_match_the_reference(name, tmp_path):
    """get_rle / rle_to_grid / load_universe / log_universe of carle_b200.CARLE on a device
    universe vs the reference's text (byte for byte up to its dropped last line)."""
    cb = _carle()
    meta, z = load(name)
    size = meta["size"]
    cells = unbits(z["cells"], size)
    env = cb.CARLE(instances=2, height=size, width=size, action_width=size // 2, action_height=size // 2,
                   device="cuda", logging=True)
    env.rules_from_string(meta["rule"])
    env.reset()
    env.instance_id = meta["instance_id"]
    env.universe = torch.from_numpy(np.stack([np.zeros_like(cells), cells]))[:, None].float()
    env.step_number = meta["step_number"]
    ref_text = bytes(z["text"]).decode("ascii")
    assert env.get_rle(env.universe[1, 0], keep_tail=False) == ref_text
    ours = env.get_rle(env.universe[1, 0])
    assert ours.startswith(ref_text[:-1]) and ours.endswith("!")
    assert np.array_equal(env.rle_to_grid(ref_text.split("\n", 3)[3]).numpy(), unbits(z["decoded"], size))
    # own text -> file -> load_universe (reads the header get_rle writes) -> same cells, packed state
    path = tmp_path / "u.rle"
    path.write_text(ours)
    env.load_universe(str(path), universe_index=0)
    assert np.array_equal(env.universe[0, 0].cpu().numpy(), cells)
    # log_universe encodes straight from the packed words
    env.action = torch.zeros(2, 1, size // 2, size // 2)
    env.log_universe(universe_index=1)
    assert env.log[-1][1] == ours
