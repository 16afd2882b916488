"""numpy restatement of the reference CARLE step and mcl grid reductions.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA
path and the source of the CPU baseline.  It is *not* a fallback; the product
never imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (in the build container) and records its
outputs in ``tests/golden/*.npz`` / ``golden.json``; ``tests/test_oracle.py``
checks this restatement against every one of those vectors, the reference's
``spaceship_duck.rle -> spaceship_step.rle`` known-answer pair and the values
pinned by the reference's own ``tests/test_env.py``.

Every function cites the reference lines (relative to ``/root/reference``) it
follows.  Cells are held as ``uint8`` 0/1 in an array ``[N, H, W]``; the
reference holds the same values as float32 ``[N, 1, H, W]``.  All arithmetic on
the grid is on small exact integers, so dtype does not change results.
"""
from __future__ import annotations

import hashlib
import numpy as np

ALLOWED = "012345678"  # carle/env.py:57


def parse_rule_digits(text):
    """carle/env.py:62-78 — keep chars '0'..'8', dedupe, sort."""
    return sorted({int(ch) for ch in text if ch in ALLOWED})


def rules_from_string(text):
    """carle/env.py:80-85 — split on '/', part 0 -> birth, part 1 -> survive.

    Raises IndexError when there is no '/', like the reference."""
    parts = text.split("/")
    return parse_rule_digits(parts[0]), parse_rule_digits(parts[1])


def window_geometry(height, width, action_height, action_width):
    """carle/env.py:119-132 (set_action_padding) + :179 (ZeroPad2d use).

    Returns ``(action_width, action_height, row0, col0)`` after the reference's
    odd-size adjustment, where an action element ``[r, c]`` (dim 2 = r is checked
    against action_width, dim 3 = c against action_height, env.py:172-177) lands
    on universe cell ``[row0 + r, col0 + c]``.  ZeroPad2d pads the LAST dim with
    its first two numbers, so the 'height' padding is applied to columns and the
    'width' padding to rows (the reference's axis swap).  Raises ValueError when
    the padded action would not have the universe's shape (the reference then
    dies with a torch RuntimeError at the XOR, env.py:182)."""
    asym_w = (width - action_width) % 2
    asym_h = (height - action_height) % 2
    aw = action_width - (width % 2)
    ah = action_height - (height % 2)
    wp = (width - aw) // 2
    hp = (height - ah) // 2
    rows = aw + 2 * wp + asym_w      # dim 2 after padding
    cols = ah + 2 * hp + asym_h      # dim 3 after padding
    if rows != height or cols != width:
        raise ValueError(
            f"padded action is {rows}x{cols}, universe is {height}x{width}")
    return aw, ah, wp, hp


def neighbour_count(u):
    """carle/env.py:95-104, 219 — 3x3 Moore sum (centre excluded), circular in
    both axes.  ``u``: uint8 [N, H, W] -> uint8 counts 0..8."""
    total = np.zeros(u.shape, dtype=np.uint8)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                continue
            total += np.roll(u, shift=(dr, dc), axis=(1, 2))
    return total


def life_like_update(u, birth, survive):
    """carle/env.py:219-229 — next = (1-u)*[count in B] + u*[count in S].

    Empty birth or survive list -> TypeError (functools.reduce on an empty
    sequence, env.py:221-224)."""
    if len(birth) == 0 or len(survive) == 0:
        raise TypeError("reduce() of empty iterable with no initial value")
    cnt = neighbour_count(u)
    born = np.isin(cnt, np.asarray(birth, dtype=np.uint8))
    stay = np.isin(cnt, np.asarray(survive, dtype=np.uint8))
    return np.where(u != 0, stay, born).astype(np.uint8)


def digest(u):
    """SURVEY.md §8(c) digest convention: first 16 hex of
    sha256(packbits(universe as uint8, C order, bitorder big))."""
    return hashlib.sha256(np.packbits(np.asarray(u, dtype=np.uint8).ravel())
                          .tobytes()).hexdigest()[:16]


class OracleCARLE:
    """Restatement of ``class CARLE`` (carle/env.py:15-242), numpy only."""

    def __init__(self, width=256, height=256, action_width=64, action_height=64,
                 instances=1):
        # carle/env.py:21-59
        self.width, self.height = width, height
        self.instances = instances
        (self.action_width, self.action_height,
         self.row0, self.col0) = window_geometry(height, width, action_height,
                                                 action_width)
        self.birth = [3]
        self.survive = [2, 3]
        self.universe = None
        self.step_number = 0
        self.steps_since_action = 0

    def rules_from_string(self, text):
        self.birth, self.survive = rules_from_string(text)

    def reset(self):
        """carle/env.py:134-148 — all-dead universe; rules are NOT reset."""
        self.universe = np.zeros((self.instances, self.height, self.width),
                                 dtype=np.uint8)
        self.step_number = 0
        self.steps_since_action = 0
        return self.universe

    def _window_action(self, action):
        """carle/env.py:152-177 — coerce to 4-D, optional centre crop of a
        grid-sized action, dimension asserts."""
        a = np.asarray(action)
        while a.ndim < 4:
            a = a[None]
        if a.shape[3] > self.action_width and a.shape[1] < self.width:
            off_y = (self.width - self.action_width) // 2
            off_x = (self.height - self.action_height) // 2
            a = a[:, :, off_y:-off_y, off_x:-off_x]
        assert a.shape[2] == self.action_width, "action width is wrong"
        assert a.shape[3] == self.action_height, "action height is wrong"
        return a

    def apply_action(self, action):
        """carle/env.py:150-182 — zero-pad to the grid, logical XOR (any
        non-zero value toggles; batch-1 actions broadcast)."""
        a = self._window_action(action)
        toggles = (a[:, 0] != 0)
        r0, c0 = self.row0, self.col0
        win = self.universe[:, r0:r0 + self.action_width,
                            c0:c0 + self.action_height]
        self.universe = self.universe.copy()
        self.universe[:, r0:r0 + self.action_width,
                      c0:c0 + self.action_height] = \
            np.logical_xor(win != 0, toggles).astype(np.uint8)

    def step(self, action):
        """carle/env.py:188-242."""
        a = np.asarray(action, dtype=np.float32)
        if not a.sum():                                   # env.py:191,200
            self.steps_since_action += 1
        self.apply_action(a)                              # env.py:197/206
        if np.float32(a.mean(dtype=np.float32)) == np.float32(1.0):  # :208
            obs = self.reset()                            # env.py:216
        else:
            self.universe = life_like_update(self.universe, self.birth,
                                             self.survive)
            self.step_number += 1                         # env.py:230
            obs = self.universe
        reward = np.zeros((self.instances, 1), dtype=np.float32)   # :238
        done = np.zeros((self.instances, 1), dtype=np.float32)     # :239
        info = [{}] * self.instances                               # :240
        return obs, reward, done, info


# --------------------------------------------------------------------------
# mcl.py grid reductions
# --------------------------------------------------------------------------

def outside_window_mask(env):
    """carle/mcl.py:749-754 — 1 outside the action window, 0 inside
    (ones window zero-padded by env.action_padding, then 1 - that)."""
    m = np.ones((env.height, env.width), dtype=np.float32)
    m[env.row0:env.row0 + env.action_width,
      env.col0:env.col0 + env.action_height] = 0.0
    return m


def speed_sums(universe, mask):
    """carle/mcl.py:773-779 numerators — exact integers:
    live = sum(u); sh = sum(i * mask * u); sw = sum(j * mask * u)."""
    u = universe.astype(np.int64)
    h, w = u.shape[1:]
    mi = mask.astype(np.int64)
    live = u.sum(axis=(1, 2))
    sh = (u * (np.arange(h)[:, None] * mi)).sum(axis=(1, 2))
    sw = (u * (np.arange(w)[None, :] * mi)).sum(axis=(1, 2))
    return live, sh, sw


class OracleSpeedDetector:
    """carle/mcl.py:730-799."""

    def __init__(self, env):
        self.env = env
        self.mask = outside_window_mask(env)
        self.center_of_mass = None
        self.speed = None

    def reset(self):
        return self.env.reset()     # centre of mass persists, mcl.py:65-69

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        live, sh, sw = speed_sums(self.env.universe, self.mask)
        denom = live.astype(np.float32) + np.float32(1e-7)        # mcl.py:777
        com = np.stack([sh.astype(np.float32) / denom,
                        sw.astype(np.float32) / denom])           # (2, N)
        if self.center_of_mass is None:                           # mcl.py:784
            self.center_of_mass = com
        else:
            velocity = self.center_of_mass - com                  # mcl.py:787
            speed = np.sqrt(np.sum(np.square(velocity), dtype=np.float32),
                            dtype=np.float32)                     # mcl.py:789
            self.speed = speed
            self.center_of_mass = com
            reward = reward + speed                               # mcl.py:795
        self.live_cells = live
        return obs, reward, done, info


def corner_masks(height, width):
    """carle/mcl.py:206-217 (python slicing semantics kept, including the
    empty slices for ii < 4)."""
    reward_mask = np.zeros((height, width), dtype=np.float32)
    punish_mask = np.zeros((height, width), dtype=np.float32)
    reward_mask[:16, :16] = 1.0
    for ii in range(96):
        reward_mask[ii - 4:ii + 4, ii - 4:ii + 4] = 1.0
    punish_mask[-64:, -64:] = -1.0
    punish_mask[:64, -64:] = -1.0
    return reward_mask, punish_mask


class OracleCornerBonus:
    """carle/mcl.py:197-231."""

    def __init__(self, env):
        self.env = env
        self.reward_mask, self.punish_mask = corner_masks(env.height, env.width)

    def reset(self):
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        o = obs.astype(np.float32)
        reward = reward + (self.reward_mask * o).sum(-1).sum(-1)[:, None]
        reward = reward + (self.punish_mask * o).sum(-1).sum(-1)[:, None]
        return obs, reward.astype(np.float32), done, info


class OraclePufferDetector:
    """carle/mcl.py:804-853."""

    def __init__(self, env, growth_threshold=512):
        self.env = env
        self.cells = []
        self.growth_threshold = growth_threshold

    def reset(self):
        return self.env.reset()

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self.live_cells = float(self.env.universe.sum(dtype=np.int64))  # :832
        if not np.asarray(action).sum():                                # :835
            self.cells.append(self.live_cells)
            if len(self.cells) > self.growth_threshold:
                slope = self.cells[-1] - self.cells[0]
                self.cells.pop(0)
                if slope > 0.01:
                    reward = reward + 1
        else:
            self.cells = []
        return obs, reward, done, info


def parsimony(reward, action):
    """carle/mcl.py:102-103 — 100*reward / max(sum_{1,2,3} action, 100).
    (N,1)/(N,) broadcasts to (N,N) for N>1, as in the reference."""
    s = np.asarray(action, dtype=np.float32).sum(axis=(1, 2, 3))
    return np.float32(100.0) * reward / np.maximum(s, np.float32(100.0))


def random_agent_action(rng_uniform, toggle_rate=0.1):
    """carle/agents.py:38-40 — 1.0 * (uniform <= 0.1)."""
    return (rng_uniform <= toggle_rate).astype(np.float32)


# ---------------------------------------------------------------------------------------
# MorphoBonus (carle/mcl.py:107-195)
# ---------------------------------------------------------------------------------------
def morpho_patterns(grids):
    """carle/mcl.py:146-172 (add_rle_pattern).  ``grids``: iterable of 0/1 arrays [H, W] as
    ``rle_to_grid`` returns them.  Each is padded (left 1, right 1, top 2, bottom 1), cut to its
    top-left 8x8, dead cells -> -1, live cells -> 15 / (number of live cells), and contributes six
    orientations: itself, flipped vertically, flipped horizontally, transposed + each flip,
    transposed.  Returns float32 [6 * len(grids), 8, 8]."""
    out = []
    for g in grids:
        g = np.asarray(g, dtype=np.float32)
        padded = np.pad(g, ((2, 1), (1, 1)))[:8, :8].copy()
        padded[padded == 0] = -1.0
        live = padded == 1
        padded[live] *= np.float32(15.0) / np.float32(live.sum())
        t = padded.T
        out += [padded, padded[::-1, :], padded[:, ::-1], t[::-1, :], t[:, ::-1], t]
    return np.stack(out).astype(np.float32)


def morpho_scores(grid, patterns):
    """carle/mcl.py:176-185 — valid (un-padded) cross-correlation of ``grid`` [N, H, W] with every
    pattern, then max and min over patterns and positions.  Returns (max [N], min [N]) float32."""
    grid = np.asarray(grid, dtype=np.float32)
    n, h, w = grid.shape
    windows = np.lib.stride_tricks.sliding_window_view(grid, (8, 8), axis=(1, 2))   # [N, H-7, W-7, 8, 8]
    conv = np.einsum("nijrc,prc->npij", windows, patterns.astype(np.float32), optimize=True)
    return (conv.reshape(n, -1).max(axis=1).astype(np.float32),
            conv.reshape(n, -1).min(axis=1).astype(np.float32))


class OracleMorphoBonus:
    """carle/mcl.py:107-195.  ``action`` must broadcast against the universe in ``abs(universe -
    action)`` (mcl.py:176): grid-sized, or window-sized when the window is the whole grid."""

    def __init__(self, env, pattern_grids):
        self.env, self.inner_env = env, env
        self.reward_scale = 1.0
        self.target_patterns = morpho_patterns(pattern_grids)

    def reset(self):
        return self.env.reset()

    def step(self, action):
        a = np.asarray(action, dtype=np.float32)
        a = a.reshape((-1,) + a.shape[-2:])
        my_grid = np.abs(self.inner_env.universe.astype(np.float32) - a)        # mcl.py:176
        mx, mn = morpho_scores(my_grid, self.target_patterns)                  # :177, 181-182
        obs, reward, done, info = self.env.step(action)                        # :179
        return obs, reward + self.reward_scale * (mx + mn)[:, None], done, info   # :185


# ---------------------------------------------------------------------------------------
# RLE text (carle/env.py:260-328 rle_to_grid, 408-464 get_rle)
# ---------------------------------------------------------------------------------------
def get_rle(cells, birth, survive, height, width, instance_id, step_number, action=False):
    """carle/env.py:408-464, including its quirk: the last partial line (< 70 characters) is
    never flushed before the closing '!' (env.py:453-455)."""
    cells = np.asarray(cells)
    rle = "#C exp_id={} \n".format(instance_id)
    rle += "#C step={} ({}) \n".format(step_number, "action" if action else "universe")
    rle += "x = 0, y = 0, rule = B" + "".join(str(b) for b in birth) + "/S" + \
        "".join(str(s) for s in survive) + ":T{}, {}\n".format(height, width)
    line = ""
    for row in cells:
        jj, state, run = 0, row[0], 1
        while jj < len(row) - 1:
            jj += 1
            if row[jj] == state:
                run += 1
            else:
                line += str(run) + "bo"[int(state)]
                if len(line) > 69:
                    rle += line + "\n"
                    line = ""
                state, run = row[jj], 1
        line += str(run) + "bo"[int(state)] + "$"
        if len(line) > 69:
            rle += line + "\n"
            line = ""
    return rle + "!"


def rle_to_grid(rle, height, width):
    """carle/env.py:260-328 — token walk: counts accumulate until b / o / $ / !; newlines are
    skipped (a count may straddle one)."""
    grid = np.zeros((height, width), dtype=np.uint8)
    ii = jj = 0
    temp = ""
    for ch in rle:
        if ch == "\n":
            continue
        temp += ch
        tag = ch.lower()
        if tag in "bo":
            run = 1 if len(temp) == 1 else int(temp[:-1])
            if tag == "o":
                grid[ii, jj:jj + run] = 1
            jj += run
            temp = ""
        elif ch == "$":
            ii += int(temp[:-1]) if len(temp) > 1 else 1
            jj = 0
            temp = ""
        elif ch == "!":
            temp = ""
    return grid
