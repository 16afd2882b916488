#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    python tests/golden/make_golden.py

It imports ``carle`` from ``/root/reference`` with the two non-invasive shims of
SURVEY.md §8(c) (stub modules for matplotlib/skimage, which the hot path never
uses, and ``set_neighborhood`` executed under ``torch.no_grad()`` because the
reference targets torch 1.5), drives the reference's own ``CARLE`` /
``SpeedDetector`` / ``CornerBonus`` / ``PufferDetector`` / ``ParsimonyBonus`` on
its torch CPU path, and records inputs and outputs as small ``.npz`` fixtures
plus ``golden.json`` (manifest, digests).  Nothing from the reference's sources
is copied; only its *outputs* are stored.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.path.insert(0, REF)
    import carle.env as ref_env
    import carle.mcl as ref_mcl
    orig = ref_env.CARLE.set_neighborhood

    def patched(self):
        with torch.no_grad():
            orig(self)
    ref_env.CARLE.set_neighborhood = patched
    return ref_env, ref_mcl


def digest(u):
    return hashlib.sha256(np.packbits(np.asarray(u, dtype=np.uint8).ravel())
                          .tobytes()).hexdigest()[:16]


def bits(t):
    """torch float [N,1,H,W] (0/1) -> packed uint8 [N,H,ceil(W/8)], LSB-first."""
    a = (t.detach().cpu().numpy()[:, 0] != 0).astype(np.uint8)
    return np.packbits(a, axis=-1, bitorder="little")


def main():
    ref_env, ref_mcl = import_reference()
    CARLE = ref_env.CARLE
    manifest = {"torch": torch.__version__, "cases": [], "digests": {}}

    def save(name, meta, **arrays):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        meta = dict(meta)
        meta["name"] = name
        manifest["cases"].append(meta)

    # ----- SURVEY §8(c) digests G1..G5 (re-derived live from the reference) ----
    def rollout(seed, n, size, win, steps, rule=None, wrapper=None, fill=None,
                zero_actions=False, record_every=False):
        torch.manual_seed(seed)
        env = CARLE(instances=n, height=size, width=size, action_height=win,
                    action_width=win, device="cpu")
        inner = env
        if wrapper is not None:
            env = wrapper(env)
        if rule is not None:
            inner.rules_from_string(rule)
        env.reset()
        init = None
        if fill is not None:
            inner.universe = (torch.rand(n, 1, size, size) < fill).float()
            init = bits(inner.universe)
        acts, pops, rewards, states = [], [], [], []
        obs = None
        for _ in range(steps):
            if zero_actions:
                a = torch.zeros(n, 1, win, win)
            else:
                a = 1.0 * (torch.rand(n, 1, win, win) <= 0.1)
            acts.append(np.packbits((a.numpy()[:, 0] != 0).astype(np.uint8),
                                    axis=-1, bitorder="little"))
            obs, r, d, info = env.step(a)
            pops.append(int(obs.sum().item()))
            rewards.append(r.detach().cpu().numpy().astype(np.float32).copy())
            if record_every:
                states.append(bits(obs))
        return dict(init=init, actions=np.stack(acts), pops=pops,
                    rewards=np.stack(rewards), final=bits(obs),
                    digest=digest(obs.numpy()), states=states,
                    obs=obs)

    g1 = rollout(0, 1, 64, 32, 256)
    save("g1", dict(kind="rollout", seed=0, n=1, size=64, win=32, steps=256,
                    rule="B3/S23", digest=g1["digest"], pops=g1["pops"]),
         actions=g1["actions"], final=g1["final"])
    g2 = rollout(1, 8, 128, 32, 64)
    save("g2", dict(kind="rollout", seed=1, n=8, size=128, win=32, steps=64,
                    rule="B3/S23", digest=g2["digest"], pops=g2["pops"]),
         actions=g2["actions"], final=g2["final"])
    g3 = rollout(2, 4, 256, 64, 32, rule="B368/S245",
                 wrapper=ref_mcl.SpeedDetector)
    save("g3", dict(kind="rollout", seed=2, n=4, size=256, win=64, steps=32,
                    rule="B368/S245", wrapper="SpeedDetector",
                    digest=g3["digest"], pops=g3["pops"],
                    reward_sum=float(np.sum(g3["rewards"][:, 0, 0],
                                            dtype=np.float64)),
                    last_reward_hex=float(g3["rewards"][-1, 0, 0]).hex()),
         actions=g3["actions"], final=g3["final"], rewards=g3["rewards"])
    g4 = rollout(3, 4, 64, 32, 32, fill=0.5)
    save("g4", dict(kind="rollout", seed=3, n=4, size=64, win=32, steps=32,
                    rule="B3/S23", digest=g4["digest"], pops=g4["pops"]),
         init=g4["init"], actions=g4["actions"], final=g4["final"])
    # G5: free run, digests at generations 1, 4, 8, 16
    torch.manual_seed(4)
    env = CARLE(instances=2, height=256, width=256, device="cpu")
    env.reset()
    env.universe = (torch.rand(2, 1, 256, 256) < 0.35).float()
    init5 = bits(env.universe)
    g5 = {}
    zero = torch.zeros(2, 1, 64, 64)
    snaps = {}
    for gen in range(1, 17):
        obs = env.step(zero)[0]
        if gen in (1, 4, 8, 16):
            g5[str(gen)] = dict(digest=digest(obs.numpy()),
                                pop=int(obs.sum().item()))
            snaps[f"gen{gen}"] = bits(obs)
    save("g5", dict(kind="freerun", seed=4, n=2, size=256, win=64, rule="B3/S23",
                    gens=g5), init=init5, **snaps)
    manifest["digests"] = dict(G1=g1["digest"], G2=g2["digest"], G3=g3["digest"],
                               G4=g4["digest"], G5=g5)

    # ----- shape / rule sweep: random soup + Bernoulli(0.1) actions, every state
    sweep = [
        # (size, win, n, rule, steps)
        (16, 8, 3, "B3/S23", 6),
        (32, 16, 5, "B36/S23", 6),
        (64, 32, 3, "B3678/S34678", 5),
        (64, 31, 2, "B3/S23", 5),            # odd window: extra pad bottom/right
        (96, 32, 2, "B368/S245", 5),
        (100, 50, 2, "B3/S23", 5),           # not a multiple of 32
        (128, 32, 3, "B0/S8", 4),            # B0 rule
        (128, 64, 2, "B1357/S1357", 4),
        (160, 64, 2, "B2/S0", 4),
        (192, 64, 1, "B3/S012345678", 4),
        (224, 64, 1, "B012345678/S012345678", 3),
        (256, 64, 2, "B368/S245", 4),
        (320, 64, 1, "B3/S23", 3),           # wider than the warp-resident path
        (512, 64, 1, "B36/S23", 3),
        (48, 48, 2, "B3/S23", 4),            # window == grid
        (6, 2, 2, "B3/S23", 4),              # tiny torus
    ]
    rng_rules = np.random.RandomState(7)
    for _ in range(6):                       # random B/S masks
        b = [k for k in range(9) if rng_rules.rand() < 0.4] or [3]
        s = [k for k in range(9) if rng_rules.rand() < 0.4] or [2]
        sweep.append((64, 32, 2, "B" + "".join(map(str, b)) + "/S" +
                      "".join(map(str, s)), 4))
    for idx, (size, win, n, rule, steps) in enumerate(sweep):
        r = rollout(100 + idx, n, size, win, steps, rule=rule, fill=0.4,
                    record_every=True)
        save(f"sweep{idx:02d}", dict(kind="sweep", seed=100 + idx, n=n, size=size,
                                     win=win, steps=steps, rule=rule,
                                     digest=r["digest"], pops=r["pops"]),
             init=r["init"], actions=r["actions"], states=np.stack(r["states"]))

    # ----- action semantics -------------------------------------------------
    # (a) non-binary values toggle; batch-1 action broadcasts over N instances
    torch.manual_seed(200)
    env = CARLE(instances=3, height=64, width=64, action_height=32,
                action_width=32, device="cpu")
    env.reset()
    env.universe = (torch.rand(3, 1, 64, 64) < 0.3).float()
    init = bits(env.universe)
    vals = torch.tensor([0.0, 0.5, -2.0, 7.0, 1.0, 0.0, 0.0, 1e-30])
    a = vals[torch.randint(0, len(vals), (1, 1, 32, 32))]
    obs = env.step(a)[0]
    save("action_values", dict(kind="action_values", n=3, size=64, win=32,
                               rule="B3/S23"),
         init=init, action=a.numpy(), final=bits(obs))

    # (b) master reset: all-ones fires, all-2.0 does not, one zero does not
    env = CARLE(instances=2, height=64, width=64, action_height=32,
                action_width=32, device="cpu")
    env.reset()
    torch.manual_seed(201)
    env.universe = (torch.rand(2, 1, 64, 64) < 0.3).float()
    init = bits(env.universe)
    seq, outs, stepnos = [], [], []
    a0 = 1.0 * (torch.rand(2, 1, 32, 32) <= 0.1)
    ones = torch.ones(2, 1, 32, 32)
    twos = 2.0 * torch.ones(2, 1, 32, 32)
    almost = torch.ones(2, 1, 32, 32)
    almost[1, 0, 5, 7] = 0.0
    for a in (a0, twos, almost, a0, ones, a0, a0):
        seq.append(a.numpy().copy())
        outs.append(bits(env.step(a)[0]))
        stepnos.append(env.step_number)
    save("master_reset", dict(kind="master_reset", n=2, size=64, win=32,
                              rule="B3/S23", step_numbers=stepnos),
         init=init, actions=np.stack(seq), states=np.stack(outs))

    # (b2) the reset predicate is `mean(action) == 1.0` and the no-action predicate is
    #      `not sum(action)` (env.py:191, 208): actions whose elements are neither 0 nor 1
    env = CARLE(instances=2, height=64, width=64, action_height=32,
                action_width=32, device="cpu")
    env.reset()
    torch.manual_seed(204)
    env.universe = (torch.rand(2, 1, 64, 64) < 0.3).float()
    init = bits(env.universe)
    ii, jj = torch.meshgrid(torch.arange(32), torch.arange(32), indexing="ij")
    checker = ((ii + jj) % 2).float()[None, None].repeat(2, 1, 1, 1)
    a0 = 1.0 * (torch.rand(2, 1, 32, 32) <= 0.1)
    zero_two = 2.0 * checker                           # mean exactly 1.0 -> reset
    half = 0.5 + checker                               # 0.5 / 1.5, mean 1.0 -> reset
    mixed = torch.ones(2, 1, 32, 32)
    mixed[1] = zero_two[1]                             # instance 0 all ones, instance 1 0/2
    cancel = 2.0 * checker - 1.0                       # +1 / -1: sum 0 (counts as "no action"),
    #                                                    every cell toggles, mean 0 -> no reset
    near = zero_two.clone()
    near[0, 0, 3, 4] = 0.0                             # mean just below 1.0 -> no reset
    over = zero_two.clone()
    over[1, 0, 2, 2] = 2.0                             # mean just above 1.0 -> no reset
    zeros = torch.zeros(2, 1, 32, 32)
    seq, outs, stepnos, since = [], [], [], []
    for a in (a0, zero_two, a0, half, a0, mixed, a0, cancel, near, over, zeros, a0):
        seq.append(a.numpy().copy())
        outs.append(bits(env.step(a)[0]))
        stepnos.append(env.step_number)
        since.append(env.steps_since_action)
    save("master_reset_mean", dict(kind="master_reset_mean", n=2, size=64, win=32,
                                   rule="B3/S23", step_numbers=stepnos,
                                   steps_since_action=since),
         init=init, actions=np.stack(seq), states=np.stack(outs))
    # the same predicate on a batch-1 (broadcast) action and on the 256 / 64 geometry
    env = CARLE(instances=3, height=256, width=256, action_height=64,
                action_width=64, device="cpu")
    env.reset()
    torch.manual_seed(205)
    env.universe = (torch.rand(3, 1, 256, 256) < 0.3).float()
    init = bits(env.universe)
    ii, jj = torch.meshgrid(torch.arange(64), torch.arange(64), indexing="ij")
    checker1 = ((ii + jj) % 2).float()[None, None]
    b0 = 1.0 * (torch.rand(1, 1, 64, 64) <= 0.1)
    seq, outs, stepnos, since = [], [], [], []
    for a in (b0, 2.0 * checker1, b0, 3.0 * checker1, 0.5 + checker1, b0):
        seq.append(a.numpy().copy())
        outs.append(bits(env.step(a)[0]))
        stepnos.append(env.step_number)
        since.append(env.steps_since_action)
    save("master_reset_mean_b1", dict(kind="master_reset_mean", n=3, size=256, win=64,
                                      rule="B3/S23", step_numbers=stepnos,
                                      steps_since_action=since),
         init=init, actions=np.stack(seq), states=np.stack(outs))

    # (c) grid-sized action is centre-cropped (env.py:164-169)
    env = CARLE(instances=2, height=64, width=64, action_height=32,
                action_width=32, device="cpu")
    env.reset()
    torch.manual_seed(202)
    big = 1.0 * (torch.rand(2, 1, 64, 64) <= 0.2)
    obs = env.step(big)[0]
    save("grid_sized_action", dict(kind="grid_sized_action", n=2, size=64, win=32,
                                   rule="B3/S23"),
         action=big.numpy(), final=bits(obs))

    # (d) non-square window on a square grid: action is [N,1,aw,ah]
    env = CARLE(instances=2, height=64, width=64, action_height=16,
                action_width=32, device="cpu")
    env.reset()
    torch.manual_seed(203)
    env.universe = (torch.rand(2, 1, 64, 64) < 0.3).float()
    init = bits(env.universe)
    a = 1.0 * (torch.rand(2, 1, 32, 16) <= 0.3)
    obs = env.step(a)[0]
    save("nonsquare_window", dict(kind="nonsquare_window", n=2, size=64, aw=32,
                                  ah=16, rule="B3/S23"),
         init=init, action=a.numpy(), final=bits(obs))

    # (e) action placement probes: single toggles, zero-generation view via
    #     apply_action only
    probes = []
    for size, win in ((16, 8), (64, 32), (64, 31), (128, 32), (256, 64)):
        env = CARLE(instances=1, height=size, width=size, action_height=win,
                    action_width=win, device="cpu")
        env.reset()
        for (r, c) in ((0, 0), (win - 1, win - 1), (0, win - 1), (3, 1)):
            env.reset()
            a = torch.zeros(1, 1, win, win)
            a[0, 0, r, c] = 1.0
            env.apply_action(a)
            pos = torch.nonzero(env.universe[0, 0]).numpy().tolist()
            probes.append(dict(size=size, win=win, r=r, c=c, cell=pos[0]))
    manifest["placement_probes"] = probes

    # ----- known-answer pair shipped by the reference ------------------------
    env = CARLE(instances=1, height=16, width=16, action_height=8, action_width=8,
                device="cpu")
    env.reset()
    env.load_universe(os.path.join(REF, "carle", "spaceship_duck.rle"))
    duck = bits(env.universe)
    obs = env.step(torch.zeros(1, 1, 8, 8))[0]
    env2 = CARLE(instances=1, height=16, width=16, action_height=8,
                 action_width=8, device="cpu")
    env2.reset()
    env2.load_universe(os.path.join(REF, "carle", "spaceship_step.rle"))
    assert torch.equal(obs, env2.universe), "reference fixture pair mismatch"
    save("spaceship", dict(kind="spaceship", size=16, win=8, rule="B3/S23"),
         duck=duck, step=bits(env2.universe))

    # ----- reduction wrappers -------------------------------------------------
    def wrapped(name, wrapper_cls, seed, n, size, win, rule, steps, fill,
                zero_every=None, tweak=None):
        torch.manual_seed(seed)
        inner = CARLE(instances=n, height=size, width=size, action_height=win,
                      action_width=win, device="cpu")
        env = wrapper_cls(inner)
        if tweak:
            tweak(env)
        inner.rules_from_string(rule)
        env.reset()
        inner.universe = (torch.rand(n, 1, size, size) < fill).float()
        init = bits(inner.universe)
        acts, rewards, states = [], [], []
        for t in range(steps):
            if zero_every is not None and zero_every(t):
                a = torch.zeros(n, 1, win, win)
            else:
                a = 1.0 * (torch.rand(n, 1, win, win) <= 0.1)
            acts.append(np.packbits((a.numpy()[:, 0] != 0).astype(np.uint8),
                                    axis=-1, bitorder="little"))
            obs, r, d, info = env.step(a)
            rewards.append(np.asarray(r.detach().cpu().numpy(),
                                      dtype=np.float32).copy())
            states.append(bits(obs))
        save(name, dict(kind="wrapper", wrapper=wrapper_cls.__name__, seed=seed,
                        n=n, size=size, win=win, rule=rule, steps=steps),
             init=init, actions=np.stack(acts), rewards=np.stack(rewards),
             final=states[-1])

    wrapped("speed_64", ref_mcl.SpeedDetector, 300, 5, 64, 32, "B3/S23", 12, 0.3)
    wrapped("speed_128", ref_mcl.SpeedDetector, 301, 3, 128, 32, "B368/S245", 8,
            0.3)
    wrapped("speed_256", ref_mcl.SpeedDetector, 302, 2, 256, 64, "B368/S245", 6,
            0.2)
    wrapped("corner_256", ref_mcl.CornerBonus, 303, 3, 256, 64, "B3/S23", 6, 0.4)
    wrapped("corner_128", ref_mcl.CornerBonus, 304, 2, 128, 32, "B3/S23", 6, 0.4)

    def small_window(env):
        env.growth_threshold = 4
    wrapped("puffer_64", ref_mcl.PufferDetector, 305, 2, 64, 32,
            "B3/S012345678", 14, 0.05, zero_every=lambda t: t not in (0, 9),
            tweak=small_window)

    # ParsimonyBonus on top of CornerBonus (non-zero inner reward), N=1 and N=3
    for n in (1, 3):
        torch.manual_seed(310 + n)
        inner = CARLE(instances=n, height=256, width=256, device="cpu")
        env = ref_mcl.ParsimonyBonus(ref_mcl.CornerBonus(inner))
        env.reset()
        inner.universe = (torch.rand(n, 1, 256, 256) < 0.4).float()
        init = bits(inner.universe)
        acts, rewards = [], []
        for t in range(4):
            p = 0.01 if t % 2 == 0 else 0.2
            a = 1.0 * (torch.rand(n, 1, 64, 64) <= p)
            acts.append(np.packbits((a.numpy()[:, 0] != 0).astype(np.uint8),
                                    axis=-1, bitorder="little"))
            obs, r, d, info = env.step(a)
            rewards.append(r.detach().cpu().numpy().astype(np.float32).copy())
        save(f"parsimony_n{n}", dict(kind="parsimony", n=n, size=256, win=64,
                                     rule="B3/S23", steps=4,
                                     reward_shape=list(rewards[0].shape)),
             init=init, actions=np.stack(acts), rewards=np.stack(rewards),
             final=bits(obs))

    # ----- RLE codec: the reference's own get_rle text and rle_to_grid grids ----
    # (env.py:408-464 / 260-328).  The reference cannot read back its own header
    # (env.py:349 raises ValueError on "S23:T64, 64"), so rle_to_grid is fed the
    # body, i.e. everything after the third header line.
    def rle_case(name, cells, rule="B3/S23", step_number=7):
        size = cells.shape[-1]
        env = CARLE(instances=1, height=size, width=size, action_height=size // 2,
                    action_width=size // 2, device="cpu")
        env.rules_from_string(rule)
        env.reset()
        env.instance_id = "1234567890"
        env.step_number = step_number
        env.universe = torch.as_tensor(cells, dtype=torch.float32).reshape(1, 1, size, size)
        text = env.get_rle(env.universe[0, 0])
        action_text = env.get_rle(env.universe[0, 0, :size // 2, :size // 2], action=True)
        body = text.split("\n", 3)[3]
        grid = env.rle_to_grid(body)
        save(name, dict(kind="rle", size=size, rule=rule, step_number=step_number,
                        instance_id="1234567890"),
             cells=np.packbits(np.asarray(cells, dtype=np.uint8), axis=-1, bitorder="little"),
             text=np.frombuffer(text.encode("ascii"), dtype=np.uint8),
             action_text=np.frombuffer(action_text.encode("ascii"), dtype=np.uint8),
             decoded=np.packbits(grid.numpy().astype(np.uint8), axis=-1, bitorder="little"))

    torch.manual_seed(400)
    rle_case("rle_soup_64", (torch.rand(64, 64) < 0.35).numpy())
    rle_case("rle_sparse_128", (torch.rand(128, 128) < 0.02).numpy(), rule="B368/S245")
    g = np.zeros((32, 32), dtype=np.uint8)
    g[3, 4] = g[4, 5] = g[5, 3] = g[5, 4] = g[5, 5] = 1           # one glider, mostly blank rows
    rle_case("rle_glider_32", g, step_number=0)
    rle_case("rle_full_16", np.ones((16, 16), dtype=np.uint8))

    # ----- MorphoBonus (mcl.py:107-195).  Its default patterns glider_1.rle / glider_2.rle are
    # not shipped upstream, so the two glider phases are written to a temp dir (standard RLE
    # header, which the reference's read_rle does parse) and add_default_patterns is pointed
    # there; everything else is the reference's class as it stands. ---------------------------
    import tempfile
    tmp = tempfile.mkdtemp()
    phases = {"glider_1.rle": "bob$2bo$3o!", "glider_2.rle": "obo$b2o$bo!"}
    for fname, body in phases.items():
        with open(os.path.join(tmp, fname), "w") as f:
            f.write("x = 3, y = 3, rule = B3/S23\n" + body + "\n")

    def default_patterns(self):
        self.add_rle_pattern(os.path.join(tmp, "glider_1.rle"))
        self.add_rle_pattern(os.path.join(tmp, "glider_2.rle"))
    ref_mcl.MorphoBonus.add_default_patterns = default_patterns

    def morpho_case(name, seed, n, size, win, steps, fill, grid_sized):
        torch.manual_seed(seed)
        inner = CARLE(instances=n, height=size, width=size, action_height=win,
                      action_width=win, device="cpu")
        env = ref_mcl.MorphoBonus(inner)
        env.reset()
        inner.universe = (torch.rand(n, 1, size, size) < fill).float()
        # a few real gliders so that the maximum is a full match somewhere
        glider = torch.tensor([[0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=torch.float32)
        for k in range(n):
            inner.universe[k, 0, 5 + 3 * k:8 + 3 * k, 9:12] = glider
        init = bits(inner.universe)
        a_size = size if grid_sized else win
        acts, rewards, states = [], [], []
        for t in range(steps):
            a = 1.0 * (torch.rand(n, 1, a_size, a_size) <= (0.0 if t == 1 else 0.05))
            acts.append(np.packbits((a.numpy()[:, 0] != 0).astype(np.uint8), axis=-1,
                                    bitorder="little"))
            obs, r, d, info = env.step(a)
            rewards.append(r.detach().cpu().numpy().astype(np.float32).copy())
            states.append(bits(obs))
        save(name, dict(kind="morpho", seed=seed, n=n, size=size, win=win, steps=steps,
                        rule="B3/S23", grid_sized=grid_sized,
                        pattern_count=int(env.target_patterns.shape[0])),
             init=init, actions=np.stack(acts), rewards=np.stack(rewards), final=states[-1],
             patterns=env.target_patterns.detach().cpu().numpy().astype(np.float32))

    morpho_case("morpho_64_grid_action", 410, 3, 64, 32, 5, 0.08, True)
    morpho_case("morpho_32_full_window", 411, 2, 32, 32, 4, 0.15, False)
    morpho_case("morpho_128_grid_action", 412, 2, 128, 32, 3, 0.3, True)

    # ----- values pinned by the reference's own tests/test_env.py -------------
    env = CARLE()
    env.birth_rule_from_string("asdfasdfB0357*!@#!@$%")
    env.survive_rule_from_string("S2468")
    manifest["rule_parser"] = dict(birth=env.birth, survive=env.survive)
    env.rules_from_string("23/3")
    manifest["rule_parser_23_3"] = dict(birth=env.birth, survive=env.survive)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", len(manifest["cases"]), "cases; digests:",
          json.dumps(manifest["digests"]))


if __name__ == "__main__":
    main()
