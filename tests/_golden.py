"""Helpers shared by the oracle (CPU) and parity (GPU) tests: golden-case loading."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN_DIR, "golden.json")) as _f:
    MANIFEST = json.load(_f)

CASES = {c["name"]: c for c in MANIFEST["cases"]}


def load(name):
    return CASES[name], np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def unbits(packed, width):
    """packed uint8 [..., ceil(W/8)] LSB-first -> uint8 0/1 [..., W]."""
    return np.unpackbits(packed, axis=-1, bitorder="little")[..., :width]


def action_from_bits(packed, win_cols):
    """packed actions [N, aw, ceil(ah/8)] -> float32 [N, 1, aw, ah]."""
    return unbits(packed, win_cols).astype(np.float32)[:, None]


def by_kind(kind):
    return [c["name"] for c in MANIFEST["cases"] if c["kind"] == kind]
