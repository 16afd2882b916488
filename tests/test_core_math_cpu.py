"""Pins carle_b200/csrc/ca_core.cuh — the bit-sliced arithmetic the kernels are built
from — on the CPU: the same header is compiled with g++ (tests/cpu_twin/twin.cpp, test
infrastructure only) and compared with the oracle.  The kernels' data movement (warp
shuffles, tiling) is covered by the GPU parity tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import carle_oracle as oc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_twin", "twin.cpp")
OUT = os.path.join(HERE, "cpu_twin", "_twin.so")

RULES = {1: ([3], [2, 3]), 2: ([3, 6, 8], [2, 4, 5]), 3: ([3, 6], [2, 3]),
         4: ([3, 6, 7, 8], [3, 4, 6, 7, 8])}


def _mask(vals):
    return sum(1 << v for v in vals)


@pytest.fixture(scope="module")
def twin():
    core = os.path.join(HERE, "..", "carle_b200", "csrc", "ca_core.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC),
                                                              os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                               SRC, "-o", OUT])
    lib = ctypes.CDLL(OUT)
    lib.twin_rule_word.restype = ctypes.c_uint32
    lib.twin_rule_word.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
    lib.twin_exhaustive_dynamic.restype = ctypes.c_long
    lib.twin_rule_triples.argtypes = [ctypes.c_int, ctypes.c_void_p]
    lib.twin_rule_triples.restype = None
    lib.twin_bit_index_sum.restype = ctypes.c_uint32
    lib.twin_bit_index_sum.argtypes = [ctypes.c_uint32]
    lib.twin_csa_bad.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.twin_strip_sums.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p]
    lib.twin_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                              ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
    return lib


def test_static_rule_tables(twin):
    for mode, (b, s) in RULES.items():
        word = twin.twin_rule_word(_mask(b), _mask(s), mode)
        for p in range(32):
            t0, k0, t1, k1, x = p & 1, (p >> 1) & 1, (p >> 2) & 1, (p >> 3) & 1, p >> 4
            sum9 = t0 + 2 * (k0 + t1) + 4 * k1
            if (x == 0 and sum9 > 8) or (x == 1 and sum9 < 1):
                continue
            want = ((sum9 - 1) in s) if x else (sum9 in b)
            assert ((word >> p) & 1) == int(want), (mode, p)


def test_builtin_rules_from_row_triples_all_inputs(twin):
    """The short LOP3 networks that go from the three row triples to the next state
    (ca::life_from_triples: 7 LOP3; morley_ / highlife_from_triples: 8; the path every kernel takes
    for the built-in rules) on all 2^7 combinations of (cell, three row triples) against the
    definition: born if the 3x3 sum including the centre is in B, survives if sum - 1 is in S
    (carle/env.py:219-229)."""
    for mode, (b, s) in RULES.items():
        out = (ctypes.c_uint32 * 4)()
        twin.twin_rule_triples(mode, out)
        for p in range(128):
            x = p & 1
            lo = [(p >> k) & 1 for k in (1, 2, 3)]
            hi = [(p >> k) & 1 for k in (4, 5, 6)]
            sum9 = sum(lo) + 2 * sum(hi)
            if x and sum9 < 1:
                continue                                   # unreachable: the centre counts itself
            want = int(((sum9 - 1) in s) if x else (sum9 in b))
            assert ((out[p // 32] >> (p % 32)) & 1) == want, (mode, p)


def test_dynamic_rule_all_262144_rules(twin):
    assert twin.twin_exhaustive_dynamic() == 0


def test_bit_index_sum(twin):
    rng = np.random.default_rng(0)
    for v in list(rng.integers(0, 2**32, size=200, dtype=np.uint64)) + [0, 1, 2**31, 2**32 - 1]:
        v = int(v)
        assert twin.twin_bit_index_sum(v) == sum(b for b in range(32) if (v >> b) & 1)


def _pack(u):
    n, h, w = u.shape
    wpr = (w + 31) // 32
    pad = np.zeros((n, h, wpr * 32), dtype=np.uint8)
    pad[:, :, :w] = u
    return np.ascontiguousarray(
        np.packbits(pad, axis=-1, bitorder="little").view("<u4").reshape(n, h, wpr))


def _unpack(p, w):
    b = np.unpackbits(p.view(np.uint8), axis=-1, bitorder="little")
    return b.reshape(p.shape[0], p.shape[1], -1)[:, :, :w]


@pytest.mark.parametrize("size", [2, 6, 16, 31 + 1, 34, 64, 100, 128, 160])
def test_generation_matches_oracle(twin, size):
    rng = np.random.default_rng(size)
    u = (rng.random((2, size, size)) < 0.45).astype(np.uint8)
    cases = [(m, b, s) for m, (b, s) in RULES.items()]
    for _ in range(6):
        b = [k for k in range(9) if rng.random() < 0.4] or [3]
        s = [k for k in range(9) if rng.random() < 0.4] or [2]
        cases.append((0, b, s))
    for mode, b, s in cases:
        cur = u
        for gen in range(3):
            packed = _pack(cur)
            out = np.zeros_like(packed)
            twin.twin_step(packed.ctypes.data, out.ctypes.data, 2, size, size, _mask(b),
                           _mask(s), mode)
            want = oc.life_like_update(cur, b, s)
            got = _unpack(out, size)
            assert np.array_equal(got, want), (mode, b, s, gen)
            # dynamic path must agree with the static instantiation of the same rule
            if mode:
                out2 = np.zeros_like(packed)
                twin.twin_step(packed.ctypes.data, out2.ctypes.data, 2, size, size,
                               _mask(b), _mask(s), 0)
                assert np.array_equal(out, out2)
            cur = want


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 12, 16, 32, 64])
def test_carry_save_column_counts(twin, n):
    rng = np.random.default_rng(n)
    for density in (0.0, 0.3, 0.9, 1.0):
        bits = (rng.random((n, 32)) < density).astype(np.uint8)
        words = np.ascontiguousarray(np.packbits(bits, axis=-1, bitorder="little").view("<u4"))
        assert twin.twin_csa_bad(words.ctypes.data, n) == 0


@pytest.mark.parametrize("wpl,r,awin", [(8, 2, 64), (8, 4, 64), (8, 8, 64), (4, 2, 32),
                                        (4, 4, 32), (2, 2, 32), (4, 1, 96)])
def test_strip_sums_match_oracle(twin, wpl, r, awin):
    """ca::strip_lane_sums (the fused SpeedDetector sums of step_strip_kernel, carry-save
    formulation) == oracle speed_sums, including all-live and empty universes."""
    size = 32 * wpl
    rng = np.random.default_rng(wpl * 100 + r)
    ref = oc.OracleCARLE(width=size, height=size, action_width=awin, action_height=awin,
                         instances=1)
    mask = oc.outside_window_mask(ref)
    for density in (0.0, 0.07, 0.5, 1.0):
        u = (rng.random((1, size, size)) < density).astype(np.uint8)
        live, sh, sw = oc.speed_sums(u, mask)
        packed = _pack(u)
        out = np.zeros(4, dtype=np.uint32)
        assert twin.twin_strip_sums(packed.ctypes.data, wpl, r, awin, out.ctypes.data) == 0
        row0 = (size - awin) // 2
        inside = int(u[0, row0:row0 + awin, row0:row0 + awin].sum())
        assert [int(v) for v in out] == [int(live[0]), int(sh[0]), int(sw[0]), inside]
