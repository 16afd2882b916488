"""CPU-side checks of the boundary: the shared library builds, loads, exports exactly
the symbols include/carle_b200.h declares, validates arguments without a GPU, and the
host-side helpers (rule masks, mask packing, RLE codec) behave.  No compute calls."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from carle_b200 import build, _lib
    build.build()                      # no-op when up to date
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from carle_b200 import _lib
    header = open(os.path.join(ROOT, "include", "carle_b200.h")).read()
    declared = set(re.findall(r"CARLE_API\s+[\w\s\*]+?\b(carle_\w+)\s*\(", header))
    assert declared, "no CARLE_API declarations found"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.carle_version() >= 100


def test_argument_validation_without_gpu(lib):
    from carle_b200 import _lib
    h = ctypes.c_void_p()
    rc = lib.carle_create(ctypes.byref(h), 0, 0, 64, 64, 32, 32)       # zero instances
    assert rc == _lib.CARLE_EINVAL and b"non-positive" in lib.carle_last_error()
    rc = lib.carle_create(ctypes.byref(h), 0, 1, 64, 64, 32, 32)
    if not torch.cuda.is_available():
        # the product has no CPU path: creating a handle must fail loudly
        assert rc == _lib.CARLE_ENODEV
        assert b"no CPU path" in lib.carle_last_error()
        with pytest.raises(_lib.CarleLibraryError):
            _lib.check(rc, "carle_create")
    assert lib.carle_set_rule(None, 8, 12) == _lib.CARLE_EINVAL
    assert lib.carle_step(None, None, None, None, 1, None, None, None, None) == _lib.CARLE_EINVAL


def test_env_requires_cuda():
    import carle_b200
    from carle_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CarleLibraryError):
        carle_b200.CARLE()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "carle_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_rule_mask_and_pack_mask():
    from carle_b200.env import _rule_mask
    from carle_b200.mcl import pack_mask
    assert _rule_mask([3]) == 0x008 and _rule_mask([2, 3]) == 0x00C
    assert _rule_mask([3, 6, 8]) == 0x148 and _rule_mask([2, 4, 5]) == 0x034
    assert _rule_mask([]) == 0 and _rule_mask([9, -1, 3]) == 0x008
    rng = np.random.default_rng(0)
    for h, w in ((4, 32), (5, 70), (3, 16), (2, 256)):
        m = (rng.random((h, w)) < 0.5)
        words = pack_mask(torch.from_numpy(m), "cpu").numpy().view(np.uint32)
        assert words.shape == (h, (w + 31) // 32)
        for r in range(h):
            for c in range(w):
                assert ((words[r, c // 32] >> (c % 32)) & 1) == int(m[r, c])
        assert (words[:, -1] >> ((w - 1) % 32 + 1) == 0).all() or w % 32 == 0


def _fake_env(h, w):
    from carle_b200 import rle
    from carle_b200 import _lib
    env = types.SimpleNamespace(height=h, width=w, birth=[3], survive=[2, 3],
                                instance_id="0", step_number=7, _lib=_lib.load())
    for name in ("rle_to_grid", "rle_to_packed", "rle_body", "read_rle", "get_rle"):
        setattr(env, name, types.MethodType(getattr(rle, name), env))
    return env


def test_rle_roundtrip_and_reference_fixture_format(tmp_path):
    env = _fake_env(16, 16)
    # the LWSS phase shipped by the reference as carle/spaceship_duck.rle
    path = tmp_path / "duck.rle"
    path.write_text("#CXRLE Pos=7,-7 Gen=36\nx = 6, y = 4, rule = B36/S125\n3b2o$3ob2o$5o$b3o!\n")
    body = env.read_rle(str(path))
    assert env.birth == [3, 6] and env.survive == [1, 2, 5]
    grid = env.rle_to_grid(body).numpy()
    want = np.zeros((16, 16))
    want[0, 3:5] = 1
    want[1, 0:3] = 1
    want[1, 4:6] = 1
    want[2, 0:5] = 1
    want[3, 1:4] = 1
    assert np.array_equal(grid, want)
    # encode -> decode round trip, including the header get_rle itself writes
    rng = np.random.default_rng(1)
    cells = (rng.random((16, 16)) < 0.4).astype(np.float32)
    text = env.get_rle(torch.from_numpy(cells))
    assert text.startswith("#C exp_id=0 \n#C step=7 (universe) \nx = 0, y = 0, rule = B36/S125:T16, 16\n")
    assert text.endswith("!")
    path2 = tmp_path / "own.rle"
    path2.write_text(text)
    env2 = _fake_env(16, 16)
    body2 = env2.read_rle(str(path2))
    assert env2.birth == [3, 6] and env2.survive == [1, 2, 5]
    assert np.array_equal(env2.rle_to_grid(body2).numpy(), cells)
    # multi-row '$' runs and bare tags
    assert np.array_equal(env.rle_to_grid("o2$2bo!").numpy()[:3, :3],
                          np.array([[1, 0, 0], [0, 0, 0], [0, 0, 1]]))


def _golden_names(kind):
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return [c["name"] for c in json.load(f)["cases"] if c["kind"] == kind]


@pytest.mark.parametrize("name", _golden_names("rle"))
def test_library_rle_codec_matches_the_reference_text(name):
    """carle_rle_encode_host / carle_rle_decode_host (host C++ over packed words, no GPU) against
    the text the reference's get_rle emitted and the grid its rle_to_grid decoded."""
    import _cases as cs
    from carle_b200 import rle

    def encode(cells, keep_tail):
        env = _fake_env(*cells.shape)
        return env.rle_body(rle.pack_cells(cells), cells.shape[0], cells.shape[1], keep_tail)

    def decode(text, h, w):
        return _fake_env(h, w).rle_to_grid(text).numpy().astype(np.uint8)

    cs.check_rle_codec(name, encode, decode)


def test_library_rle_codec_edge_cases(lib):
    from carle_b200 import _lib, rle
    env = _fake_env(8, 70)                          # a width that is no multiple of 32
    rng = np.random.default_rng(5)
    cells = (rng.random((8, 70)) < 0.5).astype(np.uint8)
    assert np.array_equal(env.rle_to_grid(env.rle_body(rle.pack_cells(cells), 8, 70)).numpy(), cells)
    # counts straddling a line break, upper-case tags, cells outside the grid are dropped
    assert np.array_equal(env.rle_to_grid("1\n2O$B3o!").numpy()[:2, :12],
                          np.array([[1] * 12, [0, 1, 1, 1] + [0] * 8]))
    assert env.rle_to_grid("100o20$5o!").numpy().sum() == 70
    # sizing call, too-small buffer, bad arguments
    words = rle.pack_cells(cells)
    ptr = words.ctypes.data_as(ctypes.c_void_p)
    need = lib.carle_rle_encode_host(ptr, 8, 70, _lib.RLE_KEEP_TAIL, None, 0)
    buf = ctypes.create_string_buffer(4)
    assert lib.carle_rle_encode_host(ptr, 8, 70, _lib.RLE_KEEP_TAIL, buf, 4) == need and buf.raw == b"\0" * 4
    assert lib.carle_rle_encode_host(None, 8, 70, 0, None, 0) == _lib.CARLE_EINVAL
    assert lib.carle_rle_decode_host(None, 0, 8, 70, None) == _lib.CARLE_EINVAL


def test_jit_probe_compiles_specialised_kernels_without_a_gpu():
    """NVRTC specialisation of the one-launch step kernels for an arbitrary rule (jit.cu): the
    embedded kernel headers compile for sm_100a for every fused shape; no GPU involved."""
    import ctypes
    from carle_b200 import _lib
    lib = _lib.load()
    try:
        ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc is not installed here")
    for shape in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
        size = ctypes.c_int64(0)
        rc = lib.carle_jit_probe(shape, 0b001001000, 0b000100110, ctypes.byref(size))   # B36/S125
        assert rc == 0, _lib.last_error()[:2000]
        assert size.value > 10000
    assert lib.carle_jit_probe(11, 8, 12, None) == _lib.CARLE_EINVAL
    assert lib.carle_jit_probe(1, 0, 12, None) == _lib.CARLE_ERULE
    assert lib.carle_jit_loaded() == 0                 # probing never loads anything


@pytest.mark.parametrize("isa", ["auto", "avx2", "sse2", "entrywise", "streams3"])
def test_host_side_action_packing(lib, isa, monkeypatch):
    """carle_pack_action_host (host threads, no device): float32 / uint8 actions -> grid-aligned packed
    words + the three flags, against numpy, for word-aligned and unaligned windows, ragged widths,
    one and several threads, every instruction-set path, the flat multi-stream walk of word-aligned windows
    (several grabs per thread, a left-over behind the streams) and the entry-by-entry walk."""
    from carle_b200 import _lib
    if isa in ("sse2", "avx2"):
        monkeypatch.setenv("CARLE_HOST_PACK_ISA", isa)
    elif isa == "entrywise":
        monkeypatch.setenv("CARLE_HOST_PACK_FLAT", "0")
    elif isa == "streams3":
        monkeypatch.setenv("CARLE_HOST_PACK_STREAMS", "3")
        monkeypatch.setenv("CARLE_HOST_PACK_PREFETCH", "0")
    rng = np.random.default_rng(3)
    for aw, ah, bit0, batch, threads in ((64, 64, 0, 1100, 4), (32, 32, 16, 300, 3), (30, 30, 3, 17, 1),
                                         (5, 77, 31, 9, 2), (64, 64, 0, 1, 8), (33, 32, 0, 7, 2), (3, 96, 0, 700, 3)):
        awpr = (bit0 + ah + 31) // 32
        for kind in ("f32", "u8", "ones", "zeros", "nonbinary", "nan", "odd_floats"):
            if kind == "u8":
                a = (rng.random((batch, aw, ah)) < 0.1).astype(np.uint8)
                a[0, 0, 0] = 7                                    # any non-zero byte toggles
            elif kind == "ones":
                a = np.ones((batch, aw, ah), dtype=np.float32)
            elif kind == "zeros":
                a = np.zeros((batch, aw, ah), dtype=np.float32)
            else:
                a = (rng.random((batch, aw, ah)) < 0.1).astype(np.float32)
                if kind == "nonbinary":
                    a[batch // 2, aw - 1, ah - 1] = 0.5
                if kind == "nan":
                    a[0, 0, ah // 2] = np.nan
                if kind == "odd_floats":
                    # -0.0 is zero (no toggle, not binary-breaking), a denormal / inf / negative toggles
                    flat = a.reshape(-1)
                    flat[::7] = -0.0
                    flat[3::11] = np.float32(1e-42)
                    flat[5::13] = np.inf
                    flat[1::17] = -1.0
            out = np.full((batch, aw, awpr), 0xDEADBEEF, dtype=np.uint32)
            flags = (ctypes.c_int32 * 3)()
            rc = lib.carle_pack_action_host(aw, ah, awpr, bit0, a.ctypes.data_as(ctypes.c_void_p),
                                            _lib.U8 if a.dtype == np.uint8 else _lib.F32, batch,
                                            out.ctypes.data_as(ctypes.c_void_p), flags, threads)
            assert rc == 0, _lib.last_error()
            bits = np.zeros((batch, aw, awpr * 32), dtype=np.uint8)
            bits[:, :, bit0:bit0 + ah] = a != 0
            want = np.packbits(bits, axis=-1, bitorder="little").view("<u4").reshape(batch, aw, awpr)
            assert np.array_equal(out, want), (aw, ah, bit0, kind)
            ones = (a == 1)
            assert bool(flags[0]) == (not ones.all()) and bool(flags[1]) == bool((a != 0).any())
            assert bool(flags[2]) == bool(((a != 0) & ~ones).any()), kind
    assert lib.carle_pack_action_host(4, 4, 9, 0, None, 0, 1, None, None, 1) == _lib.CARLE_EINVAL
    # the variant that also ships the words needs a device
    a = np.zeros((2, 32, 32), dtype=np.float32)
    out = np.zeros((2, 32, 1), dtype=np.uint32)
    flags = (ctypes.c_int32 * 3)()
    args = (32, 32, 1, 0, a.ctypes.data_as(ctypes.c_void_p), _lib.F32, 2, out.ctypes.data_as(ctypes.c_void_p), flags, 1)
    assert lib.carle_pack_action_host_copy(*args, None, 0, None) == _lib.CARLE_EINVAL
    if not torch.cuda.is_available():
        assert lib.carle_pack_action_host_copy(*args, out.ctypes.data_as(ctypes.c_void_p), 0, None) == _lib.CARLE_ENODEV
