"""Host-side I/O helpers of ``CARLE`` (reference side layer, carle/env.py:244-513):
text render, RLE decode/encode, per-step log, PNG frame.

These are the on-disk formats beside the hot path, not the hot path: they run on a
host copy of ONE universe's PACKED WORDS (H * ceil(W/32) * 4 bytes cross the bus, not the
float batch), and the run-length codec itself is the library's C++ one over those words
(``carle_rle_encode_host`` / ``carle_rle_decode_host``, include/carle_b200.h) -- a 256 x 256
universe encodes in tens of microseconds instead of the reference's per-cell Python loop.
The functions are written as methods (first argument ``self`` is the ``CARLE`` instance) and
attached to the class in ``env.py``.  Two upstream defects are fixed rather than copied:
``read_rle`` parses the ``rule = B3/S23:T16, 16`` header that ``get_rle`` itself emits
(upstream raises ValueError on it, env.py:349), and ``get_rle`` flushes the last partial
line instead of dropping it (env.py:453-455; ``keep_tail=False`` reproduces the reference's
text byte for byte -- pinned by tests/golden ``rle_*``).
"""
import ctypes
import os
import re
import struct
import time
import zlib

import numpy as np
import torch

from . import _lib

_RULE = re.compile(r"rule\s*=\s*([^\s,]+)")


def render(self):
    """env.py:244-258 — print instance 0 as text."""
    grid = self.instance_cells(0)
    os.system("clear")
    print("\n CA Universe")
    for row in grid:
        print("")
        print("".join("o" if v else " " for v in row), end="")
    time.sleep(0.125)


def pack_cells(cells):
    """0/1 array ``[H, W]`` -> uint32 ``[H, ceil(W/32)]`` in the library's state layout."""
    cells = np.asarray(cells) != 0
    h, w = cells.shape
    wpr = (w + 31) // 32
    padded = np.zeros((h, wpr * 32), dtype=np.uint8)
    padded[:, :w] = cells
    return np.ascontiguousarray(np.packbits(padded, axis=-1, bitorder="little").view("<u4"))


def unpack_cells(words, width):
    """uint32 ``[H, WPR]`` -> uint8 0/1 ``[H, width]``."""
    words = np.ascontiguousarray(words).view(np.uint32)
    bits = np.unpackbits(words.view(np.uint8), axis=-1, bitorder="little")
    return bits.reshape(words.shape[0], -1)[:, :width]


def rle_body(self, words, height, width, keep_tail=True):
    """Run tokens + line breaks + ``!`` of one packed grid (env.py:424-462) -- the library's
    host codec.  ``keep_tail=False`` drops the last partial line exactly as upstream does."""
    words = np.ascontiguousarray(words, dtype=np.uint32)
    flags = _lib.RLE_KEEP_TAIL if keep_tail else 0
    ptr = words.ctypes.data_as(ctypes.c_void_p)
    need = self._lib.carle_rle_encode_host(ptr, height, width, flags, None, 0)
    if need < 0:
        _lib.check(int(need), "carle_rle_encode_host")
    buf = ctypes.create_string_buffer(int(need))
    self._lib.carle_rle_encode_host(ptr, height, width, flags, buf, need)
    return buf.raw[:need].decode("ascii")


def rle_to_packed(self, rle, height=None, width=None):
    """RLE body -> uint32 numpy ``[H, ceil(W/32)]`` (top-left anchored)."""
    height = self.height if height is None else height
    width = self.width if width is None else width
    text = rle.encode("ascii", "replace")
    words = np.zeros((height, (width + 31) // 32), dtype=np.uint32)
    _lib.check(self._lib.carle_rle_decode_host(text, len(text), height, width,
                                               words.ctypes.data_as(ctypes.c_void_p)),
               "carle_rle_decode_host")
    return words


def rle_to_grid(self, rle):
    """env.py:260-328 — decode an RLE body into a float ``[H, W]`` grid (top-left
    anchored; ``b`` dead run, ``o`` live run, ``$`` end of row(s), ``!`` end)."""
    words = self.rle_to_packed(rle)
    return torch.from_numpy(unpack_cells(words, self.width).astype(np.float32))


def read_rle(self, filepath):
    """env.py:330-382 — read an RLE file: sets ``birth``/``survive`` from its ``rule``
    header and returns the body text."""
    body = []
    seen_rule = False
    with open(filepath, "r") as f:
        for line in f.readlines():
            if seen_rule:
                body.append(line)
            elif "rule" in line:
                m = _RULE.search(line)
                parts = m.group(1).split("/") if m else ["B3", "S23"]
                survive_part = parts[-1].split(":")[0]
                self.birth = sorted({int(c) for c in parts[0] if c.isdigit() and c != "9"})
                self.survive = sorted({int(c) for c in survive_part
                                       if c.isdigit() and c != "9"})
                seen_rule = True
    return "".join(body)


def read_csv(self, filepath):
    """env.py:384-388 — not implemented upstream either."""
    print("warning, read_csv not implemented yet")
    return ""


def load_universe(self, filepath, universe_index=0):
    """env.py:390-406 — load an RLE pattern into one instance (top-left anchored).  The text is
    decoded straight into packed words and copied over the instance's rows."""
    text = self.read_rle(filepath) if "rle" in filepath[-4:] else self.read_csv(filepath)
    if self._packed is None:
        raise AttributeError("universe is undefined before reset() (as upstream)")
    words = self.rle_to_packed(text)
    self._absorb_view()
    self._packed[universe_index].copy_(torch.from_numpy(words.view(np.int32)))
    self._view, self._view_stale = None, True


def _rle_header(self, action):
    kind = "action" if action else "universe"
    return (f"#C exp_id={self.instance_id} \n"
            f"#C step={self.step_number} ({kind}) \n"
            "x = 0, y = 0, rule = B" + "".join(str(b) for b in self.birth) +
            "/S" + "".join(str(s) for s in self.survive) +
            f":T{self.height}, {self.width}\n")


def get_rle(self, universe, action=False, keep_tail=True):
    """env.py:408-464 — RLE text (with the reference's header) of one 2-D grid."""
    cells = torch.as_tensor(universe).squeeze().detach().cpu().numpy() != 0
    if cells.ndim != 2:
        raise ValueError("get_rle expects one 2-D universe")
    return _rle_header(self, action) + self.rle_body(pack_cells(cells), cells.shape[0],
                                                     cells.shape[1], keep_tail)


def log_universe(self, universe_index=0):
    """env.py:466-476 — append ``[action_rle, universe_rle]`` to ``self.log``."""
    # the universe goes packed words -> text: nothing is unpacked on the way
    self._absorb_view()
    words = self._packed[universe_index].cpu().numpy().view(np.uint32)
    rle_universe = _rle_header(self, False) + self.rle_body(words, self.height, self.width)
    action = self.action.to_float() if hasattr(self.action, "to_float") else self.action
    rle_action = self.get_rle(torch.as_tensor(action)[universe_index, 0, :, :], action=True)
    self.log.append([rle_action, rle_universe])


def save_log(self):
    """env.py:479-491."""
    with open(f"./logs/carle_log{self.instance_id}.csv", "w") as f:
        f.write("action,universe,\n")
        for entry in self.log:
            for item in entry:
                f.write('"' + item + '"' + ",")
            f.write("\n")


def save_rle(self, rle):
    """env.py:495-500."""
    with open(f"./logs/universe{self.instance_id}_step{self.step_number}.rle", "w") as f:
        f.write(rle)


def _png_gray8(pixels):
    h, w = pixels.shape
    raw = b"".join(b"\x00" + pixels[r].tobytes() for r in range(h))

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def save_frame(self):
    """env.py:504-513 — 8-bit PNG of instance 0 (self-contained encoder; upstream
    uses scikit-image, which this image does not ship)."""
    pixels = np.uint8(255) * self.instance_cells(0)
    path = f"./frames/frame{self.instance_id}_step{self.step_number}.png"
    with open(path, "wb") as f:
        f.write(_png_gray8(np.ascontiguousarray(pixels)))
