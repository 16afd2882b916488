"""Host-side I/O helpers of ``CARLE`` (reference side layer, carle/env.py:244-513):
text render, RLE decode/encode, per-step log, PNG frame.

These are the on-disk formats beside the hot path, not the hot path: they run on a
host copy of ONE universe, decoded from its packed words (``CARLE.instance_cells``), so
``logging=True`` costs one small D2H copy per step instead of unpacking the whole batch.  The functions are written as
methods (first argument ``self`` is the ``CARLE`` instance) and attached to the class
in ``env.py``.  Two upstream defects are fixed rather than copied: ``read_rle`` parses
the ``rule = B3/S23:T16, 16`` header that ``get_rle`` itself emits (upstream raises
ValueError on it, env.py:349), and ``get_rle`` flushes the last partial line instead of
dropping it (env.py:447-462).
"""
import os
import re
import struct
import time
import zlib

import numpy as np
import torch

_TOKEN = re.compile(r"(\d*)([bBoO$!])")
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


def rle_to_grid(self, rle):
    """env.py:260-328 — decode an RLE body into a float ``[H, W]`` grid (top-left
    anchored; ``b`` dead run, ``o`` live run, ``$`` end of row(s), ``!`` end)."""
    grid = torch.zeros(self.height, self.width)
    row = col = 0
    for count, tag in _TOKEN.findall(rle.replace("\n", "")):
        run = int(count) if count else 1
        tag = tag.lower()
        if tag == "b":
            col += run
        elif tag == "o":
            grid[row, col:col + run] = 1
            col += run
        elif tag == "$":
            row += run
            col = 0
        else:                       # "!"
            break
    return grid


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
    """env.py:390-406 — load an RLE pattern into one instance (top-left anchored)."""
    text = self.read_rle(filepath) if "rle" in filepath[-4:] else self.read_csv(filepath)
    grid = self.rle_to_grid(text)
    universe = self.universe
    assert universe.shape[2] == grid.shape[0] and universe.shape[3] == grid.shape[1], \
        "tried to load the wrong size universe"
    universe[universe_index, 0, :, :] = grid.to(universe.device)


def _encode_rows(cells):
    """Run-length tokens of a 2-D 0/1 array, one ``$`` per row, every count explicit
    (the reference writes ``1o`` not ``o``)."""
    tokens = []
    names = ("b", "o")
    for row in cells:
        change = np.flatnonzero(np.diff(row)) + 1
        starts = np.concatenate(([0], change))
        ends = np.concatenate((change, [row.shape[0]]))
        runs = [f"{e - s}{names[int(row[s])]}" for s, e in zip(starts, ends)]
        runs[-1] += "$"
        tokens.extend(runs)
    return tokens


def get_rle(self, universe, action=False):
    """env.py:408-464 — RLE text (with the reference's header) of one 2-D grid."""
    cells = (torch.as_tensor(universe).squeeze().detach().cpu().numpy() != 0).astype(np.int8)
    kind = "action" if action else "universe"
    out = [f"#C exp_id={self.instance_id} \n",
           f"#C step={self.step_number} ({kind}) \n",
           "x = 0, y = 0, rule = B" + "".join(str(b) for b in self.birth) +
           "/S" + "".join(str(s) for s in self.survive) +
           f":T{self.height}, {self.width}\n"]
    line = ""
    for tok in _encode_rows(cells):
        line += tok
        if len(line) > 69:
            out.append(line + "\n")
            line = ""
    out.append(line)            # upstream drops this last partial line
    out.append("!")
    return "".join(out)


def log_universe(self, universe_index=0):
    """env.py:466-476 — append ``[action_rle, universe_rle]`` to ``self.log``."""
    rle_universe = self.get_rle(self.instance_cells(universe_index))
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
