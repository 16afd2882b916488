#!/usr/bin/env python
"""Opcode histogram of the hot kernels in libcarle_b200.so (cuobjdump -sass), for profiles/:
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt
Shows that the TMA / mbarrier / LOP3 machinery DESIGN.md describes is what is in the binary:
UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier ops, ACQBULK / ELECT,
LOP3.LUT = the bit-sliced adders and rule networks, VOTE / FSETP = the ballot ingest of float32 actions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "carle_b200", "lib", "libcarle_b200.so")
#: (label, regex on the demangled kernel name)
HOT = [
    ("step_strip_kernel 256x256 Morley float32 actions (headline, configs[2])",
     r"step_strip_kernel<\(int\)8, \(int\)4, \(int\)64, carle::StaticRule<\(unsigned int\)328, \(unsigned int\)52>, float, \(int\)1>"),
    ("step_strip_kernel 256x256 Life float32 actions",
     r"step_strip_kernel<\(int\)8, \(int\)4, \(int\)64, carle::StaticRule<\(unsigned int\)8, \(unsigned int\)12>, float, \(int\)1>"),
    ("step_strip_kernel 256x256 Morley packed actions",
     r"step_strip_kernel<\(int\)8, \(int\)4, \(int\)64, carle::StaticRule<\(unsigned int\)328, \(unsigned int\)52>, carle::PackedWords, \(int\)1>"),
    ("step_stream_kernel 128x128 Life float32 actions (configs[1])",
     r"step_stream_kernel<\(int\)4, carle::StaticRule<\(unsigned int\)8, \(unsigned int\)12>, float, \(int\)1, \(int\)8, \(int\)2, \(bool\)0>"),
    ("step_stream_kernel 64x64 Life float32 actions (configs[3] shards)",
     r"step_stream_kernel<\(int\)2, carle::StaticRule<\(unsigned int\)8, \(unsigned int\)12>, float, \(int\)1, \(int\)16, \(int\)2"),
    ("step_tiled_kernel Life, 256-row register tiles (configs[4])",
     r"step_tiled_kernel<carle::StaticRule<\(unsigned int\)8, \(unsigned int\)12>, \(int\)8>"),
    ("step_warp_kernel 128x128 Life (K generations per launch)",
     r"step_warp_kernel<\(int\)4, carle::StaticRule<\(unsigned int\)8, \(unsigned int\)12>>"),
]
SHOW = ["UBLKCP", "UTMALDG", "SYNCS", "ACQBULK", "ELECT", "LDGSTS", "LOP3", "SHF", "POPC", "VOTE", "SHFL", "REDUX",
        "FFMA2", "FSETP", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "BAR", "MEMBAR", "IMAD", "SEL", "PRMT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                           capture_output=True, text=True).stdout.splitlines()
    bodies = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(bodies)} sm_100a kernels; opcode counts are STATIC SASS instructions")
    whole = collections.Counter()
    for body in bodies:
        for op in re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", body, re.M):
            whole[op] += 1
    print("# whole library: " + ", ".join(f"{k} {whole[k]}" for k in ("UBLKCP", "UTMALDG", "SYNCS", "ACQBULK",
                                                                     "ELECT", "LOP3", "FFMA2", "POPC", "REDUX")))
    for label, pattern in HOT:
        hits = [(n, b) for n, b in zip(names, bodies) if re.search(pattern, n)]
        if not hits:
            print(f"\n== {label}: NOT FOUND ({pattern})")
            continue
        name, body = hits[0]
        ops = collections.Counter(
            re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", body, re.M))
        print(f"\n== {label}\n   {name[:150]}")
        print(f"   {sum(ops.values())} instructions; " + ", ".join(f"{k} {ops[k]}" for k in SHOW if ops.get(k)))


if __name__ == "__main__":
    main()
