#!/usr/bin/env python3
"""Status-code parity on malformed input: GPU (C ABI) vs the oracle, over truncations, bit flips and byte stomps of
corpus and synthetic frames.  Prints a confusion summary of (oracle status, gpu status) pairs that differ."""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, workloads as W

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
golden = os.path.join(ROOT, "tests", "golden")
index = json.load(open(os.path.join(golden, "corpus_index.json")))
blob = open(os.path.join(golden, "corpus_frames.bin"), "rb").read()
base = [(blob[e["frame_off"]:e["frame_off"] + e["frame_len"]], e["orig_len"]) for e in index if e["orig_len"] <= 200000]
f2, o2 = W.config2_text_frames(6, 20000)
base += [(f, len(o)) for f, o in zip(f2, o2)]
f3, o3 = W.small_alphabet_frames(10)
base += [(f, len(o)) for f, o in zip(f3, o3)]
frames, caps = [], []
for f, n in base:
    if len(f) < 12:
        continue
    for _ in range(6):
        b = bytearray(f)
        kind = rng.integers(0, 4)
        if kind == 0:
            b = b[:int(rng.integers(1, len(b)))]
        elif kind == 1:
            pos = int(rng.integers(4, len(b))); b[pos] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            pos = int(rng.integers(4, len(b))); b[pos] = int(rng.integers(0, 256))
        else:
            pos = int(rng.integers(4, min(len(b), 40))); b[pos] ^= 0xFF
        frames.append(bytes(b)); caps.append(4 * n + 4096)
ctx = czb.Context(0)
outs, res = ctx.decode_batch(frames, caps, 0)
conf = collections.Counter()
okmis = 0
same = 0
for i, f in enumerate(frames):
    st, want, _ = O.decode_frame(f, dst_cap=caps[i])
    g = res[i].status
    if (st == 0) != (g == 0) or (st == 0 and outs[i] != want):
        okmis += 1
        conf[(czb.status_name(st), czb.status_name(g), "OK-MISMATCH")] += 1
    elif st == g:
        same += 1
    else:
        conf[(czb.status_name(st), czb.status_name(g))] += 1
print(f"{len(frames)} malformed/valid frames: identical status {same} ({100.0 * same / len(frames):.1f} %), ok/not-ok mismatches {okmis}")
for k, v in conf.most_common(25):
    print(v, k)
