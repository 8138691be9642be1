#!/usr/bin/env python3
"""Stress of the CTA-per-frame executors: many frames of very different kinds (text, long-window copies, periodic patterns with
tiny offsets, literal-heavy blocks with zero runs and raw blocks, low-entropy binary), each forced through k_exec_flow (narrow /
wide shape) or k_exec_big, several seeds, outputs compared with the originals and XXH64 with the trailer.
usage: [CZB_BIG_CLS=0 CZB_BIG_SEQ_BYTES=0 [CZB_FLOW_WIDE=1 | CZB_BIG_FLOW=0]] python scripts/flow_stress.py [seeds]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, workloads as W


def periodic(rng, n):
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    while pos < n:
        kind = rng.integers(0, 4)
        m = int(min(n - pos, rng.integers(50, 60000)))
        if kind == 0:      # one byte repeated (offset 1)
            out[pos:pos + m] = rng.integers(0, 256)
        elif kind == 1:    # short period
            p = rng.integers(0, 256, size=int(rng.integers(2, 40)), dtype=np.uint8)
            out[pos:pos + m] = np.resize(p, m)
        elif kind == 2:    # text
            out[pos:pos + m] = np.frombuffer(W.synth_text(m, int(rng.integers(1 << 30))), dtype=np.uint8)
        else:              # copy of something earlier, far or near
            if pos > 100:
                s = int(rng.integers(0, pos - 1))
                m = min(m, pos - s)
                out[pos:pos + m] = out[s:s + m]
            else:
                out[pos:pos + m] = rng.integers(32, 40, size=m, dtype=np.uint8)
        pos += m
    return out.tobytes()


def make(seed):
    rng = np.random.default_rng(seed)
    origs = []
    for _ in range(10):
        origs.append(W.synth_text(int(rng.integers(60000, 3 << 20)), int(rng.integers(1 << 30))))
    for _ in range(10):
        origs.append(periodic(rng, int(rng.integers(60000, 2 << 20))))
    f3, o3 = W.config3_literal_heavy(3, frame_size=1 << 20, seed=seed)
    f5, o5 = W.config5_mixed_sizes(12, seed=seed, lo=50000, hi=3 << 20)
    for _ in range(6):
        n = int(rng.integers(100000, 1 << 20))
        origs.append((rng.integers(0, 4, size=n, dtype=np.uint8) * 17).astype(np.uint8).tobytes())
    frames = []
    for k, o in enumerate(origs):
        cz = W.Compressor(level=int(rng.choice([1, 3, 3, 5, 9])), checksum=True, window_log=int(rng.choice([17, 20, 23])) if k % 3 == 0 else None)
        frames.append(cz.compress(o))
    if seed % 2 == 0:
        f4, o4 = W.config4_long_window(1, total=(7 << 20) + int(rng.integers(0, 1 << 20)), seed=seed)
        frames += f4; origs += o4
    return frames + f3 + f5, origs + o3 + o5


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    ctx = czb.Context(0)
    L = czb.load_library()
    t0 = time.time()
    total = 0
    for seed in range(100, 100 + seeds):
        frames, origs = make(seed)
        for rep in range(2):
            outs, res = ctx.decode_batch(frames, [len(o) for o in origs], api.FLAG_VERIFY_CHECKSUM)
            for i, (o, g, r) in enumerate(zip(origs, outs, res)):
                assert r.status == 0, (seed, i, czb.status_name(r.status))
                assert g == o, (seed, i, len(o))
                assert r.checksum_calculated == r.checksum_from_data, (seed, i)
            total += sum(len(o) for o in origs)
    wd = (C.c_uint32 * 16)()
    L.czb_debug_flow_watchdog(wd)
    assert wd[0] == 0, list(wd)
    print(f"flow stress ok: {seeds} seeds, {total / 1e6:.0f} MB decoded and verified, {time.time() - t0:.1f} s, env "
          + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CZB_")))


if __name__ == "__main__":
    main()
