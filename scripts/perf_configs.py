#!/usr/bin/env python3
"""Device-timed throughput of the other BASELINE configs (3: literal-heavy, 4: long window, 5: mixed sizes),
each checked against the original bytes.  Parity-test configs, not bench lines (bench.py is config 2)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, workloads as W


def run(name, frames, origs, reps, ctx, dev):
    n0 = len(frames)
    n = n0 * reps
    flens = np.array([len(f) for f in frames], dtype=np.int64)
    olens = np.array([len(o) for o in origs], dtype=np.int64)
    foff = np.concatenate([[0], np.cumsum((flens + 15) & ~15)])
    ooff = np.concatenate([[0], np.cumsum((olens + 15) & ~15)])
    host = np.zeros(int(foff[-1]), dtype=np.uint8)
    for i, f in enumerate(frames):
        host[foff[i]:foff[i] + len(f)] = np.frombuffer(f, dtype=np.uint8)
    src = torch.from_numpy(host).to(dev).repeat(reps)
    dst = torch.empty(int(ooff[-1]) * reps, dtype=torch.uint8, device=dev)
    idx = np.arange(n)
    d = np.zeros((n, 4), dtype=np.uint64)
    d[:, 0] = src.data_ptr() + (idx // n0) * int(foff[-1]) + foff[idx % n0]
    d[:, 1] = flens[idx % n0]
    d[:, 2] = dst.data_ptr() + (idx // n0) * int(ooff[-1]) + ooff[idx % n0]
    d[:, 3] = olens[idx % n0]
    descs = torch.from_numpy(d.view(np.uint8).reshape(-1)).to(dev)
    results = torch.zeros(n * C.sizeof(api.FrameResult), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    for _ in range(2):
        ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, 0, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    K = 3
    for _ in range(K):
        ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, 0, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ctx.profile_enable(True)  # per-kernel-class times of one more pass (events around every launch)
    ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, 0, st.cuda_stream)
    torch.cuda.synchronize()
    kernel_ms = {k: round(v[0], 3) for k, v in ctx.profile_collect().items()}
    ctx.profile_enable(False)
    res = results.cpu().numpy().view(np.dtype([("status", "<i4"), ("b", "<u4"), ("br", "<u8"), ("bw", "<u8"), ("cs", "<u8"), ("w", "<u8"),
                                               ("c1", "<u4"), ("c2", "<u4"), ("h", "<i4"), ("f", "<i4")]))
    nocheck = bool(os.environ.get("CZB_PERF_NOCHECK"))  # what-if builds that compute something wrong on purpose
    assert nocheck or (res["status"] == 0).all(), name
    for k in (() if nocheck else (0, n - 1)):
        o = int(d[k, 2] - dst.data_ptr())
        assert dst[o:o + int(olens[k % n0])].cpu().numpy().tobytes() == origs[k % n0], (name, k)
    out = {"config": name, "frames": n, "out_GB": float(olens.sum() * reps / 1e9), "ratio": float(olens.sum() / flens.sum()),
           "ms": ms, "GBps": float(olens.sum() * reps / ms / 1e6), "kernel_ms": kernel_ms}
    print(json.dumps(out), flush=True)
    return out


def main():
    dev = torch.device("cuda", 0)
    ctx = czb.Context(0)
    outs = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""  # "3", "4" or "5": just that configuration
    if only in ("", "3"):
        f, o = W.config3_literal_heavy(16)
        outs.append(run("config3 literal-heavy 1 MiB frames", f, o, 64, ctx, dev))
    if only in ("", "5"):
        f, o = W.config5_mixed_sizes(512, hi=4 << 20)
        outs.append(run("config5 mixed 1 KiB..4 MiB", f, o, 16, ctx, dev))
    if only == "6":  # not a BASELINE config: a mid-size batch of equal frames, to check the k_exec_big share rule
        f, o = W.config2_text_frames(64, 262144)
        outs.append(run("equal 256 KiB text frames", f, o, int(os.environ.get("CZB_PERF_REPS", "64")), ctx, dev))
    if only == "4b":  # the same long-window frames as a batch large enough to fill the machine (throughput rather than per-frame latency)
        f, o = W.config4_long_window(2, total=17 << 20)
        outs.append(run("config4 long window 17 MiB frames, 512-frame batch", f, o, 256, ctx, dev))
    if only in ("", "4"):
        f, o = W.config4_long_window(2, total=17 << 20)
        outs.append(run("config4 long window 17 MiB frames", f, o, 32, ctx, dev))
    json.dump(outs, open(os.path.join(ROOT, "gpurun_out", "perf_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
