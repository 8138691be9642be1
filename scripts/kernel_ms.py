#!/usr/bin/env python3
"""Per-kernel-class milliseconds for one 131072-frame wave of the config-2 shape, WITHOUT checking results.
Only for what-if builds that deliberately compute something wrong (e.g. a kernel with a stage compiled out)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, workloads as W

dev = torch.device("cuda", 0)
frames, origs = W.config2_text_frames(512, 65536)
n0, reps = len(frames), 256
n = n0 * reps
flens = np.array([len(f) for f in frames], dtype=np.int64)
foff = np.concatenate([[0], np.cumsum((flens + 15) & ~15)])
host = np.zeros(int(foff[-1]), dtype=np.uint8)
for i, f in enumerate(frames):
    host[foff[i]:foff[i] + len(f)] = np.frombuffer(f, dtype=np.uint8)
src = torch.from_numpy(host).to(dev).repeat(reps)
dst = torch.empty(n * 65536, dtype=torch.uint8, device=dev)
idx = np.arange(n)
d = np.zeros((n, 4), dtype=np.uint64)
d[:, 0] = src.data_ptr() + (idx // n0) * int(foff[-1]) + foff[idx % n0]
d[:, 1] = flens[idx % n0]
d[:, 2] = dst.data_ptr() + idx * 65536
d[:, 3] = 65536
descs = torch.from_numpy(d.view(np.uint8).reshape(-1)).to(dev)
results = torch.zeros(n * C.sizeof(api.FrameResult), dtype=torch.uint8, device=dev)
ctx = czb.Context(0)
st = torch.cuda.current_stream(dev)
ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, 0, st.cuda_stream)
torch.cuda.synchronize()
ctx.profile_enable(True)
K = 3
for _ in range(K):
    ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, 0, st.cuda_stream)
torch.cuda.synchronize()
prof = ctx.profile_collect()
print(" ".join(f"{k}={v[0] / K:.2f}" for k, v in prof.items()))
