"""-m gpu: out-of-bounds evidence without compute-sanitizer (the tool is closed on this pool: profiles/r02_sanitizer_closed.log).

Two kinds of guard zones, both filled with a pattern before decoding and checked after it:
  * around every frame's output span in the caller's device buffer (the tests lay the spans out with 64-byte gaps and give the
    decoder EXACT capacities), catching any store outside [dst, dst + dst_cap);
  * behind the used part of the library's literal / sequence / block scratch (CZB_GUARD=1, czb_debug_guard_faults), catching a
    kernel that writes past its slice (k_huff's 16-byte literal stores, k_fse's 32-byte record stores).
Run over the corpus, the synthetic configs and mutated frames, through the warp-per-frame executor and with every frame
forced through the CTA-per-frame executors (k_exec_flow, and k_exec_big with CZB_BIG_FLOW=0)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, workloads as W

pytestmark = pytest.mark.gpu
GAP = 64
PAT = 0xA5


def _frames(corpus):
    frames = [corpus.frame(i) for i in range(len(corpus))]
    caps = [e["orig_len"] for e in corpus.index]
    for gen in (lambda: W.config2_text_frames(8, 65536), lambda: W.config3_literal_heavy(2, frame_size=1 << 19),
                lambda: W.config5_mixed_sizes(24, hi=1 << 20), lambda: W.small_alphabet_frames(8)):
        f, o = gen()
        frames += f; caps += [len(x) for x in o]
    # malformed frames too: a failing frame must not scribble outside its span either
    rng = np.random.default_rng(5)
    for i in (11, 33, 35, 97):
        b = bytearray(corpus.frame(i))
        for _ in range(3):
            b[int(rng.integers(8, len(b)))] ^= 1 << int(rng.integers(0, 8))
        frames.append(bytes(b)); caps.append(corpus.index[i]["orig_len"])
        frames.append(corpus.frame(i)[: len(corpus.frame(i)) * 2 // 3]); caps.append(corpus.index[i]["orig_len"])
    return frames, caps


@pytest.mark.parametrize("mode", ["default", "flow-forced", "flow-wide-forced", "big-forced"])
def test_no_store_leaves_its_span_and_scratch_guards_stay_intact(corpus, monkeypatch, mode):
    import torch
    monkeypatch.setenv("CZB_GUARD", "1")
    if mode != "default":
        monkeypatch.setenv("CZB_BIG_CLS", "0"); monkeypatch.setenv("CZB_BIG_SEQ_BYTES", "0")
    if mode == "big-forced":
        monkeypatch.setenv("CZB_BIG_FLOW", "0")
    if mode == "flow-wide-forced":
        monkeypatch.setenv("CZB_FLOW_WIDE", "1")
    ctx = czb.Context(0)
    frames, caps = _frames(corpus)
    n = len(frames)
    dev = torch.device("cuda", 0)
    flens = np.array([len(f) for f in frames], dtype=np.int64)
    soff = np.concatenate([[0], np.cumsum((flens + 15) & ~15)])
    src_h = np.zeros(int(soff[-1]) + 16, dtype=np.uint8)
    for i, f in enumerate(frames):
        src_h[soff[i]:soff[i] + len(f)] = np.frombuffer(f, dtype=np.uint8)
    src = torch.from_numpy(src_h).to(dev)
    caps_a = np.array(caps, dtype=np.int64)
    # spans at odd alignments on purpose: span i starts GAP + (i % 7) bytes after the previous one ends
    starts = np.zeros(n, dtype=np.int64)
    pos = GAP
    for i in range(n):
        starts[i] = pos
        pos += int(caps_a[i]) + GAP + (i % 7)
    dst = torch.full((pos + GAP,), PAT, dtype=torch.uint8, device=dev)
    d = np.zeros((n, 4), dtype=np.uint64)
    d[:, 0] = src.data_ptr() + soff[:-1]
    d[:, 1] = flens
    d[:, 2] = dst.data_ptr() + starts
    d[:, 3] = caps_a
    descs = torch.from_numpy(d.view(np.uint8).reshape(-1)).to(dev)
    results = torch.zeros(n * C.sizeof(api.FrameResult), dtype=torch.uint8, device=dev)
    for _ in range(2):  # twice: the second call reuses scratch that is exactly big enough
        ctx.decode_batch_device(descs.data_ptr(), results.data_ptr(), n, api.FLAG_VERIFY_CHECKSUM, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dst.cpu().numpy()
    res = (api.FrameResult * n).from_buffer_copy(results.cpu().numpy().tobytes())
    mask = np.ones(out.size, dtype=bool)
    for i in range(n):
        st, want, _ = O.decode_frame(frames[i], dst_cap=caps[i])
        assert res[i].status == st, (i, czb.status_name(st), czb.status_name(res[i].status))
        if st == 0:
            assert out[starts[i]:starts[i] + len(want)].tobytes() == want, i
        mask[starts[i]:starts[i] + caps[i]] = False
    assert (out[mask] == PAT).all(), f"{int((out[mask] != PAT).sum())} bytes outside the frames' spans were overwritten"
    assert ctx.guard_faults() == 0
    ctx.close()


def test_planned_call_needs_no_host_sync_and_can_be_graph_captured(corpus):
    """czb_plan_batch_device + czb_decode_batch_device_planned: same results as the ordinary call, no planning read-back; the
    planned call is captured into a CUDA graph and replayed."""
    import torch
    ctx = czb.Context(0)
    frames, caps = _frames(corpus)
    frames, caps = frames[:140], caps[:140]   # valid frames only (corpus + synthetic)
    n = len(frames)
    dev = torch.device("cuda", 0)
    flens = np.array([len(f) for f in frames], dtype=np.int64)
    soff = np.concatenate([[0], np.cumsum((flens + 15) & ~15)])
    src_h = np.zeros(int(soff[-1]) + 16, dtype=np.uint8)
    for i, f in enumerate(frames):
        src_h[soff[i]:soff[i] + len(f)] = np.frombuffer(f, dtype=np.uint8)
    src = torch.from_numpy(src_h).to(dev)
    caps_a = np.array(caps, dtype=np.int64)
    doff = np.concatenate([[0], np.cumsum((caps_a + 15) & ~15)])
    dst = torch.zeros(int(doff[-1]) + 64, dtype=torch.uint8, device=dev)
    d = np.zeros((n, 4), dtype=np.uint64)
    d[:, 0] = src.data_ptr() + soff[:-1]; d[:, 1] = flens; d[:, 2] = dst.data_ptr() + doff[:-1]; d[:, 3] = caps_a
    descs = torch.from_numpy(d.view(np.uint8).reshape(-1)).to(dev)
    res_a = torch.zeros(n * C.sizeof(api.FrameResult), dtype=torch.uint8, device=dev)
    res_b = torch.zeros_like(res_a)
    s = torch.cuda.Stream(dev)
    ctx.decode_batch_device(descs.data_ptr(), res_a.data_ptr(), n, api.FLAG_VERIFY_CHECKSUM, s.cuda_stream)
    s.synchronize()
    want = dst.clone()
    plan = ctx.plan_batch_device(descs.data_ptr(), n, s.cuda_stream)
    dst.zero_()
    ctx.decode_batch_device_planned(plan, descs.data_ptr(), res_b.data_ptr(), n, api.FLAG_VERIFY_CHECKSUM, s.cuda_stream)
    s.synchronize()
    assert torch.equal(dst, want) and torch.equal(res_a, res_b)
    res = (api.FrameResult * n).from_buffer_copy(res_b.cpu().numpy().tobytes())
    assert all(r.status == 0 and r.checksum_calculated == r.checksum_from_data for r in res)
    # graph capture + two replays
    g = torch.cuda.CUDAGraph()
    dst.zero_(); res_b.zero_()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s, capture_error_mode="relaxed"):
        ctx.decode_batch_device_planned(plan, descs.data_ptr(), res_b.data_ptr(), n, api.FLAG_VERIFY_CHECKSUM, s.cuda_stream)
    for _ in range(2):
        dst.zero_(); res_b.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(dst, want) and torch.equal(res_a, res_b)
    ctx.plan_destroy(plan)
    ctx.close()
