"""Batch splitter, output sizing and multi-GPU sharding (SURVEY.md section 8 rows f2 and e).

CPU part: czb_split_frames_host over the 100 corpus frames concatenated with skippable frames interleaved
(src/frame.cairo:160-166), truncation and capacity behaviour, and the largest-first partitioner.  GPU part (-m gpu):
czb_frame_sizes_* gives orig_len for every corpus frame -- including the 50 that carry no Frame_Content_Size --
without executing it; the device splitter agrees with the host one; czb_decode_batch_multi over the visible devices."""
import ctypes as C
import struct

import numpy as np
import pytest

import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api, sharding


def _skippable(k, payload):
    return struct.pack("<II", 0x184D2A50 + (k & 15), len(payload)) + payload


def _concat(corpus):
    parts, offs = [], []
    pos = 0
    for i in range(len(corpus)):
        if i % 3 == 0:
            sk = _skippable(i, bytes([i & 255]) * (i % 7))
            parts.append(sk); pos += len(sk)
        f = corpus.frame(i)
        offs.append((pos, len(f)))
        parts.append(f); pos += len(f)
    parts.append(_skippable(5, b""))
    return b"".join(parts), offs


def test_split_host_finds_every_frame_and_skips_skippable_ones(corpus):
    buf, offs = _concat(corpus)
    st, spans, skipped, used = czb.split_frames(buf)
    assert st == 0 and used == len(buf) and skipped == 34 + 1
    assert [(s.offset, s.length) for s in spans] == offs
    n_fcs_less = 0
    for s, e in zip(spans, corpus.index):
        info = czb.frame_header_info(buf[s.offset:s.offset + s.length])
        assert s.window_size == info.window_size and s.fcs_present == info.fcs_present and s.has_checksum_flag == 1
        if s.fcs_present:
            assert s.content_size == e["orig_len"]
        else:
            n_fcs_less += 1
    assert n_fcs_less == 50  # (SURVEY Appendix B says 51; a direct count of the descriptors gives 50)


def test_split_host_capacity_truncation_and_garbage(corpus):
    buf, offs = _concat(corpus)
    # capacity: the walk stops at the frame that does not fit and can be resumed from `consumed`
    st, spans, skipped, used = czb.split_frames(buf, cap=10)
    assert st == 0 and len(spans) == 10 and used == offs[9][0] + offs[9][1]
    st2, more, _, used2 = czb.split_frames(buf[used:])
    assert st2 == 0 and len(more) == 90 and used + used2 == len(buf)
    # truncated last frame: the good ones are reported, the status says why the walk stopped, consumed = where
    cut = offs[57][0] + offs[57][1] // 2
    st, spans, _, used = czb.split_frames(buf[:cut])
    assert st == 100 and len(spans) == 57 and used <= offs[57][0]  # CZS_PANIC_TRUNCATED; a skippable frame may precede
    # garbage after a frame: BadMagicNumber (frame.cairo:168-170), consumed = start of the garbage
    f0 = corpus.frame(1)
    st, spans, _, used = czb.split_frames(f0 + b"\x01\x02\x03\x04\x05\x06\x07\x08")
    assert st == 7 and len(spans) == 1 and used == len(f0)
    assert czb.split_frames(b"")[0] == 0


def test_partitioner_is_largest_first_and_matches_the_python_one():
    rng = np.random.default_rng(3)
    costs = [int(x) for x in np.exp(rng.uniform(np.log(1024), np.log(4 << 20), size=3000))]
    for world in (1, 2, 3, 4, 8):
        shard_of, load = czb.partition_frames(costs, world)
        ref = sharding.partition_frames(costs, world)
        got = [sorted(i for i, s in enumerate(shard_of) if s == r) for r in range(world)]
        assert got == ref
        assert load == [sum(costs[i] for i in g) for g in got]
        assert max(load) - min(load) <= max(costs)  # LPT bound
    assert czb.partition_frames([], 4) == ([], [0, 0, 0, 0])


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_frame_sizes_without_execution_equal_orig_len_for_all_corpus_frames(corpus):
    from gpu_common import ctx
    frames = [corpus.frame(i) for i in range(len(corpus))]
    res = ctx().frame_sizes(frames)
    fcs_less = 0
    for r, e, f in zip(res, corpus.index, frames):
        assert r.status == 0, (e["name"], czb.status_name(r.status))
        assert r.bytes_written == e["orig_len"], e["name"]
        assert r.bytes_read == len(f) and r.finished == 1
        fcs_less += 0 if czb.frame_header_info(f).fcs_present else 1
    assert fcs_less == 50
    # sizes drive the layout of a real decode: exact capacities, nothing to spare
    outs, res2 = ctx().decode_batch(frames, [r.bytes_written for r in res], api.FLAG_VERIFY_CHECKSUM)
    for r, r2, e in zip(res, res2, corpus.index):
        assert r2.status == 0 and r2.bytes_written == r.bytes_written and r2.checksum_calculated == r2.checksum_from_data, e["name"]
    # malformed frames: the sequence-section / header errors show up as statuses, the batch is not poisoned
    bad = [frames[1][:len(frames[1]) // 2], b"", frames[5]]
    rb = ctx().frame_sizes(bad)
    assert rb[0].status != 0 and rb[1].status != 0 and rb[2].status == 0 and rb[2].bytes_written == corpus.index[5]["orig_len"]


@pytest.mark.gpu
def test_split_device_agrees_with_host(corpus):
    import torch
    from gpu_common import ctx
    buf, offs = _concat(corpus)
    L = czb.load_library()
    d_buf = torch.frombuffer(bytearray(buf), dtype=torch.uint8).cuda()
    cap = 128
    d_spans = torch.zeros(cap * C.sizeof(api.FrameSpan), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros(4, dtype=torch.int64, device="cuda")
    rc = L.czb_split_frames_device(ctx().handle, d_buf.data_ptr(), len(buf), d_spans.data_ptr(), cap, d_counts.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    n, skipped, used, st = [int(x) for x in d_counts.cpu()]
    assert (n, skipped, used, st) == (100, 35, len(buf), 0)
    spans = (api.FrameSpan * cap).from_buffer_copy(d_spans.cpu().numpy().tobytes())
    assert [(spans[i].offset, spans[i].length) for i in range(n)] == offs


@pytest.mark.gpu
def test_decode_batch_multi_over_the_visible_devices(corpus):
    import torch
    import oracle_lib as O
    nd = torch.cuda.device_count()
    frames = [corpus.frame(i) for i in range(len(corpus))]
    caps = [e["orig_len"] + 8 for e in corpus.index]
    for devices in ([0], list(range(min(nd, 4)))):
        m = czb.MultiContext(devices)
        outs, res, stats = m.decode_batch(frames, caps, api.FLAG_VERIFY_CHECKSUM)
        for f, cap, out, r, e in zip(frames, caps, outs, res, corpus.index):
            st, want, ores = O.decode_frame(f, dst_cap=cap)
            assert r.status == st == 0 and out == want, e["name"]
            assert r.checksum_calculated == ores.checksum_calculated == r.checksum_from_data
        assert sum(s.frames for s in stats) == len(frames) and [s.device for s in stats] == devices
        assert sum(s.bytes_out for s in stats) == sum(e["orig_len"] for e in corpus.index)
        if len(devices) > 1:
            loads = [s.bytes_in + s.bytes_out for s in stats]
            assert min(loads) > 0
        m.close()
