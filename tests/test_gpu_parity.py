"""-m gpu: the CUDA path, called through the C ABI, against the oracle and the golden fixtures."""
import ctypes as C
import struct

import numpy as np
import pytest

import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200 import api
from cairo_zstd_b200 import workloads as W
from cairo_zstd_b200.frame_decoder import ByteSlice
from gpu_common import compare_with_oracle, ctx, gpu_decode, sha

pytestmark = pytest.mark.gpu


# ---- config 1: the reference's corpus (src/tests/decoding.cairo:4-21) ----
def test_corpus_reference_test_set_bytes(corpus):
    idx = [i for i, e in enumerate(corpus.index) if e["in_reference_test_set"]]
    outs, res = gpu_decode([corpus.frame(i) for i in idx], [2048] * len(idx))
    for k, i in enumerate(idx):
        assert res[k].status == 0, (corpus.index[i]["name"], czb.status_name(res[k].status))
        assert outs[k] == corpus.small_original(i), corpus.index[i]["name"]
        assert res[k].finished == 1
        assert res[k].checksum_from_data == res[k].checksum_calculated == corpus.index[i]["trailer_xxh64_low32"]


def test_corpus_all_100_one_batch(corpus):
    frames = [corpus.frame(i) for i in range(len(corpus))]
    caps = [e["orig_len"] + 8 for e in corpus.index]
    outs, res = gpu_decode(frames, caps)
    bad = []
    for i, e in enumerate(corpus.index):
        ok = (res[i].status == 0 and len(outs[i]) == e["orig_len"] and sha(outs[i]) == e["orig_sha256"]
              and res[i].checksum_calculated == e["trailer_xxh64_low32"] and res[i].bytes_read == e["frame_len"])
        if not ok:
            bad.append((e["name"], czb.status_name(res[i].status), res[i].bytes_written, e["orig_len"]))
    assert not bad, bad[:10]


def test_corpus_against_oracle_fields(corpus):
    frames = [corpus.frame(i) for i in range(len(corpus))]
    caps = [e["orig_len"] + 8 for e in corpus.index]
    compare_with_oracle(frames, caps, label="corpus")


def test_corpus_exact_capacity_and_too_small(corpus):
    i = next(k for k, e in enumerate(corpus.index) if e["orig_len"] > 5000)
    f, n = corpus.frame(i), corpus.index[i]["orig_len"]
    outs, res = gpu_decode([f, f], [n, n - 1])
    assert res[0].status == 0 and sha(outs[0]) == corpus.index[i]["orig_sha256"]
    assert res[1].status == 102  # CZS_DST_TOO_SMALL


# ---- stage-level parity: literals and sequences against the oracle's trace ----
@pytest.mark.parametrize("name", ["z000000", "z000033", "z000097", "z000035"])
def test_stage_intermediates_match_oracle_trace(corpus, name):
    i = next(k for k, e in enumerate(corpus.index) if e["name"] == name)
    f = corpus.frame(i)
    outs, res = gpu_decode([f], [corpus.index[i]["orig_len"] + 8], flags=0)
    assert res[0].status == 0
    blocks, lits, seqs = ctx().debug_last_wave()
    st, out, ores, trace = O.decode_frame(f, dst_cap=corpus.index[i]["orig_len"] + 8, trace=True)
    gblocks = [b for b in blocks if b.block_type != 3]
    assert len(gblocks) == len(trace)
    for k, (g, t) in enumerate(zip(gblocks, trace)):
        assert g.block_type == t["block_type"], k
        if g.block_type != 2:
            continue
        assert (g.lit_type, g.regen_size, g.n_seq) == (t["lit_type"], t["regen_size"], t["n_seq"]), k
        if g.lit_type >= 2:
            got = lits[g.lit_off:g.lit_off + g.regen_size]
            assert got == t["lits"], f"block {k}: literals differ"
        for s in range(g.n_seq):
            ll, ml, off = seqs[3 * (g.seq_off + s):3 * (g.seq_off + s) + 3]
            assert (ll, ml) == t["seqs"][s][:2], (k, s)
            if off < 0xF0000000:  # symbolic offsets are resolved by k_exec
                assert off == t["seqs"][s][3], (k, s)


# ---- synthetic configs at sizes the oracle finishes quickly ----
def test_config2_text_frames_small_batch():
    frames, origs = W.config2_text_frames(96)
    outs, res = compare_with_oracle(frames, [len(o) for o in origs], label="config2")
    assert all(a == b for a, b in zip(outs, origs))


def test_config3_literal_heavy():
    frames, origs = W.config3_literal_heavy(3)
    outs, _ = compare_with_oracle(frames, [len(o) + 3 for o in origs], label="config3")
    assert all(a == b for a, b in zip(outs, origs))


def test_config4_long_window():
    frames, origs = W.config4_long_window(1, total=17 << 20)
    outs, res = compare_with_oracle(frames, [len(o) for o in origs], label="config4")
    assert outs[0] == origs[0] and res[0].window_size == 8 << 20


def test_config5_mixed_sizes():
    frames, origs = W.config5_mixed_sizes(40, hi=2 << 20)
    outs, _ = compare_with_oracle(frames, [len(o) + 1 for o in origs], label="config5")
    assert all(a == b for a, b in zip(outs, origs))


def test_direct_weight_headers_follow_rfc_order():
    frames, origs = W.small_alphabet_frames(60)
    outs, _ = compare_with_oracle(frames, [len(o) for o in origs], label="direct-weights")
    assert all(a == b for a, b in zip(outs, origs))


def test_huffman_streams_with_non_rfc_split_decode_like_the_reference():
    """The reference concatenates the four streams and checks only the total
    (literals_section_decoder.cairo:203-240), so any split that sums to regenerated_size decodes."""
    import handmade as H
    splits = [[[0, 1], [1, 1], [0, 0], [1, 0]], [[0, 1, 1], [1], [0, 0], [1, 0]],
              [[0, 1, 1, 1, 0, 1, 1, 0, 0, 1, 0], [1], [0], [1, 0, 1]], [[0, 1], [1, 1], [0, 0], [1, 0, 1]],
              [[1] * 40, [0] * 3, [1, 0] * 9, [0]]]
    built = [H.huf4_two_symbol_frame(s) for s in splits]
    outs, res = compare_with_oracle([f for f, _ in built], [64] * len(built), label="non-rfc-split")
    assert all(r.status == 0 for r in res)
    assert outs == [e for _, e in built]


def test_small_workspace_budget_splits_into_many_waves():
    """A 24 MB workspace forces the planner to shrink waves far below the batch; results must not depend on it."""
    frames, origs = W.config2_text_frames(24, 65536)
    frames, origs = frames * 30, origs * 30  # 720 frames, ~47 MB of output
    small = czb.Context(0, 24 << 20)
    outs, res = small.decode_batch(frames, [len(o) for o in origs], api.FLAG_VERIFY_CHECKSUM)
    assert all(r.status == 0 and r.checksum_calculated == r.checksum_from_data for r in res)
    assert outs == origs


def test_packed_host_path_chunks_and_error_frames(monkeypatch):
    """czb_decode_batch_host_packed (the e2e path of bench.py): several staging chunks, one broken frame in the middle."""
    import ctypes as C
    monkeypatch.setenv("CZB_HOST_CHUNK_MB", "4")
    frames, origs = W.config2_text_frames(32, 65536)
    frames, origs = list(frames) * 8, list(origs) * 8  # 256 frames, 16 MB out: several 4 MB chunks
    bad = len(frames) // 2
    frames[bad] = frames[bad][: len(frames[bad]) // 2]
    n = len(frames)
    src = np.frombuffer(b"".join(frames), dtype=np.uint8).copy()
    src_off = np.zeros(n + 1, dtype=np.uint64); src_off[1:] = np.cumsum([len(f) for f in frames])
    dst_off = np.zeros(n + 1, dtype=np.uint64); dst_off[1:] = np.cumsum([len(o) for o in origs])
    dst = np.zeros(int(dst_off[-1]), dtype=np.uint8)
    results = (api.FrameResult * n)()
    packed = czb.Context(0)
    packed.decode_batch_packed(src.ctypes.data, src_off, dst.ctypes.data, dst_off, n, C.addressof(results), api.FLAG_VERIFY_CHECKSUM)
    for k in range(n):
        if k == bad:
            assert results[k].status == O.decode_frame(frames[k], dst_cap=len(origs[k]))[0] != 0
            continue
        assert results[k].status == 0 and results[k].bytes_written == len(origs[k])
        assert dst[int(dst_off[k]):int(dst_off[k + 1])].tobytes() == origs[k], k


def test_tightly_packed_buffers_any_alignment(corpus):
    """The packed path keeps the caller's offsets on the device: corpus frames and outputs back to back, so that sources
    and destinations start at every alignment (unaligned match copies, flushes and the unaligned XXH64 read path)."""
    import ctypes as C
    frames = [corpus.frame(i) for i in range(len(corpus.index))]
    lens = [e["orig_len"] for e in corpus.index]
    n = len(frames)
    src = np.frombuffer(b"\x00" + b"".join(frames), dtype=np.uint8).copy()
    src_off = np.ones(n + 1, dtype=np.uint64); src_off[1:] += np.cumsum([len(f) for f in frames]).astype(np.uint64)
    dst_off = np.full(n + 1, 3, dtype=np.uint64); dst_off[1:] += np.cumsum(lens).astype(np.uint64)
    dst = np.zeros(int(dst_off[-1]) + 8, dtype=np.uint8)
    results = (api.FrameResult * n)()
    ctx().decode_batch_packed(src.ctypes.data, src_off, dst.ctypes.data, dst_off, n, C.addressof(results), api.FLAG_VERIFY_CHECKSUM)
    assert len({int(o) & 7 for o in dst_off[:-1]}) == 8
    for k in range(n):
        assert results[k].status == 0 and results[k].bytes_written == lens[k], k
        assert sha(dst[int(dst_off[k]):int(dst_off[k + 1])].tobytes()) == corpus.index[k]["orig_sha256"], k
        assert results[k].has_checksum and results[k].checksum_calculated == results[k].checksum_from_data, k


def test_cta_per_frame_executor_forced_on_everything(corpus, monkeypatch):
    """k_exec_big normally takes only large frames with sparse sequences; force every frame through it (valid, multi-block,
    raw/RLE blocks, malformed) and compare with the oracle, status codes included."""
    monkeypatch.setenv("CZB_BIG_CLS", "0")
    monkeypatch.setenv("CZB_BIG_SEQ_BYTES", "0")
    big = czb.Context(0)
    frames = [corpus.frame(i) for i in range(len(corpus.index))]
    caps = [e["orig_len"] + 16 for e in corpus.index]
    f2, o2 = W.config2_text_frames(8, 65536)
    f3, o3 = W.config3_literal_heavy(3)
    f5, o5 = W.config5_mixed_sizes(10, hi=1 << 20)
    for fs, os_ in ((f2, o2), (f3, o3), (f5, o5)):
        frames += list(fs); caps += [len(o) for o in os_]
    rng = np.random.default_rng(3)
    for i in (5, 20, 43, 63):
        f = corpus.frame(i)
        for _ in range(10):
            b = bytearray(f); pos = int(rng.integers(4, len(b))); b[pos] ^= 1 << int(rng.integers(0, 8))
            frames.append(bytes(b)); caps.append(4 * corpus.index[i]["orig_len"] + 64)
        frames.append(f[: len(f) // 2]); caps.append(corpus.index[i]["orig_len"] + 64)
    outs, res = big.decode_batch(frames, caps, api.FLAG_VERIFY_CHECKSUM)
    for k, f in enumerate(frames):
        st, want, r = O.decode_frame(f, dst_cap=caps[k])
        assert res[k].status == st, (k, czb.status_name(st), czb.status_name(res[k].status))
        if st == 0:
            assert outs[k] == want and res[k].bytes_read == r.bytes_read and res[k].blocks_decoded == r.blocks_decoded


def _giant_and_mixed_frames():
    """Frames whose chunks mix tiny and very long segments: periodic data (one self-overlapping match of ~100 KiB per
    block, offsets 1, 2 and 1000), long literal runs followed by long repeats, and text with 4 KiB repeats spliced in
    (so that tile-sized chunks and long ones alternate inside a block)."""
    rng = np.random.default_rng(11)
    cz = W.Compressor()
    blob = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
    text = W.synth_text(300000, 7)
    spliced = bytearray()
    pos = 0
    while pos < len(text) - 9000:
        spliced += text[pos:pos + 3000]
        spliced += text[max(0, pos - 20000):max(0, pos - 20000) + 4096]  # a long repeat of older material
        pos += 3000
    origs = [b"a" * 300000, b"ab" * 150000, blob * 300, blob + b"\x00" * 5000 + blob * 40 + bytes(rng.integers(0, 256, 70000, dtype=np.uint8)) + blob * 100,
             bytes(spliced), bytes(spliced[:70000]) * 4]
    return [cz.compress(o) for o in origs], origs


def test_giant_sequences_and_mixed_chunk_lengths():
    frames, origs = _giant_and_mixed_frames()
    outs, _ = compare_with_oracle(frames, [len(o) for o in origs], label="giant")
    for o, w in zip(outs, origs):
        assert o == w


def test_giant_sequences_through_the_cta_per_frame_executor(monkeypatch):
    monkeypatch.setenv("CZB_BIG_CLS", "0")
    monkeypatch.setenv("CZB_BIG_SEQ_BYTES", "0")
    big = czb.Context(0)
    frames, origs = _giant_and_mixed_frames()
    outs, res = big.decode_batch(frames, [len(o) for o in origs], api.FLAG_VERIFY_CHECKSUM)
    for k, o in enumerate(origs):
        assert res[k].status == 0 and outs[k] == o, k
        assert res[k].checksum_calculated == res[k].checksum_from_data


def test_large_frames_inside_a_large_batch_of_small_ones():
    """More frames than k_exec_big has resident CTAs: only the frames well above the mean go to it (share rule), the rest
    to k_exec, side by side on two streams; every output is checked."""
    cz = W.Compressor()
    small = [W.synth_text(4096, 1000 + i) for i in range(1200)]
    large = [W.synth_text(512 * 1024, 5000 + i) for i in range(3)]
    origs = small[:600] + large[:1] + small[600:] + large[1:]
    frames = [cz.compress(o) for o in origs]
    outs, res = gpu_decode(frames, [len(o) for o in origs])
    for k, o in enumerate(origs):
        assert res[k].status == 0 and outs[k] == o, k
        assert res[k].checksum_calculated == res[k].checksum_from_data, k


def test_empty_and_tiny_frames(corpus):
    cz = W.Compressor()
    origs = [b"", b"a", b"ab" * 3, b"\x00" * 70000, bytes(range(256)) * 3]
    frames = [cz.compress(o) for o in origs]
    outs, _ = compare_with_oracle(frames, [max(len(o), 1) + 4 for o in origs], label="tiny")
    assert outs == origs
    empties = [i for i, e in enumerate(corpus.index) if e["orig_len"] == 0]
    assert len(empties) == 3
    outs, res = gpu_decode([corpus.frame(i) for i in empties], [0, 1, 16])
    assert all(r.status == 0 and r.bytes_written == 0 for r in res)


def test_frames_without_checksum_or_fcs():
    cz = W.Compressor(checksum=False)
    o = W.synth_text(30000, 5)
    f = cz.compress(o)
    outs, res = compare_with_oracle([f], [len(o)], label="nochk")
    assert outs[0] == o and res[0].has_checksum == 0 and res[0].finished == 1


# ---- malformed input: per-frame status, the batch is not poisoned ----
def test_truncations_and_bit_flips_match_oracle_status(corpus):
    rng = np.random.default_rng(11)
    frames, caps = [], []
    for i in (1, 5, 20, 43, 63):
        f = corpus.frame(i)
        n = corpus.index[i]["orig_len"] + 64
        for cut in (len(f) - 1, len(f) - 4, len(f) - 5, len(f) // 2, 7, 5, 3):
            if 0 < cut < len(f):
                frames.append(f[:cut]); caps.append(n)
        for _ in range(12):
            b = bytearray(f)
            pos = int(rng.integers(4, len(b)))
            b[pos] ^= 1 << int(rng.integers(0, 8))
            frames.append(bytes(b)); caps.append(n * 4)
        frames.append(f); caps.append(n)
    outs, res = gpu_decode(frames, caps)
    mism = []
    for k, f in enumerate(frames):
        st, want, _ = O.decode_frame(f, dst_cap=caps[k])
        if (st == 0) != (res[k].status == 0) or (st == 0 and outs[k] != want):
            mism.append((k, czb.status_name(st), czb.status_name(res[k].status)))
    assert not mism, mism[:10]
    # leaf status codes are identical (see also tests/test_gpu_fuzz.py)
    diff = [(k, czb.status_name(O.decode_frame(f, dst_cap=caps[k])[0]), czb.status_name(res[k].status)) for k, f in enumerate(frames)
            if O.decode_frame(f, dst_cap=caps[k])[0] != res[k].status]
    assert not diff, diff[:10]


def test_header_level_errors_in_batch(corpus):
    good = corpus.frame(1)
    bad = [b"", b"\x00\x01\x02\x03\x04", struct.pack("<II", 0x184D2A50, 4) + b"abcd", struct.pack("<IB", 0xFD2FB528, 0xE0)]
    frames = [good] + bad + [good]
    outs, res = gpu_decode(frames, [4096] * len(frames))
    assert res[0].status == 0 and res[-1].status == 0 and outs[0] == outs[-1]
    for k, b in enumerate(bad):
        assert res[1 + k].status == O.decode_frame(b)[0]


# ---- FrameDecoder mirror: reads like src/tests/decoding.cairo:4-21 ----
def _test_decode(source_bytes, expected):
    source = ByteSlice(source_bytes)
    state = czb.FrameDecoderState.new(source)
    frame_decoder = czb.FrameDecoder.new(state)
    frame_decoder.decode_blocks(source, czb.BlockDecodingStrategy.All())
    assert frame_decoder.is_finished(), "not finished"
    result = frame_decoder.collect()
    assert frame_decoder.get_checksum_from_data() == frame_decoder.get_calculated_checksum(), "checksums do not match"
    assert result == expected, "wrong decoding result"
    assert len(source) == 0


def test_frame_decoder_mirror_on_reference_test_set(corpus):
    for i, e in enumerate(corpus.index):
        if e["in_reference_test_set"]:
            _test_decode(corpus.frame(i), corpus.small_original(i))


def test_frame_decoder_incremental_matches_oracle(corpus):
    """UptoBlocks / UptoBytes + collect()/can_collect() (frame_decoder.cairo:202-243) against the oracle's
    restatement of the same state machine, call by call."""
    L = O.lib()
    for name, strat_kind, n in (("z000033", 1, 7), ("z000035", 2, 300000), ("z000011", 1, 1)):
        i = next(k for k, e in enumerate(corpus.index) if e["name"] == name)
        f = corpus.frame(i)
        source = ByteSlice(f)
        dec = czb.FrameDecoder.new(czb.FrameDecoderState.new(source))
        used = C.c_size_t(); st = C.c_int32()
        ofd = L.oracle_fd_new(f, len(f), C.byref(used), 0, C.byref(st))
        opos = used.value
        obuf = C.create_string_buffer(corpus.index[i]["orig_len"] + 8)
        total = b""
        strat = czb.BlockDecodingStrategy(strat_kind, n)
        for rounds in range(100000):
            fin = dec.decode_blocks(source, strat)
            ofin = C.c_int32(); oused = C.c_size_t()
            assert L.oracle_fd_decode_blocks(ofd, f[opos:], len(f) - opos, C.byref(oused), strat_kind, n, C.byref(ofin)) == 0
            opos += oused.value
            assert source.pos == opos and fin == bool(ofin.value)
            assert dec.can_collect() == L.oracle_fd_can_collect(ofd)
            got = dec.collect()
            wrote = C.c_size_t()
            orc = L.oracle_fd_collect(ofd, obuf, len(obuf), C.byref(wrote))
            assert (got is not None) == (orc == 1)
            if got is not None:
                assert got == obuf.raw[:wrote.value]
                total += got
            r = O.OracleResult(); L.oracle_fd_getters(ofd, C.byref(r))
            assert dec.blocks_decoded() == r.blocks_decoded and dec.bytes_read_from_source() == r.bytes_read
            assert dec.get_calculated_checksum() == r.checksum_calculated
            if fin:
                break
        assert sha(total) == corpus.index[i]["orig_sha256"]
        assert dec.is_finished() and dec.get_checksum_from_data() == dec.get_calculated_checksum()
        L.oracle_fd_free(ofd)
