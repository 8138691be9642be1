"""-m gpu parity tests for the rest of the FrameDecoder handle surface:
FrameDecoderStateTrait::reset incl. the 100 MiB WindowSizeTooBig check (src/frame_decoder.cairo:78-105, :92-94),
FrameDecoderTrait::read (:328-334) and decode_from_to (:245-326) with small feeds, including the
RingBuffer::len quirk (src/decoding/ring_buffer.cairo:20-22), call by call against the oracle's restatement
of the same state machine (oracle_fd_*).  Everything goes through the C ABI (czb_fd_*)."""
import ctypes as C
import struct

import pytest

import oracle_lib as O
import cairo_zstd_b200 as czb
from cairo_zstd_b200.frame_decoder import ByteSlice
from gpu_common import sha

pytestmark = pytest.mark.gpu

CZS_WINDOW_SIZE_TOO_BIG = 11


def _idx(corpus, name):
    return next(k for k, e in enumerate(corpus.index) if e["name"] == name)


def _oracle_new(f, flags=0):
    L = O.lib()
    used, st = C.c_size_t(), C.c_int32()
    h = L.oracle_fd_new(f, len(f), C.byref(used), flags, C.byref(st))
    return h, used.value, st.value


def _getters_equal(dec, ofd):
    r = O.OracleResult()
    O.lib().oracle_fd_getters(ofd, C.byref(r))
    assert dec.blocks_decoded() == r.blocks_decoded
    assert dec.bytes_read_from_source() == r.bytes_read
    assert dec.content_size() == r.content_size
    assert dec.is_finished() == bool(r.finished)
    got = dec.get_checksum_from_data()
    assert (got is not None) == bool(r.has_checksum)
    if got is not None:
        assert got == r.checksum_from_data
    assert dec.get_calculated_checksum() == r.checksum_calculated


def test_fd_reset_reuses_the_handle_like_the_reference(corpus):
    """new(A) -> decode -> reset(B) -> decode -> reset(A) ...: every getter equals the oracle's after each call."""
    L = O.lib()
    names = ["z000001", "z000033", "z000062", "z000009", "z000035"]  # incl. an empty original and a single-segment frame
    frames = [corpus.frame(_idx(corpus, n)) for n in names]
    source = ByteSlice(frames[0])
    state = czb.FrameDecoderState.new(source)
    dec = czb.FrameDecoder.new(state)
    ofd, opos, st = _oracle_new(frames[0])
    assert st == 0 and source.pos == opos
    obuf = C.create_string_buffer(1 << 21)
    for k, f in enumerate(frames):
        if k:
            source = ByteSlice(f)
            state.reset(source)
            dec.reset(state)
            used = C.c_size_t()
            assert L.oracle_fd_reset(ofd, f, len(f), C.byref(used)) == 0
            opos = used.value
            assert source.pos == opos
            _getters_equal(dec, ofd)  # counters cleared by reset (:78-105)
        fin = dec.decode_blocks(source, czb.BlockDecodingStrategy.All())
        ofin, oused = C.c_int32(), C.c_size_t()
        assert L.oracle_fd_decode_blocks(ofd, f[opos:], len(f) - opos, C.byref(oused), 0, 0, C.byref(ofin)) == 0
        assert fin == bool(ofin.value) and source.pos == opos + oused.value
        assert dec.can_collect() == L.oracle_fd_can_collect(ofd)
        got = dec.collect()
        wrote = C.c_size_t()
        assert L.oracle_fd_collect(ofd, obuf, len(obuf), C.byref(wrote)) == 1
        assert got == obuf.raw[:wrote.value]
        assert sha(got) == corpus.index[_idx(corpus, names[k])]["orig_sha256"]
        _getters_equal(dec, ofd)
    L.oracle_fd_free(ofd)


def test_fd_reset_rejects_windows_above_100_mib_but_new_accepts_them(corpus):
    """frame_decoder.cairo:92-94: only `reset` has the 100 MiB limit; `new` (:54-76) does not."""
    L = O.lib()
    good = corpus.frame(_idx(corpus, "z000001"))
    # magic, descriptor 0 (no single segment, no checksum, no FCS field), window descriptor exponent 17 -> 128 MiB
    big = struct.pack("<IBB", 0xFD2FB528, 0x00, 17 << 3) + b"\x01\x00\x00"
    just_ok = struct.pack("<IBB", 0xFD2FB528, 0x00, (16 << 3) | 4) + b"\x01\x00\x00"  # 64 MiB * 1.5 = 96 MiB <= 100 MiB
    # new accepts both
    for f in (big, just_ok):
        s = ByteSlice(f)
        czb.FrameDecoderState.new(s)
        h, used, st = _oracle_new(f)
        assert st == 0 and s.pos == used == 6
        L.oracle_fd_free(h)
    # reset: too big -> WindowSizeTooBig on both sides; 96 MiB passes
    source = ByteSlice(good)
    state = czb.FrameDecoderState.new(source)
    ofd, _, st = _oracle_new(good)
    assert st == 0
    with pytest.raises(czb.FrameDecoderError) as ei:
        state.reset(ByteSlice(big))
    used = C.c_size_t()
    ost = L.oracle_fd_reset(ofd, big, len(big), C.byref(used))
    assert ei.value.status == ost == CZS_WINDOW_SIZE_TOO_BIG
    s2 = ByteSlice(just_ok)
    state.reset(s2)
    assert L.oracle_fd_reset(ofd, just_ok, len(just_ok), C.byref(used)) == 0 and s2.pos == used.value
    # every other header-level error comes out of reset exactly as out of new
    for bad in (b"", b"\x28\xb5\x2f", struct.pack("<I", 0x184D2A50) + b"\x04\x00\x00\x00abcd", struct.pack("<IB", 0xFD2FB528, 0x00),
                struct.pack("<IBB", 0xFD2FB528, 0xC0, 0x00)):
        with pytest.raises(czb.FrameDecoderError) as ei:
            state.reset(ByteSlice(bad))
        assert ei.value.status == L.oracle_fd_reset(ofd, bad, len(bad), C.byref(used))
    L.oracle_fd_free(ofd)


@pytest.mark.parametrize("name,feed", [("z000033", 4096), ("z000035", 20000), ("z000011", 1 << 20), ("z000001", 7)])
def test_decode_from_to_with_small_feeds_matches_oracle(corpus, name, feed):
    """decode_from_to (:245-326): the caller hands over whatever source bytes it has; the decoder consumes whole blocks
    only, appends what may leave the window to `target` and reports (bytes read, bytes written)."""
    L = O.lib()
    i = _idx(corpus, name)
    f = corpus.frame(i)
    source = ByteSlice(f)
    dec = czb.FrameDecoder.new(czb.FrameDecoderState.new(source))
    ofd, opos, st = _oracle_new(f)
    assert st == 0 and source.pos == opos
    pos = opos
    obuf = C.create_string_buffer(corpus.index[i]["orig_len"] + 64)
    total, ototal = bytearray(), bytearray()
    window = feed
    for _ in range(200000):
        chunk = f[pos:pos + window]
        target = bytearray()
        C.memset(obuf, 0, len(obuf))
        rl, wr = dec.decode_from_to(chunk, target, cap=len(obuf))
        orl, owr = C.c_size_t(), C.c_size_t()
        assert L.oracle_fd_decode_from_to(ofd, chunk, len(chunk), obuf, len(obuf), C.byref(orl), C.byref(owr)) == 0
        assert (rl, wr) == (orl.value, owr.value), (name, pos, rl, wr, orl.value, owr.value)
        # `read` reports the amount it asked for (RingBuffer::len ignores head): the bytes really drained can be fewer, so
        # both sides start from zeroed buffers and the whole reported span is compared
        assert bytes(target) == obuf.raw[:owr.value], (name, pos)
        total += target
        _getters_equal(dec, ofd)
        assert dec.can_collect() == L.oracle_fd_can_collect(ofd)
        pos += rl
        if dec.is_finished() and pos >= len(f):
            break
        if rl == 0:
            assert window < len(f) * 2 + 16, "no progress"
            window *= 2  # not even one whole block in the feed: hand over more
        else:
            window = feed
    assert pos == len(f)
    # incremental, not quadratic: however small the feeds, every block is executed on the device exactly once
    runs, blocks = dec.device_work()
    assert blocks == dec.blocks_decoded(), (runs, blocks, dec.blocks_decoded())
    # drain the rest the way a caller would: read() until nothing comes
    for _ in range(4):
        t = bytearray()
        C.memset(obuf, 0, len(obuf))
        n = dec.read(t, cap=len(obuf))
        on = L.oracle_fd_read(ofd, obuf, len(obuf))
        assert n == on and bytes(t) == obuf.raw[:on]
        total += t
        _getters_equal(dec, ofd)
    L.oracle_fd_free(ofd)
    if name == "z000001":  # output smaller than the window: nothing leaves before the end, then everything at once, in order
        assert sha(bytes(total[:corpus.index[i]["orig_len"]])) == corpus.index[i]["orig_sha256"]


def test_read_after_partial_decode_follows_the_ring_buffer_len_quirk(corpus):
    """read (:328-334) after UptoBlocks: drains only what lies beyond the window until the frame is finished; because
    RingBuffer::len ignores `head` (ring_buffer.cairo:20-22) a second read() asks for the same amount again."""
    L = O.lib()
    i = _idx(corpus, "z000033")  # 986 blocks, window smaller than the output
    f = corpus.frame(i)
    source = ByteSlice(f)
    dec = czb.FrameDecoder.new(czb.FrameDecoderState.new(source))
    ofd, opos, st = _oracle_new(f)
    obuf = C.create_string_buffer(corpus.index[i]["orig_len"] + 64)
    total = bytearray()
    strat = czb.BlockDecodingStrategy.UptoBlocks(97)
    for _ in range(100):
        fin = dec.decode_blocks(source, strat)
        ofin, oused = C.c_int32(), C.c_size_t()
        assert L.oracle_fd_decode_blocks(ofd, f[opos:], len(f) - opos, C.byref(oused), 1, 97, C.byref(ofin)) == 0
        opos += oused.value
        assert source.pos == opos and fin == bool(ofin.value)
        for _ in range(2):  # twice: the second call exercises the quirk
            t = bytearray()
            C.memset(obuf, 0, len(obuf))
            n = dec.read(t, cap=len(obuf))
            on = L.oracle_fd_read(ofd, obuf, len(obuf))
            assert n == on and bytes(t) == obuf.raw[:on]
            total += t
            _getters_equal(dec, ofd)
        if fin:
            break
    assert dec.is_finished()
    L.oracle_fd_free(ofd)
