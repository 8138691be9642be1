"""Python mirror of the reference's FrameDecoder traits (src/frame_decoder.cairo), over the C ABI.

Names, argument meaning and error behaviour follow the Cairo so that the parity tests read like
the reference's own tests (src/tests/decoding.cairo:4-21):

    source = ByteSlice(data)
    state = FrameDecoderState.new(source)          # advances `source` past the frame header
    dec = FrameDecoder.new(state)
    dec.decode_blocks(source, BlockDecodingStrategy.All())
    assert dec.is_finished()
    out = dec.collect()
    assert dec.get_checksum_from_data() == dec.get_calculated_checksum()

Cairo `Result::Err` becomes FrameDecoderError(status); `Option` becomes value-or-None.
"""
import ctypes as C

from .api import Context, load_library, status_name

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class FrameDecoderError(Exception):
    """FrameDecoderError (src/frame_decoder.cairo:39-48), flattened to a czs_status leaf code."""

    def __init__(self, status):
        super().__init__(status_name(status))
        self.status = int(status)


class ByteSlice:
    """`ref source: @ByteArraySlice` (src/utils/byte_array.cairo:9-13): a borrowed view the callee re-points."""

    def __init__(self, data: bytes, start: int = 0):
        self.data = bytes(data)
        self.pos = start

    def remaining(self) -> bytes:
        return self.data[self.pos:]

    def __len__(self):
        return len(self.data) - self.pos


class BlockDecodingStrategy:
    """src/frame_decoder.cairo:33-37"""

    def __init__(self, kind, n=0):
        self.kind, self.n = kind, n

    @staticmethod
    def All():
        return BlockDecodingStrategy(0)

    @staticmethod
    def UptoBlocks(n):
        return BlockDecodingStrategy(1, n)

    @staticmethod
    def UptoBytes(n):
        return BlockDecodingStrategy(2, n)


class FrameDecoderState:
    """FrameDecoderStateTrait (src/frame_decoder.cairo:52-106)."""

    def __init__(self, handle, ctx):
        self._h, self._ctx = handle, ctx

    @staticmethod
    def new(source: ByteSlice, ctx: Context = None) -> "FrameDecoderState":
        ctx = ctx or default_context()
        L = load_library()
        h, used = C.c_void_p(), C.c_uint64()
        rem = source.remaining()
        st = L.czb_fd_new(ctx.handle, rem, len(rem), C.byref(used), C.byref(h))
        if st != 0:
            raise FrameDecoderError(st)
        source.pos += used.value
        return FrameDecoderState(h, ctx)

    def reset(self, source: ByteSlice):
        used = C.c_uint64()
        rem = source.remaining()
        st = load_library().czb_fd_reset(self._h, rem, len(rem), C.byref(used))
        if st != 0:
            raise FrameDecoderError(st)
        source.pos += used.value

    def __del__(self):
        try:
            if self._h:
                load_library().czb_fd_free(self._h)
                self._h = None
        except Exception:
            pass


class FrameDecoder:
    """FrameDecoderTrait (src/frame_decoder.cairo:108-335)."""

    def __init__(self, state: FrameDecoderState):
        self.state = state
        self._L = load_library()

    @staticmethod
    def new(state: FrameDecoderState) -> "FrameDecoder":
        return FrameDecoder(state)

    def init(self, state: FrameDecoderState):
        self.reset(state)

    def reset(self, state: FrameDecoderState):
        self.state = state

    @property
    def _h(self):
        return self.state._h

    def content_size(self) -> int:
        return self._L.czb_fd_content_size(self._h)

    def get_checksum_from_data(self):
        v = C.c_uint32()
        return v.value if self._L.czb_fd_get_checksum_from_data(self._h, C.byref(v)) else None

    def get_calculated_checksum(self):
        v = C.c_uint32()
        return v.value if self._L.czb_fd_get_calculated_checksum(self._h, C.byref(v)) else None

    def bytes_read_from_source(self) -> int:
        return self._L.czb_fd_bytes_read_from_source(self._h)

    def is_finished(self) -> bool:
        return bool(self._L.czb_fd_is_finished(self._h))

    def blocks_decoded(self) -> int:
        return self._L.czb_fd_blocks_decoded(self._h)

    def decode_blocks(self, source: ByteSlice, strat: BlockDecodingStrategy) -> bool:
        used, fin = C.c_uint64(), C.c_int32()
        rem = source.remaining()
        st = self._L.czb_fd_decode_blocks(self._h, rem, len(rem), C.byref(used), strat.kind, strat.n, C.byref(fin))
        source.pos += used.value
        if st != 0:
            raise FrameDecoderError(st)
        return bool(fin.value)

    def can_collect(self) -> int:
        return self._L.czb_fd_can_collect(self._h)

    def collect(self):
        cap = max(self.can_collect(), 1)
        buf = C.create_string_buffer(cap)
        wrote = C.c_uint64()
        rc = self._L.czb_fd_collect(self._h, buf, cap, C.byref(wrote))
        if rc < 0:
            raise FrameDecoderError(-rc)
        return buf.raw[: wrote.value] if rc == 1 else None

    def decode_from_to(self, source: bytes, target: bytearray, cap: int = 1 << 26):
        buf = C.create_string_buffer(cap)
        rl, wr = C.c_uint64(), C.c_uint64()
        st = self._L.czb_fd_decode_from_to(self._h, source, len(source), buf, cap, C.byref(rl), C.byref(wr))
        if st != 0:
            raise FrameDecoderError(st)
        # `read` reports the amount it was asked to drain; the bytes actually present bound the copy
        target += buf.raw[: wr.value]
        return rl.value, wr.value

    def device_work(self):
        """(device runs, blocks executed on the device over all runs): an incremental decode executes every block once."""
        runs, blocks = C.c_uint64(), C.c_uint64()
        self._L.czb_debug_fd_device_work(self._h, C.byref(runs), C.byref(blocks))
        return runs.value, blocks.value

    def read(self, target: bytearray, cap: int = 1 << 26) -> int:
        buf = C.create_string_buffer(cap)
        n = self._L.czb_fd_read(self._h, buf, cap)
        if n < 0:
            raise FrameDecoderError(-n)
        target += buf.raw[:n]
        return n
