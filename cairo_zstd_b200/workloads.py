"""Seeded input generators for the BASELINE.json configs.

Compression uses the host's libzstd (runtime library only, loaded with ctypes)
strictly to *produce inputs*; it is never used to decode on the product path.
All generators are deterministic in (seed, parameters).
"""
import ctypes as C
import numpy as np

# Advanced-API parameter ids (zstd.h is not installed)
ZSTD_c_compressionLevel = 100
ZSTD_c_windowLog = 101
ZSTD_c_checksumFlag = 201

_z = None


def libzstd():
    global _z
    if _z is None:
        z = C.CDLL("libzstd.so.1")
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_compress2.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_decompress.restype = C.c_size_t
        _z = z
    return _z


class Compressor:
    """One reusable ZSTD_CCtx: level, checksum flag and optional windowLog."""

    def __init__(self, level=3, checksum=True, window_log=None):
        z = libzstd()
        self.z = z
        self.ctx = z.ZSTD_createCCtx()
        z.ZSTD_CCtx_setParameter(self.ctx, ZSTD_c_compressionLevel, level)
        z.ZSTD_CCtx_setParameter(self.ctx, ZSTD_c_checksumFlag, 1 if checksum else 0)
        if window_log is not None:
            z.ZSTD_CCtx_setParameter(self.ctx, ZSTD_c_windowLog, window_log)
        self._buf = None

    def compress(self, data) -> bytes:
        data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
        bound = self.z.ZSTD_compressBound(len(data))
        if self._buf is None or len(self._buf) < bound:
            self._buf = C.create_string_buffer(bound)
        src = (C.c_char * len(data)).from_buffer_copy(data) if len(data) else None
        n = self.z.ZSTD_compress2(self.ctx, self._buf, bound, src, len(data))
        if self.z.ZSTD_isError(n):
            raise RuntimeError("ZSTD_compress2 failed")
        return self._buf.raw[:n]

    def __del__(self):
        try:
            self.z.ZSTD_freeCCtx(self.ctx)
        except Exception:
            pass


def libzstd_decompress(frame: bytes, cap: int) -> bytes:
    """Independent cross-check only (tests); never on the product path."""
    z = libzstd()
    dst = C.create_string_buffer(max(cap, 1))
    n = z.ZSTD_decompress(dst, cap, frame, len(frame))
    if z.ZSTD_isError(n):
        raise RuntimeError("ZSTD_decompress failed")
    return dst.raw[:n]


# ---------------------------------------------------------------------------
# raw material
# ---------------------------------------------------------------------------
def _vocabulary(rng, n_words=4096, max_len=11):
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    p = 1.0 / (np.arange(len(letters)) + 2.0)
    p /= p.sum()
    lens = rng.integers(2, max_len, size=n_words)
    table = np.full((n_words, max_len + 1), ord(" "), dtype=np.uint8)
    chars = rng.choice(letters, size=(n_words, max_len), p=p)
    for i in range(max_len):
        col = chars[:, i]
        table[:, i] = np.where(i < lens, col, ord(" "))
    return table, lens


def synth_text(n_bytes: int, seed: int) -> bytes:
    """Pseudo-English: a few thousand words from a skewed letter distribution, Zipf word
    frequencies, joined by spaces (occasional punctuation/newlines).  ASCII only, so the
    direct-weight Huffman header never appears."""
    rng = np.random.default_rng(seed)
    table, lens = _vocabulary(rng)
    n_words = table.shape[0]
    pw = 1.0 / (np.arange(n_words) + 2.7)
    pw /= pw.sum()
    need = n_bytes // 4 + 64
    ids = rng.choice(n_words, size=need, p=pw)
    rows = table[ids]
    width = lens[ids] + 1
    mask = np.arange(table.shape[1])[None, :] < width[:, None]
    # sprinkle punctuation at ~1/12 of the word ends
    punct = rng.random(need) < (1.0 / 12.0)
    pc = rng.choice(np.frombuffer(b".,;\n", dtype=np.uint8), size=need)
    rows = rows.copy()
    idx = np.nonzero(punct)[0]
    rows[idx, lens[ids[idx]]] = pc[idx]
    out = rows[mask]
    while out.size < n_bytes:
        out = np.concatenate([out, out])
    return out[:n_bytes].tobytes()


def skewed_bytes(n_bytes: int, seed: int, lo=32, n_sym=64) -> np.ndarray:
    """i.i.d. bytes from a geometric-ish distribution over [lo, lo+n_sym): Huffman-friendly,
    almost no matches (literal-heavy config)."""
    rng = np.random.default_rng(seed)
    p = 0.93 ** np.arange(n_sym)
    p /= p.sum()
    return (rng.choice(n_sym, size=n_bytes, p=p) + lo).astype(np.uint8)


# ---------------------------------------------------------------------------
# config builders: each returns (frames: list[bytes], originals: list[bytes])
# ---------------------------------------------------------------------------
def config2_text_frames(n_distinct: int, frame_size: int = 65536, seed: int = 1234, level: int = 3):
    """BASELINE config 2: independent frames of `frame_size` bytes of synthetic text, level 3,
    checksum flag on."""
    cz = Compressor(level=level, checksum=True)
    frames, origs = [], []
    # one big text per 64 frames, then sliced: keeps generation fast and frames distinct
    group = 64
    for g in range(0, n_distinct, group):
        k = min(group, n_distinct - g)
        text = synth_text(k * frame_size, seed + g)
        for i in range(k):
            o = text[i * frame_size:(i + 1) * frame_size]
            origs.append(o)
            frames.append(cz.compress(o))
    return frames, origs


def config3_literal_heavy(n_frames: int, frame_size: int = 1 << 20, seed: int = 77, level: int = 3):
    """BASELINE config 3: >= 256 KiB frames so blocks are 128 KiB; skewed bytes (4-stream Huffman,
    later blocks Treeless), with spliced zero runs (RLE blocks) and PRNG bytes (Raw blocks)."""
    cz = Compressor(level=level, checksum=True)
    frames, origs = [], []
    for f in range(n_frames):
        rng = np.random.default_rng(seed + 1000 * f)
        data = skewed_bytes(frame_size, seed + f)
        blk = 128 * 1024
        nblk = frame_size // blk
        if nblk >= 4:
            z = int(rng.integers(1, nblk - 1))
            data[z * blk:(z + 1) * blk] = 0  # -> RLE block
            r = int(rng.integers(1, nblk - 1))
            if r == z:
                r = (r % (nblk - 2)) + 1
            if r != z:
                data[r * blk:(r + 1) * blk] = rng.integers(0, 256, size=blk, dtype=np.uint8)  # -> Raw block
        # a short text island so at least one block has a real sequence section
        t = np.frombuffer(synth_text(4096, seed + f), dtype=np.uint8)
        data[100:100 + t.size] = t
        o = data.tobytes()
        origs.append(o)
        frames.append(cz.compress(o))
    return frames, origs


def config4_long_window(n_frames: int = 1, total: int = 17 << 20, seed: int = 5, window_log: int = 23, level: int = 3):
    """BASELINE config 4: windowLog 23 (8 MiB window); later regions repeat material ~6 MiB
    earlier, so matches reach far back and repeat offsets are common."""
    cz = Compressor(level=level, checksum=True, window_log=window_log)
    frames, origs = [], []
    for f in range(n_frames):
        rng = np.random.default_rng(seed + f)
        base = np.frombuffer(synth_text(6 << 20, seed + 31 * f), dtype=np.uint8)
        data = np.empty(total, dtype=np.uint8)
        data[:base.size] = base
        pos = base.size
        while pos < total:
            n = int(min(total - pos, rng.integers(200, 4000)))
            if rng.random() < 0.8:
                src = pos - (6 << 20) + int(rng.integers(-65536, 65536))
                src = max(0, min(src, pos - n))
                data[pos:pos + n] = data[src:src + n]
            else:
                data[pos:pos + n] = np.frombuffer(synth_text(n, int(rng.integers(1 << 30))), dtype=np.uint8)
            pos += n
        o = data.tobytes()
        origs.append(o)
        frames.append(cz.compress(o))
    return frames, origs


def config5_mixed_sizes(n_frames: int, seed: int = 99, lo: int = 1024, hi: int = 4 << 20, level: int = 3):
    """BASELINE config 5: frame sizes log-uniform in [1 KiB, 4 MiB], text, level 3."""
    rng = np.random.default_rng(seed)
    cz = Compressor(level=level, checksum=True)
    sizes = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n_frames)).astype(np.int64)
    frames, origs = [], []
    for i, s in enumerate(sizes):
        o = synth_text(int(s), seed + 7 * i)
        origs.append(o)
        frames.append(cz.compress(o))
    return frames, origs


def small_alphabet_frames(n_frames: int, seed: int = 3, level: int = 3):
    """Low-valued byte alphabets (3..24 symbols, values 0..23): these make libzstd emit DIRECT
    (4-bit) Huffman weight headers, the case where huff0_decoder.cairo:302 and RFC 8878 differ
    (SURVEY.md section 0, Appendix C)."""
    rng = np.random.default_rng(seed)
    cz = Compressor(level=level, checksum=True)
    frames, origs = [], []
    for i in range(n_frames):
        nsym = int(rng.integers(3, 25))
        size = int(rng.choice([400, 2000, 20000]))
        p = rng.random(nsym) ** 2 + 0.01
        p /= p.sum()
        o = rng.choice(nsym, size=size, p=p).astype(np.uint8).tobytes()
        origs.append(o)
        frames.append(cz.compress(o))
    return frames, origs
