"""Hand-assembled zstd frames for corners no encoder produces (RFC 8878 field layouts; the reference's
parsers: src/frame.cairo:152-284, src/decoding/block_decoder.cairo:237-278,
src/decoding/literals_section_decoder.cairo:41-120 (header), :203-240 (streams), huff0_decoder.cairo:288-318)."""
import struct

MAGIC = struct.pack("<I", 0xFD2FB528)


def rev_stream(bits):
    """Reverse bitstream that yields `bits` (first decoded first): last bit lowest, end marker on top."""
    v = 1
    for b in bits:
        v = (v << 1) | (b & 1)
    return v.to_bytes((v.bit_length() + 7) // 8, "little")


def huf4_two_symbol_frame(streams):
    """Single-segment frame, one compressed block, no sequences; literals are a 4-stream Huffman section over
    the alphabet {0x00, 0x01} with direct weights (1, implicit 1) so that every symbol is one bit.
    `streams` = four lists of 0/1; their lengths need not follow the RFC's (regen+3)/4 split."""
    assert len(streams) == 4
    regen = sum(len(s) for s in streams)
    enc = [rev_stream(s) for s in streams]
    tree = bytes([128, 0x10])  # one explicit weight: symbol 0 -> 1 (high nibble first)
    jump = struct.pack("<HHH", len(enc[0]), len(enc[1]), len(enc[2]))
    payload = tree + jump + b"".join(enc)
    comp = len(payload)
    lit_hdr = (2 | (1 << 2) | (regen << 4) | (comp << 14)).to_bytes(3, "little")
    block = lit_hdr + payload + b"\x00"  # 0 sequences
    bh = (1 | (2 << 1) | (len(block) << 3)).to_bytes(3, "little")
    fhd = bytes([0x20, regen])  # single segment, 1-byte frame content size, no checksum
    expected = bytes(b for s in streams for b in s)
    return MAGIC + fhd + bh + block, expected


class BitWriter:
    """LSB-first bit writer (the forward reader's order, bit_reader.cairo:38-44)."""

    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, bits):
        self.v |= (value & ((1 << bits) - 1)) << self.n
        self.n += bits

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def fse_normalized_counts(log, probs):
    """RFC 8878 4.1.1 table description for `probs` (count per symbol; -1 = "less than one"; they must sum to 1 << log, counting
    -1 as 1): the inverse of FSETable::read_probabilities (src/fse/fse_decoder.cairo:258-368)."""
    w = BitWriter()
    w.put(log - 5, 4)
    remaining = 1 << log
    i = 0
    while i < len(probs) and remaining > 0:
        p = probs[i]
        value = p + 1
        max_remaining = remaining + 1
        bits = max_remaining.bit_length()
        low_threshold = ((1 << bits) - 1) - max_remaining
        mask = (1 << (bits - 1)) - 1
        if value < low_threshold:
            w.put(value, bits - 1)
        elif value <= mask:
            w.put(value, bits)
        else:
            w.put(value + low_threshold, bits)
        remaining -= p if p > 0 else (1 if p == -1 else 0)
        i += 1
        if p == 0:  # repeat flags: how many more zero-probability symbols follow, 2 bits at a time, 3 = "and more"
            run = 0
            while i + run < len(probs) and probs[i + run] == 0:
                run += 1
            i += run
            while run >= 3:
                w.put(3, 2)
                run -= 3
            w.put(run, 2)
    assert remaining == 0, remaining
    return w.bytes()


def frame_with_weight_table(log, probs, weight_stream: bytes, huf_stream: bytes, regen: int):
    """Single-segment frame, one compressed block, no sequences; literals are a 1-stream Huffman section whose tree is described
    by FSE-compressed weights with the given normalized counts (any accuracy log 5..20) and the given raw weight bitstream."""
    desc = fse_normalized_counts(log, probs)
    tree = desc + weight_stream
    assert len(tree) < 128
    payload = bytes([len(tree)]) + tree + huf_stream
    comp = len(payload)
    assert regen < 1024 and comp < 1024
    lit_hdr = (2 | (0 << 2) | (regen << 4) | (comp << 14)).to_bytes(3, "little")
    block = lit_hdr + payload + b"\x00"
    bh = (1 | (2 << 1) | (len(block) << 3)).to_bytes(3, "little")
    if regen < 256:
        fhd = bytes([0x20, regen])
    else:
        fhd = bytes([0x60]) + struct.pack("<H", regen - 256)
    return MAGIC + fhd + bh + block


def greedy_sequences_frame(n_seq, seed, history_blocks=1, max_of_code=24, window_log=24):
    """A frame whose sequence section consumes as many bits per sequence as valid input can: `history_blocks` raw blocks of
    128 KiB (so that large offsets are legal), then ONE compressed block with raw literals and `n_seq` sequences whose three FSE
    tables are built so that every state update reads the table's full accuracy log (9 + 9 + 8 bits): the symbols in use are
    "less than one" symbols (one cell each at the top of the table, baseline 0, num_bits = log:
    src/fse/fse_decoder.cairo:169-189, :377-400), so the encoder is free to name any next state.  Offset codes grow with the
    output (up to `max_of_code` extra bits); literal lengths use code 22 (3 extra bits), match lengths code 32 (1 extra bit).
    ~50 bits per sequence against ~25 for text: several 16-byte chunks of the backward bitstream per four sequences, which is what
    the ring look-after of k_fse has to keep up with.  Returns (frame, expected output)."""
    import random
    rng = random.Random(seed)
    out = bytearray()
    body = b""
    for _ in range(history_blocks):
        raw = bytes(rng.getrandbits(8) for _ in range(1 << 17))
        body += (0 | (0 << 1) | ((1 << 17) << 3)).to_bytes(3, "little") + raw
        out += raw
    # tables: symbol 0 takes every cell the used symbols leave
    LL_LOG, OF_LOG, ML_LOG = 9, 8, 9
    ll_probs = [511] + [0] * 21 + [-1]                      # code 22: baseline 32, 3 bits
    ml_probs = [511] + [0] * 31 + [-1]                      # code 32: baseline 35, 1 bit
    of_codes = list(range(2, max_of_code + 1))             # offset value (1 << N) + extra >= 4: never a repeat code
    of_probs = [256 - len(of_codes), 0] + [-1] * len(of_codes)
    ll_state = (1 << LL_LOG) - 1                            # the only "less than one" symbol sits in the top cell
    ml_state = (1 << ML_LOG) - 1
    of_state = {c: (1 << OF_LOG) - 1 - k for k, c in enumerate(of_codes)}  # top cells in symbol order
    bits = []
    def put(v, n):
        bits.extend((v >> (n - 1 - k)) & 1 for k in range(n))
    lits = bytearray()
    seqs = []
    produced = len(out)
    for _ in range(n_seq):
        ll = 32 + rng.randrange(8)
        ml = 35 + rng.randrange(2)
        before = produced + ll
        n = min(max_of_code, (before + 3).bit_length() - 1)
        if rng.random() < 0.1:
            n = rng.randrange(2, n + 1)
        extra = rng.randrange(0, min((1 << n) - 1, before + 3 - (1 << n)) + 1)
        seqs.append((ll, ml, n, extra))
        produced = before + ml
    put(ll_state, LL_LOG); put(of_state[seqs[0][2]], OF_LOG); put(ml_state, ML_LOG)   # initial states: LL, OF, ML
    for i, (ll, ml, n, extra) in enumerate(seqs):
        put(extra, n); put(ml - 35, 1); put(ll - 32, 3)        # extra bits: offset, match length, literal length
        if i + 1 < len(seqs):
            put(ll_state, LL_LOG); put(ml_state, ML_LOG); put(of_state[seqs[i + 1][2]], OF_LOG)  # state updates: LL, ML, OF
        new = bytes(rng.getrandbits(8) for _ in range(ll))
        lits += new
        out += new
        off = (1 << n) + extra - 3
        assert 0 < off <= len(out)
        for _ in range(ml):
            out.append(out[-off])
    assert len(lits) < (1 << 20)
    lit_hdr = (0 | (3 << 2) | (len(lits) << 4)).to_bytes(3, "little")
    assert 128 <= n_seq < 0x7F00
    seq_hdr = bytes([(n_seq >> 8) + 128, n_seq & 255, (2 << 6) | (2 << 4) | (2 << 2)])
    block = (lit_hdr + bytes(lits) + seq_hdr + fse_normalized_counts(LL_LOG, ll_probs) + fse_normalized_counts(OF_LOG, of_probs)
             + fse_normalized_counts(ML_LOG, ml_probs) + rev_stream(bits))
    body += (1 | (2 << 1) | (len(block) << 3)).to_bytes(3, "little") + block
    fhd = bytes([0x80, (window_log - 10) << 3]) + struct.pack("<I", len(out))   # 4-byte frame content size, window descriptor, no checksum
    return MAGIC + fhd + body, bytes(out)
