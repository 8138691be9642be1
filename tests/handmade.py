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
