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
