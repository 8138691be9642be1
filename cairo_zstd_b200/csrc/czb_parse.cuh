// czb_parse.cuh -- header parsers shared by the device scan kernels and the host-side
// handle API (all __host__ __device__, byte-wise loads, bounds checked).
#pragma once
#include <stdint.h>

#include "czb_internal.cuh"

namespace czb {

struct FrameHeader {
    uint8_t descriptor;
    uint8_t window_descriptor;
    uint32_t dict_id;
    uint64_t fcs;
    uint32_t hdr_len;
    uint8_t fcs_bytes;
};

// read_frame_header, src/frame.cairo:152-284.  Field order: magic, descriptor, window byte,
// dictionary id, frame content size (Appendix A.1).
__host__ __device__ inline int32_t parse_frame_header(const uint8_t* p, uint64_t n, FrameHeader& h) {
    if (n < 4) return CZS_MAGIC_NUMBER_READ_ERROR;
    uint32_t magic = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    uint32_t i = 4;
    if (magic >= 0x184D2A50u && magic <= 0x184D2A5Fu) {  // :160-166
        if (n < 8) return CZS_FRAME_DESCRIPTOR_READ_ERROR;
        return CZS_SKIP_FRAME;
    }
    if (magic != 0xFD2FB528u) return CZS_BAD_MAGIC_NUMBER;
    if (n < i + 1) return CZS_FRAME_DESCRIPTOR_READ_ERROR;
    uint8_t d = p[i++];
    h.descriptor = d; h.window_descriptor = 0; h.dict_id = 0; h.fcs = 0;
    const bool single = (d >> 5) & 1;
    if (!single) {
        if (n < i + 1) return CZS_WINDOW_DESCRIPTOR_READ_ERROR;
        h.window_descriptor = p[i++];
    }
    const uint32_t dflag = d & 3;
    const uint32_t dl = dflag == 3 ? 4 : dflag;  // frame.cairo:76-90
    if (dl) {
        if (n < i + dl) return CZS_DICTIONARY_ID_READ_ERROR;
        uint32_t id = 0;
        for (uint32_t k = 0; k < dl; k++) id |= (uint32_t)p[i + k] << (8 * k);
        i += dl;
        h.dict_id = id;  // 0 is treated as absent :228-230
    }
    const uint32_t flag = d >> 6;
    const uint32_t fl = flag == 0 ? (single ? 1u : 0u) : (flag == 1 ? 2u : (flag == 2 ? 4u : 8u));  // :56-74
    h.fcs_bytes = (uint8_t)fl;
    if (fl) {
        if (n < i + fl) return CZS_DICTIONARY_ID_READ_ERROR;  // sic, :245-270
        uint64_t v = 0;
        for (uint32_t k = 0; k < fl; k++) v |= (uint64_t)p[i + k] << (8 * k);
        i += fl;
        if (fl == 2) v += 256;  // :273-275
        h.fcs = v;
    }
    h.hdr_len = i;
    return CZS_OK;
}

// FrameHeader::window_size, src/frame.cairo:106-129, then the usize conversion in
// FrameDecoderStateTrait::new (frame_decoder.cairo:71) which traps above u32.
__host__ __device__ inline int32_t frame_window_size(const FrameHeader& h, bool apply_reset_limit, uint64_t& ws) {
    if ((h.descriptor >> 5) & 1) {
        ws = h.fcs;
    } else {
        uint64_t exp = h.window_descriptor >> 3, mant = h.window_descriptor & 7;
        uint64_t base = 1ull << (10 + exp);
        uint64_t w = base + (base / 8) * mant;
        if (w < 1024) return CZS_WINDOW_TOO_SMALL;
        if (w >= 4123168604160ull) return CZS_WINDOW_TOO_BIG;
        ws = w;
    }
    if (apply_reset_limit && ws > 100ull * 1024 * 1024) return CZS_WINDOW_SIZE_TOO_BIG;  // frame_decoder.cairo:92-94
    if (ws > 0xFFFFFFFFull) return CZS_PANIC_INTERNAL;
    return CZS_OK;
}

struct ParsedBlock {
    int32_t hdr_status;     // != OK: header unreadable (pseudo block)
    uint8_t type, last;
    uint32_t size;          // Block_Size
    uint32_t content;       // bytes of content in the source
    // compressed blocks
    int32_t pre_status, seqhdr_status;
    uint8_t lit_type, n_streams, modes;
    uint32_t regen, lit_hdr, lit_payload;
    uint32_t n_seq, seq_hdr;
};

// read_block_header (block_decoder.cairo:237-278, :284-321) at src[pos..], followed by the
// section headers of a Compressed block: LiteralsSection::parse_from_header
// (literals_section.cairo:81-175), the size split of decompress_block (block_decoder.cairo:139-214)
// and SequencesHeader::parse_from_header (sequence_section.cairo:77-114).
__host__ __device__ inline void parse_block_at(const uint8_t* src, uint64_t len, uint64_t pos, ParsedBlock& b) {
    b.hdr_status = CZS_OK; b.pre_status = CZS_OK; b.seqhdr_status = CZS_OK;
    b.lit_type = 0; b.n_streams = 0; b.modes = 0; b.regen = 0; b.lit_hdr = 0; b.lit_payload = 0; b.n_seq = 0; b.seq_hdr = 0;
    b.type = BT_ERROR; b.last = 0; b.size = 0; b.content = 0;
    if (len - pos < 3) { b.hdr_status = CZS_PANIC_TRUNCATED; return; }  // r.slice(0,3) :240
    const uint8_t a = src[pos], b1 = src[pos + 1], c = src[pos + 2];
    const uint32_t t = (a >> 1) & 3;
    if (t == 3) { b.hdr_status = CZS_FOUND_RESERVED_BLOCK; return; }
    const uint32_t size = (a >> 3) | ((uint32_t)b1 << 5) | ((uint32_t)c << 13);
    if (size > MAX_BLOCK_SIZE) { b.hdr_status = CZS_BLOCK_SIZE_TOO_LARGE; return; }
    b.last = a & 1; b.size = size;
    b.content = (t == BT_RLE) ? 1u : size;
    if (len - pos - 3 < b.content) { b.hdr_status = CZS_PANIC_TRUNCATED; return; }  // slice/at asserts :98, :105, :146
    b.type = (uint8_t)t;
    if (t != BT_COMPRESSED) return;

    const uint8_t* raw = src + pos + 3;
    uint32_t rlen = size;
    if (rlen == 0) { b.pre_status = CZS_LIT_GET_BITS_ERROR; return; }
    const uint8_t b0 = raw[0];
    const uint32_t lt = b0 & 3, sf = (b0 >> 2) & 3;
    uint32_t need;
    if (lt <= 1) need = (sf == 0 || sf == 2) ? 1 : (sf == 1 ? 2 : 3);
    else need = (sf <= 1) ? 3 : (sf == 2 ? 4 : 5);
    if (rlen < need) { b.pre_status = CZS_LIT_NOT_ENOUGH_BYTES; return; }
    uint32_t regen, comp = 0;
    if (lt <= 1) {
        if (sf == 0 || sf == 2) regen = b0 >> 3;
        else if (sf == 1) regen = (b0 >> 4) + ((uint32_t)raw[1] << 4);
        else regen = (b0 >> 4) + ((uint32_t)raw[1] << 4) + ((uint32_t)raw[2] << 12);
    } else {
        b.n_streams = sf == 0 ? 1 : 4;
        if (sf <= 1) { regen = (b0 >> 4) + (((uint32_t)raw[1] & 0x3f) << 4); comp = (raw[1] >> 6) + ((uint32_t)raw[2] << 2); }
        else if (sf == 2) { regen = (b0 >> 4) + ((uint32_t)raw[1] << 4) + (((uint32_t)raw[2] & 3) << 12); comp = (raw[2] >> 2) + ((uint32_t)raw[3] << 6); }
        else { regen = (b0 >> 4) + ((uint32_t)raw[1] << 4) + (((uint32_t)raw[2] & 0x3f) << 12); comp = (raw[2] >> 6) + ((uint32_t)raw[3] << 2) + ((uint32_t)raw[4] << 10); }
    }
    b.lit_type = (uint8_t)lt; b.regen = regen; b.lit_hdr = need;
    const uint32_t upper = lt >= 2 ? comp : (lt == LT_RLE ? 1u : regen);
    b.lit_payload = upper;
    rlen -= need;
    if (rlen < upper) { b.pre_status = CZS_MALFORMED_SECTION_HEADER; return; }
    rlen -= upper;
    const uint8_t* s = raw + need + upper;
    // sequences header
    if (rlen == 0) { b.seqhdr_status = CZS_SEQ_HDR_NOT_ENOUGH_BYTES; return; }
    const uint8_t s0 = s[0];
    uint32_t br;
    if (s0 == 0) { b.n_seq = 0; b.seq_hdr = 1; return; }
    else if (s0 <= 127) { if (rlen < 2) { b.seqhdr_status = CZS_SEQ_HDR_NOT_ENOUGH_BYTES; return; } b.n_seq = s0; br = 1; }
    else if (s0 <= 254) { if (rlen < 3) { b.seqhdr_status = CZS_SEQ_HDR_NOT_ENOUGH_BYTES; return; } b.n_seq = (((uint32_t)s0 - 128) << 8) + s[1]; br = 2; }
    else { if (rlen < 4) { b.seqhdr_status = CZS_SEQ_HDR_NOT_ENOUGH_BYTES; return; } b.n_seq = (uint32_t)s[1] + ((uint32_t)s[2] << 8) + 0x7F00u; br = 3; }
    b.modes = s[br];
    b.seq_hdr = br + 1;
}

// End of the zstd frame that starts at src[0]: header, blocks up to the last one, checksum trailer.
__host__ __device__ inline int32_t frame_end_walk(const uint8_t* src, uint64_t len, const FrameHeader& fh, uint64_t& frame_len) {
    uint64_t pos = fh.hdr_len;
    for (;;) {
        if (len - pos < 3) return CZS_PANIC_TRUNCATED;
        const uint8_t a = src[pos];
        const uint32_t t = (a >> 1) & 3;
        if (t == 3) return CZS_FOUND_RESERVED_BLOCK;
        const uint32_t size = (a >> 3) | ((uint32_t)src[pos + 1] << 5) | ((uint32_t)src[pos + 2] << 13);
        if (size > MAX_BLOCK_SIZE) return CZS_BLOCK_SIZE_TOO_LARGE;
        const uint32_t content = t == BT_RLE ? 1u : size;
        if (len - pos - 3 < content) return CZS_PANIC_TRUNCATED;
        pos += 3 + content;
        if (a & 1) break;
    }
    if ((fh.descriptor >> 2) & 1) { if (len - pos < 4) return CZS_PANIC_TRUNCATED; pos += 4; }
    frame_len = pos;
    return CZS_OK;
}

// Walk a buffer of concatenated frames.  Skippable frames (magic 0x184D2A50..5F followed by a
// 4-byte little-endian size, frame.cairo:160-166 reports them as SkipFrame) are stepped over and counted; every zstd frame
// is appended to spans[] until `cap` is reached.  Stops at the end of the buffer (CZS_OK), when spans is full (CZS_OK,
// pos < len) or at the first frame that cannot be delimited (its status; pos = where it starts).
__host__ __device__ inline int32_t split_frames_walk(const uint8_t* buf, uint64_t len, czb_frame_span* spans, uint64_t cap,
                                                     uint64_t& n, uint64_t& skipped, uint64_t& pos) {
    n = 0; skipped = 0; pos = 0;
    while (pos < len && n < cap) {
        FrameHeader fh{};
        const int32_t st = parse_frame_header(buf + pos, len - pos, fh);
        if (st == CZS_SKIP_FRAME) {
            const uint8_t* p = buf + pos + 4;
            const uint64_t user = (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24);
            if (len - pos - 8 < user) return CZS_PANIC_TRUNCATED;
            pos += 8 + user; skipped++;
            continue;
        }
        if (st != CZS_OK) return st;
        uint64_t ws = 0, flen = 0;
        int32_t st2 = frame_window_size(fh, false, ws);
        if (st2 == CZS_OK) st2 = frame_end_walk(buf + pos, len - pos, fh, flen);
        if (st2 != CZS_OK) return st2;
        czb_frame_span sp;
        sp.offset = pos; sp.length = flen; sp.content_size = fh.fcs; sp.window_size = ws;
        sp.fcs_present = fh.fcs_bytes != 0; sp.has_checksum_flag = (fh.descriptor >> 2) & 1;
        spans[n++] = sp;
        pos += flen;
    }
    return CZS_OK;
}

}  // namespace czb
