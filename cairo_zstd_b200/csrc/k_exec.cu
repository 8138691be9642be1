// k_exec.cu -- block loop and sequence execution: one warp per frame, blocks in order.
//
// Reference path: the block loop of FrameDecoder::decode_blocks (src/frame_decoder.cairo:156-222),
// decode_block_content (src/decoding/block_decoder.cairo:77-137; Raw :97-103, RLE :104-123),
// the tail of decompress_block (:216-232), execute_sequences (src/decoding/sequence_execution.cairo:12-83),
// DecodeBuffer::push / repeat (src/decoding/decode_buffer.cairo:57-133) and collect() (:224-231):
// the append-only RingBuffer (ring_buffer.cairo:6) is simply the caller's dst span.
//
// Mapping: sequences are taken 32 at a time (lane = sequence).  A warp prefix sum over
// literal/match lengths gives every sequence its literal source and output offsets; then the
// chunk's output span is produced output-centrically: lane l computes output byte base+l, base+32+l, ...
// by locating the segment (literal run or match) that owns it.  A match byte whose source lies
// inside the same chunk is chased back through earlier segments until it reaches a literal or
// already-written output, so overlapping matches (decode_buffer.cairo:101-120) and matches on
// fresh output need no ordering between lanes, and all dst stores are coalesced.
// HBM traffic per frame: literals + 12 B/sequence in, decoded bytes out, plus match re-reads that
// mostly hit L1/L2 (recent output).
#include "czb_internal.cuh"

namespace czb {

constexpr int EXEC_WARPS = 4;

struct ExecWarpSmem {
    uint32_t bound[65];   // bound[2i] = first output byte of sequence i's literal run, [2i+1] = of its match, [64] = span
    uint32_t lit_src[32]; // literal source offset of sequence i
    uint32_t off[32];     // actual match offset of sequence i
};

__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n) {
    const unsigned lane = lane_id();
    // 16-byte path when both sides share alignment
    if (n >= 512 && ((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
        if (lane < head) dst[lane] = src[lane];
        dst += head; src += head; n -= head;
        const uint32_t nv = n >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t i = lane; i < nv; i += 32) d4[i] = s4[i];
        dst += nv << 4; src += nv << 4; n &= 15;
    }
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ void warp_fill(uint8_t* dst, uint8_t byte, uint32_t n) {
    const unsigned lane = lane_id();
    if (n >= 512) {
        const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
        if (lane < head) dst[lane] = byte;
        dst += head; n -= head;
        const uint32_t w = byte * 0x01010101u;
        const uint4 v = make_uint4(w, w, w, w);
        const uint32_t nv = n >> 4;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t i = lane; i < nv; i += 32) d4[i] = v;
        dst += nv << 4; n &= 15;
    }
    for (uint32_t i = lane; i < n; i += 32) dst[i] = byte;
}

__global__ void __launch_bounds__(EXEC_WARPS * 32) k_exec(const czb_frame_desc* __restrict__ descs, const FrameInfo* __restrict__ infos,
                                                           uint64_t count, BlockDesc* __restrict__ blocks,
                                                           const uint8_t* __restrict__ lit_scratch, const Seq* __restrict__ seq_scratch,
                                                           czb_frame_result* __restrict__ results) {
    __shared__ ExecWarpSmem smem[EXEC_WARPS];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t f = (uint64_t)blockIdx.x * EXEC_WARPS + warp;
    if (f >= count) return;
    ExecWarpSmem& sm = smem[warp];
    const FrameInfo fi = infos[f];
    if (fi.status != CZS_OK) return;  // k_header_results already reported it
    const czb_frame_desc fd = descs[f];
    const uint8_t* src = fd.src;
    uint8_t* dst = fd.dst;
    const uint64_t cap = fd.dst_cap < MAX_FRAME_OUT ? fd.dst_cap : MAX_FRAME_OUT;

    uint64_t out = 0;  // bytes appended so far == DecodeBuffer.len() == total_output_counter
    uint32_t hist[3] = {1, 4, 8};
    int32_t status = CZS_OK;
    uint32_t n_done = 0;
    uint64_t bytes_read = fi.hdr_len;
    bool finished = false;

    for (uint32_t k = 0; k < fi.n_blocks && status == CZS_OK; k++) {
        const BlockDesc d = blocks[fi.block_base + k];
        const uint64_t out_before = out;
        if (d.type == BT_ERROR) { status = d.pre_status; break; }
        if (d.type == BT_RAW) {
            if (out + d.size > cap) { status = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED; break; }
            warp_copy(dst + out, src + d.src_off, d.size);
            out += d.size; bytes_read += 3ull + d.size;
        } else if (d.type == BT_RLE) {
            if (out + d.size > cap) { status = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED; break; }
            warp_fill(dst + out, src[d.src_off], d.size);
            out += d.size; bytes_read += 4;
        } else {
            // error order of decompress_block (:139-235): literals header, literals, sequences header, sequences, execution
            if (d.pre_status != CZS_OK) { status = d.pre_status; break; }
            if (d.lit_type >= LT_COMPRESSED && d.huf_status != CZS_OK) { status = d.huf_status; break; }
            if (d.seqhdr_status != CZS_OK) { status = d.seqhdr_status; break; }
            if (d.n_seq && d.fse_status != CZS_OK) { status = d.fse_status; break; }
            const uint8_t* lits = d.lit_type >= LT_COMPRESSED ? lit_scratch + d.lit_off : src + d.lit_src_off;
            const bool lit_rle = d.lit_type == LT_RLE;
            const uint32_t rle_byte = lit_rle ? src[d.lit_src_off] : 0;
            const uint32_t n_lit = d.regen;
            uint32_t lit_pos = 0;
            __syncwarp();
            const Seq* seqs = seq_scratch + d.seq_off;
            for (uint32_t s0 = 0; s0 < d.n_seq; s0 += 32) {
                const uint32_t i = s0 + lane;
                const bool have = i < d.n_seq;
                uint32_t ll = 0, ml = 0, off = 1;
                if (have) { const Seq q = seqs[i]; ll = q.ll; ml = q.ml; off = q.off; if (sym_is(off)) off = sym_resolve(off, hist); }
                // warp prefix sums: literal offsets and output offsets
                uint32_t lsum = ll, osum = ll + ml;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, lsum, o), b = __shfl_up_sync(0xFFFFFFFFu, osum, o);
                    if ((int)lane >= o) { lsum += a; osum += b; }
                }
                // note: u32 sums cannot wrap: 32 * (131071 + 131074) < 2^32
                const uint32_t my_lit = lit_pos + lsum - ll;      // literals_copy_counter before this sequence
                const uint32_t my_out = osum - ll - ml;           // output offset of the literal run, relative to chunk base
                const uint64_t before_match = out + my_out + ll;  // DecodeBuffer.len() when repeat() is called
                int32_t err = CZS_OK;
                if (have) {
                    if (ll > 0 && (uint64_t)my_lit + ll > n_lit) err = CZS_EXEC_NOT_ENOUGH_BYTES_FOR_SEQUENCE;  // :29-37
                    else if (off == 0) err = CZS_EXEC_ZERO_OFFSET;                                            // :47-49
                    else if (ml > 0 && off > before_match)                                                      // decode_buffer.cairo:65-93
                        err = (before_match <= fi.window) ? CZS_NOT_ENOUGH_BYTES_IN_DICTIONARY : CZS_OFFSET_TOO_BIG;
                    else if (before_match + ml > cap) err = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED;
                }
                const unsigned errm = __ballot_sync(0xFFFFFFFFu, err != CZS_OK);
                if (errm) { status = __shfl_sync(0xFFFFFFFFu, err, __ffs(errm) - 1); break; }
                const uint32_t span = __shfl_sync(0xFFFFFFFFu, osum, 31);
                const uint32_t lit_used = __shfl_sync(0xFFFFFFFFu, lsum, 31);
                sm.bound[2 * lane] = my_out; sm.bound[2 * lane + 1] = my_out + ll;
                sm.lit_src[lane] = my_lit; sm.off[lane] = off;
                if (lane == 0) sm.bound[64] = span;
                __syncwarp();
                uint8_t* obase = dst + out;
                for (uint32_t p0 = 0; p0 < span; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    if (p < span) {
                        uint32_t q = p;
                        uint32_t byte;
                        for (;;) {
                            // largest kk in [0,63] with bound[kk] <= q (empty segments share a bound with their successor)
                            uint32_t kk = 0;
#pragma unroll
                            for (int stp = 32; stp > 0; stp >>= 1) if (sm.bound[kk + stp] <= q) kk += stp;
                            const uint32_t sq = kk >> 1, rel = q - sm.bound[kk];
                            if ((kk & 1) == 0) { byte = lit_rle ? rle_byte : lits[sm.lit_src[sq] + rel]; break; }
                            const uint32_t o = sm.off[sq];
                            const uint32_t r = rel >= o ? rel % o : rel;  // overlapping match = periodic pattern
                            const int64_t srcpos = (int64_t)sm.bound[kk] + r - (int64_t)o;  // relative to chunk base
                            if (srcpos < 0) { byte = obase[srcpos]; break; }
                            q = (uint32_t)srcpos;  // inside this chunk: chase to an earlier segment
                        }
                        obase[p] = (uint8_t)byte;
                    }
                }
                __syncwarp();
                out += span; lit_pos += lit_used;
            }
            if (status != CZS_OK) break;
            if (d.n_seq) {  // history after this block (resolved against the history it started from)
                uint32_t nh[3];
#pragma unroll
                for (int j = 0; j < 3; j++) nh[j] = sym_is(d.hist_out[j]) ? sym_resolve(d.hist_out[j], hist) : d.hist_out[j];
                hist[0] = nh[0]; hist[1] = nh[1]; hist[2] = nh[2];
            }
            // rest literals (:72-78), or all literals when there are no sequences (block_decoder.cairo:229-232)
            const uint32_t rest = n_lit - lit_pos;
            if (rest) {
                if (out + rest > cap) { status = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED; break; }
                if (lit_rle) warp_fill(dst + out, (uint8_t)rle_byte, rest);
                else warp_copy(dst + out, lits + lit_pos, rest);
                out += rest;
            }
            bytes_read += 3ull + d.size;
        }
        n_done++;
        if (lane == 0) blocks[fi.block_base + k].out_bytes = (uint32_t)(out - out_before);
        if (d.last) {
            finished = true;
            if ((fi.descriptor >> 2) & 1) bytes_read += 4;  // the trailer was verified present by the scan (else a pseudo block follows)
        }
        __syncwarp();
    }
    // A frame whose last block is followed by a missing checksum trailer: the pseudo block reports the trap.
    if (lane == 0) {
        czb_frame_result r;
        r.status = status;
        r.blocks_decoded = n_done;
        r.bytes_read = bytes_read;
        r.bytes_written = status == CZS_OK ? out : 0;
        r.content_size = fi.fcs;
        r.window_size = fi.window;
        r.checksum_from_data = fi.checksum;
        r.checksum_calculated = 0;
        r.has_checksum = fi.has_checksum;
        r.finished = (status == CZS_OK && finished && (!((fi.descriptor >> 2) & 1) || fi.has_checksum)) ? 1 : 0;
        results[f] = r;
    }
}

void launch_exec(const LaunchCtx& lc, const czb_frame_desc* descs, const FrameInfo* infos, uint64_t first, uint64_t count,
                 BlockDesc* blocks, const uint8_t* lit_scratch, const Seq* seq_scratch, czb_frame_result* results) {
    if (!count) return;
    const unsigned grid = (unsigned)((count + EXEC_WARPS - 1) / EXEC_WARPS);
    k_exec<<<grid, EXEC_WARPS * 32, 0, lc.stream>>>(descs + first, infos + first, count, blocks, lit_scratch, seq_scratch, results + first);
    ++*lc.launches;
}

}  // namespace czb
