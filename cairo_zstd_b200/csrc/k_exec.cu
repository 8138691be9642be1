// k_exec.cu -- block loop and sequence execution: one warp per frame, blocks in order.
//
// Reference path: the block loop of FrameDecoder::decode_blocks (src/frame_decoder.cairo:156-222),
// decode_block_content (src/decoding/block_decoder.cairo:77-137; Raw :97-103, RLE :104-123),
// the tail of decompress_block (:216-232), execute_sequences (src/decoding/sequence_execution.cairo:12-83),
// DecodeBuffer::push / repeat (src/decoding/decode_buffer.cairo:57-133) and collect() (:224-231):
// the append-only RingBuffer (ring_buffer.cairo:6) is simply the caller's dst span.
//
// Mapping: sequences are taken 32 at a time (lane = sequence).  A warp prefix sum over literal/match lengths gives
// every sequence its literal source and output offsets.
//   * Tile path (the chunk's output span fits 1 KiB, the common case): the first 16 bytes of every literal run and of
//     every match whose source precedes the chunk go into a shared-memory tile with per-lane unaligned 16-byte copies;
//     everything sparse -- tails beyond 16 bytes, matches that read the chunk's own output (in sequence order,
//     overlapping ones as a repeated pattern, decode_buffer.cairo:101-120) -- is done by the whole warp, one item at
//     a time; the tile is flushed with aligned 16-byte stores.
//   * Row path (longer spans): each of the 64 segments publishes its start and a source delta in shared memory; output
//     is produced in rows of 128 bytes, lane l owns the aligned word at row + 4l; segment ownership of every row byte
//     comes from one warp max-scan over "segment id at its start byte" marks; each byte is one gather; bytes whose
//     source lies inside the row being built, overlapping matches and RLE literals chase the source back through
//     the row's segment map.
// HBM traffic per frame: literals + 8 B/sequence in, decoded bytes out; match sources are recent output (L1/L2 when the
// in-flight working set allows, DRAM otherwise: profiles/r01_final_ncu_summary.md).
#include "czb_internal.cuh"

namespace czb {

constexpr int EXEC_WARPS = 4;
#ifndef EXEC_MIN_CTAS
#define EXEC_MIN_CTAS 7
#endif
#ifndef EXEC_CTAS_PER_SM
#define EXEC_CTAS_PER_SM 7
#endif
constexpr uint32_t EXEC_ROW = 128;
#ifndef EXEC_TILE_PATH
#define EXEC_TILE_PATH 1
#endif
constexpr uint32_t EXEC_TILE = 1024;

struct ExecWarpSmem {
    uint32_t bound[66];      // bound[2i] = first output byte of sequence i's literal run, [2i+1] = of its match, [64] = span
    int segdelta[66];        // per segment: source index = output position + delta (literal buffer for even ids, dst for odd);
                             // [64] is the "past the end" pseudo segment
    unsigned long long segbase[66];  // indexed by id = segment index + 1: address of the source byte for output position 0
    uint32_t segthr[66];     // indexed by id: a byte at row-relative... see gather: fast iff (p - rlo) < segthr[id]
#if EXEC_TILE_PATH
    __align__(16) uint8_t tile[EXEC_TILE + 48];  // a chunk's whole output span (sequence-centric path)
#endif
    __align__(4) uint8_t rowmap[EXEC_ROW];  // (segment id + 1) at each non-empty segment's start byte inside the row
    __align__(4) uint8_t krow[EXEC_ROW];    // (segment id + 1) owning each row byte
};

// dst[0..n) = src[0..n): 16-byte stores to aligned dst; src may have any alignment (aligned
// 32-bit loads + funnel shifts).  Only aligned words containing at least one source byte are read.
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n) {
    const unsigned lane = lane_id();
    if (n >= 64) {
        const uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
        if (lane < head) dst[lane] = src[lane];
        dst += head; src += head; n -= head;
        const uint32_t nv = n >> 4;
        const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
        const uint32_t sh = (uint32_t)(sa & 3) * 8;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~uintptr_t(3));
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint32_t* w = sw + 4 * i;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
            const uint32_t w4 = sh ? w[4] : 0u;
            uint4 v;
            v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
            v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
            d4[i] = v;
        }
        dst += nv << 4; src += nv << 4; n &= 15;
    }
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

__device__ __forceinline__ void warp_fill(uint8_t* dst, uint8_t byte, uint32_t n) {
    const unsigned lane = lane_id();
    if (n >= 64) {
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

#if EXEC_TILE_PATH
// 16 source bytes starting at src (any alignment, global or shared memory) as four little-endian words.  Only the
// aligned 32-bit words that hold one of the first n bytes are read.
struct Vec16 { uint32_t v[4]; };
template <bool CG = false>  // CG: read through L2 (data another warp of the CTA has just written)
__device__ __forceinline__ Vec16 load16_unaligned(const uint8_t* __restrict__ src, uint32_t n) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint32_t mis = (uint32_t)(a & 3), sh = mis * 8, need = n + mis;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
    auto ld = [&](int k) -> uint32_t { return CG ? __ldcg(w + k) : w[k]; };
    const uint32_t w0 = n ? ld(0) : 0u;
    const uint32_t w1 = need > 4 ? ld(1) : 0u, w2 = need > 8 ? ld(2) : 0u, w3 = need > 12 ? ld(3) : 0u, w4 = need > 16 ? ld(4) : 0u;
    Vec16 r;
    r.v[0] = __funnelshift_r(w0, w1, sh); r.v[1] = __funnelshift_r(w1, w2, sh);
    r.v[2] = __funnelshift_r(w2, w3, sh); r.v[3] = __funnelshift_r(w3, w4, sh);
    return r;
}
// t[0..n) = the first n bytes of x (byte stores into the shared-memory tile).  Groups of four bytes are skipped
// warp-uniformly when no lane needs them, so short segments do not pay for sixteen predicated stores.
__device__ __forceinline__ void store16_to_tile(uint8_t* t, const Vec16& x, uint32_t n) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (g == 0 || __any_sync(0xFFFFFFFFu, n > 4u * g)) {
#pragma unroll
            for (int k = 4 * g; k < 4 * g + 4; k++) if ((uint32_t)k < n) t[k] = (uint8_t)(x.v[g] >> (8 * (k & 3)));
        }
    }
}

// All lanes copy n bytes src -> tile + dst_off for the lane `j` that owns the job (arguments are taken from lane j).
template <bool CG = false>
__device__ __forceinline__ void coop_copy_to_tile(uint8_t* tile, unsigned lane, int j, uint32_t dst_off, const uint8_t* src, uint32_t n) {
    const uint32_t d = __shfl_sync(0xFFFFFFFFu, dst_off, j), cnt = __shfl_sync(0xFFFFFFFFu, n, j);
    const unsigned long long sp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)reinterpret_cast<uintptr_t>(src), j);
    const uint8_t* s = reinterpret_cast<const uint8_t*>((uintptr_t)sp);
    for (uint32_t i = lane; i < cnt; i += 32) tile[d + i] = CG ? __ldcg(s + i) : s[i];
}

// the same without the votes
__device__ __forceinline__ void store16_to_tile_all(uint8_t* t, const Vec16& x, uint32_t n) {
#pragma unroll
    for (int k = 0; k < 16; k++) if ((uint32_t)k < n) t[k] = (uint8_t)(x.v[k >> 2] >> (8 * (k & 3)));
}

// Sequence-centric execution of one chunk whose output span fits the tile: lane = sequence.  The first 16 bytes of
// every literal run and of every match whose source is already in dst are copied by their own lane (all lanes in
// parallel, uniform control flow); tails and the matches that depend on output not yet in dst are done by the whole
// warp, one at a time, in sequence order.  The tile is flushed with aligned 16-byte stores.
//
// avail_rel (<= 0) and wait_prev exist for k_exec_big, where several warps work on consecutive chunks of one frame:
// output below obase + avail_rel is complete in dst when the call starts; wait_prev() returns once everything
// below obase is.  The one-warp-per-frame kernel passes 0 and a no-op.
// Which frames get a whole CTA (k_exec_big) instead of one warp (k_exec): large ones (>= 2^big_cls compressed bytes) that
// either have sparse sequences (>= big_seq_bytes compressed bytes per sequence: literal-heavy data, long matches; the
// warps then rarely wait for each other: 1 MiB literal-heavy frames 253 -> 371 GB/s) or are a large share of the wave
// (>= 1/512 of its compressed bytes: one warp would still be on that frame long after the others have finished; with
// dense short matches the in-order commit chain limits the gain: 17 MiB long-window frames 12 -> 16 GB/s, while
// thousands of 1..4 MiB text frames are faster one warp each, 115 vs 103 GB/s).
__device__ __forceinline__ bool frame_is_big(const FrameInfo& fi, uint32_t big_cls, uint32_t big_seq_bytes, uint64_t wave_share_bytes) {
    return fi.size_cls >= big_cls && (fi.n_seq * big_seq_bytes <= fi.src_end || fi.src_end >= wave_share_bytes);
}

struct NoWait { __device__ __forceinline__ void operator()() const {} };
template <bool CG_LOADS, typename WaitPrev>
__device__ __forceinline__ void exec_chunk_tile(uint8_t* tile_base, uint8_t* obase, const uint8_t* __restrict__ lits, bool lit_rle,
                                                uint32_t rle_byte, unsigned lane, uint32_t ll, uint32_t ml, uint32_t off,
                                                uint32_t my_lit, uint32_t segA, uint32_t span, int avail_rel, WaitPrev wait_prev) {
    const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(obase) & 15);
    uint8_t* tile = tile_base + a0;  // tile[p] = output byte at chunk-relative position p
    const uint32_t segM = segA + ll;
    const bool indep = ml > 0 && (int)(segM + ml) - (int)off <= avail_rel;  // whole source is already in dst
    const uint8_t* msrc = obase + ((int64_t)segM - (int64_t)off);
    // first 16 bytes of every literal run and independent match: all loads are issued before the stores
    {
        const uint32_t nl = lit_rle ? 0u : (ll < 16u ? ll : 16u), nm = indep ? (ml < 16u ? ml : 16u) : 0u;
        const Vec16 xl = load16_unaligned(lits + my_lit, nl), xm = load16_unaligned<CG_LOADS>(msrc, nm);
        store16_to_tile(tile + segA, xl, nl);      // literal runs average under three bytes: later groups are usually skipped
        store16_to_tile_all(tile + segM, xm, nm);  // matches average nine: some lane always needs every group, votes only cost
    }
    // Tails beyond the first 16 bytes are rare (a few per cent of the segments) and may be long: the whole warp
    // copies each one instead of every lane looping in lockstep for the longest.
    for (unsigned m = __ballot_sync(0xFFFFFFFFu, !lit_rle && ll > 16u); m; m &= m - 1)
        coop_copy_to_tile(tile, lane, __ffs(m) - 1, segA + 16u, lits + my_lit + 16, ll - 16u);
    if (lit_rle) for (uint32_t k = 0; __any_sync(0xFFFFFFFFu, k < ll); k++) if (k < ll) tile[segA + k] = (uint8_t)rle_byte;
    for (unsigned m = __ballot_sync(0xFFFFFFFFu, indep && ml > 16u); m; m &= m - 1)
        coop_copy_to_tile<CG_LOADS>(tile, lane, __ffs(m) - 1, segM + 16u, msrc + 16, ml - 16u);
    __syncwarp();
    wait_prev();
    bool done = indep || ml == 0;
    if (CG_LOADS) {
        // k_exec_big: everything below obase is in dst now.  Matches whose source ends there but was not available
        // when the chunk started are mutually independent: one more per-lane pass instead of one warp pass each.
        const bool late = !done && segM + ml <= off;
        const uint32_t nm = late ? (ml < 16u ? ml : 16u) : 0u;
        if (__any_sync(0xFFFFFFFFu, late)) {
            store16_to_tile(tile + segM, load16_unaligned<true>(msrc, nm), nm);
            for (unsigned m = __ballot_sync(0xFFFFFFFFu, late && ml > 16u); m; m &= m - 1)
                coop_copy_to_tile<true>(tile, lane, __ffs(m) - 1, segM + 16u, msrc + 16, ml - 16u);
            __syncwarp();
        }
        done = done || late;
    }
    // Matches that read this chunk's own output (a few per chunk): in sequence order, the whole warp on each one, so
    // every source byte is final when it is read.  A match that overlaps itself (offset < length,
    // decode_buffer.cairo:101-120) repeats its first `offset` source bytes, which lie before its destination.
    for (unsigned U = __ballot_sync(0xFFFFFFFFu, !done); U; U &= U - 1) {
        const int j = __ffs(U) - 1;
        const uint32_t dM = __shfl_sync(0xFFFFFFFFu, segM, j), n = __shfl_sync(0xFFFFFFFFu, ml, j), o = __shfl_sync(0xFFFFFFFFu, off, j);
        const int s0 = (int)dM - (int)o;  // chunk-relative source start, may lie before the chunk (already in dst)
        if (o >= n && s0 >= 0) {  // the usual case: source inside the tile, no self-overlap
            for (uint32_t i = lane; i < n; i += 32) tile[dM + i] = tile[(uint32_t)s0 + i];
        } else if (o >= n) {
            for (uint32_t i = lane; i < n; i += 32) { const int q = s0 + (int)i; tile[dM + i] = q < 0 ? (CG_LOADS ? __ldcg(obase + q) : obase[q]) : tile[q]; }
        } else {
            for (uint32_t i = lane; i < n; i += 32) { const int q = s0 + (int)(i % o); tile[dM + i] = q < 0 ? (CG_LOADS ? __ldcg(obase + q) : obase[q]) : tile[q]; }
        }
        __syncwarp();
    }
    // flush: aligned 16-byte stores (tile index and dst address agree modulo 16)
    const uint32_t head = span < ((16 - a0) & 15) ? span : ((16 - a0) & 15);
    if (lane < head) obase[lane] = tile[lane];
    const uint32_t body = span - head, nv = body >> 4, tail = body & 15;
    const uint4* t4 = reinterpret_cast<const uint4*>(tile + head);
    uint4* g4 = reinterpret_cast<uint4*>(obase + head);
    for (uint32_t v = lane; v < nv; v += 32) g4[v] = t4[v];
    if (lane < tail) obase[head + (nv << 4) + lane] = tile[head + (nv << 4) + lane];
    __syncwarp();
}
#endif

__global__ void __launch_bounds__(EXEC_WARPS * 32, EXEC_MIN_CTAS) k_exec(const czb_frame_desc* __restrict__ descs, const FrameInfo* __restrict__ infos,
                                                           uint64_t wave_share_bytes, uint32_t big_cls, uint32_t big_seq_bytes, uint32_t n_exec, WaveCounters* __restrict__ counters, const uint32_t* __restrict__ exec_order,
                                                           BlockDesc* __restrict__ blocks,
                                                           const uint8_t* __restrict__ lit_scratch, const Seq* __restrict__ seq_scratch,
                                                           czb_frame_result* __restrict__ results) {
    __shared__ ExecWarpSmem smem[EXEC_WARPS];
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    ExecWarpSmem& sm = smem[warp];
    // Persistent warps pulling frames from a queue: frames differ in size by orders of magnitude
    // (1 KiB .. tens of MiB), so a static frame -> warp map would leave most warps idle at the tail.
    for (;;) {
    unsigned int fq = 0;
    if (lane == 0) fq = atomicAdd(&counters->exec_next, 1u);
    const uint64_t qpos = __shfl_sync(0xFFFFFFFFu, fq, 0);
    if (qpos >= n_exec) break;
    const uint64_t f = exec_order[qpos];  // largest frames first
    const FrameInfo fi = infos[f];
    if (fi.status != CZS_OK) continue;  // k_header_results already reported it
    if (frame_is_big(fi, big_cls, big_seq_bytes, wave_share_bytes)) continue;  // k_exec_big's
    const czb_frame_desc fd = descs[f];
    const uint8_t* src = fd.src;
    uint8_t* dst = fd.dst;
    const uint64_t cap = fd.dst_cap < MAX_FRAME_OUT ? fd.dst_cap : MAX_FRAME_OUT;
    const int32_t cap_status = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED;

    uint64_t out = 0;  // bytes appended so far == DecodeBuffer.len() == total_output_counter
    uint32_t h0 = 1, h1 = 4, h2 = 8;  // scratch.cairo:35
    int32_t status = CZS_OK;
    uint32_t n_done = 0;
    uint64_t bytes_read = fi.hdr_len;
    bool finished = false;

    for (uint32_t k = 0; k < fi.n_blocks && status == CZS_OK; k++) {
        const BlockDesc d = blocks[fi.block_base + k];
        const uint64_t out_before = out;
        if (d.type == BT_ERROR) { status = d.pre_status; break; }
        if (d.type == BT_RAW) {
            if (out + d.size > cap) { status = cap_status; break; }
            warp_copy(dst + out, src + d.src_off, d.size);
            out += d.size; bytes_read += 3ull + d.size;
        } else if (d.type == BT_RLE) {
            if (out + d.size > cap) { status = cap_status; break; }
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
            // records are read exactly once: stream them (evict-first) so that L2 keeps the frame's recent output instead
            Seq rec_next = lane < d.n_seq ? __ldcs(seqs + lane) : 0ull;
            for (uint32_t s0 = 0; s0 < d.n_seq; s0 += 32) {
                const uint32_t i = s0 + lane;
                const bool have = i < d.n_seq;
                const Seq rec = rec_next;
                if (i + 32 < d.n_seq) rec_next = __ldcs(seqs + i + 32);  // next chunk's record is in flight while this chunk executes
                uint32_t ll = 0, ml = 0, off = 1;
                if (have) { ll = seq_ll(rec); ml = seq_ml(rec); off = off29_resolve(seq_off29(rec), h0, h1, h2); }
                // warp prefix sums: literal offsets and output offsets (u32 cannot wrap: 32 * (131071 + 131074) < 2^32)
                uint32_t lsum = ll, osum = ll + ml;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, lsum, o), b = __shfl_up_sync(0xFFFFFFFFu, osum, o);
                    if ((int)lane >= o) { lsum += a; osum += b; }
                }
                const uint32_t my_lit = lit_pos + lsum - ll;      // literals_copy_counter before this sequence
                const uint32_t my_out = osum - ll - ml;           // output offset of the literal run, relative to chunk base
                const uint32_t span = __shfl_sync(0xFFFFFFFFu, osum, 31);
                const uint32_t lit_used = __shfl_sync(0xFFFFFFFFu, lsum, 31);
                // Errors are rare: one cheap test that is exact for "some lane fails" (totals for the literal and
                // capacity checks, one unsigned compare per lane for offset 0 / offset beyond the output so far;
                // out <= cap < 2^28, so 32 bits suffice), then the reference's checks in its order only if it fires.
                const uint32_t before_match32 = (uint32_t)out + my_out + ll;
                const bool suspicious = (uint64_t)lit_pos + lit_used > n_lit || out + span > cap || (have && off - 1u >= before_match32);
                if (__any_sync(0xFFFFFFFFu, suspicious)) {
                    const uint64_t before_match = out + my_out + ll;  // DecodeBuffer.len() when repeat() is called
                    int32_t err = CZS_OK;
                    if (have) {
                        if (ll > 0 && (uint64_t)my_lit + ll > n_lit) err = CZS_EXEC_NOT_ENOUGH_BYTES_FOR_SEQUENCE;  // :29-37
                        else if (off == 0) err = CZS_EXEC_ZERO_OFFSET;                                            // :47-49
                        else if (ml > 0 && off > before_match)                                                      // decode_buffer.cairo:65-93
                            err = (before_match <= fi.window) ? CZS_NOT_ENOUGH_BYTES_IN_DICTIONARY : CZS_OFFSET_TOO_BIG;
                        else if (before_match + ml > cap) err = cap_status;
                    }
                    const unsigned errm = __ballot_sync(0xFFFFFFFFu, err != CZS_OK);
                    if (errm) { status = __shfl_sync(0xFFFFFFFFu, err, __ffs(errm) - 1); break; }
                }
#if EXEC_TILE_PATH
                if (span <= EXEC_TILE) {  // the common case: short segments, span of a few hundred bytes
                    exec_chunk_tile<false>(sm.tile, dst + out, lits, lit_rle, rle_byte, lane, ll, ml, off, my_lit, my_out, span, 0, NoWait{});
                    out += span; lit_pos += lit_used;
                    continue;
                }
#endif
                // general path (long segments): publish the 64 segments of this chunk
                const uint32_t segA = my_out, segM = my_out + ll;
                sm.bound[2 * lane] = segA; sm.bound[2 * lane + 1] = segM;
                sm.segdelta[2 * lane] = (int)my_lit - (int)segA;
                sm.segdelta[2 * lane + 1] = -(int)off;
                uint8_t* obase = dst + out;
                // gather tables, indexed by id = segment index + 1 (0 = before the chunk, 65 = past the end):
                // source address of output position p is segbase[id] + p; the byte may be fetched directly iff
                // (p - rlo) < segthr[id]: always for literals, only while the source precedes the row for matches,
                // never for overlapping matches and RLE literals (those take the per-byte path).
                sm.segbase[2 * lane + 1] = (unsigned long long)(lits + my_lit) - segA;
                sm.segbase[2 * lane + 2] = (unsigned long long)obase - off;
                sm.segthr[2 * lane + 1] = lit_rle ? 0u : 0xFFFFFFFFu;
                sm.segthr[2 * lane + 2] = off < ml ? 0u : off;
                if (lane == 0) {
                    sm.bound[64] = span;
                    sm.segbase[0] = sm.segbase[65] = (unsigned long long)obase; sm.segthr[0] = sm.segthr[65] = 0u;
                }
                const uint32_t wrapmask = __ballot_sync(0xFFFFFFFFu, have && off < ml);  // overlapping matches
                const int a = (int)(reinterpret_cast<uintptr_t>(obase) & 3);
                uint32_t carry = 0;
                for (int r = -a; r < (int)span; r += (int)EXEC_ROW) {
                    // ---- which segment owns each byte of the row ----
                    *reinterpret_cast<uint32_t*>(&sm.rowmap[4 * lane]) = 0u;
                    __syncwarp();
                    if (ll && (uint32_t)((int)segA - r) < EXEC_ROW) sm.rowmap[(int)segA - r] = (uint8_t)(2 * lane + 1);
                    if (ml && (uint32_t)((int)segM - r) < EXEC_ROW) sm.rowmap[(int)segM - r] = (uint8_t)(2 * lane + 2);
                    if (lane == 0 && (uint32_t)((int)span - r) < EXEC_ROW) sm.rowmap[(int)span - r] = 65;  // pseudo segment: past the end
                    __syncwarp();
                    uint32_t x = *reinterpret_cast<const uint32_t*>(&sm.rowmap[4 * lane]);
                    x = __vmaxu4(x, x << 8);
                    x = __vmaxu4(x, x << 16);  // running maximum inside the word (ids grow with position)
                    uint32_t tot = x >> 24;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, tot, o); if ((int)lane >= o) tot = max(tot, t); }
                    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, tot, 1);
                    if (lane == 0) excl = 0;
                    x = __vmaxu4(x, max(excl, carry) * 0x01010101u);
                    carry = __shfl_sync(0xFFFFFFFFu, x >> 24, 31);
                    *reinterpret_cast<uint32_t*>(&sm.krow[4 * lane]) = x;
                    __syncwarp();
                    // ---- gather the four bytes of this lane's word ----
                    const int p0 = r + 4 * (int)lane;
                    const int rlo = r > 0 ? r : 0;  // match sources at or beyond this position are being built in this row
                    const uint32_t pr0 = (uint32_t)(p0 - rlo);
                    uint32_t word = 0, slow_mask = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t kj = (x >> (8 * j)) & 0xFFu;
                        const unsigned long long bs = sm.segbase[kj];
                        const uint32_t thr = sm.segthr[kj];
                        const uint32_t pj = (uint32_t)(p0 + j);
                        const bool valid = pj < span;
                        const bool fast = (pr0 + (uint32_t)j) < thr;
                        uint32_t b = 0;
                        if (valid && fast) b = *reinterpret_cast<const uint8_t*>(bs + pj);
                        word |= b << (8 * j);
                        if (valid && !fast) slow_mask |= 1u << j;
                    }
                    if (__any_sync(0xFFFFFFFFu, slow_mask != 0)) {
                        for (int j = 0; j < 4; j++) {
                            if (!((slow_mask >> j) & 1)) continue;
                            int p = p0 + j;
                            uint32_t kq = ((x >> (8 * j)) & 0xFFu) - 1u, byte;
                            for (;;) {
                                const int dlt = sm.segdelta[kq];
                                if (!(kq & 1u)) { byte = lit_rle ? rle_byte : lits[p + dlt]; break; }
                                int q = p + dlt;  // p - offset
                                if ((wrapmask >> (kq >> 1)) & 1u) {  // overlapping match = periodic pattern (decode_buffer.cairo:101-120)
                                    const uint32_t o = (uint32_t)(-dlt), seg0 = sm.bound[kq];
                                    q = (int)seg0 + (int)(((uint32_t)p - seg0) % o) - (int)o;
                                }
                                if (q < rlo) { byte = obase[q]; break; }  // earlier rows / chunks are already in dst
                                kq = (uint32_t)sm.krow[q - r] - 1u;        // same row, strictly earlier byte: chase
                                p = q;
                            }
                            word |= byte << (8 * j);
                        }
                    }
                    if (p0 >= 0 && (uint32_t)(p0 + 3) < span) *reinterpret_cast<uint32_t*>(obase + p0) = word;
                    else {
#pragma unroll
                        for (int j = 0; j < 4; j++) if ((uint32_t)(p0 + j) < span) obase[p0 + j] = (uint8_t)(word >> (8 * j));
                    }
                    __syncwarp();
                }
                out += span; lit_pos += lit_used;
            }
            if (status != CZS_OK) break;
            if (d.n_seq) {  // history after this block (resolved against the history it started from)
                const uint32_t n0 = sym_is(d.hist_out[0]) ? sym_resolve(d.hist_out[0], h0, h1, h2) : d.hist_out[0];
                const uint32_t n1 = sym_is(d.hist_out[1]) ? sym_resolve(d.hist_out[1], h0, h1, h2) : d.hist_out[1];
                const uint32_t n2 = sym_is(d.hist_out[2]) ? sym_resolve(d.hist_out[2], h0, h1, h2) : d.hist_out[2];
                h0 = n0; h1 = n1; h2 = n2;
            }
            // rest literals (:72-78), or all literals when there are no sequences (block_decoder.cairo:229-232)
            const uint32_t rest = n_lit - lit_pos;
            if (rest) {
                if (out + rest > cap) { status = cap_status; break; }
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
    __syncwarp();
    }  // frame loop
}


// ---------------------------------------------------------------------------------------
// k_exec_big: one CTA per large frame.  A frame's blocks and sequences are one sequential stream, so the
// one-warp-per-frame kernel leaves a 17 MiB frame to a single warp.  Here BIG_WARPS warps take consecutive
// 32-sequence chunks of a block round-robin.  A per-block pre-pass (chunk totals, one scan) gives every chunk its
// literal and output offsets up front.  Chunks commit in order through two shared counters:
//   * when a chunk starts, everything below `committed_out` is complete in dst, so matches whose source ends
//     below it are copied right away (the common case for long-window data);
//   * the rest (sources in chunks still in flight, or in this chunk) wait until all earlier chunks have
//     committed and are then done in sequence order by the whole warp, exactly as in k_exec.
// Same checks, same statuses, same results as k_exec; the first failing chunk in sequence order wins.
// ---------------------------------------------------------------------------------------
#ifndef CZB_BIG_WARPS
#define CZB_BIG_WARPS 4  // swept 4/8/16: literal-heavy 1 MiB frames 398/367/346 GB/s, 17 MiB long-window frames 15.9/16.5/15.4 GB/s
#endif
constexpr int BIG_WARPS = CZB_BIG_WARPS;
constexpr uint32_t BIG_MAX_CHUNKS = 3072;  // n_seq <= 0x7F00 + 0xFFFF (sequence_section.cairo) -> at most 3065 chunks of 32

struct BigSmem {
    uint32_t chunk_lit[BIG_MAX_CHUNKS + 1];  // exclusive prefix of literal bytes per chunk, [n] = block total
    uint32_t chunk_out[BIG_MAX_CHUNKS + 1];  // exclusive prefix of output bytes per chunk
    __align__(16) uint8_t tile[BIG_WARPS][EXEC_TILE + 48];
    unsigned long long committed_out;        // every output byte below this offset is in dst
    uint32_t committed_chunks;               // chunks (numbered through the whole frame) committed so far
    uint32_t err_chunk;                      // smallest failing chunk number, NONE32 if none
    int32_t err_status;
};

__device__ __forceinline__ void cta_copy(uint8_t* dst, const uint8_t* src, uint64_t n, unsigned warp) {
    const uint64_t per = ((n + BIG_WARPS - 1) / BIG_WARPS + 15) & ~15ull;
    const uint64_t lo = (uint64_t)warp * per;
    if (lo < n) warp_copy(dst + lo, src + lo, (uint32_t)(n - lo < per ? n - lo : per));
}
__device__ __forceinline__ void cta_fill(uint8_t* dst, uint8_t byte, uint64_t n, unsigned warp) {
    const uint64_t per = ((n + BIG_WARPS - 1) / BIG_WARPS + 15) & ~15ull;
    const uint64_t lo = (uint64_t)warp * per;
    if (lo < n) warp_fill(dst + lo, byte, (uint32_t)(n - lo < per ? n - lo : per));
}

__global__ void __launch_bounds__(BIG_WARPS * 32) k_exec_big(const czb_frame_desc* __restrict__ descs, const FrameInfo* __restrict__ infos,
                                                           uint64_t wave_share_bytes, uint32_t big_cls, uint32_t big_seq_bytes,
                                                           const uint32_t* __restrict__ exec_order, BlockDesc* __restrict__ blocks,
                                                           const uint8_t* __restrict__ lit_scratch, const Seq* __restrict__ seq_scratch,
                                                           czb_frame_result* __restrict__ results) {
    extern __shared__ __align__(16) uint8_t big_raw[];
    BigSmem& sm = *reinterpret_cast<BigSmem*>(big_raw);
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t f = exec_order[blockIdx.x];
    const FrameInfo fi = infos[f];
    if (fi.status != CZS_OK) return;  // k_header_results already reported it
    if (!frame_is_big(fi, big_cls, big_seq_bytes, wave_share_bytes)) return;  // k_exec's
    const czb_frame_desc fd = descs[f];
    const uint8_t* src = fd.src;
    uint8_t* dst = fd.dst;
    const uint64_t cap = fd.dst_cap < MAX_FRAME_OUT ? fd.dst_cap : MAX_FRAME_OUT;
    const int32_t cap_status = fd.dst_cap < MAX_FRAME_OUT ? CZS_DST_TOO_SMALL : CZS_UNSUPPORTED;
    volatile unsigned long long* v_out = &sm.committed_out;
    volatile uint32_t* v_chunks = &sm.committed_chunks;
    if (threadIdx.x == 0) { sm.committed_out = 0; sm.committed_chunks = 0; sm.err_chunk = NONE32; sm.err_status = CZS_OK; }
    __syncthreads();

    // every thread tracks the frame-level state identically (all of it is uniform)
    uint64_t out = 0;
    uint32_t h0 = 1, h1 = 4, h2 = 8;  // scratch.cairo:35
    int32_t status = CZS_OK;
    uint32_t n_done = 0, gc_base = 0;
    uint64_t bytes_read = fi.hdr_len;
    bool finished = false;

    for (uint32_t k = 0; k < fi.n_blocks && status == CZS_OK; k++) {
        const BlockDesc d = blocks[fi.block_base + k];
        const uint64_t out_before = out;
        if (d.type == BT_ERROR) { status = d.pre_status; break; }
        if (d.type == BT_RAW) {
            if (out + d.size > cap) { status = cap_status; break; }
            cta_copy(dst + out, src + d.src_off, d.size, warp);
            out += d.size; bytes_read += 3ull + d.size;
        } else if (d.type == BT_RLE) {
            if (out + d.size > cap) { status = cap_status; break; }
            cta_fill(dst + out, src[d.src_off], d.size, warp);
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
            const Seq* seqs = seq_scratch + d.seq_off;
            const uint32_t n_chunks = (d.n_seq + 31) / 32;
            // ---- pre-pass: literal and output bytes of every chunk, then one exclusive scan ----
            for (uint32_t c = warp; c < n_chunks; c += BIG_WARPS) {
                const uint32_t i = c * 32 + lane;
                uint32_t ll = 0, ml = 0;
                if (i < d.n_seq) { const Seq rec = seqs[i]; ll = seq_ll(rec); ml = seq_ml(rec); }
                uint32_t ls = ll, os = ll + ml;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { ls += __shfl_xor_sync(0xFFFFFFFFu, ls, o); os += __shfl_xor_sync(0xFFFFFFFFu, os, o); }
                if (lane == 0) { sm.chunk_lit[c] = ls; sm.chunk_out[c] = os; }
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t cl = 0, co = 0;
                for (uint32_t b = 0; b < n_chunks; b += 32) {
                    const uint32_t c = b + lane;
                    const uint32_t vl = c < n_chunks ? sm.chunk_lit[c] : 0u, vo = c < n_chunks ? sm.chunk_out[c] : 0u;
                    uint32_t il = vl, io = vo;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, il, o), b2 = __shfl_up_sync(0xFFFFFFFFu, io, o);
                        if ((int)lane >= o) { il += a; io += b2; }
                    }
                    if (c < n_chunks) { sm.chunk_lit[c] = cl + il - vl; sm.chunk_out[c] = co + io - vo; }
                    cl += __shfl_sync(0xFFFFFFFFu, il, 31); co += __shfl_sync(0xFFFFFFFFu, io, 31);
                }
                if (lane == 0) { sm.chunk_lit[n_chunks] = cl; sm.chunk_out[n_chunks] = co; }
            }
            __syncthreads();
            const uint32_t lit_total = sm.chunk_lit[n_chunks], out_total = sm.chunk_out[n_chunks];
            // ---- chunks, round-robin over the warps, committed in order ----
            for (uint32_t c = warp; c < n_chunks; c += BIG_WARPS) {
                const uint32_t gc = gc_base + c;
                const uint32_t i = c * 32 + lane;
                const bool have = i < d.n_seq;
                uint32_t ll = 0, ml = 0, off = 1;
                if (have) { const Seq rec = __ldcs(seqs + i); ll = seq_ll(rec); ml = seq_ml(rec); off = off29_resolve(seq_off29(rec), h0, h1, h2); }
                uint32_t lsum = ll, osum = ll + ml;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, lsum, o), b = __shfl_up_sync(0xFFFFFFFFu, osum, o);
                    if ((int)lane >= o) { lsum += a; osum += b; }
                }
                const uint32_t lit_pos = sm.chunk_lit[c];
                const uint64_t out_c = out + sm.chunk_out[c];
                const uint32_t my_lit = lit_pos + lsum - ll, my_out = osum - ll - ml;
                const uint32_t span = __shfl_sync(0xFFFFFFFFu, osum, 31);
                // the reference's checks in its order (see k_exec); a failing chunk writes nothing but still commits
                const uint64_t before_match = out_c + my_out + ll;
                int32_t err = CZS_OK;
                if (have) {
                    if (ll > 0 && (uint64_t)my_lit + ll > n_lit) err = CZS_EXEC_NOT_ENOUGH_BYTES_FOR_SEQUENCE;  // :29-37
                    else if (off == 0) err = CZS_EXEC_ZERO_OFFSET;                                            // :47-49
                    else if (ml > 0 && off > before_match)                                                      // decode_buffer.cairo:65-93
                        err = (before_match <= fi.window) ? CZS_NOT_ENOUGH_BYTES_IN_DICTIONARY : CZS_OFFSET_TOO_BIG;
                    else if (before_match + ml > cap) err = cap_status;
                }
                const unsigned errm = __ballot_sync(0xFFFFFFFFu, err != CZS_OK);
                auto wait_prev = [&]() {
                    while (*v_chunks != gc) __nanosleep(32);
                    __threadfence_block();
                };
                if (errm) {
                    const int32_t e = __shfl_sync(0xFFFFFFFFu, err, __ffs(errm) - 1);
                    wait_prev();  // every earlier chunk has committed: if one of them failed, its number is already there
                    if (lane == 0 && gc < sm.err_chunk) { sm.err_chunk = gc; sm.err_status = e; }
                } else if (span <= EXEC_TILE) {
                    const unsigned long long avail = *v_out;  // a lower bound is fine: it only grows
                    __threadfence_block();
                    const uint64_t behind = out_c - (avail < out_c ? avail : out_c);
                    const int avail_rel = -(int)(behind < 0x40000000ull ? behind : 0x40000000ull);
                    exec_chunk_tile<true>(sm.tile[warp], dst + out_c, lits, lit_rle, rle_byte, lane, ll, ml, off, my_lit, my_out, span, avail_rel, wait_prev);
                } else {
                    // long segments: once everything before the chunk is in dst, one sequence at a time, the whole warp
                    // copying straight into dst (overlapping matches as a repeated pattern, decode_buffer.cairo:101-120)
                    wait_prev();
                    for (int j = 0; j < 32; j++) {
                        const uint32_t jl = __shfl_sync(0xFFFFFFFFu, ll, j), jm = __shfl_sync(0xFFFFFFFFu, ml, j), jo = __shfl_sync(0xFFFFFFFFu, off, j);
                        const uint32_t jlit = __shfl_sync(0xFFFFFFFFu, my_lit, j), jout = __shfl_sync(0xFFFFFFFFu, my_out, j);
                        uint8_t* o = dst + out_c + jout;
                        if (jl) { if (lit_rle) warp_fill(o, (uint8_t)rle_byte, jl); else warp_copy(o, lits + jlit, jl); }
                        o += jl;
                        if (jm) {
                            __syncwarp();
                            if (jo >= jm) warp_copy(o, o - jo, jm);
                            else for (uint32_t t = lane; t < jm; t += 32) o[t] = __ldcg(o - jo + (t % jo));
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) { *v_out = out_c + span; __threadfence_block(); *v_chunks = gc + 1; }
            }
            __syncthreads();
            if (sm.err_chunk != NONE32) { status = sm.err_status; break; }
            gc_base += n_chunks;
            out += out_total;
            if (d.n_seq) {  // history after this block (resolved against the history it started from)
                const uint32_t n0 = sym_is(d.hist_out[0]) ? sym_resolve(d.hist_out[0], h0, h1, h2) : d.hist_out[0];
                const uint32_t n1 = sym_is(d.hist_out[1]) ? sym_resolve(d.hist_out[1], h0, h1, h2) : d.hist_out[1];
                const uint32_t n2 = sym_is(d.hist_out[2]) ? sym_resolve(d.hist_out[2], h0, h1, h2) : d.hist_out[2];
                h0 = n0; h1 = n1; h2 = n2;
            }
            // rest literals (:72-78), or all literals when there are no sequences (block_decoder.cairo:229-232)
            const uint32_t rest = n_lit - lit_total;
            if (rest) {
                if (out + rest > cap) { status = cap_status; break; }
                if (lit_rle) cta_fill(dst + out, (uint8_t)rle_byte, rest, warp);
                else cta_copy(dst + out, lits + lit_total, rest, warp);
                out += rest;
            }
            bytes_read += 3ull + d.size;
        }
        n_done++;
        if (threadIdx.x == 0) blocks[fi.block_base + k].out_bytes = (uint32_t)(out - out_before);
        if (d.last) {
            finished = true;
            if ((fi.descriptor >> 2) & 1) bytes_read += 4;
        }
        __syncthreads();  // the block's output (incl. raw/rle/rest copies by all warps) is complete in dst
        if (threadIdx.x == 0) { *v_out = out; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
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

static int exec_persistent_ctas() {
    static int n = 0;
    if (!n) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        n = sms * EXEC_CTAS_PER_SM;
    }
    return n;
}

void launch_exec(const LaunchCtx& lc, const czb_frame_desc* descs, const FrameInfo* infos, uint64_t first, uint64_t count, uint32_t n_big_cls,
                 uint32_t n_exec, uint32_t big_cls, uint32_t big_seq_bytes, uint64_t wave_src_bytes, WaveCounters* counters,
                 const uint32_t* exec_order, BlockDesc* blocks, const uint8_t* lit_scratch, const Seq* seq_scratch,
                 czb_frame_result* results) {
    if (!count || !n_exec) return;
    const uint64_t wave_share_bytes = big_seq_bytes ? wave_src_bytes / 512 + 1 : 0;  // big_seq_bytes == 0 (test knob): every large frame
    // exec_order lists the wave's frames largest size class first.  k_exec_big gets one CTA for each of the first
    // n_big_cls entries (the frames of at least 2^big_cls bytes) and takes those that pass frame_is_big; k_exec walks
    // the whole list and skips exactly those.
    const uint64_t want = ((uint64_t)n_exec + EXEC_WARPS - 1) / EXEC_WARPS;
    const unsigned grid = (unsigned)(want < (uint64_t)exec_persistent_ctas() ? want : (uint64_t)exec_persistent_ctas());
    k_exec<<<grid, EXEC_WARPS * 32, 0, lc.stream>>>(descs + first, infos + first, wave_share_bytes, big_cls, big_seq_bytes, n_exec, counters, exec_order, blocks, lit_scratch, seq_scratch, results + first);
    ++*lc.launches;
    if (n_big_cls) {
        k_exec_big<<<n_big_cls, BIG_WARPS * 32, sizeof(BigSmem), lc.stream>>>(descs + first, infos + first, wave_share_bytes, big_cls, big_seq_bytes, exec_order, blocks, lit_scratch, seq_scratch, results + first);
        ++*lc.launches;
    }
}

int setup_exec_attributes() {
    return (int)cudaFuncSetAttribute(k_exec_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem));
}

}  // namespace czb
