// czb_fse_build.cuh -- FSE normalized-count parsing and warp-cooperative table construction.
//
// Reference: FSETable::read_probabilities (src/fse/fse_decoder.cairo:258-368),
// build_decoding_table (:156-256), calc_baseline_and_numbits (:377-400), next_position (:371-375).
//
// Table entry format (16 bit): symbol << 10 | next_state, with
//     num_bits  = log - floor(log2(next_state))
//     base_line = (next_state << num_bits) - (1 << log)
// where next_state = prob[symbol] + (number of earlier cells holding the same symbol).  This is
// algebraically the reference's calc_baseline_and_numbits(table_size, prob, k) (checked for every
// (log, prob, k) in tests/test_formulas.py) and halves the shared-memory footprint versus storing
// (base_line, num_bits, symbol) -- shared memory per block is what bounds how many sequence
// streams an SM can decode at once.
// Symbols are clamped to 63: every code above 35 (LL), 52 (ML) or 31 (OF) is an error downstream.
#pragma once
#include "czb_internal.cuh"

namespace czb {

constexpr int FSE_MAX_LOG = 9;       // LL/ML limit (sequence_section_decoder.cairo:397-399); OF is 8
constexpr int FSE_MAX_SYMBOLS = 256;

__device__ __forceinline__ uint16_t fse_entry(uint32_t sym, uint32_t next_state) {
    return (uint16_t)(((sym > 63u ? 63u : sym) << 10) | next_state);
}
__device__ __forceinline__ uint32_t fse_entry_sym(uint32_t e) { return e >> 10; }
__device__ __forceinline__ uint32_t fse_entry_nbits(uint32_t e, uint32_t log) { return log - (31u - (uint32_t)__clz(e & 1023u)); }
__device__ __forceinline__ uint32_t fse_entry_base(uint32_t e, uint32_t nb, uint32_t log) { return ((e & 1023u) << nb) - (1u << log); }

// Serial (one lane).  probs[] receives up to FSE_MAX_SYMBOLS entries (nullptr: only measure the description); n_probs counts all of them.
// A description whose accuracy log passes max_log (the reference's check) but exceeds FSE_MAX_LOG -- only possible
// for Huffman weights, where the reference passes 100 (huff0_decoder.cairo:176) -- is still parsed to the end so
// that every error the reference would raise is raised; if it parses cleanly CZS_UNSUPPORTED is returned with
// bytes_read set (the caller may still prefer its own "used too many bytes" error).
// probs32 (optional): receives the counts of an oversize description (accuracy log > FSE_MAX_LOG, where they do not fit 16 bits);
// the function then returns CZS_OK for it and the caller builds a table in global memory (k_huff_prep, Huffman weights only).
__device__ inline int32_t fse_read_probabilities(const uint8_t* p, int len, int max_log, int16_t* probs, int& n_probs, int& log,
                                                 int& bytes_read, int32_t* probs32 = nullptr) {
    FwdBits br{p, len, 0};
    uint32_t v;
    n_probs = 0;
    bytes_read = 0;
    if (!br.get(4, v)) return CZS_FSE_GET_BITS_ERROR;
    log = 5 + (int)v;
    if (log > max_log) return CZS_FSE_ACC_LOG_TOO_BIG;
    const bool oversize = log > FSE_MAX_LOG;
    const uint32_t sum = 1u << log;
    uint32_t counter = 0;
    while (counter < sum) {
        const uint32_t max_remaining = sum - counter + 1;
        const int bits = (int)highest_bit_set(max_remaining);
        uint32_t unchecked;
        if (!br.get(bits, unchecked)) return CZS_FSE_GET_BITS_ERROR;
        const uint32_t low_threshold = ((1u << bits) - 1u) - max_remaining;
        const uint32_t mask = (1u << (bits - 1)) - 1u;
        const uint32_t small = unchecked & mask;
        uint32_t value;
        if (small < low_threshold) { br.idx -= 1; value = small; }
        else if (unchecked > mask) value = unchecked - low_threshold;
        else value = unchecked;
        const int prob = (int)value - 1;
        if (n_probs < FSE_MAX_SYMBOLS) { if (!oversize) { if (probs) probs[n_probs] = (int16_t)prob; } else if (probs32) probs32[n_probs] = prob; }
        n_probs++;
        if (prob != 0) {
            counter += prob > 0 ? (uint32_t)prob : 1u;
        } else {
            for (;;) {
                uint32_t skip;
                if (!br.get(2, skip)) return CZS_FSE_GET_BITS_ERROR;
                for (uint32_t k = 0; k < skip; k++) {
                    if (n_probs < FSE_MAX_SYMBOLS) { if (!oversize) { if (probs) probs[n_probs] = 0; } else if (probs32) probs32[n_probs] = 0; }
                    n_probs++;
                }
                if (skip != 3) break;
            }
        }
    }
    if (counter != sum) return CZS_FSE_PROBABILITY_COUNTER_MISMATCH;
    if (n_probs > 256) return CZS_FSE_TOO_MANY_SYMBOLS;
    bytes_read = (br.idx + 7) >> 3;
    return (oversize && !probs32) ? CZS_UNSUPPORTED : CZS_OK;
}

// ---- oversize tables (accuracy log 10..20): only the Huffman-weight description can have one (the reference passes a limit of 100,
// huff0_decoder.cairo:176; RFC 8878 allows 6, no encoder exceeds it).  Same construction as fse_build_table_warp with 32-bit
// entries (symbol << 24 | next_state) and all arrays in global memory: one scratch area per context, taken by one warp at a time.
constexpr int FSE_BIG_MAX_LOG = 20;  // 5 + 15: the largest value the 4-bit field can give (fse_decoder.cairo:265-276)
struct FseBigScratch {
    int lock;
    int32_t probs[FSE_MAX_SYMBOLS];
    uint32_t table[1u << FSE_BIG_MAX_LOG];
    uint8_t rank_sym[1u << FSE_BIG_MAX_LOG];
};
__device__ __forceinline__ uint32_t fse_big_sym(uint32_t e) { return e >> 24; }
__device__ __forceinline__ uint32_t fse_big_nbits(uint32_t e, uint32_t log) { return log - (31u - (uint32_t)__clz(e & 0xFFFFFFu)); }
__device__ __forceinline__ uint32_t fse_big_base(uint32_t e, uint32_t nb, uint32_t log) { return ((e & 0xFFFFFFu) << nb) - (1u << log); }
__device__ inline void fse_build_table_big(int32_t* probs, int n, int log, uint32_t* table, uint8_t* rank_sym) {
    const unsigned lane = lane_id();
    const int size = 1 << log;
    int neg_idx = size, cum = 0;
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + (int)lane;
        const int pr = s < n ? probs[s] : 0;
        const unsigned negm = __ballot_sync(0xFFFFFFFFu, pr == -1);
        if (pr == -1) table[neg_idx - 1 - __popc(negm & lanemask_lt())] = ((uint32_t)s << 24) | 1u;
        neg_idx -= __popc(negm);
        const int pos = pr > 0 ? pr : 0;
        int incl = pos;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += t; }
        const int start = cum + incl - pos;
        for (int q = 0; q < 32; q++) {  // a count can be 2^19 here: the whole warp fills each symbol's range
            const int sq = __shfl_sync(0xFFFFFFFFu, start, q), pq = __shfl_sync(0xFFFFFFFFu, pos, q);
            for (int k = (int)lane; k < pq; k += 32) rank_sym[sq + k] = (uint8_t)(s0 + q);
        }
        cum += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (pr == -1) probs[s] = 1;
    }
    __syncwarp();
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int rank_base = 0;
    for (int j0 = 0; j0 < size; j0 += 32) {
        const int j = j0 + (int)lane;
        const int pos = (int)(((long long)j * step) & mask);
        const bool ok = pos < neg_idx;
        const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
        if (ok) table[pos] = (uint32_t)rank_sym[rank_base + __popc(okm & lanemask_lt())] << 24;
        rank_base += __popc(okm);
    }
    __syncwarp();
    for (int i0 = 0; i0 < neg_idx; i0 += 32) {
        const int i = i0 + (int)lane;
        const bool ok = i < neg_idx;
        const uint32_t s = ok ? (table[i] >> 24) : 0xFFFFu;
        const unsigned same = __match_any_sync(0xFFFFFFFFu, s);
        const int before = __popc(same & lanemask_lt());
        if (ok) table[i] = (s << 24) | (uint32_t)(probs[s] + before);
        __syncwarp();
        if (ok && before == 0) probs[s] = probs[s] + __popc(same);
        __syncwarp();
    }
}

// Warp-cooperative table build.  probs[0..n) in shared memory (modified: becomes the running
// next-state counter), rank_sym: scratch of (1<<log) bytes, table: (1<<log) entries.
__device__ inline void fse_build_table_warp(int16_t* probs, int n, int log, uint16_t* table, uint8_t* rank_sym) {
    const unsigned lane = lane_id();
    const int size = 1 << log;
    // 1. "less than one" symbols take cells from the top in symbol order (:169-189); positive
    //    symbols get their rank range in spread order.
    int neg_idx = size;
    int cum = 0;
    for (int s0 = 0; s0 < n; s0 += 32) {
        const int s = s0 + (int)lane;
        const int pr = s < n ? (int)probs[s] : 0;
        const unsigned negm = __ballot_sync(0xFFFFFFFFu, pr == -1);
        if (pr == -1) table[neg_idx - 1 - __popc(negm & lanemask_lt())] = fse_entry((uint32_t)s, 1u);
        neg_idx -= __popc(negm);
        int pos = pr > 0 ? pr : 0;
        int incl = pos;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += t; }
        const int start = cum + incl - pos;
        for (int k = 0; k < pos; k++) rank_sym[start + k] = (uint8_t)s;
        cum += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (pr == -1) probs[s] = 1;  // running next-state for completeness (never read again)
    }
    __syncwarp();
    // 2. spread (:191-226): the walk p -> (p + step) & mask visits every cell once; cells in the
    //    "less than one" region are skipped, the r-th accepted cell takes the r-th ranked symbol.
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int rank_base = 0;
    for (int j0 = 0; j0 < size; j0 += 32) {
        const int j = j0 + (int)lane;
        const int pos = (j * step) & mask;
        const bool ok = pos < neg_idx;
        const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
        if (ok) table[pos] = (uint16_t)((uint32_t)rank_sym[rank_base + __popc(okm & lanemask_lt())] << 10);
        rank_base += __popc(okm);
    }
    __syncwarp();
    // 3. next-state numbering in cell order (:231-255)
    for (int i0 = 0; i0 < neg_idx; i0 += 32) {
        const int i = i0 + (int)lane;
        const bool ok = i < neg_idx;
        const uint32_t s = ok ? (uint32_t)(table[i] >> 10) : 0xFFFFu;
        const unsigned same = __match_any_sync(0xFFFFFFFFu, s);
        const int before = __popc(same & lanemask_lt());
        if (ok) {
            const int ns = (int)probs[s] + before;
            table[i] = fse_entry(s, (uint32_t)ns);
        }
        __syncwarp();
        if (ok && before == 0) probs[s] = (int16_t)(probs[s] + __popc(same));
        __syncwarp();
    }
}

// Predefined distributions (sequence_section_decoder.cairo:417-456, :493-525, :561-617; RFC 8878 3.1.1.3.2.2.1-3).
__device__ __constant__ int8_t kLLDefault[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__device__ __constant__ int8_t kOFDefault[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
__device__ __constant__ int8_t kMLDefault[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};

}  // namespace czb
