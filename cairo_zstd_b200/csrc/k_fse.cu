// k_fse.cu -- sequence section decode: FSE tables in shared memory, one lane per block.
//
// Reference path: decode_sequences, maybe_update_fse_tables, decode_sequences_{with,without}_rle,
// lookup_ll_code / lookup_ml_code (src/decoding/sequence_section_decoder.cairo:35-647),
// FSETable / FSEDecoder (src/fse/fse_decoder.cairo:64-400), BitReaderReversed
// (src/decoding/bit_reader_reverse.cairo), and do_offset_history
// (src/decoding/sequence_execution.cairo:85-129), which is folded in here because it is a serial
// chain over the same sequences.
//
// Mapping: the interleaved LL/OF/ML state machine is one dependency chain per block and cannot be
// split, so parallelism comes from blocks: a CTA owns 27 blocks, its 4 warps build the 81 tables
// cooperatively, then warp 0 decodes with lane = block.  Throughput is bounded by
// (blocks resident per SM) / (cycles per sequence step of a warp that is alone on its scheduler);
// 16-bit table entries (czb_fse_build.cuh) keep a block's three tables at <= 2.5 KiB so 81 blocks
// fit per SM (three CTAs).  HBM traffic: the bitstream in (a few bytes per sequence, through
// per-lane cp.async rings) and one packed 8-byte record per sequence out to scratch.
#include <type_traits>

#include "czb_fse_build.cuh"
#include "czb_internal.cuh"

namespace czb {

#ifndef CZB_FSE_WARPS
#define CZB_FSE_WARPS 4  // table-building warps per CTA (6 / 8 measured: see DESIGN.md)
#endif
constexpr int FSE_WARPS = CZB_FSE_WARPS;
#ifndef CZB_FSE_QUAD
#define CZB_FSE_QUAD 1  // four records leave as two 16-byte stores (half the store sectors); 0: one 8-byte store per sequence
#endif
#ifndef CZB_FSE_SLOTS
#define CZB_FSE_SLOTS 27
#endif
constexpr int FSE_SLOTS = CZB_FSE_SLOTS;  // 27 * 2560 B of tables + scratch = ~75 KB -> three CTAs (81 decode lanes) per SM
constexpr int FSE_SLOT_ENTRIES = 512 + 512 + 256;  // LL (log<=9), ML (log<=9), OF (log<=8)
constexpr int FSE_LL_OFS = 0, FSE_ML_OFS = 512, FSE_OF_OFS = 1024;

struct FseSlot {
    const uint8_t* bits;   // sequence bitstream
    Seq* out;
    uint32_t bits_len;
    uint32_t n_seq;
    uint32_t blk;
    int32_t status;
    uint16_t tbl[3];       // LL, OF, ML: entry index of the table inside FseSmem::entries (flat)
    int8_t log[3];         // accuracy log; 0 = RLE (one entry); -1 = never initialised
    uint8_t first_in_frame;
    uint8_t any_rle;
};

struct alignas(16) FseWarpTmp {
    int16_t probs[FSE_MAX_SYMBOLS];
    uint8_t rank_sym[1 << FSE_MAX_LOG];
};
static_assert(sizeof(FseWarpTmp) * FSE_WARPS >= FSE_SLOTS * (RevBitsWin::RING + 16), "phase-2 rings (+ one 16-byte dummy slot per lane) reuse phase-1 scratch");

struct FseSmem {
    uint16_t entries[FSE_SLOTS * FSE_SLOT_ENTRIES + 64 + 32 + 64];  // per-slot tables, then predefined LL, OF, ML
    FseSlot slot[FSE_SLOTS];
    FseWarpTmp tmp[FSE_WARPS];
    uint32_t ll_code[64];  // base | bits << 20 (lookup_ll_code :299-345); bit 31 = code beyond the table -> (0,255)
    uint32_t ml_code[64];  // (lookup_ml_code :347-395)
};
constexpr int FSE_PREDEF_LL = FSE_SLOTS * FSE_SLOT_ENTRIES, FSE_PREDEF_OF = FSE_PREDEF_LL + 64, FSE_PREDEF_ML = FSE_PREDEF_OF + 32;

__device__ __constant__ uint32_t kLLBase[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
__device__ __constant__ uint8_t kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__device__ __constant__ uint32_t kMLBase[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
__device__ __constant__ uint8_t kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

__device__ __forceinline__ int stream_max_log(int s) { return s == 1 ? 8 : 9; }  // LL 9, OF 8, ML 9 (:397-399)

// Bytes a table description of `mode` occupies at p (for skipping to a later stream's description).
// Lane 0 only.  Returns false if the description cannot be parsed.
__device__ inline bool skip_description(int mode, int s, const uint8_t* p, int len, FseWarpTmp& t, int& bytes) {
    bytes = 0;
    if (mode == MODE_RLE) { if (len < 1) return false; bytes = 1; return true; }
    if (mode == MODE_FSE) {
        int n_probs, log;
        return fse_read_probabilities(p, len, stream_max_log(s), t.probs, n_probs, log, bytes) == CZS_OK;
    }
    return true;
}

__global__ void __launch_bounds__(FSE_WARPS * 32, 3) k_fse(const czb_frame_desc* __restrict__ descs, BlockDesc* __restrict__ blocks,
                                                         const uint32_t* __restrict__ items,
                                                         const WaveCounters* __restrict__ counters, Seq* __restrict__ seq_scratch) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FseSmem& sm = *reinterpret_cast<FseSmem*>(smem_raw);
    const uint32_t n_items = counters->n_fse;
    const unsigned warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t first = blockIdx.x * FSE_SLOTS;
    if (first >= n_items) return;  // whole CTA
    FseWarpTmp& tmp = sm.tmp[warp];

    // ---- predefined tables and code tables, once per CTA ----
    if (warp < 3) {
        const int n = warp == 0 ? 36 : (warp == 1 ? 29 : 53);
        const int8_t* src = warp == 0 ? kLLDefault : (warp == 1 ? kOFDefault : kMLDefault);
        for (int i = lane; i < n; i += 32) tmp.probs[i] = src[i];
        __syncwarp();
        fse_build_table_warp(tmp.probs, n, warp == 1 ? 5 : 6,
                             sm.entries + (warp == 0 ? FSE_PREDEF_LL : (warp == 1 ? FSE_PREDEF_OF : FSE_PREDEF_ML)), tmp.rank_sym);
    } else if (warp == 3) {
        for (int i = lane; i < 64; i += 32) sm.ll_code[i] = i < 36 ? (kLLBase[i] | ((uint32_t)kLLBits[i] << 20)) : 0x80000000u;
        for (int i = lane; i < 64; i += 32) sm.ml_code[i] = i < 53 ? (kMLBase[i] | ((uint32_t)kMLBits[i] << 20)) : 0x80000000u;
    }
    __syncwarp();

    // ---- phase 1: maybe_update_fse_tables per block (:405-647), one block per warp at a time ----
    for (int s = warp; s < FSE_SLOTS; s += FSE_WARPS) {
        FseSlot& sl = sm.slot[s];
        const uint32_t item = first + s;
        if (item >= n_items) { if (lane == 0) { sl.status = CZS_NOT_DECODED; sl.blk = NONE32; sl.n_seq = 0; } continue; }
        const uint32_t bi = items[item];
        const BlockDesc d = blocks[bi];
        const uint8_t* fsrc = descs[d.frame].src;
        const uint32_t modes[3] = {(uint32_t)d.modes >> 6, ((uint32_t)d.modes >> 4) & 3u, ((uint32_t)d.modes >> 2) & 3u};
        uint32_t cursor = d.seq_src_off;
        const uint32_t end = d.seq_src_off + d.seq_src_len;
        int32_t st = CZS_OK;
        bool any_rle = false;
        const int region[3] = {s * FSE_SLOT_ENTRIES + FSE_LL_OFS, s * FSE_SLOT_ENTRIES + FSE_OF_OFS, s * FSE_SLOT_ENTRIES + FSE_ML_OFS};
        const int predef[3] = {FSE_PREDEF_LL, FSE_PREDEF_OF, FSE_PREDEF_ML};
        const int predef_log[3] = {6, 5, 6};
        for (int k = 0; k < 3 && st == CZS_OK; k++) {
            uint32_t mode = modes[k];
            const uint8_t* p = fsrc + cursor;
            int plen = (int)(end - cursor);
            const bool own = mode != MODE_REPEAT;
            bool usable = true;
            if (!own) {  // Repeat: re-derive the table from the block that last set it
                const uint32_t sb = d.tbl_src_blk[k];
                if (sb == NONE32) { usable = false; if (lane == 0) { sl.log[k] = -1; sl.tbl[k] = (uint16_t)region[k]; } }
                else {
                    const BlockDesc sd = blocks[sb];
                    const uint8_t* ssrc = fsrc;  // same frame
                    const uint32_t sm3[3] = {(uint32_t)sd.modes >> 6, ((uint32_t)sd.modes >> 4) & 3u, ((uint32_t)sd.modes >> 2) & 3u};
                    uint32_t scur = sd.seq_src_off;
                    const uint32_t send = sd.seq_src_off + sd.seq_src_len;
                    int ok = 1;
                    if (lane == 0) {
                        for (int q = 0; q < k && ok; q++) {
                            int bytes = 0;
                            ok = skip_description((int)sm3[q], q, ssrc + scur, (int)(send - scur), tmp, bytes) ? 1 : 0;
                            scur += (uint32_t)bytes;
                        }
                    }
                    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
                    scur = __shfl_sync(0xFFFFFFFFu, scur, 0);
                    if (!ok) { usable = false; if (lane == 0) { sl.log[k] = -1; sl.tbl[k] = (uint16_t)region[k]; } }  // source block failed the frame earlier
                    mode = sm3[k];
                    p = ssrc + scur;
                    plen = (int)(send - scur);
                }
            }
            if (!usable) continue;
            if (mode == MODE_PREDEFINED) {
                if (lane == 0) { sl.tbl[k] = (uint16_t)predef[k]; sl.log[k] = (int8_t)predef_log[k]; }
            } else if (mode == MODE_RLE) {
                if (plen < 1) {
                    if (own) st = k == 0 ? CZS_MISSING_BYTE_FOR_RLE_LL_TABLE : (k == 1 ? CZS_MISSING_BYTE_FOR_RLE_OF_TABLE : CZS_MISSING_BYTE_FOR_RLE_ML_TABLE);
                    else if (lane == 0) { sl.log[k] = -1; sl.tbl[k] = (uint16_t)region[k]; }
                } else {
                    if (lane == 0) { sm.entries[region[k]] = fse_entry(p[0], 1u); sl.tbl[k] = (uint16_t)region[k]; sl.log[k] = 0; }
                    any_rle = true;
                    if (own) cursor += 1;
                }
            } else {  // MODE_FSE
                int n_probs = 0, log = 0, used = 0;
                int32_t pst = CZS_OK;
                if (lane == 0) pst = fse_read_probabilities(p, plen, stream_max_log(k), tmp.probs, n_probs, log, used);
                pst = __shfl_sync(0xFFFFFFFFu, pst, 0);
                n_probs = __shfl_sync(0xFFFFFFFFu, n_probs, 0);
                log = __shfl_sync(0xFFFFFFFFu, log, 0);
                used = __shfl_sync(0xFFFFFFFFu, used, 0);
                __syncwarp();
                if (pst != CZS_OK) {
                    if (own) st = pst;
                    else if (lane == 0) { sl.log[k] = -1; sl.tbl[k] = (uint16_t)region[k]; }
                } else {
                    fse_build_table_warp(tmp.probs, n_probs, log, sm.entries + region[k], tmp.rank_sym);
                    if (lane == 0) { sl.tbl[k] = (uint16_t)region[k]; sl.log[k] = (int8_t)log; }
                    if (own) cursor += (uint32_t)used;
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            sl.blk = bi; sl.n_seq = d.n_seq; sl.status = st;
            sl.bits = fsrc + cursor; sl.bits_len = end - cursor;
            sl.out = seq_scratch + d.seq_off;
            sl.first_in_frame = d.first_in_frame; sl.any_rle = any_rle ? 1 : 0;
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp != 0) return;

    // ---- phase 2: lane = block ----
    if (lane >= FSE_SLOTS) return;
    const FseSlot& sl = sm.slot[lane];
    if (sl.blk == NONE32) return;
    int32_t st = sl.status;
    uint32_t h0, h1, h2;
    if (sl.first_in_frame) { h0 = 1; h1 = 4; h2 = 8; }  // scratch.cairo:35
    else { h0 = sym_enc(0); h1 = sym_enc(1); h2 = sym_enc(2); }
    // The decode loop exists twice.  The fast form only notes THAT something went wrong (two ORs per step); a block for
    // which it did is decoded again by the exact form, which records WHICH error came first, in the reference's order
    // (two compares and selects per step: 8 % of the kernel when it was always on).
    const uint32_t h_in0 = h0, h_in1 = h1, h_in2 = h2;
    uint32_t ml_total = 0;
    auto decode = [&](auto exact_tag) -> int32_t {
    constexpr bool EXACT = decltype(exact_tag)::value;
    int32_t st = CZS_OK;
    uint32_t trouble = 0;  // fast form: bit 31 set <=> a bad code or an over-read happened somewhere
    h0 = h_in0; h1 = h_in1; h2 = h_in2;
    ml_total = 0;
    {
        // phase 1's scratch is dead now (the other warps have left): it becomes the lanes' bitstream rings
        RevBitsWin br;
        const bool init_ok = br.init(sl.bits, (int)sl.bits_len, (uint32_t)__cvta_generic_to_shared(sm.tmp) + lane * RevBitsWin::RING);
        if (!init_ok) st = CZS_SEQ_EXTRA_PADDING;  // :46-64
        else if (sl.log[0] < 0 || sl.log[1] < 0 || sl.log[2] < 0) st = CZS_FSE_TABLE_IS_UNINITIALIZED;  // fse_decoder.cairo:82-84
        else {
            const uint32_t logLL = (uint32_t)sl.log[0], logOF = (uint32_t)sl.log[1], logML = (uint32_t)sl.log[2];
            const uint32_t mLL = (1u << logLL) - 1u, mOF = (1u << logOF) - 1u, mML = (1u << logML) - 1u;
            const uint16_t* tLL = sm.entries + sl.tbl[0];
            const uint16_t* tOF = sm.entries + sl.tbl[1];
            const uint16_t* tML = sm.entries + sl.tbl[2];
            // init order LL, OF, ML (:207-218)
            uint32_t eLL = tLL[br.get((int)logLL)];
            uint32_t eOF = tOF[br.get((int)logOF)];
            uint32_t eML = tML[br.get((int)logML)];
            const uint32_t n_seq = sl.n_seq;
            Seq* out = sl.out;
            const int32_t short_status = sl.any_rle ? CZS_SEQ_NOT_ENOUGH_BYTES_FOR_NUM_SEQUENCES : CZS_PANIC_INTERNAL;
            // One sequence (:223-286).  MORE = false is the last sequence: states are not updated (:258).
            //
            // The lane runs alone on its scheduler most of the time, so a step costs the sum of its issue stalls.
            // Hence: (a) all lookups are software-pipelined -- the next states, their num_bits, their code words
            // (lookup_ll_code / lookup_ml_code) and the next 32 stream bits are requested as soon as their inputs
            // exist and consumed one step later; (b) no data-dependent branches -- errors are recorded (first one
            // wins) and the loop runs on, which is safe because table indices are masked, records stay inside
            // this block's slice of the scratch and ring reads wrap; the ring refill is a predicated cp.async.
            uint32_t lle = sm.ll_code[fse_entry_sym(eLL)], mle = sm.ml_code[fse_entry_sym(eML)];
            RevBitsWin::Raw64 win = br.window64_raw();
            const uint32_t dummy = (uint32_t)__cvta_generic_to_shared(sm.tmp) + FSE_SLOTS * RevBitsWin::RING + lane * 16u;
            uint32_t nbLL = fse_entry_nbits(eLL, logLL), nbML = fse_entry_nbits(eML, logML), nbOF = fse_entry_nbits(eOF, logOF);
            Seq quad[4];  // four records leave as two 16-byte stores: a block's slice of the scratch is 32-byte aligned (k_fill_blocks)
            auto step = [&](uint32_t i, auto more_tag, auto refill_tag, auto quad_tag) {
                constexpr bool MORE = decltype(more_tag)::value;
                constexpr int REFILLS = decltype(refill_tag)::value;  // ring slots to look after in this step
                constexpr int QUAD = decltype(quad_tag)::value;       // position inside the group of four (-1: store it alone)
                br.step_sync();
                const uint32_t ofc = fse_entry_sym(eOF);
                // :235-237; codes beyond the tables give (0,255) -> TooManyBits
                if (EXACT) {
                    const bool bad_code = ((ofc >> 5) | ((lle | mle) >> 31)) != 0;
                    const int32_t code_status = ofc >= 32 ? CZS_SEQ_UNSUPPORTED_OFFSET : CZS_SEQ_GET_BITS_ERROR;
                    st = (st == CZS_OK && bad_code) ? code_status : st;
                } else {
                    trouble |= (ofc << 26) | lle | mle;  // ofc >= 32 puts its bit 5 at bit 31
                }
                const uint32_t llb = (lle >> 20) & 31u, mlb = (mle >> 20) & 31u, ofb = ofc & 31u;
                const uint32_t extras = ofb + mlb + llb;  // <= 63 bits, read in the order OF, ML, LL (:239)
                // The three state updates (LL, ML, OF, <= 26 bits, :258-276) come from their own 32-bit window below the
                // extra bits, whose position is known from the start of the step: no "does it all fit in 32 bits" branch.
                RevBitsWin::Raw winB;
                if (MORE) winB = br.window_raw_at(br.P - (int)extras);
                const uint32_t xh = __funnelshift_r(win.w1, win.w2, win.sh), xl = __funnelshift_r(win.w0, win.w1, win.sh);
                const uint32_t ofv = shr_clamp(xh, 32u - ofb);
                const uint32_t y = __funnelshift_l(xl, xh, ofb);
                const uint32_t mlv = shr_clamp(y, 32u - mlb), llv = shr_clamp(y << mlb, 32u - llb);
                uint32_t aLL = 0, aML = 0, aOF = 0;
                if (MORE) {
                    const uint32_t z = RevBitsWin::window_of(winB);
                    aLL = shr_clamp(z, 32u - nbLL); aML = shr_clamp(z << nbLL, 32u - nbML); aOF = shr_clamp(z << (nbLL + nbML), 32u - nbOF);
                    br.P -= (int)(extras + nbLL + nbML + nbOF);
#pragma unroll
                    for (int r = 0; r < REFILLS; r++) br.refill_nobranch(dummy);
                    win = br.window64_raw();  // the next step's bits: three LDS issued next to the state lookups below
                } else {
                    br.P -= (int)extras;
                }
                const uint32_t ll = (lle & 0xFFFFFu) + llv, ml = (mle & 0xFFFFFu) + mlv;
                if (MORE) {  // issue the next-state lookups now; they complete under the history/pack/store work below
                    eLL = tLL[(fse_entry_base(eLL, nbLL, logLL) + aLL) & mLL];
                    eML = tML[(fse_entry_base(eML, nbML, logML) + aML) & mML];
                    eOF = tOF[(fse_entry_base(eOF, nbOF, logOF) + aOF) & mOF];
                }
                const uint32_t v = (1u << ofb) + ofv;  // :243
                // do_offset_history (sequence_execution.cairo:85-129) with selects only:
                // idx 0,1,2 = history slot, 3 = h0 - 1 (reachable only when ll == 0)
                const uint32_t idx = v - (ll != 0);
                const bool rep = v <= 3;
                uint32_t cand = h0 - 1u;
                cand = idx == 2 ? h2 : cand;
                cand = idx == 1 ? h1 : cand;
                cand = idx == 0 ? h0 : cand;
                const uint32_t nz = min(v - 3u, REAL_OFF_CLAMP);
                const uint32_t act = rep ? cand : nz;
                h2 = (rep & (idx <= 1)) ? h2 : h1;
                h1 = (rep & (idx == 0)) ? h1 : h0;
                h0 = act;
                ml_total += ml;  // the block's output size is regen + sum(ml): what czb_frame_sizes_* reports without executing
                const Seq rec = (Seq)ll | ((Seq)ml << 17) | ((Seq)off29_pack(act) << 35);
                if constexpr (QUAD < 0) __stcs(out + i, rec);  // written once, read by a later kernel: streaming store
                else quad[QUAD] = rec;
                if (MORE) {
                    lle = sm.ll_code[fse_entry_sym(eLL)]; mle = sm.ml_code[fse_entry_sym(eML)];
                    nbLL = fse_entry_nbits(eLL, logLL); nbML = fse_entry_nbits(eML, logML); nbOF = fse_entry_nbits(eOF, logOF);
                }
                // :281-283; the no-RLE variant traps on the unwrap at :279 instead
                if (EXACT) st = (st == CZS_OK && br.rem() < 0) ? short_status : st;
                // (fast form: rem only ever decreases, so one look after the loop tells whether it went negative on the way)
            };
            // Four steps consume at most 356 bits, i.e. leave at most three 16-byte chunks behind: the ring is looked after
            // once per four steps (three always-issued cp.async) instead of once per step.
            using R0 = std::integral_constant<int, 0>; using R1 = std::integral_constant<int, 1>; using R3 = std::integral_constant<int, 3>;
            using Q0 = std::integral_constant<int, 0>; using Q1 = std::integral_constant<int, 1>; using Q2 = std::integral_constant<int, 2>;
            using Q3 = std::integral_constant<int, 3>; using QN = std::integral_constant<int, -1>;
            uint32_t i = 0;
            for (; i + 4 < n_seq; i += 4) {
#if CZB_FSE_QUAD
                step(i, std::true_type{}, R0{}, Q0{}); step(i + 1, std::true_type{}, R0{}, Q1{});
                step(i + 2, std::true_type{}, R0{}, Q2{}); step(i + 3, std::true_type{}, R3{}, Q3{});
                uint4* o4 = reinterpret_cast<uint4*>(out + i);
                __stcs(o4, make_uint4((uint32_t)quad[0], (uint32_t)(quad[0] >> 32), (uint32_t)quad[1], (uint32_t)(quad[1] >> 32)));
                __stcs(o4 + 1, make_uint4((uint32_t)quad[2], (uint32_t)(quad[2] >> 32), (uint32_t)quad[3], (uint32_t)(quad[3] >> 32)));
#else
                step(i, std::true_type{}, R0{}, QN{}); step(i + 1, std::true_type{}, R0{}, QN{});
                step(i + 2, std::true_type{}, R0{}, QN{}); step(i + 3, std::true_type{}, R3{}, QN{});
#endif
            }
            for (; i + 1 < n_seq; i++) step(i, std::true_type{}, R1{}, QN{});
            step(i, std::false_type{}, R0{}, QN{});
            if (!EXACT && ((trouble >> 31) || br.rem() < 0)) st = CZS_NOT_DECODED;  // placeholder: the exact form decides
            if (st == CZS_OK && br.rem() > 0) st = CZS_SEQ_EXTRA_BITS;  // :292-296
        }
    }
    return st;
    };
    if (st == CZS_OK) {
        st = decode(std::false_type{});
        if (st == CZS_NOT_DECODED) st = decode(std::true_type{});
    }
    BlockDesc& d = blocks[sl.blk];
    d.fse_status = st;
    d.hist_out[0] = h0; d.hist_out[1] = h1; d.hist_out[2] = h2;
    d.ml_sum = ml_total;
}

void launch_fse(const LaunchCtx& lc, const czb_frame_desc* descs, BlockDesc* blocks, const uint32_t* items,
                const WaveCounters* counters, uint32_t max_items, Seq* seq_scratch) {
    if (!max_items) return;
    const unsigned grid = (max_items + FSE_SLOTS - 1) / FSE_SLOTS;
    k_fse<<<grid, FSE_WARPS * 32, sizeof(FseSmem), lc.stream>>>(descs, blocks, items, counters, seq_scratch);
    ++*lc.launches;
}

int setup_fse_attributes() {
    return (int)cudaFuncSetAttribute(k_fse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FseSmem));
}

}  // namespace czb
