// k_exec_flow.cu -- sequence execution of one LARGE frame by a whole CTA, ordered by data flow instead of by commit order.
//
// Reference path: the same as k_exec.cu -- the block loop of FrameDecoder::decode_blocks (src/frame_decoder.cairo:156-222),
// decode_block_content (src/decoding/block_decoder.cairo:77-137), execute_sequences
// (src/decoding/sequence_execution.cairo:12-83), DecodeBuffer::push / repeat (src/decoding/decode_buffer.cairo:57-133).
// Same checks, same statuses, same results; only the schedule differs.
//
// Why: a frame's sequences are one sequential stream.  k_exec gives a frame one warp; k_exec_big (k_exec.cu) gives it four
// warps that take 32-sequence chunks round-robin and COMMIT THEM IN ORDER -- every chunk waits for its predecessor before it
// touches a match whose source is not complete yet, so one frame moves at one chunk per ~1850 cycles however many warps
// there are (17 MiB long-window frames: 0.3 GB/s per frame).  But a match does not need "everything before it": it needs
// its own source bytes.  Here every output byte of the frame's recent past has a READY bit in shared memory; a chunk's
// literal runs and matches are written into a CTA-wide window as soon as their own sources are ready, in rounds (multi-
// round resolution of back-references), and the warps never wait for each other's chunks as such.  Retirement (the copy of
// a finished chunk to dst, the advance of the "everything below is in dst" mark) stays in order, off every critical path;
// it exists so that window slots can be reused and older sources can be read from dst.
//
//   window      FLOW_WIN bytes, indexed by the low bits of the frame position.
//   ready       one bit per byte for two window generations (index = position mod 2 * FLOW_WIN): a chunk that completes
//               clears the bits its slots will need one generation later, so a bit is never stale when it is looked at.
//   chunks      claimed in order from a counter.  A chunk may start once its span lies within FLOW_INFLIGHT of the retired
//               mark; sources older than FLOW_REACH below the chunk are therefore retired and are read from dst (L2).
//   long chunks (span > FLOW_SLICE, rare) run alone: after everything before them has retired, straight into dst; the window
//               restarts above them.
// Progress: the oldest unfinished chunk only depends on retired bytes and on its own earlier sequences, so it always
// completes; everybody else waits only on earlier positions.
#include <cstdio>
#include <cstdlib>

#include "czb_exec.cuh"
#include "czb_internal.cuh"

// Two shapes of the same kernel:
//   narrow  16 warps, 64 KiB window, 4 KiB slices: two CTAs per SM, for waves with up to 2 x SMs large frames;
//   wide    32 warps, 128 KiB window, 8 KiB slices: one CTA per SM, a single frame moves ~25 % faster (64 long-window frames
//           of 17 MiB: 33 -> 41 GB/s), for waves with at most one large frame per SM.
#ifndef CZB_FLOW_WARPS
#define CZB_FLOW_WARPS 16
#endif
#ifndef CZB_FLOW_MIN_CTAS
#define CZB_FLOW_MIN_CTAS 2
#endif
#ifndef CZB_FLOW_WIN_LOG
#define CZB_FLOW_WIN_LOG 16  // swept: 15 / 16 -> 27.5 / 33 GB/s on 64 long-window frames (with the slice below)
#endif
#ifndef CZB_FLOW_SLICE
#define CZB_FLOW_SLICE 4096  // 2048 / 4096 / 8192: a longer chunk runs alone (everybody waits), a larger slice leaves fewer chunks in flight
#endif

#define FLOW_NS flow_narrow
#define FLOW_P_WARPS CZB_FLOW_WARPS
#define FLOW_P_MIN_CTAS CZB_FLOW_MIN_CTAS
#define FLOW_P_WIN_LOG CZB_FLOW_WIN_LOG
#define FLOW_P_SLICE CZB_FLOW_SLICE
#include "k_exec_flow_impl.cuh"
#undef FLOW_NS
#undef FLOW_P_WARPS
#undef FLOW_P_MIN_CTAS
#undef FLOW_P_WIN_LOG
#undef FLOW_P_SLICE

#define FLOW_NS flow_wide
#define FLOW_P_WARPS 32
#define FLOW_P_MIN_CTAS 1
#define FLOW_P_WIN_LOG 17
#define FLOW_P_SLICE 8192
#include "k_exec_flow_impl.cuh"
#undef FLOW_NS
#undef FLOW_P_WARPS
#undef FLOW_P_MIN_CTAS
#undef FLOW_P_WIN_LOG
#undef FLOW_P_SLICE

namespace czb {

void launch_exec_flow(const LaunchCtx& lc, unsigned n_ctas, const czb_frame_desc* descs, const FrameInfo* infos, BigRule rule, const uint32_t* exec_order,
                      BlockDesc* blocks, const uint8_t* lit_scratch, const Seq* seq_scratch, czb_frame_result* results, const FrameResume* resume, bool wide) {
    if (wide) flow_wide::launch_exec_flow(lc, n_ctas, descs, infos, rule, exec_order, blocks, lit_scratch, seq_scratch, results, resume);
    else flow_narrow::launch_exec_flow(lc, n_ctas, descs, infos, rule, exec_order, blocks, lit_scratch, seq_scratch, results, resume);
}
int setup_exec_flow_attributes() {
    const int rc = flow_narrow::setup_exec_flow_attributes();
    return rc ? rc : flow_wide::setup_exec_flow_attributes();
}

}  // namespace czb

extern "C" int czb_debug_flow_watchdog(unsigned int* out16) {  // tests / debugging only: {code, cta, warp, 8 values}; code 0 = never fired
    if (!out16) return CZS_BAD_ARGUMENT;
    unsigned int a[16], b[16];
    if (czb::flow_narrow::read_watchdog(a) != CZS_OK || czb::flow_wide::read_watchdog(b) != CZS_OK) return CZS_CUDA_ERROR;
    for (int i = 0; i < 16; i++) out16[i] = a[0] ? a[i] : b[i];
    return CZS_OK;
}
