// Compiles include/frame_decoder.hpp (the header-only C++ mirror of the reference's FrameDecoder traits) against the C
// ABI and runs the reference's own test body (_test_decode, src/tests/decoding.cairo:4-21) on the frames of a pack file:
//   u32 n, then n x { u32 frame_len, frame bytes, u32 orig_len, orig bytes }   (written by tests/test_cpp_mirror.py)
// Also walks one frame with UptoBlocks(1) + collect() and with decode_from_to (:202-214, :245-326).
// Exit code 0 = every frame decoded to its original with matching checksums.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

#include "frame_decoder.hpp"

static uint32_t rd32(const std::vector<uint8_t>& b, size_t& p) { uint32_t v; memcpy(&v, b.data() + p, 4); p += 4; return v; }

static bool test_decode(czb_context* ctx, const std::vector<uint8_t>& src, const std::vector<uint8_t>& expected) {
    czb::ByteSlice source(src.data(), src.size());
    auto state = czb::FrameDecoderState::make(ctx, source);
    czb::FrameDecoder frame_decoder(std::move(state));
    frame_decoder.decode_blocks(source, czb::BlockDecodingStrategy::All());
    if (!frame_decoder.is_finished()) { fprintf(stderr, "not finished\n"); return false; }
    auto result = frame_decoder.collect();
    if (!result) { fprintf(stderr, "collect() = None\n"); return false; }
    if (frame_decoder.get_checksum_from_data() != frame_decoder.get_calculated_checksum()) { fprintf(stderr, "checksums do not match\n"); return false; }
    if (*result != expected) { fprintf(stderr, "wrong decoding result\n"); return false; }
    return source.len == 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s pack.bin\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<uint8_t> b((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (b.size() < 4) { fprintf(stderr, "empty pack\n"); return 2; }
    czb_context* ctx = nullptr;
    if (int rc = czb_context_create(0, 0, &ctx)) { fprintf(stderr, "czb_context_create: %s\n", czs_status_name(rc)); return 3; }
    size_t p = 0;
    const uint32_t n = rd32(b, p);
    int bad = 0;
    std::vector<uint8_t> biggest_src, biggest_orig;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t fl = rd32(b, p); std::vector<uint8_t> f(b.begin() + p, b.begin() + p + fl); p += fl;
        const uint32_t ol = rd32(b, p); std::vector<uint8_t> o(b.begin() + p, b.begin() + p + ol); p += ol;
        try { if (!test_decode(ctx, f, o)) { bad++; fprintf(stderr, "frame %u failed\n", i); } }
        catch (const czb::FrameDecoderError& e) { bad++; fprintf(stderr, "frame %u: %s\n", i, e.what()); }
        if (o.size() >= biggest_orig.size()) { biggest_src = f; biggest_orig = o; }
    }
    // incremental strategies on the largest frame of the pack
    try {
        czb::ByteSlice source(biggest_src.data(), biggest_src.size());
        czb::FrameDecoder dec(czb::FrameDecoderState::make(ctx, source));
        std::vector<uint8_t> total;
        while (!dec.decode_blocks(source, czb::BlockDecodingStrategy::UptoBlocks(1))) {
            if (auto part = dec.collect()) total.insert(total.end(), part->begin(), part->end());
        }
        if (auto part = dec.collect()) total.insert(total.end(), part->begin(), part->end());
        if (total != biggest_orig || !dec.is_finished()) { bad++; fprintf(stderr, "UptoBlocks walk differs\n"); }
        czb::ByteSlice s2(biggest_src.data(), biggest_src.size());
        auto st2 = czb::FrameDecoderState::make(ctx, s2);
        czb::FrameDecoder dec2(std::move(st2));
        std::vector<uint8_t> target;
        auto rw = dec2.decode_from_to(s2, target);
        if (rw.first != s2.len || target != biggest_orig || !dec2.is_finished()) { bad++; fprintf(stderr, "decode_from_to differs\n"); }
        // error path: a truncated frame throws the reference's leaf status
        czb::ByteSlice s3(biggest_src.data(), biggest_src.size() / 2);
        czb::FrameDecoder dec3(czb::FrameDecoderState::make(ctx, s3));
        bool threw = false;
        try { dec3.decode_blocks(s3, czb::BlockDecodingStrategy::All()); } catch (const czb::FrameDecoderError& e) { threw = e.status != CZS_OK; }
        if (!threw) { bad++; fprintf(stderr, "truncated frame did not throw\n"); }
    } catch (const czb::FrameDecoderError& e) { bad++; fprintf(stderr, "incremental: %s\n", e.what()); }
    czb_context_destroy(ctx);
    printf("%u frames, %d failures\n", n, bad);
    return bad ? 1 : 0;
}
