// frame_decoder.hpp -- header-only C++ mirror of the reference's FrameDecoder traits over the C ABI.
//
// The reference is compiled Cairo (src/frame_decoder.cairo); its toolchain is not available in this
// image, so the host side above the C ABI is C++.  Names, argument meaning and error behaviour follow
// the Cairo: Result::Err becomes a thrown czb::FrameDecoderError carrying the czs_status leaf code,
// Option becomes std::optional, `ref source: @ByteArraySlice` becomes a ByteSlice the callee advances.
//
//   czb::ByteSlice source(data, len);
//   auto state = czb::FrameDecoderState::make(ctx, source);        // FrameDecoderStateTrait::new  :54-76
//   czb::FrameDecoder dec(std::move(state));                       // FrameDecoderTrait::new       :109-111
//   dec.decode_blocks(source, czb::BlockDecodingStrategy::All());  //                              :156-222
//   assert(dec.is_finished());                                     //                              :144-150
//   std::optional<std::vector<uint8_t>> out = dec.collect();       //                              :224-231
//   assert(dec.get_checksum_from_data() == dec.get_calculated_checksum());
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cairo_zstd_b200.h"

namespace czb {

struct FrameDecoderError : std::runtime_error {  // FrameDecoderError, src/frame_decoder.cairo:39-48
    int status;
    explicit FrameDecoderError(int s) : std::runtime_error(czs_status_name(s)), status(s) {}
};

struct ByteSlice {  // ByteArraySlice, src/utils/byte_array.cairo:9-13
    const uint8_t* data;
    uint64_t len;
    ByteSlice(const uint8_t* p, uint64_t n) : data(p), len(n) {}
    void advance(uint64_t n) { data += n; len -= n; }
};

struct BlockDecodingStrategy {  // src/frame_decoder.cairo:33-37
    int kind;
    uint32_t n;
    static BlockDecodingStrategy All() { return {CZB_STRATEGY_ALL, 0}; }
    static BlockDecodingStrategy UptoBlocks(uint32_t n) { return {CZB_STRATEGY_UPTO_BLOCKS, n}; }
    static BlockDecodingStrategy UptoBytes(uint32_t n) { return {CZB_STRATEGY_UPTO_BYTES, n}; }
};

class FrameDecoderState {  // FrameDecoderStateTrait, :52-106
public:
    static FrameDecoderState make(czb_context* ctx, ByteSlice& source) {
        czb_frame_decoder* h = nullptr;
        uint64_t used = 0;
        int st = czb_fd_new(ctx, source.data, source.len, &used, &h);
        if (st != CZS_OK) throw FrameDecoderError(st);
        source.advance(used);
        return FrameDecoderState(h);
    }
    void reset(ByteSlice& source) {
        uint64_t used = 0;
        int st = czb_fd_reset(h_, source.data, source.len, &used);
        if (st != CZS_OK) throw FrameDecoderError(st);
        source.advance(used);
    }
    FrameDecoderState(FrameDecoderState&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    FrameDecoderState& operator=(FrameDecoderState&& o) noexcept { std::swap(h_, o.h_); return *this; }
    FrameDecoderState(const FrameDecoderState&) = delete;
    ~FrameDecoderState() { if (h_) czb_fd_free(h_); }
    czb_frame_decoder* handle() const { return h_; }
private:
    explicit FrameDecoderState(czb_frame_decoder* h) : h_(h) {}
    czb_frame_decoder* h_;
};

class FrameDecoder {  // FrameDecoderTrait, :108-335
public:
    explicit FrameDecoder(FrameDecoderState state) : state_(std::move(state)) {}
    void init(FrameDecoderState state) { reset(std::move(state)); }
    void reset(FrameDecoderState state) { state_ = std::move(state); }
    uint64_t content_size() const { return czb_fd_content_size(h()); }
    std::optional<uint32_t> get_checksum_from_data() const { uint32_t v; return czb_fd_get_checksum_from_data(h(), &v) ? std::optional<uint32_t>(v) : std::nullopt; }
    std::optional<uint32_t> get_calculated_checksum() const { uint32_t v; return czb_fd_get_calculated_checksum(h(), &v) ? std::optional<uint32_t>(v) : std::nullopt; }
    uint64_t bytes_read_from_source() const { return czb_fd_bytes_read_from_source(h()); }
    bool is_finished() const { return czb_fd_is_finished(h()) != 0; }
    uint32_t blocks_decoded() const { return czb_fd_blocks_decoded(h()); }
    bool decode_blocks(ByteSlice& source, BlockDecodingStrategy strat) {
        uint64_t used = 0;
        int32_t fin = 0;
        int st = czb_fd_decode_blocks(h(), source.data, source.len, &used, strat.kind, strat.n, &fin);
        source.advance(used);
        if (st != CZS_OK) throw FrameDecoderError(st);
        return fin != 0;
    }
    uint64_t can_collect() const { return czb_fd_can_collect(h()); }
    std::optional<std::vector<uint8_t>> collect() {
        std::vector<uint8_t> buf(can_collect());
        uint64_t wrote = 0;
        int rc = czb_fd_collect(h(), buf.data(), buf.size(), &wrote);
        if (rc < 0) throw FrameDecoderError(-rc);
        if (rc == 0) return std::nullopt;
        buf.resize(wrote);
        return buf;
    }
    std::pair<uint64_t, uint64_t> decode_from_to(ByteSlice source, std::vector<uint8_t>& target, uint64_t cap = 1ull << 26) {
        std::vector<uint8_t> buf(cap);
        uint64_t rl = 0, wr = 0;
        int st = czb_fd_decode_from_to(h(), source.data, source.len, buf.data(), cap, &rl, &wr);
        if (st != CZS_OK) throw FrameDecoderError(st);
        target.insert(target.end(), buf.begin(), buf.begin() + wr);
        return {rl, wr};
    }
    uint64_t read(std::vector<uint8_t>& target, uint64_t cap = 1ull << 26) {
        std::vector<uint8_t> buf(cap);
        int64_t n = czb_fd_read(h(), buf.data(), cap);
        if (n < 0) throw FrameDecoderError((int)-n);
        target.insert(target.end(), buf.begin(), buf.begin() + n);
        return (uint64_t)n;
    }
private:
    czb_frame_decoder* h() const { return state_.handle(); }
    FrameDecoderState state_;
};

}  // namespace czb
