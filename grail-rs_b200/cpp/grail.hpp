// grail.hpp -- header-only C++ facade over include/grail_cuda.h that mirrors the reference's operator chain
//   elems.sequence(voice).jitter(seed, voice).synthesize()         (reference src/lib.rs:936-953, 781-801, 582-600)
// for hosts written in C++.  Errors become exceptions carrying grail_cuda_last_error(); no CPU path.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/grail_cuda.h"

namespace grail {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(std::string(grail_cuda_status_string(s)) + ": " + m), status(s) {}
};

class Context {
public:
    explicit Context(int device = 0)
    {
        const int rc = grail_cuda_create(device, &ctx_);
        if (rc) throw Error(rc, "grail_cuda_create");
    }
    ~Context() { grail_cuda_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    grail_ctx* get() const { return ctx_; }
    void check(int rc) const { if (rc) throw Error(rc, grail_cuda_last_error(ctx_)); }
private:
    grail_ctx* ctx_ = nullptr;
};

// the Voice scalars of reference src/lib.rs:696-717 that the path reads
struct Voice {
    float sample_rate, jitter_frequency, jitter_delta_frequency, jitter_delta_formant_frequency, jitter_delta_amplitude;
};

class Synthesize {   // Iterator<Item = f32>: next() returns false at the end of the utterance
public:
    explicit Synthesize(std::vector<float> buf) : buf_(std::move(buf)) {}
    bool next(float& x) { if (pos_ >= buf_.size()) return false; x = buf_[pos_++]; return true; }
    const std::vector<float>& samples() const { return buf_; }
private:
    std::vector<float> buf_;
    size_t pos_ = 0;
};

class Jitter {
public:
    Jitter(std::vector<grail_seq_elem> e, float sample_rate, uint32_t seed, Voice v) : elems_(std::move(e)), rate_(sample_rate), seed_(seed), v_(v) {}
    Synthesize synthesize(Context& ctx) const
    {
        grail_voice_params vp{ rate_, v_.jitter_frequency, v_.jitter_delta_frequency, v_.jitter_delta_formant_frequency,
                               v_.jitter_delta_amplitude, seed_, 0u };
        const uint32_t offs[2] = { 0u, (uint32_t)elems_.size() };
        uint64_t n = 0;
        int rc = grail_cuda_count_samples(elems_.data(), offs, &vp, 1, &n);
        if (rc) throw Error(rc, "grail_cuda_count_samples");
        std::vector<float> out(n);
        const uint64_t oo[2] = { 0, n };
        ctx.check(grail_cuda_synthesize_batch(ctx.get(), elems_.data(), offs, &vp, 1, out.data(), oo, 0));
        return Synthesize(std::move(out));
    }
private:
    std::vector<grail_seq_elem> elems_;
    float rate_;
    uint32_t seed_;
    Voice v_;
};

class Sequencer {
public:
    Sequencer(std::vector<grail_seq_elem> e, Voice v) : elems_(std::move(e)), v_(v) {}
    Jitter jitter(uint32_t seed, Voice v) && { return Jitter(std::move(elems_), v_.sample_rate, seed, v); }
private:
    std::vector<grail_seq_elem> elems_;
    Voice v_;
};

inline Sequencer sequence(std::vector<grail_seq_elem> elems, Voice voice) { return Sequencer(std::move(elems), voice); }

} // namespace grail
