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

// ------------------------------------------------------------------------------------------------
// Lazy chain over an ARBITRARY, possibly infinite upstream (examples/interactive.rs:31-38: the source is
// `repeat_with(...)`, it never ends, and the audio callback pulls a few hundred samples at a time).
// `Upstream` is any callable `bool(grail_seq_elem&)`: true = one more SequenceElem, false = the upstream is over.
// StreamSynthesize::next is the reference's `Iterator<Item = f32>::next`: it hands samples out of a window buffer and
// refills the window through grail_cuda_stream_{push,pull}, pulling from the upstream only as far as the stream's
// one-element look-ahead needs.  It never throws: an error ends the iterator (next() returns false) and is kept in
// error(), which is what the Rust facade's `Option<f32>` does (the reference's hot path has no error channel).
// This is the exact adaptor logic of grail-rs-cuda's `Synthesize<T>` (rust/grail-rs-cuda/src/lib.rs), which cannot be
// compiled in the build image; tests/cpp/stream_test.cpp runs THIS one against an infinite generator.
// ------------------------------------------------------------------------------------------------
template <class Upstream>
class StreamSynthesize {
public:
    StreamSynthesize(Context& ctx, Upstream up, float sample_rate, uint32_t seed, Voice v, size_t window = 2048)
        : ctx_(&ctx), up_(std::move(up)), window_(window ? window : 1)
    {
        vp_ = grail_voice_params{ sample_rate, v.jitter_frequency, v.jitter_delta_frequency,
                                  v.jitter_delta_formant_frequency, v.jitter_delta_amplitude, seed, 0u };
    }
    StreamSynthesize(StreamSynthesize&& o) noexcept
        : ctx_(o.ctx_), up_(std::move(o.up_)), vp_(o.vp_), window_(o.window_), s_(o.s_), buf_(std::move(o.buf_)), pos_(o.pos_),
          n_(o.n_), up_done_(o.up_done_), ended_(o.ended_), status_(o.status_), err_(std::move(o.err_))
    {
        o.s_ = nullptr;
        o.ended_ = true;
    }
    StreamSynthesize& operator=(StreamSynthesize&&) = delete;   // (closures are not assignable)
    StreamSynthesize(const StreamSynthesize&) = delete;
    StreamSynthesize& operator=(const StreamSynthesize&) = delete;
    ~StreamSynthesize() { release(); }

    bool next(float& x)
    {
        if (pos_ >= n_ && !refill()) return false;
        x = buf_[pos_++];
        return true;
    }
    int status() const { return status_; }              // GRAIL_OK unless an error ended the iterator
    const std::string& error() const { return err_; }

private:
    void release() { if (s_) { grail_cuda_stream_free(s_); s_ = nullptr; } }
    bool fail(int rc)
    {
        status_ = rc;
        err_ = grail_cuda_last_error(ctx_->get());
        ended_ = true;
        return false;
    }
    bool refill()
    {
        if (ended_) return false;
        if (!s_) {                                       // created lazily: a chain that is never pulled costs nothing
            const int rc = grail_cuda_stream_new(ctx_->get(), &vp_, &s_);
            if (rc) return fail(rc);
            buf_.resize(window_);
        }
        for (;;) {
            uint64_t got = 0;
            int rc = grail_cuda_stream_pull(s_, buf_.data(), window_, &got);
            if (rc) return fail(rc);
            if (got) { pos_ = 0; n_ = (size_t)got; return true; }
            if (up_done_) { ended_ = true; return false; }   // finished and drained: None, like the reference
            // the stream has run dry: it needs one more upstream element (its look-ahead) before it can go on
            grail_seq_elem e;
            if (up_(e)) rc = grail_cuda_stream_push(s_, &e, 1);
            else { up_done_ = true; rc = grail_cuda_stream_finish(s_); }
            if (rc) return fail(rc);
        }
    }
    Context* ctx_ = nullptr;
    Upstream up_;
    grail_voice_params vp_{};
    size_t window_ = 2048;
    grail_stream* s_ = nullptr;
    std::vector<float> buf_;
    size_t pos_ = 0, n_ = 0;
    bool up_done_ = false, ended_ = false;
    int status_ = GRAIL_OK;
    std::string err_;
};

template <class Upstream>
class LazyJitter {
public:
    LazyJitter(Upstream up, float sample_rate, uint32_t seed, Voice v) : up_(std::move(up)), rate_(sample_rate), seed_(seed), v_(v) {}
    // `window`: samples synthesized per device round trip (an audio callback's buffer size is a good value)
    StreamSynthesize<Upstream> synthesize(Context& ctx, size_t window = 2048) &&
    {
        return StreamSynthesize<Upstream>(ctx, std::move(up_), rate_, seed_, v_, window);
    }
private:
    Upstream up_;
    float rate_;
    uint32_t seed_;
    Voice v_;
};
template <class Upstream>
class LazySequencer {
public:
    LazySequencer(Upstream up, Voice v) : up_(std::move(up)), v_(v) {}
    LazyJitter<Upstream> jitter(uint32_t seed, Voice v) && { return LazyJitter<Upstream>(std::move(up_), v_.sample_rate, seed, v); }
private:
    Upstream up_;
    Voice v_;
};
// sequence_from(upstream, voice).jitter(seed, voice).synthesize(ctx): nothing is pulled until the first next()
template <class Upstream>
LazySequencer<Upstream> sequence_from(Upstream up, Voice voice) { return LazySequencer<Upstream>(std::move(up), voice); }

// channel duplication, `.flat_map(move |x| std::iter::repeat(x).take(num_channels))` (examples/interactive.rs:38)
template <class It>
class RepeatChannels {
public:
    RepeatChannels(It it, unsigned channels) : it_(std::move(it)), ch_(channels ? channels : 1) {}
    bool next(float& x)
    {
        if (left_ == 0) {
            if (!it_.next(cur_)) return false;
            left_ = ch_;
        }
        --left_;
        x = cur_;
        return true;
    }
    It& inner() { return it_; }
private:
    It it_;
    unsigned ch_, left_ = 0;
    float cur_ = 0.0f;
};

} // namespace grail
