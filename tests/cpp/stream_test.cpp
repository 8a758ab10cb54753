// The lazy streaming adaptor of the C++ facade (grail-rs_b200/cpp/grail.hpp, StreamSynthesize) in the shape of
// examples/interactive.rs:31-48: an INFINITE upstream (`repeat_with`), the chain built lazily, channel duplication,
// and an audio callback that pulls a fixed number of frames.  A chain that collected its upstream would hang here.
//   stream_test <n_callbacks> <frames_per_callback> <channels>
// prints "pulled <n> upstream_calls <k> checksum <c> first_equal_channels <0|1>" and the samples of channel 0 to
// argv[4] (raw f32) when given; exits 3 without a device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../grail-rs_b200/cpp/grail.hpp"

static grail_seq_elem voiced(float f0)
{
    grail_seq_elem a{};
    a.has_elem = 1; a.length = 0.05f; a.blend_length = 0.05f;
    a.elem.frequency = f0 / 44100.0f;
    const float ff[8] = { 910, 1271, 2851, 3213, 1200, 2000, 3000, 4000 }, bw[8] = { 60, 160, 180, 200, 100, 100, 100, 100 };
    const float amp[8] = { 0.4f, 0.35f, 0.25f, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 8; ++i) {
        a.elem.formant_freq[i] = ff[i] / 44100.0f;
        a.elem.formant_bw[i] = bw[i] / 44100.0f;
        a.elem.formant_smooth[i] = 1600.0f / 44100.0f;
        a.elem.formant_breath[i] = 0.2f;
        a.elem.formant_turb[i] = 0.1f;
        a.elem.formant_amp[i] = amp[i];
    }
    return a;
}

int main(int argc, char** argv)
{
    if (grail_cuda_device_count() == 0) {
        std::printf("no-device\n");
        return 3;
    }
    const long n_callbacks = argc > 1 ? std::atol(argv[1]) : 50;
    const long frames = argc > 2 ? std::atol(argv[2]) : 441;
    const unsigned channels = argc > 3 ? (unsigned)std::atoi(argv[3]) : 2;
    grail::Voice v{ 44100.0f, 16.0f / 44100.0f, 6.0f / 44100.0f, 6.0f / 44100.0f, 0.2f };
    grail::Context ctx(0);
    long upstream_calls = 0;
    // repeat_with: never ends; silence, then voiced elements of alternating pitch (element k depends only on k)
    auto source = [&upstream_calls](grail_seq_elem& e) -> bool {
        const long k = upstream_calls++;
        if (k % 5 == 0) { e = grail_seq_elem{}; e.length = 0.05f; e.blend_length = 0.05f; }
        else e = voiced(k % 2 ? 120.0f : 150.0f);
        return true;
    };
    auto it = grail::RepeatChannels<grail::StreamSynthesize<decltype(source)>>(
        grail::sequence_from(source, v).jitter(0u, v).synthesize(ctx, (size_t)frames), channels);
    if (upstream_calls != 0) { std::printf("the chain pulled its upstream before the first next()\n"); return 1; }
    std::vector<float> ch0;
    double sum = 0.0;
    bool equal = true;
    for (long cb = 0; cb < n_callbacks; ++cb) {          // the audio callback: `for i in data { *i = iterator.next().unwrap_or(0.0) }`
        for (long f = 0; f < frames; ++f) {
            float first = 0.0f;
            for (unsigned c = 0; c < channels; ++c) {
                float x = 0.0f;
                if (!it.next(x)) { std::printf("iterator ended: %s\n", it.inner().error().c_str()); return 1; }
                if (c == 0) { first = x; ch0.push_back(x); sum += std::fabs((double)x); }
                else equal = equal && (x == first);
            }
        }
    }
    if (argc > 4) {
        FILE* fp = std::fopen(argv[4], "wb");
        if (!fp) return 1;
        std::fwrite(ch0.data(), sizeof(float), ch0.size(), fp);
        std::fclose(fp);
    }
    std::printf("pulled %zu upstream_calls %ld checksum %.9g first_equal_channels %d\n", ch0.size(), upstream_calls, sum, equal ? 1 : 0);
    return std::isfinite(sum) ? 0 : 1;
}
