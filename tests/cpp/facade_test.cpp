// Exercises the C++ facade (grail-rs_b200/cpp/grail.hpp) the way examples/cli.rs:175-184 uses the reference:
//   elems.sequence(voice).jitter(seed, voice).synthesize()  drained into a vector.
// Prints "n_samples checksum" so the Python test can compare with the ctypes path; exits 3 when there is no device
// (the library has no CPU path).
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../grail-rs_b200/cpp/grail.hpp"

int main()
{
    if (grail_cuda_device_count() == 0) {
        std::printf("no-device\n");
        return 3;
    }
    // two hand-made phonemes: a silence and one voiced element with 3 formants
    grail_seq_elem sil{}, a{};
    sil.has_elem = 0; sil.length = 0.05f; sil.blend_length = 0.05f;
    a.has_elem = 1; a.length = 0.08f; a.blend_length = 0.08f;
    a.elem.frequency = 120.0f / 44100.0f;
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
    grail::Voice v{ 44100.0f, 16.0f / 44100.0f, 6.0f / 44100.0f, 6.0f / 44100.0f, 0.2f };
    grail::Context ctx(0);
    grail::Synthesize it = grail::sequence({ sil, a, a }, v).jitter(7u, v).synthesize(ctx);
    double sum = 0.0;
    size_t n = 0;
    float x;
    while (it.next(x)) { sum += std::fabs((double)x); ++n; }
    std::printf("%zu %.9g\n", n, sum);
    return (n > 0 && std::isfinite(sum)) ? 0 : 1;
}
