// CPU-only check of vocoderproject_b200/csrc/vp_wav.hpp: reads a WAV, prints its shape and a few samples as JSON, and
// re-writes it as float32 and PCM16 (tests/test_wav.py compares against Python's own decoders).
#include <cstdio>
#include <string>
#include <vector>

#include "vp_wav.hpp"

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: wav_roundtrip in.wav out_f32.wav out_pcm16.wav\n"); return 2; }
    try {
        const vpb200::WavData w = vpb200::wav_read(argv[1]);
        std::vector<const float*> ch;
        for (const auto& c : w.ch) ch.push_back(c.data());
        vpb200::wav_write(argv[2], w.sampleRate, ch, w.frames, false);
        vpb200::wav_write(argv[3], w.sampleRate, ch, w.frames, true);
        double sum = 0;
        for (const auto& c : w.ch) for (float v : c) sum += (double)v * v;
        std::printf("{\"sample_rate\": %d, \"channels\": %d, \"frames\": %zu, \"energy\": %.9e}\n", w.sampleRate, w.channels, w.frames, sum);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
