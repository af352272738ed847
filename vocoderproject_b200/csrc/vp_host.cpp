// vp_host -- headless host for the B200 batch engine: the role VocoderAudioProcessor plays inside a DAW
// (Source/PluginProcessor.cpp:144-184, :203-234), for nStreams plug-in instances at once, written against the C++
// facade (vp_facade.hpp) and linked to libvp_engine.so through the C ABI only.
//
//   vp_host [--streams S] [--seconds T] [--fs 44100] [--block 1024] [--blocks-per-call K] [--key 12] [--voc 0|1]
//           [--pitch 0|1] [--gain-voice dB] [--gain-synth dB] [--device d] [--flavour f]
//
// Generates the seeded synthetic inputs (vp_synth_host), runs them block by block (K host blocks per processBlock
// call) and prints one JSON line: audio-seconds per wall second, CRC-32 of the left output (for cross-checks against
// the Python mirror), latency. Exit code 0 on success; any engine error is fatal (no CPU fallback).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vp_facade.hpp"

static uint32_t crc32_bytes(const unsigned char* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

int main(int argc, char** argv) {
    int S = 8, B = 1024, K = 1, keyPitch = 12, voc = 1, pitch = 1, device = 0, flavour = 0;
    double seconds = 2.0, fs = 44100.0;
    float gainVoice = -60.f, gainSynth = -60.f;
    for (int i = 1; i < argc; ++i) {
        auto arg = [&](const char* name) { return !strcmp(argv[i], name) && i + 1 < argc; };
        if (arg("--streams")) S = atoi(argv[++i]);
        else if (arg("--seconds")) seconds = atof(argv[++i]);
        else if (arg("--fs")) fs = atof(argv[++i]);
        else if (arg("--block")) B = atoi(argv[++i]);
        else if (arg("--blocks-per-call")) K = atoi(argv[++i]);
        else if (arg("--key")) keyPitch = atoi(argv[++i]);
        else if (arg("--voc")) voc = atoi(argv[++i]);
        else if (arg("--pitch")) pitch = atoi(argv[++i]);
        else if (arg("--gain-voice")) gainVoice = (float)atof(argv[++i]);
        else if (arg("--gain-synth")) gainSynth = (float)atof(argv[++i]);
        else if (arg("--device")) device = atoi(argv[++i]);
        else if (arg("--flavour")) flavour = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--help")) {
            printf("usage: vp_host [--streams S] [--seconds T] [--fs Hz] [--block B] [--blocks-per-call K] [--key 0..12] "
                   "[--voc 0|1] [--pitch 0|1] [--gain-voice dB] [--gain-synth dB] [--device d] [--flavour 0|1|2]\n");
            return 0;
        } else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    const size_t nBlocks = (size_t)(fs * seconds) / B / K * K;
    const size_t n = nBlocks * B;
    if (n == 0 || S <= 0) { fprintf(stderr, "nothing to do\n"); return 2; }
    try {
        std::vector<float> voice((size_t)S * n), sl((size_t)S * n), sr((size_t)S * n), outL((size_t)S * n), outR((size_t)S * n);
        if (vp_synth_host(fs, flavour, 0, S, n, n, voice.data(), sl.data(), sr.data()) != VP_OK) throw vpb200::Error(VP_E_ARG, "vp_synth_host");
        vpb200::VocoderBatchProcessor proc(device);
        proc.params.keyPitch = keyPitch; proc.params.vocBool = voc; proc.params.pitchBool = pitch;
        proc.params.gainVoice = gainVoice; proc.params.gainSynth = gainSynth;
        proc.prepareToPlay(fs, B, S, K);
        // block buffers [S][K*B] like a host's AudioBuffer per instance
        const size_t m = (size_t)K * B;
        std::vector<float> bv(S * m), bl(S * m), br(S * m), ol(S * m), orr(S * m);
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t b = 0; b < nBlocks; b += K) {
            for (int s = 0; s < S; ++s) {
                memcpy(&bv[s * m], &voice[(size_t)s * n + b * B], m * sizeof(float));
                memcpy(&bl[s * m], &sl[(size_t)s * n + b * B], m * sizeof(float));
                memcpy(&br[s * m], &sr[(size_t)s * n + b * B], m * sizeof(float));
            }
            proc.processBlock(bv.data(), bl.data(), br.data(), ol.data(), orr.data(), m, K);
            for (int s = 0; s < S; ++s) {
                memcpy(&outL[(size_t)s * n + b * B], &ol[s * m], m * sizeof(float));
                memcpy(&outR[(size_t)s * n + b * B], &orr[s * m], m * sizeof(float));
            }
        }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t launches = 0;
        vp_engine_get_stats(proc.myBuffer.engine.h, &launches, nullptr, nullptr);
        printf("{\"host\": \"vp_host\", \"streams\": %d, \"sample_rate\": %.1f, \"block\": %d, \"blocks_per_call\": %d, \"blocks\": %zu, "
               "\"latency_samples\": %d, \"audio_s_per_s\": %.3f, \"wall_s\": %.4f, \"kernel_launches\": %llu, "
               "\"crc_outL\": \"%08x\", \"crc_outR\": \"%08x\"}\n",
               S, fs, B, K, nBlocks, proc.getLatencySamples(), (double)S * n / fs / wall, wall, (unsigned long long)launches,
               crc32_bytes((const unsigned char*)outL.data(), outL.size() * 4), crc32_bytes((const unsigned char*)outR.data(), outR.size() * 4));
    } catch (const vpb200::Error& e) {
        fprintf(stderr, "vp_host: %s\n", e.what());
        return 1;
    }
    return 0;
}
