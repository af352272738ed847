// Host side of the synthetic input generator (vp_synth.h): the per-stream parameter table and the plain fill loop.
// Header-only so that two translation units build it from one definition: the engine (vp_synth_host / vp_synth_device of
// the C ABI) and tools/inputgen.cpp, the stand-alone generator that bench.py's reference arm loads instead of the product
// library. Host code, double arithmetic + libm for the table, vp_synth.h's explicitly rounded float arithmetic per sample.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "vp_synth.h"

static inline void vps_make_streams(double fs, int flavour, int first, int S, size_t nSamples, std::vector<vp_synth_stream>& out) {
    static const double vowels[5][4] = {{700, 1220, 2600, 3300}, {400, 2000, 2550, 3400}, {300, 2300, 3000, 3500},
                                        {450, 800, 2830, 3300}, {325, 700, 2530, 3400}};
    static const double bws[4] = {130, 70, 160, 250};
    out.resize(S);
    for (int i = 0; i < S; ++i) {
        vp_synth_stream& p = out[i];
        memset(&p, 0, sizeof p);
        const uint32_t seed = 0x5EED0000u + (uint32_t)(first + i);
        auto u = [&](int k) { return (double)vps_hash(seed ^ 0xA5A5A5A5u, (uint64_t)k) / 4294967296.0; };
        p.seed = seed;
        const double f0 = 120.0 * pow(330.0 / 120.0, u(0));
        p.f0inc = (float)(f0 / fs);
        p.glideInc = (uint32_t)((0.1 + 0.2 * u(1)) / fs * 4294967296.0);
        p.vibInc = (uint32_t)((5.0 + u(2)) / fs * 4294967296.0);
        p.glidePh0 = (uint32_t)(u(3) * 4294967296.0);
        p.vibPh0 = (uint32_t)(u(4) * 4294967296.0);
        p.glideDepth = (float)(3.0 / 12.0);
        p.vibDepth = (float)(0.3 / 12.0);
        p.tilt = (float)(0.90 + 0.05 * u(5));
        const int vw = (int)(u(6) * 5.0) % 5;
        const double scale = 0.9 + 0.25 * u(7);
        for (int k = 0; k < VPS_NFORMANTS; ++k) {
            const double fr = vowels[vw][k] * scale, bw = bws[k];
            const double r = exp(-3.14159265358979323846 * bw / fs), th = 2.0 * 3.14159265358979323846 * fr / fs;
            const double a1 = -2.0 * r * cos(th), a2 = r * r;
            p.a1[k] = (float)a1; p.a2[k] = (float)a2;
            p.b0[k] = (float)(1.0 + a1 + a2);
        }
        p.noiseAmp = (flavour == 1) ? 1e-4f : 1e-2f;
        static const double chord[VPS_NSAW + 1] = {0, 4, 7, 12, 10};
        const int root = (int)(u(8) * 12.0) % 12;
        for (int k = 0; k <= VPS_NSAW; ++k) {
            const double f = 130.8127826502993 * pow(2.0, (root + chord[k]) / 12.0);
            p.sawInc[k] = (uint32_t)(f / fs * 4294967296.0);
            p.sawPh0[k] = (uint32_t)(u(9 + k) * 4294967296.0);
        }
        p.sawAmp = 0.25f / VPS_NSAW;
        p.muteStart = p.muteEnd = 0;
        if (flavour == 2 && nSamples > 0) {
            p.muteStart = (int64_t)(nSamples * (0.35 + 0.1 * u(20)));
            p.muteEnd = p.muteStart + (int64_t)(fs * (0.25 + 0.2 * u(21)));
        }
        // calibrate the output gain so that the voiced peak sits near 0.5
        p.gain = 1.0f;
        const float keepNoise = p.noiseAmp;
        p.noiseAmp = 0.f;
        const int64_t ms = p.muteStart, me = p.muteEnd;
        p.muteStart = p.muteEnd = 0;
        vp_synth_state stt;
        vps_init(&p, &stt);
        float peak = 1e-9f;
        for (int j = 0; j < 4096; ++j) {
            float a, b, c;
            vps_step(&p, &stt, j, &a, &b, &c);
            if (j >= 1024) peak = fmaxf(peak, fabsf(a));
        }
        p.gain = 0.5f / peak;
        p.noiseAmp = keepNoise;
        p.muteStart = ms; p.muteEnd = me;
    }
}

// voice / synthL / synthR: [S][stride] floats (synthL, synthR optional). Returns false on bad arguments.
static inline bool vps_fill_host(double fs, int flavour, int first, int S, size_t nSamples, size_t stride, float* voice,
                                 float* synthL, float* synthR) {
    if (!voice || S <= 0 || stride < nSamples || !(fs > 0)) return false;
    std::vector<vp_synth_stream> ps;
    vps_make_streams(fs, flavour, first, S, nSamples, ps);
    for (int s = 0; s < S; ++s) {
        vp_synth_state st;
        vps_init(&ps[s], &st);
        for (size_t i = 0; i < nSamples; ++i) {
            float a, b, c;
            vps_step(&ps[s], &st, (int64_t)i, &a, &b, &c);
            voice[(size_t)s * stride + i] = a;
            if (synthL) synthL[(size_t)s * stride + i] = b;
            if (synthR) synthR[(size_t)s * stride + i] = c;
        }
    }
    return true;
}
