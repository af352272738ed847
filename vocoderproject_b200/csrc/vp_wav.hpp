// vp_wav.hpp -- minimal RIFF/WAVE reader and writer for the batch front-end (pushing real recordings
// through the engine, the use case of the reference's notebook, Notebook/*.ipynb cells 3 and 22, which loads WAVs as
// float arrays). Header-only, no dependencies. Reads PCM 8/16/24/32-bit integer and IEEE float 32/64 (plain and
// WAVE_FORMAT_EXTENSIBLE), any channel count; writes IEEE float32 or PCM16. Integer PCM maps to [-1, 1) by 2^-(bits-1),
// the convention of the usual decoders (8-bit is unsigned, offset 128).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vpb200 {

struct WavData {
    int sampleRate = 0;
    int channels = 0;
    size_t frames = 0;
    std::vector<std::vector<float>> ch;  // planar [channels][frames]
};

namespace wavdetail {
inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline void wr32(unsigned char* p, uint32_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = (v >> 24) & 255; }
inline void wr16(unsigned char* p, uint16_t v) { p[0] = v & 255; p[1] = (v >> 8) & 255; }
}  // namespace wavdetail

inline WavData wav_read(const std::string& path) {
    using namespace wavdetail;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<unsigned char> buf;
    {
        std::fseek(f, 0, SEEK_END);
        const long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (sz < 12) { std::fclose(f); throw std::runtime_error(path + ": not a RIFF/WAVE file"); }
        buf.resize((size_t)sz);
        const size_t got = std::fread(buf.data(), 1, buf.size(), f);
        std::fclose(f);
        if (got != buf.size()) throw std::runtime_error(path + ": short read");
    }
    if (std::memcmp(buf.data(), "RIFF", 4) != 0 || std::memcmp(buf.data() + 8, "WAVE", 4) != 0)
        throw std::runtime_error(path + ": not a RIFF/WAVE file");
    int fmtTag = 0, channels = 0, rate = 0, bits = 0, blockAlign = 0;
    const unsigned char* data = nullptr;
    size_t dataLen = 0;
    size_t pos = 12;
    while (pos + 8 <= buf.size()) {
        const unsigned char* ck = buf.data() + pos;
        size_t len = rd32(ck + 4);
        const size_t body = pos + 8;
        if (body + len > buf.size()) len = buf.size() - body;  // truncated last chunk (streamed files): take what is there
        if (!std::memcmp(ck, "fmt ", 4) && len >= 16) {
            fmtTag = rd16(ck + 8); channels = rd16(ck + 10); rate = (int)rd32(ck + 12); blockAlign = rd16(ck + 20); bits = rd16(ck + 22);
            if (fmtTag == 0xFFFE && len >= 26) fmtTag = rd16(ck + 8 + 24);  // WAVE_FORMAT_EXTENSIBLE: sub-format GUID's first 2 bytes
        } else if (!std::memcmp(ck, "data", 4)) {
            data = buf.data() + body; dataLen = len;
            break;
        }
        pos = body + len + (len & 1);
    }
    if (!data || channels <= 0 || rate <= 0) throw std::runtime_error(path + ": missing fmt / data chunk");
    const int bytes = bits / 8;
    if (!((fmtTag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32)) || (fmtTag == 3 && (bits == 32 || bits == 64))))
        throw std::runtime_error(path + ": unsupported sample format (tag " + std::to_string(fmtTag) + ", " + std::to_string(bits) + " bits)");
    if (blockAlign < bytes * channels) blockAlign = bytes * channels;
    WavData w;
    w.sampleRate = rate; w.channels = channels; w.frames = dataLen / (size_t)blockAlign;
    w.ch.assign((size_t)channels, std::vector<float>(w.frames));
    for (size_t i = 0; i < w.frames; ++i) {
        const unsigned char* fr = data + i * (size_t)blockAlign;
        for (int c = 0; c < channels; ++c) {
            const unsigned char* p = fr + (size_t)c * bytes;
            float v;
            if (fmtTag == 3) {
                if (bits == 32) { std::memcpy(&v, p, 4); }
                else { double d; std::memcpy(&d, p, 8); v = (float)d; }
            } else if (bits == 8) v = ((int)p[0] - 128) * (1.0f / 128.0f);
            else if (bits == 16) v = (float)(int16_t)rd16(p) * (1.0f / 32768.0f);
            else if (bits == 24) { int32_t x = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24) >> 8; v = (float)x * (1.0f / 8388608.0f); }
            else v = (float)((double)(int32_t)rd32(p) * (1.0 / 2147483648.0));
            w.ch[(size_t)c][i] = v;
        }
    }
    return w;
}

// planar channels (all of `frames` samples) -> interleaved IEEE float32 (pcm16 = false) or dithered-less rounded PCM16
inline void wav_write(const std::string& path, int sampleRate, const std::vector<const float*>& ch, size_t frames, bool pcm16 = false) {
    using namespace wavdetail;
    const int channels = (int)ch.size(), bytes = pcm16 ? 2 : 4;
    const uint64_t dataLen64 = (uint64_t)frames * channels * bytes;
    if (channels <= 0 || dataLen64 > 0xFFFFFF00ull) throw std::runtime_error(path + ": too large for a RIFF file");
    const uint32_t dataLen = (uint32_t)dataLen64;
    unsigned char h[44];
    std::memcpy(h, "RIFF", 4); wr32(h + 4, 36 + dataLen); std::memcpy(h + 8, "WAVEfmt ", 8); wr32(h + 16, 16);
    wr16(h + 20, pcm16 ? 1 : 3); wr16(h + 22, (uint16_t)channels); wr32(h + 24, (uint32_t)sampleRate);
    wr32(h + 28, (uint32_t)sampleRate * channels * bytes); wr16(h + 32, (uint16_t)(channels * bytes)); wr16(h + 34, (uint16_t)(bytes * 8));
    std::memcpy(h + 36, "data", 4); wr32(h + 40, dataLen);
    std::vector<unsigned char> body(dataLen);
    for (size_t i = 0; i < frames; ++i)
        for (int c = 0; c < channels; ++c) {
            unsigned char* p = body.data() + (i * channels + c) * bytes;
            const float v = ch[(size_t)c][i];
            if (pcm16) {
                float s = std::nearbyint(v * 32768.0f);
                s = s > 32767.0f ? 32767.0f : (s < -32768.0f ? -32768.0f : s);
                wr16(p, (uint16_t)(int16_t)s);
            } else std::memcpy(p, &v, 4);
        }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    const bool ok = std::fwrite(h, 1, 44, f) == 44 && std::fwrite(body.data(), 1, body.size(), f) == body.size();
    std::fclose(f);
    if (!ok) throw std::runtime_error("short write to " + path);
}

}  // namespace vpb200
