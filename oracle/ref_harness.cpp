// TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
//
// Headless driver around the reference's own, unmodified C++ (compiled in
// place from /root/reference/Source by oracle/Makefile into oracle/_ref/).
// It instantiates the real VocoderAudioProcessor, feeds it blocks exactly as a
// DAW would (PluginProcessor.cpp:144-184 prepareToPlay, :203-234 processBlock)
// and exposes a C ABI so tests/ and bench.py's cpu_baseline / --impl reference
// legs can call it through ctypes.
//
// Two drive modes:
//   log == 0 : the true processBlock (the ground truth and the timed baseline).
//   log == 1 : processBlock's 30 lines and the 25-line PitchProcess::process /
//              VocoderProcess::process schedulers restated here (private members
//              reached with `#define private public`) so that per-frame integer
//              decisions can be logged; tests assert log==1 audio is bit-identical
//              to log==0 audio.
//
// Build flags that matter (oracle/Makefile): -include math.h -include stdlib.h
// (restores ::abs(double)), -ffp-contract=off, and the
// zero-filling operator new below (stale std::vector slots read by
// PitchProcess.cpp:818 become deterministic zeros).
#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

void* operator new(std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](std::size_t n) {
    void* p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

#define private public
#include "PluginProcessor.h"
#undef private

static int g_window = 0;  // 0 = "sine" (what prepareToPlay passes, PluginProcessor.cpp:165), 1 = "hann"

extern "C" {

// Selects the vocoder window type for the instances created afterwards (process-global). "hann" is reachable in the
// reference only by editing the string literal in prepareToPlay; the harness calls the reference's own setWindows("hann")
// right after prepareToPlay instead.
void vpref_set_window(int w) { g_window = w; }

// Mirrors include/vp_engine.h's vp_params field for field (same order).
struct vpref_params {
    float gainPitch, gainVoice, gainSynth, gainVoc;  // dB
    int lpcVoice, lpcPitch, lpcSynth;
    int keyPitch;  // Notes::key, 12 = chromatic
    int pitchBool, vocBool;
};

#define VPREF_MAX_MARKS 64
struct vpref_pitch_frame {
    int frame;        // index of the processChunkStart call
    int startSample;  // PitchProcess::startSample at that call
    int block;        // host block index
    int gated;        // 1 = silence gate fired (PitchProcess.cpp:208-214)
    int period, periodNew, prevVoicedPeriod;
    int note;  // index of closestFreq in notes.freq, -1 if unvoiced / no marks
    int nAn, nSt;
    int anStale;  // storage slot anMarks[size] (what PitchProcess.cpp:818 can read)
    int anMarks[VPREF_MAX_MARKS];
    int stMarks[VPREF_MAX_MARKS];
    double pitch, closestFreq, beta;
};
struct vpref_voc_frame {
    int frame, startSample, block, gated;
    double EeVoice, EeSynth, g;
};

static void set_params(VocoderAudioProcessor& p, const vpref_params& q) {
    auto set = [&](const char* id, float v) { p.treeState.getRawParameterValue(id)->store(v); };
    set("gainPitch", q.gainPitch);
    set("gainVoice", q.gainVoice);
    set("gainSynth", q.gainSynth);
    set("gainVoc", q.gainVoc);
    set("lpcVoice", (float)q.lpcVoice);
    set("lpcPitch", (float)q.lpcPitch);
    set("lpcSynth", (float)q.lpcSynth);
    set("keyPitch", (float)q.keyPitch);
    set("pitchBool", q.pitchBool ? 1.f : 0.f);
    set("vocBool", q.vocBool ? 1.f : 0.f);
}

void vpref_default_params(vpref_params* q) {
    q->gainPitch = 0.f; q->gainVoice = -60.f; q->gainSynth = -60.f; q->gainVoc = 0.f;
    q->lpcVoice = 40; q->lpcPitch = 15; q->lpcSynth = 5; q->keyPitch = 12;
    q->pitchBool = 1; q->vocBool = 1;
}

// Sizes prepareToPlay derives (PluginProcessor.cpp:160-176), for cross-checks.
struct vpref_sizes { int hopV, wlenV, hopP, frameLenP, chunk, tauMax, latency, keep, inSize, outSize, anCap, nFreq; };

static void fill_sizes(VocoderAudioProcessor& p, vpref_sizes* s) {
    s->hopV = p.vocoderProcess.hop; s->wlenV = p.vocoderProcess.wlen;
    s->hopP = p.pitchProcess.hop; s->frameLenP = p.pitchProcess.frameLen;
    s->chunk = p.pitchProcess.chunkSize; s->tauMax = p.pitchProcess.tauMax;
    s->latency = p.myBuffer.latency; s->keep = p.myBuffer.samplesToKeep;
    s->inSize = p.myBuffer.inSize; s->outSize = p.myBuffer.outSize;
    s->anCap = (int)p.pitchProcess.anMarks.capacity();
    s->nFreq = (int)p.pitchProcess.notes.freq.size();
}

static void log_pitch_frame(VocoderAudioProcessor& p, int frame, int block, int startSample, bool gated,
                            std::vector<vpref_pitch_frame>& log) {
    PitchProcess& pp = p.pitchProcess;
    vpref_pitch_frame f;
    std::memset(&f, 0, sizeof f);
    f.frame = frame; f.startSample = startSample; f.block = block; f.gated = gated ? 1 : 0;
    f.period = pp.period; f.periodNew = pp.periodNew; f.prevVoicedPeriod = pp.prevVoicedPeriod;
    f.pitch = pp.pitch; f.closestFreq = pp.closestFreq; f.beta = pp.beta;
    f.note = -1;
    if (!gated && pp.pitch > 1 && !pp.anMarks.empty()) {
        const auto& fr = pp.notes.freq;
        for (size_t i = 0; i < fr.size(); ++i)
            if (fr[i] == pp.closestFreq) f.note = (int)i;
        // U6: closestFreq may be the popped slot just past the table's end.
        if (f.note < 0) f.note = (int)fr.size();
    }
    f.nAn = (int)pp.anMarks.size(); f.nSt = (int)pp.stMarks.size();
    for (int i = 0; i < f.nAn && i < VPREF_MAX_MARKS; ++i) f.anMarks[i] = pp.anMarks[i];
    for (int i = 0; i < f.nSt && i < VPREF_MAX_MARKS; ++i) f.stMarks[i] = pp.stMarks[i];
    f.anStale = (pp.anMarks.capacity() > pp.anMarks.size()) ? pp.anMarks.data()[pp.anMarks.size()] : 0;
    log.push_back(f);
}

// processBlock restated with logging (PluginProcessor.cpp:203-234,
// VocoderProcess.cpp:173-183, PitchProcess.cpp:166-196).
static void process_block_logged(VocoderAudioProcessor& p, AudioBuffer<float>& buffer, int block,
                                 std::vector<vpref_pitch_frame>& plog, std::vector<vpref_voc_frame>& vlog,
                                 int& pframe, int& vframe) {
    auto nOutputChannels = p.getTotalNumOutputChannels();
    auto voiceBuffer = p.getBusBuffer(buffer, true, 0);
    auto synthBuffer = p.getBusBuffer(buffer, true, 1);
    MyBuffer& mb = p.myBuffer;
    mb.fillInputBuffers(voiceBuffer, synthBuffer);

    if (p.treeState.getRawParameterValue("vocBool")->load()) {
        VocoderProcess& vp = p.vocoderProcess;
        while (vp.startSample < mb.getSamplesPerBlock()) {
            vpref_voc_frame f;
            std::memset(&f, 0, sizeof f);
            f.frame = vframe++; f.startSample = vp.startSample; f.block = block;
            double rv = Decibels::gainToDecibels(mb.getRMSLevelVoiceFull());
            double rs = Decibels::gainToDecibels(mb.getRMSLevelSynthFull());
            f.gated = (rv < vp.silenceThresholdDb || rs < vp.silenceThresholdDb) ? 1 : 0;
            vp.processWindow(mb);
            f.EeVoice = vp.EeVoice; f.EeSynth = vp.EeSynth; f.g = vp.g;
            vlog.push_back(f);
            vp.startSample += vp.hop;
        }
        vp.startSample -= mb.getSamplesPerBlock();
    }

    if (p.treeState.getRawParameterValue("pitchBool")->load()) {
        PitchProcess& pp = p.pitchProcess;
        auto start = [&]() {
            bool gated = Decibels::gainToDecibels(mb.getRMSLevelVoiceFull()) < pp.silenceThresholdDb;
            pp.processChunkStart(mb);
            log_pitch_frame(p, pframe++, block, pp.startSample, gated, plog);
        };
        while (pp.startSample < mb.getSamplesPerBlock()) {
            if (pp.nChunk % pp.chunksPerFrame == pp.chunksPerFrame - 1) {
                pp.processChunkCont(mb);
                pp.nChunk = 0;
                start();
                pp.nChunk += 1;
                pp.nChunk %= pp.chunksPerFrame;
            } else if (pp.nChunk == 0) {
                start();
                pp.nChunk += 1;
            } else {
                pp.processChunkCont(mb);
                pp.nChunk += 1;
            }
            pp.startSample += pp.chunkSize;
        }
        pp.startSample -= mb.getSamplesPerBlock();
    } else {
        p.pitchProcess.silence();
    }

    auto gainVoice = p.treeState.getRawParameterValue("gainVoice");
    auto gainSynth = p.treeState.getRawParameterValue("gainSynth");
    if (gainVoice->load() > -59.0) mb.addDryVoice(Decibels::decibelsToGain(gainVoice->load(), -59.0f));
    if (gainSynth->load() > -59.0) mb.addSynth(Decibels::decibelsToGain(gainSynth->load(), -59.0f));
    mb.fillOutputBuffer(buffer, nOutputChannels);
}

// Run one stream. voice/synthL/synthR: nBlocks*B floats each (synthR may be
// NULL -> copy of synthL). outL/outR: nBlocks*B floats. plog/vlog may be NULL
// when log == 0. Parameter automation: before block schedBlock[i] the i-th
// entry of sched is stored into the parameter tree (what a DAW's automation
// does between two processBlock calls); q applies from prepareToPlay on.
// Returns 0, or -1 on bad arguments.
static int run_impl(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                    const vpref_params* q, const vpref_params* sched, const int* schedBlock, int nSched, int log,
                    float* outL, float* outR, vpref_sizes* sizes, vpref_pitch_frame* plog, int plogCap, int* nP,
                    vpref_voc_frame* vlog, int vlogCap, int* nV) {
    if (!voice || !synthL || !q || !outL || B <= 0 || nBlocks < 0) return -1;
    if (nSched > 0 && (!sched || !schedBlock)) return -1;
    if (!synthR) synthR = synthL;
    VocoderAudioProcessor proc;
    set_params(proc, *q);
    proc.prepareToPlay(fs, B);
    if (g_window == 1) proc.vocoderProcess.setWindows("hann");  // the branch prepareToPlay never selects (VocoderProcess.cpp:116-124)
    if (sizes) fill_sizes(proc, sizes);
    AudioBuffer<float> buf(3, B);
    MidiBuffer midi;
    std::vector<vpref_pitch_frame> pl;
    std::vector<vpref_voc_frame> vl;
    int pframe = 0, vframe = 0;
    for (int b = 0; b < nBlocks; ++b) {
        for (int i = 0; i < nSched; ++i)
            if (schedBlock[i] == b) set_params(proc, sched[i]);
        std::memcpy(buf.getWritePointer(0), voice + (size_t)b * B, sizeof(float) * B);
        std::memcpy(buf.getWritePointer(1), synthL + (size_t)b * B, sizeof(float) * B);
        std::memcpy(buf.getWritePointer(2), synthR + (size_t)b * B, sizeof(float) * B);
        if (log) process_block_logged(proc, buf, b, pl, vl, pframe, vframe);
        else proc.processBlock(buf, midi);
        std::memcpy(outL + (size_t)b * B, buf.getReadPointer(0), sizeof(float) * B);
        if (outR) std::memcpy(outR + (size_t)b * B, buf.getReadPointer(1), sizeof(float) * B);
    }
    if (nP) *nP = (int)pl.size();
    if (nV) *nV = (int)vl.size();
    if (plog) for (int i = 0; i < (int)pl.size() && i < plogCap; ++i) plog[i] = pl[i];
    if (vlog) for (int i = 0; i < (int)vl.size() && i < vlogCap; ++i) vlog[i] = vl[i];
    return 0;
}

int vpref_run(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
              const vpref_params* q, int log, float* outL, float* outR, vpref_sizes* sizes,
              vpref_pitch_frame* plog, int plogCap, int* nP, vpref_voc_frame* vlog, int vlogCap, int* nV) {
    return run_impl(fs, B, nBlocks, voice, synthL, synthR, q, nullptr, nullptr, 0, log, outL, outR, sizes, plog, plogCap, nP,
                    vlog, vlogCap, nV);
}

int vpref_run_sched(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                    const vpref_params* q, const vpref_params* sched, const int* schedBlock, int nSched, int log,
                    float* outL, float* outR, vpref_sizes* sizes, vpref_pitch_frame* plog, int plogCap, int* nP,
                    vpref_voc_frame* vlog, int vlogCap, int* nV) {
    return run_impl(fs, B, nBlocks, voice, synthL, synthR, q, sched, schedBlock, nSched, log, outL, outR, sizes, plog, plogCap,
                    nP, vlog, vlogCap, nV);
}

// Notes table as the reference builds it (Notes.cpp:43-70), incl. the popped
// slot just past the end (U6). Returns the table size.
int vpref_notes(int key, double fMin, double fMax, double* freq, int cap, double* popped) {
    Notes n;
    n.prepare((Notes::key)key, fMin, fMax);
    int sz = (int)n.freq.size();
    for (int i = 0; i < sz && i < cap; ++i) freq[i] = n.freq[i];
    if (popped) *popped = n.freq.data()[sz];
    return sz;
}

double vpref_closest_freq(int key, double fMin, double fMax, double pitch) {
    Notes n;
    n.prepare((Notes::key)key, fMin, fMax);
    return n.getClosestFreq(pitch, (Notes::key)key);
}

// CPU baseline: S independent plug-in instances over disjoint streams on
// nThreads host threads, timing processBlock only (inputs resident, outputs
// written). Layout: voice[S][n], synthL[S][n], synthR[S][n] (synthR may be
// NULL), out[S][2][n] (may be NULL -> discarded), n = nBlocks*B.
// Returns wall seconds of the processing region.
double vpref_bench(double fs, int B, int nBlocks, int S, const float* voice, const float* synthL,
                   const float* synthR, const vpref_params* q, int nThreads, float* out) {
    if (nThreads < 1) nThreads = 1;
    const size_t n = (size_t)nBlocks * B;
    std::vector<std::unique_ptr<VocoderAudioProcessor>> procs(S);
    for (int s = 0; s < S; ++s) {
        procs[s].reset(new VocoderAudioProcessor());
        set_params(*procs[s], *q);
        procs[s]->prepareToPlay(fs, B);
        if (g_window == 1) procs[s]->vocoderProcess.setWindows("hann");
    }
    std::atomic<int> next(0);
    auto work = [&]() {
        AudioBuffer<float> buf(3, B);
        MidiBuffer midi;
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= S) break;
            const float* v = voice + (size_t)s * n;
            const float* l = synthL + (size_t)s * n;
            const float* r = synthR ? synthR + (size_t)s * n : l;
            for (int b = 0; b < nBlocks; ++b) {
                std::memcpy(buf.getWritePointer(0), v + (size_t)b * B, sizeof(float) * B);
                std::memcpy(buf.getWritePointer(1), l + (size_t)b * B, sizeof(float) * B);
                std::memcpy(buf.getWritePointer(2), r + (size_t)b * B, sizeof(float) * B);
                procs[s]->processBlock(buf, midi);
                if (out) {
                    std::memcpy(out + ((size_t)s * 2 + 0) * n + (size_t)b * B, buf.getReadPointer(0), sizeof(float) * B);
                    std::memcpy(out + ((size_t)s * 2 + 1) * n + (size_t)b * B, buf.getReadPointer(1), sizeof(float) * B);
                }
            }
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nThreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
