// Test harness for vocoderproject_b200/csrc/vp_juce_adapter.hpp: a JUCE AudioProcessor (against the oracle's JUCE stub,
// oracle/stub/JuceLibraryCode/JuceHeader.h -- test infrastructure) whose prepareToPlay / processBlock forward to the
// B200 engine, driven block by block like a DAW drives the reference plug-in. Written for this repo: the parameter ids,
// ranges and defaults restate createParameterLayout (Source/PluginProcessor.cpp:37-73), the bus layout :22-28.
//
//   b200_plugin in.f32 out.f32 fs B nBlocks [id=value ...] [@block id=value ...]
//   in.f32 : nBlocks*B frames of 3 interleaved-by-plane floats: voice[n], synthL[n], synthR[n] (planar)
//   out.f32: outL[n], outR[n] (planar)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "JuceHeader.h"
#include "vp_juce_adapter.hpp"

class B200VocoderProcessor : public AudioProcessor {
public:
    B200VocoderProcessor()
        : AudioProcessor(BusesProperties().withInput("Input", AudioChannelSet::mono(), true)
                             .withOutput("Output", AudioChannelSet::stereo(), true)
                             .withInput("Sidechain", AudioChannelSet::stereo())),
          treeState(*this, nullptr, "PARAMETERS", layout()) {}

    static AudioProcessorValueTreeState::ParameterLayout layout() {
        std::vector<std::unique_ptr<RangedAudioParameter>> v;
        for (const char* id : {"gainPitch", "gainVoc"}) v.push_back(std::make_unique<AudioParameterFloat>(id, id, -60.0f, 6.0f, 0.0f));
        for (const char* id : {"gainVoice", "gainSynth"}) v.push_back(std::make_unique<AudioParameterFloat>(id, id, -60.0f, 6.0f, -60.0f));
        v.push_back(std::make_unique<AudioParameterInt>("lpcVoice", "lpcVoice", 2, 100, 40));
        v.push_back(std::make_unique<AudioParameterInt>("lpcPitch", "lpcPitch", 2, 100, 15));
        v.push_back(std::make_unique<AudioParameterInt>("lpcSynth", "lpcSynth", 2, 30, 5));
        v.push_back(std::make_unique<AudioParameterChoice>(
            "keyPitch", "keyPitch", StringArray("A", "A#", "B", "C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "Chrom"), 12));
        v.push_back(std::make_unique<AudioParameterBool>("pitchBool", "pitchBool", true));
        v.push_back(std::make_unique<AudioParameterBool>("vocBool", "vocBool", true));
        return {v.begin(), v.end()};
    }

    void prepareToPlay(double fs, int B) override { dsp.prepareToPlay(*this, fs, B); }
    void processBlock(AudioBuffer<float>& b, MidiBuffer&) override { dsp.processBlock(*this, b); }

    void releaseResources() override {}
    AudioProcessorEditor* createEditor() override { return nullptr; }
    bool hasEditor() const override { return false; }
    const String getName() const override { return "B200Vocoder"; }
    bool acceptsMidi() const override { return false; }
    bool producesMidi() const override { return false; }
    bool isMidiEffect() const override { return false; }
    double getTailLengthSeconds() const override { return 0.0; }
    int getNumPrograms() override { return 1; }
    int getCurrentProgram() override { return 0; }
    void setCurrentProgram(int) override {}
    const String getProgramName(int) override { return {}; }
    void changeProgramName(int, const String&) override {}
    void getStateInformation(MemoryBlock&) override {}
    void setStateInformation(const void*, int) override {}

    AudioProcessorValueTreeState treeState;
    vpb200::JuceDsp dsp;
};

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: b200_plugin in.f32 out.f32 fs B nBlocks [id=value ...] [@block id=value ...]\n"); return 2; }
    const double fs = std::atof(argv[3]);
    const int B = std::atoi(argv[4]), nBlocks = std::atoi(argv[5]);
    const size_t n = (size_t)B * (size_t)nBlocks;
    std::vector<float> in(3 * n), out(2 * n);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(in.data(), sizeof(float), in.size(), f) != in.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::fclose(f);
    try {
        B200VocoderProcessor proc;
        std::multimap<int, std::pair<std::string, float>> sched;
        int at = -1;
        for (int i = 6; i < argc; ++i) {
            std::string a = argv[i];
            if (a[0] == '@') { at = std::atoi(a.c_str() + 1); continue; }
            const size_t eq = a.find('=');
            if (eq == std::string::npos) { std::fprintf(stderr, "bad argument %s\n", a.c_str()); return 2; }
            const std::string id = a.substr(0, eq);
            const float val = (float)std::atof(a.c_str() + eq + 1);
            if (at < 0) proc.treeState.getRawParameterValue(id.c_str())->store(val);
            else sched.insert({at, {id, val}});
        }
        proc.prepareToPlay(fs, B);
        AudioBuffer<float> buf(3, B);
        MidiBuffer midi;
        for (int b = 0; b < nBlocks; ++b) {
            auto r = sched.equal_range(b);
            for (auto it = r.first; it != r.second; ++it) proc.treeState.getRawParameterValue(it->second.first.c_str())->store(it->second.second);
            for (int c = 0; c < 3; ++c) std::memcpy(buf.getWritePointer(c), in.data() + (size_t)c * n + (size_t)b * B, sizeof(float) * B);
            proc.processBlock(buf, midi);
            for (int c = 0; c < 2; ++c) std::memcpy(out.data() + (size_t)c * n + (size_t)b * B, buf.getReadPointer(c), sizeof(float) * B);
        }
        std::printf("{\"latency_samples\": %d, \"blocks\": %d}\n", proc.getLatencySamples(), nBlocks);
    } catch (const vpb200::Error& e) {
        std::fprintf(stderr, "%s (code %d; there is no CPU fallback)\n", e.what(), e.code);
        return 1;
    }
    f = std::fopen(argv[2], "wb");
    if (!f || std::fwrite(out.data(), sizeof(float), out.size(), f) != out.size()) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 2; }
    std::fclose(f);
    return 0;
}
