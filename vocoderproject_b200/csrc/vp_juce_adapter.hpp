// vp_juce_adapter.hpp -- the two bodies a JUCE plug-in swaps in for the reference's DSP members.
//
// Include AFTER JuceHeader.h. `Proc` is the plug-in's juce::AudioProcessor subclass; the adapter uses exactly what
// VocoderAudioProcessor's own prepareToPlay / processBlock use (Source/PluginProcessor.cpp:144-184, :203-234):
//   p.treeState.getRawParameterValue(id)->load()   the ten parameters of createParameterLayout (:37-73)
//   p.getBusBuffer(buffer, true, 0 | 1)            voice (mono, bus 0) and side-chain (stereo, bus 1)   (:209-210)
//   p.setLatencySamples(n)                         (:183)
// and the in-place processing buffer: channels 0..1 are overwritten with the output (MyBuffer.cpp:113-133).
//
//   class VocoderAudioProcessor : public AudioProcessor {
//       ...
//       AudioProcessorValueTreeState treeState;
//       vpb200::JuceDsp dsp;                                        // instead of myBuffer / vocoderProcess / pitchProcess
//       void prepareToPlay (double fs, int B) override      { dsp.prepareToPlay (*this, fs, B); }
//       void processBlock (AudioBuffer<float>& b, MidiBuffer&) override { dsp.processBlock (*this, b); }
//   };
//
// One plug-in instance = one stream (nStreams = 1); a host that runs many instances on one GPU uses
// vpb200::VocoderBatchProcessor (vp_facade.hpp) directly.
#pragma once
#include <vector>

#include "vp_facade.hpp"

namespace vpb200 {

class JuceDsp {
public:
    explicit JuceDsp(int device = 0) : dsp(device) {}

    // the ten atomics -> vp_params, at the call (the reference's DSP classes load them at their read sites; the
    // engine applies each one at the same granularity, see vp_engine_set_params in include/vp_engine.h)
    template <class Proc>
    static vp_params pullParams(const Proc& p) {
        auto v = [&p](const char* id) { return p.treeState.getRawParameterValue(id)->load(); };
        vp_params q;
        q.gainPitch = v("gainPitch"); q.gainVoice = v("gainVoice"); q.gainSynth = v("gainSynth"); q.gainVoc = v("gainVoc");
        q.lpcVoice = (int)v("lpcVoice"); q.lpcPitch = (int)v("lpcPitch"); q.lpcSynth = (int)v("lpcSynth");
        q.keyPitch = (int)v("keyPitch");
        q.pitchBool = v("pitchBool") != 0.0f ? 1 : 0;
        q.vocBool = v("vocBool") != 0.0f ? 1 : 0;
        return q;
    }

    template <class Proc>
    void prepareToPlay(Proc& p, double sampleRate, int samplesPerBlock) {
        dsp.params = pullParams(p);
        fs = sampleRate; B = samplesPerBlock;
        // same size derivation, :160-176; rows sized for the ends of the lpcVoice / lpcSynth ranges (:53-59), like the
        // reference's orderMax vectors, so that the order knobs can move while audio runs
        dsp.prepareToPlay(sampleRate, samplesPerBlock, /*nStreams*/ 1, /*maxBlocksPerCall*/ 1, 100, 30);
        p.setLatencySamples(dsp.getLatencySamples());                      // :183
        outL.assign((size_t)samplesPerBlock, 0.0f);
        outR.assign((size_t)samplesPerBlock, 0.0f);
        prepared = true;
    }

    template <class Proc, class Buffer>
    void processBlock(Proc& p, Buffer& buffer) {
        auto voice = p.getBusBuffer(buffer, true, 0);   // :209
        auto synth = p.getBusBuffer(buffer, true, 1);   // :210
        const int n = buffer.getNumSamples();
        if (!prepared || n != B) {                      // a host that changes its block size calls prepareToPlay first
            B = n;
            prepareToPlay(p, fs, n);
        }
        // every parameter is pushed per block and takes effect where the reference reads it (orders per vocoder frame,
        // enables per block, ...): no restart, no state loss when a knob or a bypass button moves
        const vp_params q = pullParams(p);
        dsp.params = q;
        dsp.processBlock(voice.getReadPointer(0), synth.getReadPointer(0), synth.getReadPointer(1), outL.data(), outR.data(),
                         (size_t)n);                    // :212-232 in one engine call
        // fillOutputBuffer (MyBuffer.cpp:113-133): output channels 0 and 1 of the in-place buffer
        float* o0 = buffer.getWritePointer(0);
        for (int i = 0; i < n; ++i) o0[i] = outL[(size_t)i];
        if (buffer.getNumChannels() > 1) {
            float* o1 = buffer.getWritePointer(1);
            for (int i = 0; i < n; ++i) o1[i] = outR[(size_t)i];
        }
    }

    VocoderBatchProcessor dsp;

private:
    std::vector<float> outL, outR;
    double fs = 44100.0;
    int B = 0;
    bool prepared = false;
};

}  // namespace vpb200
