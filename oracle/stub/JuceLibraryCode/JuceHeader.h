// TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
//
// Minimal stand-in for JUCE 5.4.7's JuceHeader.h so that the reference's own
// Source/*.cpp (under /root/reference, never copied here) compile unmodified
// and VocoderAudioProcessor::processBlock runs headlessly. Only the JUCE
// surface the DSP path touches carries arithmetic; it restates JUCE's
// documented behaviour. JUCE is not vendored by the
// reference (Vocoder.jucer:3 pins 5.4.7, .gitignore:1-2), so this file IS the
// pin for that boundary ("parity unpinned at the JUCE boundary").
#pragma once
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <initializer_list>

#define JUCE_CALLTYPE
#define JucePlugin_Name "Vocoder"
#define JUCE_DECLARE_NON_COPYABLE_WITH_LEAK_DETECTOR(C) \
    C(const C&) = delete;                               \
    C& operator=(const C&) = delete;

namespace juce {

template <typename... T> void ignoreUnused(const T&...) {}

template <typename T> struct MathConstants {
    static constexpr T pi = static_cast<T>(3.141592653589793238L);
    static constexpr T twoPi = static_cast<T>(2 * 3.141592653589793238L);
};

class String {
public:
    String() {}
    String(const char* s) : s_(s) {}
    String(const std::string& s) : s_(s) {}
    const std::string& str() const { return s_; }
private:
    std::string s_;
};

class StringArray {
public:
    template <typename... A> StringArray(A... a) : v_{String(a)...} {}
    int size() const { return (int)v_.size(); }
private:
    std::vector<String> v_;
};

class MemoryBlock {};
class MidiBuffer {};
class AudioProcessorEditor {};

struct ScopedNoDenormals {
    // JUCE sets FTZ/DAZ for the scope. The synthetic fixtures never produce
    // denormals in double, so this is a no-op here.
    ScopedNoDenormals() {}
};

// --- AudioBuffer<T> -------------------------------------------------------
template <typename T> class AudioBuffer {
public:
    AudioBuffer() {}
    AudioBuffer(int nc, int ns) { setSize(nc, ns); }
    // Non-owning view (what AudioProcessor::getBusBuffer hands out).
    AudioBuffer(T* const* chans, int nc, int start, int ns) : nc_(nc), ns_(ns), isClear_(false) {
        ptr_.resize(nc);
        for (int c = 0; c < nc; ++c) ptr_[c] = chans[c] + start;
    }
    void setSize(int nc, int ns) {
        nc_ = nc; ns_ = ns;
        store_.assign((size_t)nc * (size_t)ns, T(0));
        ptr_.resize(nc);
        for (int c = 0; c < nc; ++c) ptr_[c] = store_.data() + (size_t)c * ns;
        isClear_ = false;
    }
    int getNumChannels() const { return nc_; }
    int getNumSamples() const { return ns_; }
    const T* getReadPointer(int ch) const { return ptr_[ch]; }
    const T* getReadPointer(int ch, int start) const { return ptr_[ch] + start; }
    T* getWritePointer(int ch) { isClear_ = false; return ptr_[ch]; }
    T* getWritePointer(int ch, int start) { isClear_ = false; return ptr_[ch] + start; }
    T* const* getArrayOfWritePointers() { isClear_ = false; return ptr_.data(); }
    T getSample(int ch, int i) const { return ptr_[ch][i]; }
    void setSample(int ch, int i, T v) { ptr_[ch][i] = v; isClear_ = false; }
    void addSample(int ch, int i, T v) { ptr_[ch][i] += v; isClear_ = false; }
    void clear() {
        if (!isClear_) {
            for (int c = 0; c < nc_; ++c) std::fill(ptr_[c], ptr_[c] + ns_, T(0));
            isClear_ = true;
        }
    }
    void clear(int ch, int start, int n) {
        if (!isClear_) std::fill(ptr_[ch] + start, ptr_[ch] + start + n, T(0));
    }
    template <typename U>
    void addFrom(int dCh, int dStart, const AudioBuffer<U>& src, int sCh, int sStart, int n, T gain = T(1)) {
        if (gain == T(0) || n <= 0 || src.hasBeenCleared()) return;
        T* d = ptr_[dCh] + dStart;
        const U* s = src.getReadPointer(sCh, sStart);
        if (isClear_) {
            isClear_ = false;
            for (int i = 0; i < n; ++i) d[i] = s[i] * gain;
        } else {
            for (int i = 0; i < n; ++i) d[i] += s[i] * gain;
        }
    }
    T getRMSLevel(int ch, int start, int n) const {
        if (n <= 0 || ch < 0 || ch >= nc_ || isClear_) return T(0);
        const T* d = ptr_[ch] + start;
        double sum = 0.0;
        for (int i = 0; i < n; ++i) { const double s = (double)d[i]; sum += s * s; }
        return static_cast<T>(std::sqrt(sum / n));
    }
    bool hasBeenCleared() const { return isClear_; }
private:
    int nc_ = 0, ns_ = 0;
    bool isClear_ = false;
    std::vector<T> store_;
    std::vector<T*> ptr_;
};

// --- Decibels ---------------------------------------------------------------
struct Decibels {
    template <typename T> static T decibelsToGain(T dB, T minusInf = T(-100)) {
        return dB > minusInf ? std::pow(T(10.0), dB * T(0.05)) : T();
    }
    template <typename T> static T gainToDecibels(T g, T minusInf = T(-100)) {
        return g > T() ? std::max(minusInf, static_cast<T>(std::log10(g)) * T(20.0)) : minusInf;
    }
};

// --- dsp::WindowingFunction ---------------------------------------------------
namespace dsp {
template <typename F> struct WindowingFunction {
    enum WindowingMethod { rectangular = 0, triangular, hann, hamming };
    static void fillWindowingTables(F* w, size_t size, WindowingMethod m, bool normalise = true) {
        (void)normalise;
        if (m == hann) {
            for (size_t i = 0; i < size; ++i) {
                F c2 = std::cos(static_cast<F>(2 * i) * MathConstants<F>::pi / static_cast<F>(size - 1));
                w[i] = static_cast<F>(0.5 - 0.5 * c2);
            }
        } else {
            for (size_t i = 0; i < size; ++i) w[i] = F(1);
        }
    }
};
}  // namespace dsp

// --- parameters -------------------------------------------------------------
template <typename T> struct NormalisableRange { T start{}, end{}; };

class RangedAudioParameter {
public:
    RangedAudioParameter(const String& id, float lo, float hi, float def) : id_(id.str()), lo_(lo), hi_(hi), v_(def) {}
    virtual ~RangedAudioParameter() {}
    std::string id_;
    float lo_, hi_;
    std::atomic<float> v_;
};
struct AudioParameterFloat : RangedAudioParameter {
    AudioParameterFloat(const String& id, const String&, float lo, float hi, float def) : RangedAudioParameter(id, lo, hi, def) {}
};
struct AudioParameterInt : RangedAudioParameter {
    AudioParameterInt(const String& id, const String&, int lo, int hi, int def) : RangedAudioParameter(id, (float)lo, (float)hi, (float)def) {}
};
struct AudioParameterChoice : RangedAudioParameter {
    AudioParameterChoice(const String& id, const String&, const StringArray& c, int def) : RangedAudioParameter(id, 0.f, (float)(c.size() - 1), (float)def) {}
};
struct AudioParameterBool : RangedAudioParameter {
    AudioParameterBool(const String& id, const String&, bool def) : RangedAudioParameter(id, 0.f, 1.f, def ? 1.f : 0.f) {}
};

class AudioProcessor;
class AudioProcessorValueTreeState {
public:
    struct ParameterLayout {
        ParameterLayout() {}
        template <typename It> ParameterLayout(It b, It e) {
            for (; b != e; ++b) params.push_back(std::move(*b));
        }
        std::vector<std::unique_ptr<RangedAudioParameter>> params;
    };
    AudioProcessorValueTreeState(AudioProcessor&, void*, const String&, ParameterLayout layout) {
        for (auto& p : layout.params) { std::string id = p->id_; map_[id] = std::move(p); }
    }
    std::atomic<float>* getRawParameterValue(const String& id) const { return &map_.at(id.str())->v_; }
    NormalisableRange<float> getParameterRange(const String& id) const {
        const auto& p = map_.at(id.str());
        return {p->lo_, p->hi_};
    }
private:
    std::map<std::string, std::unique_ptr<RangedAudioParameter>> map_;
};

// --- AudioProcessor ---------------------------------------------------------
struct AudioChannelSet {
    int n = 0;
    static AudioChannelSet mono() { return {1}; }
    static AudioChannelSet stereo() { return {2}; }
    static AudioChannelSet disabled() { return {0}; }
    bool isDisabled() const { return n == 0; }
    bool operator==(const AudioChannelSet& o) const { return n == o.n; }
};

class AudioProcessor {
public:
    struct BusesProperties {
        std::vector<int> in, out;
        BusesProperties withInput(const String&, const AudioChannelSet& s, bool = true) const {
            BusesProperties r = *this; r.in.push_back(s.n); return r;
        }
        BusesProperties withOutput(const String&, const AudioChannelSet& s, bool = true) const {
            BusesProperties r = *this; r.out.push_back(s.n); return r;
        }
    };
    struct BusesLayout {
        std::vector<AudioChannelSet> inputBuses, outputBuses;
        AudioChannelSet getMainInputChannelSet() const { return inputBuses.empty() ? AudioChannelSet() : inputBuses[0]; }
        AudioChannelSet getMainOutputChannelSet() const { return outputBuses.empty() ? AudioChannelSet() : outputBuses[0]; }
        AudioChannelSet getChannelSet(bool isInput, int bus) const {
            const auto& v = isInput ? inputBuses : outputBuses;
            return bus < (int)v.size() ? v[bus] : AudioChannelSet();
        }
    };
    AudioProcessor() {}
    explicit AudioProcessor(const BusesProperties& p) : props_(p) {}
    virtual ~AudioProcessor() {}
    virtual void prepareToPlay(double, int) = 0;
    virtual void releaseResources() = 0;
    virtual bool isBusesLayoutSupported(const BusesLayout&) const { return true; }
    virtual void processBlock(AudioBuffer<float>&, MidiBuffer&) = 0;
    virtual AudioProcessorEditor* createEditor() = 0;
    virtual bool hasEditor() const = 0;
    virtual const String getName() const = 0;
    virtual bool acceptsMidi() const = 0;
    virtual bool producesMidi() const = 0;
    virtual bool isMidiEffect() const = 0;
    virtual double getTailLengthSeconds() const = 0;
    virtual int getNumPrograms() = 0;
    virtual int getCurrentProgram() = 0;
    virtual void setCurrentProgram(int) = 0;
    virtual const String getProgramName(int) = 0;
    virtual void changeProgramName(int, const String&) = 0;
    virtual void getStateInformation(MemoryBlock&) = 0;
    virtual void setStateInformation(const void*, int) = 0;

    int getTotalNumInputChannels() const { int n = 0; for (int c : props_.in) n += c; return n; }
    int getTotalNumOutputChannels() const { int n = 0; for (int c : props_.out) n += c; return n; }
    void setLatencySamples(int l) { latency_ = l; }
    int getLatencySamples() const { return latency_; }
    AudioProcessorEditor* getActiveEditor() const { return nullptr; }
    // Bus b of the in-place processing buffer: channels [offset, offset + n).
    template <typename T> AudioBuffer<T> getBusBuffer(AudioBuffer<T>& buf, bool isInput, int bus) const {
        const auto& v = isInput ? props_.in : props_.out;
        int off = 0;
        for (int b = 0; b < bus; ++b) off += v[b];
        return AudioBuffer<T>(buf.getArrayOfWritePointers() + off, v[bus], 0, buf.getNumSamples());
    }
private:
    BusesProperties props_;
    int latency_ = 0;
};

}  // namespace juce

namespace foleys {
struct MagicProcessorState {
    MagicProcessorState(juce::AudioProcessor&, juce::AudioProcessorValueTreeState&) {}
    void getStateInformation(juce::MemoryBlock&) {}
    void setStateInformation(const void*, int, juce::AudioProcessorEditor*) {}
};
struct MagicPluginEditor : juce::AudioProcessorEditor {
    MagicPluginEditor(MagicProcessorState&, const char*, int) {}
};
}  // namespace foleys

namespace BinaryData {
static const char* const vocodergui_final = "";
static const int vocodergui_finalSize = 0;
}  // namespace BinaryData

using namespace juce;
