// Stand-alone synthetic input generator for bench.py's reference arm: the same definitions as the engine's vp_synth_host
// (vocoderproject_b200/csrc/vp_synth.h + vp_synth_host.hpp), built without CUDA into tools/_build/libvp_inputgen.so, so that
// `bench.py --impl reference` times the reference's CPU path without loading the product library at all.
//   g++ -O2 -ffp-contract=off -fPIC -shared -std=c++17 -o tools/_build/libvp_inputgen.so tools/inputgen.cpp
#include "../vocoderproject_b200/csrc/vp_synth_host.hpp"

extern "C" int vpgen_synth_host(double fs, int flavour, int first, int S, size_t nSamples, size_t stride, float* voice,
                                float* synthL, float* synthR) {
    return vps_fill_host(fs, flavour, first, S, nSamples, stride, voice, synthL, synthR) ? 0 : -1;
}
