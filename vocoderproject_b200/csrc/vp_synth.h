// Synthetic input generator (SURVEY.md section 8(d) "Synthetic inputs"):
// voice = glottal pulse train with glide + vibrato through a formant
// resonator cascade plus aspiration noise; side-chain = sawtooth chord.
//
// One definition, compiled for the host (plain C++) and for the device
// (CUDA): every floating-point operation goes through VPS_MUL/VPS_ADD/VPS_FMA
// so that host and device round identically (no implicit FMA contraction), the
// oscillators are integer phase accumulators, and there are no libm calls in
// the per-sample path. The per-stream parameter table is built on the host
// (vp_synth_make_streams, uses libm) and consumed unchanged by both.
#ifndef VP_SYNTH_H
#define VP_SYNTH_H
#include <stdint.h>

#if defined(__CUDACC__)
#define VPS_HD __host__ __device__ __forceinline__
#else
#define VPS_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define VPS_MUL(a, b) __fmul_rn((a), (b))
#define VPS_ADD(a, b) __fadd_rn((a), (b))
#else
// host side is compiled with -ffp-contract=off (g++) / -fmad=false is not
// needed because nvcc never contracts host code.
#define VPS_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define VPS_ADD(a, b) ((float)((float)(a) + (float)(b)))
#endif

#define VPS_NFORMANTS 4
#define VPS_NSAW 4

typedef struct vp_synth_stream {
    uint32_t seed;
    float f0inc;       // base f0 in cycles/sample
    uint32_t glideInc; // LFO phase increments (2^32 = one cycle)
    uint32_t vibInc;
    uint32_t glidePh0, vibPh0;
    float glideDepth;  // in octaves (semitones/12)
    float vibDepth;
    float tilt;        // one-pole glottal tilt coefficient
    float b0[VPS_NFORMANTS], a1[VPS_NFORMANTS], a2[VPS_NFORMANTS];
    float gain;        // output gain after the cascade
    float noiseAmp;    // aspiration noise amplitude (linear)
    uint32_t sawInc[VPS_NSAW + 1]; // side-chain sawtooth increments; [VPS_NSAW] only on R
    uint32_t sawPh0[VPS_NSAW + 1];
    float sawAmp;
    // optional silent stretch [muteStart, muteEnd) in samples (gate fixtures)
    int64_t muteStart, muteEnd;
} vp_synth_stream;

typedef struct vp_synth_state {
    uint32_t glidePh, vibPh;
    float ph, carry, tiltY;
    float y1[VPS_NFORMANTS], y2[VPS_NFORMANTS];
    uint32_t sawPh[VPS_NSAW + 1];
} vp_synth_state;

VPS_HD uint32_t vps_hash(uint32_t seed, uint64_t i) {
    uint64_t z = (uint64_t)seed * 0x9E3779B97F4A7C15ull + i * 0xBF58476D1CE4E5B9ull + 0x94D049BB133111EBull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}

// smooth LFO in [-1, 1] from a 32-bit phase: triangle shaped by t(1.5 - 0.5 t^2)
VPS_HD float vps_lfo(uint32_t ph) {
    int32_t s = (int32_t)ph;                 // [-2^31, 2^31)
    int32_t a = s < 0 ? -(s + 1) : s;        // [0, 2^31)
    float t = VPS_ADD(VPS_MUL((float)a, 9.313225746154785e-10f), -1.0f); // a/2^30 - 1 in [-1, 1)
    float t2 = VPS_MUL(t, t);
    return VPS_MUL(t, VPS_ADD(1.5f, VPS_MUL(-0.5f, t2)));
}

// 2^x for |x| <= 0.5 (degree-6 Taylor of exp(x ln 2))
VPS_HD float vps_exp2(float x) {
    float y = VPS_MUL(x, 0.6931471805599453f);
    float p = 1.0f / 720.0f;
    p = VPS_ADD(VPS_MUL(p, y), 1.0f / 120.0f);
    p = VPS_ADD(VPS_MUL(p, y), 1.0f / 24.0f);
    p = VPS_ADD(VPS_MUL(p, y), 1.0f / 6.0f);
    p = VPS_ADD(VPS_MUL(p, y), 0.5f);
    p = VPS_ADD(VPS_MUL(p, y), 1.0f);
    p = VPS_ADD(VPS_MUL(p, y), 1.0f);
    return p;
}

VPS_HD void vps_init(const vp_synth_stream* s, vp_synth_state* st) {
    st->glidePh = s->glidePh0; st->vibPh = s->vibPh0;
    st->ph = 0.0f; st->carry = 0.0f; st->tiltY = 0.0f;
    for (int k = 0; k < VPS_NFORMANTS; ++k) { st->y1[k] = 0.0f; st->y2[k] = 0.0f; }
    for (int k = 0; k <= VPS_NSAW; ++k) st->sawPh[k] = s->sawPh0[k];
}

// One sample of (voice, synthL, synthR) at sample index i.
VPS_HD void vps_step(const vp_synth_stream* s, vp_synth_state* st, int64_t i, float* voice, float* sl, float* sr) {
    // --- f0 trajectory
    float oct = VPS_ADD(VPS_MUL(s->glideDepth, vps_lfo(st->glidePh)), VPS_MUL(s->vibDepth, vps_lfo(st->vibPh)));
    st->glidePh += s->glideInc; st->vibPh += s->vibInc;
    float inc = VPS_MUL(s->f0inc, vps_exp2(oct));
    // --- impulse train with linear-interpolated sub-sample placement
    float x = st->carry;
    st->carry = 0.0f;
    st->ph = VPS_ADD(st->ph, inc);
    if (st->ph >= 1.0f) {
        st->ph = VPS_ADD(st->ph, -1.0f);
        float frac = st->ph / inc;  // IEEE division, identical on host and device
        if (frac > 1.0f) frac = 1.0f;
        x = VPS_ADD(x, frac);
        st->carry = VPS_ADD(1.0f, -frac);
    }
    // --- glottal tilt (one pole) and formant cascade (two-pole resonators)
    st->tiltY = VPS_ADD(x, VPS_MUL(s->tilt, st->tiltY));
    float v = st->tiltY;
    for (int k = 0; k < VPS_NFORMANTS; ++k) {
        float y = VPS_ADD(VPS_ADD(VPS_MUL(s->b0[k], v), VPS_MUL(-s->a1[k], st->y1[k])), VPS_MUL(-s->a2[k], st->y2[k]));
        st->y2[k] = st->y1[k]; st->y1[k] = y;
        v = y;
    }
    uint32_t h = vps_hash(s->seed, (uint64_t)i);
    float nz = VPS_MUL((float)(int32_t)h, 4.656612873077393e-10f);  // [-1, 1)
    v = VPS_ADD(VPS_MUL(s->gain, v), VPS_MUL(s->noiseAmp, nz));
    if (i >= s->muteStart && i < s->muteEnd) v = 0.0f;
    *voice = v;
    // --- side-chain: naive sawtooth chord
    float l = 0.0f;
    for (int k = 0; k < VPS_NSAW; ++k) {
        l = VPS_ADD(l, VPS_MUL((float)(int32_t)st->sawPh[k], 4.656612873077393e-10f));
        st->sawPh[k] += s->sawInc[k];
    }
    float r = VPS_ADD(l, VPS_MUL((float)(int32_t)st->sawPh[VPS_NSAW], 4.656612873077393e-10f));
    st->sawPh[VPS_NSAW] += s->sawInc[VPS_NSAW];
    *sl = VPS_MUL(s->sawAmp, l);
    *sr = VPS_MUL(s->sawAmp, r);
}

#endif  // VP_SYNTH_H
