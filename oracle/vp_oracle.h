/* TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
 *
 * CPU restatement, in plain C and in double precision, of the DSP core that
 * VocoderAudioProcessor::processBlock drives in DamRsn/VocoderProject:
 * MyBuffer ring buffering, LPC, VocoderProcess, PitchProcess, Notes.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference). It is written on the flat "delayed timeline" of SURVEY.md
 * App. A.1 instead of ring buffers, but executes block by block and chunk by
 * chunk in the reference's own order, so its output is pinned bit-for-bit
 * against oracle/_ref (the reference's C++ compiled in place) by
 * tests/test_oracle.py and against tests/golden/ (fixtures produced by
 * oracle/_ref in the authoring container).
 */
#ifndef VP_ORACLE_H
#define VP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as include/vp_engine.h's vp_params and as
 * oracle/ref_harness.cpp's vpref_params. Gains in dB (float, like the
 * plug-in's std::atomic<float> parameters, PluginProcessor.cpp:37-73). */
typedef struct {
    float gainPitch, gainVoice, gainSynth, gainVoc;
    int lpcVoice, lpcPitch, lpcSynth;
    int keyPitch; /* Notes::key, Notes.h:26; 12 = chromatic */
    int pitchBool, vocBool;
} vpo_params;

#define VPO_MAX_MARKS 64
typedef struct {
    int frame, startSample, block, gated;
    int period, periodNew, prevVoicedPeriod;
    int note; /* index into the note table, -1 if unvoiced / no marks */
    int nAn, nSt;
    int anStale; /* storage slot anMarks[size] (PitchProcess.cpp:818 reads it) */
    int anMarks[VPO_MAX_MARKS];
    int stMarks[VPO_MAX_MARKS];
    double pitch, closestFreq, beta;
} vpo_pitch_frame;

typedef struct {
    int frame, startSample, block, gated;
    double EeVoice, EeSynth, g;
} vpo_voc_frame;

typedef struct {
    int hopV, wlenV, hopP, frameLenP, chunk, tauMax, latency, keep, inSize, outSize, anCap, nFreq;
} vpo_sizes;

/* Flags the restatement raises where the reference has undefined behaviour
 * (SURVEY.md App. B). A set bit means the oracle's output is NOT authoritative
 * for that run. */
#define VPO_UB_CLOSEST_PREV 1 /* U2  PitchProcess.cpp:812 -> :694 */
#define VPO_UB_YIN_END 2      /* U3  PitchProcess.cpp:435 */
#define VPO_UB_PREV_EMPTY 4   /* U4  PitchProcess.cpp:487 */
#define VPO_UB_INTERP 8       /* U5  PitchProcess.cpp:856 */
#define VPO_UB_CAPACITY 16    /* mark vector would have re-allocated */
#define VPO_UB_ASSERT 32      /* a reference assert(false) site was reached */

/* Defined-behaviour mode (process-global, not thread-safe: set it before the runs it should apply to). 0 = the reference as
 * it runs (default), 1 = every undefined-behaviour site takes its bounds-correct reading (the engine's VP_MODE_DEFINED).
 * vpo_defined_deviations: how often, since vpo_set_defined, a defined-mode choice differed from the mode-0 one. */
/* Vocoder window type of the runs that follow (process-global): 0 = "sine" (what prepareToPlay passes), 1 = "hann"
 * (VocoderProcess.cpp:116-124: rectangular analysis, Hann synthesis). */
void vpo_set_window(int hann);
void vpo_set_defined(int on);
long vpo_defined_deviations(void);

void vpo_default_params(vpo_params* p);
void vpo_sizes_for(double fs, int B, int key, vpo_sizes* s);

/* Equivalent of: construct the plug-in, set parameters, prepareToPlay(fs, B),
 * then nBlocks processBlock calls (PluginProcessor.cpp:144-184, :203-234).
 * voice/synthL/synthR: nBlocks*B floats (synthR NULL -> synthL).
 * outL/outR: nBlocks*B floats (outR may be NULL). Logs may be NULL.
 * Returns 0 on success, <0 on bad arguments; *ubFlags (may be NULL) receives
 * the OR of VPO_UB_* bits. */
int vpo_process(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                const vpo_params* params, float* outL, float* outR, vpo_sizes* sizes,
                vpo_pitch_frame* plog, int plogCap, int* nP, vpo_voc_frame* vlog, int vlogCap, int* nV,
                int* ubFlags);

/* Same with parameter automation: sched[i] replaces the parameters before block schedBlock[i]. */
int vpo_process_sched(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                      const vpo_params* params, const vpo_params* sched, const int* schedBlock, int nSched,
                      float* outL, float* outR, vpo_sizes* sizes,
                      vpo_pitch_frame* plog, int plogCap, int* nP, vpo_voc_frame* vlog, int vlogCap, int* nV,
                      int* ubFlags);

/* Notes.cpp:43-70 table; returns its size, *popped = the popped slot (U6). */
int vpo_notes(int key, double fMin, double fMax, double* freq, int cap, double* popped);

/* CPU "port" baseline: S streams over nThreads pthreads. Layouts as
 * oracle/ref_harness.cpp's vpref_bench. Returns wall seconds. */
double vpo_bench(double fs, int B, int nBlocks, int S, const float* voice, const float* synthL,
                 const float* synthR, const vpo_params* params, int nThreads, float* out);

#ifdef __cplusplus
}
#endif
#endif
