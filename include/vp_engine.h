/* vp_engine.h -- C ABI of the B200 batch engine for VocoderProject's DSP core.
 *
 * Drop-in boundary for the path VocoderAudioProcessor::processBlock drives
 * (reference: Source/PluginProcessor.cpp:203-234) -- MyBuffer ring buffering,
 * LPC, VocoderProcess, PitchProcess, Notes -- batched over independent
 * streams. Plain pointers and sizes only; no C++/CUDA/torch types. The C++
 * facade with the reference's class names lives in
 * vocoderproject_b200/csrc/vp_facade.hpp and is implemented on top of this
 * ABI; INTEGRATION.md shows the PluginProcessor-side binding.
 *
 * Semantics. One "stream" is one plug-in instance (1 mono voice + stereo
 * side-chain in, stereo out). vp_engine_prepare is prepareToPlay(sampleRate,
 * samplesPerBlock) for every stream; each vp_engine_process* call is, per
 * stream, nBlocks further consecutive processBlock calls. Results do
 * depend (sparsely) on samplesPerBlock exactly as in the reference (whole-ring
 * RMS gate, MyBuffer.cpp:258-261; PSOLA look-ahead test, PitchProcess.cpp:
 * 799-803), hence the "virtual block size" argument of vp_engine_prepare.
 *
 * All functions return VP_OK (0) or a negative VP_E_* code; none throws, none
 * aborts. There is no CPU fallback: without a usable CUDA device every
 * compute entry point returns VP_E_CUDA.
 */
#ifndef VP_ENGINE_H
#define VP_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_OK 0
#define VP_E_ARG (-1)     /* bad argument */
#define VP_E_STATE (-2)   /* call order (e.g. process before prepare) */
#define VP_E_CUDA (-3)    /* CUDA runtime error / no device; see vp_last_error */
#define VP_E_NOMEM (-4)   /* device or host allocation failed */
#define VP_E_RANGE (-5)   /* parameter outside the plug-in's range */

/* The ten plug-in parameters (Source/PluginProcessor.cpp:37-73), pushed by
 * the caller instead of pulled through audioProcPtr->treeState. Gains in dB
 * as float, exactly what the plug-in's std::atomic<float> holds. */
typedef struct vp_params {
    float gainPitch; /* [-60, 6]  default 0    PitchProcess.cpp:336 */
    float gainVoice; /* [-60, 6]  default -60  PluginProcessor.cpp:226 (dry voice, off <= -59) */
    float gainSynth; /* [-60, 6]  default -60  PluginProcessor.cpp:229 (dry synth, off <= -59) */
    float gainVoc;   /* [-60, 6]  default 0    VocoderProcess.cpp:291 */
    int lpcVoice;    /* [2, 100]  default 40   VocoderProcess.cpp:139 */
    int lpcPitch;    /* [2, 100]  default 15   PitchProcess.cpp:70 (read in prepare only) */
    int lpcSynth;    /* [2, 30]   default 5    VocoderProcess.cpp:155 */
    int keyPitch;    /* Notes::key 0..12 (A..G#, 12 = chromatic), default 12; Notes.h:26 */
    int pitchBool;   /* default 1 */
    int vocBool;     /* default 1 */
} vp_params;

/* Sizes prepareToPlay derives from the sample rate
 * (PluginProcessor.cpp:160-176, PitchProcess.cpp:100-107, MyBuffer.cpp:46-48). */
typedef struct vp_sizes {
    int hopV, wlenV;           /* vocoder hop / frame length */
    int hopP, frameLenP, chunk;/* pitch hop / frame length / chunk */
    int tauMin, tauMax;        /* YIN lag search range [tauMin, tauMax) */
    int latency, keep;         /* plug-in latency, samplesToKeep */
    int inSize, outSize;       /* MyBuffer ring lengths for this block size */
    int anCap;                 /* capacity of the pitch-mark vectors */
    int nFreq;                 /* entries in the note table for keyPitch */
} vp_sizes;

/* Per pitch frame decisions (one record per processChunkStart call,
 * PitchProcess.cpp:203-247). */
#define VP_MAX_MARKS 24
#define VP_NO_MARK (-2147483647 - 1)
#define VP_PF_GATED 1u       /* silence gate fired (PitchProcess.cpp:208-214) */
#define VP_PF_VOICED 2u      /* pitch > 1 */
#define VP_PF_HAS_MARKS 4u   /* anMarks non-empty -> frame is synthesised */
#define VP_PF_NEAR_GATE 16u      /* gate level within eps of -60 dB */
#define VP_PF_NEAR_YIN 32u       /* a deciding YIN comparison within eps (after the FP64 re-check) */
#define VP_PF_UB 64u             /* the reference has undefined behaviour here */
#define VP_PF_YIN_RECHECKED 128u /* FP32 YIN was inconclusive; decided by the FP64 pass */
typedef struct vp_pitch_frame {
    uint32_t flags;
    int32_t period;           /* detected period in samples (0 = unvoiced) */
    int32_t periodPsola;      /* period else prevVoicedPeriod (PitchProcess.cpp:669-672) */
    int32_t periodNew;        /* synthesis mark spacing */
    int32_t note;             /* snapped note index in the key's table, -1 if none */
    int32_t nAn, nSt;
    int32_t anStale;          /* storage slot anMarks[nAn] (PitchProcess.cpp:818) */
    int32_t nAnOv;
    int32_t prevAnLast;       /* last mark of the previous frame that does not overlap this one, in this frame's
                                 coordinates (what PitchProcess.cpp:812 means to fall back to); VP_NO_MARK if none */
    int32_t anMarks[VP_MAX_MARKS];
    int32_t stMarks[VP_MAX_MARKS];
    double beta;              /* closestFreq / pitch, carried over unvoiced frames */
} vp_pitch_frame;

typedef struct vp_engine vp_engine;

/* ---- lifecycle ----------------------------------------------------------- */
void vp_default_params(vp_params* p);
/* Host-only helper: the sizes prepareToPlay would derive. */
int vp_sizes_for(double sampleRate, int samplesPerBlock, int keyPitch, vp_sizes* out);
/* Number of visible CUDA devices (0 when there is none). */
int vp_device_count(void);

int vp_engine_create(vp_engine** e, int device);
void vp_engine_destroy(vp_engine* e);
const char* vp_last_error(const vp_engine* e);

/* Largest lpcVoice / lpcSynth that vp_engine_set_params may be given on a RUNNING stream (the reference re-reads both orders
 * at every vocoder frame, VocoderProcess.cpp:193-194, and sizes its vectors for the ends of the parameter ranges, :50-57).
 * Call before vp_engine_prepare: the coefficient rows of the workspace are sized for it. Without it the maximum is what the
 * parameters say at vp_engine_prepare. 0 = no reservation. */
int vp_engine_reserve_orders(vp_engine* e, int maxLpcVoice, int maxLpcSynth);

/* prepareToPlay(sampleRate, samplesPerBlock) for nStreams independent
 * plug-in instances. maxBlocks bounds nBlocks of later process calls.
 * workspaceBytes: device memory the engine may use for intermediates
 * (0 = default); it decides how many streams are in flight per pass. */
int vp_engine_prepare(vp_engine* e, double sampleRate, int samplesPerBlock, int nStreams, int maxBlocks,
                      size_t workspaceBytes);
/* Replaces the reference's parameter pulls (treeState.getRawParameterValue(id)->load(); VocoderProcess.cpp:193-194,291,
 * PitchProcess.cpp:70,206,336, PluginProcessor.cpp:212-230). Before prepare: anything in the plug-in's ranges. Between
 * process calls of a running stream (automation) every parameter takes effect where the reference reads it: gainVoc per
 * vocoder frame, gainPitch per emitted chunk, gainVoice / gainSynth per block, keyPitch per pitch frame, lpcVoice / lpcSynth
 * per vocoder frame (frames in flight keep the order they started with; VP_E_STATE beyond vp_engine_reserve_orders / the
 * orders at prepare), vocBool / pitchBool per block: a block with vocBool off skips VocoderProcess::process (its frame grid
 * freezes relative to the block, frames in flight still come out), a block with pitchBool off is PitchProcess::silence()
 * (PluginProcessor.cpp:214-221, PitchProcess.cpp:146-158). lpcPitch is read in PitchProcess::prepare only (:70): a change on
 * a running stream is ignored until the next vp_engine_prepare, as in the reference. */
int vp_engine_set_params(vp_engine* e, const vp_params* p);
/* Behaviour at the places where the reference's C++ has undefined behaviour:
 *   VP_MODE_PARITY  (default) what the reference build actually does: the stale vector slot of PitchProcess.cpp:818 and the
 *                   popped table slot of Notes.cpp:99 are modelled, the other sites keep going and the frame is flagged VP_PF_UB;
 *   VP_MODE_DEFINED every such site takes the bounds-correct reading of what the code says it wants:
 *                   :818  the completeness test looks at the mark it is about to return (the last one), not one past the end;
 *                   :812  "last non-overlapping mark of the previous frame" = prevAnMarks[size - nOv - 1] (vp_pitch_frame.prevAnLast),
 *                         the frame's first mark when there is none;
 *                   :435  the descent ends at the last lag;   :487  "previous frame voiced" without previous marks searches the
 *                         frame like a first voiced frame;    :856  the interval search restarts at begin();
 *                   Notes.cpp:99  above the note table the closest note is the last one.
 *                   No frame is flagged VP_PF_UB for these sites. Identical to VP_MODE_PARITY wherever none of them is reached
 *                   with a different outcome. The CPU oracle has the same switch (oracle/vp_oracle.h).
 * May be called between process calls; takes effect with the next call. */
/* Window type of VocoderProcess::prepare (VocoderProcess.cpp:35-71, setWindows :95-135): VP_WINDOW_SINE is what
 * prepareToPlay passes (PluginProcessor.cpp:165); VP_WINDOW_HANN is the class's other branch (:116-124: no analysis window,
 * Hann synthesis window). Before vp_engine_prepare (or after vp_engine_reset, followed by vp_engine_prepare). */
#define VP_WINDOW_SINE 0
#define VP_WINDOW_HANN 1
int vp_engine_set_window(vp_engine* e, int window);

#define VP_MODE_PARITY 0
#define VP_MODE_DEFINED 1
int vp_engine_set_mode(vp_engine* e, int mode);
int vp_engine_get_sizes(const vp_engine* e, vp_sizes* out);
/* Layout the engine chose in vp_engine_prepare: streams processed per pass (nStreams / streamsPerPass passes per call),
 * carried input history per stream (samples) and the device workspace in use. Any pointer may be NULL. */
int vp_engine_get_info(const vp_engine* e, int* streamsPerPass, int* historySamples, size_t* workspaceBytes);

/* Back to the state right after vp_engine_prepare (= prepareToPlay): rings cleared, histories and pitch marks
 * forgotten, block counter 0. Consecutive vp_engine_process_* calls otherwise CONTINUE the streams: calling with
 * nBlocks = a then nBlocks = b gives the output of one call with nBlocks = a + b (the carried per-stream state is what
 * MyBuffer / VocoderProcess / PitchProcess keep between processBlock calls). */
int vp_engine_reset(vp_engine* e);

/* Host-only helper (no CUDA): the frame grids a sequence of process calls produces -- call c processes nBlocks[c] blocks
 * with params[c] in force (only vocBool, pitchBool, lpcVoice, lpcSynth matter). For hosts that want to know how many frame
 * records a call will report, and for the CPU tests of the grid bookkeeping. */
typedef struct vp_call_plan {
    long long firstBlock;        /* blocks processed before this call */
    int offV, nFramesV;          /* first new vocoder frame (samples after the call's start) and how many start in the call */
    int carriedV;                /* vocoder frames of earlier calls that lie on this call's grid */
    int rowOrderV, rowOrderS;    /* width - 1 of the call's coefficient rows (>= lpcVoice / lpcSynth in force) */
    int offP, nFramesP;          /* same for the pitch frames */
    int vocMix, pitchMix;        /* the call's output contains vocoder / pitch-corrector samples */
    int rowsOrphaned, orphansLive; /* vocoder frames left off the grid by a vocBool-off block: moved at this call / emitting in it */
    int carryPosP[2], carryChunksP[2]; /* carried pitch frames: start relative to the call, chunks of theirs that are ever processed */
} vp_call_plan;
int vp_grid_plan(double sampleRate, int samplesPerBlock, int nCalls, const int* nBlocks, const vp_params* params,
                 vp_call_plan* out);

/* ---- processing ----------------------------------------------------------- */
/* Device-resident batch. Layouts (row stride = strideSamples floats):
 *   voice  [nStreams][stride]   synthL, synthR [nStreams][stride]
 *   outL, outR [nStreams][stride]
 * synthR may be NULL: the right side-chain channel then equals the left one (the
 * vocoder analyses channel 0 only, VocoderProcess.cpp:211,218; channel 1 is only
 * heard through the dry side-chain mix, gainSynth > -59 dB). When it is given its
 * ring is carried on every call whatever gainSynth is, as in the reference
 * (MyBuffer.cpp:69-92), so gainSynth can be automated on mid-stream. outR may be
 * NULL (then only L is written; L == R whenever gainSynth <= -59 dB).
 * The outputs must NOT overlap the inputs (no in-place use: the mix reads the
 * delayed inputs while the outputs are being written) -- VP_E_ARG otherwise.
 * Asynchronous on the engine's stream; vp_engine_sync waits.
 * If a process call fails after its first pass has been issued (VP_E_CUDA /
 * VP_E_NOMEM), part of the streams' carried state has advanced: every later
 * process call returns VP_E_STATE until vp_engine_reset or vp_engine_prepare. */
int vp_engine_process_device(vp_engine* e, int nBlocks, const float* voice, const float* synthL,
                             const float* synthR, float* outL, float* outR, size_t strideSamples);

/* Host buffers (pageable or pinned, see vp_host_alloc), same layouts. Streams
 * are moved in slices with H2D / compute / D2H overlapped on three CUDA
 * streams; returns when the outputs are in host memory. */
int vp_engine_process_host(vp_engine* e, int nBlocks, const float* voice, const float* synthL,
                           const float* synthR, float* outL, float* outR, size_t strideSamples);

/* Same call with 16-bit PCM host arrays (int16 [nStreams][stride]): the samples cross the host link as 2 bytes and are
 * converted on the device exactly as the WAV front-end converts on the host (csrc/vp_wav.hpp: x / 32768 in,
 * clamp(rint(y * 32768)) out), so the result is bit-identical to converting on the host and calling
 * vp_engine_process_host, then quantising its output. For hosts whose audio is PCM anyway (files, capture devices): half
 * the link bytes of the float call, and the link is what bounds the end-to-end rate. Not a format of the
 * reference's processBlock (JUCE buffers are float): an extension of the WAV front-end of SURVEY.md 8(f)#3. */
int vp_engine_process_host_pcm16(vp_engine* e, int nBlocks, const int16_t* voice, const int16_t* synthL,
                                 const int16_t* synthR, int16_t* outL, int16_t* outR, size_t strideSamples);

int vp_engine_sync(vp_engine* e);

/* Low-latency streaming, one host block per call (BASELINE config 5). The engine owns pinned host buffers
 * voice / synthL / outL, each float[nStreams][samplesPerBlock]: fill the inputs, call vp_engine_stream_block, read
 * outL when it returns (L == R; gainSynth must be off; synthR is not carried). Equivalent to
 * vp_engine_process_host(e, 1, ...). The block's H2D copies, kernels and D2H copy are replayed from a CUDA graph
 * captured the first time each block phase (position of the block on the frame grids) occurs. */
int vp_engine_stream_buffers(vp_engine* e, float** voice, float** synthL, float** outL);
int vp_engine_stream_block(vp_engine* e);
int vp_engine_stream_stats(const vp_engine* e, uint64_t* graphLaunches, uint64_t* graphCaptures);

/* Decisions of the most recent process call for one stream: up to cap
 * records, *nFrames = frames available. */
int vp_engine_get_pitch_frames(vp_engine* e, int stream, vp_pitch_frame* out, int cap, int* nFrames);
/* Per vocoder frame of the most recent call: gate flag, excitation energies
 * and gain (VocoderProcess.cpp:199-204, :264-272). Arrays may be NULL. */
int vp_engine_get_voc_frames(vp_engine* e, int stream, int cap, int* nFrames, uint8_t* gated, double* EeVoice,
                             double* EeSynth, double* g);

/* Counters since prepare: kernels launched, frames re-decided in FP64. */
int vp_engine_get_stats(const vp_engine* e, uint64_t* kernelLaunches, uint64_t* yinRechecked,
                        uint64_t* yinFrames);
/* Device time (ms) of the most recent process_device call, measured with
 * CUDA events on the engine's stream; per-stage breakdown optional. */
#define VP_NSTAGES 16
int vp_engine_last_timing(vp_engine* e, float* totalMs, float* stageMs /* [VP_NSTAGES] or NULL */);
/* Number of timed intervals (= kernel launches of that stage) behind each stageMs entry. */
int vp_engine_last_timing_counts(vp_engine* e, int* stageCount /* [VP_NSTAGES] */);
const char* vp_stage_name(int stage);
/* accumulate != 0: vp_engine_last_timing sums over every process call since this reset
 * (total = first call's start .. last call's end); 0 (default): most recent call only. */
int vp_engine_timing_reset(vp_engine* e, int accumulate);
/* Device timers on the engine's stream (CUDA events), slots 0..7: bracket any sequence of
 * asynchronous calls; _elapsed_ms waits for slotB and returns the time since slotA. */
int vp_engine_timer_record(vp_engine* e, int slot);
int vp_engine_timer_elapsed_ms(vp_engine* e, int slotA, int slotB, float* ms);

/* ---- memory helpers -------------------------------------------------------- */
int vp_host_alloc(void** p, size_t bytes);   /* pinned host memory */
void vp_host_free(void* p);
int vp_device_alloc(vp_engine* e, void** p, size_t bytes);
void vp_device_free(vp_engine* e, void* p);
int vp_memcpy_h2d(vp_engine* e, void* dst, const void* src, size_t bytes);
int vp_memcpy_d2h(vp_engine* e, void* dst, const void* src, size_t bytes);

/* ---- synthetic inputs -------------------------------------- */
/* flavour: 0 = breathy voice (-40 dBFS aspiration), 1 = clean (-80 dBFS),
 * 2 = breathy with a silent stretch (exercises the gates). Streams are
 * numbered firstStream.. so shards generate disjoint inputs. */
int vp_synth_host(double sampleRate, int flavour, int firstStream, int nStreams, size_t nSamples,
                  size_t strideSamples, float* voice, float* synthL, float* synthR);
int vp_synth_device(vp_engine* e, double sampleRate, int flavour, int firstStream, int nStreams,
                    size_t nSamples, size_t strideSamples, float* voice, float* synthL, float* synthR);

/* ---- measurement helpers ---------------------------------------------------- */
/* Issue-rate microbenchmarks on the engine's device: FP32 FMA and FP64 FMA
 * lane-operations per second (the roofline denominators of DESIGN.md). */
int vp_measure_peaks(vp_engine* e, double* fp32FmaPerSec, double* fp64FmaPerSec);
/* Same, with the operand pattern of the engine's inner loops (acc_i = fma(x, b_i, acc_i): two distinct register
 * operands per FMA instead of one) -- the rate a real multiply-accumulate loop can reach. */
int vp_measure_peaks2(vp_engine* e, double* fp32FmaPerSec, double* fp64FmaPerSec);

#ifdef __cplusplus
}
#endif
#endif /* VP_ENGINE_H */
