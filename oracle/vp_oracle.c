/* TEST INFRASTRUCTURE ONLY (oracle/). Not shipped, not on the product path.
 * See vp_oracle.h. Paths in comments are relative to /root/reference.
 *
 * Pinning: bit-identical float32 output and identical integer decisions vs
 * oracle/_ref (the reference's own C++ built in place) on every fixture in
 * tests/test_oracle.py, incl. the SURVEY.md App. E known-answer table.
 */
#define _GNU_SOURCE
#include "vp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define VPO_PI_VOC 3.14159265               /* VocoderProcess.cpp:13 (truncated pi) */
#define VPO_PI 3.14159265358979323846       /* juce::MathConstants<double>::pi */
#define ORDER_MAX 100                       /* PluginProcessor.cpp:53-59 range ends */
#define SLOTS 128                           /* >= any mark-vector capacity */

typedef struct {
    /* geometry (PluginProcessor.cpp:160-176, MyBuffer.cpp:46-48) */
    double fs;
    int B, hopV, wlenV, hopP, L, c, tauMax, lat, keep, inSize, anCap;
    long n; /* nBlocks*B */
    const float *voice, *s0, *s1;
    double *out0, *out1; /* flat output timeline, n + lat + slack */
    vpo_params prm;
    int ub;
    /* vocoder (VocoderProcess.h:36-107) */
    double *wV, *wS, *eV, *eS, *oV; /* analysis / synthesis window of the vocoder (the same for "sine") */
    double rV[ORDER_MAX + 1], aV[ORDER_MAX + 1], apV[ORDER_MAX + 1];
    double rS[ORDER_MAX + 1], aS[ORDER_MAX + 1], apS[ORDER_MAX + 1];
    double HV[10], HS[10], EeV, EeS, g;
    /* notes (Notes.h:31-41) */
    double freq[SLOTS];
    int nFreq;
    /* pitch (PitchProcess.h:95-157) */
    double fMin, fMax, delta, yinTol;
    int order, nChunk, period, prevPeriod, prevVoicedPeriod, periodNew, stMarkIdx, nAnOv, nStOv;
    double pitch, prevPitch, prevVoicedPitch, closestFreq, prevClosestFreq, beta;
    int an[SLOTS], nAn, pan[SLOTS], nPan, st[SLOTS], nSt, pst[SLOTS], nPst;
    double aP[ORDER_MAX + 1], apP[ORDER_MAX + 1], rP[ORDER_MAX + 1];
    double *yin, *eFrame, *outE, *yF, *stW, *psW, *perS, *xI;
    int eLen;
    long pFrameStart; /* delayed position of the current frame's chunk 0 */
    /* logs */
    vpo_pitch_frame* plog; int plogCap, nP;
    vpo_voc_frame* vlog; int vlogCap, nV;
} vpo_t;

/* ---- MyBuffer ------------------------------------------------------------ */

/* MyBuffer.cpp:142-153 getVoiceSample on the delayed timeline u (App. A.1):
 * input time t = u - latency; the rings start zeroed (MyBuffer.cpp:56-58). */
static inline double xv(const vpo_t* o, long u) {
    long t = u - o->lat;
    return (t >= 0 && t < o->n) ? (double)o->voice[t] : 0.0;
}
static inline double xs(const vpo_t* o, const float* s, long u) {
    long t = u - o->lat;
    return (t >= 0 && t < o->n) ? (double)s[t] : 0.0;
}

/* MyBuffer.cpp:258-261 / :299-302 + juce::AudioBuffer::getRMSLevel: RMS over
 * the whole ring, summed in ring-memory order. Ring slot of input time t is
 * (keep + latency + t) mod inSize (inCounter init, MyBuffer.cpp:60). */
static double ring_rms(const vpo_t* o, const float* x, int block) {
    long tmax = (long)(block + 1) * o->B - 1; /* newest sample in the ring */
    long posmax = (o->keep + o->lat + tmax) % o->inSize;
    double sum = 0.0;
    for (long pos = 0; pos < o->inSize; ++pos) {
        long back = (posmax - pos + o->inSize) % o->inSize;
        long t = tmax - back;
        double s = (t >= 0) ? (double)x[t] : 0.0;
        sum += s * s;
    }
    return sqrt(sum / o->inSize);
}

/* juce::Decibels::gainToDecibels<double> (default -100 dB floor). */
static double gain_to_db(double g) {
    if (g > 0.0) { double d = log10(g) * 20.0; return d > -100.0 ? d : -100.0; }
    return -100.0;
}
/* juce::Decibels::decibelsToGain<float>(dB, -59.0f). */
static float db_to_gain_f(float db) { return db > -59.0f ? powf(10.0f, db * 0.05f) : 0.0f; }

/* ---- LPC ------------------------------------------------------------------ */

/* LPC.cpp:44-97 biaisedAutoCorr: n-outer, m-inner, ((x[n]w[n]) * x[n+m]) * w[n+m]. */
static void autocorr(const vpo_t* o, const float* x, long u0, int order, int wlen, const double* w, double* r) {
    for (int m = 0; m <= order; ++m) r[m] = 0.0;
    for (int n = 0; n < wlen; ++n) {
        double tmp = xs(o, x, u0 + n) * (w ? w[n] : 1.0);
        for (int m = 0; m <= order; ++m)
            if (n < wlen - m) r[m] += tmp * xs(o, x, u0 + n + m) * (w ? w[m + n] : 1.0);
    }
    for (int m = 0; m <= order; ++m) r[m] /= (double)wlen;
}

/* LPC.cpp:107-148 levinsonDurbin. */
static void levinson(const double* r, double* a, double* ap, int order) {
    if (fabs(r[0]) < pow(10, -9)) {
        /* std::fill over the whole vector (size orderMax+1) */
        for (int i = 0; i <= ORDER_MAX; ++i) a[i] = 0.0;
        a[0] = 1.0;
        return;
    }
    a[0] = 1.0;
    a[1] = r[1] / r[0];
    for (int p = 2; p <= order; ++p) {
        for (int j = 1; j < p; ++j) ap[j] = a[j];
        double rho_a = 0.0, r_a = 0.0;
        for (int i = 1; i < p; ++i) { rho_a += r[p - i] * a[i]; r_a += r[i] * a[i]; }
        double k = (r[p] - rho_a) / (r[0] - r_a);
        for (int i = 1; i < p; ++i) a[i] = ap[i] - k * ap[p - i];
        a[p] = k;
    }
    for (int i = 1; i <= order; ++i) a[i] *= -1.;
}

/* ---- defined-behaviour mode ------------------------------------------------------------------
 * 0 (default): the reference as it runs, undefined-behaviour sites modelled (U1, U6) or flagged (U2-U5).
 * 1: every such site takes the bounds-correct reading of what the code says it wants -- the same rules as the engine's
 *    VP_MODE_DEFINED (include/vp_engine.h), so the two can be compared. vpo_defined_deviations() counts how often a
 *    defined-mode choice differed from what mode 0 would have done on the same state (0 => both modes agree exactly). */
static int g_window = 0; /* 0 = "sine" (prepareToPlay), 1 = "hann" (VocoderProcess.cpp:116-124) */
void vpo_set_window(int w) { g_window = w ? 1 : 0; }
static int g_defined = 0;
static long g_deviations = 0;
void vpo_set_defined(int on) { g_defined = on ? 1 : 0; g_deviations = 0; }
long vpo_defined_deviations(void) { return g_deviations; }

/* ---- Notes ---------------------------------------------------------------- */

/* Notes.cpp:43-70 buildFreqVect. freq[nFreq] keeps the popped value (U6). */
static void notes_build(double* freq, int* nFreq, int key, double fMin, double fMax) {
    static const int intervals[7] = {2, 2, 1, 2, 2, 2, 1};
    int n = 0, i = 0;
    double f = 27.5;
    f = f * pow(2, (double)key / 12.0);
    double factorSemiTone = pow(2, 1.0 / 12);
    while (n == 0 || freq[n - 1] < fMax) {
        if (key != 12) f = f * pow(factorSemiTone, intervals[i % 7]);
        else f = f * factorSemiTone;
        if (f > fMin && n < SLOTS) freq[n++] = f;
        i += 1;
    }
    *nFreq = n - 1; /* pop_back: the slot keeps its value */
}

/* Notes.cpp:79-110 getClosestFreq (lower_bound + <= tie rule). */
static double notes_closest(const vpo_t* o, double pitch, int* idxOut) {
    int lo = 0, hi = o->nFreq;
    while (lo < hi) { int mid = (lo + hi) / 2; if (o->freq[mid] < pitch) lo = mid + 1; else hi = mid; }
    int idx = lo, pick;
    if (idx > 0) pick = (fabs(o->freq[idx] - pitch) <= fabs(o->freq[idx - 1] - pitch)) ? idx : idx - 1;
    else pick = idx;
    if (g_defined && idx == o->nFreq) { /* Notes.cpp:99 reads freq[size]: above the table the closest note is the last one */
        if (pick != idx - 1) g_deviations++;
        pick = idx - 1;
    }
    *idxOut = pick;
    return o->freq[pick];
}

/* ---- VocoderProcess ------------------------------------------------------- */

/* VocoderProcess.cpp:235-251 filterFIR (zero initial state, taps on windowed samples). */
static void voc_fir(const vpo_t* o, const float* x, long u0, double* e, const double* a, int order, double* E) {
    *E = 0.0;
    for (int i = 0; i < o->wlenV; ++i) {
        e[i] = a[0] * xs(o, x, u0 + i) * o->wV[i];
        for (int k = 1; k <= order; ++k) {
            if (i - k >= 0) e[i] += xs(o, x, u0 + i - k) * o->wV[i - k] * a[k];
            else break;
        }
        *E += e[i] * e[i];
    }
}

/* VocoderProcess.cpp:190-223 processWindow + :260-297 filterIIR for the frame
 * starting at delayed position u0, processed in host block `block`. */
static void voc_window(vpo_t* o, long u0, int block, int gated) {
    int ov = o->prm.lpcVoice, os = o->prm.lpcSynth;
    vpo_voc_frame* lg = NULL;
    if (o->vlog && o->nV < o->vlogCap) lg = &o->vlog[o->nV];
    if (lg) { memset(lg, 0, sizeof *lg); lg->frame = o->nV; lg->startSample = (int)(u0 - (long)block * o->B); lg->block = block; lg->gated = gated; }
    o->nV++;
    if (!gated) {
        autocorr(o, o->voice, u0, ov, o->wlenV, o->wV, o->rV);
        levinson(o->rV, o->aV, o->apV, ov);
        autocorr(o, o->s0, u0, os, o->wlenV, o->wV, o->rS);
        levinson(o->rS, o->aS, o->apS, os);
        voc_fir(o, o->voice, u0, o->eV, o->aV, ov, &o->EeV);
        voc_fir(o, o->s0, u0, o->eS, o->aS, os, &o->EeS);
        /* filterIIR: shift histories (:301-327), gain, all-pole recursion */
        for (int i = 9; i > 0; --i) { o->HV[i] = o->HV[i - 1]; o->HS[i] = o->HS[i - 1]; }
        o->HV[0] = o->EeV; o->HS[0] = o->EeS;
        if (o->EeS > pow(10, -4)) {
            double sv = 0, ss = 0;
            for (int i = 0; i < 10; ++i) sv += o->HV[i];
            for (int i = 0; i < 10; ++i) ss += o->HS[i];
            o->g = sqrt(sv / ss);
        } else o->g = 0.0;
        for (int i = 0; i < o->wlenV; ++i) {
            o->oV[i] = o->g * o->eS[i];
            for (int k = 1; k <= ov; ++k) {
                if (i - k >= 0) o->oV[i] -= o->oV[i - k] * o->aV[k];
                else break;
            }
        }
        float gv = db_to_gain_f(o->prm.gainVoc);
        for (int i = 0; i < o->wlenV; ++i) {
            double v = gv * o->oV[i] * o->wS[i];
            o->out0[u0 + i] += v;
            o->out1[u0 + i] += v;
        }
    }
    if (lg) { lg->EeVoice = o->EeV; lg->EeSynth = o->EeS; lg->g = o->g; }
}

/* ---- PitchProcess --------------------------------------------------------- */

/* PitchProcess.cpp:752-776 argExt (valley: first index of the minimum). */
static int arg_min(const vpo_t* o, long p, int i0, int i1) {
    double ext = xv(o, p + i0);
    int arg = i0;
    for (int i = i0 + 1; i < i1; ++i)
        if (xv(o, p + i) < ext) { ext = xv(o, p + i); arg = i; }
    return arg;
}

static void an_push(vpo_t* o, int v) { if (o->nAn >= o->anCap) o->ub |= VPO_UB_CAPACITY; if (o->nAn < SLOTS) o->an[o->nAn++] = v; }
static void an_insert_front(vpo_t* o, int v) {
    if (o->nAn >= o->anCap) o->ub |= VPO_UB_CAPACITY;
    if (o->nAn >= SLOTS) return;
    for (int i = o->nAn; i > 0; --i) o->an[i] = o->an[i - 1];
    o->an[0] = v; o->nAn++;
}
static void st_push(vpo_t* o, int v) { if (o->nSt >= o->anCap) o->ub |= VPO_UB_CAPACITY; if (o->nSt < SLOTS) o->st[o->nSt++] = v; }

/* PitchProcess.cpp:350-403 computeYinTemp + :411-448 yin. */
static void yin(vpo_t* o, long p) {
    o->prevPeriod = o->period; o->prevPitch = o->pitch;
    if (o->pitch > 1) { o->prevVoicedPeriod = o->period; o->prevVoicedPitch = o->pitch; }
    o->pitch = 0; o->period = 0;
    double* y = o->yin;
    int tauMax = o->tauMax;
    for (int k = 0; k < tauMax; ++k) y[k] = 0.0;
    for (int i = 0; i < o->L; ++i) {
        double vi = xv(o, p + i - tauMax);
        for (int k = 0; k < tauMax; ++k) { double d = vi - xv(o, p + i - tauMax + k); y[k] += d * d; }
    }
    y[0] = 1.0;
    double tmp = 0;
    for (int k = 1; k < tauMax; ++k) { tmp += y[k]; y[k] *= k / tmp; }
    int tau = (int)floor(o->fs / o->fMax);
    while (tau < tauMax) {
        if (y[tau] < o->yinTol) {
            for (;;) {
                if (tau + 1 >= tauMax) { if (!g_defined) o->ub |= VPO_UB_YIN_END; break; } /* U3: reads yinTemp[tauMax]; defined: the descent ends at the last lag */
                if (!(y[tau + 1] < y[tau])) break;
                tau += 1;
                if (tau + 1 >= tauMax) break;
            }
            o->pitch = o->fs / tau; o->period = tau;
            break;
        } else tau += 1;
    }
}

/* PitchProcess.cpp:455-567 pitchMarks. */
static void pitch_marks(vpo_t* o, long p) {
    int L = o->L;
    memcpy(o->pan, o->an, sizeof(int) * (size_t)o->nAn); o->nPan = o->nAn; /* prevAnMarks = anMarks */
    o->nAn = 0;                                                            /* clear(): slots keep values */
    o->nAnOv = 0;
    for (int i = 0; i < o->nPan; ++i) o->pan[i] -= o->hopP;
    for (int i = 0; i < o->nPan; ++i) if (o->pan[i] >= 0) o->nAnOv += 1;
    int searchLeft = 0, t, l_lim, r_lim, lastMark, sw_c, sw_f;
    if (o->pitch > 1) {
        sw_c = (int)floor(o->delta * o->period);
        sw_f = (int)ceil((2.0 - o->delta) * o->period);
        /* defined: "previous frame voiced" needs previous marks to continue from; without any (a gated frame cleared them,
         * :208-214) the frame is searched like the first voiced frame after an unvoiced one */
        if (o->prevPitch > 1 && !(g_defined && o->nPan == 0)) {
            if (o->nAnOv == 0) {
                if (o->nPan == 0) { o->ub |= VPO_UB_PREV_EMPTY; lastMark = 0; } /* U4 */
                else lastMark = o->pan[o->nPan - 1];
                int mn = o->prevPeriod < o->period ? o->prevPeriod : o->period;
                int mx = o->prevPeriod > o->period ? o->prevPeriod : o->period;
                int a1 = (int)floor(o->delta * mn), a2 = (int)ceil((2 - o->delta) * mx);
                l_lim = lastMark + (sw_c < a1 ? sw_c : a1); if (l_lim < 0) l_lim = 0;
                r_lim = lastMark + (sw_f > a2 ? sw_f : a2); if (r_lim > L) r_lim = L;
                t = arg_min(o, p, l_lim, r_lim);
            } else t = o->pan[o->nPan - o->nAnOv];
        } else { if (o->prevPitch > 1) g_deviations++; searchLeft = 1; t = arg_min(o, p, 0, L); }
        an_push(o, t);
        while (o->an[o->nAn - 1] + sw_c < L) {
            int bk = o->an[o->nAn - 1];
            if (bk + sw_f < L) an_push(o, arg_min(o, p, bk + sw_c, bk + sw_f));
            else {
                if (bk + o->period < L) { an_push(o, arg_min(o, p, bk + sw_c, L)); break; }
                else break;
            }
        }
        if (searchLeft) {
            while (o->an[0] - sw_c > 0) {
                int fr = o->an[0];
                if (fr - sw_f >= 0) an_insert_front(o, arg_min(o, p, fr - sw_f, fr - sw_c));
                else {
                    if (fr - o->period >= 0) { an_insert_front(o, arg_min(o, p, 0, fr - sw_c)); break; }
                    else break;
                }
            }
        }
    } else if (o->nPan > 0) {
        if (o->nAnOv > 0) for (int i = 0; i < o->nAnOv; ++i) an_push(o, o->pan[o->nPan - o->nAnOv + i]);
        else an_push(o, o->pan[o->nPan - 1] + o->prevVoicedPeriod);
        if (o->prevVoicedPeriod <= 0) { o->ub |= VPO_UB_ASSERT; return; }
        while (o->an[o->nAn - 1] + o->prevVoicedPeriod < L) an_push(o, o->an[o->nAn - 1] + o->prevVoicedPeriod);
    }
}

/* PitchProcess.cpp:573-658 placeStMarks. *noteIdx: table index or -1. */
static void place_st_marks(vpo_t* o, int* noteIdx) {
    *noteIdx = -1;
    memcpy(o->pst, o->st, sizeof(int) * (size_t)o->nSt); o->nPst = o->nSt;
    o->nSt = 0; o->nStOv = 0;
    for (int i = 0; i < o->nPst; ++i) o->pst[i] -= o->hopP;
    if (o->nAn == 0) return;
    for (int i = 0; i < o->nPst; ++i) if (o->pst[i] >= 0) o->nStOv += 1;
    int firstMark;
    o->prevClosestFreq = o->closestFreq;
    if (o->pitch > 1) {
        o->closestFreq = notes_closest(o, o->pitch, noteIdx);
        o->beta = o->closestFreq / o->pitch;
        o->periodNew = (int)round(o->period / o->beta);
    } else { o->closestFreq = 0; o->periodNew = o->prevVoicedPeriod; }
    if (o->periodNew <= 0) { o->ub |= VPO_UB_ASSERT; return; }
    if (o->pitch > 1) {
        if (o->prevPitch > 1) {
            if (o->nStOv > 0) firstMark = o->pst[o->nPst - o->nStOv];
            else if (o->nPst == 0) { if (!g_defined) o->ub |= VPO_UB_PREV_EMPTY; firstMark = o->an[0]; } /* defined: start from the first analysis mark */
            else if (o->pst[o->nPst - 1] + o->periodNew >= 0) firstMark = o->pst[o->nPst - 1] + o->periodNew;
            else firstMark = o->an[0];
        } else firstMark = o->an[0];
    } else {
        if (o->nPst == 0) return;
        if (o->nStOv > 0) firstMark = o->pst[o->nPst - o->nStOv];
        else {
            int n = 1;
            while (o->pst[o->nPst - 1] + n * o->periodNew < 0) n += 1;
            firstMark = o->pst[o->nPst - 1] + n * o->periodNew;
        }
    }
    st_push(o, firstMark);
    while (o->st[o->nSt - 1] + o->periodNew < o->L) st_push(o, o->st[o->nSt - 1] + o->periodNew);
}

/* PitchProcess.cpp:280-302 filterFIR. startSample: block-relative start of
 * the current chunk; blockBase: delayed position of the block start. */
static void pitch_fir(vpo_t* o, long blockBase, int startSample, int startIdxBuf, int nFilt, int startIdxE) {
    int minBufIdx = -o->keep;
    for (int i = 0; i < nFilt; ++i) {
        int idx = startSample + startIdxBuf + i;
        double e = o->aP[0] * xv(o, blockBase + idx);
        for (int k = 1; k <= o->order; ++k) {
            if (idx - k >= minBufIdx) e += xv(o, blockBase + idx - k) * o->aP[k];
            else break;
        }
        if (startIdxE + i < o->eLen) o->eFrame[startIdxE + i] = e;
    }
}

/* PitchProcess.cpp:307-322 filterIIR for chunk nChunk. */
static void pitch_iir(vpo_t* o) {
    int shift = o->nChunk * o->c;
    for (int i = 0; i < o->c; ++i) {
        o->yF[i + shift] = o->outE[i + shift];
        for (int k = 1; k <= o->order; ++k) {
            if (i + shift - k >= 0) o->yF[i + shift] -= o->yF[i + shift - k] * o->aP[k];
            else break;
        }
    }
}

/* PitchProcess.cpp:328-342 fillOutputBuffer for chunk nChunk at delayed position P. */
static void pitch_out(vpo_t* o, long P) {
    float gain = db_to_gain_f(o->prm.gainPitch);
    for (int i = 0; i < o->c; ++i) {
        double v = o->yF[i + o->nChunk * o->c] * o->stW[i + o->nChunk * o->c] * gain;
        o->out0[P + i] += v;
        o->out1[P + i] += v;
    }
}

/* PitchProcess.cpp:788-831 getClosestAnMarkIdx. lookahead = bufferIdxMax - startSample. */
static int closest_an(vpo_t* o, int stMark, int T, int lookahead, int* bad) {
    int lo = 0, hi = o->nAn;
    while (lo < hi) { int mid = (lo + hi) / 2; if (o->an[mid] < stMark) lo = mid + 1; else hi = mid; }
    int idx = lo, sz = o->nAn, nc = o->nChunk * o->c;
    *bad = 0;
    if (idx > 0 && idx < sz) {
        if (abs(o->an[idx] - stMark) <= abs(o->an[idx - 1] - stMark) && o->an[idx] + T - nc < lookahead) return idx;
        else if (o->an[idx - 1] + T - nc < lookahead) return idx - 1;
        else if (idx - 2 > 0) return idx - 2;
        else return -o->nAnOv - 1;
    } else if (idx == 0) return idx;
    else { /* idx == size: reads the stale storage slot an[size] (U1) */
        if (g_defined) {
            /* the completeness test is meant for the mark it is about to return: the last one */
            const int parity = (o->an[idx] + T - nc < lookahead) ? idx - 1 : (idx - 2 >= 0 ? idx - 2 : -99);
            const int pick = (o->an[idx - 1] + T - nc < lookahead) ? idx - 1 : (idx - 2 >= 0 ? idx - 2 : idx - 1);
            if (pick != parity) g_deviations++;
            return pick;
        }
        if (o->an[idx] + T - nc < lookahead) return idx - 1;
        else if (idx - 2 >= 0) return idx - 2;
        else { *bad = 1; return 0; }
    }
}

/* PitchProcess.cpp:842-870 interp. */
static void interp(vpo_t* o, const double* x, const double* y, int len, int startIdx, int stopIdx) {
    int search = 0;
    for (int i = startIdx; i < stopIdx; ++i) {
        if (i >= x[0] && i <= x[len - 1]) {
            int lo = search, hi = len;
            if (lo < 0) { if (!g_defined) o->ub |= VPO_UB_INTERP; lo = 0; } /* U5; defined: the search restarts at begin() */
            while (lo < hi) { int mid = (lo + hi) / 2; if (x[mid] < (double)i) lo = mid + 1; else hi = mid; }
            int lb = lo;
            search = lb - 1;
            double value;
            if (lb > 0) value = y[lb - 1] + (y[lb] - y[lb - 1]) / (x[lb] - x[lb - 1]) * (i - x[lb - 1]);
            else value = y[lb];
            o->outE[i] += value;
        } else if (i > x[len - 1]) break;
    }
}

/* PitchProcess.cpp:665-741 psola for the current chunk. */
static void psola(vpo_t* o, int startSample) {
    int T = (o->pitch > 1) ? o->period : o->prevVoicedPeriod;
    int len = 2 * T + 1;
    if (T <= 0) { o->ub |= VPO_UB_ASSERT; return; }
    /* fillPsolaWindow :878-882 (juce hann, size 2T+1) */
    for (int i = 0; i < len; ++i) o->psW[i] = 0.5 - 0.5 * cos((double)(2 * i) * VPO_PI / (double)(len - 1));
    int lookahead = (o->lat + o->B) - startSample;
    while (o->stMarkIdx < o->nSt) {
        int stMark = o->st[o->stMarkIdx];
        if (stMark - T >= (o->nChunk + 1) * o->c) break;
        int bad, clAnMark;
        int clIdx = closest_an(o, stMark, T, lookahead, &bad);
        if (bad) { o->ub |= VPO_UB_ASSERT; clAnMark = o->an[0]; }
        else if (clIdx >= 0) clAnMark = o->an[clIdx];
        else if (g_defined) {
            /* :812 "take last element of non overlapping of prevAnMarks": prevAnMarks[size - nOv - 1] (the code's sign slip
             * makes it size + nOv + 1); without such a mark, the frame's first one */
            const int k = o->nPan - o->nAnOv - 1;
            clAnMark = (k >= 0) ? o->pan[k] : o->an[0];
            g_deviations++;
        }
        else { o->ub |= VPO_UB_CLOSEST_PREV; clAnMark = 0; } /* U2: out-of-bounds prevAnMarks read */
        int first = (o->stMarkIdx == 0), last = (o->stMarkIdx == o->nSt - 1);
        for (int j = 0; j < len; ++j) {
            int ei = o->keep + clAnMark - T + j;
            double e = (ei >= 0 && ei < o->eLen) ? o->eFrame[ei] : 0.0;
            int win;
            if (!first && !last) win = 1;
            else if (first) win = (j >= T);
            else win = (j < T);
            o->perS[j] = win ? e * o->psW[j] : e;
            o->xI[j] = stMark + (-T + j) / o->beta;
        }
        int startIdx = (int)floor(o->xI[0]); if (startIdx < 0) startIdx = 0;
        int stopIdx = (int)ceil(o->xI[len - 1]); if (stopIdx > o->L) stopIdx = o->L;
        interp(o, o->xI, o->perS, len, startIdx, stopIdx);
        o->stMarkIdx += 1;
    }
}

/* PitchProcess.cpp:203-247 processChunkStart at delayed position p (block, startSample). */
static void chunk_start(vpo_t* o, long p, int block, int startSample, int gated) {
    vpo_pitch_frame* lg = NULL;
    if (o->plog && o->nP < o->plogCap) lg = &o->plog[o->nP];
    int frameNo = o->nP++;
    int note = -1;
    o->pFrameStart = p;
    if (gated) {
        o->nAn = 0; o->prevPitch = 0;
    } else {
        for (int i = 0; i < o->eLen; ++i) o->eFrame[i] = 0.0;
        for (int i = 0; i < o->L; ++i) { o->outE[i] = 0.0; o->yF[i] = 0.0; }
        yin(o, p);
        pitch_marks(o, p);
        place_st_marks(o, &note);
        if (o->nAn > 0) {
            autocorr(o, o->voice, p, o->order, o->L, NULL, o->rP);
            levinson(o->rP, o->aP, o->apP, o->order);
            pitch_fir(o, (long)block * o->B, startSample, -o->keep, o->keep + o->L, 0);
            o->stMarkIdx = 0;
            psola(o, startSample);
            pitch_iir(o);
        }
        pitch_out(o, p);
    }
    if (lg) {
        memset(lg, 0, sizeof *lg);
        lg->frame = frameNo; lg->startSample = startSample; lg->block = block; lg->gated = gated;
        lg->period = o->period; lg->periodNew = o->periodNew; lg->prevVoicedPeriod = o->prevVoicedPeriod;
        lg->pitch = o->pitch; lg->closestFreq = o->closestFreq; lg->beta = o->beta;
        lg->note = (!gated && o->pitch > 1 && o->nAn > 0) ? note : -1;
        lg->nAn = o->nAn; lg->nSt = o->nSt;
        for (int i = 0; i < o->nAn && i < VPO_MAX_MARKS; ++i) lg->anMarks[i] = o->an[i];
        for (int i = 0; i < o->nSt && i < VPO_MAX_MARKS; ++i) lg->stMarks[i] = o->st[i];
        lg->anStale = (o->nAn < o->anCap) ? o->an[o->nAn] : 0;
    }
}

/* PitchProcess.cpp:253-271 processChunkCont for chunk nChunk at delayed position P. */
static void chunk_cont(vpo_t* o, long P, int block, int startSample) {
    if (o->nAn > 0) {
        pitch_fir(o, (long)block * o->B, startSample, o->L - o->c, o->c, o->keep + o->L + (o->nChunk - 1) * o->c);
        psola(o, startSample);
        pitch_iir(o);
        pitch_out(o, P);
    }
}

/* ---- driver ---------------------------------------------------------------- */

void vpo_default_params(vpo_params* p) {
    p->gainPitch = 0.f; p->gainVoice = -60.f; p->gainSynth = -60.f; p->gainVoc = 0.f;
    p->lpcVoice = 40; p->lpcPitch = 15; p->lpcSynth = 5; p->keyPitch = 12; p->pitchBool = 1; p->vocBool = 1;
}

/* PluginProcessor.cpp:160-176, PitchProcess.cpp:100-107, MyBuffer.cpp:46-48. */
void vpo_sizes_for(double fs, int B, int key, vpo_sizes* s) {
    double ratio = fs / 44100.0;
    s->hopV = (int)floor(128.0 * ratio); s->wlenV = 4 * s->hopV;
    int c256 = (int)floor(256.0 * ratio);
    s->hopP = 3 * c256; s->frameLenP = 4 * c256; s->chunk = s->frameLenP - s->hopP;
    s->tauMax = (int)ceil(fs / 100.0);
    s->latency = s->frameLenP > s->wlenV ? s->frameLenP : s->wlenV;
    s->keep = s->frameLenP;
    s->inSize = s->keep + B + s->latency; s->outSize = B + s->latency;
    s->anCap = (int)ceil(s->frameLenP * 800.0 / fs) + 1;
    double fr[SLOTS]; int nf;
    notes_build(fr, &nf, key, 100.0, 800.0);
    s->nFreq = nf;
}

int vpo_notes(int key, double fMin, double fMax, double* freq, int cap, double* popped) {
    double fr[SLOTS]; int nf;
    notes_build(fr, &nf, key, fMin, fMax);
    for (int i = 0; i < nf && i < cap; ++i) freq[i] = fr[i];
    if (popped) *popped = fr[nf];
    return nf;
}

int vpo_process(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                const vpo_params* params, float* outL, float* outR, vpo_sizes* sizes,
                vpo_pitch_frame* plog, int plogCap, int* nP, vpo_voc_frame* vlog, int vlogCap, int* nV,
                int* ubFlags) {
    return vpo_process_sched(fs, B, nBlocks, voice, synthL, synthR, params, NULL, NULL, 0, outL, outR, sizes, plog, plogCap,
                             nP, vlog, vlogCap, nV, ubFlags);
}

/* Same, with parameter automation: before block schedBlock[i] the parameter tree takes sched[i] (every read site of
 * the reference loads the atomics afresh: VocoderProcess.cpp:193-194,291, PitchProcess.cpp:206,336,
 * PluginProcessor.cpp:212-230; lpcPitch is only read in prepare, PitchProcess.cpp:70). */
int vpo_process_sched(double fs, int B, int nBlocks, const float* voice, const float* synthL, const float* synthR,
                      const vpo_params* params, const vpo_params* sched, const int* schedBlock, int nSched,
                      float* outL, float* outR, vpo_sizes* sizes,
                      vpo_pitch_frame* plog, int plogCap, int* nP, vpo_voc_frame* vlog, int vlogCap, int* nV,
                      int* ubFlags) {
    if (!voice || !synthL || !params || !outL || B <= 0 || nBlocks < 0 || fs <= 0) return -1;
    if (nSched > 0 && (!sched || !schedBlock)) return -1;
    for (int i = 0; i < nSched; ++i)
        if (sched[i].lpcVoice > ORDER_MAX || sched[i].lpcSynth > ORDER_MAX || sched[i].lpcVoice < 1 || sched[i].lpcSynth < 1) return -2;
    if (params->lpcVoice > ORDER_MAX || params->lpcPitch > ORDER_MAX || params->lpcSynth > ORDER_MAX ||
        params->lpcVoice < 1 || params->lpcPitch < 1 || params->lpcSynth < 1) return -2;
    vpo_t* o = (vpo_t*)calloc(1, sizeof(vpo_t));
    if (!o) return -3;
    vpo_sizes sz;
    vpo_sizes_for(fs, B, params->keyPitch, &sz);
    if (sizes) *sizes = sz;
    o->fs = fs; o->B = B; o->hopV = sz.hopV; o->wlenV = sz.wlenV; o->hopP = sz.hopP; o->L = sz.frameLenP;
    o->c = sz.chunk; o->tauMax = sz.tauMax; o->lat = sz.latency; o->keep = sz.keep; o->inSize = sz.inSize;
    o->anCap = sz.anCap; o->n = (long)nBlocks * B;
    o->voice = voice; o->s0 = synthL; o->s1 = synthR ? synthR : synthL;
    o->prm = *params; o->plog = plog; o->plogCap = plogCap; o->vlog = vlog; o->vlogCap = vlogCap;
    size_t outLen = (size_t)o->n + (size_t)o->lat + (size_t)o->wlenV + 16;
    o->out0 = (double*)calloc(outLen, sizeof(double));
    o->out1 = (double*)calloc(outLen, sizeof(double));
    o->wV = (double*)calloc((size_t)o->wlenV, sizeof(double));
    o->wS = (double*)calloc((size_t)o->wlenV, sizeof(double));
    o->eV = (double*)calloc((size_t)o->wlenV, sizeof(double));
    o->eS = (double*)calloc((size_t)o->wlenV, sizeof(double));
    o->oV = (double*)calloc((size_t)o->wlenV, sizeof(double));
    o->eLen = o->inSize + 3 * o->c; /* PitchProcess.cpp:136 */
    o->yin = (double*)calloc((size_t)o->tauMax + 1, sizeof(double));
    o->eFrame = (double*)calloc((size_t)o->eLen, sizeof(double));
    o->outE = (double*)calloc((size_t)o->L, sizeof(double));
    o->yF = (double*)calloc((size_t)o->L, sizeof(double));
    o->stW = (double*)calloc((size_t)o->L, sizeof(double));
    o->psW = (double*)calloc(2 * (size_t)o->tauMax + 3, sizeof(double));
    o->perS = (double*)calloc(2 * (size_t)o->tauMax + 3, sizeof(double));
    o->xI = (double*)calloc(2 * (size_t)o->tauMax + 3, sizeof(double));

    /* VocoderProcess::prepare (:35-71) + setWindows "sine" (:95-135) */
    {
        double overlap = (double)(o->wlenV - o->hopV) / (double)o->wlenV;
        double factor = 1.0;
        if (fabs(overlap - 0.75) < pow(10, -10)) factor = 1.0 / sqrt(2);
        for (int i = 0; i < o->wlenV; ++i) {
            if (g_window == 1) { /* "hann" (:116-124): juce hann table (not normalised) x factor for synthesis, no analysis window */
                o->wS[i] = (0.5 - 0.5 * cos((double)(2 * i) * VPO_PI / (double)(o->wlenV - 1))) * factor;
                o->wV[i] = 1.0;
            } else {
                o->wV[i] = factor * sin((i + 0.5) * VPO_PI_VOC / (double)o->wlenV);
                o->wS[i] = o->wV[i];
            }
        }
        o->g = 0.0; o->EeS = 1.0; o->EeV = 0.0;
        for (int i = 0; i <= ORDER_MAX; ++i) { o->rV[i] = o->aV[i] = o->apV[i] = 1.0; o->rS[i] = o->aS[i] = o->apS[i] = 1.0; }
    }
    /* PitchProcess::prepare (:62-128) + buildWindows (:889-905) */
    {
        o->fMin = 100; o->fMax = 800; o->delta = 0.94; o->yinTol = 0.25; o->beta = 1;
        o->order = params->lpcPitch;
        notes_build(o->freq, &o->nFreq, params->keyPitch, o->fMin, o->fMax);
        double overlap = ((double)(o->L - o->hopP)) / ((double)o->L);
        int h = (int)round(overlap * o->L);
        for (int i = 0; i < o->L; ++i) o->stW[i] = 1.0;
        for (int i = 0; i < 2 * h; ++i) {
            double w = 0.5 - 0.5 * cos((double)(2 * i) * VPO_PI / (double)(2 * h - 1));
            if (i < h) o->stW[i] = w; else o->stW[o->L - 2 * h + i] = w;
        }
    }

    int startV = 0, startP = 0; /* VocoderProcess::startSample, PitchProcess::startSample */
    for (int b = 0; b < nBlocks; ++b) {
        long base = (long)b * B;
        for (int i = 0; i < nSched; ++i)
            if (schedBlock[i] == b) {
                const int keyOld = o->prm.keyPitch, ordP = o->prm.lpcPitch;
                o->prm = sched[i];
                o->prm.lpcPitch = ordP; /* read in prepare only */
                /* Notes::getClosestFreq rebuilds the table when the key differs (Notes.cpp:83-88) */
                if (o->prm.keyPitch != keyOld) notes_build(o->freq, &o->nFreq, o->prm.keyPitch, o->fMin, o->fMax);
            }
        params = &o->prm;
        const float gVoice = params->gainVoice, gSynth = params->gainSynth;
        int gateV = -1, gateS = -1; /* lazily evaluated ring RMS gates for this block */
        /* VocoderProcess::process :173-183 */
        if (params->vocBool) {
            while (startV < B) {
                if (gateV < 0) { gateV = gain_to_db(ring_rms(o, voice, b)) < -60.0; gateS = gain_to_db(ring_rms(o, synthL, b)) < -60.0; }
                voc_window(o, base + startV, b, gateV || gateS);
                startV += o->hopV;
            }
            startV -= B;
        }
        /* PitchProcess::process :166-196 */
        if (params->pitchBool) {
            while (startP < B) {
                if (gateV < 0) gateV = gain_to_db(ring_rms(o, voice, b)) < -60.0;
                long P = base + startP;
                if (o->nChunk % 4 == 3) {
                    chunk_cont(o, P, b, startP);
                    o->nChunk = 0;
                    chunk_start(o, P, b, startP, gateV);
                    o->nChunk += 1; o->nChunk %= 4;
                } else if (o->nChunk == 0) {
                    chunk_start(o, P, b, startP, gateV);
                    o->nChunk += 1;
                } else {
                    chunk_cont(o, P, b, startP);
                    o->nChunk += 1;
                }
                startP += o->c;
            }
            startP -= B;
        } else {
            /* PitchProcess::silence :146-158 */
            o->nAn = 0; o->nSt = 0; o->prevPitch = 0; o->prevPeriod = 0; o->pitch = 0; o->period = 0;
        }
        /* PluginProcessor.cpp:226-230 dry voice / synth mix (MyBuffer.cpp:309-448) */
        if (gVoice > -59.0) {
            double g = db_to_gain_f(gVoice);
            for (int i = 0; i < B; ++i) { double v = xv(o, base + i) * g; o->out0[base + i] += v; o->out1[base + i] += v; }
        }
        if (gSynth > -59.0) {
            double g = db_to_gain_f(gSynth);
            for (int i = 0; i < B; ++i) { o->out0[base + i] += xs(o, o->s0, base + i) * g; o->out1[base + i] += xs(o, o->s1, base + i) * g; }
        }
        /* MyBuffer::fillOutputBuffer :113-133 */
        for (int i = 0; i < B; ++i) {
            outL[base + i] = (float)o->out0[base + i];
            if (outR) outR[base + i] = (float)o->out1[base + i];
        }
    }
    if (nP) *nP = o->nP;
    if (nV) *nV = o->nV;
    if (ubFlags) *ubFlags = o->ub;
    free(o->out0); free(o->out1); free(o->wV); free(o->wS); free(o->eV); free(o->eS); free(o->oV); free(o->yin);
    free(o->eFrame); free(o->outE); free(o->yF); free(o->stW); free(o->psW); free(o->perS); free(o->xI);
    free(o);
    return 0;
}

/* ---- CPU "port" baseline ---------------------------------------------------- */
typedef struct {
    double fs; int B, nBlocks, S; const float *voice, *sl, *sr; const vpo_params* prm; float* out;
    int* next; pthread_mutex_t* mu;
} bench_arg;

static void* bench_worker(void* v) {
    bench_arg* a = (bench_arg*)v;
    size_t n = (size_t)a->nBlocks * a->B;
    float* tmp = a->out ? NULL : (float*)malloc(sizeof(float) * n * 2);
    for (;;) {
        pthread_mutex_lock(a->mu);
        int s = (*a->next)++;
        pthread_mutex_unlock(a->mu);
        if (s >= a->S) break;
        float* oL = a->out ? a->out + ((size_t)s * 2) * n : tmp;
        float* oR = a->out ? a->out + ((size_t)s * 2 + 1) * n : tmp + n;
        vpo_process(a->fs, a->B, a->nBlocks, a->voice + s * n, a->sl + s * n, a->sr ? a->sr + s * n : NULL, a->prm,
                    oL, oR, NULL, NULL, 0, NULL, NULL, 0, NULL, NULL);
    }
    free(tmp);
    return NULL;
}

double vpo_bench(double fs, int B, int nBlocks, int S, const float* voice, const float* synthL,
                 const float* synthR, const vpo_params* params, int nThreads, float* out) {
    if (nThreads < 1) nThreads = 1;
    int next = 0;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    bench_arg a = {fs, B, nBlocks, S, voice, synthL, synthR, params, out, &next, &mu};
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nThreads);
    for (int t = 1; t < nThreads; ++t) pthread_create(&th[t], NULL, bench_worker, &a);
    bench_worker(&a);
    for (int t = 1; t < nThreads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
