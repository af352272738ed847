// Pitch-corrector path kernels (sm_100a): YIN period detection, pitch-mark
// chain + note snapping, LPC residual + PSOLA, all-pole resynthesis + OLA.
// Reference behaviour: Source/PitchProcess.cpp:166-342 (scheduling, filters),
// :350-448 (YIN), :455-658 (marks), :665-870 (PSOLA); Source/Notes.cpp:79-110.
// SURVEY.md App. A.4 / App. B describe the semantics that are restated here.
#include "vp_common.cuh"

// ===========================================================================
// YIN (PitchProcess.cpp:350-448). One CTA per (frame, stream).
//   d[k]  = sum_{i<L} (x[q+i] - x[q+i+k])^2,  q = p - tauMax,  k < tauMax
//   d'[k] = d[k] * k / sum_{j<=k} d[j]
// Direct (a-b)^2 form: all terms positive, so FP32 accumulation keeps ~1e-6
// relative accuracy even in deep minima. Thread tile: R consecutive lags x one
// i-range, sliding window of R samples in registers (2 LDS per 2R FP32 ops;
// odd R -> conflict-free window loads). CMND prefix sum, threshold search and
// descent run on warp 0 in FP64. A frame whose deciding comparisons are within
// yinEps (relative) is queued for the FP64 re-check kernel, which runs the same
// code with T = double.
// ===========================================================================
#define YF_RECHECK 1u
#define YF_NEAR 2u
#define YF_UB 4u
#define YF_DONE64 8u

template <int R, typename T, bool MASK>
__device__ __forceinline__ void yin_round(const float* __restrict__ xs, int nb, int nEnd, int k0, T* W, T* acc) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int n = nb + u;
        if (u > 0) W[(u + R - 1) % R] = (T)xs[n + k0 + R - 1];
        const T a = (T)xs[n];
        const T msk = (!MASK || n < nEnd) ? (T)1 : (T)0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            T d = a - W[(u + j) % R];
            if (MASK) d *= msk;
            acc[j] = fma(d, d, acc[j]);
        }
    }
    W[(R - 1) % R] = (T)xs[nb + R + k0 + R - 1];
}

template <int R, typename T>
__global__ void __launch_bounds__(512) k_yin(VPGeom g, const float* __restrict__ voice, const uint8_t* __restrict__ gate,
                                             int* __restrict__ period, uint32_t* __restrict__ yflags,
                                             int* __restrict__ list, int* __restrict__ listCount, int maxList, int LT,
                                             int IG, int xsLen, bool fromList) {
    extern __shared__ double smd[];
    const int tauMax = g.tauMax, L = g.L;
    const int lagPad = LT * R;
    double* dp = smd;                                 // [lagPad] d then d'
    float* xs = (float*)(smd + lagPad);               // [xsLen]
    T* part = (T*)(xs + ((xsLen + 3) & ~3));          // [IG][lagPad]

    for (int item = blockIdx.x;; item += gridDim.x) {
        long long fidx;
        if (fromList) {
            int cnt = *listCount;
            if (cnt > maxList) cnt = maxList;
            if (item >= cnt) return;
            fidx = list[item];
        } else {
            if (item >= g.nFramesP) return;
            fidx = (long long)blockIdx.y * g.nFramesP + item;
        }
        const int s = (int)(fidx / g.nFramesP), f = (int)(fidx - (long long)s * g.nFramesP);
        const long long p = (long long)f * g.hopP;
        const int b = (int)(p / g.B);
        if (gate[(size_t)s * g.nBlocks + b] & VP_GATE_VOICE) {  // gated: yin() is not run (PitchProcess.cpp:208-214)
            if (threadIdx.x == 0) { period[fidx] = 0; yflags[fidx] = 0; }
            if (!fromList) return;
            continue;
        }
        const float* v = voice + (size_t)s * g.stride;
        const long long q = p - tauMax;
        __syncthreads();
        for (int j = threadIdx.x; j < xsLen; j += blockDim.x) xs[j] = (j < L + tauMax) ? vp_x(v, q + j, g.lat, g.n) : 0.0f;
        __syncthreads();
        {
            const int lt = threadIdx.x % LT, ig = threadIdx.x / LT;
            const int iLen = ((L + IG - 1) / IG + R - 1) / R * R;  // multiple of R
            const int i0 = ig * iLen;
            int iEnd = i0 + iLen;
            if (iEnd > L) iEnd = L;
            const int k0 = lt * R;
            T W[R], acc[R];
#pragma unroll
            for (int j = 0; j < R; ++j) { acc[j] = (T)0; W[j] = (i0 < L) ? (T)xs[i0 + k0 + j] : (T)0; }
            int nb = i0;
            for (; nb + R <= iEnd; nb += R) yin_round<R, T, false>(xs, nb, iEnd, k0, W, acc);
            if (nb < iEnd) yin_round<R, T, true>(xs, nb, iEnd, k0, W, acc);
#pragma unroll
            for (int j = 0; j < R; ++j) part[ig * lagPad + k0 + j] = acc[j];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < lagPad; k += blockDim.x) {
            double d = 0.0;
            for (int ig = 0; ig < IG; ++ig) d += (double)part[ig * lagPad + k];
            dp[k] = d;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int per = (tauMax + 31) / 32;
            const int kA = lane * per, kB = min(kA + per, tauMax);
            // cumulative mean normalisation (PitchProcess.cpp:396-402); d'[0] = 1 and the sum starts at k = 1
            double loc = 0.0;
            for (int k = max(kA, 1); k < kB; ++k) loc += dp[k];
            double incl = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            double run = incl - loc;
            const double total = __shfl_sync(0xffffffffu, incl, 31);
            for (int k = max(kA, 1); k < kB; ++k) {
                run += dp[k];
                dp[k] = dp[k] * ((double)k / run);
            }
            if (lane == 0) dp[0] = 1.0;
            __syncwarp();
            // first tau >= tauMin with d'[tau] < yinTol (PitchProcess.cpp:429-433)
            const double tol = 0.25;
            int first = 0x7fffffff;
            for (int k = max(kA, g.tauMin); k < kB; ++k)
                if (dp[k] < tol) { first = k; break; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            int per_ = 0;
            unsigned fl = 0;
            double margin = 1e300;
            if (first != 0x7fffffff && total > 0.0) {
                // descent while d'[tau+1] < d'[tau]  (PitchProcess.cpp:435-440)
                if (first + 1 >= tauMax) fl |= YF_UB;  // U3: reads yinTemp[tauMax]
                int stop = 0x7fffffff;
                for (int k = max(kA, first); k < kB; ++k)
                    if (k + 1 >= tauMax || !(dp[k + 1] < dp[k])) { stop = k; break; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) stop = min(stop, __shfl_xor_sync(0xffffffffu, stop, o));
                per_ = stop;
                for (int k = max(kA, g.tauMin); k < kB; ++k) {
                    if (k <= first) margin = fmin(margin, fabs(dp[k] - tol) / tol);
                    if (k >= first && k <= stop && k + 1 < tauMax)
                        margin = fmin(margin, fabs(dp[k + 1] - dp[k]) / fmax(fabs(dp[k]), 1e-300));
                }
            } else if (total > 0.0) {
                for (int k = max(kA, g.tauMin); k < kB; ++k) margin = fmin(margin, fabs(dp[k] - tol) / tol);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) margin = fmin(margin, __shfl_xor_sync(0xffffffffu, margin, o));
            if (lane == 0) {
                const bool is64 = sizeof(T) == 8;
                if (margin < (is64 ? 1e-12 : g.yinEps)) fl |= is64 ? YF_NEAR : YF_RECHECK;
                if (is64) fl |= YF_DONE64;
                period[fidx] = per_;
                yflags[fidx] = fl;
                if (!is64 && (fl & YF_RECHECK)) {
                    const int slot = atomicAdd(listCount, 1);
                    if (slot < maxList) list[slot] = (int)fidx;
                }
            }
        }
        if (!fromList) return;
    }
}

struct YinCfg { int R, LT, IG; };
static YinCfg yin_cfg(int tauMax) {
    YinCfg best = {7, 64, 4};
    int bestWaste = 1 << 30;
    const int Rs[6] = {5, 7, 9, 11, 13, 15};
    for (int i = 0; i < 6; ++i) {
        const int R = Rs[i];
        int LT = ((tauMax + R - 1) / R + 31) / 32 * 32;
        const int waste = LT * R - tauMax;
        if (LT <= 128 && waste <= bestWaste) { bestWaste = waste; best.R = R; best.LT = LT; }
    }
    best.IG = 256 / best.LT;
    if (best.IG < 1) best.IG = 1;
    return best;
}

template <typename T>
static void yin_dispatch(cudaStream_t st, const VPGeom& g, int S, const float* voice, const uint8_t* gate, int* period,
                         uint32_t* yflags, int* list, int* listCount, int maxList, bool fromList) {
    const YinCfg c = yin_cfg(g.tauMax);
    const int lagPad = c.LT * c.R;
    const int xsLen = ((g.L + c.R * 2 + lagPad + c.R + 8) + 3) & ~3;
    const size_t smem = (size_t)lagPad * 8 + (size_t)xsLen * 4 + (size_t)c.IG * lagPad * sizeof(T);
    const int threads = c.LT * c.IG;
    dim3 grid = fromList ? dim3(148 * 4) : dim3(g.nFramesP, S);
#define YIN_CASE(RR)                                                                                               \
    case RR:                                                                                                       \
        cudaFuncSetAttribute(k_yin<RR, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);               \
        k_yin<RR, T><<<grid, threads, smem, st>>>(g, voice, gate, period, yflags, list, listCount, maxList, c.LT, \
                                                  c.IG, xsLen, fromList);                                          \
        break;
    switch (c.R) {
        YIN_CASE(5) YIN_CASE(7) YIN_CASE(9) YIN_CASE(11) YIN_CASE(13) YIN_CASE(15)
    }
#undef YIN_CASE
}

void vp_launch_yin(cudaStream_t st, const VPGeom& g, int S, const float* voice, const uint8_t* gate, int* period,
                   uint32_t* yflags, int* recheckList, int* recheckCount, int maxList) {
    yin_dispatch<float>(st, g, S, voice, gate, period, yflags, recheckList, recheckCount, maxList, false);
}
void vp_launch_yin_recheck(cudaStream_t st, const VPGeom& g, int S, const float* voice, const uint8_t* gate, int* period,
                           uint32_t* yflags, const int* recheckList, const int* recheckCount, int maxList) {
    yin_dispatch<double>(st, g, S, voice, gate, period, yflags, const_cast<int*>(recheckList),
                         const_cast<int*>(recheckCount), maxList, true);
}

// ===========================================================================
// YIN, correlation form (default path). d[k] = A + B(k) - 2 C(k) with
//   A = sum_{i<L} x[q+i]^2,  B(k) = sum_{i<L} x[q+k+i]^2,  C(k) = sum_{i<L} x[q+i] x[q+k+i]
// -- one FFMA per (i, k) pair instead of FADD + FFMA, and because hop = 3c and
// L = 4c (PluginProcessor.cpp:168-170) C(k) splits into four chunk partials
//   Pm(k) = sum_{j in chunk m} x[j] x[j+k],  chunk m = [m c - tauMax, (m+1) c - tauMax)
// of which frame f uses m = 3f .. 3f+3 and shares the last with frame f+1:
// 0.75 L tauMax FFMAs per frame instead of 2 L tauMax lane-ops.
//   k_yin_corr   : FP32, warp <-> chunk, lane <-> 15 consecutive lags (odd stride: conflict-free
//                  window loads; the a[n] load is a broadcast), 15-deep register window.
//   k_yin_decide : one warp per frame, FP64: energies by running sums, CMND, threshold search and
//                  descent, each deciding comparison checked against a rigorous bound on the FP32
//                  accumulation error of C(k) (|err d[k]| <= c 2^-24 (A + B(k))); a frame with any
//                  comparison inside its bound goes to the FP64 direct-form kernel above.
// ===========================================================================
#define YC_R 15
#define YC_LAGS (32 * YC_R)  // lags per warp pass
#define YC_CH 8              // chunks (= warps) per CTA

__global__ void __launch_bounds__(32 * YC_CH) k_yin_corr(VPGeom g, const float* __restrict__ voice, float* __restrict__ P,
                                                          int nChunks, int lagPad) {
    extern __shared__ float xs[];  // [YC_CH * c + lagPad + 16]
    const int c = g.c, tauMax = g.tauMax;
    const int s = blockIdx.y;
    const int m0 = blockIdx.x * YC_CH;
    const float* v = voice + (size_t)s * g.stride;
    const long long u0 = (long long)m0 * c - tauMax;  // delayed position of xs[0]
    const int span = YC_CH * c + lagPad + 16;
    for (int j = threadIdx.x; j < span; j += blockDim.x) xs[j] = vp_x(v, u0 + j, g.lat, g.n);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = m0 + warp;
    if (m >= nChunks) return;
    const float* xa = xs + warp * c;
    float* out = P + ((size_t)s * nChunks + m) * (size_t)lagPad;
    for (int k0 = lane * YC_R; k0 < lagPad; k0 += YC_LAGS) {
        const float* xw = xa + k0;
        float acc[YC_R], W[YC_R];
#pragma unroll
        for (int r = 0; r < YC_R; ++r) { acc[r] = 0.0f; W[r] = xw[r]; }
        int n = 0;
        for (; n + YC_R <= c; n += YC_R) {
#pragma unroll
            for (int u = 0; u < YC_R; ++u) {
                const float a = xa[n + u];
#pragma unroll
                for (int r = 0; r < YC_R; ++r) acc[r] = fmaf(a, W[(u + r) % YC_R], acc[r]);
                W[u % YC_R] = xw[n + u + YC_R];
            }
        }
#pragma unroll
        for (int u = 0; u < YC_R; ++u) {  // tail: fewer than YC_R samples left
            const float a = (n + u < c) ? xa[n + u] : 0.0f;
#pragma unroll
            for (int r = 0; r < YC_R; ++r) acc[r] = fmaf(a, W[(u + r) % YC_R], acc[r]);
            W[u % YC_R] = xw[n + u + YC_R];
        }
#pragma unroll
        for (int r = 0; r < YC_R; ++r) out[k0 + r] = acc[r];
    }
}

#define YD_WARPS 4

__global__ void __launch_bounds__(32 * YD_WARPS) k_yin_decide(VPGeom g, const float* __restrict__ voice,
                                                              const uint8_t* __restrict__ gate, const float* __restrict__ P,
                                                              int nChunks, int lagPad, int tauPad, int S,
                                                              int* __restrict__ period, uint32_t* __restrict__ yflags,
                                                              int* __restrict__ list, int* __restrict__ listCount, int maxList) {
    extern __shared__ double smd[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long fidx = (long long)blockIdx.x * YD_WARPS + warp;
    if (fidx >= (long long)S * g.nFramesP) return;
    double* dp = smd + (size_t)warp * 2 * tauPad;  // d then d'
    double* er = dp + tauPad;                      // absolute error bound of d'[k]
    float* xlo = (float*)(smd + (size_t)YD_WARPS * 2 * tauPad) + (size_t)warp * 2 * tauPad;  // x[q + i], i < tauMax
    float* xhi = xlo + tauPad;                                                               // x[q + L + i]
    const int tauMax = g.tauMax, L = g.L, c = g.c;
    const int s = (int)(fidx / g.nFramesP), f = (int)(fidx - (long long)s * g.nFramesP);
    const long long p = (long long)f * g.hopP;
    const int b = (int)(p / g.B);
    if (gate[(size_t)s * g.nBlocks + b] & VP_GATE_VOICE) {  // gated: yin() is not run (PitchProcess.cpp:208-214)
        if (lane == 0) { period[fidx] = 0; yflags[fidx] = 0; }
        return;
    }
    const float* v = voice + (size_t)s * g.stride;
    const long long q = p - tauMax;
    // ---- A = sum x[q+i]^2 (exact products in double), edge samples for the running B(k)
    double A = 0.0;
    for (int i = lane; i < L; i += 32) {
        const float x = vp_x(v, q + i, g.lat, g.n);
        A = fma((double)x, (double)x, A);
        if (i < tauMax) xlo[i] = x;
    }
    for (int i = lane; i < tauMax; i += 32) xhi[i] = vp_x(v, q + L + i, g.lat, g.n);
    A = vp_warp_sum(A);
    // ---- C(k) = four chunk partials
    const float* P0 = P + ((size_t)s * nChunks + (size_t)3 * f) * (size_t)lagPad;
    for (int k = lane; k < tauMax; k += 32) {
        const double cc = ((double)P0[k] + (double)P0[(size_t)lagPad + k]) + ((double)P0[(size_t)2 * lagPad + k] + (double)P0[(size_t)3 * lagPad + k]);
        dp[k] = cc;
    }
    __syncwarp();
    // ---- lane owns consecutive lags [kA, kB): B(k) = A + sum_{i<k} (xhi[i]^2 - xlo[i]^2)
    const int per = (tauMax + 31) / 32;
    const int kA = lane * per, kB = min(kA + per, tauMax);
    double locDelta = 0.0;
    for (int k = kA; k < kB; ++k) { const double h = xhi[k], l = xlo[k]; locDelta += h * h - l * l; }
    double inc = locDelta;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    double run = inc - locDelta;  // sum of deltas before kA
    const double beta = (double)c * 5.9604644775390625e-08 * 1.25;  // c * 2^-24, 25 % head room
    double locD = 0.0, locE = 0.0;
    for (int k = kA; k < kB; ++k) {
        const double Bk = A + run;
        const double h = xhi[k], l = xlo[k];
        run += h * h - l * l;
        const double d = (k == 0) ? 0.0 : (A + Bk) - 2.0 * dp[k];
        const double e = beta * (A + Bk);
        dp[k] = d;
        er[k] = e;
        if (k >= 1) { locD += d; locE += e; }
    }
    double incD = locD, incE = locE;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incD, o), t2 = __shfl_up_sync(0xffffffffu, incE, o);
        if (lane >= o) { incD += t; incE += t2; }
    }
    double runD = incD - locD, runE = incE - locE;
    const double total = __shfl_sync(0xffffffffu, incD, 31);
    const double energy = __shfl_sync(0xffffffffu, incE, 31);
    bool shaky = false;  // cumulative sum not resolved against its own error bound
    for (int k = max(kA, 1); k < kB; ++k) {
        runD += dp[k];
        runE += er[k];
        const double d = dp[k], e = er[k];
        const double dn = d * ((double)k / runD);  // cumulative mean normalisation (PitchProcess.cpp:396-402)
        double en;
        if (runD > 2.0 * runE) en = e * ((double)k / runD) + fabs(dn) * (runE / runD) * 1.01;
        else { en = 1e300; if (runE > 0.0 && k >= g.tauMin) shaky = true; }
        dp[k] = dn;
        er[k] = en;
    }
    if (lane == 0) { dp[0] = 1.0; er[0] = 0.0; }
    __syncwarp();
    // ---- first tau >= tauMin with d'[tau] < yinTol, then the descent (PitchProcess.cpp:429-440)
    const double tol = 0.25;
    int first = 0x7fffffff;
    for (int k = max(kA, g.tauMin); k < kB; ++k)
        if (dp[k] < tol) { first = k; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    int per_ = 0;
    unsigned fl = 0;
    bool unsafe = false;
    const bool have = (first != 0x7fffffff) && total > 0.0;
    if (energy > 0.0) {  // all-zero input is exact: d = 0, d' = NaN, unvoiced
        const int kLast = have ? first : tauMax - 1;
        for (int k = max(kA, g.tauMin); k < kB && k <= kLast; ++k)
            if (!(fabs(dp[k] - tol) > er[k])) unsafe = true;
        if (shaky) unsafe = true;
    }
    if (have) {
        if (first + 1 >= tauMax) fl |= YF_UB;  // U3: reads yinTemp[tauMax]
        int stop = 0x7fffffff;
        for (int k = max(kA, first); k < kB; ++k)
            if (k + 1 >= tauMax || !(dp[k + 1] < dp[k])) { stop = k; break; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) stop = min(stop, __shfl_xor_sync(0xffffffffu, stop, o));
        per_ = stop;
        for (int k = max(kA, first); k < kB && k <= stop; ++k)
            if (k + 1 < tauMax && !(fabs(dp[k + 1] - dp[k]) > er[k] + er[k + 1])) unsafe = true;
    }
    unsafe = __any_sync(0xffffffffu, unsafe);
    if (lane == 0) {
        if (unsafe) fl |= YF_RECHECK;
        period[fidx] = per_;
        yflags[fidx] = fl;
        if (unsafe) {
            const int slot = atomicAdd(listCount, 1);
            if (slot < maxList) list[slot] = (int)fidx;
        }
    }
}

int vp_yin_corr_lagpad(const VPGeom& g) { return (g.tauMax + YC_LAGS - 1) / YC_LAGS * YC_LAGS; }
int vp_yin_corr_chunks(const VPGeom& g) { return 3 * g.nFramesP + 1; }

void vp_launch_yin_corr(cudaStream_t st, const VPGeom& g, int S, const float* voice, float* P) {
    const int lagPad = vp_yin_corr_lagpad(g), nChunks = vp_yin_corr_chunks(g);
    const size_t smem = (size_t)(YC_CH * g.c + lagPad + 16) * sizeof(float);
    cudaFuncSetAttribute(k_yin_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid((nChunks + YC_CH - 1) / YC_CH, S);
    k_yin_corr<<<grid, 32 * YC_CH, smem, st>>>(g, voice, P, nChunks, lagPad);
}

void vp_launch_yin_decide(cudaStream_t st, const VPGeom& g, int S, const float* voice, const uint8_t* gate, const float* P,
                          int* period, uint32_t* yflags, int* recheckList, int* recheckCount, int maxList) {
    const int lagPad = vp_yin_corr_lagpad(g), nChunks = vp_yin_corr_chunks(g);
    const int tauPad = (g.tauMax + 3) & ~3;
    const size_t smem = (size_t)YD_WARPS * 2 * tauPad * (sizeof(double) + sizeof(float));
    cudaFuncSetAttribute(k_yin_decide, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const long long tot = (long long)S * g.nFramesP;
    k_yin_decide<<<(unsigned)((tot + YD_WARPS - 1) / YD_WARPS), 32 * YD_WARPS, smem, st>>>(
        g, voice, gate, P, nChunks, lagPad, tauPad, S, period, yflags, recheckList, recheckCount, maxList);
}

// ===========================================================================
// Pitch-mark chain (PitchProcess.cpp:455-658): sequential over the frames of a
// stream, so one warp per stream. The four mark vectors are modelled as
// storage-slot arrays, one slot per lane: push_back writes slot[size],
// insert(begin) shifts right, clear() keeps the values (the reference reads
// slot[size] at PitchProcess.cpp:818).
// ===========================================================================
struct ArgMin { float v; int i; };

__device__ __forceinline__ int marks_argmin(const float* __restrict__ v, long long p, int i0, int i1, int lat, long long n,
                                            int lane) {
    // first index of the minimum over [i0, i1) (strict '<', PitchProcess.cpp:752-764); empty range -> i0
    float bv = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int i = i0 + lane; i < i1; i += 32) {
        const float x = vp_x(v, p + i, lat, n);
        if (x < bv) { bv = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    return (bi == 0x7fffffff) ? i0 : bi;
}

__global__ void __launch_bounds__(128) k_marks(VPGeom g, VPTables tb, const float* __restrict__ voice,
                                               const uint8_t* __restrict__ gate, const int* __restrict__ periodArr,
                                               const uint32_t* __restrict__ yflags, vp_pitch_frame* __restrict__ frames,
                                               int S) {
    const int lane = threadIdx.x & 31;
    const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (s >= S) return;
    const float* v = voice + (size_t)s * g.stride;
    const uint8_t* gt = gate + (size_t)s * g.nBlocks;
    const int L = g.L, hop = g.hopP, cap = g.anCap;
    // PitchProcess state (PitchProcess.cpp:76-92)
    int period = 0, prevPeriod = 0, prevVoicedPeriod = 0, periodNew = 0;
    bool voiced = false, prevVoiced = false;
    double beta = 1.0;
    int an = 0, pan = 0, st = 0, pst = 0;  // this lane's storage slot of each vector
    int nAn = 0, nPan = 0, nSt = 0, nPst = 0;
#define SLOT(arr, i) __shfl_sync(0xffffffffu, (arr), (i))
#define AN_PUSH(val)                                   \
    do {                                               \
        if (nAn >= cap) ub = true;                     \
        if (nAn < VP_SLOTS - 1) { if (lane == nAn) an = (val); ++nAn; } \
    } while (0)
#define ST_PUSH(val)                                   \
    do {                                               \
        if (nSt >= cap) ub = true;                     \
        if (nSt < VP_SLOTS - 1) { if (lane == nSt) st = (val); ++nSt; } \
    } while (0)
    for (int f = 0; f < g.nFramesP; ++f) {
        const long long p = (long long)f * hop;
        const int b = (int)(p / g.B);
        const size_t fidx = (size_t)s * g.nFramesP + f;
        vp_pitch_frame* rec = frames + fidx;
        unsigned flags = 0;
        bool ub = false;
        int note = -1, nAnOv = 0;
        const uint8_t gb = gt[b];
        if (gb & VP_GATE_NEAR) flags |= VP_PF_NEAR_GATE;
        if (gb & VP_GATE_VOICE) {
            // PitchProcess.cpp:208-214: anMarks.clear(); prevPitch = 0; return (pitch, stMarks stay stale)
            nAn = 0;
            flags |= VP_PF_GATED;
        } else {
            // ---- yin() state roll (PitchProcess.cpp:415-425)
            prevPeriod = period;
            prevVoiced = voiced;
            if (voiced) prevVoicedPeriod = period;
            period = periodArr[fidx];
            voiced = period > 0;
            const uint32_t yf = yflags[fidx];
            if (yf & YF_NEAR) flags |= VP_PF_NEAR_YIN;
            if (yf & YF_UB) ub = true;
            if (yf & YF_DONE64) flags |= VP_PF_YIN_RECHECKED;
            // ---- pitchMarks() (PitchProcess.cpp:455-567)
            pan = an; nPan = nAn;  // prevAnMarks = anMarks (element copy)
            nAn = 0;               // clear(): slots keep their values
            pan -= hop;
            nAnOv = __popc(__ballot_sync(0xffffffffu, lane < nPan && pan >= 0));
            if (voiced) {
                const int sw_c = (int)floor(0.94 * period);
                const int sw_f = (int)ceil((2.0 - 0.94) * period);
                bool searchLeft = false;
                int t;
                if (prevVoiced) {
                    if (nAnOv == 0) {
                        int lastMark = 0;
                        if (nPan == 0) ub = true;  // U4: prevAnMarks.back() on an empty vector
                        else lastMark = SLOT(pan, nPan - 1);
                        const int mn = min(prevPeriod, period), mx = max(prevPeriod, period);
                        const int l_lim = max(lastMark + min(sw_c, (int)floor(0.94 * mn)), 0);
                        const int r_lim = min(lastMark + max(sw_f, (int)ceil((2 - 0.94) * mx)), L);
                        t = marks_argmin(v, p, l_lim, r_lim, g.lat, g.n, lane);
                    } else {
                        t = SLOT(pan, nPan - nAnOv);
                    }
                } else {
                    searchLeft = true;
                    t = marks_argmin(v, p, 0, L, g.lat, g.n, lane);
                }
                AN_PUSH(t);
                for (;;) {
                    const int bk = SLOT(an, nAn - 1);
                    if (!(bk + sw_c < L)) break;
                    if (bk + sw_f < L) {
                        const int m = marks_argmin(v, p, bk + sw_c, bk + sw_f, g.lat, g.n, lane);
                        AN_PUSH(m);
                    } else {
                        if (bk + period < L) {
                            const int m = marks_argmin(v, p, bk + sw_c, L, g.lat, g.n, lane);
                            AN_PUSH(m);
                        }
                        break;
                    }
                    if (nAn >= VP_SLOTS - 1) break;
                }
                if (searchLeft) {
                    for (;;) {
                        const int fr = SLOT(an, 0);
                        if (!(fr - sw_c > 0)) break;
                        int m;
                        bool last = false;
                        if (fr - sw_f >= 0) m = marks_argmin(v, p, fr - sw_f, fr - sw_c, g.lat, g.n, lane);
                        else if (fr - period >= 0) { m = marks_argmin(v, p, 0, fr - sw_c, g.lat, g.n, lane); last = true; }
                        else break;
                        // insert(begin): slots [0, nAn) move one to the right
                        if (nAn >= cap) ub = true;
                        if (nAn >= VP_SLOTS - 1) break;
                        const int up = __shfl_up_sync(0xffffffffu, an, 1);
                        if (lane >= 1 && lane <= nAn) an = up;
                        if (lane == 0) an = m;
                        ++nAn;
                        if (last) break;
                    }
                }
            } else if (nPan > 0) {
                if (nAnOv > 0) {
                    for (int i = 0; i < nAnOv; ++i) { const int m = SLOT(pan, nPan - nAnOv + i); AN_PUSH(m); }
                } else {
                    const int m = SLOT(pan, nPan - 1) + prevVoicedPeriod;
                    AN_PUSH(m);
                }
                if (prevVoicedPeriod <= 0) ub = true;
                else
                    for (;;) {
                        const int bk = SLOT(an, nAn - 1);
                        if (!(bk + prevVoicedPeriod < L) || nAn >= VP_SLOTS - 1) break;
                        AN_PUSH(bk + prevVoicedPeriod);
                    }
            }
            // ---- placeStMarks() (PitchProcess.cpp:573-658)
            pst = st; nPst = nSt;
            nSt = 0;
            pst -= hop;
            if (nAn > 0) {
                const int nStOv = __popc(__ballot_sync(0xffffffffu, lane < nPst && pst >= 0));
                if (voiced) {
                    beta = tb.lutBeta[period];
                    periodNew = tb.lutPeriodNew[period];
                    note = tb.lutNote[period];
                } else {
                    periodNew = prevVoicedPeriod;
                }
                bool place = true;
                int firstMark = 0;
                if (periodNew <= 0) { ub = true; place = false; }
                else if (voiced) {
                    if (prevVoiced) {
                        if (nStOv > 0) firstMark = SLOT(pst, nPst - nStOv);
                        else if (nPst == 0) { ub = true; firstMark = SLOT(an, 0); }
                        else {
                            const int bk = SLOT(pst, nPst - 1);
                            firstMark = (bk + periodNew >= 0) ? bk + periodNew : SLOT(an, 0);
                        }
                    } else firstMark = SLOT(an, 0);
                } else {
                    if (nPst == 0) place = false;
                    else if (nStOv > 0) firstMark = SLOT(pst, nPst - nStOv);
                    else {
                        const int bk = SLOT(pst, nPst - 1);
                        int nn = 1;
                        while (bk + nn * periodNew < 0) nn += 1;
                        firstMark = bk + nn * periodNew;
                    }
                }
                if (place) {
                    ST_PUSH(firstMark);
                    int bk = firstMark;
                    while (bk + periodNew < L && nSt < VP_SLOTS - 1) { bk += periodNew; ST_PUSH(bk); }
                }
            }
            if (voiced) flags |= VP_PF_VOICED;
            if (nAn > 0) flags |= VP_PF_HAS_MARKS;
        }
        if (ub) flags |= VP_PF_UB;
        // ---- record
        const int stale = (nAn < cap) ? SLOT(an, nAn) : 0;
        if (lane < VP_MAX_MARKS) {
            rec->anMarks[lane] = (lane < nAn) ? an : 0;
            rec->stMarks[lane] = (lane < nSt) ? st : 0;
        }
        if (lane == 0) {
            rec->flags = flags;
            rec->period = (flags & VP_PF_GATED) ? period : period;
            rec->periodPsola = voiced ? period : prevVoicedPeriod;
            rec->periodNew = periodNew;
            rec->note = ((flags & VP_PF_VOICED) && nAn > 0) ? note : -1;
            rec->nAn = nAn;
            rec->nSt = nSt;
            rec->anStale = stale;
            rec->nAnOv = nAnOv;
            rec->beta = beta;
        }
    }
#undef SLOT
#undef AN_PUSH
#undef ST_PUSH
}

void vp_launch_marks(cudaStream_t st, const VPGeom& g, const VPTables& tb, int S, const float* voice,
                     const uint8_t* gate, const int* period, const uint32_t* yflags, vp_pitch_frame* frames) {
    const int threads = 128;
    const long long tot = (long long)S * 32;
    k_marks<<<(unsigned)((tot + threads - 1) / threads), threads, 0, st>>>(g, tb, voice, gate, period, yflags, frames, S);
}

// ===========================================================================
// Per pitch frame with marks: LPC (rectangular window, PitchProcess.cpp:233),
// residual e = A(z) x (PitchProcess.cpp:235, :258-259, :280-302), PSOLA on the
// residual (PitchProcess.cpp:665-741, :788-870) -> the IIR input outE, with the
// reference's chunk-by-chunk visibility made explicit:
// a grain handled while chunk n is current only lands on samples i >= n*c, and
// can only see residual samples up to L + n*c.
// One CTA per (frame, stream), FP64.
// ===========================================================================
#define PF_THREADS 256
#define PF_R 4

template <int R>
__device__ __forceinline__ void pf_ac_task(const double* __restrict__ xd, int n0, int segLen, int m0, int L, double* acc) {
    // rectangular-window autocorrelation partials: sum_{n in seg} x[n] x[n+m], x = 0 beyond L (padded)
    double W[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { acc[j] = 0.0; W[j] = xd[n0 + m0 + j]; }
    const int nEnd = n0 + segLen;
    for (int nb = n0; nb < nEnd; nb += R) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int n = nb + u;
            if (u > 0) W[(u + R - 1) % R] = xd[n + m0 + R - 1];
            const double a = (n < nEnd) ? xd[n] : 0.0;
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fma(a, W[(u + j) % R], acc[j]);
        }
        W[(R - 1) % R] = xd[nb + R + m0 + R - 1];
    }
}

__global__ void __launch_bounds__(PF_THREADS) k_pitch_frame(VPGeom g, VPTables tb, const float* __restrict__ voice,
                                                            vp_pitch_frame* __restrict__ frames,
                                                            double* __restrict__ aP, double* __restrict__ outE,
                                                            int segLen, int xdLen, int eLen) {
    extern __shared__ double smd[];
    const int f = blockIdx.x, s = blockIdx.y;
    const size_t fidx = (size_t)s * g.nFramesP + f;
    vp_pitch_frame* rec = frames + fidx;
    const unsigned flags = rec->flags;
    if ((flags & VP_PF_GATED) || !(flags & VP_PF_HAS_MARKS)) return;
    const int L = g.L, c = g.c, P = g.ordP, tauMax = g.tauMax;
    double* xd = smd;                 // [xdLen]  frame samples (double), zero padded   -- reused as outE later
    double* e = xd + xdLen;           // [eLen]   residual, index idx + tauMax, idx in [-tauMax, L + 3c)
    double* r = e + eLen;             // [P + 1]
    double* a = r + (VP_ORDER_MAX + 1);  // [P + 1]
    __shared__ int sAn[VP_MAX_MARKS + 1], sSt[VP_MAX_MARKS];
    const float* v = voice + (size_t)s * g.stride;
    const long long p = (long long)f * g.hopP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nWarps = PF_THREADS / 32;

    for (int j = tid; j < xdLen; j += PF_THREADS) xd[j] = (j < L) ? (double)vp_x(v, p + j, g.lat, g.n) : 0.0;
    if (tid < VP_MAX_MARKS) { sAn[tid] = rec->anMarks[tid]; sSt[tid] = rec->stMarks[tid]; }
    if (tid == 0) sAn[VP_MAX_MARKS] = 0;
    __syncthreads();
    // ---- autocorrelation: warp = lag group (R lags), lane = segment
    const int G = (P + 1 + PF_R - 1) / PF_R;
    for (int grp = warp; grp < G; grp += nWarps) {
        double acc[PF_R];
        pf_ac_task<PF_R>(xd, lane * segLen, segLen, grp * PF_R, L, acc);
#pragma unroll
        for (int j = 0; j < PF_R; ++j) {
            const double t = vp_warp_sum(acc[j]);
            const int m = grp * PF_R + j;
            if (lane == 0 && m <= P) r[m] = t / (double)L;
        }
    }
    __syncthreads();
    // ---- Levinson-Durbin (LPC.cpp:107-148), serial: thread 0
    if (tid == 0) {
        a[0] = 1.0;
        if (fabs(r[0]) < 1e-9) {
            for (int i = 1; i <= P; ++i) a[i] = 0.0;
        } else {
            a[1] = r[1] / r[0];
            for (int q = 2; q <= P; ++q) {
                double rho = 0.0, ra = 0.0;
                for (int i = 1; i < q; ++i) { rho = fma(r[q - i], a[i], rho); ra = fma(r[i], a[i], ra); }
                const double k = (r[q] - rho) / (r[0] - ra);
                for (int i = 1; 2 * i <= q; ++i) {
                    const double t1 = a[i], t2 = a[q - i];
                    a[i] = fma(-k, t2, t1);
                    if (i != q - i) a[q - i] = fma(-k, t1, t2);
                }
                a[q] = k;
            }
            for (int i = 1; i <= P; ++i) a[i] = -a[i];
        }
        double* ap = aP + fidx * (size_t)(P + 1);
        for (int i = 0; i <= P; ++i) ap[i] = a[i];
    }
    __syncthreads();
    // ---- residual over frame-relative idx in [-tauMax, L + 3c) (full taps; App. A.4 #4)
    for (int j = tid; j < eLen; j += PF_THREADS) {
        const long long u = p + (j - tauMax);
        double acc = 0.0;
        for (int k = 0; k <= P; ++k) acc = fma(a[k], (double)vp_x(v, u - k, g.lat, g.n), acc);
        e[j] = acc;
    }
    double* oE = xd;  // reuse
    __syncthreads();
    for (int i = tid; i < L; i += PF_THREADS) oE[i] = 0.0;
    __syncthreads();
    // ---- PSOLA, chunk by chunk
    const int T = rec->periodPsola, nSt = rec->nSt, nAn = rec->nAn, nAnOv = rec->nAnOv;
    const double beta = rec->beta;
    const int stale = rec->anStale;
    bool ub = false;
    if (T > 0 && T < tauMax) {
        const double* hann = tb.hann + tb.hannOff[T];
        int stIdx = 0;
        for (int n = 0; n < 4; ++n) {
            const long long Pn = p + (long long)n * c;
            if (Pn >= g.n) break;
            const int startSample = (int)(Pn % g.B);
            const int lookahead = g.lat + g.B - startSample;  // bufferIdxMax - startSample (PitchProcess.cpp:800)
            const int eValid = L + n * c;                     // residual filtered so far
            while (stIdx < nSt) {
                const int stMark = sSt[stIdx];
                if (stMark - T >= (n + 1) * c) break;
                // getClosestAnMarkIdx (PitchProcess.cpp:788-831)
                int lo = 0, hi = nAn;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (sAn[mid] < stMark) lo = mid + 1; else hi = mid; }
                const int idx = lo, nc = n * c;
                int cl;
                if (idx > 0 && idx < nAn) {
                    if (abs(sAn[idx] - stMark) <= abs(sAn[idx - 1] - stMark) && sAn[idx] + T - nc < lookahead) cl = idx;
                    else if (sAn[idx - 1] + T - nc < lookahead) cl = idx - 1;
                    else if (idx - 2 > 0) cl = idx - 2;
                    else cl = -nAnOv - 1;
                } else if (idx == 0) cl = 0;
                else {
                    if (stale + T - nc < lookahead) cl = idx - 1;  // anMarks[size]: stale storage slot (U1)
                    else if (idx - 2 >= 0) cl = idx - 2;
                    else { cl = 0; ub = true; }
                }
                int clAn;
                if (cl >= 0) clAn = sAn[cl];
                else { clAn = 0; ub = true; }  // U2: out-of-bounds prevAnMarks read in the reference
                const bool first = (stIdx == 0), last = (stIdx == nSt - 1);
                const double x0 = (double)stMark + (double)(-T) / beta;
                const double xEnd = (double)stMark + (double)(T) / beta;
                const int startIdx = max((int)floor(x0), 0);
                const int stopIdx = min((int)ceil(xEnd), L);
                if (x0 >= 0.0 && x0 == floor(x0)) ub = true;  // U5
                const int eBase = clAn - T + tauMax;           // e index of grain sample j = 0
                // thread <-> output index i is fixed (i mod PF_THREADS) so that successive grains
                // accumulate into oE[i] in mark order without synchronisation
                for (int i = (startIdx / PF_THREADS) * PF_THREADS + tid; i < stopIdx; i += PF_THREADS) {
                    const double di = (double)i;
                    if (i < startIdx || !(di >= x0 && di <= xEnd)) continue;
                    int j = (int)ceil((double)T + (di - (double)stMark) * beta);
                    j = max(0, min(j, 2 * T));
                    // lower_bound on x[j] = stMark + (j - T) / beta (PitchProcess.cpp:850-853)
                    while (j > 0 && (double)stMark + (double)(j - 1 - T) / beta >= di) --j;
                    while (j < 2 * T && (double)stMark + (double)(j - T) / beta < di) ++j;
                    double val;
                    {
                        const int ej = eBase + j;
                        const int rel = clAn - T + j;  // frame-relative index of grain sample j
                        double y1 = (ej >= 0 && ej < eLen && rel < eValid) ? e[ej] : 0.0;
                        const bool w1 = (!first && !last) || (first ? (j >= T) : (j < T));
                        if (w1) y1 *= hann[j];
                        if (j > 0) {
                            double y0 = (ej - 1 >= 0 && ej - 1 < eLen && rel - 1 < eValid) ? e[ej - 1] : 0.0;
                            const bool w0 = (!first && !last) || (first ? (j - 1 >= T) : (j - 1 < T));
                            if (w0) y0 *= hann[j - 1];
                            const double xa = (double)stMark + (double)(j - 1 - T) / beta;
                            const double xb = (double)stMark + (double)(j - T) / beta;
                            val = y0 + (y1 - y0) / (xb - xa) * (di - xa);
                        } else val = y1;
                    }
                    if (i >= nc) oE[i] += val;  // earlier chunks were already filtered (App. A.4 #6)
                }
                ++stIdx;
            }
        }
    } else ub = true;
    __syncthreads();
    double* dst = outE + fidx * (size_t)L;
    for (int i = tid; i < L; i += PF_THREADS) dst[i] = oE[i];
    if (ub && tid == 0) rec->flags = flags | VP_PF_UB;
}

void vp_launch_pitch_frame(cudaStream_t st, const VPGeom& g, const VPTables& tb, int S, const float* voice,
                           vp_pitch_frame* frames, double* aP, double* outE) {
    int segLen = (g.L + 31) / 32;
    if ((segLen & 1) == 0) ++segLen;  // odd -> conflict-free 64-bit loads across the 32 segments
    const int G = (g.ordP + 1 + PF_R - 1) / PF_R;
    int xdLen = 32 * segLen + G * PF_R + 2 * PF_R + 2;
    if (xdLen < g.L) xdLen = g.L;
    const int eLen = g.tauMax + g.L + 3 * g.c;
    const size_t smem = ((size_t)xdLen + eLen + 2 * (VP_ORDER_MAX + 1)) * sizeof(double);
    cudaFuncSetAttribute(k_pitch_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(g.nFramesP, S);
    k_pitch_frame<<<grid, PF_THREADS, smem, st>>>(g, tb, voice, frames, aP, outE, segLen, xdLen, eLen);
}

// ===========================================================================
// All-pole resynthesis of the PSOLA'd residual (PitchProcess.cpp:307-322),
// synthesis window and overlap-add (PitchProcess.cpp:328-342). The recursion is
// serial in i and restarts at every frame, so one thread per frame; a warp
// moves 32-sample slabs of its 32 frames through shared memory so that global
// reads of outE and writes of the output are coalesced 128/256-byte rows.
// Chunks 1..2 are private to a frame (plain stores); chunk 0 / chunk 3 overlap
// the neighbouring frame (exactly two contributors -> deterministic float
// atomics on a zeroed buffer).
// ===========================================================================
#define PI_WARPS 2

template <int P>
__global__ void __launch_bounds__(32 * PI_WARPS) k_pitch_iir(VPGeom g, VPTables tb, const vp_pitch_frame* __restrict__ frames,
                                                             const double* __restrict__ aP, const double* __restrict__ outE,
                                                             float* __restrict__ outP, long long nFramesTot) {
    __shared__ double tin[PI_WARPS][32][33];
    __shared__ float tout[PI_WARPS][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long f0 = ((long long)blockIdx.x * PI_WARPS + warp) * 32;
    if (f0 >= nFramesTot) return;
    const long long fidx = f0 + lane;
    const int L = g.L, c = g.c;
    const int order = (P > 0) ? P : g.ordP;
    bool active = false;
    int s = 0, f = 0, nSteps = 0;
    if (fidx < nFramesTot) {
        s = (int)(fidx / g.nFramesP);
        f = (int)(fidx - (long long)s * g.nFramesP);
        const unsigned fl = frames[fidx].flags;
        active = !(fl & VP_PF_GATED) && (fl & VP_PF_HAS_MARKS);
        const long long p = (long long)f * g.hopP;
        for (int n = 0; n < 4; ++n) if (p + (long long)n * c < g.n) nSteps += c;  // chunks processed in this run
    }
    if (!active) nSteps = 0;
    constexpr int PA = (P > 0) ? P : VP_ORDER_MAX;
    double a[PA + 1], h[PA];
    for (int k = 0; k <= PA; ++k) a[k] = 0.0;
    for (int k = 0; k < PA; ++k) h[k] = 0.0;
    if (active) {
        const double* ap = aP + (size_t)fidx * (order + 1);
        for (int k = 0; k <= order; ++k) a[k] = ap[k];
    }
    // per-lane metadata shared through shuffles for the cooperative slab moves
    const long long myP = (long long)f * g.hopP;
    const double gp = (double)g.gainPitchF;
    const int nSlabs = (L + 31) / 32;
    int maxSteps = nSteps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxSteps = max(maxSteps, __shfl_xor_sync(0xffffffffu, maxSteps, o));
    for (int sl = 0; sl < nSlabs; ++sl) {
        const int i0 = sl * 32;
        if (i0 >= maxSteps) break;
        // load: row fr of the tile <- outE[frame f0+fr][i0 .. i0+32)
        for (int fr = 0; fr < 32; ++fr) {
            const int steps = __shfl_sync(0xffffffffu, nSteps, fr);
            double val = 0.0;
            if (i0 + lane < steps) val = outE[(size_t)(f0 + fr) * L + i0 + lane];
            tin[warp][fr][lane] = val;
        }
        __syncwarp();
        if (P > 0) {
            // statically indexed circular history: slab length 32 and order P are unrolled jointly below
        }
        for (int j = 0; j < 32; ++j) {
            const int i = i0 + j;
            double acc = tin[warp][lane][j];
            if (P > 0) {
#pragma unroll
                for (int k = 1; k <= PA; ++k) acc = fma(-a[k], h[k - 1], acc);
#pragma unroll
                for (int k = PA - 1; k > 0; --k) h[k] = h[k - 1];
                h[0] = acc;
            } else {
                for (int k = 1; k <= order && k <= i; ++k) acc = fma(-a[k], h[(i - k) % order], acc);
                h[i % order] = acc;
            }
            const double w = (i < L) ? tb.stP[i] : 0.0;
            tout[warp][lane][j] = (float)(acc * w * gp);
        }
        __syncwarp();
        for (int fr = 0; fr < 32; ++fr) {
            const int steps = __shfl_sync(0xffffffffu, nSteps, fr);
            const long long pf = __shfl_sync(0xffffffffu, myP, fr);
            const int sf = __shfl_sync(0xffffffffu, s, fr);
            const int i = i0 + lane;
            if (i < steps) {
                const long long u = pf + i;
                if (u < g.n) {
                    float* o = outP + (size_t)sf * g.wstride + u;
                    const float val = tout[warp][fr][lane];
                    if (i < c || i >= 3 * c) atomicAdd(o, val);
                    else *o = val;
                }
            }
        }
        __syncwarp();
    }
}

void vp_launch_pitch_iir(cudaStream_t st, const VPGeom& g, const VPTables& tb, int S, const vp_pitch_frame* frames,
                         const double* aP, const double* outE, float* outP) {
    const long long tot = (long long)S * g.nFramesP;
    const unsigned grid = (unsigned)((tot + 32 * PI_WARPS - 1) / (32 * PI_WARPS));
    if (g.ordP == 15) k_pitch_iir<15><<<grid, 32 * PI_WARPS, 0, st>>>(g, tb, frames, aP, outE, outP, tot);
    else k_pitch_iir<0><<<grid, 32 * PI_WARPS, 0, st>>>(g, tb, frames, aP, outE, outP, tot);
}
