// Developer microbenchmark (not part of the product): the inner loop of k_yin_corr (vp_pitch.cu) alone -- R FP32
// accumulators, an R-deep register window, one broadcast load + one strided load per R FFMAs -- at 1..8 resident warps per
// scheduler, in the formulations that were considered: two-level / single-level accumulation, R = 13 / 15 / 17, and the
// window refilled by 64-bit loads. Prints cycles per (sample x R lags) step per warp; the issue bound is R + 2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_yin tools/ubench_yin.cu && ./tools/ubench_yin
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int SUB, bool TWO>
__global__ void __launch_bounds__(1024, 1) k_yin_loop(float* sink, int iters, long long* cyc, int c) {
    extern __shared__ float xs[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) xs[i] = 1e-3f * ((i * 7) % 113);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, lg = lane & 15;
    const float* xa = xs + (warp & 7) * 64 + half * 2576;   // the two half-warps 16 banks apart, like the kernel's sub-spans
    float tot = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float* xw = xa + lg * R + (it & 1) * R * 16;
        float acc[R], acc2[R], W[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = 0.f; acc2[r] = 0.f; W[r] = xw[r]; }
        int n = 0;
        while (n + R <= c) {
            const int nSub = TWO ? min(n + SUB, c) : c;
            for (; n + R <= nSub; n += R) {
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    const float a = xa[n + u];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(a, W[(u + r) % R], acc[r]);
                    W[u % R] = xw[n + u + R];
                }
            }
            if (TWO) {
#pragma unroll
                for (int r = 0; r < R; ++r) { acc2[r] += acc[r]; acc[r] = 0.f; }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) tot += acc2[r] + acc[r];
    }
    const long long t1 = clock64();
    if (tot == 123456789.f) sink[0] = tot;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

// the same sums with the `a` samples fetched two at a time (one 64-bit broadcast load per two steps)
template <int R, int SUB>
__global__ void __launch_bounds__(1024, 1) k_yin_loop_a2(float* sink, int iters, long long* cyc, int c) {
    extern __shared__ float xs[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) xs[i] = 1e-3f * ((i * 7) % 113);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, lg = lane & 15;
    const float* xa = xs + (warp & 7) * 64 + half * 2576;
    float tot = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const float* xw = xa + lg * R + (it & 1) * R * 16;
        float acc[R], acc2[R], W[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = 0.f; acc2[r] = 0.f; W[r] = xw[r]; }
        int n = 0;
        while (n + 2 * R <= c) {
            const int nSub = min(n + SUB, c);
            for (; n + 2 * R <= nSub; n += 2 * R) {
#pragma unroll
                for (int u = 0; u < 2 * R; u += 2) {
                    const float2 a2 = *reinterpret_cast<const float2*>(xa + n + u);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(a2.x, W[(u + r) % R], acc[r]);
                    W[u % R] = xw[n + u + R];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = fmaf(a2.y, W[(u + 1 + r) % R], acc[r]);
                    W[(u + 1) % R] = xw[n + u + 1 + R];
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) { acc2[r] += acc[r]; acc[r] = 0.f; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) tot += acc2[r] + acc[r];
    }
    const long long t1 = clock64();
    if (tot == 123456789.f) sink[0] = tot;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

static float* g_sink;
static long long* g_cyc;
static int g_sms;

template <typename F>
static void run(const char* name, int iters, double steps, int R, F launch) {
    printf("%-44s", name);
    for (int w = 1; w <= 8; ++w) {
        long long best = 1LL << 62;
        for (int rep = 0; rep < 3; ++rep) {
            launch(g_sms, 128 * w);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf(" [%s]", cudaGetErrorString(e)); cudaGetLastError(); best = -1; break; }
            long long c;
            cudaMemcpy(&c, g_cyc, 8, cudaMemcpyDeviceToHost);
            if (c < best) best = c;
        }
        const double per = (double)best / iters / steps / w;  // cycles per step per resident warp of a scheduler
        printf(" w%d %5.2f (%3.0f%%)", w, per, 100.0 * (R + 2) / per);
    }
    printf("\n");
}

#define RUN(NAME, K, R, C) \
    cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4); \
    run(NAME, 400, (double)((C) / (R) * (R)), R, [&](int b, int t) { K<<<b, t, 8192 * 4>>>(g_sink, 400, g_cyc, C); })

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    g_sms = p.multiProcessorCount;
    cudaMalloc(&g_sink, 64);
    cudaMalloc(&g_cyc, 64);
    printf("%s, %d SMs; resident warps per scheduler -> cycles per step (R FFMA + 2 LDS) per warp (issue bound R + 2 = 100%%)\n", p.name, g_sms);
    RUN("R=15 two-level (kernel), c=278", (k_yin_loop<15, 60, true>), 15, 278);
    RUN("R=15 single-level, c=278", (k_yin_loop<15, 60, false>), 15, 278);
    RUN("R=13 two-level, c=278", (k_yin_loop<13, 52, true>), 13, 278);
    RUN("R=17 two-level, c=278", (k_yin_loop<17, 68, true>), 17, 278);
    RUN("R=17 single-level, c=278", (k_yin_loop<17, 68, false>), 17, 278);
    RUN("R=15 two-level, a by 64-bit loads, c=270", (k_yin_loop_a2<15, 60>), 15, 270);
    return 0;
}
