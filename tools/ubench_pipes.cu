// Developer microbenchmark (not part of the product): FP64 DFMA and FP32 FFMA / FFMA2 issue rate on one GPU as a
// function of resident warps per scheduler and independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dfma(double* sink, int iters) {
    double a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x + i; b[i] = 1e-3 * (i + 1 + (threadIdx.x & 3)); }
    double x = 0.999 + 1e-6 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fma(x, b[i], a[i]);
        x = -x;
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += a[i];
    if (r == 123456789.0) sink[0] = r;
}

// transposed-IIR-like: 40 DFMAs all depending on one fresh value per step, that value depending on the first DFMA of the previous step
template <int P>
__global__ void k_iir(double* sink, int iters) {
    double a[P + 1], st[P + 1];
#pragma unroll
    for (int i = 0; i <= P; ++i) { a[i] = 1e-3 * (i + 1 + (threadIdx.x & 3)); st[i] = 0.0; }
    double e = 1e-3 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const double o = e + st[0];
#pragma unroll
        for (int k = 0; k < P; ++k) st[k] = fma(-a[k + 1], o, st[k + 1]);
        e = -e;
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i <= P; ++i) r += st[i];
    if (r == 123456789.0) sink[0] = r;
}

// stream-synthesis-like step: FIR(PS) + IIR(P) + output scaling (+ optional shuffles), VAR selects features
template <int P, int PS, int VAR>
__global__ void k_synth(float* sink, int iters) {
    double a[P + 1], st[P + 1], as[PS + 1], t[PS + 1];
#pragma unroll
    for (int i = 0; i <= P; ++i) { a[i] = 1e-3 * (i + 1 + (threadIdx.x & 3)); st[i] = 0.0; }
#pragma unroll
    for (int i = 0; i <= PS; ++i) { as[i] = 1e-2 * (i + 1); t[i] = 0.0; }
    float xs = 1e-3f * threadIdx.x;
    double w = 0.7, gv = 1.0;
    float acc = 0.f;
    const int phi = threadIdx.x & 3;
    for (int it = 0; it < iters; it += 4) {
        double e[4];
        float c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double x = (double)xs * w;
            e[j] = fma(as[0], x, t[0]);
#pragma unroll
            for (int q = 0; q < PS; ++q) t[q] = fma(as[q + 1], x, t[q + 1]);
            xs = -xs;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double ov = e[j] + st[0];
#pragma unroll
            for (int k = 0; k < P; ++k) st[k] = fma(-a[k + 1], ov, st[k + 1]);
            if (VAR >= 1) c[j] = (float)(gv * ov * w); else c[j] = 0.f;
        }
        if (VAR >= 2) {
            const bool hi = (phi & 2) != 0, od = (phi & 1) != 0;
            const float s0 = hi ? c[0] : c[2], s1 = hi ? c[1] : c[3];
            const float k0 = hi ? c[2] : c[0], k1 = hi ? c[3] : c[1];
            const float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
            const float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
            const float snd = od ? r0 : r1, kp = od ? r1 : r0;
            acc += kp + __shfl_xor_sync(0xffffffffu, snd, 1);
        } else if (VAR >= 1) acc += c[0] + c[1] + c[2] + c[3];
    }
    double r = acc;
#pragma unroll
    for (int i = 0; i <= P; ++i) r += st[i];
    if (r == 123456789.0) sink[0] = (float)r;
}

template <int CH>
__global__ void k_ffma(float* sink, int iters) {
    float a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x + i; b[i] = 1e-3f * (i + 1 + (threadIdx.x & 3)); }
    float x = 0.999f + 1e-6f * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = fmaf(x, b[i], a[i]);
        x = -x;
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += a[i];
    if (r == 123456789.0f) sink[0] = r;
}

template <int CH>  // CH float2 chains = 2 CH FMAs per step
__global__ void k_ffma2(float* sink, int iters) {
    float2 a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = make_float2(threadIdx.x + i, i); b[i] = make_float2(1e-3f * (i + 1), 2e-3f * (i + 1 + (threadIdx.x & 3))); }
    float2 x = make_float2(0.999f + 1e-6f * threadIdx.x, 0.999f + 1e-6f * threadIdx.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) a[i] = __ffma2_rn(x, b[i], a[i]);
        x.x = -x.x; x.y = -x.y;
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += a[i].x + a[i].y;
    if (r == 123456789.0f) sink[0] = r;
}

// sliding-window correlation pattern of the engine's autocorrelation / YIN loops: acc[j] += x_u * W[(u + j) % R]
// (ROT = 1: every accumulator meets every window register) versus the same FMAs with a fixed pairing (ROT = 0)
template <typename T, int R, int ROT>
__global__ void k_window(T* sink, int iters) {
    T acc[R], W[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { acc[j] = (T)0; W[j] = (T)1e-3 * (T)(j + 1 + (threadIdx.x & 3)); }
    T x = (T)0.999 + (T)1e-6 * (T)threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
#pragma unroll
            for (int j = 0; j < R; ++j) acc[j] = fma(x, W[ROT ? (u + j) % R : j], acc[j]);
            x = -x;
        }
    }
    T r = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) r += acc[j] + W[j];
    if (r == (T)123456789) sink[0] = r;
}

// the YIN correlation inner loop with its shared-memory traffic: a = xa[n] (broadcast), window refilled from xw[n + R]
// (lane stride R). VEC = 1: the a values come as one 64-bit broadcast load per two steps.
template <int R, int VEC>
__global__ void __launch_bounds__(256, 4) k_yinloop(float* sink, int c, int reps) {
    __shared__ float xs[8 * 288 + 512];
    for (int j = threadIdx.x; j < 8 * 288 + 512; j += blockDim.x) xs[j] = 1e-3f * (float)((j * 7 + blockIdx.x) % 113);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xa = xs + warp * 288;
    const float* xw = xa + lane * R;
    float tot = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        float acc[R], W[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = 0.f; W[r] = xw[r]; }
        for (int n = 0; n + R <= c; n += R) {
            if (VEC && (R % 2 == 1)) {
                // odd R: steps pair up across two bodies; keep it simple: scalar loads for the odd tail
            }
#pragma unroll
            for (int u = 0; u < R; ++u) {
                float a;
                if (VEC && (u % 2 == 0) && u + 1 < R) {
                    const float2 a2 = *reinterpret_cast<const float2*>(xa + ((n + u) & ~1));
                    a = ((n + u) & 1) ? a2.y : a2.x;
                } else a = xa[n + u];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fmaf(a, W[(u + r) % R], acc[r]);
                W[u % R] = xw[n + u + R];
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) tot += acc[r];
    }
    if (tot == 123456789.f) sink[0] = tot;
}

template <typename F>
static double timeit(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    void* sink;
    cudaMalloc(&sink, 64);
    printf("%s, %d SMs, %d kHz\n", p.name, sms, p.clockRate);
    const int iters = 20000;
    const int wps[] = {1, 2, 3, 4, 8};
    for (int wi = 0; wi < 5; ++wi) {
        const int w = wps[wi];
        const int threads = 128 * w;  // w warps per scheduler (4 schedulers), one CTA per SM
        const double n = (double)sms * threads * iters;
        double t;
        t = timeit([&] { k_dfma<8><<<sms, threads>>>((double*)sink, iters); });
        printf("warps/sched %d  DFMA ch8  %6.2f T/s", w, n * 8 / t / 1e12);
        t = timeit([&] { k_dfma<20><<<sms, threads>>>((double*)sink, iters); });
        printf("  ch20 %6.2f", n * 20 / t / 1e12);
        t = timeit([&] { k_dfma<40><<<sms, threads>>>((double*)sink, iters); });
        printf("  ch40 %6.2f", n * 40 / t / 1e12);
        t = timeit([&] { k_iir<40><<<sms, threads>>>((double*)sink, iters); });
        printf("  iir40 %6.2f", n * 40 / t / 1e12);
        t = timeit([&] { k_iir<15><<<sms, threads>>>((double*)sink, iters); });
        printf("  iir15 %6.2f", n * 15 / t / 1e12);
        if (w <= 2) {
            t = timeit([&] { k_synth<40, 5, 0><<<sms, threads>>>((float*)sink, iters); });
            printf("  synth0 %6.2f", n * 46 / t / 1e12);
            t = timeit([&] { k_synth<40, 5, 1><<<sms, threads>>>((float*)sink, iters); });
            printf("  synth1 %6.2f", n * 46 / t / 1e12);
            t = timeit([&] { k_synth<40, 5, 2><<<sms, threads>>>((float*)sink, iters); });
            printf("  synth2 %6.2f", n * 46 / t / 1e12);
        }
        t = timeit([&] { k_ffma<16><<<sms, threads>>>((float*)sink, iters); });
        printf(" | FFMA ch16 %6.2f", n * 16 / t / 1e12);
        t = timeit([&] { k_ffma2<8><<<sms, threads>>>((float*)sink, iters); });
        printf("  FFMA2 ch8x2 %6.2f T FMA/s\n", n * 16 / t / 1e12);
    }
    {
        const int threads = 256, blocks = sms * 4, it2 = 2000;
        const double n = (double)blocks * threads * it2;
        double t;
        t = timeit([&] { k_window<float, 15, 0><<<blocks, threads>>>((float*)sink, it2); });
        printf("window R=15 FP32 fixed pairing    %6.2f T FMA/s\n", n * 225 / t / 1e12);
        t = timeit([&] { k_window<float, 15, 1><<<blocks, threads>>>((float*)sink, it2); });
        printf("window R=15 FP32 rotating pairing %6.2f T FMA/s\n", n * 225 / t / 1e12);
        {
            const int c = 270, reps = 400;
            const double nf = (double)blocks * threads * reps * (c / 15) * 225.0;
            t = timeit([&] { k_yinloop<15, 0><<<blocks, threads>>>((float*)sink, c, reps); });
            printf("YIN loop R=15 with LDS (a broadcast + window refill) %6.2f T FMA/s\n", nf / t / 1e12);
        }
        t = timeit([&] { k_window<double, 14, 0><<<blocks, threads>>>((double*)sink, it2); });
        printf("window R=14 FP64 fixed pairing    %6.2f T FMA/s\n", n * 196 / t / 1e12);
        t = timeit([&] { k_window<double, 14, 1><<<blocks, threads>>>((double*)sink, it2); });
        printf("window R=14 FP64 rotating pairing %6.2f T FMA/s\n", n * 196 / t / 1e12);
    }
    return 0;
}
