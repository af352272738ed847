// Developer microbenchmark (not part of the product): what a non-DFMA instruction costs next to DFMAs on one SM
// sub-partition, and what the float <-> double conversions cost. Answers the design questions of the FP64 vocoder kernels:
//   * does an FFMA / ALU op / LDS issue "for free" in the second cycle of a DFMA, or does it add a cycle?
//   * which pipe do F2F.F64.F32 / F2F.F32.F64 occupy, and at what rate?
//   * the synthesis step (order-40 transposed IIR + order-5 FIR + window + 4-lane overlap-add) in four formulations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mix tools/ubench_mix.cu && ./ubench_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// ND DFMAs + NF FFMAs + NA ALU ops + NL shared loads + NU float->double + NW double->float per iteration, all independent
template <int ND, int NF, int NA, int NL, int NU, int NW>
__global__ void k_mix(float* sink, int iters, long long* cyc, const double* __restrict__ coef) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1e-3f * i;
    __syncthreads();
    double d[ND > 0 ? ND : 1], db[ND > 0 ? ND : 1];
    float f[NF > 0 ? NF : 1], fb[NF > 0 ? NF : 1];
    unsigned u[NA > 0 ? NA : 1];
    float cu[NU > 0 ? NU : 1];
    double cw[NW > 0 ? NW : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) { d[i] = threadIdx.x + i; db[i] = coef[i * 32 + (threadIdx.x & 3)]; }
#pragma unroll
    for (int i = 0; i < NF; ++i) { f[i] = threadIdx.x + i; fb[i] = (float)coef[i * 32 + 4 + (threadIdx.x & 3)]; }
#pragma unroll
    for (int i = 0; i < NA; ++i) u[i] = threadIdx.x * 2654435761u + i;
#pragma unroll
    for (int i = 0; i < NU; ++i) cu[i] = 1.0f + 1e-3f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < NW; ++i) cw[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    double x = 0.999 + 1e-6 * threadIdx.x;
    float y = 0.999f + 1e-6f * threadIdx.x;
    unsigned accI = 0;
    float accL = 0.f;
    int idx = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) d[i] = fma(x, db[i], d[i]);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fmaf(y, fb[i], f[i]);
#pragma unroll
        for (int i = 0; i < NA; ++i) u[i] = (u[i] ^ (u[i] >> 7)) + 0x9e3779b9u;  // SHF + LOP3/IADD3: ALU pipe
#pragma unroll
        for (int i = 0; i < NL; ++i) accL += sm[(idx + 33 * i) & 1023];
#pragma unroll
        for (int i = 0; i < NU; ++i) { const double t = (double)cu[i]; accI ^= (unsigned)__double2hiint(t); cu[i] = __uint_as_float(__float_as_uint(cu[i]) + 1u); }
#pragma unroll
        for (int i = 0; i < NW; ++i) { const float t = (float)cw[i]; accI ^= __float_as_uint(t); cw[i] = __hiloint2double(__double2hiint(cw[i]), __double2loint(cw[i]) + 64); }
        x = -x; y = -y; idx += 7;
    }
    const long long t1 = clock64();
    double r = accL + (double)accI;
#pragma unroll
    for (int i = 0; i < ND; ++i) r += d[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) r += f[i];
#pragma unroll
    for (int i = 0; i < NA; ++i) r += u[i];
    if (r == 123456789.0) sink[0] = (float)r;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

// exact float -> double without F2F (normal numbers and zero; denormals flush to zero): integer ops only
__device__ __forceinline__ double f2d_bits(float v) {
    const unsigned b = __float_as_uint(v);
    const unsigned t = b & 0x7fffffffu;
    unsigned hi = (t >> 3) + 0x38000000u;
    if (t < 0x00800000u) hi = 0u;
    return __hiloint2double((int)(hi | (b & 0x80000000u)), (int)(b << 29));
}

// One synthesis step per sample: VAR 0 = all FP64 (round-1 kernel), 1 = FP32 FIR + F2F up + FP32 output scaling,
// 2 = as 1 with the integer float->double, 3 = as 1 with x and the window coming from shared memory as float4.
template <int P, int PS, int VAR>
__global__ void k_synth(float* sink, int iters, long long* cyc, const double* __restrict__ coef) {
    __shared__ __align__(16) float xsm[2048];
    __shared__ __align__(16) float wsm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) { xsm[i] = 1e-3f * ((i * 7) % 113); wsm[i] = 0.5f + 1e-4f * (i % 97); }
    __syncthreads();
    double a[P + 1], st[P + 1];
#pragma unroll
    for (int i = 0; i <= P; ++i) { a[i] = coef[i * 32 + (threadIdx.x & 3)]; st[i] = 0.0; }
    const int phi = threadIdx.x & 3;
    float acc = 0.f;
    float xs = 1e-3f * threadIdx.x;
    const long long t0 = clock64();
    if (VAR == 0) {
        double as[PS + 1], t[PS + 1];
#pragma unroll
        for (int i = 0; i <= PS; ++i) { as[i] = coef[i * 32 + 8 + (threadIdx.x & 3)]; t[i] = 0.0; }
        const double w = coef[20];
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
                c[j] = (float)(ov * w);
            }
            const bool hi = (phi & 2) != 0, od = (phi & 1) != 0;
            const float s0 = hi ? c[0] : c[2], s1 = hi ? c[1] : c[3];
            const float k0 = hi ? c[2] : c[0], k1 = hi ? c[3] : c[1];
            const float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
            const float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
            const float snd = od ? r0 : r1, kp = od ? r1 : r0;
            acc += kp + __shfl_xor_sync(0xffffffffu, snd, 1);
        }
    } else {
        float as[PS + 1], t[PS + 1];
#pragma unroll
        for (int i = 0; i <= PS; ++i) { as[i] = (float)coef[i * 32 + 8 + (threadIdx.x & 3)]; t[i] = 0.0f; }
        const float wf = (float)coef[20];
        int pos = (threadIdx.x >> 2) * 4;
        for (int it = 0; it < iters; it += 4) {
            float e[4], c[4], xv[4], wv[4];
            if (VAR == 3) {
                const float4 x4 = *reinterpret_cast<const float4*>(xsm + (pos & 2047));
                const float4 w4 = *reinterpret_cast<const float4*>(wsm + ((pos + 4 * phi * 128) & 2047));
                xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
                wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
                pos += 4;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { xv[j] = xs; xs = -xs; wv[j] = wf; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x = xv[j] * wv[j];
                e[j] = fmaf(as[0], x, t[0]);
#pragma unroll
                for (int q = 0; q < PS; ++q) t[q] = fmaf(as[q + 1], x, t[q + 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double ed = (VAR == 2) ? f2d_bits(e[j]) : (double)e[j];
                const double ov = ed + st[0];
#pragma unroll
                for (int k = 0; k < P; ++k) st[k] = fma(-a[k + 1], ov, st[k + 1]);
                c[j] = (float)ov * wv[j];
            }
            const bool hi = (phi & 2) != 0, od = (phi & 1) != 0;
            const float s0 = hi ? c[0] : c[2], s1 = hi ? c[1] : c[3];
            const float k0 = hi ? c[2] : c[0], k1 = hi ? c[3] : c[1];
            const float r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 2);
            const float r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
            const float snd = od ? r0 : r1, kp = od ? r1 : r0;
            acc += kp + __shfl_xor_sync(0xffffffffu, snd, 1);
        }
    }
    const long long t1 = clock64();
    double r = acc;
#pragma unroll
    for (int i = 0; i <= P; ++i) r += st[i];
    if (r == 123456789.0) sink[0] = (float)r;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

// autocorrelation inner loop: R accumulators, R-deep register window, 2 x 64-bit shared loads per R DFMAs
template <int R>
__global__ void k_acloop(float* sink, int iters, long long* cyc, int segLen) {
    extern __shared__ double xd[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xw = xd + warp * 1024;
    for (int i = lane; i < 1024; i += 32) xw[i] = 1e-3 * ((i * 7 + warp) % 113);
    __syncwarp();
    const int seg = lane & 7, grp = lane >> 3;
    const int n0 = seg * segLen, m0 = grp * (R - (grp ? 1 : 0));
    double tot = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double acc[R], W[R];
#pragma unroll
        for (int j = 0; j < R; ++j) { acc[j] = 0.0; W[j] = xw[n0 + m0 + j]; }
        const double* p = xw + n0;
        const double* pw = p + m0 + R - 1;
        for (int r = 0; r < segLen / R; ++r, p += R, pw += R) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                if (u > 0) W[(u + R - 1) % R] = pw[u];
                const double a = p[u];
#pragma unroll
                for (int j = 0; j < R; ++j) acc[j] = fma(a, W[(u + j) % R], acc[j]);
            }
            W[(R - 1) % R] = pw[R];
        }
#pragma unroll
        for (int j = 0; j < R; ++j) tot += acc[j];
    }
    const long long t1 = clock64();
    if (tot == 123456789.0) sink[0] = (float)tot;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

static float* g_sink;
static long long* g_cyc;
static int g_sms;
static double* g_coef;

template <typename F>
static void run(const char* name, int iters, double dfmaPerIter, F launch) {
    const int wps[] = {1, 2, 3, 4};
    printf("%-34s", name);
    for (int wi = 0; wi < 4; ++wi) {
        const int w = wps[wi];
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
        // cycles per iteration per resident warp of a scheduler; FP64 pipe share if a DFMA holds the pipe 2 cycles
        const double per = (double)best / iters / w;
        printf("  w%d %7.1f cyc (%4.0f%%)", w, per, dfmaPerIter > 0 ? 100.0 * 2.0 * dfmaPerIter / per : 0.0);
    }
    printf("\n");
}

#define MIX(ND, NF, NA, NL, NU, NW) \
    run("mix D" #ND " F" #NF " A" #NA " L" #NL " U" #NU " W" #NW, 4000, ND, [&](int b, int t) { k_mix<ND, NF, NA, NL, NU, NW><<<b, t>>>(g_sink, 4000, g_cyc, g_coef); })

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    g_sms = p.multiProcessorCount;
    cudaMalloc(&g_sink, 64);
    cudaMalloc(&g_cyc, 64);
    {
        double h[128 * 32];
        for (int i = 0; i < 128 * 32; ++i) h[i] = 1e-3 * (1 + i % 37);
        cudaMalloc(&g_coef, sizeof h);
        cudaMemcpy(g_coef, h, sizeof h, cudaMemcpyHostToDevice);
    }
    printf("%s, %d SMs; columns: resident warps per scheduler -> cycles per iteration per warp (DFMA share of those cycles at 2 cyc/DFMA)\n",
           p.name, g_sms);
    MIX(40, 0, 0, 0, 0, 0);
    MIX(40, 10, 0, 0, 0, 0);
    MIX(40, 20, 0, 0, 0, 0);
    MIX(40, 40, 0, 0, 0, 0);
    MIX(40, 80, 0, 0, 0, 0);
    MIX(0, 80, 0, 0, 0, 0);
    MIX(40, 0, 10, 0, 0, 0);
    MIX(40, 0, 20, 0, 0, 0);
    MIX(40, 0, 40, 0, 0, 0);
    MIX(40, 0, 0, 6, 0, 0);
    MIX(40, 0, 0, 12, 0, 0);
    MIX(0, 0, 0, 0, 8, 0);
    MIX(0, 0, 0, 0, 0, 8);
    MIX(40, 0, 0, 0, 4, 0);
    MIX(40, 0, 0, 0, 8, 0);
    MIX(40, 0, 0, 0, 0, 4);
    MIX(40, 0, 0, 0, 0, 8);
    MIX(40, 12, 8, 1, 1, 1);
    run("synth<40,5> all-FP64 (r1)", 8000, 40, [&](int b, int t) { k_synth<40, 5, 0><<<b, t>>>(g_sink, 8000, g_cyc, g_coef); });
    run("synth<40,5> FP32 FIR + F2F", 8000, 40, [&](int b, int t) { k_synth<40, 5, 1><<<b, t>>>(g_sink, 8000, g_cyc, g_coef); });
    run("synth<40,5> FP32 FIR + int f2d", 8000, 40, [&](int b, int t) { k_synth<40, 5, 2><<<b, t>>>(g_sink, 8000, g_cyc, g_coef); });
    run("synth<40,5> FP32 FIR, smem x/w", 8000, 40, [&](int b, int t) { k_synth<40, 5, 3><<<b, t>>>(g_sink, 8000, g_cyc, g_coef); });
    cudaFuncSetAttribute(k_acloop<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_acloop<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    run("autocorr loop R=14 seg 70", 200, 14.0 * 70, [&](int b, int t) { k_acloop<14><<<b, t, (t / 32) * 1024 * 8>>>(g_sink, 200, g_cyc, 70); });
    run("autocorr loop R=21 seg 63", 200, 21.0 * 63, [&](int b, int t) { k_acloop<21><<<b, t, (t / 32) * 1024 * 8>>>(g_sink, 200, g_cyc, 63); });
    return 0;
}
