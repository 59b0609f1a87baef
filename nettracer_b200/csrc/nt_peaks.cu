// nt_peaks.cu — issue-rate micro-benchmarks: the roofline denominators for the
// intersect-and-shade path, which is bound by the FP64 / FP32 pipes, not by HBM or tensor cores.
// MEASURED_PEAKS.json carries no FP32/FP64 vector figure, so nt_measure_peaks() measures them on the
// device it will be compared with, in the same process, at whatever clock the GPU is running.
//   *_fma   : chains of fused multiply-add, 2 flops per instruction (the pipe's nominal peak)
//   *_nofma : alternating multiply / add with the round-to-nearest intrinsics, 1 flop per
//             instruction — the ceiling of the strict mode, which may not fuse (SPEC-PROVISIONAL §0)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "../../include/nettracer_b200.h"

namespace {

constexpr int ILP = 8;

template <typename T> __device__ __forceinline__ T fma_(T a, T b, T c);
template <> __device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
template <> __device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <typename T> __device__ __forceinline__ T mul_(T a, T b);
template <> __device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T add_(T a, T b);
template <> __device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }

// FMA = true: ILP independent fma chains, 2*ILP flops per inner step.
// FMA = false: ILP independent mul-then-add chains, 2*ILP flops (2*ILP instructions) per step.
template <typename T, bool FMA>
__global__ void __launch_bounds__(256) peak_kernel(int iters, T seed, T *sink, long long *cycles) {
    // cycles[2b] = SM clock ticks, cycles[2b + 1] = nanoseconds (%globaltimer) that block b spent in the loop: the clock
    // estimate divides two intervals taken by the SAME block (round 1 divided one block's ticks by the whole launch's
    // event time and was low by the number of waves)
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + (T)(threadIdx.x + i);
    const T b = (T)0.9999, c = (T)0.0001;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = FMA ? fma_<T>(v[i], b, c) : add_<T>(mul_<T>(v[i], b), c);
    }
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == (T)123456.789) *sink = s; // never true; keeps the chains alive
    if (threadIdx.x == 0) { cycles[2 * blockIdx.x] = t1 - t0; cycles[2 * blockIdx.x + 1] = (long long)(g1 - g0); }
}

template <typename T, bool FMA>
int run(int sm_count, double *gflops, double *mhz) {
    const int blocks = sm_count * 8, threads = 256;
    T *sink = nullptr;
    long long *cyc = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&sink, sizeof(T)) || cudaMalloc(&cyc, sizeof(long long) * 2 * blocks)) return 1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000;
    float ms = 0;
    for (int pass = 0; pass < 3; ++pass) { // calibrate towards ~40 ms, then measure
        cudaEventRecord(e0);
        peak_kernel<T, FMA><<<blocks, threads>>>(iters, (T)1.0, sink, cyc);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); cudaFree(cyc); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass < 2) {
            double scale = 40.0 / (ms > 1e-3 ? ms : 1e-3);
            if (scale > 50) scale = 50;
            iters = (int)(iters * scale) + 1;
        }
    }
    const double flops = (double)blocks * threads * (double)iters * 4 * ILP * 2;
    *gflops = flops / (ms * 1e-3) / 1e9;
    if (mhz) {
        long long h[16];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double best = 0; // the longest-running of the first 8 blocks gives the best-resolved ratio
        long long ns = 0;
        for (int b = 0; b < 8; ++b)
            if (h[2 * b + 1] > ns) { ns = h[2 * b + 1]; best = (double)h[2 * b] / (double)h[2 * b + 1] * 1e3; }
        *mhz = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink); cudaFree(cyc);
    return cudaGetLastError() != cudaSuccess;
}

} // namespace

extern "C" int nt_measure_peaks(int device, nt_peaks *out) {
    if (!out) return NT_ERR_INVALID;
    memset(out, 0, sizeof *out);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return NT_ERR_NO_DEVICE; }
    if (cudaSetDevice(device) != cudaSuccess) return NT_ERR_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    out->sm_count = sms;
    int bad = 0;
    bad |= run<double, true>(sms, &out->f64_fma_gflops, &out->sm_clock_mhz_est);
    bad |= run<double, false>(sms, &out->f64_nofma_gflops, nullptr);
    bad |= run<float, true>(sms, &out->f32_fma_gflops, nullptr);
    bad |= run<float, false>(sms, &out->f32_nofma_gflops, nullptr);
    return bad ? NT_ERR_CUDA : NT_OK;
}
