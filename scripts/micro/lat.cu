// Dependent-issue latency of FP64/FP32 ops on the current GPU (development aid).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void lat(double *out, long long *cyc, int n, double a, double b) {
    double x = a; float xf = (float)a, bf = (float)b;
    long long t0 = clock64();
    asm volatile("" : "+d"(x), "+f"(xf));
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = __dadd_rn(x, b);
            if (OP == 1) x = __dmul_rn(x, b);
            if (OP == 2) x = __fma_rn(x, b, a);
            if (OP == 3) xf = __fmaf_rn(xf, bf, bf);
            if (OP == 4) xf = __fadd_rn(xf, bf);
        }
    }
    asm volatile("" : "+d"(x), "+f"(xf));
    long long t1 = clock64();
    out[threadIdx.x] = x + xf;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *o; long long *c, h;
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 8);
    const char *names[] = { "DADD", "DMUL", "DFMA", "FFMA", "FADD" };
    for (int warps = 1; warps <= 8; warps *= 2)
        for (int op = 0; op < 5; ++op) {
            int n = 4096;
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) lat<0><<<1, 32 * warps>>>(o, c, n, 1.0, 1e-9);
                if (op == 1) lat<1><<<1, 32 * warps>>>(o, c, n, 1.0, 1.0000001);
                if (op == 2) lat<2><<<1, 32 * warps>>>(o, c, n, 1.0, 0.999);
                if (op == 3) lat<3><<<1, 32 * warps>>>(o, c, n, 1.0, 0.999);
                if (op == 4) lat<4><<<1, 32 * warps>>>(o, c, n, 1.0, 1e-9);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%s warps/block=%d: %.2f cycles per dependent op (per warp)\n", names[op], warps, (double)h / (n * 16.0));
        }
    return 0;
}
