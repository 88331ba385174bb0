// Microbenchmark: issue rate of FFMA vs FFMA2 (packed fp32x2) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            else a[i] = __ffma2_rn(a[i], m, c);
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float2* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float2));
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int warps = 1; warps <= 16; warps *= 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        dim3 grid(148 * 4), block(32 * warps);
        if (mode == 0) k<0><<<grid, block>>>(out, 100, make_float2(1, 2)); else k<1><<<grid, block>>>(out, 100, make_float2(1, 2));
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, block>>>(out, iters, make_float2(1, 2)); else k<1><<<grid, block>>>(out, iters, make_float2(1, 2));
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)grid.x * block.x * iters * 16.0;
        printf("%s warps/CTA=%2d (x4 CTA/SM): %.3f ms  %.1f TFLOP/s (2*fma)\n", mode ? "FFMA2" : "FFMA ", warps, ms, 2 * fma / ms / 1e9);
    }
    return 0;
}
