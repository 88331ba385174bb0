// Micro-benchmark: cycles per sample of the flanger's register-history serial run (one warp, lock-step),
// variants: full loop / no stores / no record loads.   nvcc -arch=sm_100a -O3 -o serial_chain serial_chain.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int VAR, int SPIN>
__global__ void k(const float4* __restrict__ rec_g, float* out, long long* cyc, int ngroups, int reps) {
    extern __shared__ __align__(16) float sm[];
    float4* coef = reinterpret_cast<float4*>(sm);                 // ngroups * 4 records
    float* ring = sm + ngroups * 16;
    float* itb = ring + ngroups * 4 + 64;
    const int lane = threadIdx.x & 31;
    volatile int* flag = reinterpret_cast<volatile int*>(itb + ngroups * 4);
    if (threadIdx.x == 0) *flag = 0;
    __syncthreads();
    if (threadIdx.x >= 32) {
        // companion warps: poll a shared-memory flag like the producers of fc_cta_kernel waiting for the consumer
        // SPIN = 1: bare polling loop; 2: polling with __nanosleep(32); 3: __nanosleep(200)
        int seen = 0;
        while (!seen) {
            seen = *flag;
            if (SPIN == 2 && !seen) __nanosleep(32);
            if (SPIN == 3 && !seen) __nanosleep(200);
        }
        return;
    }
    for (int i = lane; i < ngroups * 4; i += 32) coef[i] = rec_g[i];
    for (int i = lane; i < ngroups * 4 + 64; i += 32) ring[i] = 0.01f * i;
    __syncwarp();
    const float fb = 0.5f;
    long long t0 = clock64();
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        float w[K + 5];
#pragma unroll
        for (int j = 1; j <= K + 1; ++j) w[3 + j] = ring[32 - j];
        float4* dst = reinterpret_cast<float4*>(ring + 32);
        float4* ito = reinterpret_cast<float4*>(itb);
        float4 cur[4], nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = coef[u];
#pragma unroll 1
        for (int g = 0; g < ngroups; ++g) {
            const int gn = min(g + 1, ngroups - 1);
            if (VAR != 2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nxt[u] = coef[4 * gn + u];
            }
            float its[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 cf = cur[u];
                const float far = (K == 1) ? cf.w : __fmul_rn(cf.w, w[3 - u + K + 1]);
                const float it = __fadd_rn(__fmul_rn(cf.y, w[3 - u + (K == 1 ? 1 : K - 1)]),
                                           __fadd_rn(__fmul_rn(cf.z, w[3 - u + (K == 1 ? 2 : K)]), far));
                its[u] = it;
                w[3 - u] = __fadd_rn(cf.x, __fmul_rn(fb, it));
            }
            if (VAR != 1 && lane == 0) {
                dst[g] = make_float4(w[3], w[2], w[1], w[0]);
                ito[g] = make_float4(its[0], its[1], its[2], its[3]);
            }
#pragma unroll
            for (int kk = K + 4; kk >= 4; --kk) w[kk] = w[kk - 4];
            if (VAR != 2) {
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
            }
        }
        acc += w[4];
    }
    long long t1 = clock64();
    if (lane == 0) { *cyc = t1 - t0; *out = acc; *flag = 1; }
}

template <int K, int VAR, int SPIN>
void run(const char* name, const float4* rec, float* out, long long* cyc, int ngroups, int reps) {
    size_t smem = (ngroups * 16 + ngroups * 4 + 64 + ngroups * 4 + 4) * 4;
    const int threads = SPIN ? 128 : 32;
    k<K, VAR, SPIN><<<1, threads, smem>>>(rec, out, cyc, ngroups, reps);
    cudaDeviceSynchronize();
    k<K, VAR, SPIN><<<1, threads, smem>>>(rec, out, cyc, ngroups, reps);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s K=%d: %.2f cycles/sample\n", name, K, (double)c / ((double)reps * ngroups * 4));
}

int main() {
    const int ngroups = 32, reps = 2000;
    float4* rec; float* out; long long* cyc;
    cudaMalloc(&rec, ngroups * 4 * 16); cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    float4 h[128];
    for (int i = 0; i < 128; ++i) h[i] = make_float4(0.001f * i, 0.3f, 0.2f, 0.1f);
    cudaMemcpy(rec, h, sizeof(h), cudaMemcpyHostToDevice);
    run<1, 0, 0>("full loop", rec, out, cyc, ngroups, reps);
    run<1, 1, 0>("no stores", rec, out, cyc, ngroups, reps);
    run<1, 2, 0>("no record loads", rec, out, cyc, ngroups, reps);
    run<2, 0, 0>("full loop", rec, out, cyc, ngroups, reps);
    run<4, 0, 0>("full loop", rec, out, cyc, ngroups, reps);
    run<8, 0, 0>("full loop", rec, out, cyc, ngroups, reps);
    run<1, 0, 1>("+3 warps polling smem", rec, out, cyc, ngroups, reps);
    run<1, 0, 2>("+3 warps poll+nanosleep32", rec, out, cyc, ngroups, reps);
    run<1, 0, 3>("+3 warps poll+nanosleep200", rec, out, cyc, ngroups, reps);
    run<4, 0, 1>("+3 warps polling smem", rec, out, cyc, ngroups, reps);
    run<4, 0, 2>("+3 warps poll+nanosleep32", rec, out, cyc, ngroups, reps);
    return 0;
}
