// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M = 128, SS operands in SWIZZLE_128B shared memory) as a
// function of N, of the operand kind (tf32 / f16) and of how many threads issue.  No loads: operands are whatever
// shared memory holds.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
    return ((uint64_t)kDescHi << 32) | (((addr & 0x3FFFFu) >> 4) | (1u << 16));
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(a), "l"(b), "r"(idesc), "r"(acc)
                     : "memory");
}

// mode 0: one issuer, all MMAs into the same accumulator; mode 1: one issuer, round robin over 512/N accumulators;
// mode 2: two issuers (warps 1 and 3) on disjoint accumulators
template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int fmt = KIND == 0 ? 2 : 1;      // tf32 : bf16
    const uint32_t idesc = (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const int nacc = 512 / N;
    const bool issuer = lane == 0 && (warp == 1 || (mode == 2 && warp == 3));
    if (issuer) {
        const int w = warp == 1 ? 0 : 1;
        const int acc_lo = (mode == 2) ? w * (nacc / 2) : 0;
        const int acc_n = (mode == 0) ? 1 : (mode == 2 ? max(nacc / 2, 1) : nacc);
        const long long t0 = clock64();
        int a = 0;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                mma<KIND>(tmem + (acc_lo + a) * N, desc(base + (k >> 2) * 16384 + (k & 3) * 32),
                          desc(base + 65536 + (k >> 2) * 32768 + (k & 3) * 32), idesc, 1u);
            }
            a = (a + 1 == acc_n) ? 0 : a + 1;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[w])) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar[w])) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[w] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    const int smem = 65536 + 65536 + 2048;
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int grid : {1, 148})
            for (int mode = 0; mode < 3; ++mode)
                for (int N : {64, 128, 192, 256}) {
                    if (mode == 2 && 512 / N < 2) continue;
                    out[0] = out[1] = 0;
                    if (kind == 0) rate_kernel<0><<<grid, 128, smem>>>(N, iters, mode, out);
                    else rate_kernel<1><<<grid, 128, smem>>>(N, iters, mode, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    const double total = (mode == 2 ? 2.0 : 1.0) * iters * 8;
                    const double cyc = (double)(out[0] > out[1] ? out[0] : out[1]);
                    printf("%s grid=%3d mode=%d N=%3d: %7.1f cycles per MMA (ideal %d)\n", kind == 0 ? "tf32" : "bf16", grid, mode, N,
                           cyc / total, N / 2);
                }
    return 0;
}
