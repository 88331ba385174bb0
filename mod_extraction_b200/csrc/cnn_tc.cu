// 5x13 time-dilated convolution (64 -> 64 channels) + 2x1 max-pool + PReLU of the LFO-net on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32, operands staged by TMA, accumulators in tensor memory.
// Reference: nn.Conv2d(64, 64, (5, 13), dilation=(1, d), padding="same") -> nn.MaxPool2d((2, 1)) -> nn.PReLU(64),
// mod_extraction/models.py:187-190 (layers 2..6 of configs/models/spectral_2dcnn.yml, 97 % of the network's flops).
//
// Implicit GEMM, no im2col buffer.  Activations are channels-last (B, H, W, 64), so for one kernel tap (kh, kw)
// the A operand of a tile of 128 consecutive frames is the box  x[b, h + kh - 2, w0 + (kw - 6) d : +128, 0:64]:
// a plain 4-D TMA tile whose out-of-range rows / frames are zero-filled by the TMA unit -- that IS
// padding="same".  The B operand is the 64 x 64 weight slice of the tap.  K = 64 channels per tap, 65 taps.
//
// A CTA owns one POOLED output row: conv rows h0 = 2 hp and h0 + 1 accumulate side by side in tensor memory
// (columns 0..63 and 64..127, lane = frame), so the max-pool is a per-thread max of two TMEM reads, and every
// A tile (input row h0 - 2 + r, r = 0..5) is used twice: with kernel row r for the upper conv row and kernel row
// r - 1 for the lower one.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warps 2..5 = epilogue
// (tcgen05.ld -> max -> bias -> PReLU -> 128-byte stores of the channels-last output).  3-stage ring of
// {A 32 KB, B(kh=r) 16 KB, B(kh=r-1) 16 KB}, SWIZZLE_128B in both the tensor maps and the UMMA descriptors.
#include "common.cuh"

#include <cuda.h>

namespace modfx {
namespace {

constexpr int kC = 64;                    // channels in = channels out = K per tap = UMMA N
constexpr int kKH = 5, kKW = 13;
constexpr int kTileW = 128;               // frames per CTA = UMMA M
constexpr int kStages = 3;
constexpr int kPanelA = kTileW * 128;     // bytes: 128 frames x 32 channels (one 128-byte swizzle row per frame)
constexpr int kPanelB = kC * 128;         // bytes: 64 output channels x 32 input channels
constexpr int kStageA = 2 * kPanelA;      // 32 KB
constexpr int kStageB = 2 * kPanelB;      // 16 KB
constexpr int kStageBytes = kStageA + 2 * kStageB;        // 64 KB
constexpr int kThreads = 192;
constexpr int kTmemCols = 128;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 128 /* barriers */;

// kind::tf32 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(kTileW >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded spin: a broken pipeline traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__global__ void __launch_bounds__(kThreads, 1) conv_tf32_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                const __grid_constant__ CUtensorMap tm_w,
                                                                float* __restrict__ y, int H, int W, int dil,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ prelu) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                   // SWIZZLE_128B wants 1024-byte aligned tiles
    unsigned char* gen = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(gen + kStages * kStageBytes);
    const uint32_t bar0 = base + kStages * kStageBytes;
    // barriers: full[s] = bar0 + 8 s, empty[s] = bar0 + 8 (kStages + s), accumulators ready = bar0 + 16 kStages
    const uint32_t bar_acc = bar0 + 16 * kStages;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * kStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int w0 = blockIdx.x * kTileW;
    const int hp = blockIdx.y;
    const int b = blockIdx.z;
    const int h0 = 2 * hp;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar0 + 8 * s, 1);
            mbar_init(bar0 + 8 * (kStages + s), 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(const_cast<uint32_t*>(tmem_slot))),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int it = 0;
            for (int r = 0; r < kKH + 1; ++r) {
                const int hi = h0 - kKH / 2 + r;
                if (hi < 0 || hi >= H) continue;                    // a zero row adds nothing: skipped by both roles
                const uint32_t bytes = kStageA + (r < kKH ? kStageB : 0) + (r >= 1 ? kStageB : 0);
                for (int kw = 0; kw < kKW; ++kw, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(bar0 + 8 * (kStages + s), ph ^ 1u);   // slot free (passes at once the first time round)
                    const uint32_t full = bar0 + 8 * s;
                    const uint32_t st = base + s * kStageBytes;
                    mbar_expect_tx(full, bytes);
                    const int wx = w0 + (kw - kKW / 2) * dil;
                    tma_load_4d(st, &tm_x, full, 0, wx, hi, b);
                    tma_load_4d(st + kPanelA, &tm_x, full, 32, wx, hi, b);
                    if (r < kKH) {
                        const int tap = r * kKW + kw;
                        tma_load_3d(st + kStageA, &tm_w, full, 0, 0, tap);
                        tma_load_3d(st + kStageA + kPanelB, &tm_w, full, 32, 0, tap);
                    }
                    if (r >= 1) {
                        const int tap = (r - 1) * kKW + kw;
                        tma_load_3d(st + kStageA + kStageB, &tm_w, full, 0, 0, tap);
                        tma_load_3d(st + kStageA + kStageB + kPanelB, &tm_w, full, 32, 0, tap);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            int it = 0;
            uint32_t acc_up = 0, acc_lo = 0;
            for (int r = 0; r < kKH + 1; ++r) {
                const int hi = h0 - kKH / 2 + r;
                if (hi < 0 || hi >= H) continue;
                for (int kw = 0; kw < kKW; ++kw, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(bar0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kStageBytes;
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {               // 8 TF32 = 32 bytes of K per instruction
                            const uint64_t ad = umma_desc(st + p * kPanelA + k * 32);
                            if (r < kKH) {
                                umma_tf32(tmem, ad, umma_desc(st + kStageA + p * kPanelB + k * 32), acc_up);
                                acc_up = 1;
                            }
                            if (r >= 1) {
                                umma_tf32(tmem + kC, ad, umma_desc(st + kStageA + kStageB + p * kPanelB + k * 32), acc_lo);
                                acc_lo = 1;
                            }
                        }
                    }
                    umma_commit(bar0 + 8 * (kStages + s));          // frees the slot once these MMAs have read it
                }
            }
            umma_commit(bar_acc);                                   // both accumulators complete
        }
    } else {
        // ===== epilogue: lane = frame, column = output channel =====
        mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                                     // the TMEM lane quadrant this warp may read
        const int w = w0 + 32 * q + lane;
        float* yp = y + (((int64_t)b * (H / 2) + hp) * W + w) * kC;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t up[32], lo[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + 32 * half;
            tmem_ld32(taddr, up);
            tmem_ld32(taddr + kC, lo);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (w < W) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = 32 * half + j + i;
                        const float v = fmaxf(__uint_as_float(up[j + i]), __uint_as_float(lo[j + i])) + __ldg(bias + c);
                        o[i] = v > 0.0f ? v : __ldg(prelu + c) * v;
                    }
                    *reinterpret_cast<float4*>(yp + 32 * half + j) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

int cnn_conv_tf32(const float* x, float* y, int B, int H, int W, int dil, const float* weight, const float* bias,
                  const float* prelu, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(MODFX_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    MODFX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                  "x, y and weight must be 16-byte aligned");
    CUtensorMap tm_x, tm_w;
    {
        // channels-last activations (B, H, W, 64): box = 32 channels x 128 frames of one row of one example
        const cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)kC * 4, (cuuint64_t)W * kC * 4, (cuuint64_t)H * W * kC * 4};
        const cuuint32_t box[4] = {32, (cuuint32_t)kTileW, 1, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        const CUresult r = enc(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(MODFX_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
    }
    {
        // weights (KH, KW, Cout, Cin) = (tap, out channel, in channel): box = 32 in x 64 out channels of one tap
        const cuuint64_t dims[3] = {(cuuint64_t)kC, (cuuint64_t)kC, (cuuint64_t)(kKH * kKW)};
        const cuuint64_t strides[2] = {(cuuint64_t)kC * 4, (cuuint64_t)kC * kC * 4};
        const cuuint32_t box[3] = {32, (cuuint32_t)kC, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        const CUresult r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(weight), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(MODFX_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    }
    MODFX_CUDA_OK(cudaFuncSetAttribute(conv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    conv_tf32_kernel<<<dim3((W + kTileW - 1) / kTileW, H / 2, B), kThreads, kSmemBytes, stream>>>(tm_x, tm_w, y, H, W, dil,
                                                                                                   bias, prelu);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}

}  // namespace modfx
