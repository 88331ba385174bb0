// LFO synthesis (make_mod_signal, reference mod_extraction/modulations.py:16-57) and
// linear resampling (util.linear_interpolate_last_dim, reference mod_extraction/util.py:15-29)
// as stand-alone batched kernels.  In the render path both are fused into the effect kernels
// (fc.cu); these entry points serve callers that want the signals themselves
// (make_rand_mod_signal, the phaser ground-truth LFO of datasets.py:442-450, targets at 345 frames).
#include "common.cuh"

namespace modfx {
namespace {

__global__ void __launch_bounds__(256) lfo_kernel(float* __restrict__ out, int64_t n, float sr,
                                                  const float* __restrict__ freq, const float* __restrict__ phase,
                                                  const int32_t* __restrict__ shape, const float* __restrict__ exp_) {
    const int b = blockIdx.y;
    const LfoDesc d = make_lfo_desc(freq[b], phase[b], shape[b], exp_ ? exp_[b] : 1.0f, sr);
    float* o = out + (int64_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = lfo_value(d, i);
}

// ATen upsample_linear1d: see upsample_ac in common.cuh for the align_corners=True arithmetic;
// align_corners=False uses src = max(fma(scale, i + 0.5, -0.5), 0) with scale = I / O.
__global__ void __launch_bounds__(256) interp_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                     int64_t I, int64_t O, float scale, int align_corners) {
    const float* xi = in + (int64_t)blockIdx.y * I;
    float* yo = out + (int64_t)blockIdx.y * O;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < O; i += (int64_t)gridDim.x * blockDim.x) {
        float src;
        if (align_corners) src = __fmul_rn(scale, (float)i);
        else src = fmaxf(__fmaf_rn(scale, __fadd_rn((float)i, 0.5f), -0.5f), 0.0f);
        int64_t i0 = (int64_t)src;
        if (i0 > I - 1) i0 = I - 1;
        const int64_t i1 = i0 + ((i0 < I - 1) ? 1 : 0);
        const float l1 = __fsub_rn(src, (float)i0);
        const float l0 = __fsub_rn(1.0f, l1);
        yo[i] = __fmaf_rn(l0, xi[i0], __fmul_rn(l1, xi[i1]));
    }
}

}  // namespace
}  // namespace modfx

using namespace modfx;

extern "C" int modfx_lfo_f32(float* out, int32_t B, int64_t n, float sr, const float* freq, const float* phase,
                             const int32_t* shape, const float* exp_or_null, void* stream) {
    MODFX_REQUIRE(out && freq && phase && shape, "NULL pointer");
    MODFX_REQUIRE(B >= 0 && n >= 1 && sr > 0.0f, "bad arguments B=%d n=%lld sr=%g", B, (long long)n, sr);
    if (B == 0) return MODFX_OK;
    MODFX_REQUIRE(B <= 65535, "B=%d exceeds grid.y", B);
    int gx = (int)((n + 255) / 256);
    if (gx > 1024) gx = 1024;
    lfo_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(out, n, sr, freq, phase, shape, exp_or_null);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}

extern "C" int modfx_interp_linear_f32(const float* in, float* out, int64_t rows, int64_t I, int64_t O,
                                       int32_t align_corners, void* stream) {
    MODFX_REQUIRE(in && out, "NULL pointer");
    MODFX_REQUIRE(rows >= 0 && I >= 1 && O >= 1, "bad shape rows=%lld I=%lld O=%lld", (long long)rows, (long long)I, (long long)O);
    if (rows == 0) return MODFX_OK;
    MODFX_REQUIRE(rows <= 65535, "rows=%lld exceeds grid.y", (long long)rows);
    float scale;
    if (align_corners) scale = upsample_scale_ac(I, O);
    else scale = (float)I / (float)O;
    int gx = (int)((O + 255) / 256);
    if (gx > 2048) gx = 2048;
    interp_kernel<<<dim3(gx, (unsigned)rows), 256, 0, as_stream(stream)>>>(in, out, I, O, scale, align_corners);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}
