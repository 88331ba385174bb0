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

// ---------------------------------------------------------------------------------------------
// Control-rate helpers of the RNG-driven LFO variants (quasi-periodic, combined).  The random
// draws and the integer bookkeeping stay on the host (torch global CPU generator);
// every float32 operation of the reference happens here.
namespace modfx {
namespace {

// find_corners, reference modulations.py:219-238: a corner at i (1 <= i <= n-2) is a sign change of
// the first difference; flags are -floor(diff_l(+/-) * (diff_r + 1e-16)) compared against 1.
__global__ void __launch_bounds__(256) corners_kernel(const float* __restrict__ mod, uint8_t* __restrict__ top,
                                                      uint8_t* __restrict__ bottom, int64_t n) {
    const float* m = mod + (int64_t)blockIdx.y * n;
    uint8_t* t = top + (int64_t)blockIdx.y * n;
    uint8_t* bt = bottom + (int64_t)blockIdx.y * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t ft = 0, fb = 0;
        if (i >= 1 && i <= n - 2) {
            const float dl = __fsub_rn(m[i], m[i - 1]);
            const float dr = __fadd_rn(__fsub_rn(m[i + 1], m[i]), 1e-16f);
            const float pos = (dl > 0.0f) ? dl : 0.0f;          // (diff_l > 0) * diff_l
            const float neg = (dl < 0.0f) ? dl : 0.0f;
            ft = (-floorf(__fmul_rn(pos, dr)) == 1.0f) ? 1 : 0;
            fb = (-floorf(__fmul_rn(neg, dr)) == 1.0f) ? 1 : 0;
        }
        t[i] = ft;
        bt[i] = fb;
    }
}

// make_combined_mod_sig's overwrite loop, reference modulations.py:203-209: section s of example b
// covers out[start, start+len) with make_mod_signal(len, len, 1.0, 0.0, shape); later sections win
// at the shared end point, which is what "largest start <= i" selects.
__global__ void __launch_bounds__(256) lfo_sections_kernel(float* __restrict__ out, int64_t n,
                                                           const int32_t* __restrict__ sec_off,
                                                           const int32_t* __restrict__ sec_start,
                                                           const int32_t* __restrict__ sec_len,
                                                           const int32_t* __restrict__ sec_shape) {
    const int b = blockIdx.y;
    const int s0 = sec_off[b], s1 = sec_off[b + 1];
    if (s0 == s1) return;
    float* o = out + (int64_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int sec = -1;
        for (int s = s0; s < s1; ++s)
            if (sec_start[s] <= i && i < sec_start[s] + sec_len[s]) sec = s;
        if (sec < 0) continue;
        const int shape = sec_shape[sec];
        const float len = (float)sec_len[sec];
        const bool rect = shape == MODFX_SHAPE_RECT_COS || shape == MODFX_SHAPE_INV_RECT_COS;
        const LfoDesc d = make_lfo_desc(rect ? 0.5f : 1.0f, 0.0f, shape, 1.0f, len);
        o[i] = lfo_value(d, i - sec_start[sec]);
    }
}

// make_quasi_periodic's concatenation, reference modulations.py:139-159: output positions
// [out_start, out_start + out_len) of example b hold the first out_len points of section
// in[in_start, in_start + in_len) resampled to new_len points (align_corners=True).
__global__ void __launch_bounds__(256) stretch_sections_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               int64_t n, const int32_t* __restrict__ sec_off,
                                                               const int32_t* __restrict__ in_start,
                                                               const int32_t* __restrict__ in_len,
                                                               const int32_t* __restrict__ new_len,
                                                               const int32_t* __restrict__ out_start) {
    const int b = blockIdx.y;
    const int s0 = sec_off[b], s1 = sec_off[b + 1];
    const float* x = in + (int64_t)b * n;
    float* o = out + (int64_t)b * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (s0 == s1) {         // fewer than two corners: signal returned unchanged (modulations.py:136-137)
            o[i] = x[i];
            continue;
        }
        int sec = s0;
        for (int s = s0; s < s1; ++s)
            if (out_start[s] <= i) sec = s;
        const int j = (int)i - out_start[sec];
        const int I = in_len[sec], O = new_len[sec];
        const float* xs = x + in_start[sec];
        o[i] = (I == O) ? xs[j] : upsample_ac(xs, I, upsample_scale_ac_dev(I, O), j);
    }
}

}  // namespace
}  // namespace modfx

extern "C" int modfx_find_corners_f32(const float* mod, uint8_t* top, uint8_t* bottom, int64_t rows, int64_t n,
                                      void* stream) {
    MODFX_REQUIRE(mod && top && bottom, "NULL pointer");
    MODFX_REQUIRE(rows >= 0 && n >= 1, "bad shape");
    if (rows == 0) return MODFX_OK;
    MODFX_REQUIRE(rows <= 65535, "rows=%lld exceeds grid.y", (long long)rows);
    int gx = (int)((n + 255) / 256);
    if (gx > 256) gx = 256;
    corners_kernel<<<dim3(gx, (unsigned)rows), 256, 0, as_stream(stream)>>>(mod, top, bottom, n);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}

extern "C" int modfx_lfo_sections_f32(float* out, int32_t B, int64_t n, const int32_t* sec_off,
                                      const int32_t* sec_start, const int32_t* sec_len, const int32_t* sec_shape,
                                      void* stream) {
    MODFX_REQUIRE(out && sec_off && sec_start && sec_len && sec_shape, "NULL pointer");
    MODFX_REQUIRE(B >= 0 && B <= 65535 && n >= 1, "bad shape");
    if (B == 0) return MODFX_OK;
    int gx = (int)((n + 255) / 256);
    if (gx > 256) gx = 256;
    lfo_sections_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(out, n, sec_off, sec_start, sec_len, sec_shape);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}

extern "C" int modfx_stretch_sections_f32(const float* in, float* out, int32_t B, int64_t n, const int32_t* sec_off,
                                          const int32_t* in_start, const int32_t* in_len, const int32_t* new_len,
                                          const int32_t* out_start, void* stream) {
    MODFX_REQUIRE(in && out && sec_off && in_start && in_len && new_len && out_start, "NULL pointer");
    MODFX_REQUIRE(B >= 0 && B <= 65535 && n >= 1, "bad shape");
    if (B == 0) return MODFX_OK;
    int gx = (int)((n + 255) / 256);
    if (gx > 256) gx = 256;
    stretch_sections_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(in, out, n, sec_off, in_start, in_len,
                                                                       new_len, out_start);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}
