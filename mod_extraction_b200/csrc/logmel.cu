// Log-mel front end of the lfo_2dcnn extractor, fused into one kernel:
//   reflect-pad -> frame -> periodic Hann -> 1024-point real FFT -> |.|^2 -> banded mel -> clip -> log
// Replaces Spectral2DCNN.spectrogram (torchaudio MelSpectrogram) + tr.clip + tr.log, reference
// mod_extraction/models.py:170-175,199,207-208.
//
// Why no tensor cores (DESIGN.md "M1"): a dense DFT is 17x the flops of the FFT and needs 3xTF32 or
// 6xBF16 error compensation to hold the 1e-4 log-mel tolerance; the mel matrix is 0.77 % dense.
// Both are not "genuine dense contractions", so the FFT runs in registers on the FP32 pipes
// (fft_core.h) and the mel projection is a <=14-tap banded sum.
//
// Work split: one CTA of 4 warps owns a row (example x channel) and a range of frames and walks it
// 8 frames at a time: the audio span of the 8 frames is staged once in shared memory (each sample
// is read from HBM once per CTA although it belongs to 4 frames), every warp transforms two frames,
// the 8 power spectra meet in shared memory and the mel/log stage writes 8 consecutive frames of
// every mel row (32 B segments of the (R, n_mels, n_frames) output).
#include "common.cuh"
#include "fft_core.h"

namespace modfx {
namespace {

constexpr int kNfft = 1024;
constexpr int kBins = kNfft / 2 + 1;     // 513
constexpr int kWarps = 4;
constexpr int kThreads = kWarps * kWarp;
constexpr int kFB = 2 * kWarps;          // frames per CTA iteration
constexpr int kPStride = kFB;            // row of the power matrix P[bin][frame]: 8 floats = two float4
constexpr int kEStride = 33;             // padded row of the pass-1 -> pass-2 exchange buffer

struct LogMelArgs {
    const float* x;
    float* out;
    int64_t R, T;
    int hop, n_mels, n_frames;
    const float* window;
    const int32_t* fb_start;
    const int32_t* fb_count;
    const float* fb_weight;
    int fb_stride;
    int fb_taps;         // capacity of the compact in-smem weight table (>= sum of fb_count)
    float eps;
    int apply_log;
    int64_t x_row_stride, out_row_stride;
    const int32_t* row_index;
    int chunks;          // CTAs per row
    int iters_per_chunk; // iterations (of kFB frames) per CTA
};

__device__ __forceinline__ int reflect_index(int i, int T) {
    // torch.nn.functional.pad(mode="reflect"): -1 -> 1, T -> T-2
    if (i < 0) i = -i;
    if (i >= T) i = 2 * (T - 1) - i;
    return max(0, min(i, T - 1));      // frames past the end of a short clip are computed but never stored
}

// log(x) for x >= eps > 0 (never denormal after the clip): one MUFU.LG2 and one multiply, <= 3 ulp
__device__ __forceinline__ float fast_log(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.69314718055994530942f;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}

__global__ void __launch_bounds__(kThreads, 3) logmel_kernel(const LogMelArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int item = blockIdx.x / a.chunks;
    const int chunk = blockIdx.x - item * a.chunks;
    const int64_t row = a.row_index ? a.row_index[item] : item;
    const int T = (int)a.T;
    const int span = (kFB - 1) * a.hop + kNfft;      // samples staged per iteration

    float* win = smem;                         // [1024]
    float2* tw1 = reinterpret_cast<float2*>(win + kNfft);       // [16][32]  exp(-2 pi i n2 k1 / 512) as (cos, sin)
    float* P = win + kNfft + 2 * 512;          // [513][8]
    float* E = P + kBins * kPStride + 4;       // per warp: float2 Ew[32][33]  (kBins*8 + 4 keeps 16 B alignment)
    float2* Ew = reinterpret_cast<float2*>(E) + warp * (32 * kEStride);
    float* sbuf = E + kWarps * (2 * 32 * kEStride);             // [span rounded up to 4]
    float* melw = sbuf + ((span + 3) & ~3);                     // [fb_taps] band weights, bands back to back
    unsigned short* moff = reinterpret_cast<unsigned short*>(melw + a.fb_taps);    // [n_mels + 1] first weight of band m
    unsigned short* mstart = moff + a.n_mels + 2;                                  // [n_mels] first FFT bin of band m

    for (int i = tid; i < kNfft; i += kThreads) win[i] = a.window[i];
    for (int i = tid; i < 512; i += kThreads) {
        const int k1 = i >> 5, n2 = i & 31;
        float s, c;
        sincospif(-(float)(n2 * k1) / 256.0f, &s, &c);
        tw1[i] = make_float2(c, s);
    }
    // compact copy of the banded mel table (<= 14 taps per band for the reference configuration):
    // counts -> exclusive prefix (every thread sums the counts below its band from shared memory)
    for (int m = tid; m < a.n_mels; m += kThreads) mstart[m] = (unsigned short)min(a.fb_count[m], 65535);
    __syncthreads();
    for (int m = tid; m <= a.n_mels; m += kThreads) {
        int off = 0;
        for (int i = 0; i < m; ++i) off += mstart[i];
        moff[m] = (unsigned short)min(off, a.fb_taps);
    }
    __syncthreads();
    for (int m = tid; m < a.n_mels; m += kThreads) {
        mstart[m] = (unsigned short)a.fb_start[m];
        const int o = moff[m], c = moff[m + 1] - o;
        for (int j = 0; j < c; ++j) melw[o + j] = a.fb_weight[(int64_t)m * a.fb_stride + j];
    }
    // split twiddle of this lane's bins k = k1 + 16 k2: w_k = w_{k1} * w_{16 k2}
    float c1, s1;
    sincospif((float)(lane & 15) / 512.0f, &s1, &c1);

    const float* xr = a.x + row * a.x_row_stride;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) && ((a.hop & 3) == 0);
    float* orow = a.out + row * a.out_row_stride;
    const int it_begin = chunk * a.iters_per_chunk;
    const int it_end = min(it_begin + a.iters_per_chunk, (a.n_frames + kFB - 1) / kFB);
    bool prefetched = false;

    for (int it = it_begin; it < it_end; ++it) {
        const int t0 = it * kFB;
        // ---- stage the audio span of frames t0 .. t0+7 (center=True: frame t starts at t*hop - 512)
        const int s0 = t0 * a.hop - kNfft / 2;
        if (prefetched) {
            asm volatile("cp.async.wait_group 0;\n" ::);
        } else {
            for (int i = tid; i < span; i += kThreads) sbuf[i] = xr[reflect_index(s0 + i, T)];
        }
        __syncthreads();

        const int lf = 2 * warp;                        // local index of this warp's first frame
        const bool active = t0 + lf < a.n_frames;
        float2 v[32];
        if (active) {
            // ---- pass 1: lane = n2; 16-point FFT over n1 of z[32 n1 + n2], both frames
#pragma unroll
            for (int f = 0; f < 2; ++f) {
                float2 u[16];
                const float* fr = sbuf + (lf + f) * a.hop;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const float2 x2 = *reinterpret_cast<const float2*>(fr + 64 * n1 + 2 * lane);
                    const float2 w2 = *reinterpret_cast<const float2*>(win + 64 * n1 + 2 * lane);
                    u[n1] = c_mul2(x2, w2);
                }
                fft_dif<16>(u);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    const float2 t = tw1[k1 * 32 + lane];
                    Ew[(f * 16 + k1) * kEStride + lane] = c_mul_tw(u[BitRev<16>::of(k1)], t.x, t.y);
                }
            }
        }
        __syncthreads();        // every warp is done with sbuf: the next span may land in it
        {
            const int s1n = (t0 + kFB) * a.hop - kNfft / 2;
            prefetched = (it + 1 < it_end) && vec_ok && s1n >= 0 && (s1n + span <= T);
            if (prefetched) {
                for (int i = tid; i < (span >> 2); i += kThreads) cp_async16(sbuf + 4 * i, xr + s1n + 4 * i);
                asm volatile("cp.async.commit_group;\n" ::);
            }
        }
        if (active) {
            // ---- pass 2: lane = (frame f, k1); 32-point FFT over n2
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) v[n2] = Ew[lane * kEStride + n2];
            fft_dif<32>(v);
            // ---- real-FFT split + power.  Bins k and 512-k share all their intermediate terms
            // (X[512-k] uses the same sums / differences with two signs flipped), so each pair is
            // computed once: this lane handles its own bins k = k1 + 16 k2 for k2 < 16 together with the
            // mirror bins 512-k, which belong to the partner lane (f, 16-k1) at k2' = 31-k2.
            constexpr float C64[32] = MODFX_C64;
            constexpr float S64[32] = MODFX_S64;
            const int f = lane >> 4, k1 = lane & 15;
            const int partner = (lane & 16) | ((16 - k1) & 15);
            const int k1m = (16 - k1) & 15;                 // k1 of the mirror bins
            // the two 16-byte halves of a power row are swapped on every other group of 4 rows so that
            // rows r and r+4 (same banks) are read / written through different banks
            float* Pown = P + ((lf + f) ^ (((k1 >> 2) & 1) << 2));
            // mirror bins: 512 - k = k1m + 16 (31 - k2) for k1 != 0 and 16 (32 - k2) for k1 == 0, i.e. one base
            // pointer minus k2 * 16 rows in both cases
            float* Pmir = P + ((lf + f) ^ (((k1m >> 2) & 1) << 2)) + ((k1 == 0) ? 512 : (k1m + 496)) * kPStride;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const int own = BitRev<32>::of(k2);
                const int src_other = BitRev<32>::of(31 - k2);          // Z[512-k] lives in the partner lane
                const int src_self = BitRev<32>::of((32 - k2) & 31);    // ... or in this lane when k1 == 0
                float2 p = make_float2(__shfl_sync(kFull, v[src_other].x, partner),
                                       __shfl_sync(kFull, v[src_other].y, partner));
                if (k1 == 0) p = v[src_self];
                // split twiddle (cos, sin)(2 pi k / 1024) = (c1 + i s1) * (C + i S), k = k1 + 16 k2
                const float2 cs = c_mul_tw(make_float2(c1, s1), C64[k2], S64[k2]);
                const float2 pw = rfft_split_power_pair(v[own], p, cs.x, cs.y);
                Pown[(k1 + 16 * k2) * kPStride] = pw.x;
                if (!(k1 == 0 && k2 == 0)) Pmir[-k2 * 16 * kPStride] = pw.y;
            }
            if (k1 == 0) {
                // k = 0 was written above (X[0] = Re Z[0] + Im Z[0]); Nyquist and the self-paired k = 256
                const float d = v[0].x - v[0].y;                        // X[512] = Re Z[0] - Im Z[0]
                Pown[512 * kPStride] = d * d;
                const float2 zq = v[BitRev<32>::of(16)];                // Z[256]: its own mirror, X[256] = conj Z[256]
                Pown[256 * kPStride] = zq.x * zq.x + zq.y * zq.y;
            }
        }
        __syncthreads();

#ifndef MODFX_EXP_SKIP_MEL
        // ---- banded mel projection + clip + log.  A thread owns whole mel bands (a low one and its
        // mirror from the top, so the 1..14 taps balance) and all 8 frames of the iteration: per tap
        // one weight, two 16-byte reads of the power row, eight FMAs.
        {
            const int half = (a.n_mels + 1) / 2;
            for (int mm = tid; mm < half; mm += kThreads) {
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const int m = side ? (a.n_mels - 1 - mm) : mm;
                    if (side && m <= mm) break;
                    const int o = moff[m], cnt = moff[m + 1] - o;
                    const float* w = melw + o;
                    float acc[kFB];
#pragma unroll
                    for (int t = 0; t < kFB; ++t) acc[t] = 0.0f;
                    const int r0 = (int)mstart[m];
                    const char* Pbytes = reinterpret_cast<const char*>(P);
#pragma unroll 2
                    for (int j = 0; j < cnt; ++j) {
                        const float wj = w[j];
                        // row r starts at byte 32 r; its two 16-byte halves are swapped when bit 2 of r is set
                        const unsigned lin = (unsigned)(r0 + j) * 32u;
                        const unsigned lo16 = lin ^ ((lin >> 3) & 16u);
                        const float4 p0 = *reinterpret_cast<const float4*>(Pbytes + lo16);
                        const float4 p1 = *reinterpret_cast<const float4*>(Pbytes + (lo16 ^ 16u));
                        acc[0] = fmaf(wj, p0.x, acc[0]); acc[1] = fmaf(wj, p0.y, acc[1]);
                        acc[2] = fmaf(wj, p0.z, acc[2]); acc[3] = fmaf(wj, p0.w, acc[3]);
                        acc[4] = fmaf(wj, p1.x, acc[4]); acc[5] = fmaf(wj, p1.y, acc[5]);
                        acc[6] = fmaf(wj, p1.z, acc[6]); acc[7] = fmaf(wj, p1.w, acc[7]);
                    }
                    // results go through shared memory (the idle FFT exchange buffer) so that the global
                    // stores below are 32-byte segments instead of 4-byte scatters
                    float4 o0, o1;
                    if (a.apply_log) {
                        o0 = make_float4(fast_log(fmaxf(acc[0], a.eps)), fast_log(fmaxf(acc[1], a.eps)),
                                         fast_log(fmaxf(acc[2], a.eps)), fast_log(fmaxf(acc[3], a.eps)));
                        o1 = make_float4(fast_log(fmaxf(acc[4], a.eps)), fast_log(fmaxf(acc[5], a.eps)),
                                         fast_log(fmaxf(acc[6], a.eps)), fast_log(fmaxf(acc[7], a.eps)));
                    } else {
                        o0 = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        o1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    }
                    reinterpret_cast<float4*>(E)[2 * m] = o0;
                    reinterpret_cast<float4*>(E)[2 * m + 1] = o1;
                }
            }
            __syncthreads();
            // 8 consecutive lanes write the 8 frames of one band: one 32-byte segment per band row
            const int live = min(kFB, a.n_frames - t0);
            const int t = tid & (kFB - 1);
            if (t < live) {
                const float* ep = E + tid;
                float* op = orow + (int64_t)(tid >> 3) * a.n_frames + t0 + t;
                const int64_t ostep = (int64_t)(kThreads / kFB) * a.n_frames;
#pragma unroll 4
                for (int m = tid >> 3; m < a.n_mels; m += kThreads / kFB) {
                    *op = *ep;
                    op += ostep;
                    ep += kThreads;
                }
            }
        }
#endif
        // no barrier here: the next iteration's first barrier orders this read of P before the next write
        // (P is only written after the barrier that follows pass 1)
    }
}

}  // namespace
}  // namespace modfx

using namespace modfx;

extern "C" int modfx_logmel_f32(const float* x, float* out, int64_t R, int64_t T, int32_t n_fft, int32_t hop,
                                int32_t n_mels, const float* window, const int32_t* fb_start,
                                const int32_t* fb_count, const float* fb_weight, int32_t fb_stride, int32_t fb_taps,
                                float eps, int32_t apply_log, int64_t x_row_stride, int64_t out_row_stride,
                                const int32_t* row_index, int32_t n_index, void* stream) {
    MODFX_REQUIRE(x && out && window && fb_start && fb_count && fb_weight, "NULL pointer");
    MODFX_REQUIRE(R >= 0 && T >= 1, "bad shape R=%lld T=%lld", (long long)R, (long long)T);
    if (n_fft != kNfft) return fail(MODFX_ERR_UNSUPPORTED, "n_fft=%d (only 1024 is built)", n_fft);
    if (hop < 2 || hop > 512 || (hop & 1))
        return fail(MODFX_ERR_UNSUPPORTED, "hop=%d (even hops in [2, 512] are built)", hop);
    MODFX_REQUIRE(T > n_fft / 2, "reflect padding needs T > n_fft/2 (T=%lld)", (long long)T);   // torch raises too
    MODFX_REQUIRE(T < (1ll << 30), "T too long");
    MODFX_REQUIRE(n_mels >= 1 && fb_stride >= 1 && fb_taps >= 1, "bad mel table");
    if (n_mels * kFB > kWarps * 2 * 32 * kEStride || fb_taps > 8192)
        return fail(MODFX_ERR_UNSUPPORTED, "mel table too large for shared memory (n_mels=%d taps=%d)", n_mels, fb_taps);
    if (row_index) R = n_index;
    if (R <= 0) return MODFX_OK;
    LogMelArgs a{};
    a.x = x; a.out = out; a.R = R; a.T = T; a.hop = hop; a.n_mels = n_mels;
    a.n_frames = (int)(T / hop) + 1;
    a.window = window; a.fb_start = fb_start; a.fb_count = fb_count; a.fb_weight = fb_weight;
    a.fb_stride = fb_stride; a.eps = eps; a.apply_log = apply_log;
    a.fb_taps = fb_taps;
    a.x_row_stride = x_row_stride > 0 ? x_row_stride : T;
    a.out_row_stride = out_row_stride > 0 ? out_row_stride : (int64_t)n_mels * a.n_frames;
    a.row_index = row_index;
    const int iters = (a.n_frames + kFB - 1) / kFB;
    // enough CTAs to fill the chip twice over even for a handful of rows
    // One CTA per row unless there are too few rows to fill the chip a few times over (every CTA pays
    // the table set-up and its first span is staged synchronously; finer chunking measured slower).
    int chunks = 1;
    const int64_t want = 4ll * num_sms();
    if (R < want) chunks = (int)((want + R - 1) / R);
    if (chunks > iters) chunks = iters;
    a.iters_per_chunk = (iters + chunks - 1) / chunks;
    a.chunks = (iters + a.iters_per_chunk - 1) / a.iters_per_chunk;
    MODFX_REQUIRE(R * a.chunks < (1ll << 31), "grid too large");
    const int span = (kFB - 1) * hop + kNfft;
    const size_t smem = sizeof(float) * (size_t)(kNfft + 2 * 512 + kBins * kPStride + 4 + kWarps * 2 * 32 * kEStride +
                                                 ((span + 3) & ~3) + fb_taps) +
                        sizeof(unsigned short) * (size_t)(2 * n_mels + 4) + 16;
    MODFX_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logmel_kernel<<<(unsigned)(R * a.chunks), kThreads, smem, as_stream(stream)>>>(a);
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}
