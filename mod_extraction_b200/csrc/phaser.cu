// Phaser: 6 first-order TPT all-pass stages sharing one swept cutoff + feedback, as rendered by
// pedalboard.Phaser (juce::dsp::Phaser) at reference mod_extraction/datasets.py:455-482.
// PARITY UNPINNED: pedalboard==0.7.3 / JUCE are not available offline; the arithmetic follows this
// repo's own scalar CPU restatement of the JUCE algorithm (test infrastructure) and is checked against it.
//
// The filter is a 7-state linear time-varying recurrence (6 all-pass states + the fed-back output),
// strictly serial per example: ~100 dependent cycles per sample.  It is parallelised over time with
// a chunked affine scan (DESIGN.md "P1"):
//   K0  control-rate cutoff: the oscillator phase is accumulated in float32 exactly like the
//       reference (sequentially inside each 8192-sample host block), then mapped to the all-pass
//       coefficient c = 2G-1 for every 4th sample;
//   K1  for every chunk of 128 samples, 8 independent runs give the chunk's state-transition matrix
//       (7 homogeneous runs from the unit states) and its zero-state response (1 run with the audio);
//   K2  a short serial pass per example chains the 7x7 maps to get the true state at every chunk start;
//   K3  every chunk is re-run from its true initial state and writes the mixed, clipped output.
// K1 and K3 stage audio through shared memory so global traffic stays coalesced.
#include "common.cuh"

namespace modfx {
namespace {

constexpr int kStagesAP = 6;     // all-pass stages (juce::dsp::Phaser numStages)
constexpr int kUpd = 4;          // cutoff update period in samples (maxUpdateCounter)
constexpr int kChunk = 128;      // samples per scan chunk
constexpr int kCtl = kChunk / kUpd;
constexpr int kState = 7;        // s1..s6, lastOutput
constexpr int kMapFloats = 8 * kState;   // 7 columns of Phi + zero-state response
constexpr int kSFloats = 8;      // padded state record

struct PhaserArgs {
    const float* x;
    float* y;
    int N;
    float sr;
    const float *rate, *depth, *centre, *feedback, *mix;
    int block;
    const int32_t* index;
    int n_items;
    int n_ctl;       // control points per example = ceil(N / 4)
    int n_chunks;    // ceil(N / kChunk)
    float* C;        // (n_items, n_ctl)   phase, then all-pass coefficient c
    float* Mw;       // (n_items, n_chunks, 56)
    float* S;        // (n_items, n_chunks, 8)
};

__device__ __forceinline__ int example_of(const PhaserArgs& a, int item) { return a.index ? a.index[item] : item; }

// ---- K0a: oscillator phase per control point (juce::dsp::Oscillator semantics) ------------
// One thread per (example, host block).  Inside a block the phase advances by `inc` per control
// point with a wrap at 2 pi; between blocks it advances by inc * (points in the block) in one step.
__global__ void phaser_phase_kernel(const PhaserArgs a, int n_blocks) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= a.n_items * n_blocks) return;
    const int item = gid / n_blocks, blk = gid - item * n_blocks;
    const int b = example_of(a, item);
    const float two_pi = MODFX_TWO_PI_F;
    const float sr_down = (float)((double)a.sr / (double)kUpd);
    const float inc = a.rate[b] * (two_pi / sr_down);
    float phase = 0.0f;
    for (int i = 0; i < blk; ++i) {
        const int64_t s0 = (int64_t)i * a.block, s1 = min((int64_t)(i + 1) * a.block, (int64_t)a.N);
        const int n_down = (int)((s1 + kUpd - 1) / kUpd - (s0 + kUpd - 1) / kUpd);
        float next = phase + inc * (float)n_down;
        while (next >= two_pi) next -= two_pi;
        phase = next;
    }
    const int64_t s0 = (int64_t)blk * a.block, s1 = min((int64_t)(blk + 1) * a.block, (int64_t)a.N);
    const int j0 = (int)((s0 + kUpd - 1) / kUpd), j1 = (int)((s1 + kUpd - 1) / kUpd);
    float* out = a.C + (int64_t)item * a.n_ctl;
    float p = phase;
    for (int j = j0; j < j1; ++j) {
        out[j] = p;
        float next = p + inc;
        while (next >= two_pi) next -= two_pi;
        p = next;
    }
}

// ---- K0b: phase -> all-pass coefficient c = 2G - 1 ---------------------------------------------
__global__ void phaser_coef_kernel(const PhaserArgs a) {
    const int item = blockIdx.y;
    const int b = example_of(a, item);
    const float f_lo = 20.0f;
    const double f_hi_d = (0.49 * (double)a.sr < 20000.0) ? 0.49 * (double)a.sr : 20000.0;
    const float log_min = log10f(f_lo), log_max = log10f((float)f_hi_d);
    const float norm_centre = (log10f(a.centre[b]) - log_min) / (log_max - log_min);   // mapFromLog10
    const float osc_vol = a.depth[b] * 0.5f;
    float* c = a.C + (int64_t)item * a.n_ctl;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_ctl; j += gridDim.x * blockDim.x) {
        float lfo = sinf(c[j] - MODFX_PI_F) * osc_vol + norm_centre;
        lfo = fminf(fmaxf(lfo, 0.0f), 1.0f);
        const float fc = powf(10.0f, lfo * (log_max - log_min) + log_min);               // mapToLog10
        const float g = (float)tan(3.14159265358979323846 * (double)fc / (double)a.sr);  // TPT prewarp
        const float G = g / (1.0f + g);
        c[j] = 2.0f * G - 1.0f;
    }
}

// One sample of the cascade in "c form": out = c*in + (1-c)*s, s' = (1+c)*in - c*s per stage,
// algebraically the TPT all-pass (v = G(in-s); y = v+s; s' = y+v; out = 2y-in) with c = 2G-1.
__device__ __forceinline__ float cascade_step(float in, float (&s)[kStagesAP], float c, float omc, float opc) {
    float v = in;
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) {
        const float o = fmaf(c, v, omc * s[k]);
        s[k] = fmaf(opc, v, -c * s[k]);
        v = o;
    }
    return v;
}

// ---- K1: per-chunk affine map ----------------------------------------------------------------
// block = 256 threads = 32 chunks x 8 runs (7 unit states + the zero-state run driven by x)
__global__ void __launch_bounds__(256) phaser_map_kernel(const PhaserArgs a, int n_super) {
    __shared__ float xs[32][kChunk + 1];
    __shared__ float cs[32][kCtl + 1];
    const int item = blockIdx.x / n_super, sc = blockIdx.x - item * n_super;
    const int b = example_of(a, item);
    const int tid = threadIdx.x;
    const int n_base = sc * 32 * kChunk;
    const float* xr = a.x + (int64_t)b * a.N;
    for (int i = tid; i < 32 * kChunk; i += 256) {
        const int n = n_base + i;
        xs[i / kChunk][i % kChunk] = (n < a.N) ? xr[n] : 0.0f;
    }
    const float* cr = a.C + (int64_t)item * a.n_ctl;
    for (int i = tid; i < 32 * kCtl; i += 256) {
        const int j = n_base / kUpd + i;
        cs[i / kCtl][i % kCtl] = (j < a.n_ctl) ? cr[j] : 0.0f;
    }
    __syncthreads();
    const int ch = tid >> 3, run = tid & 7;
    const int chunk = sc * 32 + ch;
    if (chunk >= a.n_chunks) return;
    const int len = min(kChunk, a.N - chunk * kChunk);
    const float fbk = a.feedback[b];
    const float gate = (run == 7) ? 1.0f : 0.0f;
    float s[kStagesAP];
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) s[k] = (run == k) ? 1.0f : 0.0f;
    float last = (run == 6) ? 1.0f : 0.0f;
    float c = 0.0f, omc = 1.0f, opc = 1.0f;
    for (int j = 0; j < len; ++j) {
        if ((j & (kUpd - 1)) == 0) {
            c = cs[ch][j >> 2];
            omc = 1.0f - c;
            opc = 1.0f + c;
        }
        const float out = cascade_step(gate * xs[ch][j] - last, s, c, omc, opc);
        last = out * fbk;
    }
    float* m = a.Mw + ((int64_t)item * a.n_chunks + chunk) * kMapFloats + run * kState;
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) m[k] = s[k];
    m[6] = last;
}

// ---- K2: chain the maps: state at the start of every chunk --------------------------------------
// 8 lanes per example; lane i < 7 owns state component i.
__global__ void __launch_bounds__(256) phaser_scan_kernel(const PhaserArgs a) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int item = gid >> 3, i = gid & 7;
    const bool live = item < a.n_items;
    const int it = live ? item : 0;
    const float* m = a.Mw + (int64_t)it * a.n_chunks * kMapFloats;
    float* S = a.S + (int64_t)it * a.n_chunks * kSFloats;
    const int ii = (i < kState) ? i : 0;
    float s = 0.0f;
#pragma unroll 4
    for (int cidx = 0; cidx < a.n_chunks; ++cidx) {
        if (live) S[cidx * kSFloats + i] = s;
        const float* mc = m + (int64_t)cidx * kMapFloats;
        float acc = mc[7 * kState + ii];                       // zero-state response
#pragma unroll
        for (int r = 0; r < kState; ++r) acc = fmaf(mc[r * kState + ii], __shfl_sync(kFull, s, r, 8), acc);
        s = (i < kState) ? acc : 0.0f;
    }
}

// ---- K3: re-run every chunk from its true state, mix and clip --------------------------------------
// block = 64 threads = 64 consecutive chunks of one example (8192 samples staged in shared memory)
constexpr int kRunThreads = 64;
__global__ void __launch_bounds__(kRunThreads) phaser_run_kernel(const PhaserArgs a, int n_super) {
    extern __shared__ float sm[];
    float(*xs)[kChunk + 1] = reinterpret_cast<float(*)[kChunk + 1]>(sm);
    float(*cs)[kCtl + 1] = reinterpret_cast<float(*)[kCtl + 1]>(sm + kRunThreads * (kChunk + 1));
    const int item = blockIdx.x / n_super, sc = blockIdx.x - item * n_super;
    const int b = example_of(a, item);
    const int tid = threadIdx.x;
    const int n_base = sc * kRunThreads * kChunk;
    const float* xr = a.x + (int64_t)b * a.N;
    float* yr = a.y + (int64_t)b * a.N;
    for (int i = tid; i < kRunThreads * kChunk; i += kRunThreads) {
        const int n = n_base + i;
        xs[i / kChunk][i % kChunk] = (n < a.N) ? xr[n] : 0.0f;
    }
    const float* cr = a.C + (int64_t)item * a.n_ctl;
    for (int i = tid; i < kRunThreads * kCtl; i += kRunThreads) {
        const int j = n_base / kUpd + i;
        cs[i / kCtl][i % kCtl] = (j < a.n_ctl) ? cr[j] : 0.0f;
    }
    __syncthreads();
    const int chunk = sc * kRunThreads + tid;
    if (chunk < a.n_chunks) {
        const int len = min(kChunk, a.N - chunk * kChunk);
        const float fbk = a.feedback[b];
        const float wet = a.mix[b], dry = 1.0f - a.mix[b];
        const float* S = a.S + ((int64_t)item * a.n_chunks + chunk) * kSFloats;
        float s[kStagesAP];
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) s[k] = S[k];
        float last = S[6];
        float c = 0.0f, omc = 1.0f, opc = 1.0f;
        for (int j = 0; j < len; ++j) {
            if ((j & (kUpd - 1)) == 0) {
                c = cs[tid][j >> 2];
                omc = 1.0f - c;
                opc = 1.0f + c;
            }
            const float in = xs[tid][j];
            const float out = cascade_step(in - last, s, c, omc, opc);
            last = out * fbk;
            const float o = dry * in + wet * out;
            xs[tid][j] = fminf(fmaxf(o, -1.0f), 1.0f);          // datasets.py:472 clip
        }
    }
    __syncthreads();
    for (int i = tid; i < kRunThreads * kChunk; i += kRunThreads) {
        const int n = n_base + i;
        if (n < a.N) yr[n] = xs[i / kChunk][i % kChunk];
    }
}

int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace modfx

using namespace modfx;

extern "C" int64_t modfx_phaser_workspace_bytes(int32_t B, int64_t N) {
    if (B <= 0 || N <= 0) return 0;
    const int64_t n_ctl = (N + kUpd - 1) / kUpd, n_chunks = (N + kChunk - 1) / kChunk;
    return align_up((int64_t)B * n_ctl * 4, 256) + align_up((int64_t)B * n_chunks * kMapFloats * 4, 256) +
           align_up((int64_t)B * n_chunks * kSFloats * 4, 256);
}

extern "C" int modfx_phaser_f32(const float* x, float* y, int32_t B, int64_t N, float sr, const float* rate_hz,
                                const float* depth, const float* centre_hz, const float* feedback,
                                const float* mix, int32_t block, const int32_t* example_index, int32_t n_items,
                                void* workspace, void* stream) {
    MODFX_REQUIRE(x && y && rate_hz && depth && centre_hz && feedback && mix, "NULL pointer");
    MODFX_REQUIRE(B >= 0 && N >= 1 && N < (1ll << 30), "bad shape B=%d N=%lld", B, (long long)N);
    MODFX_REQUIRE(sr > 0.0f, "sample rate must be positive");
    if (block <= 0) block = 8192;           // pedalboard's default buffer_size
    PhaserArgs a{};
    a.x = x; a.y = y; a.N = (int)N; a.sr = sr;
    a.rate = rate_hz; a.depth = depth; a.centre = centre_hz; a.feedback = feedback; a.mix = mix;
    a.block = block;
    a.index = example_index;
    a.n_items = example_index ? n_items : B;
    if (a.n_items == 0) return MODFX_OK;
    MODFX_REQUIRE(a.n_items > 0 && workspace, "workspace is NULL or n_items=%d", a.n_items);
    a.n_ctl = (int)((N + kUpd - 1) / kUpd);
    a.n_chunks = (int)((N + kChunk - 1) / kChunk);
    char* w = static_cast<char*>(workspace);
    a.C = reinterpret_cast<float*>(w);
    w += align_up((int64_t)a.n_items * a.n_ctl * 4, 256);
    a.Mw = reinterpret_cast<float*>(w);
    w += align_up((int64_t)a.n_items * a.n_chunks * kMapFloats * 4, 256);
    a.S = reinterpret_cast<float*>(w);
    cudaStream_t s = as_stream(stream);

    const int n_blocks = (int)((N + block - 1) / block);
    {
        const int total = a.n_items * n_blocks;
        phaser_phase_kernel<<<(total + 63) / 64, 64, 0, s>>>(a, n_blocks);
    }
    {
        int gx = (a.n_ctl + 255) / 256;
        if (gx > 64) gx = 64;
        MODFX_REQUIRE(a.n_items <= 65535, "n_items=%d exceeds grid.y", a.n_items);
        phaser_coef_kernel<<<dim3(gx, a.n_items), 256, 0, s>>>(a);
    }
    {
        const int n_super = (a.n_chunks + 31) / 32;
        phaser_map_kernel<<<(unsigned)((int64_t)a.n_items * n_super), 256, 0, s>>>(a, n_super);
    }
    phaser_scan_kernel<<<(a.n_items * 8 + 255) / 256, 256, 0, s>>>(a);
    {
        const int n_super = (a.n_chunks + kRunThreads - 1) / kRunThreads;
        const size_t smem = sizeof(float) * (size_t)kRunThreads * ((kChunk + 1) + (kCtl + 1));
        MODFX_CUDA_OK(cudaFuncSetAttribute(phaser_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        phaser_run_kernel<<<(unsigned)((int64_t)a.n_items * n_super), kRunThreads, smem, s>>>(a, n_super);
    }
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}
