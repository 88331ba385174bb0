// Phaser: 6 first-order TPT all-pass stages sharing one swept cutoff + feedback, as rendered by
// pedalboard.Phaser (juce::dsp::Phaser) at reference mod_extraction/datasets.py:455-482.
// PARITY UNPINNED: pedalboard==0.7.3 / JUCE are not available offline; the arithmetic follows this
// repo's own scalar CPU restatement of the JUCE algorithm (test infrastructure) and is checked against it.
//
// The filter is a 7-state linear time-varying recurrence (6 all-pass states + the fed-back output),
// strictly serial per example: ~100 dependent cycles per sample.  It is parallelised over time with
// a chunked affine scan (DESIGN.md "P1"):
//   K0  control-rate cutoff: the oscillator phase is accumulated in float32 exactly like the
//       reference (sequentially inside each 8192-sample host block) and mapped to the all-pass
//       coefficient c = 2G-1 for every 4th sample;
//   K1  for every chunk of 128 samples, 8 independent runs give the chunk's state-transition matrix
//       (7 homogeneous runs from the unit states) and its zero-state response (1 run with the audio);
//   K2  the 7x7 maps are chained to get the true state at every chunk start: groups of 32 maps are composed
//       in parallel, the group maps chained serially, and every group re-chained from its initial state;
//   K3  every chunk is re-run from its true initial state and writes the mixed, clipped output.
// K1 and K3 stage audio through shared memory so global traffic stays coalesced.
#include "common.cuh"

#include <stdlib.h>

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

// ---- K0: control-rate all-pass coefficient ----------------------------------------------------
// juce::dsp::Oscillator semantics: inside a host block the phase advances by `inc` per control point
// with a wrap at 2 pi (sequential float32 adds, reproduced exactly); between blocks it advances by
// inc * (points in the block) in one step.  A CTA owns 32 (example, host block) pairs: warp 0 walks
// the 32 phase sequences (lane = pair) 32 control points at a time into a shared tile, then all 256
// threads map the 1024 phases to c = 2G-1 (sin, 10^x, tan: the expensive part) and store them with
// 128-byte rows.
constexpr int kK0Threads = 256;
__global__ void __launch_bounds__(kK0Threads) phaser_ctl_kernel(const PhaserArgs a, int n_blocks) {
    __shared__ float tile[32][33];
    __shared__ int j0s[32], j1s[32], items[32];
    __shared__ float vol[32], ctr[32];
    const int tid = threadIdx.x;
    const float two_pi = MODFX_TWO_PI_F;
    const float f_lo = 20.0f;
    const double f_hi_d = (0.49 * (double)a.sr < 20000.0) ? 0.49 * (double)a.sr : 20000.0;
    const float log_min = log10f(f_lo), log_max = log10f((float)f_hi_d);
    const float log_span = log_max - log_min;
    const float w0 = MODFX_PI_F / a.sr;
    float p = 0.0f, inc = 0.0f;
    if (tid < 32) {
        const int pair = blockIdx.x * 32 + tid;
        const bool live = pair < a.n_items * n_blocks;
        const int item = live ? pair / n_blocks : 0, blk = live ? pair - item * n_blocks : 0;
        const int b = example_of(a, item);
        const float sr_down = (float)((double)a.sr / (double)kUpd);
        inc = a.rate[b] * (two_pi / sr_down);
        for (int i = 0; i < blk; ++i) {
            const int64_t s0 = (int64_t)i * a.block, s1 = min((int64_t)(i + 1) * a.block, (int64_t)a.N);
            const int n_down = (int)((s1 + kUpd - 1) / kUpd - (s0 + kUpd - 1) / kUpd);
            float next = p + inc * (float)n_down;
            while (next >= two_pi) next -= two_pi;
            p = next;
        }
        const int64_t s0 = (int64_t)blk * a.block, s1 = min((int64_t)(blk + 1) * a.block, (int64_t)a.N);
        j0s[tid] = live ? (int)((s0 + kUpd - 1) / kUpd) : 0;
        j1s[tid] = live ? (int)((s1 + kUpd - 1) / kUpd) : 0;
        items[tid] = item;
        vol[tid] = a.depth[b] * 0.5f;                                           // oscVolume target
        ctr[tid] = (log10f(a.centre[b]) - log_min) / log_span;                   // mapFromLog10(centre)
    }
    const int rounds = (int)((((int64_t)a.block + kUpd - 1) / kUpd + 31) / 32) + 1;
    for (int r = 0; r < rounds; ++r) {
        __syncthreads();
        if (tid < 32) {
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                tile[tid][q] = p;
                float next = p + inc;
                while (next >= two_pi) next -= two_pi;
                p = next;
            }
        }
        __syncthreads();
        for (int e = tid; e < 32 * 32; e += kK0Threads) {
            const int t = e >> 5, q = e & 31;
            const int j = j0s[t] + 32 * r + q;
            if (j < j1s[t]) {
                float lfo = sinf(tile[t][q] - MODFX_PI_F) * vol[t] + ctr[t];
                lfo = fminf(fmaxf(lfo, 0.0f), 1.0f);
                const float fc = exp10f(lfo * log_span + log_min);               // mapToLog10
                const float g = tanf(w0 * fc);                                   // TPT prewarp
                a.C[(int64_t)items[t] * a.n_ctl + j] = 2.0f * (g / (1.0f + g)) - 1.0f;
            }
        }
    }
}

// One sample of the cascade.  With c = 2G-1 the TPT all-pass (v = G(in-s); y = v+s; s' = y+v; out = 2y-in)
// is out = c*in + (1-c)*s, s' = (1+c)*in - c*s.  Carrying w = (1-c)*s instead of s turns every stage into
// the transposed direct form II of a first-order all-pass -- two dependent FMAs:  out = c*in + w,
// w' = in - c*out  -- at the price of rescaling w by (1-c_new)/(1-c_old) when the coefficient moves (every
// 4th sample).  All kernels keep the states in this form, scaled for the coefficient of the last sample done.
__device__ __forceinline__ float cascade_step(float in, float (&w)[kStagesAP], float c) {
    float v = in;
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) {
        const float y = fmaf(c, v, w[k]);
        w[k] = fmaf(-c, y, v);
        v = y;
    }
    return v;
}

__device__ __forceinline__ void cascade_retune(float (&w)[kStagesAP], float c_old, float c_new) {
    const float r = __fdividef(1.0f - c_new, 1.0f - c_old);     // 1 - c = 2 (1 - G) in (0.06, 2]: well conditioned
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) w[k] *= r;
}

// Two runs of the cascade at once, one per half of a float2 (same coefficient): Blackwell's packed FFMA2 does the
// work of two FFMAs in one issue slot, and the map kernel is bound by instruction issue.
__device__ __forceinline__ float2 cascade_step2(float2 in, float2 (&w)[kStagesAP], float c) {
    const float2 c2 = make_float2(c, c), nc2 = make_float2(-c, -c);
    float2 v = in;
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) {
        const float2 y = __ffma2_rn(c2, v, w[k]);
        w[k] = __ffma2_rn(nc2, y, v);
        v = y;
    }
    return v;
}

__device__ __forceinline__ void cascade_retune2(float2 (&w)[kStagesAP], float c_old, float c_new) {
    const float r = __fdividef(1.0f - c_new, 1.0f - c_old);
    const float2 r2 = make_float2(r, r);
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) w[k] = __fmul2_rn(w[k], r2);
}

// ---- K1: per-chunk affine map ----------------------------------------------------------------
// block = 128 threads = 4 warps; warp p performs runs 2p and 2p+1 packed in a float2 (runs 0..6: unit state
// e_r, no input; run 7: zero state, driven by x) for 32 consecutive chunks, lane = chunk.  The run pair is
// warp-uniform, so the homogeneous warps never touch the audio.
template <bool DRIVEN>
__device__ __forceinline__ void map_run2(const float (*xs)[kChunk + 1], const float (*cs)[kCtl + 2], int ch, int len,
                                         float fbk, float2 (&s)[kStagesAP], float2& out_prev, int j_begin, float c_cur) {
    // j_begin is 0, or kUpd when the caller has already done the first control group by hand
    // out_prev holds the previous cascade outputs; lastOutput = out_prev * feedback.
    // cs[ch][0] is the coefficient of the sample before the chunk, cs[ch][1 + g] that of control group g.
    const float2 nfb = make_float2(-fbk, -fbk);
    if (len == kChunk) {                       // full chunk: no guards, one coefficient per 4 samples
#pragma unroll 2
        for (int g = j_begin / kUpd; g < kCtl; ++g) {
            const float c = cs[ch][1 + g];
            cascade_retune2(s, c_cur, c);
            c_cur = c;
#pragma unroll
            for (int q = 0; q < kUpd; ++q) {
                const float2 u = DRIVEN ? __ffma2_rn(nfb, out_prev, make_float2(0.0f, xs[ch][g * kUpd + q]))
                                        : __fmul2_rn(nfb, out_prev);
                out_prev = cascade_step2(u, s, c);
            }
        }
    } else {
        for (int j = j_begin; j < len; ++j) {
            const float c = cs[ch][1 + (j >> 2)];
            if (c != c_cur) {
                cascade_retune2(s, c_cur, c);
                c_cur = c;
            }
            const float2 u = DRIVEN ? __ffma2_rn(nfb, out_prev, make_float2(0.0f, xs[ch][j])) : __fmul2_rn(nfb, out_prev);
            out_prev = cascade_step2(u, s, c);
        }
    }
}

constexpr int kMapThreads = 128;
__global__ void __launch_bounds__(kMapThreads) phaser_map_kernel(const PhaserArgs a, int n_super) {
    __shared__ float xs[32][kChunk + 1];
    __shared__ float cs[32][kCtl + 2];
    const int item = blockIdx.x / n_super, sc = blockIdx.x - item * n_super;
    const int b = example_of(a, item);
    const int tid = threadIdx.x;
    const int n_base = sc * 32 * kChunk;
    const float* xr = a.x + (int64_t)b * a.N;
    for (int i = tid; i < 32 * kChunk; i += kMapThreads) {
        const int n = n_base + i;
        xs[i / kChunk][i % kChunk] = (n < a.N) ? xr[n] : 0.0f;
    }
    const float* cr = a.C + (int64_t)item * a.n_ctl;
    for (int i = tid; i < 32 * (kCtl + 1); i += kMapThreads) {
        const int row = i / (kCtl + 1), col = i % (kCtl + 1);          // col 0 = coefficient before the chunk
        const int j = n_base / kUpd + row * kCtl + col - 1;
        cs[row][col] = cr[max(0, min(j, a.n_ctl - 1))];
    }
    __syncthreads();
    const int pair = tid >> 5, ch = tid & 31;
    const int chunk = sc * 32 + ch;
    if (chunk >= a.n_chunks) return;
    const int len = min(kChunk, a.N - chunk * kChunk);
    const float fbk = a.feedback[b];
    float2 s[kStagesAP];
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) s[k] = make_float2((2 * pair == k) ? 1.0f : 0.0f, (2 * pair + 1 == k) ? 1.0f : 0.0f);
    float2 out_prev = make_float2(0.0f, 0.0f);
    const float c_prev = cs[ch][0];
    if (pair == 3) {
        // runs 6 and 7.  State 6 is lastOutput = out_prev * feedback: a unit lastOutput is out_prev = 1 / feedback; to
        // stay finite for feedback = 0 the unit run is fed u = -1 by hand for the first sample.  Both runs start
        // with zero all-pass states, so there is nothing to retune before the first control group.
        const float c0 = cs[ch][1];
        out_prev = cascade_step2(make_float2(-1.0f, xs[ch][0]), s, c0);
        const int n0 = min(kUpd, len);
        const float2 nfb = make_float2(-fbk, -fbk);
        for (int q = 1; q < n0; ++q)
            out_prev = cascade_step2(__ffma2_rn(nfb, out_prev, make_float2(0.0f, xs[ch][q])), s, c0);
        map_run2<true>(xs, cs, ch, len, fbk, s, out_prev, kUpd, c0);
    } else {
        map_run2<false>(xs, cs, ch, len, fbk, s, out_prev, 0, c_prev);
    }
    float* m = a.Mw + ((int64_t)item * a.n_chunks + chunk) * kMapFloats + 2 * pair * kState;
#pragma unroll
    for (int k = 0; k < kStagesAP; ++k) {
        m[k] = s[k].x;
        m[kState + k] = s[k].y;
    }
    m[6] = out_prev.x * fbk;
    m[kState + 6] = out_prev.y * fbk;
}

// ---- K2: chain the maps: state at the start of every chunk --------------------------------------
// 8 lanes per example; lane i < 7 owns state component i.
// The chain of n_chunks affine maps per example is resolved in three short passes instead of one long
// one (689 dependent steps for 2 s, 20 672 for 60 s): (a) groups of 32 consecutive chunk maps are
// composed into one map each, all groups in parallel; (b) the group maps are chained serially (22 steps
// for 2 s); (c) every group re-chains its 32 chunk maps from its now known initial state.
constexpr int kGroup = 32;
constexpr int kSerialChainMax = 4096;     // up to this many maps per example a single serial chain is used (measured faster)

// (a) compose the maps of one group: 8 lanes per group, lane r carries column r of [Phi | z]
__global__ void __launch_bounds__(256) phaser_compose_kernel(const PhaserArgs a, int n_groups, float* __restrict__ GM) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t team = gid >> 3;
    const int r = (int)(gid & 7);
    if (team >= (int64_t)a.n_items * n_groups) return;
    const int item = (int)(team / n_groups), g = (int)(team - (int64_t)item * n_groups);
    const int c0 = g * kGroup, c1 = min(c0 + kGroup, a.n_chunks);
    const float* m = a.Mw + ((int64_t)item * a.n_chunks + c0) * kMapFloats;
    float col[kState];
#pragma unroll
    for (int i = 0; i < kState; ++i) col[i] = (i == r) ? 1.0f : 0.0f;      // r == 7: the z column starts at 0
    for (int c = c0; c < c1; ++c, m += kMapFloats) {
        float nxt[kState];
#pragma unroll
        for (int i = 0; i < kState; ++i) nxt[i] = (r == 7) ? __ldg(m + 7 * kState + i) : 0.0f;
#pragma unroll
        for (int k = 0; k < kState; ++k) {
#pragma unroll
            for (int i = 0; i < kState; ++i) nxt[i] = fmaf(__ldg(m + k * kState + i), col[k], nxt[i]);
        }
#pragma unroll
        for (int i = 0; i < kState; ++i) col[i] = nxt[i];
    }
    float* o = GM + team * kMapFloats + r * kState;
#pragma unroll
    for (int i = 0; i < kState; ++i) o[i] = col[i];
}

// (b), (c) chain maps [seg * seg_len, (seg + 1) * seg_len) of every example from an initial state
// (zero when init == nullptr), writing the state at the start of every map.  8 lanes per segment;
// lane i < 7 owns state component i.
__global__ void __launch_bounds__(256) phaser_chain_kernel(const float* __restrict__ maps, int n_maps, int seg_len,
                                                           int n_items, const float* __restrict__ init,
                                                           float* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = (int)(gid & 7);
    const int n_seg = (n_maps + seg_len - 1) / seg_len;
    int64_t team = gid >> 3;
    const bool live = team < (int64_t)n_items * n_seg;
    if (!live) team = 0;
    const int item = (int)(team / n_seg), seg = (int)(team - (int64_t)item * n_seg);
    const int c0 = seg * seg_len, c1 = min(c0 + seg_len, n_maps);
    const float* m = maps + ((int64_t)item * n_maps + c0) * kMapFloats;
    float* S = out + ((int64_t)item * n_maps + c0) * kSFloats;
    const int ii = (i < kState) ? i : 0;
    const int n = c1 - c0;
    constexpr int kAhead = 4;                  // maps in flight (their coefficients do not depend on the state)
    float buf[kAhead][8];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
        const float* mc = m + (int64_t)min(d, n - 1) * kMapFloats;
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[d][r] = mc[r * kState + ii];
    }
    float s = (init && i < kState) ? init[((int64_t)item * n_seg + seg) * kSFloats + i] : 0.0f;
    // every team of the warp runs the same number of steps (the shuffles below need all 32 lanes);
    // steps past the end of a short last segment are computed and discarded
    for (int b0 = 0; b0 < seg_len; b0 += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int cidx = b0 + d;
            float cur[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) cur[r] = buf[d][r];
            {   // refill this slot with map cidx + kAhead
                const float* mc = m + (int64_t)min(cidx + kAhead, n - 1) * kMapFloats;
#pragma unroll
                for (int r = 0; r < 8; ++r) buf[d][r] = mc[r * kState + ii];
            }
            float acc = cur[7];                                        // zero-state response
#pragma unroll
            for (int r = 0; r < kState; ++r) acc = fmaf(cur[r], __shfl_sync(kFull, s, r, 8), acc);
            if (cidx < n) {
                if (live) S[cidx * kSFloats + i] = s;
                s = (i < kState) ? acc : 0.0f;
            }
        }
    }
}

// ---- K3: re-run every chunk from its true state, mix and clip --------------------------------------
// (Packing two chunks per thread like K1 was measured: 0.56 ms instead of 0.36 ms for 1365 x 2 s, because it
// halves the warps that hide the tile loads; this kernel is latency-bound, not issue-bound.)
// One thread per chunk; a warp owns 32 consecutive chunks (16 KB of audio).  The audio moves through a
// per-warp 32 x 32 shared tile: row i is filled by one coalesced 128-byte request (all lanes read chunk
// i), then lane i walks row i.  That keeps every global request at one cache line (a lane-per-chunk
// access would touch 32 lines per request and saturate the L1 tag stage).
constexpr int kRunWarps = 4;
__global__ void __launch_bounds__(32 * kRunWarps) phaser_run_kernel(const PhaserArgs a) {
    __shared__ float tile[kRunWarps][32][33];
    __shared__ float ctile[kRunWarps][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * kRunWarps + w;          // warp id = group of 32 chunks of one example
    const int groups = (a.n_chunks + 31) / 32;
    if (wid >= (int64_t)a.n_items * groups) return;
    const int item = (int)(wid / groups), grp = (int)(wid - (int64_t)item * groups);
    const int b = example_of(a, item);
    const int chunk0 = grp * 32;
    const int chunk = chunk0 + lane;
    const bool live = chunk < a.n_chunks;
    const int n_base = chunk0 * kChunk;                                // first sample of the warp's span
    const float* xr = a.x + (int64_t)b * a.N;
    float* yr = a.y + (int64_t)b * a.N;
    const float* cr = a.C + (int64_t)item * a.n_ctl;
    float(*t)[33] = tile[w];
    float(*ct)[33] = ctile[w];
    // coefficients of the 32 chunks: row i = chunk i's 32 control points
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        const int j = (n_base >> 2) + i * kCtl + lane;
        ct[i][lane] = (j < a.n_ctl) ? __ldg(cr + j) : 0.0f;
    }
    const float fbk = a.feedback[b];
    const float wet = a.mix[b], dry = 1.0f - a.mix[b];
    float s[kStagesAP];
    float last = 0.0f;
    if (live) {
        const float* S = a.S + ((int64_t)item * a.n_chunks + chunk) * kSFloats;
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) s[k] = S[k];
        last = S[6];
    } else {
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) s[k] = 0.0f;
    }
    const int len = live ? min(kChunk, a.N - chunk * kChunk) : 0;
    float c_cur = cr[max(0, min(chunk * kCtl - 1, a.n_ctl - 1))];      // coefficient of the sample before the chunk
    for (int q0 = 0; q0 < kChunk; q0 += 32) {                          // 4 sub-tiles of 32 samples
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int n = n_base + i * kChunk + q0 + lane;
            t[i][lane] = (n < a.N) ? __ldg(xr + n) : 0.0f;
        }
        __syncwarp();
        if (q0 + 32 <= len) {
#pragma unroll 2
            for (int g = 0; g < 8; ++g) {
                const float c = ct[lane][(q0 >> 2) + g];
                cascade_retune(s, c_cur, c);
                c_cur = c;
#pragma unroll
                for (int q = 0; q < kUpd; ++q) {
                    const float in = t[lane][g * kUpd + q];
                    const float out = cascade_step(in - last, s, c);
                    last = out * fbk;
                    t[lane][g * kUpd + q] = fminf(fmaxf(fmaf(wet, out, dry * in), -1.0f), 1.0f);   // datasets.py:472
                }
            }
        } else {
            for (int q = 0; q0 + q < len && q < 32; ++q) {
                const float c = ct[lane][(q0 + q) >> 2];
                if (c != c_cur) {
                    cascade_retune(s, c_cur, c);
                    c_cur = c;
                }
                const float in = t[lane][q];
                const float out = cascade_step(in - last, s, c);
                last = out * fbk;
                t[lane][q] = fminf(fmaxf(fmaf(wet, out, dry * in), -1.0f), 1.0f);
            }
        }
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int n = n_base + i * kChunk + q0 + lane;
            if (n < a.N) yr[n] = t[i][lane];
        }
    }
}

// =====================================================================================================================
// Single-pass fused phaser (the default path): one kernel reads the audio once and writes the result once.
//
// The four-kernel pipeline above moves 2.3x the algorithmic bytes (the audio is read twice, the control-rate coefficients
// and the per-chunk maps round-trip HBM).  Here a CTA owns one SEGMENT of 4096 samples of one example and does everything
// for it out of shared memory:
//   A  the segment's audio lands in shared memory (one coalesced pass); the float32 oscillator phases of its 1024
//      control points are walked chunk by chunk from chunk-start phases (phaser_phase_kernel below: a few bytes per
//      128 samples, the only thing that is precomputed) -- sequential adds with the wrap at 2 pi, like the restatement;
//   B  the transcendental map phase -> all-pass coefficient c = 2G - 1, all threads;
//   C  per 64-sample sub-chunk the 7 unit-state runs + the zero-state run (two runs per thread, packed FFMA2) give its
//      affine map; lanes = sub-chunks, so a warp's 32 lanes walk 32 different sub-chunks (padded rows: no conflicts);
//   D  the state at the segment start comes from the CTA of the previous segment of the same example through global
//      memory (decoupled look-back: CTAs take their (segment, example) from an atomic ticket in segment-major order, so a
//      predecessor always holds an earlier ticket and is running or done: no deadlock); 8 lanes chain the 64 maps, the
//      exit state is published for the next segment at once;
//   E  64 threads re-run their sub-chunk from its true entry state, mix, clip;
//   F  coalesced store -- optionally only the window [start[b], start[b] + n_out) of the row, together with the same
//      window of the dry input: the random crop of PedalboardPhaserDataset.__getitem__ (datasets.py:445-447).
// Segments past the last needed sample are never touched.  HBM traffic: 4 B in + 4 B out per sample + 64 B per segment.
constexpr int kSegChunks = 32;                       // 128-sample chunks per segment
constexpr int kSeg = kSegChunks * kChunk;            // 4096 samples
constexpr int kSub = 64;                             // samples per scan sub-chunk
constexpr int kSubs = kSeg / kSub;                   // 64 sub-chunks per segment
constexpr int kSubCtl = kSub / kUpd;                 // 16 control points per sub-chunk
constexpr int kSegCtl = kSeg / kUpd;                 // 1024
constexpr int kFusedThreads = 256;
constexpr int kRec = 8;                              // published state record: w1..w6, lastOutput, c the states are scaled for

struct FusedArgs {
    PhaserArgs a;
    const int32_t* start;       // (B,) first delivered sample of each row, or nullptr (0)
    int n_out;                  // delivered samples per row
    float* dry_out;             // (B, n_out) window of the dry input, or nullptr
    int x_compact;              // 1: row `item` of x (a compact (n_items, N) array), 0: row b like every other array
    const int64_t* x_offset;    // (n_items,) or nullptr: x is a packed ragged array, row `item` starts at x + x_offset[item]
                                // and holds start + n_out samples (the prefix that determines the delivered window)
    int n_seg;
    int* ticket;                // 1 int, zero at launch
    int* flags;                 // (n_items, n_seg), zero at launch
    float* rec;                 // (n_items, n_seg, kRec)
    float2* ph;                 // (n_items, n_chunks): phase at the chunk's first control point, phase of the one before
};

// chunk-start phases: one thread per (example, host block), juce::dsp::Oscillator semantics (see phaser_ctl_kernel).
// The float32 phase is a chain of sequential additions with the wrap at 2 pi, so a host block of 8192 samples is 2048
// dependent steps whatever is done -- what can be kept short is the step: when the increment is below 2 pi (always, for
// an LFO) the wrap is ONE conditional subtraction (p + inc < 4 pi), computed branch-free (add, subtract, select: 12 cycles
// instead of a data-dependent loop), and every block is walked exactly once: the phase of a block's last control point,
// which the next block's first chunk needs as its "control point before", is written there by the thread that walks it.
__device__ __forceinline__ float phase_advance(float p, float inc, float two_pi) {
    const float next = p + inc;
    const float wrapped = next - two_pi;
    return (next >= two_pi) ? wrapped : next;
}

__global__ void __launch_bounds__(128) phaser_phase_kernel(const FusedArgs f, int n_blocks) {
    const PhaserArgs& a = f.a;
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= a.n_items * n_blocks) return;
    const int item = pair / n_blocks, blk = pair - item * n_blocks;
    const int b = example_of(a, item);
    const int need = f.start ? min(a.N, f.start[b] + f.n_out) : a.N;
    // (a block past the window still has to hand its last phase to nobody: nothing after `need` is rendered)
    if ((int64_t)blk * a.block >= need) return;
    const float two_pi = MODFX_TWO_PI_F;
    const float sr_down = (float)((double)a.sr / (double)kUpd);
    const float inc = a.rate[b] * (two_pi / sr_down);
    float p = 0.0f;
    for (int i = 0; i < blk; ++i) {                  // whole blocks before this one: phase.advance(freq * n_down)
        float next = p + inc * (float)(a.block / kUpd);
        while (next >= two_pi) next -= two_pi;
        p = next;
    }
    const int s0 = blk * a.block, s1 = min(s0 + a.block, a.N);
    const int c0 = s0 / kChunk, c1 = (s1 + kChunk - 1) / kChunk;
    float2* out = f.ph + (int64_t)item * a.n_chunks;
    // a clip's very first control point has no predecessor and the value is unused (all filter states are zero there)
    float prev = 0.0f;
    const bool small = inc < two_pi;
    for (int c = c0; c < c1; ++c) {
        out[c].x = p;
        if (c > c0 || blk == 0) out[c].y = prev;     // the first chunk of a later block: written by the previous block's thread
        if (small) {
#pragma unroll 8
            for (int j = 0; j < kCtl; ++j) {
                prev = p;
                p = phase_advance(p, inc, two_pi);
            }
        } else {
            for (int j = 0; j < kCtl; ++j) {
                prev = p;
                float next = p + inc;
                while (next >= two_pi) next -= two_pi;
                p = next;
            }
        }
    }
    // this block's last control point is the one before the next block's first (only a full block has a successor)
    if (s1 - s0 == a.block && c1 < a.n_chunks) out[c1].y = prev;
}

__device__ __forceinline__ float phaser_coef(float phase, float vol, float ctr, float log_span, float log_min, float w0) {
    float lfo = sinf(phase - MODFX_PI_F) * vol + ctr;
    lfo = fminf(fmaxf(lfo, 0.0f), 1.0f);
    const float fc = exp10f(lfo * log_span + log_min);                   // mapToLog10
    const float g = tanf(w0 * fc);                                       // TPT prewarp
    return 2.0f * (g / (1.0f + g)) - 1.0f;
}

__global__ void __launch_bounds__(kFusedThreads) phaser_fused_kernel(const FusedArgs f) {
    extern __shared__ __align__(16) float sm[];
    float(*xs)[kSub + 1] = reinterpret_cast<float(*)[kSub + 1]>(sm);                        // [64][65] audio, then output
    float(*cs)[kSubCtl + 1] = reinterpret_cast<float(*)[kSubCtl + 1]>(sm + kSubs * (kSub + 1));   // [64][17] col 0: coefficient before
    float(*Ms)[kMapFloats + 1] = reinterpret_cast<float(*)[kMapFloats + 1]>(sm + kSubs * (kSub + 1) + kSubs * (kSubCtl + 1));   // [64][57]
    float(*Ss)[kSFloats] = reinterpret_cast<float(*)[kSFloats]>(sm + kSubs * (kSub + 1) + kSubs * (kSubCtl + 1) + kSubs * (kMapFloats + 1));
    float* phs = reinterpret_cast<float*>(Ms);                           // [1024 + 1] phases (dead before the maps are written)
    __shared__ int tile_s;
    __shared__ float entry[kRec];
    const PhaserArgs& a = f.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) tile_s = atomicAdd(f.ticket, 1);
    __syncthreads();
    const int tile = tile_s;
    const int seg = tile / a.n_items, item = tile - seg * a.n_items;     // segment-major: predecessors hold earlier tickets
    const int b = example_of(a, item);
    const int start = f.start ? f.start[b] : 0;
    const int need = min(a.N, start + f.n_out);
    const int n_base = seg * kSeg;
    if (n_base >= need) return;
    const float* xr = f.x_offset ? a.x + f.x_offset[item] : a.x + (int64_t)(f.x_compact ? item : b) * a.N;
    const int row_len = f.x_offset ? need : a.N;                          // samples that exist in this row

    // ---- A: audio + oscillator phases
    const bool vec = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) && (n_base + kSeg <= row_len);
    if (vec) {                                                            // 16-byte loads, 4 per thread
#pragma unroll
        for (int k = 0; k < kSeg / 4 / kFusedThreads; ++k) {
            const int i4 = tid + k * kFusedThreads;
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr + n_base) + i4);
            float* d = &xs[i4 >> 4][(i4 & 15) << 2];
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    } else {
        for (int i = tid; i < kSeg; i += kFusedThreads) {
            const int n = n_base + i;
            xs[i / kSub][i % kSub] = (n < row_len) ? __ldg(xr + n) : 0.0f;
        }
    }
    const float two_pi = MODFX_TWO_PI_F;
    if (tid < kSegChunks) {
        const int chunk = seg * kSegChunks + tid;
        const float sr_down = (float)((double)a.sr / (double)kUpd);
        const float inc = a.rate[b] * (two_pi / sr_down);
        float2 pp = (chunk < a.n_chunks) ? f.ph[(int64_t)item * a.n_chunks + chunk] : make_float2(0.0f, 0.0f);
        if (tid == 0) phs[0] = pp.y;                                      // control point before the segment
        float p = pp.x;
        if (inc < two_pi) {
#pragma unroll 8
            for (int q = 0; q < kCtl; ++q) {
                phs[1 + tid * kCtl + q] = p;
                p = phase_advance(p, inc, two_pi);
            }
        } else {
            for (int q = 0; q < kCtl; ++q) {
                phs[1 + tid * kCtl + q] = p;
                float next = p + inc;
                while (next >= two_pi) next -= two_pi;
                p = next;
            }
        }
    }
    __syncthreads();
    // ---- B: phase -> all-pass coefficient
    {
        const float f_lo = 20.0f;
        const double f_hi_d = (0.49 * (double)a.sr < 20000.0) ? 0.49 * (double)a.sr : 20000.0;
        const float log_min = log10f(f_lo), log_span = log10f((float)f_hi_d) - log_min;
        const float w0 = MODFX_PI_F / a.sr;
        const float vol = a.depth[b] * 0.5f;
        const float ctr = (log10f(a.centre[b]) - log_min) / log_span;
        float cv[(kSegCtl + 1 + kFusedThreads - 1) / kFusedThreads];
#pragma unroll
        for (int k = 0; k < (kSegCtl + 1 + kFusedThreads - 1) / kFusedThreads; ++k) {
            const int j = tid + k * kFusedThreads;                        // j = 0: the control point before the segment
            cv[k] = (j <= kSegCtl) ? phaser_coef(phs[j], vol, ctr, log_span, log_min, w0) : 0.0f;
        }
        __syncthreads();                                                  // phs aliases Ms: all reads done before cs / Ms are written
#pragma unroll
        for (int k = 0; k < (kSegCtl + 1 + kFusedThreads - 1) / kFusedThreads; ++k) {
            const int j = tid + k * kFusedThreads;
            if (j <= kSegCtl) {
                if (j >= 1) cs[(j - 1) / kSubCtl][1 + (j - 1) % kSubCtl] = cv[k];
                if (j % kSubCtl == 0 && j < kSegCtl) cs[j / kSubCtl][0] = cv[k];     // coefficient before sub-chunk j / 16
            }
        }
    }
    __syncthreads();
    const float fbk = a.feedback[b];
    // ---- C: affine map of every sub-chunk (runs 2p, 2p+1 per thread; run 6 = unit lastOutput, run 7 = zero state + audio)
    {
        const int pair = warp >> 1, sc = ((warp & 1) << 5) | lane;
        float2 s[kStagesAP];
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) s[k] = make_float2((2 * pair == k) ? 1.0f : 0.0f, (2 * pair + 1 == k) ? 1.0f : 0.0f);
        float2 out_prev = make_float2(0.0f, 0.0f);
        const float2 nfb = make_float2(-fbk, -fbk);
        float c_cur = cs[sc][0];
        int g0 = 0;
        if (pair == 3) {
            // unit lastOutput = out_prev of 1 / feedback: fed as u = -1 by hand for the first sample (finite for feedback 0);
            // both runs start with zero all-pass states, so nothing to retune before the first control group
            const float c0 = cs[sc][1];
            out_prev = cascade_step2(make_float2(-1.0f, xs[sc][0]), s, c0);
#pragma unroll
            for (int q = 1; q < kUpd; ++q)
                out_prev = cascade_step2(__ffma2_rn(nfb, out_prev, make_float2(0.0f, xs[sc][q])), s, c0);
            c_cur = c0;
            g0 = 1;
        }
#pragma unroll 2
        for (int g = g0; g < kSubCtl; ++g) {
            const float c = cs[sc][1 + g];
            cascade_retune2(s, c_cur, c);
            c_cur = c;
#pragma unroll
            for (int q = 0; q < kUpd; ++q) {
                const float2 u = (pair == 3) ? __ffma2_rn(nfb, out_prev, make_float2(0.0f, xs[sc][g * kUpd + q]))
                                             : __fmul2_rn(nfb, out_prev);
                out_prev = cascade_step2(u, s, c);
            }
        }
        float* m = &Ms[sc][2 * pair * kState];
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) {
            m[k] = s[k].x;
            m[kState + k] = s[k].y;
        }
        m[6] = out_prev.x * fbk;
        m[kState + 6] = out_prev.y * fbk;
    }
    __syncthreads();
    // ---- D: entry state of the segment (look-back), entry state of every sub-chunk, exit state published.
    // Three short passes instead of one chain of 64 dependent steps: (1) every warp composes the 8 maps of its group
    // into one, all groups at once; (2) warp 0 takes the segment's entry state from the previous segment's CTA and
    // chains the 8 group maps; (3) every warp chains its 8 sub-chunks from its group's entry state.
    float(*GM)[kMapFloats + 1] = reinterpret_cast<float(*)[kMapFloats + 1]>(Ss + kSubs);       // [8][57] group maps
    float(*GS)[kSFloats] = reinterpret_cast<float(*)[kSFloats]>(reinterpret_cast<float*>(GM) + 8 * (kMapFloats + 1));   // [8][8]
    {
        // (1) lane = (row i = lane >> 3 and i + 4, column r = lane & 7) of the running product [Phi | z]
        const int r = lane & 7, i0 = lane >> 3, i1 = i0 + 4;
        float* g = GM[warp];
        {
            const float* m = Ms[warp * 8];
            g[r * kState + i0] = m[r * kState + i0];
            if (i1 < kState) g[r * kState + i1] = m[r * kState + i1];
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 1; k < 8; ++k) {
            const float* m = Ms[warp * 8 + k];
            float a0 = (r == 7) ? m[7 * kState + i0] : 0.0f;
            float a1 = (r == 7 && i1 < kState) ? m[7 * kState + i1] : 0.0f;
#pragma unroll
            for (int j = 0; j < kState; ++j) {
                const float cj = g[r * kState + j];
                a0 = fmaf(m[j * kState + i0], cj, a0);
                if (i1 < kState) a1 = fmaf(m[j * kState + i1], cj, a1);
            }
            __syncwarp();
            g[r * kState + i0] = a0;
            if (i1 < kState) g[r * kState + i1] = a1;
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == 0) {
        float* rec = f.rec + ((int64_t)item * f.n_seg + seg) * kRec;
        if (seg > 0) {
            const int* flag = f.flags + (int64_t)item * f.n_seg + seg - 1;
            if (lane == 0) {
                while (atomicAdd(const_cast<int*>(flag), 0) == 0) __nanosleep(64);
                __threadfence();
            }
            __syncwarp();
            if (lane < kRec) entry[lane] = __ldcg(rec - kRec + lane);
        } else if (lane < kRec) {
            entry[lane] = 0.0f;
        }
        __syncwarp();
        // (2) the predecessor left its all-pass states scaled for ITS last coefficient, which is this segment's cs[0][0]
        const int i = lane & 7, ii = (i < kState) ? i : 0;
        float sv = (i < kState) ? entry[i] : 0.0f;
        if (lane < 8) {
#pragma unroll 1
            for (int gq = 0; gq < 8; ++gq) {
                GS[gq][i] = sv;
                const float* m = GM[gq];
                float acc = m[7 * kState + ii];
#pragma unroll
                for (int r = 0; r < kState; ++r) acc = fmaf(m[r * kState + ii], __shfl_sync(0xffu, sv, r, 8), acc);
                sv = (i < kState) ? acc : 0.0f;
            }
            if (i < kState) rec[i] = sv;
            else rec[7] = cs[kSubs - 1][kSubCtl];
            __threadfence();
        }
        __syncwarp();
        if (lane == 0) atomicExch(f.flags + (int64_t)item * f.n_seg + seg, 1);
    }
    // the window of the dry input goes out while warp 0 chains (xs still holds the audio)
    if (f.dry_out) {
        float* dr = f.dry_out + (int64_t)b * f.n_out;
        for (int i = tid; i < kSeg; i += kFusedThreads) {
            const int idx = n_base + i - start;
            if (idx >= 0 && idx < f.n_out && n_base + i < row_len) dr[idx] = xs[i / kSub][i % kSub];
        }
    }
    __syncthreads();
    // (3) entry state of every sub-chunk: each warp chains its group (8 lanes, lane i < 7 owns state component i)
    if (lane < 8) {
        const int i = lane, ii = (i < kState) ? i : 0;
        float sv = (i < kState) ? GS[warp][i] : 0.0f;
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
            const int sc = warp * 8 + k;
            Ss[sc][i] = sv;
            const float* m = Ms[sc];
            float acc = m[7 * kState + ii];
#pragma unroll
            for (int r = 0; r < kState; ++r) acc = fmaf(m[r * kState + ii], __shfl_sync(0xffu, sv, r, 8), acc);
            sv = (i < kState) ? acc : 0.0f;
        }
    }
    __syncthreads();
    // ---- E: re-run every sub-chunk from its true entry state, mix and clip (datasets.py:472)
    if (tid < kSubs) {
        const int sc = tid;
        float s[kStagesAP];
#pragma unroll
        for (int k = 0; k < kStagesAP; ++k) s[k] = Ss[sc][k];
        float last = Ss[sc][6];
        const float wet = a.mix[b], dry = 1.0f - a.mix[b];
        float c_cur = cs[sc][0];
#pragma unroll 2
        for (int g = 0; g < kSubCtl; ++g) {
            const float c = cs[sc][1 + g];
            cascade_retune(s, c_cur, c);
            c_cur = c;
#pragma unroll
            for (int q = 0; q < kUpd; ++q) {
                const float in = xs[sc][g * kUpd + q];
                const float out = cascade_step(in - last, s, c);
                last = out * fbk;
                xs[sc][g * kUpd + q] = fminf(fmaxf(fmaf(wet, out, dry * in), -1.0f), 1.0f);
            }
        }
    }
    __syncthreads();
    // ---- F: coalesced store of the delivered window
    float* yr = a.y + (int64_t)b * f.n_out;
    if (vec && start == 0 && n_base + kSeg <= f.n_out && ((reinterpret_cast<uintptr_t>(yr) & 15) == 0)) {
#pragma unroll
        for (int k = 0; k < kSeg / 4 / kFusedThreads; ++k) {
            const int i4 = tid + k * kFusedThreads;
            const float* d = &xs[i4 >> 4][(i4 & 15) << 2];
            reinterpret_cast<float4*>(yr + n_base)[i4] = make_float4(d[0], d[1], d[2], d[3]);
        }
    } else {
        for (int i = tid; i < kSeg; i += kFusedThreads) {
            const int n = n_base + i;
            const int idx = n - start;
            if (idx >= 0 && idx < f.n_out && n < a.N) yr[idx] = xs[i / kSub][i % kSub];
        }
    }
}

int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace modfx

using namespace modfx;

extern "C" int64_t modfx_phaser_workspace_bytes(int32_t B, int64_t N) {
    if (B <= 0 || N <= 0) return 0;
    const int64_t n_ctl = (N + kUpd - 1) / kUpd, n_chunks = (N + kChunk - 1) / kChunk;
    const int64_t multi = align_up((int64_t)B * n_ctl * 4, 256) + align_up((int64_t)B * n_chunks * kMapFloats * 4, 256) +
                          align_up((int64_t)B * n_chunks * kSFloats * 4, 256) +
                          align_up((int64_t)B * ((n_chunks + 31) / 32) * kMapFloats * 4, 256) +
                          align_up((int64_t)B * ((n_chunks + 31) / 32) * kSFloats * 4, 256);
    const int64_t n_seg = (N + kSeg - 1) / kSeg;
    const int64_t fused = 256 + align_up((int64_t)B * n_seg * 4, 256) + align_up((int64_t)B * n_seg * kRec * 4, 256) +
                          align_up((int64_t)B * n_chunks * 8, 256);
    return (multi > fused ? multi : fused) + 256;
}

namespace {

int phaser_launch(const float* x, int x_compact, const int64_t* x_offset, float* y, float* dry_out, int32_t B, int64_t N, int64_t n_out, const int32_t* start,
                  float sr, const float* rate_hz, const float* depth, const float* centre_hz, const float* feedback,
                  const float* mix, int32_t block, const int32_t* example_index, int32_t n_items, void* workspace,
                  void* stream) {
    MODFX_REQUIRE(x && y && rate_hz && depth && centre_hz && feedback && mix, "NULL pointer");
    MODFX_REQUIRE(B >= 0 && N >= 1 && N < (1ll << 30), "bad shape B=%d N=%lld", B, (long long)N);
    MODFX_REQUIRE(n_out >= 1 && n_out <= N, "bad output window n_out=%lld (row length %lld)", (long long)n_out, (long long)N);
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
    cudaStream_t s = as_stream(stream);
    char* w = static_cast<char*>(workspace);
    const int n_blocks = (int)((N + block - 1) / block);

    const char* force = getenv("MODFX_PHASER_KERNEL");
    const bool fused_ok = (block % kChunk == 0) && !(force && force[0] == 'm');
    if (fused_ok) {
        // ---- single-pass fused kernel
        FusedArgs f{};
        f.a = a;
        f.start = start; f.n_out = (int)n_out; f.dry_out = dry_out; f.x_compact = x_compact; f.x_offset = x_offset;
        f.n_seg = (int)((N + kSeg - 1) / kSeg);
        f.ticket = reinterpret_cast<int*>(w);                                     w += 256;
        f.flags = reinterpret_cast<int*>(w);
        const int64_t flag_bytes = align_up((int64_t)a.n_items * f.n_seg * 4, 256);  w += flag_bytes;
        f.rec = reinterpret_cast<float*>(w);                                      w += align_up((int64_t)a.n_items * f.n_seg * kRec * 4, 256);
        f.ph = reinterpret_cast<float2*>(w);
        MODFX_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)(256 + flag_bytes), s));
        const int pairs = a.n_items * n_blocks;
        phaser_phase_kernel<<<(pairs + 127) / 128, 128, 0, s>>>(f, n_blocks);
        const size_t smem = sizeof(float) * (size_t)(kSubs * (kSub + 1) + kSubs * (kSubCtl + 1) + kSubs * (kMapFloats + 1) + kSubs * kSFloats +
                                                     8 * (kMapFloats + 1) + 8 * kSFloats);
        const int64_t tiles = (int64_t)a.n_items * f.n_seg;
        MODFX_REQUIRE(tiles < (1ll << 31), "too many segments in one call");
        phaser_fused_kernel<<<(unsigned)tiles, kFusedThreads, smem, s>>>(f);
        MODFX_CUDA_OK(cudaGetLastError());
        return MODFX_OK;
    }
    // ---- four-kernel pipeline (host blocks that are not a multiple of 128 samples; MODFX_PHASER_KERNEL=multi)
    if (start || dry_out || n_out != N || x_compact || x_offset)
        return fail(MODFX_ERR_UNSUPPORTED, "cropped output needs a host block size that is a multiple of %d samples", kChunk);
    a.C = reinterpret_cast<float*>(w);
    w += align_up((int64_t)a.n_items * a.n_ctl * 4, 256);
    a.Mw = reinterpret_cast<float*>(w);
    w += align_up((int64_t)a.n_items * a.n_chunks * kMapFloats * 4, 256);
    a.S = reinterpret_cast<float*>(w);
    w += align_up((int64_t)a.n_items * a.n_chunks * kSFloats * 4, 256);
    {
        const int total = a.n_items * n_blocks;
        phaser_ctl_kernel<<<(total + 31) / 32, kK0Threads, 0, s>>>(a, n_blocks);
    }
    {
        const int n_super = (a.n_chunks + 31) / 32;
        phaser_map_kernel<<<(unsigned)((int64_t)a.n_items * n_super), kMapThreads, 0, s>>>(a, n_super);
    }
    if (a.n_chunks <= kSerialChainMax) {
        // short clips (689 maps for 2 s): one serial chain per example is the cheapest
        phaser_chain_kernel<<<(unsigned)(((int64_t)a.n_items * 8 + 255) / 256), 256, 0, s>>>(a.Mw, a.n_chunks, a.n_chunks,
                                                                                          a.n_items, nullptr, a.S);
    } else {
        // long clips (20 672 maps for 60 s): compose groups in parallel, chain the groups, re-chain inside groups
        const int n_groups = (a.n_chunks + kGroup - 1) / kGroup;
        float* GM = reinterpret_cast<float*>(w);
        w += align_up((int64_t)a.n_items * n_groups * kMapFloats * 4, 256);
        float* GS = reinterpret_cast<float*>(w);
        const int64_t teams = (int64_t)a.n_items * n_groups;
        phaser_compose_kernel<<<(unsigned)((teams * 8 + 255) / 256), 256, 0, s>>>(a, n_groups, GM);
        phaser_chain_kernel<<<(unsigned)(((int64_t)a.n_items * 8 + 255) / 256), 256, 0, s>>>(GM, n_groups, n_groups, a.n_items,
                                                                                          nullptr, GS);
        phaser_chain_kernel<<<(unsigned)((teams * 8 + 255) / 256), 256, 0, s>>>(a.Mw, a.n_chunks, kGroup, a.n_items, GS, a.S);
    }
    {
        const int64_t warps = (int64_t)a.n_items * ((a.n_chunks + 31) / 32);
        phaser_run_kernel<<<(unsigned)((warps + kRunWarps - 1) / kRunWarps), 32 * kRunWarps, 0, s>>>(a);
    }
    MODFX_CUDA_OK(cudaGetLastError());
    return MODFX_OK;
}

}  // namespace

extern "C" int modfx_phaser_f32(const float* x, float* y, int32_t B, int64_t N, float sr, const float* rate_hz,
                                const float* depth, const float* centre_hz, const float* feedback,
                                const float* mix, int32_t block, const int32_t* example_index, int32_t n_items,
                                void* workspace, void* stream) {
    return phaser_launch(x, 0, nullptr, y, nullptr, B, N, N, nullptr, sr, rate_hz, depth, centre_hz, feedback, mix, block, example_index,
                         n_items, workspace, stream);
}

extern "C" int modfx_phaser_crop_f32(const float* x, int32_t x_compact, float* y, float* dry_out, int32_t B, int64_t N,
                                     int64_t n_out, const int32_t* start, float sr, const float* rate_hz, const float* depth,
                                     const float* centre_hz, const float* feedback, const float* mix, int32_t block,
                                     const int32_t* example_index, int32_t n_items, void* workspace, void* stream) {
    MODFX_REQUIRE(start, "start is NULL");
    MODFX_REQUIRE(!x_compact || example_index, "x_compact needs an example_index list");
    return phaser_launch(x, x_compact ? 1 : 0, nullptr, y, dry_out, B, N, n_out, start, sr, rate_hz, depth, centre_hz, feedback, mix, block, example_index,
                         n_items, workspace, stream);
}

extern "C" int modfx_phaser_crop_packed_f32(const float* x, const int64_t* x_offset, float* y, float* dry_out, int32_t B,
                                            int64_t N_max, int64_t n_out, const int32_t* start, float sr, const float* rate_hz,
                                            const float* depth, const float* centre_hz, const float* feedback, const float* mix,
                                            int32_t block, const int32_t* example_index, int32_t n_items, void* workspace,
                                            void* stream) {
    MODFX_REQUIRE(start && x_offset, "start or x_offset is NULL");
    MODFX_REQUIRE(example_index, "a packed x needs an example_index list (row i belongs to example example_index[i])");
    return phaser_launch(x, 1, x_offset, y, dry_out, B, N_max, n_out, start, sr, rate_hz, depth, centre_hz, feedback, mix, block,
                         example_index, n_items, workspace, stream);
}
