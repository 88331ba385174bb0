/*
 * modfx_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the arithmetic of the effect-rendering hot path of
 * christhetree/mod_extraction.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (mod_extraction_b200/) never does.
 *
 * Every function cites the reference file:line it restates.  The reference is
 * pure Python/torch; all of its per-element float32 operations are separate
 * torch ops (no FMA contraction) unless noted, so this file must be built with
 * -ffp-contract=off and uses fmaf() only where the torch CPU kernel is known
 * to contract (upsample_linear1d, see modfx_oracle_interp_linear).
 *
 * Parity pinning: fx.py / modulations.py / util.py paths are pinned against
 * golden vectors generated from the reference itself (tests/golden/, made by
 * tests/golden/make_golden.py).  The phaser (modfx_oracle_phaser) restates
 * JUCE dsp::Phaser as wrapped by pedalboard 0.7.3, which is NOT available
 * offline: PARITY UNPINNED for the phaser.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define MODFX_SHAPE_COS 0
#define MODFX_SHAPE_RECT_COS 1
#define MODFX_SHAPE_INV_RECT_COS 2
#define MODFX_SHAPE_TRI 3
#define MODFX_SHAPE_SAW 4
#define MODFX_SHAPE_RSAW 5
#define MODFX_SHAPE_SQR 6

/* ---- minimal pthread parallel-for over independent examples ------------- */
static int g_threads = 0;

int modfx_oracle_num_threads(void) {
    if (g_threads > 0) return g_threads;
    const char* e = getenv("MODFX_ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 512) n = 512;
    return (int)n;
}

void modfx_oracle_set_threads(int n) { g_threads = n; }

typedef void (*item_fn)(int idx, void* ctx);
typedef struct { item_fn fn; void* ctx; int n; volatile int* next; } pf_job;

static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        const int i = __sync_fetch_and_add(j->next, 1);
        if (i >= j->n) break;
        j->fn(i, j->ctx);
    }
    return NULL;
}

static void parallel_for(int n, item_fn fn, void* ctx) {
    int nt = modfx_oracle_num_threads();
    if (nt > n) nt = n;
    volatile int next = 0;
    pf_job job = { fn, ctx, n, &next };
    if (nt <= 1) { pf_worker(&job); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    for (int t = 0; t < nt - 1; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* torch.remainder(a, b) for float32, b > 0 (ATen BinaryOpsKernel remainder:
 * fmod then shift into the sign of the divisor). */
static inline float torch_remainder_f32(float a, float b) {
    float m = fmodf(a, b);
    if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m = m + b;
    return m;
}

/*
 * MonoFlangerChorusModule.apply_effect, fx.py:72-119.
 *
 * x      (B, C, N)   dry audio
 * mod    (B, N) if mod_has_ch == 0, else (B, C, N)           fx.py:84-85
 * y      (B, C, N)   wet audio
 * Per-example derived coefficients (computed by the caller with the
 * reference's float-vs-tensor promotion rules, see oracle/oracle.py):
 *   lfo_delay[b]      = max_lfo_delay_samples * width         fx.py:98
 *   min_delay[b]      = min_delay_width * max_min_delay_samples  fx.py:97
 *   fb[b], depth[b], mix[b], one_minus_mix[b] = (1.0 - mix)   fx.py:114-117
 */
typedef struct {
    const float* x; const float* mod; int mod_has_ch; float* y;
    int B, C; int64_t N; int Mmin, Mlfo;
    const float *lfo_delay, *min_delay, *fb, *depth, *mix, *one_minus_mix;
    int interp;     /* 0: linear (the reference, fx.py:113); 1: first-order all-pass interpolation (OWN definition, see below) */
} fc_ctx;

static void fc_item(int bc, void* vctx) {
    const fc_ctx* k = (const fc_ctx*)vctx;
    const int M = k->Mmin + k->Mlfo;                /* fx.py:42 */
    const float Mf = (float)M;
    const int64_t N = k->N;
    const int b = bc / k->C;
    const float* xs = k->x + (int64_t)bc * N;
    const float* ms = k->mod_has_ch ? k->mod + (int64_t)bc * N : k->mod + (int64_t)b * N;
    float* ys = k->y + (int64_t)bc * N;
    float* buf = (float*)calloc((size_t)M, sizeof(float));          /* fx.py:92 */
    const float a = k->lfo_delay[b], d0 = k->min_delay[b];
    const float g = k->fb[b], dp = k->depth[b], mx = k->mix[b], omm = k->one_minus_mix[b];
    float it_prev = 0.0f;                           /* all-pass interpolator state (interp == 1) */
    for (int64_t n = 0; n < N; ++n) {
        const int w = (int)(n % M);                                  /* fx.py:95 */
        float d = a * ms[n];                                         /* fx.py:98 */
        d = d + d0;
        float t = (float)w - d;                                      /* fx.py:99 */
        t = t + Mf;
        const float r = torch_remainder_f32(t, Mf);
        const float pf = floorf(r);
        const float fr = r - pf;                                     /* fx.py:100 */
        int p = (int)pf;                                             /* fx.py:101 */
        if (p < 0) p = 0;
        if (p >= M) p = M - 1;        /* reference would raise in gather; stay memory-safe */
        const int q = (p + 1) % M;                                   /* fx.py:102 */
        float it;
        if (k->interp == 0) {
            const float t1 = fr * buf[q];                            /* fx.py:113 */
            const float omf = 1.0f - fr;
            const float t2 = omf * buf[p];
            it = t1 + t2;
        } else {
            /* OWN definition (the reference has linear interpolation only): the read position lies
             * delta = 1 - fr samples behind the newer tap buf[q]; first-order all-pass fractional delay
             *   it[n] = eta * (buf[q] - it[n-1]) + buf[p],  eta = (1 - delta) / (1 + delta) = fr / (2 - fr)
             * (J. O. Smith, "Physical Audio Signal Processing", all-pass interpolation), everything else as fx.py:95-118. */
            const float den = 2.0f - fr;
            const float eta = fr / den;
            const float dq = buf[q] - it_prev;
            const float e1 = eta * dq;
            it = e1 + buf[p];
            it_prev = it;
        }
        const float xn = xs[n];
        const float f1 = g * it;                                     /* fx.py:114 */
        buf[w] = xn + f1;
        const float o1 = dp * it;                                    /* fx.py:115 */
        const float o = xn + o1;
        const float m1 = omm * xn;                                   /* fx.py:117 */
        const float m2 = mx * o;
        float out = m1 + m2;
        out = out < -1.0f ? -1.0f : out;                             /* fx.py:118 */
        out = out > 1.0f ? 1.0f : out;
        ys[n] = out;
    }
    free(buf);
}

void modfx_oracle_flanger_chorus(const float* x, const float* mod, int mod_has_ch, float* y,
                                 int B, int C, int64_t N, int Mmin, int Mlfo,
                                 const float* lfo_delay, const float* min_delay,
                                 const float* fb, const float* depth,
                                 const float* mix, const float* one_minus_mix) {
    fc_ctx k = { x, mod, mod_has_ch, y, B, C, N, Mmin, Mlfo,
                 lfo_delay, min_delay, fb, depth, mix, one_minus_mix, 0 };
    parallel_for(B * C, fc_item, &k);
}

/* Same delay line with all-pass instead of linear fractional-delay interpolation (north_star "linear or all-pass";
 * no reference implementation exists: this float32 restatement is this repository's OWN definition). */
void modfx_oracle_flanger_chorus_allpass(const float* x, const float* mod, int mod_has_ch, float* y,
                                         int B, int C, int64_t N, int Mmin, int Mlfo,
                                         const float* lfo_delay, const float* min_delay,
                                         const float* fb, const float* depth,
                                         const float* mix, const float* one_minus_mix) {
    fc_ctx k = { x, mod, mod_has_ch, y, B, C, N, Mmin, Mlfo,
                 lfo_delay, min_delay, fb, depth, mix, one_minus_mix, 1 };
    parallel_for(B * C, fc_item, &k);
}

/* apply_tremolo, fx.py:13-22: ((1 - mix) * x) + (mix * mod * x).
 * Evaluation order in Python: (mix * mod_sig) first, then * x. */
typedef struct {
    const float* x; const float* mod; int mod_has_ch; float* y; int C; int64_t N;
    const float *mix, *one_minus_mix;
} tr_ctx;

static void tr_item(int bc, void* vctx) {
    const tr_ctx* k = (const tr_ctx*)vctx;
    const int b = bc / k->C;
    const int64_t N = k->N;
    const float* xs = k->x + (int64_t)bc * N;
    const float* ms = k->mod_has_ch ? k->mod + (int64_t)bc * N : k->mod + (int64_t)b * N;
    float* ys = k->y + (int64_t)bc * N;
    for (int64_t n = 0; n < N; ++n) {
        const float a = k->one_minus_mix[b] * xs[n];
        float m = k->mix[b] * ms[n];
        m = m * xs[n];
        ys[n] = a + m;
    }
}

void modfx_oracle_tremolo(const float* x, const float* mod, int mod_has_ch, float* y,
                          int B, int C, int64_t N, const float* mix, const float* one_minus_mix) {
    tr_ctx k = { x, mod, mod_has_ch, y, C, N, mix, one_minus_mix };
    parallel_for(B * C, tr_item, &k);
}

/*
 * make_mod_signal, modulations.py:16-57.  freq/phase are the values AFTER the
 * rect_cos / inv_rect_cos halving (modulations.py:26-29), done by the caller in
 * double like Python does.  torch.cumsum on a float32 CPU tensor accumulates in
 * double (acc_type) and rounds each prefix to float32.
 */
void modfx_oracle_lfo(float* out, int64_t n, float sr, float freq, float phase, int shape,
                      double exp_d) {
    const float two_pi = (float)(2.0 * M_PI);
    float inc = two_pi * freq;                 /* modulations.py:31: 2*pi*full(freq) */
    inc = inc / sr;                            /*                     ... / sr         */
    const float pi_f = (float)M_PI;
    const float half_pi_f = (float)(M_PI / 2.0);
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        acc += (double)inc;                    /* cumsum, double accumulator */
        const float arg = (float)acc + phase;
        float v;
        switch (shape) {
        case MODFX_SHAPE_COS: {                /* modulations.py:35 */
            float c = cosf(arg + pi_f);
            c = c + 1.0f;
            v = c / 2.0f;
        } break;
        case MODFX_SHAPE_RECT_COS:             /* modulations.py:37 */
            v = fabsf(cosf(arg + half_pi_f));
            break;
        case MODFX_SHAPE_INV_RECT_COS:         /* modulations.py:39 */
            v = -fabsf(cosf(arg)) + 1.0f;
            break;
        case MODFX_SHAPE_SQR: {                /* modulations.py:41-43 */
            const float c = cosf(arg + pi_f);
            const float s = (c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : 0.0f);
            v = (s + 1.0f) / 2.0f;
        } break;
        default: {                             /* saw family, modulations.py:32 */
            float saw = torch_remainder_f32(arg, two_pi);
            saw = saw / two_pi;
            if (shape == MODFX_SHAPE_SAW) v = saw;
            else if (shape == MODFX_SHAPE_RSAW) v = 1.0f - saw;   /* :48 */
            else {                                                 /* tri :50-51 */
                const float tri = 2.0f * saw;
                v = (tri > 1.0f) ? (2.0f - tri) : tri;
            }
        } break;
        }
        if (exp_d != 1.0) {                    /* modulations.py:55-56, torch.pow(Tensor, Scalar) */
            if (exp_d == 2.0) v = v * v;
            else if (exp_d == 3.0) v = (v * v) * v;
            else if (exp_d == 0.5) v = sqrtf(v);
            else v = powf(v, (float)exp_d);
        }
        out[i] = v;
    }
}

/*
 * util.linear_interpolate_last_dim, util.py:15-29 -> F.interpolate(mode="linear").
 * ATen upsample_linear1d CPU: scale = (I-1)/(O-1) (align_corners) or I/O;
 * src = scale*i (align_corners) or max(fma(scale, i+0.5, -0.5), 0); the two-tap blend
 * is contracted by the compiler to fma(w0, x0, w1*x1) in this torch build
 * (both verified bitwise against F.interpolate by tests/golden/make_golden.py).
 */
typedef struct { const float* in; float* out; int64_t I, O; int align_corners; float scale; } ip_ctx;

static void ip_item(int r, void* vctx) {
    const ip_ctx* k = (const ip_ctx*)vctx;
    const int64_t I = k->I, O = k->O;
    const float* xi = k->in + (int64_t)r * I;
    float* yo = k->out + (int64_t)r * O;
    for (int64_t i = 0; i < O; ++i) {
        float src;
        if (k->align_corners) src = k->scale * (float)i;
        else {
            src = fmaf(k->scale, (float)i + 0.5f, -0.5f);   /* contracted in the torch build */
            if (src < 0.0f) src = 0.0f;
        }
        int64_t i0 = (int64_t)src;
        if (i0 > I - 1) i0 = I - 1;
        const int64_t i1 = i0 + ((i0 < I - 1) ? 1 : 0);
        const float l1 = src - (float)i0;
        const float l0 = 1.0f - l1;
        yo[i] = fmaf(l0, xi[i0], l1 * xi[i1]);
    }
}

void modfx_oracle_interp_linear(const float* in, float* out, int64_t rows, int64_t I, int64_t O,
                                int align_corners) {
    ip_ctx k = { in, out, I, O, align_corners, 0.0f };
    if (align_corners) k.scale = (O > 1) ? (float)(I - 1) / (float)(O - 1) : 0.0f;
    else k.scale = (float)I / (float)O;
    parallel_for((int)rows, ip_item, &k);
}

/*
 * Phaser: restatement (from the published JUCE 7 sources, recalled; NOT available
 * offline => PARITY UNPINNED) of juce::dsp::Phaser<float> as driven by
 * pedalboard 0.7.3's Phaser plugin and called at datasets.py:455-482:
 *   - 6 first-order TPT all-pass stages sharing one cutoff,
 *   - cutoff updated every 4th sample from a sine oscillator running at sr/4
 *     (phase starts at 0, value sin(phase - pi)), scaled by depth*0.5, added to
 *     the log-normalised centre frequency, clamped to [0,1], mapped back to Hz
 *     on a log scale over [20, min(20000, 0.49*sr)],
 *   - cascade input x[n] - lastOut, lastOut = y[n]*feedback,
 *   - linear dry/wet mix,
 *   - host processes in blocks of `block` samples (pedalboard default 8192);
 *     the oscillator phase is advanced per block as in juce::dsp::Oscillator.
 * Output is clipped to [-1,1] (datasets.py:472).
 * All parameters are snapped at reset (no smoothing ramps are active because
 * pedalboard sets them before prepare()/reset()).
 */
typedef struct {
    const float* x; float* y; int64_t N; float sr;
    const float *rate_hz, *depth, *centre_hz, *feedback, *mix; int block;
} ph_ctx;

static void ph_item(int b, void* vctx) {
    enum { STAGES = 6, UPD = 4 };
    const ph_ctx* k = (const ph_ctx*)vctx;
    const int64_t N = k->N;
    const int block = k->block;
    const float sr = k->sr;
    const float* xs = k->x + (int64_t)b * N;
    float* ys = k->y + (int64_t)b * N;
    const float two_pi = (float)(2.0 * M_PI);
    const float pi_f = (float)M_PI;
    const float f_lo = 20.0f;
    const double f_hi_d = (0.49 * (double)sr < 20000.0) ? 0.49 * (double)sr : 20000.0;
    const float f_hi = (float)f_hi_d;
    const float log_min = log10f(f_lo), log_max = log10f(f_hi);
    /* mapFromLog10(centre, 20, f_hi) */
    const float norm_centre = (log10f(k->centre_hz[b]) - log_min) / (log_max - log_min);
    const float osc_vol = k->depth[b] * 0.5f;
    const float sr_down = (float)((double)sr / (double)UPD);
    const float inc = k->rate_hz[b] * (two_pi / sr_down);   /* frequency * baseIncrement */
    const float fbk = k->feedback[b];
    const float wet = k->mix[b], dry = 1.0f - k->mix[b];
    float s[STAGES];
    for (int j = 0; j < STAGES; ++j) s[j] = 0.0f;
    float last = 0.0f;
    float phase = 0.0f;
    float G;
    {   /* FirstOrderTPTFilter default cutoff 1000 Hz until the first update */
        const float g0 = (float)tan(M_PI * 1000.0 / (double)sr);
        G = g0 / (1.0f + g0);
    }
    int counter = 0;                                       /* updateCounter */
    for (int64_t start = 0; start < N; start += block) {
        const int64_t len = (N - start < block) ? (N - start) : block;
        int n_down = 0;                 /* control-rate updates inside this block */
        {
            int c = counter;
            for (int64_t i = 0; i < len; ++i) {
                if (c == 0) n_down++;
                c++;
                if (c == UPD) c = 0;
            }
        }
        float p = phase;
        int c = counter;
        for (int64_t i = 0; i < len; ++i) {
            if (c == 0) {
                /* osc sample: generator(phase.advance(freq) - pi), generator = sin */
                const float last_p = p;
                float next = last_p + inc;
                while (next >= two_pi) next -= two_pi;
                p = next;
                float lfo = sinf(last_p - pi_f) * osc_vol + norm_centre;
                lfo = lfo < 0.0f ? 0.0f : (lfo > 1.0f ? 1.0f : lfo);
                /* mapToLog10 */
                const float fc = powf(10.0f, lfo * (log_max - log_min) + log_min);
                const float g = (float)tan(M_PI * (double)fc / (double)sr);
                G = g / (1.0f + g);
            }
            const float in = xs[start + i];
            float out = in - last;
            for (int j = 0; j < STAGES; ++j) {
                const float v = G * (out - s[j]);
                const float yy = v + s[j];
                s[j] = yy + v;
                out = 2.0f * yy - out;
            }
            last = out * fbk;
            float o = dry * in + wet * out;
            o = o < -1.0f ? -1.0f : (o > 1.0f ? 1.0f : o);
            ys[start + i] = o;
            c++;
            if (c == UPD) c = 0;
        }
        {   /* Oscillator::process tail: phase.advance(freq * n_down) */
            float next = phase + inc * (float)n_down;
            while (next >= two_pi) next -= two_pi;
            phase = next;
        }
        counter = (int)((counter + len) % UPD);
    }
}

void modfx_oracle_phaser(const float* x, float* y, int B, int64_t N, float sr,
                         const float* rate_hz, const float* depth, const float* centre_hz,
                         const float* feedback, const float* mix, int block) {
    ph_ctx k = { x, y, N, sr, rate_hz, depth, centre_hz, feedback, mix, block > 0 ? block : 8192 };
    parallel_for(B, ph_item, &k);
}
