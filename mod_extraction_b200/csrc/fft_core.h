// In-register FFT building blocks of the log-mel kernel (host/device so the index algebra can be
// unit-tested on the CPU: tests/cpu_harness/fft_harness.cpp emulates the 32 lanes of a warp).
//
// A 1024-point real frame is transformed as a 512-point complex FFT of z[m] = x[2m] + i x[2m+1],
// factored 512 = 16 (in-lane, pass 1) x 32 (in-lane after a shared-memory transpose, pass 2):
//   Z[k1 + 16 k2] = sum_{n2<32} W32^{n2 k2} * ( W512^{n2 k1} * sum_{n1<16} z[32 n1 + n2] W16^{n1 k1} )
// followed by the real-FFT split  X[k] = (Z[k] + conj Z[512-k])/2 - i w^k (Z[k] - conj Z[512-k])/2,
// w = exp(-2 pi i / 1024).
//
// Complex numbers are float2 (re, im) and the arithmetic is written with Blackwell's packed FP32x2
// instructions (FADD2 / FMUL2 / FFMA2, sm_100 intrinsics __fadd2_rn / __fmul2_rn / __ffma2_rn): one
// issue slot per complex add and two per complex twiddle multiply (the swap and the sign pattern of the
// second product ride on operand modifiers).  The kernel is bound by instruction issue, not by the FMA
// pipe, so halving the FP instruction count is what pays.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define MODFX_HD __host__ __device__ __forceinline__
#else
#define MODFX_HD inline
struct float2 {
    float x, y;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

namespace modfx {

// ---- packed complex helpers ---------------------------------------------------------------------
MODFX_HD float2 c_add(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
MODFX_HD float2 c_sub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// element-wise a * b
MODFX_HD float2 c_mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
// element-wise a * b + c
MODFX_HD float2 c_fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(a.x * b.x + c.x, a.y * b.y + c.y);
#endif
}
// d * (c - i s) = (dr c + di s, di c - dr s)
MODFX_HD float2 c_mul_conj_tw(float2 d, float c, float s) {
    return c_fma2(d, make_float2(c, c), c_mul2(make_float2(d.y, d.x), make_float2(s, -s)));
}
// d * (c + i s) = (dr c - di s, di c + dr s)
MODFX_HD float2 c_mul_tw(float2 d, float c, float s) {
    return c_fma2(d, make_float2(c, c), c_mul2(make_float2(d.y, d.x), make_float2(-s, s)));
}
// d * (-i) = (di, -dr)
MODFX_HD float2 c_mul_negi(float2 d) { return make_float2(d.y, -d.x); }

// cos/sin(2 pi j / 32), j = 0..15 (W32^j = c - i s)
#define MODFX_C32 {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, \
                   0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, \
                   0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f, \
                   -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f}
#define MODFX_S32 {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f, \
                   0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f, \
                   1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, \
                   0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f}

// cos/sin(2 pi j / 64), j = 0..31
#define MODFX_C64 {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.0f, -0.098017140329560645f, -0.19509032201612819f, -0.29028467725446216f, -0.38268343236508973f, -0.4713967368259977f, -0.55557023301960196f, -0.63439328416364538f, -0.70710678118654746f, -0.77301045336273699f, -0.83146961230254535f, -0.88192126434835494f, -0.92387953251128674f, -0.95694033573220882f, -0.98078528040323043f, -0.99518472667219682f}
#define MODFX_S64 {0.0f, 0.098017140329560604f, 0.19509032201612825f, 0.29028467725446233f, 0.38268343236508978f, 0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f, 0.70710678118654746f, 0.77301045336273699f, 0.83146961230254524f, 0.88192126434835494f, 0.92387953251128674f, 0.95694033573220894f, 0.98078528040323043f, 0.99518472667219682f, 1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220894f, 0.92387953251128674f, 0.88192126434835505f, 0.83146961230254546f, 0.7730104533627371f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960218f, 0.47139673682599786f, 0.38268343236508989f, 0.29028467725446239f, 0.19509032201612861f, 0.098017140329560826f}

template <int N>
struct BitRev;
template <>
struct BitRev<16> {
    static MODFX_HD int of(int i) { return ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3); }
};
template <>
struct BitRev<32> {
    static MODFX_HD int of(int i) {
        return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
    }
};

// Radix-2 decimation-in-frequency FFT, N in {16, 32}, forward (exp(-i...)), in place.
// Output element k ends up at index BitRev<N>::of(k).  Fully unrolled: every index and twiddle
// is a compile-time constant, so the array stays in registers.
template <int N>
MODFX_HD void fft_dif(float2 (&v)[N]) {
    constexpr float C[16] = MODFX_C32;
    constexpr float S[16] = MODFX_S32;
#pragma unroll
    for (int half = N / 2; half >= 1; half >>= 1) {
        const int tw_stride = 16 / half;            // W_{2*half}^j = W32^{j * 16/half}
#pragma unroll
        for (int base = 0; base < N; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int a = base + j, b = base + j + half;
                const float2 va = v[a], vb = v[b];
                v[a] = c_add(va, vb);
                const float2 d = c_sub(va, vb);
                const int t = j * tw_stride;        // 0..15
                if (t == 0) v[b] = d;
                else if (t == 8) v[b] = c_mul_negi(d);
                else v[b] = c_mul_conj_tw(d, C[t], S[t]);
            }
        }
    }
}

// Real-FFT split for the bin pair (k, 512-k): z = Z[k], p = Z[512-k] (Z[512] == Z[0]),
// (c, s) = (cos, sin)(2 pi k / 1024).  Returns (|X[k]|^2, |X[512-k]|^2): the two bins share every
// intermediate term, X[512-k] only flips two signs.
MODFX_HD float2 rfft_split_power_pair(float2 z, float2 p, float c, float s) {
    const float2 hp = c_mul2(p, make_float2(0.5f, -0.5f));              // conj(p) / 2
    const float2 a = c_fma2(z, make_float2(0.5f, 0.5f), hp);            // (Z[k] + conj Z[512-k]) / 2
    const float2 b = c_fma2(z, make_float2(0.5f, 0.5f), make_float2(-hp.x, -hp.y));   // (Z[k] - conj ...) / 2
    const float2 t = c_mul_conj_tw(b, c, s);                            // (c br + s bi, c bi - s br) = (t2, t1)
    const float2 ts = make_float2(t.y, t.x);                            // (t1, t2)
    const float2 x = c_fma2(ts, make_float2(1.0f, -1.0f), a);           // X[k]     = (ar + t1,  ai - t2)
    const float2 y = c_fma2(ts, make_float2(-1.0f, 1.0f), a);           // X[512-k] = (ar - t1, -(ai + t2)) up to sign
    const float2 xx = c_mul2(x, x), yy = c_mul2(y, y);
    return make_float2(xx.x + xx.y, yy.x + yy.y);
}

}  // namespace modfx
