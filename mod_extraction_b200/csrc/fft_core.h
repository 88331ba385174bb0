// In-register FFT building blocks of the log-mel kernel (host/device so the index algebra can be
// unit-tested on the CPU: tests/cpu_harness/fft_harness.cpp emulates the 32 lanes of a warp).
//
// A 1024-point real frame is transformed as a 512-point complex FFT of z[m] = x[2m] + i x[2m+1],
// factored 512 = 16 (in-lane, pass 1) x 32 (in-lane after a shared-memory transpose, pass 2):
//   Z[k1 + 16 k2] = sum_{n2<32} W32^{n2 k2} * ( W512^{n2 k1} * sum_{n1<16} z[32 n1 + n2] W16^{n1 k1} )
// followed by the real-FFT split  X[k] = (Z[k] + conj Z[512-k])/2 - i w^k (Z[k] - conj Z[512-k])/2,
// w = exp(-2 pi i / 1024).
#pragma once

#if defined(__CUDACC__)
#define MODFX_HD __host__ __device__ __forceinline__
#else
#define MODFX_HD inline
#endif

namespace modfx {

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
// is a compile-time constant, so the arrays stay in registers.
template <int N>
MODFX_HD void fft_dif(float (&re)[N], float (&im)[N]) {
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
                const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
                re[a] = ar + br;
                im[a] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                const int t = j * tw_stride;        // 0..15
                if (t == 0) {
                    re[b] = dr;
                    im[b] = di;
                } else if (t == 8) {                // * (-i)
                    re[b] = di;
                    im[b] = -dr;
                } else {                            // (dr + i di)(c - i s)
                    const float c = C[t], s = S[t];
                    re[b] = dr * c + di * s;
                    im[b] = di * c - dr * s;
                }
            }
        }
    }
}

// Real-FFT split for one bin.  (zr,zi) = Z[k], (pr,pi) = Z[512-k] (Z[512] == Z[0]),
// (c,s) = (cos, sin)(2 pi k / 1024).  Returns |X[k]|^2.
MODFX_HD float rfft_split_power(float zr, float zi, float pr, float pi, float c, float s) {
    const float ar = zr + pr, ai = zi - pi;     // Z[k] + conj Z[512-k]
    const float br = zr - pr, bi = zi + pi;     // Z[k] - conj Z[512-k]
    const float xr = 0.5f * (ar + c * bi - s * br);
    const float xi = 0.5f * (ai - c * br - s * bi);
    return xr * xr + xi * xi;
}

}  // namespace modfx
