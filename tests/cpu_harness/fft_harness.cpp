// CPU emulation of one warp of the log-mel kernel's FFT (two frames): checks the index algebra of
// mod_extraction_b200/csrc/fft_core.h (pass 1, twiddle, transpose, pass 2, paired real-FFT split)
// against a naive double-precision DFT.
// Build: g++ -O2 -I mod_extraction_b200/csrc tests/cpu_harness/fft_harness.cpp -o /tmp/fft_harness
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_core.h"

using namespace modfx;

int main() {
    const int NF = 1024;
    std::vector<float> xw(2 * NF);
    srand(1);
    for (auto& v : xw) v = (float)rand() / RAND_MAX - 0.5f;
    std::vector<float2> tw1(16 * 32);               // exp(-2 pi i n2 k1 / 512) as (cos, sin)
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -2.0 * M_PI * (double)(n2 * k1) / 512.0;
            tw1[k1 * 32 + n2] = make_float2((float)cos(a), (float)sin(a));
        }
    constexpr float C64[32] = MODFX_C64;
    constexpr float S64[32] = MODFX_S64;
    std::vector<float2> E(32 * 33);                 // exchange buffer E[row = f*16 + k1][n2], stride 33
    for (int lane = 0; lane < 32; ++lane)           // ---- pass 1 (lane = n2)
        for (int f = 0; f < 2; ++f) {
            float2 v[16];
            for (int n1 = 0; n1 < 16; ++n1)
                v[n1] = make_float2(xw[f * NF + 64 * n1 + 2 * lane], xw[f * NF + 64 * n1 + 2 * lane + 1]);
            fft_dif<16>(v);
            for (int k1 = 0; k1 < 16; ++k1) {
                const float2 t = tw1[k1 * 32 + lane];
                E[(f * 16 + k1) * 33 + lane] = c_mul_tw(v[BitRev<16>::of(k1)], t.x, t.y);
            }
        }
    std::vector<float2> Z(32 * 32);                 // per lane registers, raw (bit-reversed) order
    for (int lane = 0; lane < 32; ++lane) {         // ---- pass 2 (lane = f*16 + k1)
        float2 v[32];
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = E[lane * 33 + n2];
        fft_dif<32>(v);
        for (int i = 0; i < 32; ++i) Z[lane * 32 + i] = v[i];
    }
    std::vector<double> power(2 * 513, -1.0);
    for (int lane = 0; lane < 32; ++lane) {         // ---- paired real split (lane = f*16 + k1)
        const int f = lane >> 4, k1 = lane & 15;
        const int partner = (f << 4) | ((16 - k1) & 15);
        const int k1m = (16 - k1) & 15;
        const float c1 = (float)cos(2.0 * M_PI * k1 / 1024.0), s1 = (float)sin(2.0 * M_PI * k1 / 1024.0);
        for (int k2 = 0; k2 < 16; ++k2) {
            const float2 z = Z[lane * 32 + BitRev<32>::of(k2)];
            const float2 p = (k1 == 0) ? Z[lane * 32 + BitRev<32>::of((32 - k2) & 31)]
                                       : Z[partner * 32 + BitRev<32>::of(31 - k2)];
            const float c = c1 * C64[k2] - s1 * S64[k2], s = s1 * C64[k2] + c1 * S64[k2];
            const float2 pw = rfft_split_power_pair(z, p, c, s);
            power[f * 513 + k1 + 16 * k2] = pw.x;
            const int km = (k1 == 0) ? 16 * (32 - k2) : (k1m + 16 * (31 - k2));
            if (!(k1 == 0 && k2 == 0)) power[f * 513 + km] = pw.y;
        }
        if (k1 == 0) {
            const float2 z0 = Z[lane * 32];
            power[f * 513 + 512] = (z0.x - z0.y) * (z0.x - z0.y);
            const float2 zq = Z[lane * 32 + BitRev<32>::of(16)];
            power[f * 513 + 256] = zq.x * zq.x + zq.y * zq.y;
        }
    }
    double worst = 0.0, scale = 0.0;
    for (int f = 0; f < 2; ++f)
        for (int k = 0; k <= 512; ++k) {
            double sr = 0, si = 0;
            for (int n = 0; n < NF; ++n) {
                const double a = -2.0 * M_PI * (double)((long)k * n % NF) / NF;
                sr += xw[f * NF + n] * cos(a);
                si += xw[f * NF + n] * sin(a);
            }
            const double ref = sr * sr + si * si;
            worst = fmax(worst, fabs(ref - power[f * 513 + k]));
            scale = fmax(scale, ref);
        }
    printf("max |power err| = %.3e (max power %.3e, rel %.3e)\n", worst, scale, worst / scale);
    return (worst / scale < 1e-5) ? 0 : 1;
}
