// CPU emulation of one warp of the log-mel kernel's FFT (two frames): checks the index algebra of
// mod_extraction_b200/csrc/fft_core.h against a naive double-precision DFT.
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
    // twiddle tables as the kernel builds them
    std::vector<float> tw1c(16 * 32), tw1s(16 * 32), tw2c(512), tw2s(512);
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -2.0 * M_PI * (double)(n2 * k1) / 512.0;
            tw1c[k1 * 32 + n2] = (float)cos(a);
            tw1s[k1 * 32 + n2] = (float)sin(a);
        }
    for (int k = 0; k < 512; ++k) {
        tw2c[k] = (float)cos(2.0 * M_PI * k / 1024.0);
        tw2s[k] = (float)sin(2.0 * M_PI * k / 1024.0);
    }
    // exchange buffer E[row = f*16 + k1][n2], padded stride 33
    std::vector<float> Er(32 * 33), Ei(32 * 33);
    for (int lane = 0; lane < 32; ++lane)          // ---- pass 1 (lane = n2)
        for (int f = 0; f < 2; ++f) {
            float re[16], im[16];
            for (int n1 = 0; n1 < 16; ++n1) {
                re[n1] = xw[f * NF + 64 * n1 + 2 * lane];
                im[n1] = xw[f * NF + 64 * n1 + 2 * lane + 1];
            }
            fft_dif<16>(re, im);
            for (int k1 = 0; k1 < 16; ++k1) {
                const float yr = re[BitRev<16>::of(k1)], yi = im[BitRev<16>::of(k1)];
                const float c = tw1c[k1 * 32 + lane], s = tw1s[k1 * 32 + lane];
                Er[(f * 16 + k1) * 33 + lane] = yr * c - yi * s;
                Ei[(f * 16 + k1) * 33 + lane] = yr * s + yi * c;
            }
        }
    std::vector<float> Zr(32 * 32), Zi(32 * 32);    // per lane registers R[k2] after un-bit-reversal
    for (int lane = 0; lane < 32; ++lane) {         // ---- pass 2 (lane = f*16 + k1)
        float re[32], im[32];
        for (int n2 = 0; n2 < 32; ++n2) {
            re[n2] = Er[lane * 33 + n2];
            im[n2] = Ei[lane * 33 + n2];
        }
        fft_dif<32>(re, im);
        for (int k2 = 0; k2 < 32; ++k2) {
            Zr[lane * 32 + k2] = re[BitRev<32>::of(k2)];
            Zi[lane * 32 + k2] = im[BitRev<32>::of(k2)];
        }
    }
    double worst = 0.0, scale = 0.0;
    for (int lane = 0; lane < 32; ++lane) {         // ---- real split (lane = f*16 + k1)
        const int f = lane >> 4, k1 = lane & 15;
        const int partner = (f << 4) | ((16 - k1) & 15);
        for (int k2 = 0; k2 <= 32; ++k2) {
            float pw;
            int k;
            if (k2 == 32) {                         // Nyquist bin, lane k1 == 0 only
                if (k1 != 0) continue;
                k = 512;
                const float v = Zr[lane * 32] - Zi[lane * 32];
                pw = v * v;
            } else {
                k = k1 + 16 * k2;
                const int src = (k1 == 0) ? ((32 - k2) & 31) : (31 - k2);
                pw = rfft_split_power(Zr[lane * 32 + k2], Zi[lane * 32 + k2], Zr[partner * 32 + src],
                                      Zi[partner * 32 + src], tw2c[k], tw2s[k]);
            }
            double sr = 0, si = 0;
            for (int n = 0; n < NF; ++n) {
                const double a = -2.0 * M_PI * (double)((long)k * n % NF) / NF;
                sr += xw[f * NF + n] * cos(a);
                si += xw[f * NF + n] * sin(a);
            }
            const double ref = sr * sr + si * si;
            worst = fmax(worst, fabs(ref - pw));
            scale = fmax(scale, ref);
        }
    }
    printf("max |power err| = %.3e (max power %.3e, rel %.3e)\n", worst, scale, worst / scale);
    return (worst / scale < 1e-5) ? 0 : 1;
}
