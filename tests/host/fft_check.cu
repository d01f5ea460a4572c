// Host-side check of csrc/fft.cuh (no GPU needed): compares against a naive f64 DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../open_speech_b200/csrc/fft.cuh"
using namespace osb;
template <int N, typename F>
static double check(F f, bool inv = false) {
    cpx v[N];
    double xr[N], xi[N];
    for (int i = 0; i < N; ++i) { xr[i] = rand() / (double)RAND_MAX - 0.5; xi[i] = rand() / (double)RAND_MAX - 0.5; v[i] = cpx{(float)xr[i], (float)xi[i]}; }
    f(v);
    double err = 0;
    for (int k = 0; k < N; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            double a = (inv ? 2 : -2) * M_PI * n * k / N;
            sr += xr[n] * cos(a) - xi[n] * sin(a);
            si += xr[n] * sin(a) + xi[n] * cos(a);
        }
        err = fmax(err, fmax(fabs(sr - v[k].x), fabs(si - v[k].y)));
    }
    return err;
}
int main() {
    double e;
    int bad = 0;
    e = check<25>([](cpx(&v)[25]) { dft25(v); }); printf("dft25 %g\n", e); bad |= e > 1e-5;
    e = check<16>([](cpx(&v)[16]) { fft_pow2<16>(v); }); printf("fft16 %g\n", e); bad |= e > 1e-5;
    e = check<32>([](cpx(&v)[32]) { fft_pow2<32>(v); }); printf("fft32 %g\n", e); bad |= e > 1e-5;
    e = check<32>([](cpx(&v)[32]) { fft_pow2<32, true>(v); }, true); printf("ifft32 %g\n", e); bad |= e > 1e-5;
    e = check<8>([](cpx(&v)[8]) { fft_pow2<8>(v); }); printf("fft8 %g\n", e); bad |= e > 1e-5;
    e = check<4>([](cpx(&v)[4]) { fft_pow2<4>(v); }); printf("fft4 %g\n", e); bad |= e > 1e-5;
    return bad;
}
// (appended) four-step 400-point check
static int check400() {
    static float xa[400], xb[400], win[400];
    static cpx tw[400], Y[kF400Plane];
    for (int i = 0; i < 400; ++i) {
        xa[i] = rand() / (float)RAND_MAX - 0.5f; xb[i] = rand() / (float)RAND_MAX - 0.5f;
        win[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / 400));
        tw[i] = cpx{(float)cos(2 * M_PI * ((i % 16) * (i / 16)) / 400), (float)sin(2 * M_PI * ((i % 16) * (i / 16)) / 400)};
    }
    for (int n2 = 0; n2 < 16; ++n2) fft400_step1(xa, xb, win, tw, n2, Y);
    for (int k1 = 0; k1 < 25; ++k1) fft400_step2(k1, Y);
    double err = 0, mx = 0;
    for (int k = 0; k <= 200; ++k) {
        double ar = 0, ai = 0, br = 0, bi = 0;
        for (int n = 0; n < 400; ++n) {
            double a = -2 * M_PI * n * k / 400;
            ar += (double)xa[n] * win[n] * cos(a); ai += (double)xa[n] * win[n] * sin(a);
            br += (double)xb[n] * win[n] * cos(a); bi += (double)xb[n] * win[n] * sin(a);
        }
        float pa, pb;
        fft400_pair_power(Y, k, &pa, &pb);
        double ra = ar * ar + ai * ai, rb = br * br + bi * bi;
        mx = fmax(mx, fmax(ra, rb));
        err = fmax(err, fmax(fabs(pa - ra), fabs(pb - rb)));
    }
    printf("fft400 pair power: max abs err %g (max power %g, rel %g)\n", err, mx, err / mx);
    return err / mx > 1e-6;
}
struct Run400 { Run400() { if (check400()) { printf("FFT400 FAILED\n"); exit(1); } } } run400;
