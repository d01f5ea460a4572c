// In-register small FFTs (forward, e^{-2 pi i nk/N}) used by the STFT kernels.
// __host__ __device__ so the index arithmetic can be checked on the CPU (tests/host/fft_check.cu)
// without a GPU.  All loops are fully unrolled with compile-time indices: arrays stay in registers
// and the twiddle constants become immediates.
#pragma once
#include <cuda_runtime.h>

namespace osb {

#define OSB_HD __host__ __device__ __forceinline__

// Complex values are 8-byte aligned register pairs: on sm_100a every operation below is ONE packed FP32 instruction
// (FADD2 / FMUL2 / FFMA2).  ptxas folds the operand shapes these helpers produce -- a scalar broadcast to both halves, a swap of the
// halves, a sign on one half -- into the instruction's operand modifiers (R.F32, .F32x2.LO_HI, .NP / .PN), so a multiplication by
// +-i costs nothing and a radix-2 butterfly with a general twiddle is three instructions instead of six.  The host versions
// (tests/host/fft_check.cu) compute the same expressions with scalar fmaf.
struct __align__(8) cpx {
    float x, y;
};
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define OSB_F2(a) make_float2((a).x, (a).y)
OSB_HD cpx osb_c(float2 v) { return cpx{v.x, v.y}; }
OSB_HD cpx cadd(cpx a, cpx b) { return osb_c(__fadd2_rn(OSB_F2(a), OSB_F2(b))); }
OSB_HD cpx csub(cpx a, cpx b) { return osb_c(__fadd2_rn(OSB_F2(a), make_float2(-b.x, -b.y))); }
OSB_HD cpx cscale(cpx a, float s) { return osb_c(__fmul2_rn(OSB_F2(a), make_float2(s, s))); }
OSB_HD cpx cfma(cpx a, float s, cpx c) { return osb_c(__ffma2_rn(OSB_F2(a), make_float2(s, s), OSB_F2(c))); }            // a s + c
OSB_HD cpx cfma_negi(cpx q, float s, cpx c) { return osb_c(__ffma2_rn(make_float2(q.y, q.x), make_float2(s, -s), OSB_F2(c))); }  // c + s (-i q)
OSB_HD cpx cfma_posi(cpx q, float s, cpx c) { return osb_c(__ffma2_rn(make_float2(q.y, q.x), make_float2(-s, s), OSB_F2(c))); }  // c + s (+i q)
OSB_HD cpx cadd_negi(cpx c, cpx q) { return osb_c(__fadd2_rn(OSB_F2(c), make_float2(q.y, -q.x))); }                     // c - i q
OSB_HD cpx cadd_posi(cpx c, cpx q) { return osb_c(__fadd2_rn(OSB_F2(c), make_float2(-q.y, q.x))); }                     // c + i q
OSB_HD cpx cscale_negi(cpx q, float s) { return osb_c(__fmul2_rn(make_float2(q.y, q.x), make_float2(s, -s))); }          // s (-i q)
OSB_HD cpx cmul2(cpx a, cpx b) { return osb_c(__fmul2_rn(OSB_F2(a), OSB_F2(b))); }                                       // element-wise
#else
OSB_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
OSB_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
OSB_HD cpx cscale(cpx a, float s) { return cpx{a.x * s, a.y * s}; }
OSB_HD cpx cfma(cpx a, float s, cpx c) { return cpx{fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)}; }
OSB_HD cpx cfma_negi(cpx q, float s, cpx c) { return cpx{fmaf(q.y, s, c.x), fmaf(q.x, -s, c.y)}; }
OSB_HD cpx cfma_posi(cpx q, float s, cpx c) { return cpx{fmaf(q.y, -s, c.x), fmaf(q.x, s, c.y)}; }
OSB_HD cpx cadd_negi(cpx c, cpx q) { return cpx{c.x + q.y, c.y - q.x}; }
OSB_HD cpx cadd_posi(cpx c, cpx q) { return cpx{c.x - q.y, c.y + q.x}; }
OSB_HD cpx cscale_negi(cpx q, float s) { return cpx{q.y * s, q.x * -s}; }
OSB_HD cpx cmul2(cpx a, cpx b) { return cpx{a.x * b.x, a.y * b.y}; }
#endif
OSB_HD cpx cneg(cpx a) { return cpx{-a.x, -a.y}; }
OSB_HD cpx cconj(cpx a) { return cpx{a.x, -a.y}; }
OSB_HD cpx cmul(cpx a, cpx b) { return cfma_posi(a, b.y, cscale(a, b.x)); }       // a b        = a b.x + (i a) b.y
OSB_HD cpx cmul_conj(cpx a, cpx w) { return cfma_negi(a, w.y, cscale(a, w.x)); }  // a conj(w)  = a w.x + (-i a) w.y
OSB_HD cpx cmul_negi(cpx a) { return cpx{a.y, -a.x}; }                            // a * (-i)

// cos/sin(2 pi k / 32), k = 0..15
#define OSB_C32 { 1.0000000000e+00f, 9.8078528040e-01f, 9.2387953251e-01f, 8.3146961230e-01f, 7.0710678119e-01f, 5.5557023302e-01f, 3.8268343237e-01f, 1.9509032202e-01f, 6.1232339957e-17f, -1.9509032202e-01f, -3.8268343237e-01f, -5.5557023302e-01f, -7.0710678119e-01f, -8.3146961230e-01f, -9.2387953251e-01f, -9.8078528040e-01f }
#define OSB_S32 { 0.0000000000e+00f, 1.9509032202e-01f, 3.8268343237e-01f, 5.5557023302e-01f, 7.0710678119e-01f, 8.3146961230e-01f, 9.2387953251e-01f, 9.8078528040e-01f, 1.0000000000e+00f, 9.8078528040e-01f, 9.2387953251e-01f, 8.3146961230e-01f, 7.0710678119e-01f, 5.5557023302e-01f, 3.8268343237e-01f, 1.9509032202e-01f }
// cos/sin(2 pi k / 25), k = 0..24
#define OSB_C25 { 1.0000000000e+00f, 9.6858316113e-01f, 8.7630668004e-01f, 7.2896862742e-01f, 5.3582679498e-01f, 3.0901699437e-01f, 6.2790519529e-02f, -1.8738131459e-01f, -4.2577929157e-01f, -6.3742398975e-01f, -8.0901699437e-01f, -9.2977648589e-01f, -9.9211470131e-01f, -9.9211470131e-01f, -9.2977648589e-01f, -8.0901699437e-01f, -6.3742398975e-01f, -4.2577929157e-01f, -1.8738131459e-01f, 6.2790519529e-02f, 3.0901699437e-01f, 5.3582679498e-01f, 7.2896862742e-01f, 8.7630668004e-01f, 9.6858316113e-01f }
#define OSB_S25 { 0.0000000000e+00f, 2.4868988716e-01f, 4.8175367410e-01f, 6.8454710593e-01f, 8.4432792550e-01f, 9.5105651630e-01f, 9.9802672843e-01f, 9.8228725073e-01f, 9.0482705247e-01f, 7.7051324278e-01f, 5.8778525229e-01f, 3.6812455268e-01f, 1.2533323356e-01f, -1.2533323356e-01f, -3.6812455268e-01f, -5.8778525229e-01f, -7.7051324278e-01f, -9.0482705247e-01f, -9.8228725073e-01f, -9.9802672843e-01f, -9.5105651630e-01f, -8.4432792550e-01f, -6.8454710593e-01f, -4.8175367410e-01f, -2.4868988716e-01f }

template <int N>
struct Log2 {
    static constexpr int value = 1 + Log2<N / 2>::value;
};
template <>
struct Log2<1> {
    static constexpr int value = 0;
};

template <int N>
OSB_HD constexpr int bitrev(int i) {
    int r = 0;
    for (int b = 0; b < Log2<N>::value; ++b) r |= ((i >> b) & 1) << (Log2<N>::value - 1 - b);
    return r;
}

// forward FFT, N in {2,4,8,16,32}, natural order in -> natural order out.  INV: conjugate twiddles.
// Every loop has a compile-time trip count so that v[] stays in registers.
template <int N, bool INV = false>
OSB_HD void fft_pow2(cpx (&v)[N]) {
    constexpr float c32[16] = OSB_C32;
    constexpr float s32[16] = OSB_S32;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        constexpr int dummy = 0;
        (void)dummy;
        const int j = bitrev<N>(i);
        if (j > i) {
            const cpx t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    }
#pragma unroll
    for (int s = 1; s <= Log2<N>::value; ++s) {
        const int len = 1 << s, half = len >> 1, step = 32 >> s;
#pragma unroll
        for (int g = 0; g < N / 2; ++g) {
            const int blk = g / half, j = g - blk * half, b = blk * len;
            const int ti = j * step;  // compile-time after unrolling: trivial twiddles cost no multiplies
            const cpx a = v[b + j], q = v[b + j + half];
            if (ti == 0) {
                v[b + j] = cadd(a, q);
                v[b + j + half] = csub(a, q);
            } else if (ti == 8) {  // -i (forward) / +i (inverse)
                v[b + j] = INV ? cadd_posi(a, q) : cadd_negi(a, q);
                v[b + j + half] = INV ? cadd_negi(a, q) : cadd_posi(a, q);
            } else {
                // hi = a + w q with w = wc -+ i ws: two packed FMAs (the +-i q operand is a modifier); lo = a - w q = 2a - hi
                const float wc = c32[ti], ws = s32[ti];
                const cpx t = cfma(q, wc, a);
                const cpx hi = INV ? cfma_posi(q, ws, t) : cfma_negi(q, ws, t);
                v[b + j] = hi;
                v[b + j + half] = cfma(a, 2.0f, cneg(hi));
            }
        }
    }
}

// 5-point DFT (Winograd form), forward
OSB_HD void dft5(cpx& a0, cpx& a1, cpx& a2, cpx& a3, cpx& a4) {
    const float c2 = 0.559016994374947f, s1 = 0.951056516295154f, s2 = 0.587785252292473f;
    const cpx t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
    const cpx t5 = cadd(t1, t2);
    const cpx b0 = cadd(a0, t5);
    const cpx m1 = cfma(t5, -0.25f, a0);
    const cpx m2 = cscale(csub(t1, t2), c2);
    const cpx r1 = cadd(m1, m2), r2 = csub(m1, m2);
    const cpx u1 = cfma(t4, s2, cscale(t3, s1));
    const cpx u2 = cfma(t4, -s1, cscale(t3, s2));
    a0 = b0;
    a1 = cadd_negi(r1, u1);  // r1 - i u1
    a4 = cadd_posi(r1, u1);
    a2 = cadd_negi(r2, u2);
    a3 = cadd_posi(r2, u2);
}

// 25-point DFT, forward, natural in -> natural out (in place)
OSB_HD void dft25(cpx (&v)[25]) {
    constexpr float c25[25] = OSB_C25;
    constexpr float s25[25] = OSB_S25;
    // n = 5a + b ; k = c + 5d.  stage 1: DFT over a for each b -> T[b][c] stored at v[5c + b]
#pragma unroll
    for (int b = 0; b < 5; ++b) dft5(v[b], v[5 + b], v[10 + b], v[15 + b], v[20 + b]);
    // twiddle W25^(b*c)
#pragma unroll
    for (int c = 1; c < 5; ++c)
#pragma unroll
        for (int b = 1; b < 5; ++b) {
            const cpx w = cpx{c25[b * c], -s25[b * c]};
            v[5 * c + b] = cmul(v[5 * c + b], w);
        }
    // stage 2: DFT over b for each c -> Y[c + 5d] at v[5c + d]
#pragma unroll
    for (int c = 0; c < 5; ++c) dft5(v[5 * c], v[5 * c + 1], v[5 * c + 2], v[5 * c + 3], v[5 * c + 4]);
    // transpose (c,d) -> k = c + 5d
    cpx o[25];
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int d = 0; d < 5; ++d) o[c + 5 * d] = v[5 * c + d];
#pragma unroll
    for (int i = 0; i < 25; ++i) v[i] = o[i];
}

// ---------------------------------------------------------------- 400-point complex FFT (25 x 16 four-step)
// Two real 400-sample frames A,B are packed as z = w*(A + iB).  n = 16*n1 + n2, k = k1 + 25*k2:
//   step1(n2): 25-point DFT over n1, times W400^(n2*k1)      -> Y[k1][n2]   (tw holds (cos, sin)(2 pi n2 k1 / 400) at [k1*16 + n2])
//   step2(k1): 16-point FFT over n2                          -> Z[k1 + 25*k2] stored at [k1][k2]
// Y/Z are interleaved complex [25][17] (rows padded to 17 complex: the 64-bit accesses of both steps are conflict-free per half-warp).
constexpr int kF400Stride = 17, kF400Plane = 25 * 17;  // in complex elements

OSB_HD void fft400_step1(const float* xa, const float* xb, const float* win, const cpx* tw, int n2, cpx* Y) {
    cpx v[25];
#pragma unroll
    for (int n1 = 0; n1 < 25; ++n1) {
        const int idx = 16 * n1 + n2;
        v[n1] = cscale(cpx{xa[idx], xb[idx]}, win[idx]);
    }
    dft25(v);
#pragma unroll
    for (int k1 = 0; k1 < 25; ++k1) Y[k1 * kF400Stride + n2] = cmul_conj(v[k1], tw[k1 * 16 + n2]);
}

OSB_HD void fft400_step2(int k1, cpx* Y) {
    cpx v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = Y[k1 * kF400Stride + i];
    fft_pow2<16>(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) Y[k1 * kF400Stride + i] = v[i];
}

OSB_HD int fft400_addr(int k) { return (k % 25) * kF400Stride + (k / 25); }

// power spectra of the two packed real frames at bin k (0..200):  X_A = (Z[k]+conj Z[N-k])/2, X_B = (Z[k]-conj Z[N-k])/(2i)
OSB_HD void fft400_pair_power(const cpx* Z, int k, float* pa, float* pb) {
    const cpx z = Z[fft400_addr(k)], y = Z[fft400_addr(k == 0 ? 0 : 400 - k)];
    const float ar = z.x + y.x, ai = z.y - y.y, br = z.y + y.y, bi = y.x - z.x;
    *pa = 0.25f * (ar * ar + ai * ai);
    *pb = 0.25f * (br * br + bi * bi);
}

// ---------------------------------------------------------------- 1024-point complex FFT by one warp (32 x 32 four-step)
// In place on two float planes [32][33].  Input x[n], n = 32*n1 + n2, sits at [n1*33 + n2]; the result Z[k],
// k = k1 + 32*k2, is left at [k1*33 + k2] (use fft1024_out_addr).  twc/tws hold cos/sin(2 pi n2 k1 / 1024) at
// [k1*32 + n2].  INV conjugates every twiddle (unnormalised inverse).  Lane l owns column l in step 1 and row l
// in step 2, so the only cross-lane hand-over is the __syncwarp between the steps.
constexpr int kF1024Stride = 33, kF1024Plane = 32 * 33;
OSB_HD int fft1024_in_addr(int n) { return (n >> 5) * kF1024Stride + (n & 31); }
OSB_HD int fft1024_out_addr(int k) { return (k & 31) * kF1024Stride + (k >> 5); }

#ifdef __CUDACC__
template <bool INV>
__device__ __forceinline__ void fft1024_warp(float* yr, float* yi, const float* twc, const float* tws, int lane) {
    {
        cpx v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) v[n1] = cpx{yr[n1 * kF1024Stride + lane], yi[n1 * kF1024Stride + lane]};
        fft_pow2<32, INV>(v);
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            const int tw = k1 * 32 + lane;
            const cpx w = cpx{twc[tw], tws[tw]};
            const cpx y = INV ? cmul(v[k1], w) : cmul_conj(v[k1], w);
            yr[k1 * kF1024Stride + lane] = y.x;
            yi[k1 * kF1024Stride + lane] = y.y;
        }
    }
    __syncwarp();
    {
        cpx v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cpx{yr[lane * kF1024Stride + i], yi[lane * kF1024Stride + i]};
        fft_pow2<32, INV>(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            yr[lane * kF1024Stride + i] = v[i].x;
            yi[lane * kF1024Stride + i] = v[i].y;
        }
    }
    __syncwarp();
}
#endif

}  // namespace osb
