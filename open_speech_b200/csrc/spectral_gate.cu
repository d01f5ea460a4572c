// Non-stationary spectral-gating noise reduction.
//
// Replaces noisereduce.reduce_noise(y=audio, sr=sr) with every default, as called by reduce_noise
// (reference src/audio/preprocessing.py:45-50; third-party noisereduce>=3.0, pyproject.toml:38; algorithm
// restated in SURVEY.md App. A.5 / oracle/stt.py spectral_gate):
//   chunks of 600,000 samples with 30,000 samples of context (zeros beyond the clip) |
//   scipy.signal.stft(nperseg 1024, hop 256, periodic Hann, boundary zeros, spectrum scaling) |
//   A=|S| ; A_s = filtfilt([b],[1,b-1],A) along time ; M = sigmoid(((A-A_s)/A_s - 2)*10) |
//   M = conv2d_same(M, tri(33) x tri(7) / sum) ; istft(S*M) ; keep the centre.
// The reference runs this in float64; here the FFTs run in float32 (relative error ~1e-6, far inside the
// 1e-4 budget) and the one-pole recurrences keep a float64 state.
//
// Kernels: k_nr_stft (two real frames per complex 32x32 four-step FFT, samples staged once per 16 frames)
//          k_nr_iir_fwd / k_nr_iir_bwd_mask (one thread per (chunk, bin), sequential in time, coalesced over bins)
//          k_nr_smooth (separable 33x7 stencil in shared memory)
//          k_nr_istft (32 frames -> 29 hop blocks per CTA, overlap-add in shared memory, one store per sample)
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft.cuh"

namespace osb {

constexpr int NF = 1024, NH = 256, NB = 513;       // n_fft, hop, one-sided bins
constexpr long long kChunk = 600000, kCtx = 30000;
constexpr int kYs = 33, kYPlane = 32 * 33;          // four-step planes [32][33]

struct NrGeom {
    long long n;       // samples per clip
    long long stride;  // samples between clips
    int n_chunks;
    long long Lc;      // padded chunk length
    int F;             // frames per chunk
    int batch;
    int fmt;
};

struct NrTables {
    float* d = nullptr;  // win[1024], cos[1024], sin[1024], ola_scale[256]
};
static std::mutex g_nr_mu;
static std::map<int, NrTables> g_nr;

static int get_nr_tables(const float** out) {
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_nr_mu);
    auto it = g_nr.find(dev);
    if (it == g_nr.end()) {
        std::vector<float> h(3 * NF + NH);
        std::vector<double> w(NF);
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < NF; ++i) {
            w[i] = 0.5 - 0.5 * std::cos(2.0 * pi * i / NF);  // get_window('hann', 1024) (periodic)
            h[i] = (float)w[i];
            // four-step twiddles W1024^(n2*k1) stored as [k1][n2]: a warp (n2 = lane) reads 32 consecutive words
            const int k1 = i >> 5, n2 = i & 31;
            h[NF + i] = (float)std::cos(2.0 * pi * (double)(n2 * k1) / NF);
            h[2 * NF + i] = (float)std::sin(2.0 * pi * (double)(n2 * k1) / NF);
        }
        for (int r = 0; r < NH; ++r) {
            // istft: x *= win.sum() (=512); x /= sum_t win^2 ; our inverse FFT is unnormalised (x 1/1024)
            const double norm = w[r] * w[r] + w[r + 256] * w[r + 256] + w[r + 512] * w[r + 512] + w[r + 768] * w[r + 768];
            h[3 * NF + r] = (float)(512.0 / 1024.0 / norm);
        }
        NrTables t;
        OSB_CUDA(cudaMalloc(&t.d, h.size() * 4));
        OSB_CUDA(cudaMemcpy(t.d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        it = g_nr.emplace(dev, t).first;
    }
    *out = it->second.d;
    return OSB_OK;
}

__device__ __forceinline__ float nr_sample(const void* audio, const NrGeom& g, int clip, int chunk, long long p) {
    // p: chunk coordinate; zeros outside the chunk (stft boundary) and outside the clip (chunk context)
    if (p < 0 || p >= g.Lc) return 0.f;
    const long long i = (long long)chunk * kChunk - kCtx + p;
    if (i < 0 || i >= g.n) return 0.f;
    if (g.fmt == OSB_FMT_PCM16) return ((float)reinterpret_cast<const int16_t*>(audio)[(long long)clip * g.stride + i] * 3.0517578125e-05f);
    return reinterpret_cast<const float*>(audio)[(long long)clip * g.stride + i];
}

// |z| through the SFU reciprocal square root (2 ulp); |S| only feeds the smoothed-threshold mask
__device__ __forceinline__ float fast_mag(float re, float im) {
    const float p = fmaf(re, re, im * im);
    return p > 0.f ? p * rsqrtf(p) : 0.f;
}

// ---------------------------------------------------------------- forward STFT
// grid (ceil(F/16), n_chunks, batch), 256 threads.  S[((clip*n_chunks+chunk)*F + t)*513 + f]
constexpr int kStftFrames = 16, kStftXs = NH * (kStftFrames - 1) + NF;  // 4864

__global__ void __launch_bounds__(256, 2) k_nr_stft(const void* __restrict__ audio, NrGeom g, const float* __restrict__ tabs,
                                                    float2* __restrict__ S, float* __restrict__ A) {
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                 // [4864]
    float* win = xs + kStftXs;      // [1024]
    float* twc = win + NF;          // [1024]
    float* tws = twc + NF;          // [1024]
    float* Y = tws + NF;            // [8][2][32*33]
    const int tid = threadIdx.x, t0 = blockIdx.x * kStftFrames, chunk = blockIdx.y, clip = blockIdx.z;
    for (int i = tid; i < 3 * NF; i += 256) win[i] = tabs[i];
    const long long p0 = (long long)NH * t0 - NF / 2;
    {
        const long long i0 = (long long)chunk * kChunk - kCtx + p0;  // clip index of the first staged sample (multiple of 8)
        const int16_t* src = reinterpret_cast<const int16_t*>(audio) + (long long)clip * g.stride + i0;
        const bool interior = g.fmt == OSB_FMT_PCM16 && p0 >= 0 && p0 + kStftXs <= g.Lc && i0 >= 0 && i0 + kStftXs <= g.n &&
                              (((uintptr_t)src) & 15) == 0;
        if (interior) {
            uint4 v[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if (tid + 256 * r < kStftXs / 8) v[r] = ld_stream_u4(src + 8 * (tid + 256 * r));
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int i8 = tid + 256 * r;
                if (i8 < kStftXs / 8) {
                    const uint32_t w[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        o[2 * k] = (float)(int16_t)(w[k] & 0xFFFF) * 3.0517578125e-05f;
                        o[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * 3.0517578125e-05f;
                    }
                    *reinterpret_cast<float4*>(xs + 8 * i8) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(xs + 8 * i8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        } else {
            for (int i = tid; i < kStftXs; i += 256) xs[i] = nr_sample(audio, g, clip, chunk, p0 + i);
        }
    }
    __syncthreads();
    {   // step 1: 8 frame pairs x 32 residues.  n = 32*n1 + n2
        const int q = tid >> 5, n2 = tid & 31;
        const float* xa = xs + (2 * q) * NH;
        const float* xb = xa + NH;
        cpx v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int idx = 32 * n1 + n2;
            const float w = win[idx];
            v[n1] = cpx{xa[idx] * w, xb[idx] * w};
        }
        fft_pow2<32>(v);
        float* yr = Y + q * 2 * kYPlane;
        float* yi = yr + kYPlane;
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            const int tw = k1 * 32 + n2;
            const cpx y = cmul(v[k1], cpx{twc[tw], -tws[tw]});
            yr[k1 * kYs + n2] = y.x;
            yi[k1 * kYs + n2] = y.y;
        }
    }
    __syncthreads();
    {   // step 2: 32-point FFT over n2 for each (pair, k1): Z[k1 + 32*k2] stored at [k1][k2]
        const int q = tid >> 5, k1 = tid & 31;
        float* yr = Y + q * 2 * kYPlane;
        float* yi = yr + kYPlane;
        cpx v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cpx{yr[k1 * kYs + i], yi[k1 * kYs + i]};
        fft_pow2<32>(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            yr[k1 * kYs + i] = v[i].x;
            yi[k1 * kYs + i] = v[i].y;
        }
    }
    __syncthreads();
    // split the packed transform into the two one-sided spectra, apply the spectrum scaling 1/sum(win) = 1/512
    const float sc = 0.5f / 512.0f;
    const long long row0 = ((long long)clip * g.n_chunks + chunk) * g.F;
    {
        const int q = tid >> 5, lane = tid & 31;  // warp q owns frame pair q
        const int ta = t0 + 2 * q;
        if (ta < g.F) {
            const float* yr = Y + q * 2 * kYPlane;
            const float* yi = yr + kYPlane;
            const bool has_b = ta + 1 < g.F;
            float2* Sa = S + (row0 + ta) * NB;
            float* Aa = A + (row0 + ta) * NB;
            for (int f = lane; f < NB; f += 32) {
                const int m = (NF - f) & (NF - 1);
                const int a0 = (f & 31) * kYs + (f >> 5), a1 = (m & 31) * kYs + (m >> 5);
                const float zr = yr[a0], zi = yi[a0], wr = yr[a1], wi = yi[a1];
                // X_a = (Z[f] + conj Z[N-f])/2 ; X_b = (Z[f] - conj Z[N-f])/(2i)
                const float ar = (zr + wr) * sc, ai = (zi - wi) * sc, br = (zi + wi) * sc, bi = (wr - zr) * sc;
                Sa[f] = make_float2(ar, ai);
                Aa[f] = fast_mag(ar, ai);
                if (has_b) {
                    Sa[NB + f] = make_float2(br, bi);
                    Aa[NB + f] = fast_mag(br, bi);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- time smoothing (filtfilt) + sigmoid mask
// one thread per (chunk, bin), sequential over frames, coalesced across bins; f64 state.
__global__ void __launch_bounds__(128) k_nr_iir_fwd(const float* __restrict__ A, float* __restrict__ Afwd, int F, long long n_rows, double b) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * NB) return;
    const long long cc = idx / NB;
    const int f = (int)(idx - cc * NB);
    const float* s = A + cc * F * NB + f;
    float* o = Afwd + cc * F * NB + f;
    const double a1 = 1.0 - b;
    double y = (double)s[0];  // lfilter_zi start: y[-1] = x[0]
#pragma unroll 32
    for (int t = 0; t < F; ++t) {
        y = fma(a1, y, b * (double)s[(long long)t * NB]);
        o[(long long)t * NB] = (float)y;
    }
}

// M may alias A (in place): thread (chunk, bin) reads A[t][f] before it writes M[t][f].  Because of that alias the
// compiler cannot hoist loads over stores, so the loop is software-pipelined by hand: 16 frames of A and Afwd are
// loaded into registers, then the recurrence + mask for those frames is computed and stored.
template <int U>
__global__ void __launch_bounds__(128) k_nr_iir_bwd_mask(const float* A, const float* __restrict__ Afwd, float* M, int F, long long n_rows,
                                                         double b) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * NB) return;
    const long long cc = idx / NB;
    const int f = (int)(idx - cc * NB);
    const float* s = A + cc * F * NB + f;
    const float* af = Afwd + cc * F * NB + f;
    float* m = M + cc * F * NB + f;
    const double a1 = 1.0 - b;
    double y = (double)af[(long long)(F - 1) * NB];
    for (int t1 = F - 1; t1 >= 0; t1 -= U) {
        float xa[U], xf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t1 - u;
            xa[u] = t >= 0 ? s[(long long)t * NB] : 0.f;
            xf[u] = t >= 0 ? af[(long long)t * NB] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t1 - u;
            if (t < 0) break;
            y = fma(a1, y, b * (double)xf[u]);
            const float as = (float)y;
            const float rel = (xa[u] - as) / as;  // 0/0 -> NaN on all-zero input, like the reference
            m[(long long)t * NB] = 1.0f / (1.0f + __expf(-(rel - 2.0f) * 10.0f));
        }
    }
}

// ---------------------------------------------------------------- 2-D mask smoothing (fftconvolve 'same')
// CTA = 32 frames x all 513 bins.  Time taps first, straight from global memory with a sliding register
// window (coalesced over bins), result into shared memory; then the 2*nf+1 frequency taps with 4 outputs per
// thread (for nf = 16: 9 LDS.128 feed 132 FMA).  grid (ceil(F/32), n_rows)
constexpr int kSmT = 32, kSmPad = 32, kSmW = 584;  // row: 32 zeros | 513 bins | zeros up to 584 floats (16 B aligned)
constexpr int kNtMax = 9, kNfMax = 32;
struct NrSmooth {
    float vf[2 * kNfMax + 1];  // centred: tap k multiplies M[f - NFT + k]; zero-padded when nf < NFT
    float vt[2 * kNtMax + 1];
    int nt;
};

template <int NFT, int NTT>
__global__ void __launch_bounds__(256, 2) k_nr_smooth(const float* __restrict__ M, float* __restrict__ Msm, int F, NrSmooth p) {
    extern __shared__ __align__(16) float tile[];  // [kSmT][kSmW]
    const int t0 = blockIdx.x * kSmT, tid = threadIdx.x;
    const long long base = (long long)blockIdx.y * F * NB;
    for (int i = tid; i < kSmT * kSmW; i += 256) tile[i] = 0.f;
    __syncthreads();
    constexpr int ntap = 2 * NTT + 1;  // vt is centred and zero-padded to this reach
    for (int f = tid; f < NB; f += 256) {
        float w[ntap];
#pragma unroll
        for (int k = 0; k < ntap - 1; ++k) {
            const int t = t0 - NTT + k;
            w[k] = (t >= 0 && t < F) ? M[base + (long long)t * NB + f] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kSmT; ++r) {
            const int t = t0 + r + NTT;  // newest frame entering the window
            w[ntap - 1] = (t < F) ? M[base + (long long)t * NB + f] : 0.f;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < ntap; ++k) acc = fmaf(p.vt[k], w[k], acc);
            tile[r * kSmW + kSmPad + f] = acc;
#pragma unroll
            for (int k = 0; k < ntap - 1; ++k) w[k] = w[k + 1];
        }
    }
    __syncthreads();
    constexpr int NFA = (NFT + 3) / 4 * 4;      // aligned left reach
    constexpr int NV = (4 + 2 * NFA) / 4;       // float4 loads per thread
    const int groups = (NB + 3) / 4;            // 129 groups of 4 bins
    for (int task = tid; task < kSmT * groups; task += 256) {
        const int r = task / groups, gq = task - r * groups;
        if (t0 + r >= F) continue;
        const float* row = tile + r * kSmW + kSmPad + 4 * gq - NFA;
        float x[4 * NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(row + 4 * q);
            x[4 * q] = v4.x; x[4 * q + 1] = v4.y; x[4 * q + 2] = v4.z; x[4 * q + 3] = v4.w;
        }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * NFT + 1; ++k) {
            const float c = p.vf[k];
            a0 = fmaf(c, x[NFA - NFT + k], a0);
            a1 = fmaf(c, x[NFA - NFT + k + 1], a1);
            a2 = fmaf(c, x[NFA - NFT + k + 2], a2);
            a3 = fmaf(c, x[NFA - NFT + k + 3], a3);
        }
        float* o = Msm + base + (long long)(t0 + r) * NB + 4 * gq;
        o[0] = a0;
        if (4 * gq + 1 < NB) { o[1] = a1; o[2] = a2; o[3] = a3; }
    }
}

// ---------------------------------------------------------------- inverse STFT + overlap-add
// grid (tiles, n_chunks, batch): tile = 29 hop blocks [j0, j0+29) <- frames [j0-1, j0+30]
constexpr int kOlaBlocks = 29, kOlaOut = kOlaBlocks * NH;  // 7424 samples

__global__ void __launch_bounds__(256, 2) k_nr_istft(const float2* __restrict__ S, const float* __restrict__ Msm, NrGeom g,
                                                     const float* __restrict__ tabs, int j_first, float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* win = sm;               // [1024]
    float* twc = win + NF;
    float* tws = twc + NF;
    float* osc = tws + NF;         // [256] overlap-add scale
    float* Y = osc + NH;           // [8][2][32*33]
    float* acc = Y + 8 * 2 * kYPlane;  // [7424]
    const int tid = threadIdx.x, chunk = blockIdx.y, clip = blockIdx.z;
    const int j0 = j_first + blockIdx.x * kOlaBlocks;
    for (int i = tid; i < 3 * NF + NH; i += 256) win[i] = tabs[i];
    for (int i = tid; i < kOlaOut; i += 256) acc[i] = 0.f;
    const long long row0 = ((long long)clip * g.n_chunks + chunk) * g.F;
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
        const int tp0 = j0 - 1 + pass * 16;
        {   // stage Z'[k] = Xa[k] + i Xb[k] for the 8 frame pairs of this pass, k stored at [k>>5][k&31] (row stride 33):
            // Xa/Xb are the masked one-sided spectra S*Msm extended by Hermitian symmetry; every S cell is read once, coalesced
            const int q = tid >> 5, lane = tid & 31;  // warp q owns pair q
            const int ta = tp0 + 2 * q, tb = ta + 1;
            const bool va = ta >= 0 && ta < g.F, vb = tb >= 0 && tb < g.F;
            float* yr = Y + q * 2 * kYPlane;
            float* yi = yr + kYPlane;
            const float2* Sa = S + (row0 + ta) * NB;
            const float* Ma = Msm + (row0 + ta) * NB;
            for (int f = lane; f < NB; f += 32) {
                float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
                if (va) { const float2 sv = Sa[f]; const float m = Ma[f]; ar = sv.x * m; ai = sv.y * m; }
                if (vb) { const float2 sv = Sa[NB + f]; const float m = Ma[NB + f]; br = sv.x * m; bi = sv.y * m; }
                if (f == 0 || f == 512) { ai = 0.f; bi = 0.f; }  // irfft ignores the imaginary part of DC / Nyquist
                const int k0 = (f >> 5) * kYs + (f & 31);
                yr[k0] = ar - bi;
                yi[k0] = ai + br;
                if (f != 0 && f != 512) {  // mirror bin N-f: conj(Xa[f]) + i conj(Xb[f])
                    const int km = NF - f, k1 = (km >> 5) * kYs + (km & 31);
                    yr[k1] = ar + bi;
                    yi[k1] = br - ai;
                }
            }
        }
        __syncthreads();
        {   // step 1 of the inverse (conjugate twiddles), in place: thread (q, b) owns column b of pair q
            const int q = tid >> 5, bb = tid & 31;
            float* yr = Y + q * 2 * kYPlane;
            float* yi = yr + kYPlane;
            cpx v[32];
#pragma unroll
            for (int a = 0; a < 32; ++a) v[a] = cpx{yr[a * kYs + bb], yi[a * kYs + bb]};
            fft_pow2<32, true>(v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int tw = c * 32 + bb;
                const cpx y = cmul(v[c], cpx{twc[tw], tws[tw]});
                yr[c * kYs + bb] = y.x;
                yi[c * kYs + bb] = y.y;
            }
        }
        __syncthreads();
        {
            const int q = tid >> 5, c = tid & 31;
            float* yr = Y + q * 2 * kYPlane;
            float* yi = yr + kYPlane;
            cpx v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = cpx{yr[c * kYs + i], yi[c * kYs + i]};
            fft_pow2<32, true>(v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {  // z[n = c + 32 d] at [c][d]
                yr[c * kYs + i] = v[i].x;
                yi[c * kYs + i] = v[i].y;
            }
        }
        __syncthreads();
        // overlap-add the 16 frames of this pass.  Sample n of frame t lands on u = 256 (t - j0) + n - 512, and with
        // n = tid + 256 j every thread only ever touches u == tid (mod 256): no conflicts, no barrier between frames.
#pragma unroll 4
        for (int lt = 0; lt < 16; ++lt) {
            const float* pl = Y + (lt >> 1) * 2 * kYPlane + (lt & 1) * kYPlane;
            const int ub = 256 * (tp0 + lt - j0 - 2) + tid;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = ub + 256 * j, n = tid + 256 * j;
                if (u >= 0 && u < kOlaOut) acc[u] = fmaf(pl[(n & 31) * kYs + (n >> 5)], win[n], acc[u]);
            }
        }
        __syncthreads();
    }
    // store the kept centre [kCtx, kCtx + keep) of the chunk
    const long long keep = (g.n_chunks == 1) ? g.n : ((g.n - (long long)chunk * kChunk) < kChunk ? (g.n - (long long)chunk * kChunk) : kChunk);
    for (int u = tid; u < kOlaOut; u += 256) {
        const long long p = (long long)NH * j0 + u;
        const long long rel = p - kCtx;
        if (rel < 0 || rel >= keep) continue;
        out[(long long)clip * g.stride + (long long)chunk * kChunk + rel] = acc[u] * osc[u & 255];
    }
}

constexpr int kStftSmem = (kStftXs + 3 * NF + 8 * 2 * kYPlane) * (int)sizeof(float);
constexpr int kSmoothSmem = kSmT * kSmW * (int)sizeof(float);
constexpr int kIstftSmem = (3 * NF + NH + 8 * 2 * kYPlane + kOlaOut) * (int)sizeof(float);

static std::vector<double> tri_filter(int n) {
    // concat(linspace(0,1,n+1,endpoint=False), linspace(1,0,n+2))[1:-1]
    std::vector<double> v;
    for (int i = 0; i < n + 1; ++i) v.push_back((double)i / (n + 1));
    for (int i = 0; i < n + 2; ++i) v.push_back(1.0 - (double)i / (n + 1));
    return std::vector<double>(v.begin() + 1, v.end() - 1);
}

int launch_spectral_gate(const void* d_audio, int fmt, long long n, long long batch, long long stride, int sr, float* d_out,
                         cudaStream_t st) {
    const float* tabs;
    int rc = get_nr_tables(&tabs);
    if (rc) return rc;
    static std::once_flag once;
    static cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
    std::call_once(once, [&] {
        e1 = cudaFuncSetAttribute(k_nr_stft, cudaFuncAttributeMaxDynamicSharedMemorySize, kStftSmem);
        e2 = cudaFuncSetAttribute(k_nr_istft, cudaFuncAttributeMaxDynamicSharedMemorySize, kIstftSmem);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_nr_smooth<16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmoothSmem);
        if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_nr_smooth<kNfMax, kNtMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmoothSmem);
    });
    OSB_CUDA(e1);
    OSB_CUDA(e2);
    // noisereduce parameters (defaults)
    const double t_frames = 2.0 * sr / (double)NH;
    const double b = (std::sqrt(1.0 + 4.0 * t_frames * t_frames) - 1.0) / (2.0 * t_frames * t_frames);
    const int nf = (int)(500.0 / (sr / (NF / 2.0)));
    const int nt = (int)(50.0 / (((double)NH / sr) * 1000.0));
    if (nf < 1 || nt < 1 || nf > kNfMax || nt > kNtMax) {
        set_error("unsupported: sample rate %d gives mask smoothing %dx%d outside the supported 1..%d x 1..%d", sr, nf, nt, kNfMax, kNtMax);
        return OSB_ERR_UNSUPPORTED;
    }
    const int nft = (nf == 16 && nt == 3) ? 16 : kNfMax;  // compile-time tap reach of the kernel instance
    NrSmooth sp{};
    sp.nt = nt;
    {
        std::vector<double> vf = tri_filter(nf), vt = tri_filter(nt);
        double sf = 0, stt = 0;
        for (double x : vf) sf += x;
        for (double x : vt) stt += x;
        for (size_t i = 0; i < vf.size(); ++i) sp.vf[(nft - nf) + i] = (float)(vf[i] / sf);
        const int ntt = (nt == 3) ? 3 : kNtMax;
        for (size_t i = 0; i < vt.size(); ++i) sp.vt[(ntt - nt) + i] = (float)(vt[i] / stt);
    }
    NrGeom g;
    g.n = n; g.stride = stride; g.fmt = fmt;
    g.n_chunks = n > kChunk ? (int)((n - 1) / kChunk + 1) : 1;
    g.Lc = n > kChunk ? kChunk + 2 * kCtx : n + 2 * kCtx;
    g.F = (int)(g.Lc / NH) + 1;
    const long long keep_max = n > kChunk ? kChunk : n;
    const int j_first = (int)(kCtx / NH);
    const int j_last = (int)((kCtx + keep_max - 1) / NH);
    const int tiles = (j_last - j_first + 1 + kOlaBlocks - 1) / kOlaBlocks;
    // scratch per clip: S (8 B) + A/M (4 B) + Afwd/Msm (4 B) per (frame, bin); process the batch in groups of <= ~24 GB
    const long long per_clip = (long long)g.n_chunks * g.F * NB;
    long long group = (24ll << 30) / (per_clip * 16);
    if (group < 1) group = 1;
    if (group > batch) group = batch;
    Scratch scr(st);
    float2* S;
    float *A, *Afwd, *M, *Msm;
    OSB_CUDA(scr.alloc(&S, (size_t)(group * per_clip)));
    OSB_CUDA(scr.alloc(&A, (size_t)(group * per_clip)));
    OSB_CUDA(scr.alloc(&Afwd, (size_t)(group * per_clip)));
    M = A;       // the mask overwrites |S| in place (same thread reads A[t][f] then writes M[t][f])
    Msm = Afwd;  // Afwd is dead once the mask exists
    for (long long c0 = 0; c0 < batch; c0 += group) {
        const int gb = (int)((batch - c0) < group ? (batch - c0) : group);
        g.batch = gb;
        const char* in = reinterpret_cast<const char*>(d_audio) + c0 * stride * (fmt == OSB_FMT_PCM16 ? 2 : 4);
        float* outp = d_out + c0 * stride;
        const long long n_rows = (long long)gb * g.n_chunks;
        OSB_LAUNCH(k_nr_stft, dim3((g.F + kStftFrames - 1) / kStftFrames, g.n_chunks, gb), 256, kStftSmem, st, (const void*)in, g, tabs, S, A);
        OSB_CHECK_LAUNCH();
        const unsigned gi = (unsigned)((n_rows * NB + 127) / 128);
        OSB_LAUNCH(k_nr_iir_fwd, gi, 128, 0, st, A, Afwd, g.F, n_rows, b);
        OSB_CHECK_LAUNCH();
        // few rows: deep register pipelining hides latency; many rows: occupancy does, and a shallower pipeline keeps it high
        if (n_rows * NB < 160000) OSB_LAUNCH(k_nr_iir_bwd_mask<16>, gi, 128, 0, st, A, Afwd, M, g.F, n_rows, b);
        else OSB_LAUNCH(k_nr_iir_bwd_mask<8>, gi, 128, 0, st, A, Afwd, M, g.F, n_rows, b);
        OSB_CHECK_LAUNCH();
        if (nft == 16 && nt == 3) OSB_LAUNCH((k_nr_smooth<16, 3>), dim3((g.F + kSmT - 1) / kSmT, (unsigned)n_rows), 256, kSmoothSmem, st, M, Msm, g.F, sp);
        else OSB_LAUNCH((k_nr_smooth<kNfMax, kNtMax>), dim3((g.F + kSmT - 1) / kSmT, (unsigned)n_rows), 256, kSmoothSmem, st, M, Msm, g.F, sp);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_nr_istft, dim3(tiles, g.n_chunks, gb), 256, kIstftSmem, st, S, Msm, g, tabs, j_first, outp);
        OSB_CHECK_LAUNCH();
    }
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_spectral_gate_dev(const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, int sample_rate, float* d_out,
                          void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n && sample_rate > 0, "bad sizes");
    if (n == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(d_audio && d_out, "null buffer");
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    return launch_spectral_gate(d_audio, fmt, n, batch, stride, sample_rate, d_out, (cudaStream_t)stream);
}

int osb_spectral_gate_host(const void* audio, int fmt, float* out, int64_t n, int sample_rate) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    if (n <= 0) return OSB_OK;
    const size_t es = fmt == OSB_FMT_PCM16 ? 2 : 4;
    void *da, *dout;
    if ((rc = ws.dev_buf(0, (size_t)n * es, &da)) || (rc = ws.dev_buf(1, (size_t)n * 4, &dout))) return rc;
    if ((rc = ws.h2d(da, audio, (size_t)n * es))) return rc;
    if ((rc = osb_spectral_gate_dev(da, fmt, n, 1, n, sample_rate, (float*)dout, ws.stream))) return rc;
    return ws.d2h(out, dout, (size_t)n * 4);
}

}  // extern "C"
