// Duration-preserving pitch shift: the "pitch" effect of apply_chain.
//
// Replaces _pitch_shift (reference src/effects/chain.py:44-48) = librosa.effects.pitch_shift(x.astype(f32), sr=sr,
// n_steps=semitones) with librosa's defaults (third-party librosa>=0.10, requirements.lock:14; algorithm restated in
// oracle/tts.py pitch_shift, PARITY UNPINNED: neither librosa nor its resampler soxr is installed and no reference test
// holds a value):
//   rate = 2^(-semitones/12)
//   D  = stft(x, n_fft 2048, hop 512, periodic Hann, centred, zero padding)            k_ps_stft   -> (|D|, angle D)
//   D' = phase_vocoder(D, rate): linear magnitude interpolation between frame pairs,     k_ps_pv
//        phase accumulated per bin in float32 from float64 wrapped increments
//   y  = istft(D', length = round(n / rate)): overlap-add / window sum-square            k_ps_istft
//   out = resample(y, sr/rate -> sr) cropped / zero-padded to n                          k_ps_resample
// The resampler is a Kaiser-windowed sinc interpolator (64 zero crossings, 512 table entries per crossing with linear
// interpolation, roll-off 0.9476, beta 14.77 - the "kaiser_best" design) standing in for soxr_hq.
//
// Real 2048-point transforms run as 1024-point complex FFTs of the even/odd-packed frame, one warp per frame
// (fft1024_warp), with the usual split / merge step.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft.cuh"

namespace osb {

constexpr int PN = 2048, PH = 512, PB = 1025, PM = 1024;
constexpr int kPsZeros = 64, kPsPrec = 512, kPsWin = kPsZeros * kPsPrec + 1;  // 32769 table entries
constexpr double kPsRolloff = 0.9475937167399596, kPsBeta = 14.769656459379492;

struct PsRagged {
    const long long* offsets;
    const long long* lens;
};

// win[2048] | twc[1024] | tws[1024] | pc[1025] | ps[1025] | pad | sinc table: win[32769] | delta[32769]
constexpr int kPsTabWin = 0, kPsTabTwc = PN, kPsTabTws = PN + PM, kPsTabPc = PN + 2 * PM, kPsTabPs = kPsTabPc + PB;
constexpr int kPsTabSmem = (kPsTabPs + PB + 3) / 4 * 4;  // floats staged in shared memory (6152)
constexpr int kPsTabSinc = kPsTabSmem, kPsTabDelta = kPsTabSinc + kPsWin, kPsTabTotal = kPsTabDelta + kPsWin;

static std::mutex g_ps_mu;
static std::map<int, float*> g_ps;

static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

static int get_ps_tables(const float** out) {
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ps_mu);
    auto it = g_ps.find(dev);
    if (it == g_ps.end()) {
        std::vector<float> h(kPsTabTotal, 0.f);
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < PN; ++i) h[kPsTabWin + i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * i / PN));
        for (int i = 0; i < PM; ++i) {
            const int k1 = i >> 5, n2 = i & 31;
            h[kPsTabTwc + i] = (float)std::cos(2.0 * pi * (double)(n2 * k1) / PM);
            h[kPsTabTws + i] = (float)std::sin(2.0 * pi * (double)(n2 * k1) / PM);
        }
        for (int k = 0; k < PB; ++k) {
            h[kPsTabPc + k] = (float)std::cos(2.0 * pi * k / PN);
            h[kPsTabPs + k] = (float)std::sin(2.0 * pi * k / PN);
        }
        // sinc_window(num_zeros 64, precision 9, kaiser(beta), rolloff): right half of the symmetric window
        std::vector<double> w(kPsWin);
        const int n = kPsZeros * kPsPrec;
        const double i0b = bessel_i0(kPsBeta);
        for (int i = 0; i <= n; ++i) {
            const double t = (double)kPsZeros * i / n;  // linspace(0, 64, n + 1)
            const double a = kPsRolloff * t;
            const double sinc = a == 0.0 ? 1.0 : std::sin(pi * a) / (pi * a);
            const double r = (double)i / n;  // kaiser(2n+1)[n + i] = I0(beta sqrt(1 - (i/n)^2)) / I0(beta)
            w[i] = kPsRolloff * sinc * bessel_i0(kPsBeta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / i0b;
        }
        for (int i = 0; i <= n; ++i) {
            h[kPsTabSinc + i] = (float)w[i];
            h[kPsTabDelta + i] = (float)(i < n ? w[i + 1] - w[i] : 0.0);
        }
        float* d = nullptr;
        OSB_CUDA(cudaMalloc(&d, h.size() * 4));
        OSB_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        it = g_ps.emplace(dev, d).first;
    }
    *out = it->second;
    return OSB_OK;
}

struct PsGeom {
    double rate;    // 2^(-semitones/12)
    double ratio;   // resampling ratio sr / (sr / rate)
    int fs;         // frame stride of the analysis buffer   (frames per utterance slot)
    int fs2;        // frame stride of the stretched buffer
    long long ys;   // sample stride of the stretched signal
};
__device__ __forceinline__ int ps_frames(long long n) { return 1 + (int)(n / PH); }
__device__ __forceinline__ int ps_out_frames(int nfr, double rate) { return (int)ceil((double)nfr / rate); }  // len(arange(0, nfr, rate))
__device__ __forceinline__ long long ps_len_stretch(long long n, double rate) { return llrint((double)n / rate); }  // round(): half to even

// ---------------------------------------------------------------- analysis: (|D|, angle D) per frame and bin
// grid (ceil(fs/8), batch), 256 threads: one warp per frame.  MP[(b*fs + t)*1025 + k]
constexpr int kPsSmemFft = (kPsTabSmem + 8 * 2 * kF1024Plane) * (int)sizeof(float);

template <typename T>
__global__ void __launch_bounds__(256) k_ps_stft(const T* __restrict__ x, PsRagged rg, PsGeom g, const float* __restrict__ tabs,
                                                 float2* __restrict__ MP) {
    extern __shared__ __align__(16) float sm[];
    float* tab = sm;
    float* Y = sm + kPsTabSmem;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, b = blockIdx.y;
    for (int i = tid; i < kPsTabSmem; i += 256) tab[i] = tabs[i];
    __syncthreads();
    const long long n = rg.lens[b], off = rg.offsets[b];
    const int t = blockIdx.x * 8 + w;
    if (n <= 0 || t >= ps_frames(n)) return;
    float* yr = Y + w * 2 * kF1024Plane;
    float* yi = yr + kF1024Plane;
    const long long p0 = (long long)t * PH - PN / 2;
    for (int m = lane; m < PM; m += 32) {
        const long long i0 = p0 + 2 * m;
        const float xe = (i0 >= 0 && i0 < n) ? (float)x[off + i0] : 0.f;
        const float xo = (i0 + 1 >= 0 && i0 + 1 < n) ? (float)x[off + i0 + 1] : 0.f;
        const int a = fft1024_in_addr(m);
        yr[a] = xe * tab[kPsTabWin + 2 * m];
        yi[a] = xo * tab[kPsTabWin + 2 * m + 1];
    }
    __syncwarp();
    fft1024_warp<false>(yr, yi, tab + kPsTabTwc, tab + kPsTabTws, lane);
    // X[k] = E[k] + W2048^k O[k],  E = (Z[k] + conj Z[-k])/2,  O = (Z[k] - conj Z[-k])/(2i)
    float2* o = MP + ((long long)b * g.fs + t) * PB;
    for (int k = lane; k < PB; k += 32) {
        const int a0 = fft1024_out_addr(k & (PM - 1)), a1 = fft1024_out_addr((PM - k) & (PM - 1));
        const float zr = yr[a0], zi = yi[a0], mr = yr[a1], mi = -yi[a1];
        const float er = 0.5f * (zr + mr), ei = 0.5f * (zi + mi);
        const float orr = 0.5f * (zi - mi), oi = -0.5f * (zr - mr);
        const float c = tab[kPsTabPc + k], s = tab[kPsTabPs + k];
        const float re = er + c * orr + s * oi, im = ei + c * oi - s * orr;
        o[k] = make_float2(sqrtf(re * re + im * im), atan2f(im, re));
    }
}

// ---------------------------------------------------------------- phase vocoder: one thread per (utterance, bin)
__global__ void __launch_bounds__(128) k_ps_pv(const float2* __restrict__ MP, PsRagged rg, PsGeom g, int batch, float2* __restrict__ D2) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)batch * PB) return;
    const int b = (int)(idx / PB), k = (int)(idx - (long long)b * PB);
    const long long n = rg.lens[b];
    if (n <= 0) return;
    const int nfr = ps_frames(n), nout = ps_out_frames(nfr, g.rate);
    const float2* in = MP + (long long)b * g.fs * PB + k;
    float2* out = D2 + (long long)b * g.fs2 * PB + k;
    const double two_pi = 6.283185307179586476925286766559;
    const double phi = (double)k * (3.14159265358979323846 * PH / (double)(PB - 1));  // linspace(0, pi*hop, 1025)[k]
    float acc = in[0].y;
    for (int t = 0; t < nout; ++t) {
        const double step = (double)t * g.rate;
        const int i0 = (int)step;
        const double alpha = step - floor(step);
        const float2 c0 = i0 < nfr ? in[(long long)i0 * PB] : make_float2(0.f, 0.f);
        const float2 c1 = i0 + 1 < nfr ? in[(long long)(i0 + 1) * PB] : make_float2(0.f, 0.f);
        const double mag = (1.0 - alpha) * (double)c0.x + alpha * (double)c1.x;
        float sn, cs;
        sincosf(acc, &sn, &cs);
        out[(long long)t * PB] = make_float2((float)((double)cs * mag), (float)((double)sn * mag));
        double d = (double)c1.y - (double)c0.y - phi;
        d -= two_pi * rint(d / two_pi);
        acc = (float)((double)acc + (phi + d));
    }
}

// ---------------------------------------------------------------- synthesis: inverse FFT, overlap-add, window sum-square
// grid (ceil(blocks/16), batch): CTA = 16 hop blocks [j0, j0+16) of the overlap-add signal <- frames [j0-3, j0+15]
constexpr int kPsOlaBlocks = 16, kPsOla = kPsOlaBlocks * PH;
constexpr int kPsSmemIstft = (kPsTabSmem + 8 * 2 * kF1024Plane + kPsOla) * (int)sizeof(float);

__global__ void __launch_bounds__(256) k_ps_istft(const float2* __restrict__ D2, PsRagged rg, PsGeom g, const float* __restrict__ tabs,
                                                  float* __restrict__ ys) {
    extern __shared__ __align__(16) float sm[];
    float* tab = sm;
    float* Y = sm + kPsTabSmem;
    float* acc = Y + 8 * 2 * kF1024Plane;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, b = blockIdx.y;
    const long long n = rg.lens[b];
    if (n <= 0) return;
    const long long len2 = ps_len_stretch(n, g.rate);
    const int j0 = blockIdx.x * kPsOlaBlocks;
    if ((long long)j0 * PH - PN / 2 >= len2) return;  // nothing of this tile is kept
    for (int i = tid; i < kPsTabSmem; i += 256) tab[i] = tabs[i];
    for (int i = tid; i < kPsOla; i += 256) acc[i] = 0.f;
    const int nout = ps_out_frames(ps_frames(n), g.rate);
    long long need = (len2 + PN + PH - 1) / PH;  // istft uses at most ceil((length + n_fft) / hop) frames
    const int nfr = (int)(need < nout ? need : nout);
    float* yr = Y + w * 2 * kF1024Plane;
    float* yi = yr + kF1024Plane;
    __syncthreads();
    for (int pass = 0; pass < 3; ++pass) {
        const int f = j0 - 3 + 8 * pass + w;
        const bool live = f >= 0 && f < nfr && f < j0 + kPsOlaBlocks;
        if (live) {
            // Z[k] = (E[k] + i O[k]) / 2 with E = X[k] + conj X[1024-k], O = (X[k] - conj X[1024-k]) W2048^(-k)
            const float2* X = D2 + ((long long)b * g.fs2 + f) * PB;
            for (int k = lane; k < PM; k += 32) {
                float2 xa = X[k], xb = X[PM - k];
                if (k == 0) { xa.y = 0.f; xb.y = 0.f; }  // irfft ignores the imaginary parts of DC and Nyquist
                const float er = xa.x + xb.x, ei = xa.y - xb.y;
                const float dr = xa.x - xb.x, di = xa.y + xb.y;
                const float c = tab[kPsTabPc + k], s = tab[kPsTabPs + k];
                const float orr = dr * c - di * s, oi = dr * s + di * c;
                const int a = fft1024_in_addr(k);
                yr[a] = 0.5f * (er - oi);
                yi[a] = 0.5f * (ei + orr);
            }
            __syncwarp();
            fft1024_warp<true>(yr, yi, tab + kPsTabTwc, tab + kPsTabTws, lane);
        }
        __syncthreads();
        // sample n of frame f lands on u = 512 (f - j0) + n; with n = tid + 256 q every thread only touches u == tid (mod 256)
        for (int lw = 0; lw < 8; ++lw) {
            const int ff = j0 - 3 + 8 * pass + lw;
            if (ff < 0 || ff >= nfr || ff >= j0 + kPsOlaBlocks) continue;
            const float* pr = Y + lw * 2 * kF1024Plane;
            const float* pi = pr + kF1024Plane;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int nn = tid + 256 * q, u = PH * (ff - j0) + nn;
                if (u < 0 || u >= kPsOla) continue;
                const int a = fft1024_out_addr(nn >> 1);
                const float v = (nn & 1) ? pi[a] : pr[a];
                acc[u] = fmaf(v * (1.0f / PM), tab[kPsTabWin + nn], acc[u]);
            }
        }
        __syncthreads();
    }
    for (int u = tid; u < kPsOla; u += 256) {
        const long long p = (long long)j0 * PH + u, o = p - PN / 2;
        if (o < 0 || o >= len2) continue;
        float wss = 0.f;
        const int f_hi = (int)(p / PH);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int f = f_hi - d;
            if (f >= 0 && f < nfr) { const float wv = tab[kPsTabWin + (int)(p - (long long)f * PH)]; wss = fmaf(wv, wv, wss); }
        }
        ys[(long long)b * g.ys + o] = wss > 1.1754944e-38f ? acc[u] / wss : acc[u];
    }
}

// ---------------------------------------------------------------- band-limited interpolation back to n samples
template <typename OUT>
__global__ void __launch_bounds__(256) k_ps_resample(const float* __restrict__ ys, PsRagged rg, PsGeom g, const float* __restrict__ tabs,
                                                     OUT* __restrict__ out) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b], off = rg.offsets[b];
    if (n <= 0) return;
    const long long len2 = ps_len_stretch(n, g.rate);
    const long long n_res = (long long)((double)len2 * g.ratio);              // int(len * ratio)
    const long long n_fix = (long long)ceil((double)len2 * g.ratio);          // fix_length(ceil(len * ratio))
    const long long valid = n_res < n_fix ? n_res : n_fix;
    const float* win = tabs + kPsTabSinc;
    const float* dlt = tabs + kPsTabDelta;
    const float* x = ys + (long long)b * g.ys;
    const double scale = g.ratio < 1.0 ? g.ratio : 1.0;
    const float gain = g.ratio < 1.0 ? (float)g.ratio : 1.0f;  // the table is scaled by the ratio when decimating
    const int istep = (int)(scale * kPsPrec);
    const double inc = 1.0 / g.ratio;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        float y = 0.f;
        if (t < valid) {
            const double tr = (double)t * inc;
            const long long nn = (long long)tr;
            {
                const double fi = scale * (tr - (double)nn) * kPsPrec;
                const int o = (int)fi;
                const float eta = (float)(fi - (double)o);
                long long imax = (kPsWin - o) / istep;
                if (imax > nn + 1) imax = nn + 1;
                float a0 = 0.f, a1 = 0.f;
                long long i = 0;
                for (; i + 1 < imax; i += 2) {
                    const int j0 = o + (int)i * istep, j1 = j0 + istep;
                    a0 = fmaf(fmaf(eta, dlt[j0], win[j0]), x[nn - i], a0);
                    a1 = fmaf(fmaf(eta, dlt[j1], win[j1]), x[nn - i - 1], a1);
                }
                if (i < imax) { const int j0 = o + (int)i * istep; a0 = fmaf(fmaf(eta, dlt[j0], win[j0]), x[nn - i], a0); }
                y = a0 + a1;
            }
            {
                const double fi = (scale - scale * (tr - (double)nn)) * kPsPrec;
                const int o = (int)fi;
                const float eta = (float)(fi - (double)o);
                long long kmax = (kPsWin - o) / istep;
                if (kmax > len2 - nn - 1) kmax = len2 - nn - 1;
                float a0 = 0.f, a1 = 0.f;
                long long k = 0;
                for (; k + 1 < kmax; k += 2) {
                    const int j0 = o + (int)k * istep, j1 = j0 + istep;
                    a0 = fmaf(fmaf(eta, dlt[j0], win[j0]), x[nn + k + 1], a0);
                    a1 = fmaf(fmaf(eta, dlt[j1], win[j1]), x[nn + k + 2], a1);
                }
                if (k < kmax) { const int j0 = o + (int)k * istep; a0 = fmaf(fmaf(eta, dlt[j0], win[j0]), x[nn + k + 1], a0); }
                y += a0 + a1;
            }
            y *= gain;
        }
        out[off + t] = (OUT)y;
    }
}

// d_in: f32 or f64 ragged batch; d_out: f32 ragged batch with the same offsets / lengths
int launch_pitch_shift(const void* d_in, bool in_f64, const long long* d_offsets, const long long* d_lens, long long batch, long long max_len,
                       int sample_rate, double semitones, float* d_out, cudaStream_t st) {
    const float* tabs;
    int rc = get_ps_tables(&tabs);
    if (rc) return rc;
    static PerDeviceOnce once;
    OSB_CUDA(once.run([&] {
        cudaError_t e1 = cudaFuncSetAttribute(k_ps_stft<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmemFft);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_ps_stft<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmemFft);
        if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_ps_istft, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmemIstft);
        return e1;
    }));
    PsGeom g;
    g.rate = std::pow(2.0, -semitones / 12.0);
    g.ratio = (double)sample_rate / ((double)sample_rate / g.rate);
    if (!(g.rate > 1.0 / 16 && g.rate < 16.0)) {
        set_error("unsupported: pitch shift of %g semitones (|shift| must stay below 48)", semitones);
        return OSB_ERR_UNSUPPORTED;
    }
    g.fs = 1 + (int)(max_len / PH);
    g.fs2 = (int)std::ceil((double)g.fs / g.rate) + 1;
    g.ys = ((long long)std::llrint((double)max_len / g.rate) + 8 + 3) / 4 * 4;
    // scratch per utterance slot: analysis + stretched spectra (8 B per cell) + stretched signal; groups of <= ~12 GB
    const long long per = ((long long)g.fs + g.fs2) * PB * 8 + g.ys * 4;
    long long group = (12ll << 30) / per;
    if (group < 1) group = 1;
    if (group > batch) group = batch;
    Scratch scr(st);
    float2 *MP, *D2;
    float* ys;
    OSB_CUDA(scr.alloc(&MP, (size_t)(group * g.fs * PB)));
    OSB_CUDA(scr.alloc(&D2, (size_t)(group * g.fs2 * PB)));
    OSB_CUDA(scr.alloc(&ys, (size_t)(group * g.ys)));
    const long long len2_max = std::llrint((double)max_len / g.rate);
    const int ola_tiles = (int)((len2_max + PN / 2 + kPsOla - 1) / kPsOla);
    for (long long b0 = 0; b0 < batch; b0 += group) {
        const int gb = (int)((batch - b0) < group ? (batch - b0) : group);
        PsRagged rg{d_offsets + b0, d_lens + b0};
        const dim3 gs((g.fs + 7) / 8, gb);
        if (in_f64) OSB_LAUNCH(k_ps_stft<double>, gs, 256, kPsSmemFft, st, (const double*)d_in, rg, g, tabs, MP);
        else OSB_LAUNCH(k_ps_stft<float>, gs, 256, kPsSmemFft, st, (const float*)d_in, rg, g, tabs, MP);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_ps_pv, (unsigned)(((long long)gb * PB + 127) / 128), 128, 0, st, MP, rg, g, gb, D2);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_ps_istft, dim3(ola_tiles, gb), 256, kPsSmemIstft, st, D2, rg, g, tabs, ys);
        OSB_CHECK_LAUNCH();
        long long per_b = (max_len + 255) / 256;
        const long long want = ((long long)OSB_NUM_SMS * 16 + gb - 1) / gb;
        if (per_b > want) per_b = want;
        if (per_b < 1) per_b = 1;
        OSB_LAUNCH(k_ps_resample<float>, dim3((unsigned)per_b, gb), 256, 0, st, ys, rg, g, tabs, d_out);
        OSB_CHECK_LAUNCH();
    }
    return OSB_OK;
}

}  // namespace osb
