// Whisper-style log-mel front-end (n_fft 400, hop 160, 80/128 Slaney mel bins).
//
// Replaces the arithmetic of faster_whisper.FeatureExtractor.__call__(waveform, padding=160)
// (third-party, pinned faster-whisper==1.2.1 in requirements.lock:7; call site
// src/backends/faster_whisper.py:245; algorithm restated in SURVEY.md App. A.4 / oracle/stt.py):
//   pad 160 zeros | reflect-pad 200 | frames 400/160 * periodic Hann | rfft | |.|^2, drop last frame |
//   mel[n_mels,201] @ P | log10(max(.,1e-10)) | max(., global_max-8) | (.+4)/4
// and, fused in front of it for the batch STT path, normalize_gain + int16 requantisation
// (src/audio/preprocessing.py:35-42, :23-25): the reference hands faster-whisper a re-quantised WAV.
//
// Design (why not a DFT-GEMM): the direct 400x402 DFT as a GEMM costs 37 MFLOP per audio-second
// (x3 for split-precision operands) and sits above the tensor ridge; a 25x16 four-step FFT that
// packs two real frames into one complex transform costs ~1.2 MFLOP per audio-second on the FP32
// pipe and keeps the stage closer to its HBM bound.  One CTA = 32 consecutive
// frames of one clip: samples are staged once in shared memory (5,360 samples serve 32 overlapping
// frames), the spectra never leave the SM, and only [n_mels x 32] floats are written, 128 B per row.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "fft.cuh"

namespace osb {

constexpr int kNfft = 400, kHop = 160, kBins = 201, kPad = 160;
constexpr int MF = 32;                             // frames per CTA
constexpr int kXs = kHop * (MF - 1) + kNfft;       // 5360 staged samples
constexpr int kPStride = MF + 2;                   // power tile [204][34]: 201 bins + 3 rows the zero-padded mel taps may touch; row = 16 x (frame q, frame q + 16)
constexpr int kPRows = kBins + 3;
constexpr int kMelWMax = 768;                      // padded mel taps staged in shared memory (128 mels: 512, 80 mels: 480)

struct MelTables {
    int n_mels = 0;
    std::vector<float> dense;   // [n_mels][201] f32 (what FeatureExtractor.mel_filters holds)
    float* d_consts = nullptr;  // win[400], tw[400] as (cos, sin) pairs
    int* d_meta = nullptr;      // [n_mels] first non-zero bin | padded tap count << 8 | offset into d_w << 16
    float* d_w = nullptr;       // taps of every filter x 0.25 (the power tile holds 4 |X|^2: exact), each zero-padded to a multiple of 4
    int n_w = 0;
};

// faster_whisper.FeatureExtractor.get_mel_filters(16000, 400, n_mels): Slaney scale + area norm
static void build_mel(int n_mels, std::vector<float>& dense) {
    const int sr = 16000;
    const double val = 1.0 / (kNfft * (1.0 / sr));
    std::vector<double> fft(kBins), mels(n_mels + 2), freqs(n_mels + 2);
    for (int k = 0; k < kBins; ++k) fft[k] = k * val;
    const double max_mel = 45.245640471924965, step = max_mel / (n_mels + 1);
    for (int i = 0; i < n_mels + 2; ++i) mels[i] = i * step;
    mels[n_mels + 1] = max_mel;
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    for (int i = 0; i < n_mels + 2; ++i)
        freqs[i] = mels[i] >= min_log_mel ? min_log_hz * std::exp(logstep * (mels[i] - min_log_mel)) : f_sp * mels[i];
    dense.assign((size_t)n_mels * kBins, 0.f);
    for (int m = 0; m < n_mels; ++m) {
        const double fd0 = freqs[m + 1] - freqs[m], fd1 = freqs[m + 2] - freqs[m + 1];
        const double enorm = 2.0 / (freqs[m + 2] - freqs[m]);
        for (int k = 0; k < kBins; ++k) {
            const double lower = -(freqs[m] - fft[k]) / fd0, upper = (freqs[m + 2] - fft[k]) / fd1;
            double w = lower < upper ? lower : upper;
            if (w < 0) w = 0;
            dense[(size_t)m * kBins + k] = (float)(w * enorm);
        }
    }
}

static std::mutex g_mel_mu;
static std::map<long long, MelTables> g_mel;

static int get_mel(int n_mels, bool need_device, const MelTables** out) {
    int dev = 0;
    if (need_device) OSB_CUDA(cudaGetDevice(&dev));
    const long long key = ((long long)(need_device ? dev + 1 : 0) << 32) | (unsigned)n_mels;
    std::lock_guard<std::mutex> lk(g_mel_mu);
    auto it = g_mel.find(key);
    if (it == g_mel.end()) {
        MelTables t;
        t.n_mels = n_mels;
        build_mel(n_mels, t.dense);
        if (need_device) {
            std::vector<float> consts(1200);
            const double pi = 3.14159265358979323846;
            for (int i = 0; i < 400; ++i) {
                consts[i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * i / 400.0));  // np.hanning(401)[:-1] -> f32
                // four-step twiddles W400^(n2*k1) stored as [k1][n2] (25 x 16): conflict-free per-lane reads
                const int k1 = i / 16, n2 = i % 16;
                consts[400 + 2 * i] = (float)std::cos(2.0 * pi * (double)(n2 * k1) / 400.0);
                consts[401 + 2 * i] = (float)std::sin(2.0 * pi * (double)(n2 * k1) / 400.0);
            }
            std::vector<int> meta(n_mels);
            std::vector<float> w;
            for (int m = 0; m < n_mels; ++m) {
                int a = -1, b = -1;
                for (int k = 0; k < kBins; ++k)
                    if (t.dense[(size_t)m * kBins + k] != 0.f) { if (a < 0) a = k; b = k; }
                const int start = a < 0 ? 0 : a, len = a < 0 ? 0 : b - a + 1, len4 = len == 0 ? 4 : (len + 3) / 4 * 4;  // >= one round of four
                meta[m] = start | (len4 << 8) | ((int)w.size() << 16);
                for (int k = 0; k < len4; ++k) w.push_back(k < len ? 0.25f * t.dense[(size_t)m * kBins + start + k] : 0.f);
            }
            if ((int)w.size() > kMelWMax) {
                set_error("internal: mel filterbank has %d padded taps (> %d)", (int)w.size(), kMelWMax);
                return OSB_ERR_UNSUPPORTED;
            }
            t.n_w = (int)w.size();
            OSB_CUDA(cudaMalloc(&t.d_consts, consts.size() * 4));
            OSB_CUDA(cudaMemcpy(t.d_consts, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice));
            OSB_CUDA(cudaMalloc(&t.d_meta, n_mels * 4)); OSB_CUDA(cudaMemcpy(t.d_meta, meta.data(), n_mels * 4, cudaMemcpyHostToDevice));
            OSB_CUDA(cudaMalloc(&t.d_w, w.size() * 4)); OSB_CUDA(cudaMemcpy(t.d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        }
        it = g_mel.emplace(key, std::move(t)).first;
    }
    *out = &it->second;
    return OSB_OK;
}

struct MelArgs {
    const void* audio;
    long long n, stride;
    int fmt, n_frames, n_mels;
    float* out;                            // [batch][n_mels][n_frames] raw log10 values
    unsigned int* gmax;                    // [batch] bits of (max log10 + 10) >= 0
    const float* gain;                     // fused normalise: per-clip gain from k_mel_gain (negative: silent clip, passed through), else null
    int requant;                           // clip / x32767 / truncate / (/32768) while staging (the chain's WAV round trip)
    float target_dbfs;
    const float* consts;
    const int* mel_meta;
    const float* mel_w;
    int n_w;
};

// normalize_gain (clip of x * gain; a silent clip passes through unclipped) followed by the WAV round trip (clip, x 32767, truncate, / 32768):
// the second clip makes the first one redundant, and a silent clip is the gain 1.0f (x * 1.0f is x), so one multiply and one clip serve both
__device__ __forceinline__ float mel_requant(const MelArgs& a, float x, float gain, bool /*silent: folded into gain by the caller*/) {
    return a.requant ? (float)__float2int_rz(__fmul_rn(fminf(fmaxf(__fmul_rn(x, gain), -1.0f), 1.0f), 32767.0f)) * 3.0517578125e-05f : x;
}
// (log_spec + 4.0) / 4.0, applied by the transform kernels themselves (x / 4.0 written as x * 0.25: the same value for every float).  The
// clamp log_spec = maximum(log_spec, max - 8.0) comes BEFORE it in the reference; the two commute exactly because rounding is monotonic:
// (max(x, t) + 4) / 4 == max((x + 4) / 4, (t + 4) / 4) for all floats, so the clamp pass only rewrites the cells it changes.
__device__ __forceinline__ float logmel_affine(float x) { return __fmul_rn(__fadd_rn(x, 4.0f), 0.25f); }

__device__ __forceinline__ float mel_sample(const MelArgs& a, int s16, float gain, bool silent) {
    return mel_requant(a, (float)s16 * 3.0517578125e-05f, gain, silent);  // /32768, exact
}

// Persistent kernel: grid = 2 CTAs per SM, each CTA walks tiles (32 frames of one clip) round-robin.  The raw
// int16 samples of the NEXT tile are fetched by the TMA unit (cp.async.bulk -> mbarrier) while the current tile is
// transformed, so the HBM latency of the staging step is off the critical path.
//   smem: [xs|P (aliased)] [win tw] [Y] [raw int16] ; P reuses the float sample buffer once step 1 is done.
constexpr int kRawBytes = kXs * 4;                                   // raw staging: 10,720 B of int16 or 21,440 B of float32
constexpr int kXsP = (((kPRows * kPStride > kXs) ? kPRows * kPStride : kXs) + 3) / 4 * 4;  // 6,732 floats (keeps raw[] 16 B aligned)
static_assert(kPStride % 2 == 0 && ((kXsP + 1200 + 16 * 2 * kF400Plane) * 4) % 16 == 0, "raw[] must be 16-byte aligned");

__device__ __forceinline__ bool mel_tile_interior(const MelArgs& a, int b, int t0, const int16_t** src) {
    const long long p0 = (long long)kHop * t0 - kNfft / 2;  // multiple of 8 samples
    const int16_t* s = reinterpret_cast<const int16_t*>(a.audio) + (long long)b * a.stride + p0;
    if (a.fmt != OSB_FMT_PCM16) s = reinterpret_cast<const int16_t*>(reinterpret_cast<const float*>(a.audio) + (long long)b * a.stride + p0);
    *src = s;
    return p0 >= 0 && p0 + kXs <= a.n && (((uintptr_t)s) & 15) == 0;
}

template <bool F32>  // input format fixed at compile time: raw staging size and the conversion loop
__global__ void __launch_bounds__(256, 2) k_logmel(MelArgs a, int tiles_per_clip, int total_tiles) {
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                       // [kXs] float samples (step 1)  |  P [201][33] (power, later phases)
    float* P = sm;
    float* win = sm + kXsP;               // [400]
    cpx* tw = reinterpret_cast<cpx*>(win + 400);  // [400] (cos, sin)
    cpx* Y = tw + 400;                    // [16][425] complex
    int16_t* raw = reinterpret_cast<int16_t*>(Y + 16 * kF400Plane);  // [kXs] int16 or float32, TMA destination
    constexpr uint32_t raw_bytes = F32 ? kXs * 4 : kXs * 2;
    __shared__ int mel_meta[128];
    __shared__ __align__(16) float mel_wsm[kMelWMax];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;

    for (int i = tid; i < 1200; i += 256) win[i] = a.consts[i];
    // rows 201..203 of P are only ever multiplied by the zero padding of the mel taps, but they must hold finite numbers:
    // they lie beyond the staged samples, so they keep whatever the previous kernel on this SM left there
    for (int i = tid; i < (kPRows - kBins) * kPStride; i += 256) P[kBins * kPStride + i] = 0.f;
    for (int m = tid; m < a.n_mels; m += 256) mel_meta[m] = a.mel_meta[m];
    for (int i = tid; i < a.n_w; i += 256) mel_wsm[i] = a.mel_w[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        const int16_t* src;
        const int tile = blockIdx.x;
        if (tile < total_tiles && mel_tile_interior(a, tile / tiles_per_clip, (tile % tiles_per_clip) * MF, &src)) {
            mbar_expect_tx(&bar, raw_bytes);
            bulk_g2s(raw, src, raw_bytes, &bar);
        }
    }
    __syncthreads();
    uint32_t parity = 0;

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_clip, t0 = (tile - b * tiles_per_clip) * MF;
        {
            // stage the tile's samples: reflect pad 200 around [x, 160 zeros]; fused normalise + requantise
            bool silent = true;
            float gain = 1.0f;
            if (a.gain) {
                gain = a.gain[b];
                silent = gain < 0.f;
                if (silent) gain = 1.0f;
            }
            const int16_t* src;
            if (mel_tile_interior(a, b, t0, &src)) {
                mbar_wait(&bar, parity);
                parity ^= 1u;
                if (F32) {
                    const float4* rf = reinterpret_cast<const float4*>(raw);
                    for (int i4 = tid; i4 < kXs / 4; i4 += 256) {
                        const float4 v = rf[i4];
                        *reinterpret_cast<float4*>(xs + 4 * i4) = make_float4(mel_requant(a, v.x, gain, silent), mel_requant(a, v.y, gain, silent),
                                                                              mel_requant(a, v.z, gain, silent), mel_requant(a, v.w, gain, silent));
                    }
                } else
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int i8 = tid + 256 * r;
                    if (i8 < kXs / 8) {
                        const uint4 v = *reinterpret_cast<const uint4*>(raw + 8 * i8);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        float o[8];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            o[2 * k] = mel_sample(a, (int)(int16_t)(w[k] & 0xFFFF), gain, silent);
                            o[2 * k + 1] = mel_sample(a, (int)(int16_t)(w[k] >> 16), gain, silent);
                        }
                        *reinterpret_cast<float4*>(xs + 8 * i8) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(xs + 8 * i8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            } else {
                const long long L = a.n + kPad;
                const long long p0 = (long long)kHop * t0 - kNfft / 2;
                for (int i = tid; i < kXs; i += 256) {
                    long long p = p0 + i;
                    while (p < 0 || p >= L) p = p < 0 ? -p : 2 * (L - 1) - p;
                    float v = 0.f;
                    if (p < a.n) {
                        if (a.fmt == OSB_FMT_PCM16) v = mel_sample(a, (int)reinterpret_cast<const int16_t*>(a.audio)[(long long)b * a.stride + p], gain, silent);
                        else v = mel_requant(a, reinterpret_cast<const float*>(a.audio)[(long long)b * a.stride + p], gain, silent);
                    }
                    xs[i] = v;
                }
            }
        }
        __syncthreads();  // xs complete; raw[] has been consumed by every thread
        if (tid == 0) {   // prefetch the next tile of this CTA while this one is transformed
            const int nt = tile + gridDim.x;
            const int16_t* src;
            if (nt < total_tiles && mel_tile_interior(a, nt / tiles_per_clip, (nt % tiles_per_clip) * MF, &src)) {
                fence_proxy_async();  // generic-proxy reads of raw[] above are ordered before the async-proxy write
                mbar_expect_tx(&bar, raw_bytes);
                bulk_g2s(raw, src, raw_bytes, &bar);
            }
        }
        {   // four-step FFT, step 1: 16 frame pairs x 16 residues = 256 tasks.  Pair q = frames (q, q + 16): the mel phase below
            // then finds the two frames of a pair side by side and still stores 64-byte runs
            const int q = tid >> 4, n2 = tid & 15;
            fft400_step1(xs + q * kHop, xs + (q + 16) * kHop, win, tw, n2, Y + q * kF400Plane);
        }
        __syncthreads();
        {   // step 2 fused with the power spectrum, one warp per frame pair (two pairs per warp): lane k1 < 25 runs the
            // 16-point FFT of row k1 and then holds Z[k1 + 25 k2], k2 = 0..15.  The mirror bin 400 - k sits in lane
            // 25 - k1 at index 15 - k2 (lane 0: its own index 16 - k2), so the split of the two packed real frames is
            // a pair of shuffles, and the power goes straight to P: no write-back of Z, no second pass.
            const int lane = tid & 31, k1 = lane < 25 ? lane : 0, src = lane == 0 ? 0 : (lane < 25 ? 25 - lane : 0);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int q = (tid >> 5) + 8 * h;
                const cpx* y = Y + q * kF400Plane + k1 * kF400Stride;
                cpx v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = y[i];
                fft_pow2<16>(v);
#pragma unroll
                for (int k2 = 0; k2 <= 8; ++k2) {
                    // general lanes: partner's Z[(25-k1) + 25 (15-k2)]; lane 0: own Z[25 (16-k2)] (k2 = 0: Z[0] itself)
                    const float sr = __shfl_sync(0xffffffffu, v[k2 < 8 ? 15 - k2 : 15].x, src);
                    const float si = __shfl_sync(0xffffffffu, v[k2 < 8 ? 15 - k2 : 15].y, src);
                    const cpx mc = lane == 0 ? cconj(v[(16 - k2) & 15]) : cpx{sr, -si};  // conj Z[400 - k]
                    const int k = k1 + 25 * k2;
                    if (lane < 25 && k < kBins) {
                        // 2 X_a = Z[k] + conj Z[N-k], 2i X_b = Z[k] - conj Z[N-k]: the tile holds 4 |X|^2 (the mel taps carry the 1/4)
                        const cpx sa = cadd(v[k2], mc), sb = csub(v[k2], mc);
                        *reinterpret_cast<float2*>(P + k * kPStride + 2 * q) = make_float2(fmaf(sa.x, sa.x, sa.y * sa.y), fmaf(sb.x, sb.x, sb.y * sb.y));
                    }
                }
            }
        }
        __syncthreads();
        // sparse mel contraction (each triangle touches a few bins; taps zero-padded to fours) + log10.  A thread owns the frames
        // (l, l + 16) of one mel row at a time: one 64-bit load and one packed FMA per tap serve both.
        const int l16 = tid & 15;
        const bool live0 = (t0 + l16) < a.n_frames, live1 = (t0 + 16 + l16) < a.n_frames;
        float vmax = -10.0f;
        float* outb = a.out + (long long)b * a.n_mels * a.n_frames + t0 + l16;
        const cpx* Pf = reinterpret_cast<const cpx*>(P) + l16;
        for (int m = tid >> 4; m < a.n_mels; m += 16) {
            const int meta = mel_meta[m];
            const int len4 = (meta >> 8) & 255;
            const float* w = mel_wsm + (meta >> 16);
            const cpx* pp = Pf + (meta & 255) * (kPStride / 2);
            cpx acc0 = cpx{0.f, 0.f}, acc1 = cpx{0.f, 0.f};
            for (int i = 0; i < len4; i += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + i);
                acc0 = cfma(pp[i * (kPStride / 2)], w4.x, acc0);
                acc1 = cfma(pp[(i + 1) * (kPStride / 2)], w4.y, acc1);
                acc0 = cfma(pp[(i + 2) * (kPStride / 2)], w4.z, acc0);
                acc1 = cfma(pp[(i + 3) * (kPStride / 2)], w4.w, acc1);
            }
            const cpx acc = cadd(acc0, acc1);
            // log10 via the SFU log2: |error| < 3e-6 on log10, 1e-6 on the output (tolerance 1e-4)
            const float v0 = __log2f(fmaxf(acc.x, 1e-10f)) * 0.30102999566398120f;
            const float v1 = __log2f(fmaxf(acc.y, 1e-10f)) * 0.30102999566398120f;
            if (live0) {
                outb[(long long)m * a.n_frames] = logmel_affine(v0);
                vmax = fmaxf(vmax, v0);
            }
            if (live1) {
                outb[(long long)m * a.n_frames + 16] = logmel_affine(v1);
                vmax = fmaxf(vmax, v1);
            }
        }
        vmax = warp_max(vmax);
        if ((tid & 31) == 0) atomicMax(a.gmax + b, __float_as_uint(vmax + 10.0f));
        __syncthreads();  // P (aliasing xs) is free again
    }
}

// ---------------------------------------------------------------- integer-staged variant: three CTAs per SM
// Whenever the samples entering the transform are 16-bit integers / 32768 -- pcm16 input, or any input that goes through the chain's
// WAV round trip (requant) -- the staged tile is kept as biased uint16 (10.7 KB instead of 21.4 KB), the power tile overwrites the
// transform planes of its own frame pair (no separate P), and the next tile's samples wait in REGISTERS (loaded while step 2 and the mel
// phase run) instead of a TMA buffer: 69.9 KB of shared memory and 80 registers, i.e. 24 warps per SM instead of 16.
//   smem: [Y / P: 16 x 425 complex] [win * 2^-15 | tw] [xs16: 5360 x uint16 = sample + 32768]
// Step 1 rebuilds the float from the uint16 with one byte permute (0x4B00'xxxx = 2^23 + u) and one packed subtract: no I2F.
constexpr int kLm16Smem = 16 * kF400Plane * 8 + 1200 * 4 + kXs * 2;  // 69,920 B
static_assert((16 * kF400Plane * 8 + 1200 * 4) % 16 == 0, "xs16 must be 16-byte aligned");

__device__ __forceinline__ float lg2_bare(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool F32>
__device__ __forceinline__ bool lm16_interior(const MelArgs& a, int b, int t0, const char** src) {
    const long long p0 = (long long)kHop * t0 - kNfft / 2;  // multiple of 8 samples
    const char* s = reinterpret_cast<const char*>(a.audio) + ((long long)b * a.stride + p0) * (F32 ? 4 : 2);
    *src = s;
    return p0 >= 0 && p0 + kXs <= a.n && (((uintptr_t)s) & 15) == 0;
}
// sample -> biased uint16 of what the reference's int16 round trip leaves (or of the pcm16 sample itself)
__device__ __forceinline__ uint32_t lm16_q(const MelArgs& a, float x, float gain) {
    return (uint32_t)(__float2int_rz(__fmul_rn(fminf(fmaxf(__fmul_rn(x, gain), -1.0f), 1.0f), 32767.0f)) + 32768);
}
__device__ __forceinline__ uint32_t lm16_q16(const MelArgs& a, int s16, float gain) {
    return a.requant ? lm16_q(a, (float)s16 * 3.0517578125e-05f, gain) : (uint32_t)(s16 + 32768);
}

template <bool F32>
__global__ void __launch_bounds__(256, 3) k_logmel16(MelArgs a, int tiles_per_clip, int total_tiles) {
    extern __shared__ __align__(16) float sm[];
    cpx* Y = reinterpret_cast<cpx*>(sm);                              // [16][425] complex; pair q's power tile later sits at Y_q as [204] x (frame q, frame q + 16)
    float* win = sm + 16 * kF400Plane * 2;                            // [400] Hann * 2^-15
    const cpx* tw = reinterpret_cast<const cpx*>(win + 400);          // [400] (cos, sin)
    uint16_t* xs16 = reinterpret_cast<uint16_t*>(win + 1200);         // [kXs]
    __shared__ int mel_meta[128];
    __shared__ __align__(16) float mel_wsm[kMelWMax];
    const int tid = threadIdx.x;
    constexpr int NV = F32 ? 6 : 3;              // 16-byte vectors per thread: 1,340 float4 or 670 uint4 per tile
    constexpr int kVecs = F32 ? kXs / 4 : kXs / 8;

    for (int i = tid; i < 400; i += 256) win[i] = a.consts[i] * 3.0517578125e-05f;  // exact: the staged samples are integers
    for (int i = tid; i < 800; i += 256) win[400 + i] = a.consts[400 + i];
    for (int m = tid; m < a.n_mels; m += 256) mel_meta[m] = a.mel_meta[m];
    for (int i = tid; i < a.n_w; i += 256) mel_wsm[i] = a.mel_w[i];

    // (clip, tile-in-clip) of the CTA's current tile and of the one it prefetches, advanced without divisions (two integer divisions
    // per tile were 3 % of the kernel's instructions)
    const int step_b = gridDim.x / tiles_per_clip, step_t = gridDim.x - step_b * tiles_per_clip;
    int cb = blockIdx.x / tiles_per_clip, ct = blockIdx.x - cb * tiles_per_clip;  // current
    int nb = cb, nt = ct;                                                        // prefetched
    uint4 pre[NV];
    bool pre_ok = false;
    auto prefetch = [&](int tile) {
        const char* src;
        pre_ok = tile < total_tiles && lm16_interior<F32>(a, nb, nt * MF, &src);
        if (pre_ok) {
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                const int i = tid + 256 * r;
                if (i < kVecs) pre[r] = ld_stream_u4(reinterpret_cast<const uint4*>(src) + i);
            }
        }
    };
    auto advance = [&](int& bb, int& tt) {
        bb += step_b;
        tt += step_t;
        if (tt >= tiles_per_clip) {
            tt -= tiles_per_clip;
            ++bb;
        }
    };
    prefetch(blockIdx.x);

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, advance(cb, ct)) {
        const int b = cb, t0 = ct * MF;
        {
            float gain = 1.0f;
            if (a.gain) {
                gain = a.gain[b];
                if (gain < 0.f) gain = 1.0f;  // silent clip: passed through (x * 1.0f is x; the WAV round trip clips it all the same)
            }
            if (pre_ok) {
#pragma unroll
                for (int r = 0; r < NV; ++r) {
                    const int i = tid + 256 * r;
                    if (i < kVecs) {
                        const uint32_t w[4] = {pre[r].x, pre[r].y, pre[r].z, pre[r].w};
                        if (F32) {
                            uint32_t o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) o[k] = lm16_q(a, __uint_as_float(w[k]), gain);
                            *reinterpret_cast<uint2*>(xs16 + 4 * i) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                        } else {
                            uint32_t o[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                o[k] = lm16_q16(a, (int)(int16_t)(w[k] & 0xFFFF), gain) | (lm16_q16(a, (int)(int16_t)(w[k] >> 16), gain) << 16);
                            *reinterpret_cast<uint4*>(xs16 + 8 * i) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
            } else {
                // clip edges: reflect pad 200 around [x, 160 zeros]
                const long long L = a.n + kPad;
                const long long p0 = (long long)kHop * t0 - kNfft / 2;
                for (int i = tid; i < kXs; i += 256) {
                    long long p = p0 + i;
                    while (p < 0 || p >= L) p = p < 0 ? -p : 2 * (L - 1) - p;
                    uint32_t v = 32768u;
                    if (p < a.n) {
                        if (F32) v = lm16_q(a, reinterpret_cast<const float*>(a.audio)[(long long)b * a.stride + p], gain);
                        else v = lm16_q16(a, (int)reinterpret_cast<const int16_t*>(a.audio)[(long long)b * a.stride + p], gain);
                    }
                    xs16[i] = (uint16_t)v;
                }
            }
        }
        __syncthreads();  // xs16 complete (and every thread is past the previous tile's mel phase: Y is free)
        {   // four-step FFT, step 1: 16 frame pairs x 16 residues = 256 tasks; pair q = frames (q, q + 16)
            const int q = tid >> 4, n2 = tid & 15;
            const uint16_t* xa = xs16 + q * kHop + n2;
            cpx v[25];
#pragma unroll
            for (int n1 = 0; n1 < 25; ++n1) {
                const int idx = 16 * n1;
                // 0x4B00'0000 | u is the float 2^23 + u: subtracting 2^23 + 2^15 leaves the signed sample, exactly
                const cpx raw2 = cpx{__uint_as_float(__byte_perm((uint32_t)xa[idx], 0x4B00u, 0x5410)),
                                     __uint_as_float(__byte_perm((uint32_t)xa[idx + 16 * kHop], 0x4B00u, 0x5410))};
                v[n1] = cscale(cadd(raw2, cpx{-8421376.0f, -8421376.0f}), win[idx + n2]);
            }
            dft25(v);
            cpx* y = Y + q * kF400Plane + n2;
#pragma unroll
            for (int k1 = 0; k1 < 25; ++k1) y[k1 * kF400Stride] = cmul_conj(v[k1], tw[k1 * 16 + n2]);
        }
        __syncthreads();
        if (!F32) {
            advance(nb, nt);
            prefetch(tile + gridDim.x);  // the next tile's samples travel while step 2 and the mel phase run
        }
        {   // step 2 fused with the power spectrum (see k_logmel); the powers of pair q overwrite its own planes
            const int lane = tid & 31, k1 = lane < 25 ? lane : 0, src = lane == 0 ? 0 : (lane < 25 ? 25 - lane : 0);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int q = (tid >> 5) + 8 * h;
                cpx* yq = Y + q * kF400Plane;
                cpx v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = yq[k1 * kF400Stride + i];
                __syncwarp();  // every lane holds its row: the planes of this pair may be overwritten
                fft_pow2<16>(v);
                float2* Pq = reinterpret_cast<float2*>(yq);
                if (lane >= 25 && lane < 28) Pq[kBins + lane - 25] = make_float2(0.f, 0.f);  // rows the zero-padded mel taps may touch
#pragma unroll
                for (int k2 = 0; k2 <= 8; ++k2) {
                    const float sr = __shfl_sync(0xffffffffu, v[k2 < 8 ? 15 - k2 : 15].x, src);
                    const float si = __shfl_sync(0xffffffffu, v[k2 < 8 ? 15 - k2 : 15].y, src);
                    const cpx mc = lane == 0 ? cconj(v[(16 - k2) & 15]) : cpx{sr, -si};  // conj Z[400 - k]
                    const int k = k1 + 25 * k2;
                    if (lane < 25 && k < kBins) {
                        const cpx sa = cadd(v[k2], mc), sb = csub(v[k2], mc);
                        Pq[k] = make_float2(fmaf(sa.x, sa.x, sa.y * sa.y), fmaf(sb.x, sb.x, sb.y * sb.y));
                    }
                }
            }
        }
        __syncthreads();
        if (F32) advance(nb, nt);
        if (F32) prefetch(tile + gridDim.x);  // float32 input: 24 registers of samples, requested once step 2's 16 complex values are dead
        // sparse mel contraction + log10: thread = (pair l16: frames l16 and l16 + 16, one mel row at a time)
        const int l16 = tid & 15;
        const bool live0 = (t0 + l16) < a.n_frames, live1 = (t0 + 16 + l16) < a.n_frames;
        float vmax = -10.0f;
        float* outb = a.out + (long long)b * a.n_mels * a.n_frames + t0 + l16;
        const cpx* Pf = Y + l16 * kF400Plane;
        float* outp = outb + (long long)(tid >> 4) * a.n_frames;
        const long long ostep = 16ll * a.n_frames;
        for (int m = tid >> 4; m < a.n_mels; m += 16, outp += ostep) {
            const int meta = mel_meta[m];
            int rounds = (meta >> 10) & 63;  // padded tap count / 4, at least 1
            const float4* w4p = reinterpret_cast<const float4*>(mel_wsm + (meta >> 16));
            const cpx* pp = Pf + (meta & 255);
            cpx acc0 = cpx{0.f, 0.f}, acc1 = cpx{0.f, 0.f};
            // most triangles are one round of four taps: a plain counted loop (ptxas unrolled the indexed form four times and paid for
            // the remainder ladder on every row)
#pragma unroll 1
            do {
                const float4 w4 = *w4p++;
                acc0 = cfma(pp[0], w4.x, acc0);
                acc1 = cfma(pp[1], w4.y, acc1);
                acc0 = cfma(pp[2], w4.z, acc0);
                acc1 = cfma(pp[3], w4.w, acc1);
                pp += 4;
            } while (--rounds > 0);
            const cpx acc = cadd(acc0, acc1);
            // log10 via the bare SFU log2 (the argument is >= 1e-10: no denormal fix-up): |error| < 3e-6 on log10, 1e-6 on the output
            const float v0 = lg2_bare(fmaxf(acc.x, 1e-10f)) * 0.30102999566398120f;
            const float v1 = lg2_bare(fmaxf(acc.y, 1e-10f)) * 0.30102999566398120f;
            if (live0) {
                outp[0] = logmel_affine(v0);
                vmax = fmaxf(vmax, v0);
            }
            if (live1) {
                outp[16] = logmel_affine(v1);
                vmax = fmaxf(vmax, v1);
            }
        }
        vmax = warp_max(vmax);
        if ((tid & 31) == 0) atomicMax(a.gmax + b, __float_as_uint(vmax + 10.0f));
        // (no barrier here: the next tile's staging only writes xs16, and its step 1 follows a barrier)
    }
}

// per-clip gain of the fused normalise, once per clip instead of once per thread and tile (log10f + powf + a double division)
__global__ void k_mel_gain(const unsigned long long* __restrict__ sumsq, const double* __restrict__ sumsq_f, long long n, float target_dbfs,
                           float* __restrict__ gain, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    bool silent = true;
    const float g = sumsq ? gain_from_meansq((double)sumsq[b] / 1073741824.0 / (double)n, target_dbfs, &silent)
                          : gain_from_meansq(sumsq_f[b] / (double)n, target_dbfs, &silent);
    gain[b] = silent ? -1.0f : g;  // a real gain is 10^x > 0
}

// log_spec = maximum(log_spec, log_spec.max() - 8.0) on the already affine-mapped features: reads every cell, writes the 16-byte groups the
// clamp changes (normalise-only clips: a few per cent; denoised clips: most -- gated cells are far below the threshold)
__device__ __forceinline__ bool logmel_clamp4(float4& x, float ty) {
    const float4 m = make_float4(fmaxf(x.x, ty), fmaxf(x.y, ty), fmaxf(x.z, ty), fmaxf(x.w, ty));
    const bool changed = m.x != x.x || m.y != x.y || m.z != x.z || m.w != x.w;
    x = m;
    return changed;
}

__global__ void __launch_bounds__(256) k_logmel_finalize(float* __restrict__ out, long long per_clip, const unsigned int* __restrict__ gmax) {
    const float ty = logmel_affine((__uint_as_float(gmax[blockIdx.y]) - 10.0f) - 8.0f);
    float* o = out + (long long)blockIdx.y * per_clip;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const bool aligned = (((uintptr_t)o) & 15) == 0;
    const long long nvec = aligned ? per_clip / 4 : 0;
    float4* o4 = reinterpret_cast<float4*>(o);
    long long v = tid;
    for (; v + 3 * nthr < nvec; v += 4 * nthr) {  // four independent 16-byte loads in flight per thread
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = o4[v + j * nthr];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (logmel_clamp4(x[j], ty)) o4[v + j * nthr] = x[j];
    }
    for (; v < nvec; v += nthr) {
        float4 x = o4[v];
        if (logmel_clamp4(x, ty)) o4[v] = x;
    }
    for (long long i = nvec * 4 + tid; i < per_clip; i += nthr) o[i] = fmaxf(o[i], ty);
}

constexpr int kLogmelSmemF32 = (kXsP + 1200 + 16 * 2 * kF400Plane) * (int)sizeof(float) + kRawBytes;  // (a complex plane = 2 x 425 floats)
constexpr int kLogmelSmemP16 = kLogmelSmemF32 - kRawBytes / 2;

// d_sumsq (pcm16 input) / d_sumsq_f (float32 input): fuse normalize_gain in front; requant: the chain's int16 round trip
int launch_logmel(const void* d_audio, int fmt, long long n, long long batch, long long stride, int n_mels, float* d_out,
                  const unsigned long long* d_sumsq, float target_dbfs, cudaStream_t st, const double* d_sumsq_f, int requant) {
    const MelTables* t;
    int rc = get_mel(n_mels, true, &t);
    if (rc) return rc;
    static PerDeviceOnce once;
    OSB_CUDA(once.run([&] {
        cudaError_t e = cudaFuncSetAttribute(k_logmel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLogmelSmemF32);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_logmel16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLm16Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_logmel16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLm16Smem);
        return e;
    }));
    const int n_frames = (int)((n + kPad) / kHop);
    Scratch scr(st);
    unsigned int* gmax;
    OSB_CUDA(scr.alloc(&gmax, (size_t)batch));
    OSB_CUDA(cudaMemsetAsync(gmax, 0, sizeof(unsigned int) * batch, st));
    MelArgs a;
    a.audio = d_audio; a.n = n; a.stride = stride; a.fmt = fmt; a.n_frames = n_frames; a.n_mels = n_mels;
    a.out = d_out; a.gmax = gmax; a.target_dbfs = target_dbfs; a.gain = nullptr;
    if (d_sumsq || d_sumsq_f) {
        float* gain;
        OSB_CUDA(scr.alloc(&gain, (size_t)batch));
        OSB_LAUNCH(k_mel_gain, (unsigned)((batch + 127) / 128), 128, 0, st, d_sumsq, d_sumsq_f, n, target_dbfs, gain, (int)batch);
        OSB_CHECK_LAUNCH();
        a.gain = gain;
    }
    a.requant = (d_sumsq || d_sumsq_f || requant) ? 1 : 0;
    a.consts = t->d_consts; a.mel_meta = t->d_meta; a.mel_w = t->d_w; a.n_w = t->n_w;
    const int tiles_per_clip = (n_frames + MF - 1) / MF;
    const long long total_tiles_ll = (long long)tiles_per_clip * batch;
    if (total_tiles_ll > 0x7fffffffLL) {
        set_error("invalid argument: too many log-mel tiles");
        return OSB_ERR_INVALID_ARG;
    }
    const int total_tiles = (int)total_tiles_ll;
    if (fmt == OSB_FMT_PCM16 || a.requant) {
        // integer-staged kernel: 3 resident CTAs per SM (70 KB of shared memory, 80 registers)
        const int persistent = 3 * take_sm_budget();
        const int grid = total_tiles < persistent ? total_tiles : persistent;
        if (fmt == OSB_FMT_PCM16) OSB_LAUNCH(k_logmel16<false>, grid, 256, kLm16Smem, st, a, tiles_per_clip, total_tiles);
        else OSB_LAUNCH(k_logmel16<true>, grid, 256, kLm16Smem, st, a, tiles_per_clip, total_tiles);
    } else {
        // float32 audio taken as it is: float staging, 2 resident CTAs per SM (108 KB of shared memory)
        const int persistent = 2 * take_sm_budget();
        const int grid = total_tiles < persistent ? total_tiles : persistent;
        OSB_LAUNCH(k_logmel<true>, grid, 256, kLogmelSmemF32, st, a, tiles_per_clip, total_tiles);
    }
    OSB_CHECK_LAUNCH();
    const long long per_clip = (long long)n_mels * n_frames;
    long long fb = (per_clip / 4 + 255) / 256;
    long long want = ((long long)OSB_NUM_SMS * 8 + batch - 1) / batch;
    if (fb > want) fb = want;
    if (fb < 1) fb = 1;
    OSB_LAUNCH(k_logmel_finalize, dim3((unsigned)fb, (unsigned)batch), 256, 0, st, d_out, per_clip, gmax);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_logmel_frames(int64_t n_samples) { return n_samples < 0 ? 0 : (int)((n_samples + kPad) / kHop); }

int osb_mel_filters(int n_mels, float* out, size_t capacity) {
    OSB_REQUIRE(n_mels == 80 || n_mels == 128, "n_mels must be 80 or 128");
    OSB_REQUIRE(out && capacity >= (size_t)n_mels * kBins, "mel filter buffer too small");
    const MelTables* t;
    int rc = get_mel(n_mels, false, &t);
    if (rc) return rc;
    memcpy(out, t->dense.data(), sizeof(float) * n_mels * kBins);
    return OSB_OK;
}

int osb_logmel_dev(const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, int n_mels, float* d_out,
                   int fuse_normalize, float target_dbfs, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    OSB_REQUIRE(n_mels == 80 || n_mels == 128, "n_mels must be 80 or 128");
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n, "bad sizes");
    OSB_REQUIRE(!fuse_normalize || fmt == OSB_FMT_PCM16, "fused normalise needs pcm16 input");
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    if (batch == 0) return OSB_OK;
    OSB_REQUIRE(n + kPad > kNfft / 2, "clip too short for reflect padding (needs n + 160 > 200)");
    OSB_REQUIRE(d_audio && d_out, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    unsigned long long* sumsq = nullptr;
    if (fuse_normalize) {
        OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
        if ((rc = launch_sumsq_pcm16((const int16_t*)d_audio, n, batch, stride, sumsq, st))) return rc;
    }
    return launch_logmel(d_audio, fmt, n, batch, stride, n_mels, d_out, sumsq, target_dbfs, st);
}

int osb_logmel_host(const void* audio, int fmt, int64_t n, int n_mels, float* out, int fuse_normalize, float target_dbfs) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(fmt == OSB_FMT_PCM16 || fmt == OSB_FMT_F32, "fmt must be OSB_FMT_PCM16 or OSB_FMT_F32");
    OSB_REQUIRE(n >= 0 && out, "bad arguments");
    const size_t es = fmt == OSB_FMT_PCM16 ? 2 : 4;
    const size_t ob = (size_t)n_mels * osb_logmel_frames(n) * 4;
    void *da, *dout;
    if ((rc = ws.dev_buf(0, (size_t)n * es + 16, &da)) || (rc = ws.dev_buf(1, ob + 16, &dout))) return rc;
    if ((rc = ws.h2d(da, audio, (size_t)n * es))) return rc;
    if ((rc = osb_logmel_dev(da, fmt, n, 1, n, n_mels, (float*)dout, fuse_normalize, target_dbfs, ws.stream))) return rc;
    return ws.d2h(out, dout, ob);
}

}  // extern "C"
