// TTS output post-processing, the effects chain, Kokoro voice-pack blending.
//
// Replaces (reference file:line):
//   trim_silence / normalize_output / process_tts_chunks   src/audio/postprocessing.py:8-40
//   apply_chain: _normalize, _reverb, _podcast_eq, _robot  src/effects/chain.py:15-74
//   KokoroBackend._blend_voices                            src/tts/backends/kokoro.py:289-308
//
// Ragged batches: utterance b lives at [offsets[b], offsets[b] + len[b]) of a flat buffer.
// Numerics follow the reference's dtype flow: float32 until the first float64-producing effect
// (reverb / podcast_eq / robot), float64 afterwards, one cast to float32 at the end (chain.py:32).
// The recurrences (exponential-IR reverb, two biquads) are linear: each CTA runs its own stretch of
// samples from a state that is exact to f64 round-off (direct FIR sum for the reverb; a 4096-sample
// zero-state warm-up for the biquads, whose poles decay below 1e-26 over that span), so all CTAs
// are independent and the time axis is parallel.
#include <cmath>
#include <cstdlib>
#include <vector>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace osb {

struct Ragged {
    const long long* offsets;  // [B]   start of utterance b in the flat buffer
    const long long* lens;     // [B]   current length of utterance b
};

// ---------------------------------------------------------------- trim + peak normalise
struct TtsStats {
    int first, last;      // first / last index with |x| > threshold (first = INT_MAX when none)
    unsigned int maxbits; // bits of max |x| (non-negative float -> ordered as uint)
    int pad;
    double sumsq;         // k_tts_stats<true>: sum of the float32 squares of ALL samples of the utterance, summed wide
};

// Utterances start anywhere in the flat buffer: a scalar head up to the first 16-byte boundary, a float4 body, a scalar tail.
struct Span4 {
    long long head, nvec;  // scalars before the aligned body; float4 groups in it
};
__device__ __forceinline__ Span4 span4(const float* p, long long n) {
    long long head = (long long)((16u - (unsigned)((uintptr_t)p & 15u)) & 15u) >> 2;
    if (head > n) head = n;
    return Span4{head, (n - head) >> 2};
}

// SUMSQ: also the utterance's sum of squares, so that an RMS normalise at the head of the effects chain needs no pass of its own
// (k_fx_sumsq_from_stats takes the trimmed edges off and applies the peak gain)
template <bool SUMSQ>
__global__ void __launch_bounds__(256) k_tts_stats(const float* __restrict__ x, Ragged rg, float thr, TtsStats* __restrict__ st) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const float* p = x + rg.offsets[b];
    int first = 0x7fffffff, last = -1;
    float mx = 0.f;
    double acc = 0.0;
    auto see = [&](float v, long long i) {
        const float a = fabsf(v);
        mx = fmaxf(mx, a);
        if (SUMSQ) acc += (double)__fmul_rn(v, v);
        if (a > thr) {
            first = min(first, (int)i);
            last = max(last, (int)i);
        }
    };
    const Span4 sp = span4(p, n);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const float4* p4 = reinterpret_cast<const float4*>(p + sp.head);
    long long v = tid;
    for (; v + 3 * nthr < sp.nvec; v += 4 * nthr) {  // four independent 16-byte loads in flight per thread
        float4 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = ld_stream_f4(p4 + v + j * nthr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = sp.head + 4 * (v + j * nthr);
            see(q[j].x, i); see(q[j].y, i + 1); see(q[j].z, i + 2); see(q[j].w, i + 3);
        }
    }
    for (; v < sp.nvec; v += nthr) {
        const float4 q = ld_stream_f4(p4 + v);
        const long long i = sp.head + 4 * v;
        see(q.x, i); see(q.y, i + 1); see(q.z, i + 2); see(q.w, i + 3);
    }
    if (tid < sp.head) see(p[tid], tid);
    for (long long i = sp.head + 4 * sp.nvec + tid; i < n; i += nthr) see(p[i], i);
    first = warp_min_i(first);
    last = warp_max_i(last);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) {
        if (last >= 0) {
            atomicMin(&st[b].first, first);
            atomicMax(&st[b].last, last);
        }
        atomicMax(&st[b].maxbits, __float_as_uint(mx));
    }
    if (SUMSQ) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(&st[b].sumsq, acc);
    }
}

__global__ void k_tts_stats_init(TtsStats* st, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) st[i] = TtsStats{0x7fffffff, -1, 0u, 0, 0.0};
}

// sumsq (optional): per-utterance sum of squares of the samples written (float32 squares summed wide, as k_fx_sumsq
// does): an RMS normalise at the head of the effects chain then needs no pass of its own
__global__ void __launch_bounds__(256) k_tts_apply(const float* __restrict__ x, Ragged rg, const TtsStats* __restrict__ st, int trim,
                                                   int normalize, float peak, float* __restrict__ y, long long* __restrict__ out_len,
                                                   double* __restrict__ sumsq) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const TtsStats s = st[b];
    long long start = 0, m = n;
    if (trim && n > 0 && s.last >= 0) {  // nothing above the threshold -> returned unchanged (postprocessing.py:12-13)
        start = s.first;
        m = (long long)s.last - s.first + 1;
    }
    const float mx = __uint_as_float(s.maxbits);
    const bool scale_on = normalize && m > 0 && mx > 1e-8f;  // float(max) <= 1e-8 -> unchanged (:21-22)
    const float scale = scale_on ? (float)((double)peak / (double)mx) : 1.0f;
    const float* p = x + rg.offsets[b] + start;
    float* q = y + rg.offsets[b];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const float v = p[i];
        const float o = scale_on ? fminf(fmaxf(__fmul_rn(v, scale), -1.0f), 1.0f) : v;
        q[i] = o;
        acc += (double)__fmul_rn(o, o);
    }
    if (sumsq) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sumsq[b], acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out_len[b] = m;
}

// trim_silence / normalize_output as a plan instead of a pass: where the kept span starts, how long it is and the
// peak gain, so that the first kernel of the effects chain can read the untouched input (osb_tts_post_fx_dev)
__global__ void k_tts_plan(Ragged rg, const TtsStats* __restrict__ st, int batch, int trim, int normalize, float peak,
                           long long* __restrict__ offs2, long long* __restrict__ lens2, float* __restrict__ pscale, int* __restrict__ pon) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const long long n = rg.lens[b];
    const TtsStats s = st[b];
    long long start = 0, m = n;
    if (trim && n > 0 && s.last >= 0) {
        start = s.first;
        m = (long long)s.last - s.first + 1;
    }
    const float mx = __uint_as_float(s.maxbits);
    const bool on = normalize && m > 0 && mx > 1e-8f;
    offs2[b] = rg.offsets[b] + start;
    lens2[b] = m;
    pscale[b] = on ? (float)((double)peak / (double)mx) : 1.0f;
    pon[b] = on ? 1 : 0;
}

// sum of squares of the post-processed utterance without materialising it (same arithmetic as k_tts_apply + k_fx_sumsq)
__global__ void __launch_bounds__(256) k_fx_sumsq_post(const float* __restrict__ x, Ragged rg, const float* __restrict__ pscale,
                                                       const int* __restrict__ pon, double* __restrict__ sumsq) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const float* p = x + rg.offsets[b];
    const bool on = pon[b] != 0;
    const float sc = pscale[b];
    double acc = 0.0;
    auto add = [&](float v) {
        const float o = on ? fminf(fmaxf(__fmul_rn(v, sc), -1.0f), 1.0f) : v;
        acc += (double)__fmul_rn(o, o);
    };
    const Span4 sp = span4(p, n);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    const float4* p4 = reinterpret_cast<const float4*>(p + sp.head);
    long long v = tid;
    for (; v + 3 * nthr < sp.nvec; v += 4 * nthr) {
        float4 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = ld_stream_f4(p4 + v + j * nthr);
#pragma unroll
        for (int j = 0; j < 4; ++j) { add(q[j].x); add(q[j].y); add(q[j].z); add(q[j].w); }
    }
    for (; v < sp.nvec; v += nthr) {
        const float4 q = ld_stream_f4(p4 + v);
        add(q.x); add(q.y); add(q.z); add(q.w);
    }
    if (tid < sp.head) add(p[tid]);
    for (long long i = sp.head + 4 * sp.nvec + tid; i < n; i += nthr) add(p[i]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sumsq[b], acc);
}

// The same quantity from the statistics pass (k_tts_stats<true>): sum of squares of the whole utterance, minus the two trimmed edges (at
// most a few hundred milliseconds each: this kernel reads only those), times the square of the peak gain.  Against squaring the gained
// samples one by one this moves the sum by ~1e-7 relative (float32 roundings of x * gain), four decimal orders inside the tolerance, and
// saves a full pass over the input.  One CTA per utterance.
__global__ void __launch_bounds__(256) k_fx_sumsq_from_stats(const float* __restrict__ x, const long long* __restrict__ offsets,
                                                             const long long* __restrict__ lens, const long long* __restrict__ offs2,
                                                             const long long* __restrict__ lens2, const TtsStats* __restrict__ st,
                                                             const float* __restrict__ pscale, double* __restrict__ sumsq) {
    const int b = blockIdx.x;
    const float* p = x + offsets[b];
    const long long n = lens[b], head = offs2[b] - offsets[b], tail0 = head + lens2[b];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < head; i += 256) acc += (double)__fmul_rn(p[i], p[i]);
    for (long long i = tail0 + threadIdx.x; i < n; i += 256) acc += (double)__fmul_rn(p[i], p[i]);
    __shared__ double part[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double edges = 0.0;
        for (int w = 0; w < 8; ++w) edges += part[w];
        const double kept = fmax(st[b].sumsq - edges, 0.0), sc = (double)pscale[b];
        sumsq[b] = kept * sc * sc;
    }
}

// ---------------------------------------------------------------- effects: normalise (RMS), robot, cast
template <typename T>
__global__ void __launch_bounds__(256) k_fx_sumsq(const T* __restrict__ x, Ragged rg, double* __restrict__ sumsq) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const T* p = x + rg.offsets[b];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T v = p[i];
        acc += (sizeof(T) == 4) ? (double)__fmul_rn((float)v, (float)v) : (double)v * (double)v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sumsq[b], acc);
}

// samples * (target_rms / rms); rms < 1e-8 -> unchanged; f32 input stays f32 (np.float32 scalar arithmetic)
template <typename T>
__global__ void __launch_bounds__(256) k_fx_scale(const T* __restrict__ x, Ragged rg, const double* __restrict__ sumsq, double target_rms,
                                                  T* __restrict__ y) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    if (n == 0) return;
    const T* p = x + rg.offsets[b];
    T* q = y + rg.offsets[b];
    bool on;
    T scale;
    if (sizeof(T) == 4) {
        const float rms = __fsqrt_rn((float)(sumsq[b] / (double)n));
        on = !(rms < 1e-8f);
        scale = (T)((float)target_rms / rms);  // python float / np.float32 -> np.float32
    } else {
        const double rms = sqrt(sumsq[b] / (double)n);
        on = !(rms < 1e-8);
        scale = (T)(target_rms / rms);
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        q[i] = on ? (T)(p[i] * scale) : p[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_fx_robot(const T* __restrict__ x, Ragged rg, double sr, double* __restrict__ y) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const T* p = x + rg.offsets[b];
    double* q = y + rg.offsets[b];
    const double w = 2.0 * 3.141592653589793 * 100.0;  // 2*np.pi*100
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        q[i] = (double)p[i] * sin(__dmul_rn(w, __ddiv_rn((double)i, sr)));
}

// final astype(float32) (f64 -> f32) or f32 copy; optionally clip/x32767/truncate to int16 (float32_to_int16)
template <typename T, bool PCM16>
__global__ void __launch_bounds__(256) k_fx_finish(const T* __restrict__ x, Ragged rg, void* __restrict__ y, const long long* __restrict__ out_offsets) {
    const int b = blockIdx.y;
    const long long n = rg.lens[b];
    const T* p = x + rg.offsets[b];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = (float)p[i];
        const long long o = (out_offsets ? out_offsets[b] : rg.offsets[b]) + i;
        if (PCM16) reinterpret_cast<int16_t*>(y)[o] = (int16_t)quant_pcm16(v);
        else reinterpret_cast<float*>(y)[o] = v;
    }
}

// ---------------------------------------------------------------- fusing the element-wise effects into their neighbours
// A normalise that precedes reverb / podcast_eq on float32 data is not materialised: the recurrence kernel multiplies
// by the per-utterance float32 gain while it loads (FxPre), which yields exactly the float32 values k_fx_scale would
// have stored.  A robot effect and the final cast that follow a recurrence kernel are applied to the float64 value it
// is about to store (FxPost): same arithmetic as k_fx_robot / k_fx_finish, one HBM round trip less each.
struct FxPre {
    const double* sumsq;  // per-utterance sum of squares (null: no deferred normalise)
    double target_rms;
    // deferred normalize_output (src/audio/postprocessing.py:17-23) of the utterance: x -> clip(x * peak_scale, -1, 1) where
    // peak_on; applied before the gain above.  Null: the input is already post-processed.
    const float* peak_scale;
    const int* peak_on;
};
struct PreOps {
    bool peak_on, gain_on;
    float peak_scale, gain;
};
struct FxPost {
    int robot;    // multiply by the 100 Hz carrier sin(2 pi 100 n / sr) in float64
    int finish;   // 0: store float64 to the chain buffer; 1: astype(float32) -> out; 2: float32_to_int16 -> out
    int period;   // carrier period in samples, sr / gcd(100, sr)
    const double* carrier;  // [period] sin(2*pi*100*(k/sr)) evaluated like numpy for n = k (host table); for n = k + m*period
                            // numpy's own value differs from it by the rounding of its growing argument (~1e-13)
    void* out;
    const long long* out_offsets;  // where utterance b starts in `out` (null: at the input offsets)
};

__device__ __forceinline__ PreOps fx_pre_ops(const FxPre& pre, int b, long long n) {
    PreOps o{false, false, 1.0f, 1.0f};
    if (pre.peak_on && pre.peak_on[b]) { o.peak_on = true; o.peak_scale = pre.peak_scale[b]; }
    if (pre.sumsq && n > 0) {
        const float rms = __fsqrt_rn((float)(pre.sumsq[b] / (double)n));
        o.gain_on = !(rms < 1e-8f);
        o.gain = (float)pre.target_rms / rms;  // python float / np.float32 -> np.float32
    }
    return o;
}
// the float32 value the materialising kernels (k_tts_apply, k_fx_scale) would have stored for this sample
__device__ __forceinline__ float fx_pre(float x, const PreOps& o) {
    if (o.peak_on) x = fminf(fmaxf(__fmul_rn(x, o.peak_scale), -1.0f), 1.0f);
    if (o.gain_on) x = __fmul_rn(x, o.gain);
    return x;
}
template <typename T>
__device__ __forceinline__ double fx_pre_d(T x, const PreOps& o) {
    return sizeof(T) == 4 ? (double)fx_pre((float)x, o) : (double)x;
}
__device__ __forceinline__ long long fx_out_off(const FxPost& post, const Ragged& rg, int b) {
    return post.out_offsets ? post.out_offsets[b] : rg.offsets[b];
}
__device__ __forceinline__ void fx_st(const FxPost& post, double* q, long long off, long long g, double v) {
    if (post.robot) v = v * post.carrier[(unsigned)g % (unsigned)post.period];  // utterances are shorter than 2^31 samples (checked by the caller)
    if (post.finish == 0) q[g] = v;
    else if (post.finish == 1) reinterpret_cast<float*>(post.out)[off + g] = (float)v;
    else reinterpret_cast<int16_t*>(post.out)[off + g] = (int16_t)quant_pcm16((float)v);
}

// ---------------------------------------------------------------- reverb: exp-decay FIR as a one-pole recurrence
// wet[n] = sum_{k<L} ir[k] x[n-k], ir[k] = c r^k  =>  wet[n] = r wet[n-1] + c (x[n] - r^L x[n-L])
// out = (1-mix) x + mix wet.  CTA = 8192 samples; wet[n0-1] comes from the direct FIR sum with the exact
// host-computed taps, so CTAs are independent.
constexpr int kRvT = 32, kRvBlock = 256 * kRvT;  // 8192 samples per CTA
constexpr int kSegStride = kRvT + 1;             // padded per-thread segment (bank-conflict free for f64)

struct ReverbArgs {
    Ragged rg;
    const double* ir;  // [L] host-computed exp(-linspace(0,6,L))/sum
    int L;
    double r, c, rL, rT, mix;
};

template <typename T>
__global__ void __launch_bounds__(256) k_fx_reverb(const T* __restrict__ x, ReverbArgs a, FxPre pre, FxPost post, double* __restrict__ y) {
    extern __shared__ __align__(16) double smd[];
    double* u = smd;                 // [256][33]
    __shared__ double red[8];
    __shared__ double wsum[8];
    const int b = blockIdx.y, tid = threadIdx.x;
    const long long n = a.rg.lens[b];
    const long long n0 = (long long)blockIdx.x * kRvBlock;
    if (n0 >= n) return;
    const T* p = x + a.rg.offsets[b];
    const PreOps po = fx_pre_ops(pre, b, n);
    // stage u[i] = c (x[n0+i] - r^L x[n0+i-L]); eight samples' loads in flight per thread
#pragma unroll 1
    for (int i0 = tid; i0 < kRvBlock; i0 += 256 * 8) {
        T xv[8], xd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long g = n0 + i0 + 256 * j;
            xv[j] = g < n ? p[g] : (T)0;
            xd[j] = (g - a.L >= 0 && g - a.L < n) ? p[g - a.L] : (T)0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + 256 * j;
            // zeros beyond the utterance stay zeros (the pre-ops map 0 to 0)
            const double v = fx_pre_d(xv[j], po), d = fx_pre_d(xd[j], po);
            u[(i / kRvT) * kSegStride + (i % kRvT)] = a.c * (v - a.rL * d);
        }
    }
    // carry-in wet[n0-1] = sum_k ir[k] x[n0-1-k]; four taps' loads in flight per thread
    double part = 0.0;
    if (n0 > 0) {
#pragma unroll 1
        for (int k0 = tid; k0 < a.L; k0 += 256 * 4) {
            T xv[4];
            double iv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + 256 * j;
                const long long g = n0 - 1 - k;
                const bool ok = k < a.L && g >= 0;
                xv[j] = ok ? p[g] : (T)0;
                iv[j] = ok ? a.ir[k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                part = fma(iv[j], fx_pre_d(xv[j], po), part);
            }
        }
    }
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    const double carry0 = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    // phase A: zero-state end value of each thread's 32-sample segment
    double* seg = u + tid * kSegStride;
    double e = 0.0;
#pragma unroll
    for (int i = 0; i < kRvT; ++i) e = fma(a.r, e, seg[i]);
    // inclusive warp scan with factor rT^(2^k)
    double f = a.rT, v = e;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fma(f, up, v);
        f *= f;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    // f == rT^32 now; carry into this warp = state after the previous warps, starting from carry0
    double cw = carry0;
    for (int w = 0; w < wid; ++w) cw = fma(f, cw, wsum[w]);
    // state at the start of this thread's segment: rT^lane * cw + (exclusive prefix)
    double ex = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) ex = 0.0;
    double pw = 1.0, bb = a.rT;
    for (int l = lane; l > 0; l >>= 1) {
        if (l & 1) pw *= bb;
        bb *= bb;
    }
    double s = fma(pw, cw, ex);
    // phase C: real run, outputs written back in place
#pragma unroll
    for (int i = 0; i < kRvT; ++i) {
        s = fma(a.r, s, seg[i]);
        seg[i] = s;
    }
    __syncthreads();
    double* q = y + a.rg.offsets[b];
#pragma unroll 1
    for (int i0 = tid; i0 < kRvBlock; i0 += 256 * 8) {
        T xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long g = n0 + i0 + 256 * j;
            xv[j] = g < n ? p[g] : (T)0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + 256 * j;
            const long long g = n0 + i;
            if (g < n) {
                // (1 - mix) * samples is a float32 product while samples is still float32 (NEP 50 weak scalar)
                const double xg = fx_pre_d(xv[j], po);
                const double dry = (sizeof(T) == 4) ? (double)__fmul_rn((float)(1.0 - a.mix), (float)xg) : (1.0 - a.mix) * xg;
                fx_st(post, q, fx_out_off(post, a.rg, b), g, dry + a.mix * u[(i / kRvT) * kSegStride + (i % kRvT)]);
            }
        }
    }
}

// ---------------------------------------------------------------- podcast EQ: two DF2T biquads, 4-state linear system
constexpr int kEqBlock = 8192, kEqWarmMax = 6144;  // samples per CTA (warm-up + outputs); the warm-up is sized per sample rate

struct EqArgs {
    Ragged rg;
    double b1[3], a1[3], b2[3], a2[3];
    int warm;         // warm-up samples (multiple of 32): |slowest pole|^warm < 1e-17
    double K[4][32];  // K[r][i]: state r at the end of a 32-sample segment for a unit impulse at its sample i (zero state before)
    double P[5][16];  // Phi_T^(2^k), k = 0..4 (row-major 4x4), T = 32 samples
    double Q[16];     // Phi_T^32
    double Qp[4][16]; // Q^(2^k), k = 0..3
};

struct St4 {
    double z[4];
};

__device__ __forceinline__ St4 mat4(const double* m, const St4& s) {
    St4 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) o.z[i] = ((m[4 * i] * s.z[0] + m[4 * i + 1] * s.z[1]) + (m[4 * i + 2] * s.z[2] + m[4 * i + 3] * s.z[3]));
    return o;
}

__device__ __forceinline__ double eq_step(const EqArgs& a, St4& s, double x) {
    // scipy lfilter, direct form II transposed, a[0] = 1
    const double y1 = a.b1[0] * x + s.z[0];
    s.z[0] = a.b1[1] * x - a.a1[1] * y1 + s.z[1];
    s.z[1] = a.b1[2] * x - a.a1[2] * y1;
    const double y2 = a.b2[0] * y1 + s.z[2];
    s.z[2] = a.b2[1] * y1 - a.a2[1] * y2 + s.z[3];
    s.z[3] = a.b2[2] * y1 - a.a2[2] * y2;
    return y2;
}

template <typename T>
__global__ void __launch_bounds__(256) k_fx_eq(const T* __restrict__ x, EqArgs a, FxPre pre, FxPost post, double* __restrict__ y) {
    extern __shared__ __align__(16) double smd[];
    double* u = smd;  // [256][33]
    __shared__ St4 wsum[8];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long n = a.rg.lens[b];
    const int kEqWarm = a.warm, kEqOut = kEqBlock - a.warm;
    const long long n0 = (long long)blockIdx.x * kEqOut;
    if (n0 >= n) return;
    const T* p = x + a.rg.offsets[b];
    const long long base = n0 - kEqWarm;
    const PreOps po = fx_pre_ops(pre, b, n);
#pragma unroll 1
    for (int i0 = tid; i0 < kEqWarm + kEqOut; i0 += 256 * 8) {
        T xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long g = base + i0 + 256 * j;
            xv[j] = (g >= 0 && g < n) ? p[g] : (T)0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + 256 * j;
            u[(i / 32) * kSegStride + (i % 32)] = fx_pre_d(xv[j], po);
        }
    }
    __syncthreads();
    double* seg = u + tid * kSegStride;
    // zero-state end state of the segment = four 32-tap dot products with the impulse-to-state table (the system is
    // linear): four independent FMA chains instead of 32 dependent filter steps
    St4 e{{0, 0, 0, 0}};
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const double xv = seg[i];
#pragma unroll
        for (int r = 0; r < 4; ++r) e.z[r] = fma(a.K[r][i], xv, e.z[r]);
    }
    // inclusive warp scan over segments: v_j = Phi^(2^k) v_{j-2^k} + v_j
    St4 v = e;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        St4 up;
#pragma unroll
        for (int i = 0; i < 4; ++i) up.z[i] = __shfl_up_sync(0xffffffffu, v.z[i], 1 << k);
        if (lane >= (1 << k)) {
            const St4 t = mat4(a.P[k], up);
#pragma unroll
            for (int i = 0; i < 4; ++i) v.z[i] += t.z[i];
        }
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    St4 cw{{0, 0, 0, 0}};  // the block starts from rest, kEqWarm samples before its first output
    for (int w = 0; w < wid; ++w) {
        const St4 t = mat4(a.Q, cw);
#pragma unroll
        for (int i = 0; i < 4; ++i) cw.z[i] = t.z[i] + wsum[w].z[i];
    }
    // fold the warp carry into lane 0 and rescan: v'_j = true end state of segment j
    St4 e2 = e;
    if (lane == 0) {
        const St4 t = mat4(a.P[0], cw);
#pragma unroll
        for (int i = 0; i < 4; ++i) e2.z[i] += t.z[i];
    }
    v = e2;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        St4 up;
#pragma unroll
        for (int i = 0; i < 4; ++i) up.z[i] = __shfl_up_sync(0xffffffffu, v.z[i], 1 << k);
        if (lane >= (1 << k)) {
            const St4 t = mat4(a.P[k], up);
#pragma unroll
            for (int i = 0; i < 4; ++i) v.z[i] += t.z[i];
        }
    }
    St4 s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double pv = __shfl_up_sync(0xffffffffu, v.z[i], 1);
        s.z[i] = lane == 0 ? cw.z[i] : pv;
    }
#pragma unroll 4
    for (int i = 0; i < 32; ++i) seg[i] = eq_step(a, s, seg[i]);
    __syncthreads();
    double* q = y + a.rg.offsets[b];
    for (int i = kEqWarm + tid; i < kEqWarm + kEqOut; i += 256) {
        const long long g = base + i;
        if (g < n) fx_st(post, q, fx_out_off(post, a.rg, b), g, u[(i / 32) * kSegStride + (i % 32)]);
    }
}

// ---------------------------------------------------------------- reverb -> podcast EQ in one kernel
// The chain [.. reverb, podcast_eq ..] exchanges float64 samples between the two recurrences; here they never leave the
// registers.  CTA = kEqBlock consecutive samples (EQ warm-up + outputs), 512 threads, thread = one 16-sample segment.
// The reverb runs over the whole block from its exact FIR carry-in (so the warm-up stretch holds true reverb output), is
// mixed with the dry signal, then the biquads run from rest as in k_fx_eq.  Same arithmetic per sample as k_fx_reverb
// followed by k_fx_eq.
//
// ncu on the earlier versions of this kernel (segments in shared memory; then float64 staging): float64 pipe 25 %, issue
// 50-56 %, LSU data pipe 58-67 %, 110-136 instructions per sample of which a quarter were float64 arithmetic -- bound by
// everything around the recurrences.  Hence:
//  * shared memory holds only the pre-processed input as float32 (one coalesced store per sample); the segment mapping
//    reads it twice (the sample and its delayed copy x[n-L], which lies in the same block except for the first L samples)
//    and, after the last biquad step, receives the finished float32 sample for the coalesced store;
//  * INTERIOR blocks (every sample they touch, delayed ones included, lies inside the utterance) carry no bounds checks;
//  * the robot carrier and the final cast are applied in the segment mapping (carrier phase advances by one per sample);
//  * the warp carry of the biquad scan is folded in with a table of Phi^lane instead of a second scan, and segments that
//    lie wholly in the warm-up skip the final biquad run (their outputs are discarded).
constexpr int kRqT = 16, kRqThreads = kEqBlock / kRqT;  // 16-sample segments, 512 threads
constexpr int kRqStride = kRqT + 1;                     // padded segment: conflict-free in both mappings
constexpr int kRqJ = kEqBlock / kRqThreads;             // 16 coalesced rounds per block
constexpr int kRqRow = (kRqThreads / kRqT) * kRqStride; // coalesced mapping: sample tid + 512 j sits kRqRow * j further on
struct EqLanePow {
    const double* tab;  // [32][16]: Phi_T^j, j = 0..31 (row-major 4x4), device memory
};
__device__ __forceinline__ int rq_idx(int i) { return i + (i >> 4); }  // sample i of the block -> padded shared-memory index
// The Phi^lane table and the robot carrier are read in the segment mapping, where neighbouring lanes are 16 entries apart:
// from global memory that is 32 sectors per load instruction (the LSU data pipe was 79 % busy with exactly this), so both
// are copied into padded shared memory once per block.  Carriers longer than kRqMaxPeriod stay in global memory.
constexpr int kRqMaxPeriod = 1024;
constexpr int kRqTabDoubles = 32 * kRqStride;

struct RqShared {
    double red[16];
    double rsum[16];
    St4 wsum[16];
    double carry[2][5];  // chained blocks: reverb state and biquad state at the end of the previous block (double-buffered by block parity)
};
// A CTA walks kRqRun consecutive blocks of one utterance.  Only the first starts from rest behind a warm-up; the others take the exact
// states their predecessor ended in (no warm-up: 8,192 outputs instead of 8,192 - warm, and no FIR carry-in sum for the reverb).
constexpr int kRqRun = 8;

template <typename T, bool INTERIOR, int FINISH>
__device__ __forceinline__ void rveq_block(const T* __restrict__ p, const ReverbArgs& ra, const EqArgs& a, const EqLanePow& lp, const FxPre& pre,
                                           const FxPost& post, double* __restrict__ y, double* smd, RqShared& sh, int b, long long n,
                                           long long n0, bool chained, int parity) {
    T* xs = reinterpret_cast<T*>(smd);          // [512][17] pre-processed input
    float* os = reinterpret_cast<float*>(smd);  // later: finished samples (FINISH != 0)
    double* tabs = reinterpret_cast<double*>(xs + kRqThreads * kRqStride);  // [32][17] Phi^lane
    double* cs = tabs + kRqTabDoubles;                                      // padded robot carrier
    const bool cs_on = post.robot && post.period <= kRqMaxPeriod;
    double* red = sh.red;
    double* rsum = sh.rsum;
    St4* wsum = sh.wsum;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int kEqWarm = chained ? 0 : a.warm;
    const long long base = n0 - kEqWarm;
    const double* carry_in = sh.carry[parity ^ 1];  // written by the previous block of this CTA (behind a block barrier)
    const PreOps po = fx_pre_ops(pre, b, n);
    const int cidx = (tid >> 4) * kRqStride + (tid & 15);
    tabs[(tid >> 4) * kRqStride + (tid & 15)] = lp.tab[tid];
    if (cs_on)
        for (int k = tid; k < post.period; k += kRqThreads) cs[rq_idx(k)] = post.carrier[k];
    // ---- pre-processed input, coalesced: zeros outside the utterance
    {
        const T* pt = p + base + tid;
#pragma unroll
        for (int jb = 0; jb < kRqJ; jb += 8) {
            T xv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long g = base + tid + kRqThreads * (jb + j);
                xv[j] = (INTERIOR || (g >= 0 && g < n)) ? pt[kRqThreads * (jb + j)] : (T)0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xs[cidx + kRqRow * (jb + j)] = sizeof(T) == 4 ? (T)fx_pre((float)xv[j], po) : xv[j];
        }
    }
    // carry-in wet[base-1] = sum_k ir[k] x[base-1-k]: eight taps' loads in flight per thread
    double part = 0.0;
    if (base > 0 && !chained) {
#pragma unroll 1
        for (int k0 = tid; k0 < ra.L; k0 += kRqThreads * 8) {
            T xv[8];
            double iv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + kRqThreads * j;
                const long long g = base - 1 - k;
                const bool ok = k < ra.L && (INTERIOR || g >= 0);
                xv[j] = ok ? p[g] : (T)0;
                iv[j] = ok ? ra.ir[k] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part = fma(iv[j], fx_pre_d(xv[j], po), part);
        }
    }
    part = warp_sum(part);
    if (lane == 0) red[wid] = part;
    __syncthreads();
    double carry0;
    {
        double r8[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) r8[w] = red[2 * w] + red[2 * w + 1];
        carry0 = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
        if (chained) carry0 = carry_in[0];
    }
    // ---- segment mapping: thread = samples [16 tid, 16 tid + 16) of the block, in registers from here on
    const int i0 = kRqT * tid;
    double uu[kRqT];
    {   // reverb input u = c x[n] - c r^L x[n-L]
        const double crl = ra.c * ra.rL;
        if (i0 >= ra.L) {  // the delayed samples are in the block
#pragma unroll
            for (int i = 0; i < kRqT; ++i) uu[i] = fma(-crl, (double)xs[rq_idx(i0 + i - ra.L)], ra.c * (double)xs[tid * kRqStride + i]);
        } else {
#pragma unroll
            for (int i = 0; i < kRqT; ++i) {
                const int il = i0 + i - ra.L;
                double d;
                if (il >= 0) {
                    d = (double)xs[rq_idx(il)];
                } else {
                    const long long g = base + il;
                    d = (INTERIOR || (g >= 0 && g < n)) ? fx_pre_d(p[(INTERIOR || g >= 0) ? g : 0], po) : 0.0;
                }
                uu[i] = fma(-crl, d, ra.c * (double)xs[tid * kRqStride + i]);
            }
        }
    }
    {   // reverb: zero-state end value, warp scan with factor rT^(2^k), carry across warps, real run, dry/wet mix
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < kRqT; ++i) e = fma(ra.r, e, uu[i]);
        double f = ra.rT, v = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = fma(f, up, v);
            f *= f;
        }
        if (lane == 31) rsum[wid] = v;
        __syncthreads();
        double cw = carry0;  // f == rT^32; unrolled with the loads up front: the chain is 15 FMAs, not 15 load-FMA round trips
        {
            double rs[15];
#pragma unroll
            for (int w = 0; w < 15; ++w) rs[w] = rsum[w];
#pragma unroll
            for (int w = 0; w < 15; ++w)
                if (w < wid) cw = fma(f, cw, rs[w]);
        }
        double ex = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) ex = 0.0;
        double pw = 1.0, bb = ra.rT;
        for (int l = lane; l > 0; l >>= 1) {
            if (l & 1) pw *= bb;
            bb *= bb;
        }
        double s = fma(pw, cw, ex);
#pragma unroll
        for (int i = 0; i < kRqT; ++i) {
            s = fma(ra.r, s, uu[i]);
            const T xi = xs[tid * kRqStride + i];
            // (1 - mix) * samples is a float32 product while samples is still float32 (NEP 50 weak scalar)
            const double dry = (sizeof(T) == 4) ? (double)__fmul_rn((float)(1.0 - ra.mix), (float)xi) : (1.0 - ra.mix) * (double)xi;
            uu[i] = fma(ra.mix, s, dry);
            if (!INTERIOR) {  // the reverb tail past the utterance is cut (chain.py: [:len(x)])
                const long long g = base + i0 + i;
                if (g < 0 || g >= n) uu[i] = 0.0;
            }
        }
        if (tid == kRqThreads - 1) sh.carry[parity][0] = s;  // reverb state behind the block's last sample
    }
    // ---- biquads from rest over the block: zero-state end state of the segment (impulse-to-state table), warp scan
    St4 e{{0, 0, 0, 0}};
#pragma unroll
    for (int i = 0; i < kRqT; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) e.z[r] = fma(a.K[r][i], uu[i], e.z[r]);
    }
    St4 v = e;  // v_j = Phi^(2^k) v_{j-2^k} + v_j
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        St4 up;
#pragma unroll
        for (int i = 0; i < 4; ++i) up.z[i] = __shfl_up_sync(0xffffffffu, v.z[i], 1 << k);
        if (lane >= (1 << k)) {
            const St4 t = mat4(a.P[k], up);
#pragma unroll
            for (int i = 0; i < 4; ++i) v.z[i] += t.z[i];
        }
    }
    St4 x0{{0, 0, 0, 0}};  // biquad state entering the block: rest (behind a warm-up), or what the previous block ended in
    if (chained) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x0.z[i] = carry_in[1 + i];
    }
    if (lane == 31) {
        St4 tot = v;
        if (chained && wid == 0) {  // the warp totals are scanned from rest: the entering state rides along as Phi^512 x0 on the first one
            const St4 t = mat4(a.Qp[0], x0);
#pragma unroll
            for (int i = 0; i < 4; ++i) tot.z[i] += t.z[i];
        }
        wsum[wid] = tot;
    }
    __syncthreads();  // also: every thread has read its xs[] values
    if (i0 >= kEqWarm) {  // warm-up segments are done: their outputs are discarded (warp-uniform: the warm-up is a multiple of 512)
        // state entering this warp's first segment: every warp scans the 16 warp totals itself (lane w holds warp w; four
        // rounds with Q^(2^k)) instead of waiting for one thread to chain them; the block starts from rest
        St4 cw;
        {
            St4 t = wsum[lane & 15];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                St4 up;
#pragma unroll
                for (int i = 0; i < 4; ++i) up.z[i] = __shfl_up_sync(0xffffffffu, t.z[i], 1 << k);
                if ((lane & 15) >= (1 << k)) {
                    const St4 m = mat4(a.Qp[k], up);
#pragma unroll
                    for (int i = 0; i < 4; ++i) t.z[i] += m.z[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double c = __shfl_sync(0xffffffffu, t.z[i], (wid + 15) & 15);  // inclusive total of warps 0..wid-1
                cw.z[i] = wid == 0 ? x0.z[i] : c;
            }
        }
        St4 s;            // state entering this segment = zero-carry prefix of the previous lanes + Phi^lane (warp carry)
        {
            double m[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = tabs[lane * kRqStride + i];
            const St4 t = mat4(m, cw);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double pv = __shfl_up_sync(0xffffffffu, v.z[i], 1);
                s.z[i] = t.z[i] + (lane == 0 ? 0.0 : pv);
            }
        }
        const long long g0 = base + i0;  // >= n0 >= 0, and a multiple of 16 (block starts and the warm-up are)
        unsigned ph = post.robot ? (unsigned)((unsigned long long)g0 % (unsigned)post.period) : 0u;
        // period % 16 == 0 (8, 16, 24, 48 kHz): the segment's 16 carrier values are one unbroken padded row
        const bool row16 = cs_on && (post.period & 15) == 0;
        const double* crow = cs + rq_idx((int)ph);
        double* q = y + a.rg.offsets[b] + g0;
#pragma unroll
        for (int i = 0; i < kRqT; ++i) {
            double o = eq_step(a, s, uu[i]);
            if (post.robot) {
                if (row16) {
                    o *= crow[i];
                } else {
                    o *= cs_on ? cs[rq_idx((int)ph)] : post.carrier[ph];
                    ph = ph + 1 == (unsigned)post.period ? 0u : ph + 1;
                }
            }
            if (FINISH == 0) {
                if (INTERIOR || g0 + i < n) q[i] = o;
            } else {
                os[tid * kRqStride + i] = (float)o;
            }
        }
        if (tid == kRqThreads - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sh.carry[parity][1 + i] = s.z[i];
        }
    }
    if (FINISH == 0) return;
    __syncthreads();
    // ---- coalesced store of the finished float32 samples
    const long long off = fx_out_off(post, a.rg, b) + n0 + tid;
    const int jw = kEqWarm / kRqThreads;
    const int jn = INTERIOR ? kRqJ : (int)min((long long)kRqJ, jw + (n - n0 - tid + kRqThreads - 1) / kRqThreads);
    for (int j = jw; j < jn; ++j) {
        const float o = os[cidx + kRqRow * j];
        if (FINISH == 1) reinterpret_cast<float*>(post.out)[off + kRqThreads * (j - jw)] = o;
        else reinterpret_cast<int16_t*>(post.out)[off + kRqThreads * (j - jw)] = (int16_t)quant_pcm16(o);
    }
}

// one block-uniform branch picks the body: blocks in the interior of an utterance run without a single bounds check
template <typename T, int FINISH>
__global__ void __launch_bounds__(kRqThreads, 2) k_fx_reverb_eq(const T* __restrict__ x, ReverbArgs ra, EqArgs a, EqLanePow lp, FxPre pre,
                                                                FxPost post, double* __restrict__ y) {
    extern __shared__ __align__(16) double smd[];
    __shared__ RqShared sh;
    const int b = blockIdx.y;
    const long long n = a.rg.lens[b];
    const int kEqOut = kEqBlock - a.warm;
    long long n0 = (long long)blockIdx.x * (kEqOut + (long long)(kRqRun - 1) * kEqBlock);
    const T* p = x + a.rg.offsets[b];
    for (int r = 0; r < kRqRun; ++r) {
        if (n0 >= n) return;
        if (r) __syncthreads();  // the previous block's staging buffer and carried states
        const bool chained = r > 0;
        const long long base = chained ? n0 : n0 - a.warm;
        if (base - ra.L >= 0 && base + kEqBlock <= n) rveq_block<T, true, FINISH>(p, ra, a, lp, pre, post, y, smd, sh, b, n, n0, chained, r & 1);
        else rveq_block<T, false, FINISH>(p, ra, a, lp, pre, post, y, smd, sh, b, n, n0, chained, r & 1);
        n0 += chained ? kEqBlock : kEqOut;
    }
}

template <typename T>
static int launch_reverb_eq(dim3 grid, cudaStream_t st, const T* x, const ReverbArgs& ra, const EqArgs& e, EqLanePow lp, const FxPre& pre,
                            const FxPost& post, double* dst) {
    const int smem_max = kRqThreads * kRqStride * (int)sizeof(T) + (kRqTabDoubles + kRqMaxPeriod + kRqMaxPeriod / 16 + 1) * (int)sizeof(double);
    const int smem = kRqThreads * kRqStride * (int)sizeof(T) +
                     (kRqTabDoubles + ((post.robot && post.period <= kRqMaxPeriod) ? post.period + post.period / 16 + 1 : 0)) * (int)sizeof(double);
    static PerDeviceOnce once;  // per T
    OSB_CUDA(once.run([&] {
        cudaError_t err = cudaSuccess;
        auto opt = [&](auto kern) { if (err == cudaSuccess) err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max); };
        opt(k_fx_reverb_eq<T, 0>); opt(k_fx_reverb_eq<T, 1>); opt(k_fx_reverb_eq<T, 2>);
        return err;
    }));
    switch (post.finish) {
        case 0: OSB_LAUNCH((k_fx_reverb_eq<T, 0>), grid, kRqThreads, smem, st, x, ra, e, lp, pre, post, dst); break;
        case 1: OSB_LAUNCH((k_fx_reverb_eq<T, 1>), grid, kRqThreads, smem, st, x, ra, e, lp, pre, post, dst); break;
        default: OSB_LAUNCH((k_fx_reverb_eq<T, 2>), grid, kRqThreads, smem, st, x, ra, e, lp, pre, post, dst); break;
    }
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

// ---------------------------------------------------------------- voice blend
// out[b] = sum_k w[b][k] * pack[idx[b][k]]  accumulated in order in f32 (torch: result += w * t)
__global__ void __launch_bounds__(256) k_voice_blend(const float* __restrict__ packs, long long n, const int* __restrict__ idx,
                                                     const float* __restrict__ w, int kmax, float* __restrict__ out) {
    const int b = blockIdx.y;
    const long long nvec = n / 4;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < kmax; ++k) {
            const int id = idx[b * kmax + k];
            if (id < 0) break;
            const float wk = w[b * kmax + k];
            const float4 t = reinterpret_cast<const float4*>(packs + (long long)id * n)[v];
            acc.x = __fadd_rn(acc.x, __fmul_rn(wk, t.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(wk, t.y));
            acc.z = __fadd_rn(acc.z, __fmul_rn(wk, t.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(wk, t.w));
        }
        reinterpret_cast<float4*>(out + (long long)b * n)[v] = acc;
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * 4 + threadIdx.x; i < n; i += blockDim.x) {
            float acc = 0.f;
            for (int k = 0; k < kmax; ++k) {
                const int id = idx[b * kmax + k];
                if (id < 0) break;
                acc = __fadd_rn(acc, __fmul_rn(w[b * kmax + k], packs[(long long)id * n + i]));
            }
            out[(long long)b * n + i] = acc;
        }
}

// ---------------------------------------------------------------- host side
static dim3 ragged_grid(long long max_len, long long batch, int per_thread = 4) {
    long long per = (max_len / per_thread + 255) / 256;
    long long want = ((long long)OSB_NUM_SMS * 8 + batch - 1) / batch;
    if (per > want) per = want;
    if (per < 1) per = 1;
    return dim3((unsigned)per, (unsigned)batch);
}

// simulate the 4-state cascade to get Phi_T (T samples, zero input) -- plain host doubles
static void eq_transition(const EqArgs& a, int T, double* phi /*16*/) {
    for (int col = 0; col < 4; ++col) {
        double z[4] = {0, 0, 0, 0};
        z[col] = 1.0;
        for (int t = 0; t < T; ++t) {
            const double y1 = z[0];
            z[0] = -a.a1[1] * y1 + z[1];
            z[1] = -a.a1[2] * y1;
            const double y2 = a.b2[0] * y1 + z[2];
            z[2] = a.b2[1] * y1 - a.a2[1] * y2 + z[3];
            z[3] = a.b2[2] * y1 - a.a2[2] * y2;
        }
        for (int r = 0; r < 4; ++r) phi[4 * r + col] = z[r];
    }
}

static void matmul4(const double* A, const double* B, double* C) {
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            t[4 * i + j] = s;
        }
    memcpy(C, t, sizeof(t));
}

// scipy.signal.butter(2, 80/nyq, 'high') and iirpeak(3000/nyq, Q=2), restated (bilinear transform forms)
static void podcast_eq_coeffs(double sr, EqArgs& a) {
    const double pi = 3.14159265358979323846;
    {   // 2nd-order Butterworth high-pass: analog prototype s^2 + sqrt(2) s + 1, pre-warped, bilinear (fs = 2)
        const double wn = 80.0 / (sr / 2.0);
        const double warped = 4.0 * std::tan(pi * wn / 2.0);  // 2*fs*tan(pi*wn/fs), fs = 2
        // lp2hp of the prototype poles p = exp(+-j 3pi/4): poles warped/p ; zeros at 0 (double) ; gain 1
        const double pr = warped * std::cos(3.0 * pi / 4.0), pim = warped * std::sin(3.0 * pi / 4.0);  // warped/p = warped*conj(p)/|p|^2
        const double fs2 = 4.0;
        // bilinear: z = (fs2 + s)/(fs2 - s)
        const double dr = fs2 - pr, di = -pim;                 // fs2 - p
        const double nr = fs2 + pr, ni = pim;                  // fs2 + p
        const double den = dr * dr + di * di;
        const double zr = (nr * dr + ni * di) / den, zi = (ni * dr - nr * di) / den;  // pole in z
        // gain: k * prod(fs2 - z)/prod(fs2 - p) with z = 0 (double): fs2^2 / |fs2 - p|^2
        const double k = (fs2 * fs2) / den;
        a.b1[0] = k; a.b1[1] = -2.0 * k; a.b1[2] = k;         // zeros at z = +1 (double)
        a.a1[0] = 1.0; a.a1[1] = -2.0 * zr; a.a1[2] = zr * zr + zi * zi;
    }
    {   // iirpeak(w0, Q): scipy _design_notch_peak_filter
        const double w0n = 3000.0 / (sr / 2.0);
        const double bw = (w0n / 2.0) * pi;        // bw = w0/Q with Q = 2, then *pi
        const double w0 = w0n * pi;
        const double gb = 1.0 / std::sqrt(2.0);
        const double beta = (gb / std::sqrt(1.0 - gb * gb)) * std::tan(bw / 2.0);
        const double gain = 1.0 / (1.0 + beta);
        a.b2[0] = (1.0 - gain); a.b2[1] = 0.0; a.b2[2] = -(1.0 - gain);
        a.a2[0] = 1.0; a.a2[1] = -2.0 * gain * std::cos(w0); a.a2[2] = (2.0 * gain - 1.0);
    }
}

struct FxState {
    cudaStream_t st;
    Ragged rg;
    long long batch, max_len, total;
    void* cur;       // current data
    bool f64;
    float* f32_tmp;  // scratch buffers (total elements each)
    double* d_a;
    double* d_b;
    FxPre pre{nullptr, 0.0, nullptr, nullptr};  // deferred float32 normalise(s), consumed by the next recurrence kernel
    FxPost post{0, 0, 1, nullptr, nullptr, nullptr};  // robot / final cast riding on the next recurrence kernel
    const long long* out_offsets = nullptr;  // final placement of utterance b (null: its input offset)
    bool finished = false;            // the final cast already happened inside a kernel
};

// one period of the robot carrier, cached per (device, sample rate)
static int robot_carrier(int sample_rate, const double** d_tab, int* period) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, double*> cache;
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    int a = 100, bb = sample_rate;
    while (bb) { const int t = a % bb; a = bb; bb = t; }
    *period = sample_rate / a;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, sample_rate});
    if (it == cache.end()) {
        std::vector<double> h(*period);
        const double w = 2 * 3.141592653589793 * 100;  // 2 * np.pi * 100
        for (int k = 0; k < *period; ++k) h[k] = std::sin(w * ((double)k / (double)sample_rate));
        double* d = nullptr;
        OSB_CUDA(cudaMalloc(&d, h.size() * sizeof(double)));
        OSB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        it = cache.emplace(std::make_pair(dev, sample_rate), d).first;
    }
    *d_tab = it->second;
    return OSB_OK;
}

// dynamic shared memory opt-in of the recurrence kernels, once per process (the effect entry points are called from
// several threads at once)
static cudaError_t fx_smem_attrs(int smem) {
    static PerDeviceOnce once;
    return once.run([&] {
        cudaError_t err = cudaFuncSetAttribute(k_fx_reverb<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(k_fx_reverb<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(k_fx_eq<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(k_fx_eq<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        return err;
    });
}

// exponential impulse response exp(-linspace(0, 6, L)) / sum, cached on the device per (device, L): no upload and no
// host synchronisation on the effect path after the first use
static int reverb_ir(int L, const double** d_ir, double* c0) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::pair<double*, double>> cache;
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, L});
    if (it == cache.end()) {
        std::vector<double> ir(L);
        double sum = 0.0;
        for (int k = 0; k < L; ++k) {
            // np.linspace(0, 6, L)[k] = k * (6/(L-1)), last = 6
            const double t = (L == 1) ? 0.0 : (k == L - 1 ? 6.0 : k * (6.0 / (L - 1)));
            ir[k] = std::exp(-t);
        }
        // ir /= ir.sum()  (numpy pairwise sum; the Kahan sum below agrees to the last ulp or two of ~L terms)
        double c = 0.0;
        for (int k = 0; k < L; ++k) { const double yk = ir[k] - c, t = sum + yk; c = (t - sum) - yk; sum = t; }
        for (int k = 0; k < L; ++k) ir[k] /= sum;
        double* d = nullptr;
        OSB_CUDA(cudaMalloc(&d, sizeof(double) * L));
        OSB_CUDA(cudaMemcpy(d, ir.data(), sizeof(double) * L, cudaMemcpyHostToDevice));
        it = cache.emplace(std::make_pair(dev, L), std::make_pair(d, ir[0])).first;
    }
    *d_ir = it->second.first;
    *c0 = it->second.second;
    return OSB_OK;
}

static void fx_after_recurrence(FxState& s, double* dst) {
    if (s.post.finish) s.finished = true;
    else { s.cur = dst; s.f64 = true; }
    s.pre = FxPre{nullptr, 0.0, nullptr, nullptr};
    s.post = FxPost{0, 0, 1, nullptr, nullptr, nullptr};
}

// trim / peak-normalise left to the first kernel of the chain (osb_tts_post_fx_dev)
struct TtsPlan {
    const long long* offs2;        // start of the kept span of utterance b in the untouched input
    const long long* lens2;        // its length
    const float* pscale;           // peak gain
    const int* pon;                // ... applied (with clip) where non-zero
    const long long* out_offsets;  // where the result of utterance b goes (= the utterance's offset in the untouched input)
    const long long* lens;         // untrimmed lengths
    const TtsStats* stats;         // with the utterances' sums of squares (k_tts_stats<true>)
};

// OSB_FX_UNFUSED=1 keeps reverb and podcast_eq in separate kernels (tests compare the two paths)
static bool fx_unfused() {
    const char* e = getenv("OSB_FX_UNFUSED");
    return e && e[0] == '1';
}

// effective chain: unknown types and zero pitch shifts are no-ops (chain.py:18-31, :46-47); returns the count or -1
static int fx_effective(const int* fx_types, const double* fx_p0, int n_fx, int (&idx)[64]) {
    int m = 0;
    for (int i = 0; i < n_fx; ++i) {
        const int t = fx_types[i];
        const bool live = t == OSB_FX_NORMALIZE || t == OSB_FX_REVERB || t == OSB_FX_PODCAST_EQ || t == OSB_FX_ROBOT ||
                          (t == OSB_FX_PITCH && fx_p0[i] != 0.0);
        if (!live) continue;
        if (m >= 64) return -1;
        idx[m++] = i;
    }
    return m;
}

static int fx_normalize(FxState& s, double target_lufs, Scratch& scr, bool defer, const double* have_sumsq = nullptr,
                        const TtsPlan* plan = nullptr) {
    double* sumsq;
    if (have_sumsq) {
        sumsq = const_cast<double*>(have_sumsq);  // the producer of s.cur already summed its squares (float32 data)
    } else {
        OSB_CUDA(scr.alloc(&sumsq, (size_t)s.batch));
        OSB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * s.batch, s.st));
    }
    const dim3 g = ragged_grid(s.max_len, s.batch);
    const double target_rms = std::pow(10.0, target_lufs / 20.0);
    if (!s.f64) {
        if (!have_sumsq) {
            if (plan && plan->stats) OSB_LAUNCH(k_fx_sumsq_from_stats, (unsigned)s.batch, 256, 0, s.st, (const float*)s.cur, plan->out_offsets, plan->lens, plan->offs2,
                                                plan->lens2, plan->stats, plan->pscale, sumsq);
            else if (plan) OSB_LAUNCH(k_fx_sumsq_post, g, 256, 0, s.st, (const float*)s.cur, s.rg, plan->pscale, plan->pon, sumsq);
            else OSB_LAUNCH(k_fx_sumsq<float>, g, 256, 0, s.st, (const float*)s.cur, s.rg, sumsq);
            OSB_CHECK_LAUNCH();
        }
        if (defer) {  // the next effect is a recurrence kernel: it applies the gain while loading
            s.pre.sumsq = sumsq;
            s.pre.target_rms = target_rms;
            return OSB_OK;
        }
        OSB_LAUNCH(k_fx_scale<float>, g, 256, 0, s.st, (const float*)s.cur, s.rg, sumsq, target_rms, s.f32_tmp);
        OSB_CHECK_LAUNCH();
        s.cur = s.f32_tmp;
    } else {
        double* dst = (s.cur == s.d_a) ? s.d_b : s.d_a;
        OSB_LAUNCH(k_fx_sumsq<double>, g, 256, 0, s.st, (const double*)s.cur, s.rg, sumsq);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_fx_scale<double>, g, 256, 0, s.st, (const double*)s.cur, s.rg, sumsq, target_rms, dst);
        OSB_CHECK_LAUNCH();
        s.cur = dst;
    }
    return OSB_OK;
}

static int eq_prepare(int sample_rate, EqArgs& a, double eps = 1e-17, int align = 32, int T = 32);
static int eq_lane_pow(int sample_rate, int T, const EqArgs& a, const double** d_tab);

// with_eq: the next effect is podcast_eq and runs in the same kernel (the float64 samples between them stay on chip)
static int fx_reverb(FxState& s, int sample_rate, int room_ms, double mix, Scratch& scr, bool with_eq = false) {
    int L = (int)((long long)sample_rate * room_ms / 1000);
    if (L < 1) L = 1;
    const double* d_ir;
    double ir0;
    int rc = reverb_ir(L, &d_ir, &ir0);
    if (rc) return rc;
    ReverbArgs a;
    a.rg = s.rg; a.ir = d_ir; a.L = L; a.mix = mix;
    a.r = (L == 1) ? 0.0 : std::exp(-6.0 / (L - 1));
    a.c = ir0;
    a.rL = (L == 1) ? 0.0 : std::exp(-6.0 * L / (L - 1));
    a.rT = std::pow(a.r, (double)kRvT);
    double* dst = (s.cur == s.d_a) ? s.d_b : s.d_a;
    const int smem = 256 * kSegStride * (int)sizeof(double);
    OSB_CUDA(fx_smem_attrs(smem));
    if (with_eq) {
        // The chain's result is cast to float32 (chain.py:32; 6e-8 relative) and the tolerance is 1e-4 of the peak: a start-up transient
        // below 1e-7 of the state is invisible, and the warm-up is recomputed by every block -- 1,536 instead of 2,048 samples at 24 kHz
        // (1e-11) is 8 % more outputs per block.  Multiple of 512 samples: thread t then owns outputs t, t + 512, ...
        EqArgs e;
        if ((rc = eq_prepare(sample_rate, e, 1e-7, kRqThreads, kRqT))) return rc;
        e.rg = s.rg;
        EqLanePow lp;
        if ((rc = eq_lane_pow(sample_rate, kRqT, e, &lp.tab))) return rc;
        a.rT = std::pow(a.r, (double)kRqT);
        const long long run_out = (kEqBlock - e.warm) + (long long)(kRqRun - 1) * kEqBlock;  // outputs of one CTA: kRqRun chained blocks
        const dim3 ge((unsigned)((s.max_len + run_out - 1) / run_out), (unsigned)s.batch);
        if (!s.f64) rc = launch_reverb_eq<float>(ge, s.st, (const float*)s.cur, a, e, lp, s.pre, s.post, dst);
        else rc = launch_reverb_eq<double>(ge, s.st, (const double*)s.cur, a, e, lp, s.pre, s.post, dst);
        if (rc) return rc;
        fx_after_recurrence(s, dst);
        return OSB_OK;
    }
    const dim3 g((unsigned)((s.max_len + kRvBlock - 1) / kRvBlock), (unsigned)s.batch);
    if (!s.f64) OSB_LAUNCH(k_fx_reverb<float>, g, 256, smem, s.st, (const float*)s.cur, a, s.pre, s.post, dst);
    else OSB_LAUNCH(k_fx_reverb<double>, g, 256, smem, s.st, (const double*)s.cur, a, s.pre, s.post, dst);
    OSB_CHECK_LAUNCH();
    fx_after_recurrence(s, dst);
    return OSB_OK;
}

// Phi_T^j for j = 0..31 (T-sample segments; a.P[0] = Phi_T), cached on the device per (device, sample rate, T)
static int eq_lane_pow(int sample_rate, int T, const EqArgs& a, const double** d_tab) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, double*> cache;
    int dev = 0;
    OSB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev * 64 + T, sample_rate});
    if (it == cache.end()) {
        std::vector<double> h(32 * 16, 0.0);
        for (int i = 0; i < 4; ++i) h[4 * i + i] = 1.0;
        for (int j = 1; j < 32; ++j) matmul4(a.P[0], &h[16 * (j - 1)], &h[16 * j]);
        double* d = nullptr;
        OSB_CUDA(cudaMalloc(&d, h.size() * sizeof(double)));
        OSB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        it = cache.emplace(std::make_pair(dev * 64 + T, sample_rate), d).first;
    }
    *d_tab = it->second;
    return OSB_OK;
}

// T: samples per thread segment (K, P, Q are built for it; T <= 32)
static int eq_prepare(int sample_rate, EqArgs& a, double eps, int align, int T) {
    podcast_eq_coeffs((double)sample_rate, a);
    // warm-up must outlast the slowest pole: |p|^warm < eps
    const double rad = std::sqrt(std::fmax(a.a1[2], a.a2[2]));
    const double need = rad < 1.0 ? std::log(eps) / std::log(rad) : 1e30;
    if (!(need <= kEqWarmMax)) {
        set_error("unsupported: podcast_eq at %d Hz needs a warm-up longer than %d samples", sample_rate, kEqWarmMax);
        return OSB_ERR_UNSUPPORTED;
    }
    a.warm = ((int)std::ceil(need) + align - 1) / align * align;
    eq_transition(a, T, a.P[0]);
    for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 32; ++i) a.K[r][i] = 0.0;
    for (int i = 0; i < T; ++i) {  // impulse at sample i of a segment, then zeros to its end
        double z[4] = {0, 0, 0, 0};
        for (int t = i; t < T; ++t) {
            const double x = t == i ? 1.0 : 0.0;
            const double y1 = a.b1[0] * x + z[0];
            z[0] = a.b1[1] * x - a.a1[1] * y1 + z[1];
            z[1] = a.b1[2] * x - a.a1[2] * y1;
            const double y2 = a.b2[0] * y1 + z[2];
            z[2] = a.b2[1] * y1 - a.a2[1] * y2 + z[3];
            z[3] = a.b2[2] * y1 - a.a2[2] * y2;
        }
        for (int r = 0; r < 4; ++r) a.K[r][i] = z[r];
    }
    for (int k = 1; k < 5; ++k) matmul4(a.P[k - 1], a.P[k - 1], a.P[k]);
    matmul4(a.P[4], a.P[4], a.Q);
    memcpy(a.Qp[0], a.Q, sizeof(a.Q));
    for (int k = 1; k < 4; ++k) matmul4(a.Qp[k - 1], a.Qp[k - 1], a.Qp[k]);
    return OSB_OK;
}

static int fx_eq(FxState& s, int sample_rate) {
    EqArgs a;
    int rc = eq_prepare(sample_rate, a);
    if (rc) return rc;
    a.rg = s.rg;
    const int kEqOut = kEqBlock - a.warm;
    double* dst = (s.cur == s.d_a) ? s.d_b : s.d_a;
    const dim3 g((unsigned)((s.max_len + kEqOut - 1) / kEqOut), (unsigned)s.batch);
    const int smem = 256 * kSegStride * (int)sizeof(double);
    OSB_CUDA(fx_smem_attrs(smem));
    if (!s.f64) OSB_LAUNCH(k_fx_eq<float>, g, 256, smem, s.st, (const float*)s.cur, a, s.pre, s.post, dst);
    else OSB_LAUNCH(k_fx_eq<double>, g, 256, smem, s.st, (const double*)s.cur, a, s.pre, s.post, dst);
    OSB_CHECK_LAUNCH();
    fx_after_recurrence(s, dst);
    return OSB_OK;
}

static int fx_robot(FxState& s, int sample_rate) {
    double* dst = (s.cur == s.d_a) ? s.d_b : s.d_a;
    const dim3 g = ragged_grid(s.max_len, s.batch);
    if (!s.f64) OSB_LAUNCH(k_fx_robot<float>, g, 256, 0, s.st, (const float*)s.cur, s.rg, (double)sample_rate, dst);
    else OSB_LAUNCH(k_fx_robot<double>, g, 256, 0, s.st, (const double*)s.cur, s.rg, (double)sample_rate, dst);
    OSB_CHECK_LAUNCH();
    s.cur = dst;
    s.f64 = true;
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

extern "C" {

static int tts_post_impl(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len, int trim,
                         int normalize, float threshold, float peak, float* d_out, int64_t* d_out_lens, double* d_sumsq, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(batch >= 0 && max_len >= 0, "bad sizes");
    if (batch == 0) return OSB_OK;
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    OSB_REQUIRE(d_in && d_offsets && d_lens && d_out && d_out_lens, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    TtsStats* stats;
    OSB_CUDA(scr.alloc(&stats, (size_t)batch));
    Ragged rg{(const long long*)d_offsets, (const long long*)d_lens};
    OSB_LAUNCH(k_tts_stats_init, (unsigned)((batch + 255) / 256), 256, 0, st, stats, (int)batch);
    OSB_CHECK_LAUNCH();
    const dim3 g = ragged_grid(max_len > 0 ? max_len : 1, batch);
    OSB_LAUNCH(k_tts_stats<false>, g, 256, 0, st, d_in, rg, threshold, stats);
    OSB_CHECK_LAUNCH();
    OSB_LAUNCH(k_tts_apply, g, 256, 0, st, d_in, rg, stats, trim, normalize, peak, d_out, (long long*)d_out_lens, d_sumsq);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_tts_post_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len, int trim,
                     int normalize, float threshold, float peak, float* d_out, int64_t* d_out_lens, void* stream) {
    return tts_post_impl(d_in, d_offsets, d_lens, batch, max_len, trim, normalize, threshold, peak, d_out, d_out_lens, nullptr, stream);
}

// d_in_sumsq (optional): per-utterance sum of squares of d_in, used by a normalise that is the first effective effect
static int fx_chain_impl(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len,
                         int64_t total, int sample_rate, const int* fx_types, const double* fx_p0, const double* fx_p1, int n_fx,
                         void* d_out, int out_pcm16, const double* d_in_sumsq, void* stream, const TtsPlan* plan = nullptr) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(batch >= 0 && max_len >= 0 && max_len < (1ll << 31) && total >= 0 && n_fx >= 0 && sample_rate > 0, "bad sizes");
    if (batch == 0 || total == 0) return OSB_OK;
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    OSB_REQUIRE(d_in && d_offsets && d_lens && d_out && (n_fx == 0 || (fx_types && fx_p0 && fx_p1)), "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    FxState s;
    s.st = st; s.rg = Ragged{(const long long*)d_offsets, (const long long*)d_lens};
    s.batch = batch; s.max_len = max_len > 0 ? max_len : 1; s.total = total;
    s.cur = (void*)d_in; s.f64 = false;
    if (plan) {  // the chain reads the untouched input: kept spans, peak gain applied by its first kernel
        s.rg = Ragged{plan->offs2, plan->lens2};
        s.pre.peak_scale = plan->pscale;
        s.pre.peak_on = plan->pon;
        s.out_offsets = plan->out_offsets;
    }
    OSB_CUDA(scr.alloc(&s.f32_tmp, (size_t)total));
    OSB_CUDA(scr.alloc(&s.d_a, (size_t)total));
    OSB_CUDA(scr.alloc(&s.d_b, (size_t)total));
    int idx[64];
    const int m = fx_effective(fx_types, fx_p0, n_fx, idx);
    OSB_REQUIRE(m >= 0, "more than 64 effects");
    auto is_rec = [&](int k) { return k < m && (fx_types[idx[k]] == OSB_FX_REVERB || fx_types[idx[k]] == OSB_FX_PODCAST_EQ); };
    for (int k = 0; k < m; ++k) {
        const int i = idx[k];
        int with_eq = 0;  // reverb directly followed by podcast_eq: one kernel
        if (is_rec(k)) {  // let a following robot and the final cast ride on this kernel's store
            with_eq = (fx_types[i] == OSB_FX_REVERB && k + 1 < m && fx_types[idx[k + 1]] == OSB_FX_PODCAST_EQ && !fx_unfused()) ? 1 : 0;
            const int kl = k + with_eq;  // last effect inside the kernel
            if (kl + 1 < m && fx_types[idx[kl + 1]] == OSB_FX_ROBOT) {
                if ((rc = robot_carrier(sample_rate, &s.post.carrier, &s.post.period))) return rc;
                s.post.robot = 1;
            }
            if (kl + 1 + s.post.robot == m) { s.post.finish = out_pcm16 ? 2 : 1; s.post.out = d_out; s.post.out_offsets = s.out_offsets; }
        }
        switch (fx_types[i]) {
            case OSB_FX_NORMALIZE: rc = fx_normalize(s, fx_p0[i], scr, !s.f64 && is_rec(k + 1), k == 0 ? d_in_sumsq : nullptr, k == 0 ? plan : nullptr); break;
            case OSB_FX_REVERB: { const int rob = s.post.robot; rc = fx_reverb(s, sample_rate, (int)fx_p0[i], fx_p1[i], scr, with_eq != 0); k += rob + with_eq; break; }
            case OSB_FX_PODCAST_EQ: { const int rob = s.post.robot; rc = fx_eq(s, sample_rate); k += rob; break; }
            case OSB_FX_ROBOT: rc = fx_robot(s, sample_rate); break;
            case OSB_FX_PITCH:
                // x.astype(float32) -> float32 result (chain.py:48); f32_tmp may be the input: every group is read before it is written
                rc = launch_pitch_shift(s.cur, s.f64, s.rg.offsets, s.rg.lens, s.batch, s.max_len, sample_rate, fx_p0[i], s.f32_tmp, st);
                s.cur = s.f32_tmp;
                s.f64 = false;
                break;
            default: rc = OSB_OK; break;
        }
        if (rc) return rc;
    }
    if (s.finished) return OSB_OK;
    const dim3 g = ragged_grid(s.max_len, batch);
    if (s.f64) {
        if (out_pcm16) OSB_LAUNCH((k_fx_finish<double, true>), g, 256, 0, st, (const double*)s.cur, s.rg, d_out, s.out_offsets);
        else OSB_LAUNCH((k_fx_finish<double, false>), g, 256, 0, st, (const double*)s.cur, s.rg, d_out, s.out_offsets);
    } else {
        if (out_pcm16) OSB_LAUNCH((k_fx_finish<float, true>), g, 256, 0, st, (const float*)s.cur, s.rg, d_out, s.out_offsets);
        else OSB_LAUNCH((k_fx_finish<float, false>), g, 256, 0, st, (const float*)s.cur, s.rg, d_out, s.out_offsets);
    }
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_fx_chain_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len,
                     int64_t total, int sample_rate, const int* fx_types, const double* fx_p0, const double* fx_p1, int n_fx,
                     void* d_out, int out_pcm16, void* stream) {
    return fx_chain_impl(d_in, d_offsets, d_lens, batch, max_len, total, sample_rate, fx_types, fx_p0, fx_p1, n_fx, d_out, out_pcm16, nullptr, stream);
}

int osb_tts_post_fx_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len, int64_t total,
                        int trim, int normalize, float threshold, float peak, int sample_rate, const int* fx_types, const double* fx_p0,
                        const double* fx_p1, int n_fx, float* d_post, int64_t* d_out_lens, void* d_out, int out_pcm16, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(batch >= 0 && total >= 0 && max_len >= 0 && n_fx >= 0, "bad sizes");
    if (batch == 0 || total == 0) return OSB_OK;
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    OSB_REQUIRE(d_in && d_offsets && d_lens && d_out_lens && d_out && (n_fx == 0 || (fx_types && fx_p0 && fx_p1)), "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    int idx[64];
    const int m = fx_effective(fx_types, fx_p0, n_fx, idx);
    OSB_REQUIRE(m >= 0, "more than 64 effects");
    auto is_rec = [&](int k) { return k < m && (fx_types[idx[k]] == OSB_FX_REVERB || fx_types[idx[k]] == OSB_FX_PODCAST_EQ); };
    if (is_rec(0) || (m >= 2 && fx_types[idx[0]] == OSB_FX_NORMALIZE && is_rec(1))) {
        // The chain starts ([normalise ->] reverb / podcast_eq) with a kernel that can trim, peak-normalise and clip while
        // it loads: only the statistics pass touches the input before it; the post-processed utterances are never written.
        TtsStats* stats;
        long long* offs2;
        float* pscale;
        int* pon;
        OSB_CUDA(scr.alloc(&stats, (size_t)batch));
        OSB_CUDA(scr.alloc(&offs2, (size_t)batch));
        OSB_CUDA(scr.alloc(&pscale, (size_t)batch));
        OSB_CUDA(scr.alloc(&pon, (size_t)batch));
        Ragged rg{(const long long*)d_offsets, (const long long*)d_lens};
        OSB_LAUNCH(k_tts_stats_init, (unsigned)((batch + 255) / 256), 256, 0, st, stats, (int)batch);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_tts_stats<true>, ragged_grid(max_len > 0 ? max_len : 1, batch), 256, 0, st, d_in, rg, threshold, stats);
        OSB_CHECK_LAUNCH();
        OSB_LAUNCH(k_tts_plan, (unsigned)((batch + 255) / 256), 256, 0, st, rg, stats, (int)batch, trim, normalize, peak, offs2,
                   (long long*)d_out_lens, pscale, pon);
        OSB_CHECK_LAUNCH();
        const TtsPlan plan{offs2, (const long long*)d_out_lens, pscale, pon, (const long long*)d_offsets, (const long long*)d_lens, stats};
        return fx_chain_impl(d_in, d_offsets, d_out_lens, batch, max_len, total, sample_rate, fx_types, fx_p0, fx_p1, n_fx, d_out, out_pcm16,
                             nullptr, stream, &plan);
    }
    // any other chain head: materialise the post-processed utterances (into d_post, or scratch), leaving each one's sum of squares
    double* sumsq;
    OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
    OSB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * batch, st));
    if (!d_post) OSB_CUDA(scr.alloc(&d_post, (size_t)total));
    if ((rc = tts_post_impl(d_in, d_offsets, d_lens, batch, max_len, trim, normalize, threshold, peak, d_post, d_out_lens, sumsq, stream))) return rc;
    return fx_chain_impl(d_post, d_offsets, d_out_lens, batch, max_len, total, sample_rate, fx_types, fx_p0, fx_p1, n_fx, d_out, out_pcm16, sumsq, stream);
}

int osb_voice_blend_dev(const float* d_packs, int64_t pack_elems, const int32_t* d_idx, const float* d_weights, int kmax, int64_t batch,
                        float* d_out, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(pack_elems >= 0 && kmax >= 1 && batch >= 0, "bad sizes");
    if (batch == 0 || pack_elems == 0) return OSB_OK;
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    OSB_REQUIRE(d_packs && d_idx && d_weights && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_packs & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && pack_elems % 4 == 0, "packs must be 16-byte aligned, length % 4 == 0");
    long long per = (pack_elems / 4 + 255) / 256;
    if (per > 64) per = 64;
    OSB_LAUNCH(k_voice_blend, dim3((unsigned)per, (unsigned)batch), 256, 0, (cudaStream_t)stream, d_packs, (long long)pack_elems, d_idx,
               d_weights, kmax, d_out);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_podcast_eq_coeffs(int sample_rate, double* out12) {
    OSB_REQUIRE(sample_rate > 0 && out12, "bad arguments");
    EqArgs a;
    podcast_eq_coeffs((double)sample_rate, a);
    for (int i = 0; i < 3; ++i) { out12[i] = a.b1[i]; out12[3 + i] = a.a1[i]; out12[6 + i] = a.b2[i]; out12[9 + i] = a.a2[i]; }
    return OSB_OK;
}

// ---------------------------------------------------------------- host-pointer wrappers (one utterance / one blend)
int osb_tts_post_host(const float* in, int64_t n, int trim, int normalize, float threshold, float peak, float* out, int64_t* out_len) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(out_len, "null out_len");
    *out_len = n > 0 ? n : 0;
    if (n <= 0) return OSB_OK;
    void *di, *dout, *dmeta;
    if ((rc = ws.dev_buf(0, (size_t)n * 4, &di)) || (rc = ws.dev_buf(1, (size_t)n * 4, &dout)) || (rc = ws.dev_buf(2, 64, &dmeta))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 4))) return rc;
    long long meta[3] = {0, n, 0};
    OSB_CUDA(cudaMemcpyAsync(dmeta, meta, sizeof(meta), cudaMemcpyHostToDevice, ws.stream));
    long long* dm = (long long*)dmeta;
    if ((rc = osb_tts_post_dev((const float*)di, (const int64_t*)dm, (const int64_t*)(dm + 1), 1, n, trim, normalize, threshold, peak, (float*)dout,
                               (int64_t*)(dm + 2), ws.stream))) return rc;
    long long m = 0;
    OSB_CUDA(cudaMemcpyAsync(&m, dm + 2, sizeof(long long), cudaMemcpyDeviceToHost, ws.stream));
    if ((rc = ws.sync())) return rc;
    *out_len = m;
    return ws.d2h(out, dout, (size_t)m * 4);
}

int osb_fx_chain_host(const float* in, int64_t n, int sample_rate, const int* fx_types, const double* fx_p0, const double* fx_p1, int n_fx,
                      void* out, int out_pcm16) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n <= 0) return OSB_OK;
    void *di, *dout, *dmeta;
    if ((rc = ws.dev_buf(0, (size_t)n * 4, &di)) || (rc = ws.dev_buf(1, (size_t)n * 4, &dout)) || (rc = ws.dev_buf(2, 64, &dmeta))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 4))) return rc;
    long long meta[2] = {0, n};
    OSB_CUDA(cudaMemcpyAsync(dmeta, meta, sizeof(meta), cudaMemcpyHostToDevice, ws.stream));
    long long* dm = (long long*)dmeta;
    if ((rc = osb_fx_chain_dev((const float*)di, (const int64_t*)dm, (const int64_t*)(dm + 1), 1, n, n, sample_rate, fx_types, fx_p0, fx_p1, n_fx,
                               dout, out_pcm16, ws.stream))) return rc;
    return ws.d2h(out, dout, (size_t)n * (out_pcm16 ? 2 : 4));
}

int osb_voice_blend_host(const float* const* packs, const float* weights, int k, int64_t pack_elems, float* out) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(k >= 1 && k <= 64 && packs && weights && out && pack_elems >= 0, "bad arguments");
    if (pack_elems == 0) return OSB_OK;
    const long long padded = (pack_elems + 3) / 4 * 4;
    void *dp, *dout, *dmeta;
    if ((rc = ws.dev_buf(0, (size_t)k * padded * 4, &dp)) || (rc = ws.dev_buf(1, (size_t)padded * 4, &dout)) || (rc = ws.dev_buf(2, 1024, &dmeta))) return rc;
    OSB_CUDA(cudaMemsetAsync(dp, 0, (size_t)k * padded * 4, ws.stream));
    for (int i = 0; i < k; ++i)
        OSB_CUDA(cudaMemcpyAsync((float*)dp + (size_t)i * padded, packs[i], (size_t)pack_elems * 4, cudaMemcpyHostToDevice, ws.stream));
    int idx[64];
    for (int i = 0; i < k; ++i) idx[i] = i;
    OSB_CUDA(cudaMemcpyAsync(dmeta, idx, sizeof(int) * k, cudaMemcpyHostToDevice, ws.stream));
    OSB_CUDA(cudaMemcpyAsync((char*)dmeta + 512, weights, sizeof(float) * k, cudaMemcpyHostToDevice, ws.stream));
    OSB_CUDA(cudaStreamSynchronize(ws.stream));
    if ((rc = osb_voice_blend_dev((const float*)dp, padded, (const int32_t*)dmeta, (const float*)((char*)dmeta + 512), k, 1, (float*)dout, ws.stream))) return rc;
    return ws.d2h(out, dout, (size_t)pack_elems * 4);
}

}  // extern "C"
