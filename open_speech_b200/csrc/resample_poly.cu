// Polyphase PCM16 resampler: scipy.signal.resample_poly(x_f32, up, down, padtype="line"),
// then clip(-32768, 32767) and truncation -- the arithmetic of resample_pcm16
// (reference src/streaming.py:55-91; scipy/signal/_signaltools.py resample_poly and
// scipy/signal/_upfirdn_apply.pyx _apply_impl with MODE_LINE).
//
// Bit-exact by construction: the taps are designed like scipy does (f64 firwin with a
// Kaiser(5.0) window, cast to f32, scaled by `up`, pre-padded, split per phase and flipped)
// and every output is accumulated tap by tap, oldest sample first, with separate f32
// multiply and add (no FMA contraction) -- the order of _apply_impl's inner loop.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace osb {

struct PolyFilter {
    int up = 0, down = 0, per_phase = 0, n_pre_remove = 0, n_taps = 0, n_pre_pad = 0;
    std::vector<float> taps;       // f32(firwin)*up, without padding
    std::vector<float> h_tf;       // [up][per_phase], flipped per phase
    float* d_h_tf = nullptr;
};

static double bessel_i0(double x) {
    // power series  sum ((x/2)^2k / (k!)^2); converges to f64 roundoff for |x| <= 5 in ~25 terms
    double q = x * x * 0.25, term = 1.0, sum = 1.0;
    for (int k = 1; k < 200; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < sum * 1e-18) break;
    }
    return sum;
}

static void design_taps(int up, int down, std::vector<float>& out) {
    const int max_rate = up > down ? up : down;
    const int half_len = 10 * max_rate;
    const int numtaps = 2 * half_len + 1;
    const double cutoff = 1.0 / (double)max_rate, beta = 5.0, alpha = 0.5 * (numtaps - 1);
    const double pi = 3.141592653589793;
    std::vector<double> h(numtaps);
    const double i0b = bessel_i0(beta);
    for (int i = 0; i < numtaps; ++i) {
        double m = (double)i - alpha;
        double xs = cutoff * m;
        double y = pi * (xs == 0.0 ? 1.0e-20 : xs);
        double sinc = std::sin(y) / y;
        double r = ((double)i - alpha) / alpha;
        double w = bessel_i0(beta * std::sqrt(1.0 - r * r)) / i0b;
        h[i] = cutoff * sinc * w;
    }
    // scale so that the DC gain is 1: numpy sums pairwise; for <= 8 k taps the difference to a
    // Kahan sum is far below f32 resolution of the taps
    double s = 0.0, c = 0.0;
    for (int i = 0; i < numtaps; ++i) {
        double yk = h[i] - c, t = s + yk;
        c = (t - s) - yk;
        s = t;
    }
    out.resize(numtaps);
    for (int i = 0; i < numtaps; ++i) {
        float f = (float)(h[i] / s);
        out[i] = f * (float)up;  // h *= up, in f32 like scipy (h already cast to x.dtype)
    }
}

static long long output_len(long long len_h, long long in_len, long long up, long long down) {
    return (((in_len - 1) * up + len_h) - 1) / down + 1;
}

static std::mutex g_mu;
static std::map<long long, PolyFilter> g_filters;  // key: device<<48 | up<<24 | down

static int get_filter(int up, int down, bool need_device, const PolyFilter** out) {
    int dev = 0;
    if (need_device) OSB_CUDA(cudaGetDevice(&dev));
    long long key = ((long long)(need_device ? dev + 1 : 0) << 48) | ((long long)up << 24) | (long long)down;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_filters.find(key);
    if (it == g_filters.end()) {
        PolyFilter f;
        f.up = up; f.down = down;
        design_taps(up, down, f.taps);
        f.n_taps = (int)f.taps.size();
        const int half_len = 10 * (up > down ? up : down);
        f.n_pre_pad = down - half_len % down;
        f.n_pre_remove = (half_len + f.n_pre_pad) / down;
        // h = [zeros(n_pre_pad), taps]; _pad_h: pad to a multiple of up, per phase, flipped.
        // (the n_post_pad loop of resample_poly only appends zeros: per_phase is sized per call below)
        std::vector<float> h((size_t)f.n_pre_pad + f.taps.size(), 0.0f);
        for (size_t i = 0; i < f.taps.size(); ++i) h[f.n_pre_pad + i] = f.taps[i];
        size_t padded = h.size() + ((up - h.size() % up) % up);
        f.per_phase = (int)(padded / up) + 1;  // +1 slot of head-room for n_post_pad zeros
        h.resize((size_t)f.per_phase * up, 0.0f);
        f.h_tf.assign(h.size(), 0.0f);
        for (int p = 0; p < up; ++p)
            for (int k = 0; k < f.per_phase; ++k) f.h_tf[(size_t)p * f.per_phase + k] = h[(size_t)(f.per_phase - 1 - k) * up + p];
        if (need_device) {
            OSB_CUDA(cudaMalloc(&f.d_h_tf, f.h_tf.size() * sizeof(float)));
            OSB_CUDA(cudaMemcpy(f.d_h_tf, f.h_tf.data(), f.h_tf.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        it = g_filters.emplace(key, std::move(f)).first;
    }
    *out = &it->second;
    return OSB_OK;
}

struct PolyArgs {
    const void* in;
    void* out;
    const float* h_tf;
    long long n_in, n_out, batch, in_stride, out_stride;
    int up, down, per_phase, n_pre_remove;
};

// PCM16: int16 in/out, 'line' edge extension, clip + truncate (resample_pcm16, src/streaming.py:55-91)
// F32  : float32 in/out, zero ('constant') extension, no clip  (MultiTrackComposer._resample, src/composer.py:167-173)
template <bool PCM16>
__global__ void __launch_bounds__(256) k_resample_poly(PolyArgs a) {
    const long long total = a.n_out * a.batch;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long g = tid; g < total; g += nthr) {
        const long long c = g / a.n_out, j = g - c * a.n_out;
        const int16_t* xi16 = reinterpret_cast<const int16_t*>(a.in) + c * a.in_stride;
        const float* xf = reinterpret_cast<const float*>(a.in) + c * a.in_stride;
        const long long yi = j + a.n_pre_remove;
        const long long t = yi * a.down;
        const long long x_idx = t / a.up;
        const int phase = (int)(t - x_idx * a.up);
        const float* h = a.h_tf + (size_t)phase * a.per_phase;
        float x0 = 0.f, xl = 0.f, slope = 0.f;
        if (PCM16) {
            x0 = (float)__ldg(xi16);
            xl = (float)__ldg(xi16 + a.n_in - 1);
            slope = __fdiv_rn(__fsub_rn(xl, x0), (float)(a.n_in - 1));
        }
        float acc = 0.0f;
        long long xi = x_idx - a.per_phase + 1;
        for (int k = 0; k < a.per_phase; ++k, ++xi) {
            float xv;
            if (xi < 0) xv = PCM16 ? __fadd_rn(x0, __fmul_rn((float)xi, slope)) : 0.f;
            else if (xi >= a.n_in) xv = PCM16 ? __fadd_rn(xl, __fmul_rn((float)(xi - a.n_in + 1), slope)) : 0.f;
            else xv = PCM16 ? (float)__ldg(xi16 + xi) : __ldg(xf + xi);
            acc = __fadd_rn(acc, __fmul_rn(xv, __ldg(h + k)));
        }
        if (PCM16) {
            acc = fminf(fmaxf(acc, -32768.0f), 32767.0f);
            reinterpret_cast<int16_t*>(a.out)[c * a.out_stride + j] = (int16_t)__float2int_rz(acc);
        } else {
            reinterpret_cast<float*>(a.out)[c * a.out_stride + j] = acc;
        }
    }
}

// MultiTrackComposer._mix_prepared (src/composer.py:175-189): mixed[start_k + i] += track_k[i] in track order, then clip
__global__ void __launch_bounds__(256) k_mix_tracks(const float* __restrict__ flat, const long long* __restrict__ offs,
                                                    const long long* __restrict__ lens, const long long* __restrict__ starts, int n_tracks,
                                                    long long total, float* __restrict__ out) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < n_tracks; ++k) {
            const long long i = j - starts[k];
            if (i >= 0 && i < lens[k]) acc = __fadd_rn(acc, flat[offs[k] + i]);
        }
        out[j] = fminf(fmaxf(acc, -1.0f), 1.0f);
    }
}

// _resample_to_16k (src/wyoming/tts_handler.py:37-44): np.interp(linspace(0, n-1, m), arange(n), audio) in f64 -> f32
__global__ void __launch_bounds__(256) k_interp_index_f32(const float* __restrict__ x, long long n, long long m, double step, float* __restrict__ out) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
        const double pos = (j == m - 1 && m > 1) ? (double)(n - 1) : __dmul_rn((double)j, step);
        long long i = (long long)pos;
        if (i > n - 1) i = n - 1;
        double y;
        if (i >= n - 1 || (double)i == pos) y = (double)x[i];
        else {
            const double f0 = (double)x[i], f1 = (double)x[i + 1];
            const double slope = __ddiv_rn(__dsub_rn(f1, f0), 1.0);
            y = __dadd_rn(__dmul_rn(slope, __dsub_rn(pos, (double)i)), f0);
        }
        out[j] = (float)y;
    }
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_resample_poly_taps(int up, int down, float* taps_out, int capacity, int* n_taps) {
    OSB_REQUIRE(up >= 1 && down >= 1 && !(up == 1 && down == 1), "up/down must be >= 1 and not both 1");
    OSB_REQUIRE(up < (1 << 20) && down < (1 << 20), "ratio too large");
    const PolyFilter* f;
    int rc = get_filter(up, down, false, &f);
    if (rc) return rc;
    if (n_taps) *n_taps = f->n_taps;
    if (taps_out) {
        OSB_REQUIRE(capacity >= f->n_taps, "taps buffer too small");
        memcpy(taps_out, f->taps.data(), sizeof(float) * f->n_taps);
    }
    return OSB_OK;
}

int osb_resample_poly_dev(const int16_t* d_in, int16_t* d_out, int64_t n_in, int64_t batch, int64_t in_stride,
                          int64_t out_stride, int up, int down, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(up >= 1 && down >= 1 && !(up == 1 && down == 1), "up/down must be >= 1 and not both 1 (divide by the gcd first)");
    OSB_REQUIRE(up < (1 << 20) && down < (1 << 20), "ratio too large");
    OSB_REQUIRE(n_in >= 0 && batch >= 0, "negative size");
    if (n_in == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(n_in >= 2, "n_in must be >= 2 (the single-sample case is handled by the caller, src/streaming.py:69-73)");
    const long long n_out = (n_in * up + down - 1) / down;
    OSB_REQUIRE(d_in && d_out && in_stride >= n_in && out_stride >= n_out, "bad buffers/strides");
    const PolyFilter* f;
    if ((rc = get_filter(up, down, true, &f))) return rc;
    // resample_poly appends zeros until upfirdn yields enough outputs; per_phase has one slot of slack
    long long len_h = (long long)f->n_pre_pad + f->n_taps, n_post = 0;
    while (output_len(len_h + n_post, n_in, up, down) < n_out + f->n_pre_remove) ++n_post;
    long long padded = len_h + n_post;
    padded += (up - padded % up) % up;
    if (padded / up > f->per_phase) {
        set_error("unsupported: filter needs %lld taps per phase (> %d)", padded / up, f->per_phase);
        return OSB_ERR_UNSUPPORTED;
    }
    PolyArgs a;
    a.in = d_in; a.out = d_out; a.h_tf = f->d_h_tf;
    a.n_in = n_in; a.n_out = n_out; a.batch = batch; a.in_stride = in_stride; a.out_stride = out_stride;
    a.up = up; a.down = down; a.per_phase = f->per_phase; a.n_pre_remove = f->n_pre_remove;
    OSB_LAUNCH(k_resample_poly<true>, grid_for((size_t)(n_out * batch), 256), 256, 0, (cudaStream_t)stream, a);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_resample_poly_f32_dev(const float* d_in, float* d_out, int64_t n_in, int64_t batch, int64_t in_stride, int64_t out_stride, int up,
                              int down, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(up >= 1 && down >= 1 && !(up == 1 && down == 1), "up/down must be >= 1 and not both 1 (divide by the gcd first)");
    OSB_REQUIRE(up < (1 << 20) && down < (1 << 20), "ratio too large");
    OSB_REQUIRE(n_in >= 0 && batch >= 0, "negative size");
    if (n_in == 0 || batch == 0) return OSB_OK;
    const long long n_out = (n_in * up + down - 1) / down;
    OSB_REQUIRE(d_in && d_out && in_stride >= n_in && out_stride >= n_out, "bad buffers/strides");
    const PolyFilter* f;
    if ((rc = get_filter(up, down, true, &f))) return rc;
    long long len_h = (long long)f->n_pre_pad + f->n_taps, n_post = 0;
    while (output_len(len_h + n_post, n_in, up, down) < n_out + f->n_pre_remove) ++n_post;
    long long padded = len_h + n_post;
    padded += (up - padded % up) % up;
    if (padded / up > f->per_phase) {
        set_error("unsupported: filter needs %lld taps per phase (> %d)", padded / up, f->per_phase);
        return OSB_ERR_UNSUPPORTED;
    }
    PolyArgs a;
    a.in = d_in; a.out = d_out; a.h_tf = f->d_h_tf;
    a.n_in = n_in; a.n_out = n_out; a.batch = batch; a.in_stride = in_stride; a.out_stride = out_stride;
    a.up = up; a.down = down; a.per_phase = f->per_phase; a.n_pre_remove = f->n_pre_remove;
    OSB_LAUNCH(k_resample_poly<false>, grid_for((size_t)(n_out * batch), 256), 256, 0, (cudaStream_t)stream, a);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_resample_poly_f32_host(const float* in, float* out, int64_t n_in, int up, int down) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n_in <= 0) return OSB_OK;
    OSB_REQUIRE(up >= 1 && down >= 1, "up/down must be >= 1");
    const long long n_out = (n_in * up + down - 1) / down;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, (size_t)n_in * 4, &di)) || (rc = ws.dev_buf(1, (size_t)n_out * 4, &dout))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n_in * 4))) return rc;
    if ((rc = osb_resample_poly_f32_dev((const float*)di, (float*)dout, n_in, 1, n_in, n_out, up, down, ws.stream))) return rc;
    return ws.d2h(out, dout, (size_t)n_out * 4);
}

int osb_mix_tracks_host(const float* flat, const int64_t* offsets, const int64_t* lens, const int64_t* starts, int n_tracks, int64_t flat_len,
                        int64_t total, float* out) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n_tracks >= 0 && total >= 0 && flat_len >= 0, "bad sizes");
    if (total == 0) return OSB_OK;
    OSB_REQUIRE(out && (n_tracks == 0 || (flat && offsets && lens && starts)), "null buffer");
    void *df, *dout, *dm;
    if ((rc = ws.dev_buf(0, (size_t)flat_len * 4 + 16, &df)) || (rc = ws.dev_buf(1, (size_t)total * 4, &dout)) ||
        (rc = ws.dev_buf(2, (size_t)n_tracks * 24 + 64, &dm))) return rc;
    if ((rc = ws.h2d(df, flat, (size_t)flat_len * 4))) return rc;
    long long* m = (long long*)dm;
    OSB_CUDA(cudaMemcpyAsync(m, offsets, (size_t)n_tracks * 8, cudaMemcpyHostToDevice, ws.stream));
    OSB_CUDA(cudaMemcpyAsync(m + n_tracks, lens, (size_t)n_tracks * 8, cudaMemcpyHostToDevice, ws.stream));
    OSB_CUDA(cudaMemcpyAsync(m + 2 * n_tracks, starts, (size_t)n_tracks * 8, cudaMemcpyHostToDevice, ws.stream));
    OSB_CUDA(cudaStreamSynchronize(ws.stream));  // the three small arrays are caller stack/heap memory
    OSB_LAUNCH(k_mix_tracks, grid_for((size_t)total, 256), 256, 0, ws.stream, (const float*)df, m, m + n_tracks, m + 2 * n_tracks, n_tracks,
               (long long)total, (float*)dout);
    OSB_CHECK_LAUNCH();
    return ws.d2h(out, dout, (size_t)total * 4);
}

int osb_interp_index_f32_host(const float* in, int64_t n, float* out, int64_t m) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n >= 0 && m >= 0, "bad sizes");
    if (m == 0) return OSB_OK;
    OSB_REQUIRE(n >= 1 && in && out, "need at least one input sample");
    void *di, *dout;
    if ((rc = ws.dev_buf(0, (size_t)n * 4, &di)) || (rc = ws.dev_buf(1, (size_t)m * 4, &dout))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 4))) return rc;
    const double step = m > 1 ? (double)(n - 1) / (double)(m - 1) : 0.0;  // np.linspace: delta / div
    OSB_LAUNCH(k_interp_index_f32, grid_for((size_t)m, 256), 256, 0, ws.stream, (const float*)di, (long long)n, (long long)m, step, (float*)dout);
    OSB_CHECK_LAUNCH();
    return ws.d2h(out, dout, (size_t)m * 4);
}

int osb_resample_poly_host(const int16_t* in, int16_t* out, int64_t n_in, int64_t batch, int64_t in_stride,
                           int64_t out_stride, int up, int down) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n_in <= 0 || batch <= 0) return OSB_OK;
    OSB_REQUIRE(up >= 1 && down >= 1, "up/down must be >= 1");
    const long long n_out = (n_in * up + down - 1) / down;
    size_t ib = (size_t)((batch - 1) * in_stride + n_in) * 2, ob = (size_t)((batch - 1) * out_stride + n_out) * 2;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, ib, &di)) || (rc = ws.dev_buf(1, ob, &dout))) return rc;
    if ((rc = ws.h2d(di, in, ib))) return rc;
    if ((rc = osb_resample_poly_dev((const int16_t*)di, (int16_t*)dout, n_in, batch, in_stride, out_stride, up, down, ws.stream))) return rc;
    return ws.d2h(out, dout, ob);
}

}  // extern "C"
