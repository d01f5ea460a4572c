// PCM16 <-> float32 edge and RMS gain normalisation for the STT path.
//
// Replaces (reference file:line):
//   wav_bytes_to_float32_mono   src/audio/preprocessing.py:9-20   (int16 -> f32 /32768, channel mean)
//   normalize_gain              src/audio/preprocessing.py:35-42  (RMS -> -18 dBFS gain, clip)
//   float32_mono_to_wav_bytes   src/audio/preprocessing.py:23-32  (clip, *32767, truncate)
//   float32_to_int16            src/tts/pipeline.py:32-37
//
// HBM-bound: one reduction pass (2 B/sample read) + one map pass (2 B read + 2 B write); the
// second read of a clip hits the 126 MB L2 for clips up to tens of MB.  The int16 path reduces
// sum(s^2) in exact 64-bit integers, so the result does not depend on grid shape or atomic order.
#include "common.cuh"

namespace osb {

// ---------------------------------------------------------------- converts
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const int16_t* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t nvec = n / 8;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < nvec; v += nthr) {
        uint4 w = ld_stream_u4(in + v * 8);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = ((float)(int16_t)(ws[k] & 0xFFFF) * 3.0517578125e-05f);
            f[2 * k + 1] = ((float)(int16_t)(ws[k] >> 16) * 3.0517578125e-05f);
        }
        st_stream_u4(out + v * 8, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
        st_stream_u4(out + v * 8 + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
    for (size_t i = nvec * 8 + tid; i < n; i += nthr) out[i] = (float)in[i] / 32768.0f;
}

// interleaved multi-channel int16 -> mono f32: reshape(-1,ch).mean(axis=1) in f32
// (numpy adds the ch values left to right in f32, then divides by ch)
__global__ void __launch_bounds__(256) k_pcm16_to_f32_mc(const int16_t* __restrict__ in, float* __restrict__ out, size_t frames, int ch) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < frames; i += nthr) {
        float acc = 0.0f;
        for (int c = 0; c < ch; ++c) acc = __fadd_rn(acc, ((float)in[i * ch + c] * 3.0517578125e-05f));
        out[i] = __fdiv_rn(acc, (float)ch);
    }
}

__global__ void __launch_bounds__(256) k_f32_to_pcm16(const float* __restrict__ in, int16_t* __restrict__ out, size_t n) {
    size_t nvec = n / 8;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < nvec; v += nthr) {
        uint4 a = ld_stream_u4(in + v * 8), b = ld_stream_u4(in + v * 8 + 4);
        uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q0 = quant_pcm16(__uint_as_float(ws[2 * k])), q1 = quant_pcm16(__uint_as_float(ws[2 * k + 1]));
            o[k] = (uint32_t)(q0 & 0xFFFF) | ((uint32_t)q1 << 16);
        }
        st_stream_u4(out + v * 8, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (size_t i = nvec * 8 + tid; i < n; i += nthr) out[i] = (int16_t)quant_pcm16(in[i]);
}

// ---------------------------------------------------------------- reductions
// grid = (blocks_per_clip, batch)
__global__ void __launch_bounds__(256) k_sumsq_pcm16(const int16_t* __restrict__ in, long long n, long long stride,
                                                     unsigned long long* __restrict__ sumsq) {
    const int16_t* x = in + (long long)blockIdx.y * stride;
    unsigned long long acc = 0;
    const bool aligned = (((uintptr_t)x) & 15) == 0;
    long long nvec = aligned ? n / 8 : 0;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long v = tid; v < nvec; v += nthr) {
        uint4 w = ld_stream_u4(x + v * 8);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int a = (int16_t)(ws[k] & 0xFFFF), b = (int16_t)(ws[k] >> 16);
            acc += (unsigned long long)(unsigned)(a * a) + (unsigned long long)(unsigned)(b * b);
        }
    }
    for (long long i = nvec * 8 + tid; i < n; i += nthr) {
        int a = x[i];
        acc += (unsigned long long)(unsigned)(a * a);
    }
    acc = warp_sum(acc);
    __shared__ unsigned long long sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(&sumsq[blockIdx.y], t);
    }
}

__global__ void __launch_bounds__(256) k_sumsq_f32(const float* __restrict__ in, long long n, long long stride,
                                                   double* __restrict__ sumsq) {
    const float* x = in + (long long)blockIdx.y * stride;
    double acc = 0.0;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n; i += nthr) {
        float v = x[i];
        acc += (double)__fmul_rn(v, v);  // np.square in f32, summed wide
    }
    acc = warp_sum(acc);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(&sumsq[blockIdx.y], t);
    }
}

__global__ void __launch_bounds__(256) k_gain_requant_pcm16(const int16_t* __restrict__ in, int16_t* __restrict__ out,
                                                            long long n, long long stride, const unsigned long long* __restrict__ sumsq,
                                                            int normalize, float target_dbfs) {
    const int16_t* x = in + (long long)blockIdx.y * stride;
    int16_t* y = out + (long long)blockIdx.y * stride;
    bool silent = true;
    float gain = 1.0f;
    if (normalize) gain = gain_from_meansq((double)sumsq[blockIdx.y] / 1073741824.0 / (double)n, target_dbfs, &silent);
    const bool aligned = ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0;
    long long nvec = aligned ? n / 8 : 0;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long v = tid; v < nvec; v += nthr) {
        uint4 w = ld_stream_u4(x + v * 8);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w}, o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a = ((float)(int16_t)(ws[k] & 0xFFFF) * 3.0517578125e-05f), b = ((float)(int16_t)(ws[k] >> 16) * 3.0517578125e-05f);
            int qa = quant_pcm16(apply_gain(a, gain, silent)), qb = quant_pcm16(apply_gain(b, gain, silent));
            o[k] = (uint32_t)(qa & 0xFFFF) | ((uint32_t)qb << 16);
        }
        st_stream_u4(y + v * 8, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (long long i = nvec * 8 + tid; i < n; i += nthr)
        y[i] = (int16_t)quant_pcm16(apply_gain(((float)x[i] * 3.0517578125e-05f), gain, silent));
}

template <bool OUT_PCM16>
__global__ void __launch_bounds__(256) k_gain_f32(const float* __restrict__ in, void* __restrict__ out, long long n, long long stride,
                                                  const double* __restrict__ sumsq, int normalize, float target_dbfs,
                                                  int* __restrict__ silent_flags) {
    const float* x = in + (long long)blockIdx.y * stride;
    bool silent = true;
    float gain = 1.0f;
    if (normalize) {
        // np.mean of the f32 squares: numpy sums pairwise in f32; a wide sum rounded once is
        // within 1 ulp of that (tolerance 1e-4 in the parity tests, DESIGN.md "normalise")
        gain = gain_from_meansq(sumsq[blockIdx.y] / (double)n, target_dbfs, &silent);
    }
    if (silent_flags && blockIdx.x == 0 && threadIdx.x == 0) silent_flags[blockIdx.y] = silent ? 1 : 0;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n; i += nthr) {
        float v = apply_gain(x[i], gain, silent);
        if (OUT_PCM16) reinterpret_cast<int16_t*>(out)[(long long)blockIdx.y * stride + i] = (int16_t)quant_pcm16(v);
        else reinterpret_cast<float*>(out)[(long long)blockIdx.y * stride + i] = v;
    }
}

static inline dim3 clip_grid(long long n, long long batch, int per_thread) {
    long long per_clip = (n / per_thread + 255) / 256;
    long long want = ((long long)OSB_NUM_SMS * 8 + batch - 1) / batch;  // ~8 CTAs per SM in total
    if (per_clip > want) per_clip = want;
    if (per_clip < 1) per_clip = 1;
    return dim3((unsigned)per_clip, (unsigned)batch);
}

int launch_sumsq_pcm16(const int16_t* d_in, long long n, long long batch, long long stride, unsigned long long* d_sumsq,
                       cudaStream_t st) {
    OSB_CUDA(cudaMemsetAsync(d_sumsq, 0, sizeof(unsigned long long) * batch, st));
    OSB_LAUNCH(k_sumsq_pcm16, clip_grid(n, batch, 8), 256, 0, st, d_in, n, stride, d_sumsq);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_pcm16_to_f32_dev(const int16_t* d_in, float* d_out, size_t n, int channels, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(channels >= 1, "channels must be >= 1");
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) {
        OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
        OSB_LAUNCH(k_pcm16_to_f32, grid_for(n / 8 + 1, 256), 256, 0, st, d_in, d_out, n);
    } else {
        size_t frames = n / channels;
        if (frames == 0) return OSB_OK;
        OSB_LAUNCH(k_pcm16_to_f32_mc, grid_for(frames, 256), 256, 0, st, d_in, d_out, frames, channels);
    }
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_f32_to_pcm16_dev(const float* d_in, int16_t* d_out, size_t n, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
    OSB_LAUNCH(k_f32_to_pcm16, grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, n);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_normalize_gain_pcm16_dev(const int16_t* d_in, int16_t* d_out, int64_t n, int64_t batch, int64_t stride,
                                 int normalize, float target_dbfs, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n, "bad sizes");
    if (n == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    unsigned long long* sumsq = nullptr;
    dim3 grid = clip_grid(n, batch, 8);
    if (normalize) {
        OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
        OSB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(unsigned long long) * batch, st));
        OSB_LAUNCH(k_sumsq_pcm16, grid, 256, 0, st, d_in, (long long)n, (long long)stride, sumsq);
        OSB_CHECK_LAUNCH();
    }
    OSB_LAUNCH(k_gain_requant_pcm16, grid, 256, 0, st, d_in, d_out, (long long)n, (long long)stride, sumsq, normalize, target_dbfs);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

static int normalize_gain_f32_impl(const float* d_in, void* d_out, int out_pcm16, int64_t n, int64_t batch, int64_t stride,
                                   int normalize, float target_dbfs, int* d_silent, cudaStream_t st) {
    Scratch scr(st);
    double* sumsq = nullptr;
    dim3 grid = clip_grid(n, batch, 4);
    if (normalize) {
        OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
        OSB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * batch, st));
        OSB_LAUNCH(k_sumsq_f32, grid, 256, 0, st, d_in, (long long)n, (long long)stride, sumsq);
        OSB_CHECK_LAUNCH();
    }
    if (out_pcm16) OSB_LAUNCH(k_gain_f32<true>, grid, 256, 0, st, d_in, d_out, (long long)n, (long long)stride, sumsq, normalize, target_dbfs, d_silent);
    else OSB_LAUNCH(k_gain_f32<false>, grid, 256, 0, st, d_in, d_out, (long long)n, (long long)stride, sumsq, normalize, target_dbfs, d_silent);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_normalize_gain_f32_dev(const float* d_in, void* d_out, int out_pcm16, int64_t n, int64_t batch, int64_t stride,
                               int normalize, float target_dbfs, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n, "bad sizes");
    if (n == 0 || batch == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    return normalize_gain_f32_impl(d_in, d_out, out_pcm16, n, batch, stride, normalize, target_dbfs, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- host-pointer wrappers
int osb_pcm16_to_f32_host(const int16_t* in, float* out, size_t n, int channels) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(channels >= 1, "channels must be >= 1");
    size_t frames = n / channels;
    if (frames == 0) return OSB_OK;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, n * 2, &di)) || (rc = ws.dev_buf(1, frames * 4, &dout))) return rc;
    if ((rc = ws.h2d(di, in, n * 2))) return rc;
    if ((rc = osb_pcm16_to_f32_dev((const int16_t*)di, (float*)dout, n, channels, ws.stream))) return rc;
    return ws.d2h(out, dout, frames * 4);
}

int osb_f32_to_pcm16_host(const float* in, int16_t* out, size_t n) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, n * 4, &di)) || (rc = ws.dev_buf(1, n * 2, &dout))) return rc;
    if ((rc = ws.h2d(di, in, n * 4))) return rc;
    if ((rc = osb_f32_to_pcm16_dev((const float*)di, (int16_t*)dout, n, ws.stream))) return rc;
    return ws.d2h(out, dout, n * 2);
}

int osb_normalize_gain_pcm16_host(const int16_t* in, int16_t* out, int64_t n, int normalize, float target_dbfs) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (n <= 0) return OSB_OK;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, (size_t)n * 2, &di)) || (rc = ws.dev_buf(1, (size_t)n * 2, &dout))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 2))) return rc;
    if ((rc = osb_normalize_gain_pcm16_dev((const int16_t*)di, (int16_t*)dout, n, 1, n, normalize, target_dbfs, ws.stream))) return rc;
    return ws.d2h(out, dout, (size_t)n * 2);
}

int osb_normalize_gain_f32_host(const float* in, void* out, int out_pcm16, int64_t n, int normalize, float target_dbfs,
                                int* unchanged) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    if (unchanged) *unchanged = 0;
    if (n <= 0) return OSB_OK;
    size_t ob = (size_t)n * (out_pcm16 ? 2 : 4);
    void *di, *dout, *dflag;
    if ((rc = ws.dev_buf(0, (size_t)n * 4, &di)) || (rc = ws.dev_buf(1, ob, &dout)) || (rc = ws.dev_buf(2, 16, &dflag))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 4))) return rc;
    if ((rc = normalize_gain_f32_impl((const float*)di, dout, out_pcm16, n, 1, n, normalize, target_dbfs, (int*)dflag, ws.stream))) return rc;
    int flag = 0;
    OSB_CUDA(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, ws.stream));
    if ((rc = ws.d2h(out, dout, ob))) return rc;
    if (unchanged) *unchanged = flag;
    return OSB_OK;
}

}  // extern "C"

namespace osb {
int launch_normalize_f32(const float* d_in, void* d_out, int out_pcm16, long long n, long long batch, long long stride, int normalize,
                         float target_dbfs, cudaStream_t st) {
    return ::normalize_gain_f32_impl(d_in, d_out, out_pcm16, n, batch, stride, normalize, target_dbfs, nullptr, st);
}
}  // namespace osb
