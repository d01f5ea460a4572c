// Composed STT paths: preprocess_stt_audio (reference src/audio/preprocessing.py:53-63) and the batch
// STT front-end of BASELINE configs 1 / 4 (preprocess -> WAV -> faster-whisper FeatureExtractor).
#include <cstdlib>

#include "common.cuh"

using namespace osb;

static int stt_frontend(const int16_t* d_pcm, long long n, long long batch, long long stride, int sr, int noise_reduce,
                        int normalize, int n_mels, float* d_mel, cudaStream_t st) {
    int rc;
    Scratch scr(st);
    if (!noise_reduce) {
        if (normalize) {
            // sum(s^2) per clip, then normalise + requantise fused into the log-mel sample staging
            unsigned long long* sumsq;
            OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
            if ((rc = launch_sumsq_pcm16(d_pcm, n, batch, stride, sumsq, st))) return rc;
            return launch_logmel(d_pcm, OSB_FMT_PCM16, n, batch, stride, n_mels, d_mel, sumsq, -18.0f, st);
        }
        // normalize=False still requantises (x/32768*32767, truncated): float32_mono_to_wav_bytes
        int16_t* q;
        OSB_CUDA(scr.alloc(&q, (size_t)(batch * stride)));
        if ((rc = osb_normalize_gain_pcm16_dev(d_pcm, q, n, batch, stride, 0, -18.0f, st))) return rc;
        return launch_logmel(q, OSB_FMT_PCM16, n, batch, stride, n_mels, d_mel, nullptr, -18.0f, st);
    }
    // denoise -> (normalise) -> int16 round trip -> log-mel: the inverse STFT leaves each clip's sum of squares, and the
    // log-mel kernel applies the gain and the requantisation while it stages the float32 samples, so the normalised
    // int16 clip is never written
    float* den;
    double* sumsq = nullptr;
    OSB_CUDA(scr.alloc(&den, (size_t)(batch * stride)));
    if (normalize) {
        OSB_CUDA(scr.alloc(&sumsq, (size_t)batch));
        OSB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * batch, st));
    }
    if ((rc = launch_spectral_gate(d_pcm, OSB_FMT_PCM16, n, batch, stride, sr, den, st, sumsq))) return rc;
    return launch_logmel(den, OSB_FMT_F32, n, batch, stride, n_mels, d_mel, nullptr, -18.0f, st, sumsq, 1);
}

extern "C" {

int osb_stt_frontend_dev(const int16_t* d_pcm, int64_t n, int64_t batch, int64_t stride, int sample_rate, int noise_reduce,
                         int normalize, int n_mels, float* d_mel, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    OSB_REQUIRE(n_mels == 80 || n_mels == 128, "n_mels must be 80 or 128");
    OSB_REQUIRE(n >= 0 && batch >= 0 && stride >= n && sample_rate > 0, "bad sizes");
    OSB_REQUIRE(batch <= 65535, "batch too large (<= 65535)");
    if (batch == 0) return OSB_OK;
    OSB_REQUIRE(n + 160 > 200, "clip too short");
    OSB_REQUIRE(d_pcm && d_mel, "null buffer");
    return stt_frontend(d_pcm, n, batch, stride, sample_rate, noise_reduce, normalize, n_mels, d_mel, (cudaStream_t)stream);
}

// Host-buffer entry: the batch is cut into groups; H2D of group g+1, the kernels of group g and D2H of group g-1
// run concurrently on three streams (two copy engines + SMs), so the step costs ~max(copy, compute), not their sum.
int osb_stt_frontend_host(const int16_t* pcm, int64_t n, int64_t batch, int64_t stride, int sample_rate, int noise_reduce,
                          int normalize, int n_mels, float* mel) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n > 0 && batch > 0 && stride >= n && pcm && mel, "bad arguments");
    OSB_REQUIRE(n_mels == 80 || n_mels == 128, "n_mels must be 80 or 128");
    const size_t per_mel = (size_t)n_mels * osb_logmel_frames(n);
    void *di, *dout;
    if ((rc = ws.dev_buf(0, (size_t)batch * stride * 2, &di)) || (rc = ws.dev_buf(1, (size_t)batch * per_mel * 4, &dout))) return rc;
    static thread_local cudaStream_t s_in = nullptr, s_out = nullptr;
    static thread_local int s_dev = -1;
    if (s_dev != ws.device) {
        if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_out); }
        OSB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        OSB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        s_dev = ws.device;
    }
    // clip groups (OSB_STT_HOST_GROUPS=1..32 to tune): the middle of the pipeline is PCIe-bound (the float32 features
    // leaving), so more groups = shorter fill (first H2D + first kernels) and drain; measured on 256 x 60 s: 8 groups
    // 17.0 ms, 16: 16.2 ms, 24: 16.0 ms per step
    int64_t bounds[33] = {0};
    int groups = batch >= 192 ? 24 : (batch >= 64 ? 8 : (batch >= 16 ? 4 : 1));
    if (const char* e = getenv("OSB_STT_HOST_GROUPS")) {
        const int g = atoi(e);
        if (g >= 1 && g <= 32 && g <= batch) groups = g;
    }
    for (int g = 0; g <= groups; ++g) bounds[g] = batch * g / groups;
    cudaEvent_t ev_in[32], ev_done[32];
    for (int g = 0; g < groups; ++g) {
        OSB_CUDA(cudaEventCreateWithFlags(&ev_in[g], cudaEventDisableTiming));
        OSB_CUDA(cudaEventCreateWithFlags(&ev_done[g], cudaEventDisableTiming));
    }
    rc = OSB_OK;
    for (int g = 0; g < groups && rc == OSB_OK; ++g) {
        const int64_t c0 = bounds[g], c1 = bounds[g + 1], nb = c1 - c0;
        const int16_t* hin = pcm + c0 * stride;
        int16_t* din = (int16_t*)di + c0 * stride;
        float* dmel = (float*)dout + c0 * per_mel;
        const size_t ib = (size_t)((nb - 1) * stride + n) * 2;
        cudaError_t e = cudaMemcpyAsync(din, hin, ib, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaEventRecord(ev_in[g], s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ws.stream, ev_in[g], 0);
        if (e != cudaSuccess) { rc = cuda_fail(e, "h2d group", __FILE__, __LINE__); break; }
        rc = osb_stt_frontend_dev(din, n, nb, stride, sample_rate, noise_reduce, normalize, n_mels, dmel, ws.stream);
        if (rc) break;
        e = cudaEventRecord(ev_done[g], ws.stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_out, ev_done[g], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(mel + c0 * per_mel, dmel, (size_t)nb * per_mel * 4, cudaMemcpyDeviceToHost, s_out);
        if (e != cudaSuccess) { rc = cuda_fail(e, "d2h group", __FILE__, __LINE__); break; }
    }
    cudaError_t e1 = cudaStreamSynchronize(s_in), e2 = cudaStreamSynchronize(ws.stream), e3 = cudaStreamSynchronize(s_out);
    for (int g = 0; g < groups; ++g) { cudaEventDestroy(ev_in[g]); cudaEventDestroy(ev_done[g]); }
    if (rc) return rc;
    OSB_CUDA(e1);
    OSB_CUDA(e2);
    OSB_CUDA(e3);
    return OSB_OK;
}

int osb_preprocess_stt_host(const int16_t* in, int64_t n, int channels, int sample_rate, int noise_reduce, int normalize,
                            float target_dbfs, int16_t* out) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(channels >= 1 && n >= 0 && sample_rate > 0, "bad arguments");
    const long long frames = n / channels;
    if (frames == 0) return OSB_OK;
    OSB_REQUIRE(in && out, "null buffer");
    void *di, *df, *dg, *dq;
    if ((rc = ws.dev_buf(0, (size_t)n * 2, &di)) || (rc = ws.dev_buf(1, (size_t)frames * 4, &df)) ||
        (rc = ws.dev_buf(2, (size_t)frames * 4, &dg)) || (rc = ws.dev_buf(3, (size_t)frames * 2, &dq))) return rc;
    if ((rc = ws.h2d(di, in, (size_t)n * 2))) return rc;
    if ((rc = osb_pcm16_to_f32_dev((const int16_t*)di, (float*)df, (size_t)n, channels, ws.stream))) return rc;
    const float* cur = (const float*)df;
    if (noise_reduce) {
        if ((rc = launch_spectral_gate(df, OSB_FMT_F32, frames, 1, frames, sample_rate, (float*)dg, ws.stream))) return rc;
        cur = (const float*)dg;
    }
    if ((rc = launch_normalize_f32(cur, dq, 1, frames, 1, frames, normalize, target_dbfs, ws.stream))) return rc;
    return ws.d2h(out, dq, (size_t)frames * 2);
}

}  // extern "C"
