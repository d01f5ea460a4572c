// Composed STT paths: preprocess_stt_audio (reference src/audio/preprocessing.py:53-63) and the batch
// STT front-end of BASELINE configs 1 / 4 (preprocess -> WAV -> faster-whisper FeatureExtractor).
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
    float* den;
    int16_t* q;
    OSB_CUDA(scr.alloc(&den, (size_t)(batch * stride)));
    OSB_CUDA(scr.alloc(&q, (size_t)(batch * stride)));
    if ((rc = launch_spectral_gate(d_pcm, OSB_FMT_PCM16, n, batch, stride, sr, den, st))) return rc;
    if ((rc = launch_normalize_f32(den, q, 1, n, batch, stride, normalize, -18.0f, st))) return rc;
    return launch_logmel(q, OSB_FMT_PCM16, n, batch, stride, n_mels, d_mel, nullptr, -18.0f, st);
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

int osb_stt_frontend_host(const int16_t* pcm, int64_t n, int64_t batch, int64_t stride, int sample_rate, int noise_reduce,
                          int normalize, int n_mels, float* mel) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n > 0 && batch > 0 && stride >= n && pcm && mel, "bad arguments");
    const size_t ib = (size_t)((batch - 1) * stride + n) * 2;
    const size_t ob = (size_t)batch * n_mels * osb_logmel_frames(n) * 4;
    void *di, *dout;
    if ((rc = ws.dev_buf(0, (size_t)batch * stride * 2, &di)) || (rc = ws.dev_buf(1, ob, &dout))) return rc;
    if ((rc = ws.h2d(di, pcm, ib))) return rc;
    if ((rc = osb_stt_frontend_dev((const int16_t*)di, n, batch, stride, sample_rate, noise_reduce, normalize, n_mels, (float*)dout, ws.stream))) return rc;
    return ws.d2h(mel, dout, ob);
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
