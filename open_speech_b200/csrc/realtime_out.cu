// Realtime TTS output framing: float32 24 kHz audio -> PCM16 -> session output format -> base64 text of the deltas.
//
// Replaces (reference file:line):
//   (combined * 32767).clip(-32768, 32767).astype(np.int16)     src/realtime/server.py:249   (NOT float32_to_int16: the clip
//                                                                follows the multiply, so -1.0 maps to -32767 and x < -1 to -32768)
//   encode_pcm16_to_format(pcm16, 24000, output_format)         src/realtime/server.py:251 -> audio_buffer.py:65-81
//   base64.b64encode(audio_data[i:i + 3000]) per delta          src/realtime/server.py:268-277
// 3000 is a multiple of 3, so the deltas' base64 strings are consecutive 4000-character slices of the base64 of the whole
// payload: one kernel encodes the payload once, the host slices the text.
//
// Byte movers, HBM-bound: k_rt_quant 4 B in + 2 B out per sample (128-bit loads/stores); k_base64 3 B in + 4 B out per
// group, one 128-bit store per 12 input bytes.  A response is a few hundred KB, so a call is launch-latency bound.
#include "common.cuh"

namespace osb {

__device__ __forceinline__ int quant_pcm16_rt(float x) {
    const float v = fminf(fmaxf(__fmul_rn(x, 32767.0f), -32768.0f), 32767.0f);  // np.float32 * python int -> float32, then clip
    return __float2int_rz(v);                                                    // astype(int16) truncates toward zero
}

__global__ void __launch_bounds__(256) k_rt_quant(const float* __restrict__ in, int16_t* __restrict__ out, size_t n) {
    const size_t nvec = n / 8;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (size_t v = tid; v < nvec; v += nthr) {
        const uint4 a = ld_stream_u4(in + v * 8), b = ld_stream_u4(in + v * 8 + 4);
        const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q0 = quant_pcm16_rt(__uint_as_float(ws[2 * k])), q1 = quant_pcm16_rt(__uint_as_float(ws[2 * k + 1]));
            o[k] = (uint32_t)(q0 & 0xFFFF) | ((uint32_t)q1 << 16);
        }
        st_stream_u4(out + v * 8, make_uint4(o[0], o[1], o[2], o[3]));
    }
    for (size_t i = nvec * 8 + tid; i < n; i += nthr) out[i] = (int16_t)quant_pcm16_rt(in[i]);
}

// RFC 4648 alphabet without a table: A-Z a-z 0-9 + /
__device__ __forceinline__ uint32_t b64_char(uint32_t i) {
    return i + 65u + (i > 25u ? 6u : 0u) - (i > 51u ? 75u : 0u) - (i > 61u ? 15u : 0u) + (i > 62u ? 3u : 0u);
}
// three payload bytes (b0 first) -> four characters packed little-endian (first character in the low byte)
__device__ __forceinline__ uint32_t b64_quad(uint32_t b0, uint32_t b1, uint32_t b2) {
    const uint32_t t = (b0 << 16) | (b1 << 8) | b2;
    return b64_char(t >> 18) | (b64_char((t >> 12) & 63u) << 8) | (b64_char((t >> 6) & 63u) << 16) | (b64_char(t & 63u) << 24);
}

// thread = 12 payload bytes (three aligned words) -> 16 characters (one 128-bit store); the last 1..11 bytes by thread 0 of
// the last block, with '=' padding.  `in` and `out` are 16-byte aligned (allocations of the library).
__global__ void __launch_bounds__(256) k_base64(const uint8_t* __restrict__ in, size_t n, char* __restrict__ out) {
    const size_t ngrp = n / 12;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in);
    for (size_t g = tid; g < ngrp; g += nthr) {
        const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
        uint4 o;
        o.x = b64_quad(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        o.y = b64_quad(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        o.z = b64_quad((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        o.w = b64_quad((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        *reinterpret_cast<uint4*>(out + 16 * g) = o;
    }
    if (tid == 0) {
        size_t i = ngrp * 12, o = ngrp * 16;
        for (; i + 3 <= n; i += 3, o += 4) {
            const uint32_t q = b64_quad(in[i], in[i + 1], in[i + 2]);
            out[o] = (char)(q & 255u); out[o + 1] = (char)((q >> 8) & 255u); out[o + 2] = (char)((q >> 16) & 255u); out[o + 3] = (char)(q >> 24);
        }
        if (i < n) {
            const bool two = i + 2 == n;
            const uint32_t q = b64_quad(in[i], two ? in[i + 1] : 0u, 0u);
            out[o] = (char)(q & 255u); out[o + 1] = (char)((q >> 8) & 255u);
            out[o + 2] = two ? (char)((q >> 16) & 255u) : '=';
            out[o + 3] = '=';
        }
    }
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_f32_to_pcm16_rt_dev(const float* d_in, int16_t* d_out, size_t n, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
    OSB_LAUNCH(k_rt_quant, grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream, d_in, d_out, n);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_base64_encode_dev(const uint8_t* d_in, size_t n, char* d_out, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n == 0) return OSB_OK;
    OSB_REQUIRE(d_in && d_out, "null buffer");
    OSB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "buffers must be 16-byte aligned");
    OSB_LAUNCH(k_base64, grid_for(n / 12 + 1, 256), 256, 0, (cudaStream_t)stream, d_in, n, d_out);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_realtime_tts_encode_host(const float* audio, int64_t n, int out_fmt, int64_t n_out, uint8_t* payload, char* b64) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(out_fmt == OSB_FMT_PCM16 || out_fmt == OSB_FMT_ULAW || out_fmt == OSB_FMT_ALAW, "Unsupported audio format");
    OSB_REQUIRE(n >= 0 && n_out >= 0 && (out_fmt != OSB_FMT_PCM16 || n_out == n), "bad sizes");
    if (n == 0 || n_out == 0) return OSB_OK;
    OSB_REQUIRE(audio && (payload || b64), "null buffer");
    const size_t bytes = out_fmt == OSB_FMT_PCM16 ? (size_t)n * 2 : (size_t)n_out;
    void *df, *dp, *dl, *dt;
    if ((rc = ws.dev_buf(0, (size_t)n * 4, &df)) || (rc = ws.dev_buf(1, (size_t)n * 2, &dp))) return rc;
    if ((rc = ws.h2d(df, audio, (size_t)n * 4))) return rc;
    if ((rc = osb_f32_to_pcm16_rt_dev((const float*)df, (int16_t*)dp, (size_t)n, ws.stream))) return rc;
    const void* dpay = dp;
    if (out_fmt != OSB_FMT_PCM16) {  // 24 kHz PCM16 -> 8 kHz G.711: np.interp + lin2ulaw/lin2alaw fused
        if ((rc = ws.dev_buf(2, bytes, &dl))) return rc;
        if ((rc = osb_resample_linear_dev(dp, OSB_FMT_PCM16, dl, out_fmt, n, n_out, 1, n, n_out, ws.stream))) return rc;
        dpay = dl;
    }
    if (b64) {
        const size_t chars = (bytes + 2) / 3 * 4;
        if ((rc = ws.dev_buf(3, chars, &dt))) return rc;
        if ((rc = osb_base64_encode_dev((const uint8_t*)dpay, bytes, (char*)dt, ws.stream))) return rc;
        if (payload && (rc = ws.d2h(payload, dpay, bytes))) return rc;
        return ws.d2h(b64, dt, chars);
    }
    return ws.d2h(payload, dpay, bytes);
}

}  // extern "C"
