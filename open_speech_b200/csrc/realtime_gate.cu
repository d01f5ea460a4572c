// Batched realtime gates: the per-stream speech start / stop state machines of the two streaming front doors, kept on
// the device for S concurrent streams and advanced by ONE call per tick.
//
// Replaces (reference file:line):
//   InputAudioBuffer.append / clear / commit            src/realtime/audio_buffer.py:84-166   (OpenAI-realtime WebSocket)
//   decode_audio_to_pcm16 in front of it                src/realtime/audio_buffer.py:37-58, server.py:127-140
//   StreamingSession._process_chunk / _finalize_utterance (state half)   src/streaming.py:290-355, :429-498
//
// A tick = S chunks (one per stream, same length) -> [G.711 expand] -> resample to 16 kHz (np.interp arithmetic of
// _resample_linear, or scipy's polyphase arithmetic of resample_pcm16) -> append to the stream's slice of a device arena
// -> Silero score of the chunk's full 512-sample windows (max, 0.0 when there is none; LSTM state carried per stream)
// -> integer state machine -> compact event list [stream, type, ms], in stream order.  With chunks shorter than one
// window (the reference's 20 ms case) the whole tick is one kernel launch: the last CTA to finish compacts the events.
#include "common.cuh"

namespace osb {

struct GateState {  // == osb_gate_state
    long long total_samples, silence_samples, buffered_samples;
    int in_speech, speech_start_ms;
};
struct StreamState {  // == osb_stream_state
    long long silence_samples, utterance_bytes;
    int speech_active, pad;
};

struct TickArgs {
    const int16_t* pcm;       // [S][n] 16 kHz chunk of every stream (already decoded + resampled)
    long long n, n_streams;
    GateState* st;
    int16_t* arena;           // [S][arena_stride] or null
    long long arena_stride;   // capacity in samples per stream
    const float* probs;       // [S][prob_stride] per-window probabilities of this chunk, or null
    long long prob_stride;
    int n_win;                // windows per chunk in probs (0: the VAD saw no full window -> 0.0)
    int prob_is_chunk;        // probs holds ONE value per stream that already is the chunk's probability (scripted / external VAD)
    int gated;                // 0: vad=None in the reference: buffer + clock only
    float threshold;
    long long silence_ms;
    int2* dense;              // [S] (type, ms) of this tick, type 0 = none
    int* events;              // [max_events][3]
    int* event_count;
    int max_events;
    unsigned int* ticket;     // zero-initialised; the last CTA compacts and resets it
};

constexpr int kGateStreamsPerBlock = 8;

// the reference's append(), after the bytes are in the buffer (audio_buffer.py:124-156)
__device__ __forceinline__ int2 gate_step(GateState& g, long long n, float prob, bool gated, float thr, long long silence_ms) {
    const long long cur_ms = (g.total_samples * 1000) / 16000;
    g.total_samples += n;
    int2 ev = make_int2(0, 0);
    if (!gated || n == 0) return ev;
    if (prob >= thr) {
        g.silence_samples = 0;
        if (!g.in_speech) {
            g.in_speech = 1;
            g.speech_start_ms = (int)cur_ms;
            ev = make_int2(OSB_EVT_SPEECH_STARTED, (int)cur_ms);
        }
    } else if (g.in_speech) {
        g.silence_samples += n;
        if ((g.silence_samples * 1000) / 16000 >= silence_ms) {
            g.in_speech = 0;
            g.silence_samples = 0;
            ev = make_int2(OSB_EVT_SPEECH_STOPPED, (int)cur_ms);
        }
    }
    return ev;
}

// events of one tick in stream order: [stream, type, ms]; run by the last CTA of the tick (or by a 1-CTA launch)
__device__ void compact_events(const TickArgs& a) {
    __shared__ int warp_tot[32];
    __shared__ int base_sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) base_sm = 0;
    __syncthreads();
    for (long long s0 = 0; s0 < a.n_streams; s0 += blockDim.x) {
        const long long s = s0 + tid;
        int2 e = make_int2(0, 0);
        if (s < a.n_streams) e = a.dense[s];
        const unsigned m = __ballot_sync(0xffffffffu, e.x != 0);
        const int before = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base_sm;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (e.x != 0) {
            const int k = off + before;
            if (k < a.max_events) {
                a.events[3 * k] = (int)s;
                a.events[3 * k + 1] = e.x;
                a.events[3 * k + 2] = e.y;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += warp_tot[w];
            base_sm += t;
        }
        __syncthreads();
    }
    if (tid == 0) *a.event_count = base_sm;
}

// One CTA = kGateStreamsPerBlock streams: arena append (all threads), then one thread per stream advances the machine.
__global__ void __launch_bounds__(256) k_gate_tick(TickArgs a) {
    __shared__ long long off_sm[kGateStreamsPerBlock];
    __shared__ int code_sm[kGateStreamsPerBlock];
    __shared__ bool last_sm;
    const int tid = threadIdx.x;
    const long long s0 = (long long)blockIdx.x * kGateStreamsPerBlock;
    if (tid < kGateStreamsPerBlock) {
        const long long s = s0 + tid;
        int code = 0;
        long long off = -1;
        if (s < a.n_streams) {
            GateState g = a.st[s];
            if (a.arena) {
                // BufferError semantics (audio_buffer.py:118-122): a frame larger than the buffer clears it; a frame that does not
                // fit is refused; in both cases nothing else changes
                if (a.n > a.arena_stride) { code = OSB_EVT_FRAME_TOO_LARGE; g.buffered_samples = 0; g.silence_samples = 0; a.st[s] = g; }
                else if (g.buffered_samples + a.n > a.arena_stride) code = OSB_EVT_BUFFER_FULL;
                else off = g.buffered_samples;
            }
        }
        off_sm[tid] = off;
        code_sm[tid] = code;
    }
    __syncthreads();
    if (a.arena) {
        const long long per = a.n;  // samples per stream
        for (long long i = tid; i < per * kGateStreamsPerBlock; i += blockDim.x) {
            const int ls = (int)(i / per);
            const long long k = i - (long long)ls * per, s = s0 + ls;
            if (s < a.n_streams && off_sm[ls] >= 0) a.arena[s * a.arena_stride + off_sm[ls] + k] = a.pcm[s * a.n + k];
        }
    }
    if (tid < kGateStreamsPerBlock) {
        const long long s = s0 + tid;
        if (s < a.n_streams) {
            int2 ev = make_int2(code_sm[tid], 0);
            if (code_sm[tid] == 0) {
                GateState g = a.st[s];
                float prob = 0.0f;  // SileroVAD.__call__: max over the full windows, starting from 0.0 (silero.py:75-91)
                if (a.probs) {
                    if (a.prob_is_chunk) prob = a.probs[s];
                    else
                        for (int w = 0; w < a.n_win; ++w) {
                            const float p = a.probs[s * a.prob_stride + w];
                            if (p > prob) prob = p;
                        }
                }
                if (a.arena) g.buffered_samples += a.n;
                ev = gate_step(g, a.n, prob, a.gated != 0, a.threshold, a.silence_ms);
                a.st[s] = g;
            }
            a.dense[s] = ev;
        }
    }
    // last CTA of the tick compacts the dense events
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        last_sm = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last_sm) {
        __threadfence();
        compact_events(a);
        if (tid == 0) *a.ticket = 0u;
    }
}

__global__ void k_gate_clear(GateState* st, const int* ids, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {  // InputAudioBuffer.clear(): buffer emptied, silence counter reset, clock and in_speech kept (audio_buffer.py:106-109)
        GateState& g = st[ids[i]];
        g.buffered_samples = 0;
        g.silence_samples = 0;
    }
}

// StreamingSession._process_chunk + the state half of _finalize_utterance / _transcribe_utterance (streaming.py:290-355, :357-360, :429-436, :493-498)
__global__ void __launch_bounds__(256) k_stream_gate(const float* __restrict__ probs, long long prob_stride, int n_win, int prob_is_chunk, long long n16,
                                                     long long n_streams, StreamState* st, int vad_enabled, float threshold,
                                                     long long endpointing_samples, long long max_utt_bytes, int* __restrict__ actions) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    StreamState g = st[s];
    const long long chunk_bytes = n16 * 2;
    int act = 0;
    bool work = false, finalize = false;
    if (!vad_enabled) {  // everything is speech (:297-308)
        if (!g.speech_active) { g.speech_active = 1; g.utterance_bytes = 0; act |= OSB_ACT_UTTERANCE_RESET; }
        g.utterance_bytes += chunk_bytes;
        work = true;
        finalize = g.utterance_bytes >= max_utt_bytes;
    } else {
        float prob = 0.0f;
        if (prob_is_chunk) prob = probs[s];
        else
            for (int w = 0; w < n_win; ++w) {
                const float p = probs[s * prob_stride + w];
                if (p > prob) prob = p;
            }
        if (prob >= threshold) {
            g.silence_samples = 0;
            if (!g.speech_active) { g.speech_active = 1; g.utterance_bytes = 0; act |= OSB_ACT_UTTERANCE_RESET | OSB_ACT_SPEECH_START; }
            g.utterance_bytes += chunk_bytes;
            work = true;
            finalize = g.utterance_bytes >= max_utt_bytes;
        } else if (g.speech_active) {
            g.silence_samples += n16;
            g.utterance_bytes += chunk_bytes;
            work = true;
            finalize = g.silence_samples >= endpointing_samples;
        }
    }
    if (work) {
        act |= OSB_ACT_APPEND;
        if (finalize) {
            const bool was_active = g.speech_active != 0;
            g.speech_active = 0;
            g.silence_samples = 0;
            if (g.utterance_bytes < 3200) {  // too short to transcribe: only the speech_end event, the audio stays until the next start
                if (was_active && vad_enabled) act |= OSB_ACT_SPEECH_END;
            } else {
                act |= OSB_ACT_FINALIZE;
                if (vad_enabled) act |= OSB_ACT_SPEECH_END;
                g.utterance_bytes = 0;
            }
        } else if (g.utterance_bytes >= 3200) act |= OSB_ACT_TRANSCRIBE;
    }
    st[s] = g;
    actions[s] = act;
}

// decode + resample the tick into d_pcm [S][n_out] (16 kHz)
static int tick_resample(const void* d_in, int in_fmt, long long n_in, int from_rate, int poly, long long S, long long in_stride, int16_t* d_pcm,
                         long long n_out, Scratch& scr, cudaStream_t st) {
    int rc;
    if (from_rate == 16000) {
        if (in_fmt != OSB_FMT_PCM16) {
            if (in_stride != n_in) { set_error("invalid argument: G.711 input at 16 kHz must be dense"); return OSB_ERR_INVALID_ARG; }
            return osb_g711_decode_dev((const uint8_t*)d_in, d_pcm, (size_t)(S * n_in), in_fmt, st);
        }
        OSB_CUDA(cudaMemcpy2DAsync(d_pcm, (size_t)n_out * 2, d_in, (size_t)in_stride * 2, (size_t)n_in * 2, (size_t)S, cudaMemcpyDeviceToDevice, st));
        return OSB_OK;
    }
    if (!poly) return osb_resample_linear_dev(d_in, in_fmt, d_pcm, OSB_FMT_PCM16, n_in, n_out, S, in_stride, n_out, st);
    int a = 16000, b = from_rate;
    while (b) { const int t = a % b; a = b; b = t; }
    const int up = 16000 / a, down = from_rate / a;
    const int16_t* lin = (const int16_t*)d_in;
    long long lin_stride = in_stride;
    if (in_fmt != OSB_FMT_PCM16) {  // audioop.ulaw2lin first, then resample_pcm16 (the polyphase variant of BASELINE configs[2])
        if (in_stride != n_in) { set_error("invalid argument: G.711 input of the polyphase path must be dense"); return OSB_ERR_INVALID_ARG; }
        int16_t* tmp;
        OSB_CUDA(scr.alloc(&tmp, (size_t)(S * n_in) + 8));
        if ((rc = osb_g711_decode_dev((const uint8_t*)d_in, tmp, (size_t)(S * n_in), in_fmt, st))) return rc;
        lin = tmp;
        lin_stride = n_in;
    }
    return osb_resample_poly_dev(lin, d_pcm, n_in, S, lin_stride, n_out, up, down, st);
}

}  // namespace osb

using namespace osb;

extern "C" {

int osb_gate_tick_dev(void* vad, const void* d_in, int in_fmt, int64_t n_in, int from_rate, int poly, int64_t n_streams, int64_t in_stride,
                      int16_t* d_pcm, int64_t n_out, osb_gate_state* d_state, float* d_vad_state, const float* d_prob_override, int gated,
                      int16_t* d_arena, int64_t arena_stride, float threshold, int silence_duration_ms, int32_t* d_work, int32_t* d_events,
                      int32_t* d_event_count, int max_events, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    static_assert(sizeof(GateState) == sizeof(osb_gate_state), "gate state layout");
    OSB_REQUIRE(in_fmt == OSB_FMT_PCM16 || in_fmt == OSB_FMT_ULAW || in_fmt == OSB_FMT_ALAW, "in_fmt must be PCM16, ULAW or ALAW");
    OSB_REQUIRE(n_in >= 0 && n_out >= 0 && n_streams >= 0 && in_stride >= n_in && from_rate > 0 && max_events >= 0, "bad sizes");
    if (n_streams == 0) return OSB_OK;
    OSB_REQUIRE(d_state && d_work && d_event_count && (d_events || max_events == 0), "null buffer");
    OSB_REQUIRE(n_out == 0 || (d_in && d_pcm), "null audio buffer");
    OSB_REQUIRE(!d_arena || arena_stride > 0, "arena_stride must be positive");
    OSB_REQUIRE(!(gated && !vad && !d_prob_override), "gated tick needs a VAD handle or scripted probabilities");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    if (n_out > 0 && (rc = tick_resample(d_in, in_fmt, n_in, from_rate, poly, n_streams, in_stride, d_pcm, n_out, scr, st))) return rc;
    TickArgs a{};
    a.pcm = d_pcm; a.n = n_out; a.n_streams = n_streams; a.st = reinterpret_cast<GateState*>(d_state);
    a.arena = d_arena; a.arena_stride = arena_stride; a.gated = gated; a.threshold = threshold; a.silence_ms = silence_duration_ms;
    a.dense = reinterpret_cast<int2*>(d_work) + 2;      // d_work: [0] ticket, [4..] dense events (8-byte aligned)
    a.ticket = reinterpret_cast<unsigned int*>(d_work);
    a.events = d_events; a.event_count = d_event_count; a.max_events = max_events;
    const int n_win = (int)(n_out / 512);
    if (gated && d_prob_override) {
        a.probs = d_prob_override; a.prob_is_chunk = 1;
    } else if (gated && n_win > 0) {
        OSB_REQUIRE(d_vad_state, "null VAD state");
        float* probs;
        OSB_CUDA(scr.alloc(&probs, (size_t)(n_streams * n_win)));
        if ((rc = launch_vad_score(vad, d_pcm, OSB_FMT_PCM16, n_out, n_streams, n_out, d_vad_state, probs, n_win, st))) return rc;
        a.probs = probs; a.prob_stride = n_win; a.n_win = n_win;
    }
    const unsigned grid = (unsigned)((n_streams + kGateStreamsPerBlock - 1) / kGateStreamsPerBlock);
    OSB_LAUNCH(k_gate_tick, grid, 256, 0, st, a);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int64_t osb_gate_work_bytes(int64_t n_streams) { return 16 + 8 * (n_streams > 0 ? n_streams : 0); }

int osb_gate_clear_dev(osb_gate_state* d_state, const int32_t* d_stream_ids, int n, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n <= 0) return OSB_OK;
    OSB_REQUIRE(d_state && d_stream_ids, "null buffer");
    OSB_LAUNCH(k_gate_clear, (n + 127) / 128, 128, 0, (cudaStream_t)stream, reinterpret_cast<GateState*>(d_state), d_stream_ids, n);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

int osb_stream_tick_dev(void* vad, const int16_t* d_in, int64_t n_in, int from_rate, int64_t n_streams, int64_t in_stride, int16_t* d_pcm,
                        int64_t n_out, osb_stream_state* d_state, float* d_vad_state, const float* d_prob_override, int vad_enabled,
                        float threshold, int64_t endpointing_samples, int64_t max_utterance_bytes, int32_t* d_actions, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    static_assert(sizeof(StreamState) == sizeof(osb_stream_state), "stream state layout");
    OSB_REQUIRE(n_in >= 0 && n_out >= 0 && n_streams >= 0 && in_stride >= n_in && from_rate > 0, "bad sizes");
    if (n_streams == 0) return OSB_OK;
    OSB_REQUIRE(d_state && d_actions, "null buffer");
    OSB_REQUIRE(n_out == 0 || (d_in && d_pcm), "null audio buffer");
    OSB_REQUIRE(!(vad_enabled && !vad && !d_prob_override), "VAD-enabled tick needs a VAD handle or scripted probabilities");
    cudaStream_t st = (cudaStream_t)stream;
    Scratch scr(st);
    if (n_out > 0 && (rc = tick_resample(d_in, OSB_FMT_PCM16, n_in, from_rate, 1, n_streams, in_stride, d_pcm, n_out, scr, st))) return rc;
    const float* probs = nullptr;
    int n_win = 0, is_chunk = 0;
    if (vad_enabled && d_prob_override) {
        probs = d_prob_override; is_chunk = 1;
    } else if (vad_enabled && n_out / 512 > 0) {
        OSB_REQUIRE(d_vad_state, "null VAD state");
        n_win = (int)(n_out / 512);
        float* p;
        OSB_CUDA(scr.alloc(&p, (size_t)(n_streams * n_win)));
        if ((rc = launch_vad_score(vad, d_pcm, OSB_FMT_PCM16, n_out, n_streams, n_out, d_vad_state, p, n_win, st))) return rc;
        probs = p;
    }
    OSB_LAUNCH(k_stream_gate, (unsigned)((n_streams + 255) / 256), 256, 0, st, probs, (long long)n_win, n_win, is_chunk, (long long)n_out,
               (long long)n_streams, reinterpret_cast<StreamState*>(d_state), vad_enabled, threshold, (long long)endpointing_samples,
               (long long)max_utterance_bytes, d_actions);
    OSB_CHECK_LAUNCH();
    return OSB_OK;
}

// ---------------------------------------------------------------- per-stream host entries (what the drop-in classes call)
// One InputAudioBuffer.append for ONE stream whose state lives in the caller's object, like the reference's: pcm16 @16 kHz in, state and
// LSTM state in/out, at most one event out (event[0] = OSB_EVT_* or 0, event[1] = ms).
int osb_gate_append_host(void* vad, const int16_t* pcm, int64_t n, osb_gate_state* state, float* vad_state, int gated, float threshold,
                         int silence_duration_ms, int32_t* event) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n >= 0 && state && event && (pcm || n == 0), "bad arguments");
    OSB_REQUIRE(!gated || (vad && vad_state), "a gated append needs a VAD session and its state");
    event[0] = event[1] = 0;
    void *d_audio, *d_misc;
    const size_t work = (size_t)osb_gate_work_bytes(1);
    if ((rc = ws.dev_buf(0, (size_t)n * 4 + 64, &d_audio)) || (rc = ws.dev_buf(1, 4096, &d_misc))) return rc;
    int16_t* d_in = (int16_t*)d_audio;
    int16_t* d_pcm = d_in + ((n + 7) / 8) * 8;
    uint8_t* m = (uint8_t*)d_misc;
    osb_gate_state* d_state = (osb_gate_state*)m;            // 32 B
    float* d_vs = (float*)(m + 64);                           // 1 KB
    int32_t* d_work = (int32_t*)(m + 64 + 1024);              // ticket + dense
    int32_t* d_ev = (int32_t*)(m + 64 + 1024 + 64);           // [1][3] + count
    OSB_CUDA(cudaMemsetAsync(d_work, 0, work + 64, ws.stream));
    if (n > 0 && (rc = ws.h2d(d_in, pcm, (size_t)n * 2))) return rc;
    OSB_CUDA(cudaMemcpyAsync(d_state, state, sizeof(osb_gate_state), cudaMemcpyHostToDevice, ws.stream));
    if (gated) OSB_CUDA(cudaMemcpyAsync(d_vs, vad_state, 1024, cudaMemcpyHostToDevice, ws.stream));
    if ((rc = osb_gate_tick_dev(vad, d_in, OSB_FMT_PCM16, n, 16000, 0, 1, n, d_pcm, n, d_state, d_vs, nullptr, gated, nullptr, 0, threshold,
                                silence_duration_ms, d_work, d_ev, d_ev + 3, 1, ws.stream))) return rc;
    int32_t ev[4] = {0, 0, 0, 0};
    OSB_CUDA(cudaMemcpyAsync(state, d_state, sizeof(osb_gate_state), cudaMemcpyDeviceToHost, ws.stream));
    if (gated) OSB_CUDA(cudaMemcpyAsync(vad_state, d_vs, 1024, cudaMemcpyDeviceToHost, ws.stream));
    if ((rc = ws.d2h(ev, d_ev, sizeof(ev)))) return rc;
    if (ev[3] > 0) { event[0] = ev[1]; event[1] = ev[2]; }
    return OSB_OK;
}

// One StreamingSession._process_chunk for ONE session: client-rate pcm16 chunk in, resampled chunk out (out16k, n_out samples), state
// and LSTM state in/out, *actions = OSB_ACT_* bits.
int osb_stream_chunk_host(void* vad, const int16_t* pcm, int64_t n_in, int from_rate, int16_t* out16k, int64_t n_out, osb_stream_state* state,
                          float* vad_state, int vad_enabled, float threshold, int64_t endpointing_samples, int64_t max_utterance_bytes,
                          int32_t* actions) {
    HostWs& ws = host_ws();
    int rc = ws.prepare();
    if (rc) return rc;
    OSB_REQUIRE(n_in >= 0 && n_out >= 0 && state && actions && (pcm || n_in == 0) && (out16k || n_out == 0), "bad arguments");
    OSB_REQUIRE(!vad_enabled || (vad && vad_state), "a VAD-enabled chunk needs a VAD session and its state");
    OSB_REQUIRE(from_rate == 16000 || n_in >= 2, "need at least two samples to resample");
    *actions = 0;
    void *d_audio, *d_misc;
    if ((rc = ws.dev_buf(0, (size_t)(n_in + n_out) * 2 + 64, &d_audio)) || (rc = ws.dev_buf(1, 4096, &d_misc))) return rc;
    int16_t* d_in = (int16_t*)d_audio;
    int16_t* d_pcm = d_in + ((n_in + 7) / 8) * 8;
    uint8_t* m = (uint8_t*)d_misc;
    osb_stream_state* d_state = (osb_stream_state*)m;
    float* d_vs = (float*)(m + 64);
    int32_t* d_act = (int32_t*)(m + 64 + 1024);
    if (n_in > 0 && (rc = ws.h2d(d_in, pcm, (size_t)n_in * 2))) return rc;
    OSB_CUDA(cudaMemcpyAsync(d_state, state, sizeof(osb_stream_state), cudaMemcpyHostToDevice, ws.stream));
    if (vad_enabled) OSB_CUDA(cudaMemcpyAsync(d_vs, vad_state, 1024, cudaMemcpyHostToDevice, ws.stream));
    if ((rc = osb_stream_tick_dev(vad, d_in, n_in, from_rate, 1, n_in, d_pcm, n_out, d_state, d_vs, nullptr, vad_enabled, threshold,
                                  endpointing_samples, max_utterance_bytes, d_act, ws.stream))) return rc;
    OSB_CUDA(cudaMemcpyAsync(state, d_state, sizeof(osb_stream_state), cudaMemcpyDeviceToHost, ws.stream));
    if (vad_enabled) OSB_CUDA(cudaMemcpyAsync(vad_state, d_vs, 1024, cudaMemcpyDeviceToHost, ws.stream));
    OSB_CUDA(cudaMemcpyAsync(actions, d_act, 4, cudaMemcpyDeviceToHost, ws.stream));
    if (n_out > 0) return ws.d2h(out16k, d_pcm, (size_t)n_out * 2);
    return ws.sync();
}

}  // extern "C"
