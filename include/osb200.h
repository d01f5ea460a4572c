/* libosb200 -- C ABI of the B200-native open-speech audio hot path.
 *
 * The reference (will-assistant/open-speech, pure Python) has NO FFI for this path: its
 * boundary is a set of Python callables.  Each entry point below names
 * the reference callable (file:line under /root/reference) whose arithmetic it replaces;
 * open_speech_b200/*.py keeps the reference's Python signatures and binds these symbols
 * with ctypes (INTEGRATION.md shows the stub a maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every function returns OSB_OK (0) or a negative OSB_ERR_*; osb_last_error() gives a
 *     thread-local message.  No exception crosses the boundary.
 *   - `*_dev` functions take DEVICE pointers and a cudaStream_t (as void*; NULL = legacy
 *     default stream), are asynchronous, and allocate their temporaries stream-ordered.
 *   - `*_host` functions take HOST pointers, stage through a per-thread pinned workspace,
 *     run the same kernels on a per-thread stream and return after the result is in the
 *     caller's buffer.  They are re-entrant and thread-safe (per-thread stream/workspace).
 *   - there is no CPU implementation behind any of these: without a CUDA device every
 *     call returns OSB_ERR_NO_DEVICE.
 */
#ifndef OSB200_H
#define OSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSB_VERSION 100 /* 0.1.0 */

#define OSB_OK 0
#define OSB_ERR_INVALID_ARG -1 /* -> ValueError   */
#define OSB_ERR_CUDA -2        /* -> RuntimeError */
#define OSB_ERR_NO_DEVICE -3   /* -> RuntimeError */
#define OSB_ERR_BUFFER -4      /* -> BufferError  */
#define OSB_ERR_UNSUPPORTED -5 /* -> RuntimeError */

/* sample formats on the wire (src/realtime/audio_buffer.py:47-58) */
#define OSB_FMT_PCM16 0
#define OSB_FMT_ULAW 1
#define OSB_FMT_ALAW 2
#define OSB_FMT_F32 3

/* ---------------------------------------------------------------- runtime */
int osb_version(void);
int osb_device_count(void);
int osb_init(int device); /* selects the device for the calling thread; uploads constant tables */
const char* osb_last_error(void);
/* counts kernel launches issued by this library in this process (bench.py "gpu_launches") */
uint64_t osb_launch_count(void);
/* optional per-kernel timing: CUDA event pairs around every launch, on the launching stream.
 * osb_profile_report synchronises the device and writes {"kernel": {"ms": total, "launches": n}, ...}. */
int osb_profile_enable(int on);
int osb_profile_report(char* buf, size_t capacity);
/* test aid: fills the shared memory of every SM with NaN bit patterns and synchronises the device, so that a kernel
 * relying on stale shared memory (instead of what it wrote itself) shows up as a mismatch in the next call */
int osb_debug_poison_smem(void);

/* ---------------------------------------------------------------- G.711 + linear resample
 * replaces audioop.ulaw2lin/alaw2lin/lin2ulaw/lin2alaw and _resample_linear as called from
 * decode_audio_to_pcm16 / encode_pcm16_to_format (src/realtime/audio_buffer.py:20-81).
 * Bit-exact: integer codec; np.interp arithmetic reproduced in IEEE f64 without contraction. */
int osb_g711_decode_dev(const uint8_t* d_in, int16_t* d_out, size_t n, int law, void* stream);
int osb_g711_encode_dev(const int16_t* d_in, uint8_t* d_out, size_t n, int law, void* stream);
/* `batch` independent chunks, each n_in samples -> n_out samples (n_out = int(n_in*to/from),
 * computed by the caller exactly as the reference does).  in_fmt: PCM16|ULAW|ALAW (G.711
 * expand fused in front), out_fmt: PCM16|ULAW|ALAW (G.711 compress fused behind).
 * Strides are in elements of the respective buffer. */
int osb_resample_linear_dev(const void* d_in, int in_fmt, void* d_out, int out_fmt, int64_t n_in, int64_t n_out,
                            int64_t batch, int64_t in_stride, int64_t out_stride, void* stream);
int osb_g711_decode_host(const uint8_t* in, int16_t* out, size_t n, int law);
int osb_g711_encode_host(const int16_t* in, uint8_t* out, size_t n, int law);
int osb_resample_linear_host(const void* in, int in_fmt, void* out, int out_fmt, int64_t n_in, int64_t n_out,
                             int64_t batch, int64_t in_stride, int64_t out_stride);

/* ---------------------------------------------------------------- polyphase resample
 * replaces resample_pcm16 (src/streaming.py:55-91) = scipy.signal.resample_poly(x_f32, up, down,
 * padtype="line") + clip + truncation.  up/down already divided by their gcd.  The FIR is designed
 * inside the library exactly as scipy does (firwin(20*max+1, 1/max, kaiser 5.0) -> f32 -> *up).
 * n_out = ceil(n_in*up/down).  Bit-exact w.r.t. scipy's f32 tap-by-tap accumulation. */
int osb_resample_poly_taps(int up, int down, float* taps_out, int capacity, int* n_taps);
int osb_resample_poly_dev(const int16_t* d_in, int16_t* d_out, int64_t n_in, int64_t batch, int64_t in_stride,
                          int64_t out_stride, int up, int down, void* stream);
int osb_resample_poly_host(const int16_t* in, int16_t* out, int64_t n_in, int64_t batch, int64_t in_stride,
                           int64_t out_stride, int up, int down);

/* ---------------------------------------------------------------- PCM edge + gain normalise
 * replaces wav_bytes_to_float32_mono / float32_mono_to_wav_bytes / normalize_gain
 * (src/audio/preprocessing.py:9-42) and float32_to_int16 (src/tts/pipeline.py:32-37). */
int osb_pcm16_to_f32_dev(const int16_t* d_in, float* d_out, size_t n, int channels, void* stream);
int osb_f32_to_pcm16_dev(const float* d_in, int16_t* d_out, size_t n, void* stream);
/* pcm16 [batch][stride] -> normalize_gain(target_dbfs) -> requantised pcm16 (normalize=0: requantise only) */
int osb_normalize_gain_pcm16_dev(const int16_t* d_in, int16_t* d_out, int64_t n, int64_t batch, int64_t stride,
                                 int normalize, float target_dbfs, void* stream);
/* f32 -> normalize_gain -> f32 (out_pcm16=0) or requantised pcm16 (out_pcm16=1) */
int osb_normalize_gain_f32_dev(const float* d_in, void* d_out, int out_pcm16, int64_t n, int64_t batch, int64_t stride,
                               int normalize, float target_dbfs, void* stream);
int osb_pcm16_to_f32_host(const int16_t* in, float* out, size_t n, int channels);
int osb_f32_to_pcm16_host(const float* in, int16_t* out, size_t n);
int osb_normalize_gain_pcm16_host(const int16_t* in, int16_t* out, int64_t n, int normalize, float target_dbfs);
int osb_normalize_gain_f32_host(const float* in, void* out, int out_pcm16, int64_t n, int normalize, float target_dbfs,
                                int* unchanged);

/* ---------------------------------------------------------------- Silero VAD scoring + segmenting
 * replaces SileroVAD.__call__ / is_speech / get_speech_segments (src/vad/silero.py:63-177): the
 * per-window onnxruntime session.run (:86, :149) and the integer segment state machine (:133-177).
 * Windows are raw 512-sample frames with NO 64-sample context, remainders are dropped, and the
 * LSTM state [2][128] persists across calls -- all as the reference does.
 * weights_host: flat f32 blob in the order of open_speech_b200/vad/silero.py WEIGHT_LAYOUT. */
int osb_vad_create(const float* weights_host, size_t n_floats, void** handle);
int osb_vad_destroy(void* handle);
/* front-end engine: 2 = fused persistent tcgen05 kernel, samples -> gate pre-activations on chip (default); 1 = one tcgen05 split-bf16
 * GEMM per layer; 0 = FP32 FFMA GEMMs (cross-check) */
int osb_vad_set_gemm(void* handle, int use_tcgen05);
/* recurrence engine: 1 = tensor-pipe kernel (W_hh as fp16 mma fragments in registers, h as fp16 hi + lo planes, eight streams per CTA;
 * default); 0 = the FP32 FFMA lock-step kernels (cross-check).  Both replace the LSTM cell inside session.run (:86). */
int osb_vad_set_recurrence(void* handle, int tensor_pipe);
/* batch streams, n samples each (floor(n/512) windows); d_state [batch][2][128] in/out;
 * d_probs [batch][probs_stride] out (one probability per window). */
int osb_vad_score_dev(void* handle, const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, float* d_state,
                      float* d_probs, int64_t probs_stride, void* stream);
/* integer state machine on per-window probabilities: d_segments [batch][max_seg][2] (start_ms,end_ms),
 * d_counts [batch] (may exceed max_seg: then only the first max_seg were stored). */
int osb_vad_segments_dev(const float* d_probs, int64_t probs_stride, int64_t n_win, int64_t batch, int64_t n_samples,
                         float threshold, int min_speech_ms, int silence_ms, int32_t* d_segments, int32_t* d_counts, int max_seg,
                         void* stream);
int osb_vad_score_host(void* handle, const void* audio, int fmt, int64_t n, float* state, float* probs, float* max_prob);
int osb_vad_segments_host(void* handle, const void* audio, int fmt, int64_t n, float* state, float threshold, int min_speech_ms,
                          int silence_ms, int32_t* segments, int max_seg, int* n_seg);

/* ---------------------------------------------------------------- Whisper log-mel front-end
 * replaces faster_whisper.FeatureExtractor.__call__(waveform, padding=160) (third-party; call site
 * src/backends/faster_whisper.py:245, model ctor :40-45).  n_mels = 80 | 128.  Output is
 * f32 [batch][n_mels][osb_logmel_frames(n)], frames contiguous.  fuse_normalize=1 (pcm16 input only)
 * applies normalize_gain + int16 requantisation (src/audio/preprocessing.py:35-42, :23-25) while the
 * samples are staged, i.e. the result equals FeatureExtractor(preprocess_stt_audio(clip)). */
int osb_logmel_frames(int64_t n_samples);
int osb_mel_filters(int n_mels, float* out, size_t capacity); /* f32 [n_mels][201], host buffer */
int osb_logmel_dev(const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, int n_mels, float* d_out,
                   int fuse_normalize, float target_dbfs, void* stream);
int osb_logmel_host(const void* audio, int fmt, int64_t n, int n_mels, float* out, int fuse_normalize, float target_dbfs);

/* ---------------------------------------------------------------- spectral-gating noise reduction
 * replaces reduce_noise (src/audio/preprocessing.py:45-50) = noisereduce.reduce_noise(y, sr) with all
 * defaults (non-stationary gate; STFT 1024/256; 600,000-sample chunks with 30,000 samples of context).
 * Output f32 [batch][stride] (first n of each row). */
int osb_spectral_gate_dev(const void* d_audio, int fmt, int64_t n, int64_t batch, int64_t stride, int sample_rate, float* d_out,
                          void* stream);
int osb_spectral_gate_host(const void* audio, int fmt, float* out, int64_t n, int sample_rate);

/* ---------------------------------------------------------------- composed STT paths
 * osb_preprocess_stt_host = preprocess_stt_audio (src/audio/preprocessing.py:53-63) after header parsing:
 * interleaved int16 (n values, `channels` channels) -> mono f32 -> [reduce_noise] -> [normalize_gain] ->
 * requantised int16 (n / channels values).
 * osb_stt_frontend_*: BASELINE configs 1 / 4 -- the same chain followed by the log-mel front-end that
 * faster-whisper applies to the WAV it is handed: pcm16 [batch][stride] -> f32 [batch][n_mels][frames]. */
int osb_preprocess_stt_host(const int16_t* in, int64_t n, int channels, int sample_rate, int noise_reduce, int normalize,
                            float target_dbfs, int16_t* out);
int osb_stt_frontend_dev(const int16_t* d_pcm, int64_t n, int64_t batch, int64_t stride, int sample_rate, int noise_reduce,
                         int normalize, int n_mels, float* d_mel, void* stream);
int osb_stt_frontend_host(const int16_t* pcm, int64_t n, int64_t batch, int64_t stride, int sample_rate, int noise_reduce,
                          int normalize, int n_mels, float* mel);
/* measurement aid: the H2D / D2H schedule of the *_host batch entries with no kernels in between (the floor of an end-to-end step) */
int osb_copy_floor_host(const void* in, int64_t in_bytes_per_unit, void* out, int64_t out_bytes_per_unit, int64_t batch);

/* osb_stt_full_*: the north_star chain as ONE call, no host hop between stages -- wire audio [batch][n_in] (PCM16 | ULAW | ALAW at
 * from_rate) -> decode_audio_to_pcm16 / resample_pcm16 to 16 kHz -> { SileroVAD.get_speech_segments | preprocess_stt_audio -> log-mel }.
 *   Reference chain: src/realtime/server.py:127-170 -> audio_buffer.py:37-58 -> :111-154 -> src/vad/silero.py:63-177, and
 *   src/main.py:295-296 -> src/audio/preprocessing.py:53-63 -> src/backends/faster_whisper.py:245.  Both branches read the resampled pcm16.
 *   linear_chunk = 0: whole-clip polyphase resample (resample_pcm16, as the Wyoming and streaming doors do);
 *   linear_chunk = k: the realtime door, every k input samples are one append() and are resampled on their own with np.interp arithmetic
 *   (k must divide n_in).  n16 = osb_stt_full_samples(n_in, from_rate, linear_chunk) samples per clip at 16 kHz.
 *   Outputs (each may be NULL to skip): d_pcm16k [batch][n16]; d_probs [batch][n16/512] (fresh LSTM state per clip); d_segments
 *   [batch][max_seg][2] + d_counts [batch]; d_mel [batch][n_mels][osb_logmel_frames(n16)].  vad = NULL skips the VAD branch. */
int64_t osb_stt_full_samples(int64_t n_in, int from_rate, int linear_chunk);
int osb_stt_full_dev(void* vad, const void* d_in, int in_fmt, int from_rate, int64_t n_in, int64_t batch, int64_t in_stride, int linear_chunk,
                     int noise_reduce, int normalize, int n_mels, float vad_threshold, int min_speech_ms, int silence_ms, int16_t* d_pcm16k,
                     float* d_probs, int32_t* d_segments, int32_t* d_counts, int max_seg, float* d_mel, void* stream);
int osb_stt_full_host(void* vad, const void* in, int in_fmt, int from_rate, int64_t n_in, int64_t batch, int linear_chunk, int noise_reduce,
                      int normalize, int n_mels, float vad_threshold, int min_speech_ms, int silence_ms, float* probs, int32_t* segments,
                      int32_t* counts, int max_seg, float* mel);

/* ---------------------------------------------------------------- TTS post-processing, effects, voice blend
 * Ragged batches: utterance b = flat[d_offsets[b] : d_offsets[b] + d_lens[b]] (int64 arrays on the device).
 * osb_tts_post = process_tts_chunks after concatenation: trim_silence (|x| > threshold, first..last) then
 *   normalize_output (peak -> `peak`, clip) (src/audio/postprocessing.py:8-40); output b starts at d_offsets[b],
 *   its new length goes to d_out_lens[b].
 * osb_fx_chain = apply_chain (src/effects/chain.py:15-32): ordered effects, float32 until the first float64
 *   effect, float64 afterwards, cast to float32 at the end (out_pcm16=1 additionally applies float32_to_int16,
 *   src/tts/pipeline.py:32-37).  fx_p0/fx_p1: NORMALIZE target_lufs,- | REVERB room_ms,mix | PITCH semitones,-.
 *   PITCH = _pitch_shift (src/effects/chain.py:44-48 -> librosa.effects.pitch_shift defaults): input cast to float32,
 *   STFT 2048/512 -> phase vocoder at rate 2^(-semitones/12) -> ISTFT -> band-limited resampling back to the input
 *   length; float32 result.  Parity unpinned (librosa/soxr absent): Kaiser-sinc resampler instead of soxr_hq.
 * osb_voice_blend = KokoroBackend._blend_voices (src/tts/backends/kokoro.py:289-308): result += w_k * pack_k in
 *   float32, in component order; d_idx [batch][kmax] (-1 terminates), d_weights [batch][kmax]. */
#define OSB_FX_NORMALIZE 1
#define OSB_FX_REVERB 2
#define OSB_FX_PODCAST_EQ 3
#define OSB_FX_ROBOT 4
#define OSB_FX_PITCH 5
int osb_tts_post_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len, int trim,
                     int normalize, float threshold, float peak, float* d_out, int64_t* d_out_lens, void* stream);
int osb_fx_chain_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len,
                     int64_t total, int sample_rate, const int* fx_types, const double* fx_p0, const double* fx_p1, int n_fx,
                     void* d_out, int out_pcm16, void* stream);
/* process_tts_chunks + apply_chain (+ float32_to_int16) in one call (src/main.py:848-857 runs them back to back).
 * d_out receives the chain's result at d_offsets, d_out_lens the post-trim lengths.  When the chain starts with
 * ([normalize ->] reverb | podcast_eq) its first kernel trims, peak-normalises and clips while loading and the
 * post-processed utterances are never written; otherwise they are materialised in d_post (scratch of `total` floats,
 * may be NULL; contents unspecified on return) and that pass leaves the sum of squares for a leading normalize. */
int osb_tts_post_fx_dev(const float* d_in, const int64_t* d_offsets, const int64_t* d_lens, int64_t batch, int64_t max_len, int64_t total,
                        int trim, int normalize, float threshold, float peak, int sample_rate, const int* fx_types, const double* fx_p0,
                        const double* fx_p1, int n_fx, float* d_post, int64_t* d_out_lens, void* d_out, int out_pcm16, void* stream);
int osb_voice_blend_dev(const float* d_packs, int64_t pack_elems, const int32_t* d_idx, const float* d_weights, int kmax, int64_t batch,
                        float* d_out, void* stream);
int osb_podcast_eq_coeffs(int sample_rate, double* out12); /* b_hp[3], a_hp[3], b_pk[3], a_pk[3] (host) */
int osb_tts_post_host(const float* in, int64_t n, int trim, int normalize, float threshold, float peak, float* out, int64_t* out_len);
int osb_fx_chain_host(const float* in, int64_t n, int sample_rate, const int* fx_types, const double* fx_p0, const double* fx_p1, int n_fx,
                      void* out, int out_pcm16);
int osb_voice_blend_host(const float* const* packs, const float* weights, int k, int64_t pack_elems, float* out);

/* ---------------------------------------------------------------- SURVEY 8(f) "next" rows (callers either side of the path)
 * osb_resample_poly_f32: MultiTrackComposer._resample (src/composer.py:167-173) = resample_poly(f32, up, down) with the
 *   default zero ('constant') edge; same filter design and accumulation order as osb_resample_poly, float32 in/out.
 * osb_mix_tracks: MultiTrackComposer._mix_prepared (src/composer.py:175-189): offset-add in track order, clip to [-1,1].
 * osb_interp_index_f32: wyoming _resample_to_16k (src/wyoming/tts_handler.py:37-44): np.interp on an index grid, f64, -> f32. */
/* osb_vad_extract_speech: wyoming _extract_speech_segments (src/wyoming/stt_handler.py:43-115): [polyphase to 16 kHz] -> VAD score ->
 *   segments -> gather of the speech spans at the ORIGINAL rate, all on the device; only the speech samples come back.
 *   *out_n == 0 means "no usable segment": the caller returns the original audio, as the reference does. */
int osb_vad_extract_speech_host(void* handle, const int16_t* pcm, int64_t n, int rate, float threshold, int min_speech_ms, int silence_ms,
                                int16_t* out, int64_t* out_n, int* n_segments);
int osb_resample_poly_f32_dev(const float* d_in, float* d_out, int64_t n_in, int64_t batch, int64_t in_stride, int64_t out_stride, int up,
                              int down, void* stream);
int osb_resample_poly_f32_host(const float* in, float* out, int64_t n_in, int up, int down);
int osb_mix_tracks_host(const float* flat, const int64_t* offsets, const int64_t* lens, const int64_t* starts, int n_tracks, int64_t flat_len,
                        int64_t total, float* out);
int osb_interp_index_f32_host(const float* in, int64_t n, float* out, int64_t m);

/* Realtime TTS output framing: src/realtime/server.py:249-251 and :268-277.
 * osb_f32_to_pcm16_rt: (x * 32767).clip(-32768, 32767).astype(int16) -- the realtime handler's own quantiser (clip AFTER the
 *   multiply; differs from float32_to_int16 for x <= -1).  Bit-exact.
 * osb_base64_encode: RFC 4648 text of n payload bytes, 4*ceil(n/3) characters with '=' padding, no terminator; the reference's
 *   3000-byte deltas are consecutive 4000-character slices of it.  Bit-exact vs base64.b64encode.
 * osb_realtime_tts_encode_host: float32 24 kHz -> PCM16 -> out_fmt (PCM16: as is; ULAW/ALAW: np.interp to 8 kHz + lin2ulaw/lin2alaw,
 *   n_out = int(n * (8000 / 24000)) computed by the caller as the reference does) -> payload bytes and/or their base64 text
 *   (either pointer may be null). */
int osb_f32_to_pcm16_rt_dev(const float* d_in, int16_t* d_out, size_t n, void* stream);
int osb_base64_encode_dev(const uint8_t* d_in, size_t n, char* d_out, void* stream);
int osb_realtime_tts_encode_host(const float* audio, int64_t n, int out_fmt, int64_t n_out, uint8_t* payload, char* b64);

/* ---------------------------------------------------------------- batched realtime gates (device-resident per-stream state)
 * osb_gate_tick_dev = for each of n_streams concurrent streams, one decode_audio_to_pcm16 (src/realtime/audio_buffer.py:37-58; poly=1:
 *   audioop + resample_pcm16, src/streaming.py:55-91) followed by one InputAudioBuffer.append (src/realtime/audio_buffer.py:111-156):
 *   the chunk is resampled into d_pcm [S][n_out] (n_out computed by the caller like the reference: int(n_in * (16000 / from_rate)) for
 *   the linear path, ceil(n_in*up/down) for the polyphase path), appended to the stream's arena slice (d_arena [S][arena_stride], may be
 *   NULL), scored (max over the chunk's full 512-sample windows, 0.0 when there is none; d_vad_state [S][2][128] carried), and the
 *   integer start / stop machine advances.  gated=0 is `vad=None`: buffer and clock only.  d_prob_override [S] (may be NULL) replaces the
 *   VAD's value for the chunk (scripted tests, external VADs).  Events of the tick: d_events [max_events][3] = (stream, OSB_EVT_*, ms) in
 *   stream order, *d_event_count = how many there were.  OSB_EVT_FRAME_TOO_LARGE / OSB_EVT_BUFFER_FULL are the two BufferError cases
 *   (:118-122) against arena_stride; the stream's state is then left as the reference leaves it.
 *   d_work: osb_gate_work_bytes(n_streams) bytes, zeroed once by the caller, private to the gate.
 *   d_in, d_pcm, d_events and d_event_count may also be PINNED HOST buffers (unified addressing): the kernels then read the wire bytes and
 *   write the tick's pcm16 and event list over PCIe themselves, and a tick needs no copy at all (1024 streams x 20 ms: p50 49 us instead
 *   of 66 us with cudaMemcpyAsync either side); d_state, d_vad_state, d_work and d_arena stay in device memory.
 * osb_gate_clear_dev = InputAudioBuffer.clear() / the clearing half of commit() (:106-109, :158-162) for the listed streams.
 * osb_stream_tick_dev = StreamingSession._process_chunk (src/streaming.py:290-355) for n_streams sessions: polyphase resample of the
 *   client-rate chunk, VAD, and the utterance machine including the state half of _transcribe_utterance / _finalize_utterance
 *   (:357-360, :429-436, :493-498).  d_actions [S]: OSB_ACT_* bits telling the host what the reference would do next for that session. */
typedef struct osb_gate_state {
    int64_t total_samples, silence_samples, buffered_samples;
    int32_t in_speech, speech_start_ms;
} osb_gate_state;
typedef struct osb_stream_state {
    int64_t silence_samples, utterance_bytes;
    int32_t speech_active, reserved;
} osb_stream_state;
#define OSB_EVT_SPEECH_STARTED 1
#define OSB_EVT_SPEECH_STOPPED 2
#define OSB_EVT_FRAME_TOO_LARGE 3
#define OSB_EVT_BUFFER_FULL 4
#define OSB_ACT_SPEECH_START 1     /* send {"type": "vad", "state": "speech_start"} */
#define OSB_ACT_UTTERANCE_RESET 2  /* utterance_audio = bytearray(); agreement.reset() before appending */
#define OSB_ACT_APPEND 4           /* utterance_audio.extend(chunk_16k) */
#define OSB_ACT_TRANSCRIBE 8       /* _transcribe_utterance() runs (>= 3200 bytes) */
#define OSB_ACT_FINALIZE 16        /* _finalize_utterance() transcribes and resets the utterance */
#define OSB_ACT_SPEECH_END 32      /* send {"type": "vad", "state": "speech_end"} */
int64_t osb_gate_work_bytes(int64_t n_streams);
int osb_gate_tick_dev(void* vad, const void* d_in, int in_fmt, int64_t n_in, int from_rate, int poly, int64_t n_streams, int64_t in_stride,
                      int16_t* d_pcm, int64_t n_out, osb_gate_state* d_state, float* d_vad_state, const float* d_prob_override, int gated,
                      int16_t* d_arena, int64_t arena_stride, float threshold, int silence_duration_ms, int32_t* d_work, int32_t* d_events,
                      int32_t* d_event_count, int max_events, void* stream);
int osb_gate_clear_dev(osb_gate_state* d_state, const int32_t* d_stream_ids, int n, void* stream);
int osb_stream_tick_dev(void* vad, const int16_t* d_in, int64_t n_in, int from_rate, int64_t n_streams, int64_t in_stride, int16_t* d_pcm,
                        int64_t n_out, osb_stream_state* d_state, float* d_vad_state, const float* d_prob_override, int vad_enabled,
                        float threshold, int64_t endpointing_samples, int64_t max_utterance_bytes, int32_t* d_actions, void* stream);
/* per-stream host entries behind the drop-in classes (state lives in the caller's object, like the reference's):
 * osb_gate_append_host = one InputAudioBuffer.append(pcm16 @16 kHz); event[0] = OSB_EVT_* or 0, event[1] = its ms.
 * osb_stream_chunk_host = one StreamingSession._process_chunk(client-rate chunk); out16k receives resample_pcm16(chunk) (n_out samples). */
int osb_gate_append_host(void* vad, const int16_t* pcm, int64_t n, osb_gate_state* state, float* vad_state, int gated, float threshold,
                         int silence_duration_ms, int32_t* event);
int osb_stream_chunk_host(void* vad, const int16_t* pcm, int64_t n_in, int from_rate, int16_t* out16k, int64_t n_out, osb_stream_state* state,
                          float* vad_state, int vad_enabled, float threshold, int64_t endpointing_samples, int64_t max_utterance_bytes,
                          int32_t* actions);

#ifdef __cplusplus
}
#endif
#endif /* OSB200_H */
