"""Generate tests/golden/*.npz|json by running the REFERENCE itself.  Build-container only.

Run here (``python oracle/make_golden.py``) where /root/reference is mounted; the
reference cannot travel to the GPU box, so the vectors it produces are committed
as small fixtures.  Nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import importlib.machinery
import json
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("OSB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    warnings.simplefilter("ignore", DeprecationWarning)
    import transformers  # noqa: F401  (import first: its librosa probe must not see the stub)

    stub = types.ModuleType("librosa")
    stub.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
    sys.modules.setdefault("librosa", stub)
    sys.path.insert(0, REF)


def main() -> None:
    _import_reference()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import audioop

    from open_speech_b200 import synth
    from src.audio import postprocessing as post
    from src.audio import preprocessing as pre
    from src.effects import chain
    from src.realtime import audio_buffer as ab
    from src.streaming import resample_pcm16
    from src.tts import voices
    from src.tts.pipeline import encode_wav, float32_to_int16
    from src.vad.silero import SileroVAD

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    g: dict[str, np.ndarray] = {}

    # ---- codec + linear resample (src/realtime/audio_buffer.py)
    pcm8k = synth.clip_pcm16(0.1, 8000, seed=11)  # 800 samples
    ul = audioop.lin2ulaw(pcm8k.tobytes(), 2)
    al = audioop.lin2alaw(pcm8k.tobytes(), 2)
    g["codec_ulaw_in"] = np.frombuffer(ul, np.uint8)
    g["codec_alaw_in"] = np.frombuffer(al, np.uint8)
    g["codec_ulaw_all256_8k"] = np.frombuffer(ab.decode_audio_to_pcm16(bytes(range(256)), "g711_ulaw", 8000), np.int16)
    g["codec_alaw_all256_8k"] = np.frombuffer(ab.decode_audio_to_pcm16(bytes(range(256)), "g711_alaw", 8000), np.int16)
    g["codec_ulaw_16k"] = np.frombuffer(ab.decode_audio_to_pcm16(ul, "g711_ulaw", 16000), np.int16)
    g["codec_alaw_16k"] = np.frombuffer(ab.decode_audio_to_pcm16(al, "g711_alaw", 16000), np.int16)
    g["codec_ulaw_chunk160_16k"] = np.frombuffer(ab.decode_audio_to_pcm16(ul[:160], "g711_ulaw", 16000), np.int16)
    pcm24k = synth.clip_pcm16(0.1, 24000, seed=12)
    g["codec_pcm24k_in"] = pcm24k
    g["codec_pcm24k_16k"] = np.frombuffer(ab.decode_audio_to_pcm16(pcm24k.tobytes(), "pcm16", 16000), np.int16)
    pcm16k = synth.clip_pcm16(0.1, 16000, seed=13)
    g["codec_pcm16k_in"] = pcm16k
    g["codec_enc_ulaw"] = np.frombuffer(ab.encode_pcm16_to_format(pcm16k.tobytes(), 16000, "g711_ulaw"), np.uint8)
    g["codec_enc_alaw"] = np.frombuffer(ab.encode_pcm16_to_format(pcm16k.tobytes(), 16000, "g711_alaw"), np.uint8)
    g["codec_enc_pcm16"] = np.frombuffer(ab.encode_pcm16_to_format(pcm16k.tobytes(), 16000, "pcm16"), np.int16)
    all16 = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    g["codec_lin2ulaw_all"] = np.frombuffer(audioop.lin2ulaw(all16.tobytes(), 2), np.uint8)
    g["codec_lin2alaw_all"] = np.frombuffer(audioop.lin2alaw(all16.tobytes(), 2), np.uint8)

    # ---- polyphase resample (src/streaming.py:55-91)
    for fr, n in ((8000, 800), (24000, 2400), (48000, 4800), (44100, 4410), (22050, 2205), (32000, 321)):
        x = synth.clip_pcm16(n / fr, fr, seed=20 + fr // 1000)[:n]
        x[:3] = (32767, -32768, 32767)  # make the clip/edge path do something
        g[f"poly_{fr}_in"] = x
        g[f"poly_{fr}_out"] = np.frombuffer(resample_pcm16(x.tobytes(), fr, 16000), np.int16)
    x = synth.clip_pcm16(0.05, 16000, seed=31)
    g["poly_up_in"] = x
    g["poly_up_48k_out"] = np.frombuffer(resample_pcm16(x.tobytes(), 16000, 48000), np.int16)
    g["poly_single_out"] = np.frombuffer(resample_pcm16(np.array([1000], np.int16).tobytes(), 16000, 32000), np.int16)

    # ---- STT preprocessing (src/audio/preprocessing.py)
    clip = synth.clip_pcm16(2.0, 16000, seed=41)
    wav = pre.float32_mono_to_wav_bytes(clip.astype(np.float32) / 32768.0, 16000)
    g["pre_in_pcm16"] = np.frombuffer(wav[44:], np.int16)
    outw = pre.preprocess_stt_audio(wav, noise_reduce=False, normalize=True)
    g["pre_norm_out_pcm16"] = np.frombuffer(outw[44:], np.int16)
    g["pre_header"] = np.frombuffer(outw[:44], np.uint8)
    outw2 = pre.preprocess_stt_audio(wav, noise_reduce=False, normalize=False)
    g["pre_requant_out_pcm16"] = np.frombuffer(outw2[44:], np.int16)
    a, _ = pre.wav_bytes_to_float32_mono(wav)
    g["pre_gain_f32"] = pre.normalize_gain(a).astype(np.float32)
    quiet = (rng.standard_normal(4000) * 1e-4).astype(np.float32)
    g["pre_quiet_in"] = quiet
    g["pre_quiet_gain"] = pre.normalize_gain(quiet).astype(np.float32)
    # stereo WAV path
    import io, wave
    st = np.stack([clip[:8000], clip[8000:16000]], axis=1)
    bio = io.BytesIO()
    with wave.open(bio, "wb") as wf:
        wf.setnchannels(2); wf.setsampwidth(2); wf.setframerate(16000); wf.writeframes(st.tobytes())
    g["pre_stereo_in"] = st.reshape(-1)
    g["pre_stereo_out_pcm16"] = np.frombuffer(pre.preprocess_stt_audio(bio.getvalue(), noise_reduce=False, normalize=True)[44:], np.int16)

    # ---- TTS post + effects + PCM edge
    utt = synth.tts_utterance(1.5, seed=51)
    g["tts_in"] = utt
    chunks = [utt[:9000], utt[9000:20000], utt[20000:]]
    post_out = list(post.process_tts_chunks(iter(chunks), trim=True, normalize=True))[0]
    g["tts_post_out"] = post_out
    g["tts_trim_only"] = post.trim_silence(utt)
    g["tts_norm_only"] = post.normalize_output(utt)
    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"},
          {"type": "podcast_eq"}, {"type": "robot"}]
    g["fx_chain_out"] = chain.apply_chain(post_out, 24000, fx)
    g["fx_normalize"] = chain.apply_chain(post_out, 24000, [{"type": "normalize", "target_lufs": -20}])
    for room in ("small", "medium", "large"):
        g[f"fx_reverb_{room}"] = chain.apply_chain(post_out, 24000, [{"type": "reverb", "room": room}])
    g["fx_podcast_eq"] = chain.apply_chain(post_out, 24000, [{"type": "podcast_eq"}])
    g["fx_robot"] = chain.apply_chain(post_out, 24000, [{"type": "robot"}])
    g["fx_robot_then_norm"] = chain.apply_chain(post_out, 24000, [{"type": "robot"}, {"type": "normalize", "target_lufs": -18}])
    g["tts_int16"] = float32_to_int16(g["fx_chain_out"])
    g["tts_wav_header"] = np.frombuffer(encode_wav(post_out, 24000)[:44], np.uint8)

    # ---- voice blend (src/tts/backends/kokoro.py:289-308 with a mock pipeline)
    packs = synth.voice_packs(3, seed=61)
    for i, p in enumerate(packs):
        g[f"blend_pack{i}"] = p[:8].copy()  # 8 rows are enough for a fixture
    try:
        import torch
        from unittest.mock import MagicMock
        from src.tts.backends.kokoro import KokoroBackend

        be = KokoroBackend(device="cpu")
        for name, spec, k in (("a2b1", "a(2)+b(1)", 2), ("ab", "a+b", 2), ("a3b2c1", "a(3)+b(2)+c(1)", 3)):
            mp = MagicMock()
            mp.load_voice.side_effect = [torch.from_numpy(p[:8].copy()) for p in packs[:k]]
            be._pipeline = mp
            g[f"blend_{name}"] = be._blend_voices(voices.parse_voice_spec(spec)).numpy()
        blend_src = "KokoroBackend._blend_voices"
    except Exception as e:  # pragma: no cover
        blend_src = f"unavailable: {e!r}"

    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **g)

    # ---- VAD state machines with a scripted session (as tests/test_vad.py:32-43 does)
    class Seq:
        def __init__(self, probs):
            self.probs, self.idx = probs, 0

        def run(self, _n, inputs):
            p = self.probs[self.idx % len(self.probs)]
            self.idx += 1
            return [np.array([[p]], np.float32), inputs["state"]]

    cases = []
    prng = np.random.default_rng(7)
    scripts = [
        [0.9] * 10 + [0.1] * 30,
        [0.9] * 2 + [0.1] * 30,
        [0.9] * 10 + [0.1] * 30 + [0.9] * 10 + [0.1] * 30,
        [0.5] * 8 + [0.49] * 26 + [0.7] * 9,
        [0.1] * 5 + [0.9] * 40,
        [float(np.float32(v)) for v in prng.uniform(0, 1, 300)],
        [float(np.float32(v)) for v in np.clip(0.5 + 0.5 * np.sin(np.arange(400) / 9.0) + prng.normal(0, 0.1, 400), 0, 1)],
    ]
    for probs in scripts:
        for (thr, ms, sil, extra) in ((0.5, 250, 800, 0), (0.5, 0, 100, 100), (0.35, 96, 320, 511)):
            n = len(probs) * 512 + extra
            vad = SileroVAD(Seq(probs), threshold=thr)
            segs = vad.get_speech_segments(np.zeros(n, np.int16).tobytes(), min_speech_ms=ms, silence_ms=sil)
            cases.append({"probs": probs, "n_samples": n, "threshold": thr, "min_speech_ms": ms,
                          "silence_ms": sil, "segments": [[s.start_ms, s.end_ms] for s in segs]})
    buf_cases = []
    for probs in scripts[:6]:
        for (thr, sil_ms, n_chunk) in ((0.5, 500, 640), (0.5, 100, 1600), (0.6, 300, 320)):
            class V:
                def __init__(self, ps):
                    self.ps, self.i = ps, 0

                def __call__(self, audio):
                    p = self.ps[self.i % len(self.ps)]
                    self.i += 1
                    return p

            b = ab.InputAudioBuffer(vad=V(probs), threshold=thr, silence_duration_ms=sil_ms)
            ev = []
            for i in range(len(probs)):
                for e in b.append(np.zeros(n_chunk, np.int16).tobytes()):
                    ev.append([i, e["type"], e.get("audio_start_ms", e.get("audio_end_ms"))])
            buf_cases.append({"probs": probs, "threshold": thr, "silence_duration_ms": sil_ms,
                              "chunk_samples": n_chunk, "events": ev})
    with open(os.path.join(OUT, "vad_state_machines.json"), "w") as f:
        json.dump({"source": "src/vad/silero.py get_speech_segments + src/realtime/audio_buffer.py InputAudioBuffer "
                             "driven with a scripted session", "blend_source": blend_src,
                   "segments": cases, "input_buffer": buf_cases}, f)
    print("wrote", OUT, "blend:", blend_src, "arrays:", len(g))


if __name__ == "__main__":
    main()
