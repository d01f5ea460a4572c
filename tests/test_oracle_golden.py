"""CPU: the oracle restatements reproduce the vectors the reference itself produced."""
import hashlib

import numpy as np
import pytest

from oracle import codec, resample, stt, tts, vad


def test_g711_tables_match_survey_hashes():
    allb = bytes(range(256))
    assert hashlib.sha256(codec.ulaw2lin(allb)).hexdigest() == "3dab54339e520bb2c924826e3b72a917a2b612e9fd12fc867500f1d983a75827"
    assert hashlib.sha256(codec.alaw2lin(allb)).hexdigest() == "e04788d110e58ff8c70c93b8480190d973e3b67876b6119abbaec766cc75c174"
    all16 = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16).tobytes()
    assert hashlib.sha256(codec.lin2ulaw(all16)).hexdigest() == "81d633c9e6972a18c74a58720b96cb8ca0bdd096d4060b646dd708c3b846019a"
    assert hashlib.sha256(codec.lin2alaw(all16)).hexdigest() == "38488f6fd710f4686360edc4d38639f96c491595ef93f8eb8d62d5e07ca6ce7b"


def test_codec_golden(golden):
    g = golden
    ul, al = g["codec_ulaw_in"].tobytes(), g["codec_alaw_in"].tobytes()
    assert codec.decode_audio_to_pcm16(bytes(range(256)), "g711_ulaw", 8000) == g["codec_ulaw_all256_8k"].tobytes()
    assert codec.decode_audio_to_pcm16(bytes(range(256)), "g711_alaw", 8000) == g["codec_alaw_all256_8k"].tobytes()
    assert codec.decode_audio_to_pcm16(ul, "g711_ulaw", 16000) == g["codec_ulaw_16k"].tobytes()
    assert codec.decode_audio_to_pcm16(al, "g711_alaw", 16000) == g["codec_alaw_16k"].tobytes()
    assert codec.decode_audio_to_pcm16(g["codec_pcm24k_in"].tobytes(), "pcm16", 16000) == g["codec_pcm24k_16k"].tobytes()
    p = g["codec_pcm16k_in"].tobytes()
    assert codec.encode_pcm16_to_format(p, 16000, "g711_ulaw") == g["codec_enc_ulaw"].tobytes()
    assert codec.encode_pcm16_to_format(p, 16000, "g711_alaw") == g["codec_enc_alaw"].tobytes()
    assert codec.encode_pcm16_to_format(p, 16000, "pcm16") == g["codec_enc_pcm16"].tobytes()
    all16 = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    assert np.array_equal(codec.lin2ulaw_array(all16), g["codec_lin2ulaw_all"])
    assert np.array_equal(codec.lin2alaw_array(all16), g["codec_lin2alaw_all"])


@pytest.mark.parametrize("n,m", [(160, 320), (480, 320), (2400, 1600), (1, 3), (2, 5), (7, 3), (441, 160), (3, 1), (100, 37)])
def test_interp_closed_form_equals_np_interp(n, m):
    rng = np.random.default_rng(n * 1000 + m)
    for _ in range(3):
        x = rng.integers(-32768, 32768, n).astype(np.int16)
        ref = np.interp(np.linspace(0, 1, m), np.linspace(0, 1, n), x.astype(np.float32)).astype(np.int16)
        assert np.array_equal(ref, codec.interp_explicit(x, m))


@pytest.mark.parametrize("fr", [8000, 24000, 48000, 44100, 22050, 32000])
def test_poly_golden(golden, fr):
    x = golden[f"poly_{fr}_in"].tobytes()
    assert resample.resample_pcm16(x, fr, 16000) == golden[f"poly_{fr}_out"].tobytes()
    assert resample.resample_pcm16_restated(x, fr, 16000) == golden[f"poly_{fr}_out"].tobytes()


def test_poly_edge_golden(golden):
    assert resample.resample_pcm16_restated(golden["poly_up_in"].tobytes(), 16000, 48000) == golden["poly_up_48k_out"].tobytes()
    assert resample.resample_pcm16(np.array([1000], np.int16).tobytes(), 16000, 32000) == golden["poly_single_out"].tobytes()
    assert resample.resample_pcm16(b"", 16000, 48000) == b""


def test_stt_pre_golden(golden):
    g = golden
    wav = stt.float32_mono_to_wav_bytes(g["pre_in_pcm16"].astype(np.float32) / 32768.0, 16000)
    # the reference's requantisation is not idempotent (x/32768*32767): start from its own WAV
    import io, wave
    b = io.BytesIO()
    with wave.open(b, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(16000); wf.writeframes(g["pre_in_pcm16"].tobytes())
    wav = b.getvalue()
    out = stt.preprocess_stt_audio(wav, noise_reduce=False, normalize=True)
    assert out[:44] == g["pre_header"].tobytes()
    assert out[44:] == g["pre_norm_out_pcm16"].tobytes()
    assert stt.preprocess_stt_audio(wav, noise_reduce=False, normalize=False)[44:] == g["pre_requant_out_pcm16"].tobytes()
    a, _ = stt.wav_bytes_to_float32_mono(wav)
    assert np.array_equal(stt.normalize_gain(a), g["pre_gain_f32"])
    assert np.array_equal(stt.normalize_gain(g["pre_quiet_in"]), g["pre_quiet_gain"])
    assert stt.preprocess_stt_audio(b"not a wav", noise_reduce=False, normalize=True) == b"not a wav"


def test_logmel_matches_hf_extractor():
    """Independent cross-check of the unpinned log-mel restatement."""
    transformers = pytest.importorskip("transformers")
    from open_speech_b200 import synth

    a = synth.clip_pcm16(5.0, seed=5).astype(np.float32) / 32768.0
    fe = transformers.WhisperFeatureExtractor(feature_size=128)
    hf = fe._np_extract_fbank_features(np.pad(a, (0, 160))[None], "cpu")[0]
    m = stt.logmel(a, 128)
    assert m.shape == hf.shape == (128, stt.logmel_n_frames(len(a)))
    assert np.abs(hf - m).max() < 2e-5
    assert np.abs(fe.mel_filters.T - stt.mel_filters(16000, 400, 128)).max() < 1e-7
    assert stt.logmel(np.zeros(480000, np.float32), 128).shape == (128, 3001)


def test_tts_golden(golden):
    g = golden
    utt = g["tts_in"]
    chunks = [utt[:9000], utt[9000:20000], utt[20000:]]
    post = list(tts.process_tts_chunks(iter(chunks)))[0]
    assert np.array_equal(post, g["tts_post_out"])
    assert np.array_equal(tts.trim_silence(utt), g["tts_trim_only"])
    assert np.array_equal(tts.normalize_output(utt), g["tts_norm_only"])
    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
    assert np.array_equal(tts.apply_chain(post, 24000, fx), g["fx_chain_out"])
    for room in ("small", "medium", "large"):
        assert np.array_equal(tts.apply_chain(post, 24000, [{"type": "reverb", "room": room}]), g[f"fx_reverb_{room}"])
    assert np.array_equal(tts.apply_chain(post, 24000, [{"type": "podcast_eq"}]), g["fx_podcast_eq"])
    assert np.array_equal(tts.apply_chain(post, 24000, [{"type": "robot"}]), g["fx_robot"])
    assert np.array_equal(tts.float32_to_int16(g["fx_chain_out"]), g["tts_int16"])
    assert tts.encode_wav(post, 24000)[:44] == g["tts_wav_header"].tobytes()
    assert list(tts.process_tts_chunks(iter(()))) == []


def test_blend_golden(golden):
    packs = [golden[f"blend_pack{i}"] for i in range(3)]
    for name, spec, k in (("a2b1", "a(2)+b(1)", 2), ("ab", "a+b", 2), ("a3b2c1", "a(3)+b(2)+c(1)", 3)):
        comps = tts.parse_voice_spec(spec)
        out = tts.blend_voices(packs[:k], tts.normalized_weights(comps))
        assert np.array_equal(out, golden[f"blend_{name}"]), name
    assert tts.normalized_weights(tts.parse_voice_spec("a(2)+b(1)")) == [2 / 3, 1 / 3]
    with pytest.raises(ValueError):
        tts.parse_voice_spec("a(+b")


def test_vad_state_machines_golden(golden_vad):
    for c in golden_vad["segments"]:
        segs = vad.segments_from_probs(c["probs"], c["n_samples"], c["threshold"], c["min_speech_ms"], c["silence_ms"])
        assert [[s.start_ms, s.end_ms] for s in segs] == c["segments"]
    for c in golden_vad["input_buffer"]:
        ev = vad.input_buffer_events(c["probs"], [c["chunk_samples"]] * len(c["probs"]), c["threshold"], c["silence_duration_ms"])
        assert [[i, t, ms] for i, t, ms in ev] == c["events"]


def test_silero_net_run_contract():
    net = vad.SileroNet()
    st = np.zeros((2, 1, 128), np.float32)
    out, st2 = net.run(None, {"input": np.zeros((1, 512), np.float32), "state": st, "sr": np.array(16000)})
    assert out.shape == (1, 1) and st2.shape == (2, 1, 128) and 0.0 < float(out[0][0]) < 1.0
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(512 * 5 + 100) * 0.1).astype(np.float32)
    probs, _ = net.score_stream(a)
    # window-by-window through run() == batched front + loop
    s = np.zeros((2, 1, 128), np.float32)
    for k in range(5):
        o, s = net.run(None, {"input": a[None, 512 * k:512 * (k + 1)], "state": s, "sr": np.array(16000)})
        assert abs(float(o[0][0]) - float(probs[k])) < 1e-5


def test_stream_gate_oracle_matches_reference_traces():
    """oracle.vad.stream_gate_steps == the reference's StreamingSession._process_chunk, step by step (oracle/make_golden_stream.py)."""
    import json
    import os
    from math import gcd

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stream_gate.json")) as f:
        g = json.load(f)
    n_steps = 0
    for c in g["cases"]:
        sr, n = c["sample_rate"], c["chunk_samples"]
        d = gcd(16000, sr)
        n16 = n if sr == 16000 else (n * (16000 // d) + sr // d - 1) // (sr // d)
        steps = vad.stream_gate_steps(c["probs"], n16, vad_enabled=c["vad_enabled"], threshold=c["threshold"],
                                      endpointing_samples=int(16000 * c["endpointing_ms"] / 1000), max_utterance_bytes=g["max_utterance_bytes"])
        S = c["steps"]
        for i, (act, active, sil, utt) in enumerate(steps):
            want = (S["speech_active"][i], S["silence_samples"][i], S["utterance_bytes"][i], S["speech_start"][i], S["speech_end"][i],
                    S["transcribe_calls"][i], S["final"][i])
            got = (int(active), sil, utt, int(bool(act & 1)), int(bool(act & 32)), int(bool(act & 24)), int(bool(act & 16)))
            assert got == want, (c["name"], i)
            n_steps += 1
    assert n_steps > 2000
