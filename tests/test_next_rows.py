"""composer mix path and the Wyoming float resampler -- oracle pin (CPU) and GPU parity."""
import os

import numpy as np
import pytest

from oracle import next_rows as nx


@pytest.fixture(scope="module")
def gnext():
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_next.npz")))


def _tracks(g):
    a, b = g["comp_in_24k"], g["comp_track_b"]
    return [{"samples": a, "offset_s": 0.0}, {"samples": b, "offset_s": 0.25}, {"samples": a[:5000] * 2.0, "offset_s": 0.9}]


def test_oracle_matches_reference_vectors(gnext):
    a = gnext["comp_in_24k"]
    for dst in (16000, 48000, 22050, 8000):
        assert np.array_equal(nx.composer_resample(a, 24000, dst), gnext[f"comp_resample_24k_{dst}"])
    assert np.array_equal(nx.mix_prepared(_tracks(gnext), 24000), gnext["comp_mix"])
    assert np.array_equal(nx.resample_to_16k(a, 24000), gnext["wy_24k_to_16k"])
    assert np.array_equal(nx.resample_to_16k(a[:7777], 22050), gnext["wy_22050_to_16k"])
    assert np.array_equal(nx.resample_to_16k(a, 48000), gnext["wy_48k_to_16k"])


@pytest.mark.gpu
def test_composer_gpu_bit_exact(gpu, gnext):
    from open_speech_b200 import composer

    a = gnext["comp_in_24k"]
    for dst in (16000, 48000, 22050, 8000):
        got = composer.resample(a, 24000, dst)
        assert got.dtype == np.float32 and np.array_equal(got, gnext[f"comp_resample_24k_{dst}"]), dst
    assert composer.resample(a, 24000, 24000) is a
    mixed = composer.mix_prepared(_tracks(gnext), 24000)
    assert np.array_equal(mixed, gnext["comp_mix"])
    assert np.array_equal(composer.float_to_int16(mixed), gnext["comp_int16"])
    # the reference's own unit tests (tests/test_composer_unit.py:13-41)
    pulse = np.zeros(100, dtype=np.float32)
    pulse[0] = 1.0
    m = composer.mix_prepared([{"samples": pulse, "offset_s": 0.1}], sample_rate=1000)
    assert len(m) == 200 and m[100] == 1.0 and np.allclose(m[:100], 0.0)
    t = np.arange(2400) / 24000
    s = (0.2 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    m = composer.mix_prepared([{"samples": s, "offset_s": 0.0}, {"samples": s, "offset_s": 0.0}], 24000)
    assert np.isclose(np.max(np.abs(m)), np.max(np.abs(s)) * 2, rtol=0.05)
    assert len(composer.mix_prepared([], 24000)) == 0


@pytest.mark.gpu
def test_wyoming_resample_gpu_bit_exact(gpu, gnext):
    from open_speech_b200.wyoming_audio import _resample_to_16k

    a = gnext["comp_in_24k"]
    assert np.array_equal(_resample_to_16k(a, 24000), gnext["wy_24k_to_16k"])
    assert np.array_equal(_resample_to_16k(a[:7777], 22050), gnext["wy_22050_to_16k"])
    assert np.array_equal(_resample_to_16k(a, 48000), gnext["wy_48k_to_16k"])
    assert _resample_to_16k(a, 16000) is a


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [16000, 8000, 48000])
def test_vad_gated_assembly(gpu, rate):
    """wyoming _extract_speech_segments: resample -> VAD -> segments -> gather at the original rate (8(f) row 4)."""
    from open_speech_b200 import synth
    from open_speech_b200.vad.silero import SileroVAD, VadSession
    from open_speech_b200.streaming import resample_pcm16
    from open_speech_b200.wyoming_audio import _extract_speech_segments, _pcm_to_wav
    from oracle import vad as ovad

    sess = VadSession()
    pcm = synth.clip_pcm16(20.0, rate, seed=500 + rate // 1000)
    got = np.frombuffer(_extract_speech_segments(pcm.tobytes(), rate, 2, 1, session=sess), np.int16)
    # the reference's procedure, with the GPU's own 16 kHz audio and segments: the gather must be bit-exact
    p16 = pcm.tobytes() if rate == 16000 else resample_pcm16(pcm.tobytes(), rate, 16000)
    segs = SileroVAD(sess).get_speech_segments(p16)
    spm = rate // 1000
    parts = [pcm[s.start_ms * spm: min(s.end_ms * spm, len(pcm))] for s in segs if s.start_ms * spm < len(pcm)]
    want = np.concatenate(parts) if parts else pcm
    assert len(segs) >= 2 and np.array_equal(got, want)
    assert 0 < len(got) < len(pcm)
    # end to end against the oracle network unless a probability sits within tolerance of the threshold
    probs, _ = ovad.SileroNet().score_stream(np.frombuffer(p16, np.int16).astype(np.float32) / 32768.0)
    if int((np.abs(probs - 0.5) <= 1e-3).sum()) == 0:
        osegs = ovad.segments_from_probs(probs, len(p16) // 2)
        assert [(s.start_ms, s.end_ms) for s in osegs] == [(s.start_ms, s.end_ms) for s in segs]
    # pass-through cases of the reference
    assert _extract_speech_segments(b"", rate, 2, 1, session=sess) == b""
    assert _extract_speech_segments(pcm.tobytes(), rate, 1, 1, session=sess) == pcm.tobytes()
    assert _extract_speech_segments(pcm.tobytes(), rate, 2, 1, session=None) == pcm.tobytes()
    silent = np.zeros(rate * 2, np.int16).tobytes()
    assert _extract_speech_segments(silent, rate, 2, 1, session=sess, threshold=0.99) == silent
    w = _pcm_to_wav(pcm.tobytes(), rate, 2, 1)
    assert w[:4] == b"RIFF" and len(w) == 44 + 2 * len(pcm) and int.from_bytes(w[24:28], "little") == rate


@pytest.mark.gpu
def test_conversation_render_turns(gpu):
    """src/conversation.py:96-158, audio half: effects per turn, 500 ms gaps between turns, durations, per-turn WAVs."""
    from open_speech_b200 import synth
    from open_speech_b200.conversation import SILENCE_MS, render_turns
    from oracle import tts as otts

    turns = [synth.tts_utterance(1.2, seed=70), synth.tts_utterance(0.7, seed=71), synth.tts_utterance(0.9, seed=72)]
    fx = [[{"type": "podcast_eq"}], None, [{"type": "normalize", "target_lufs": -18}, {"type": "robot"}]]
    out = render_turns(turns, fx, 24000)
    gap = np.zeros(int(24000 * SILENCE_MS / 1000), np.float32)
    ref_parts = [otts.apply_chain(t, 24000, f) if f else t for t, f in zip(turns, fx)]
    ref = np.concatenate([ref_parts[0], gap, ref_parts[1], gap, ref_parts[2]])
    assert out["merged"].dtype == np.float32 and out["merged"].shape == ref.shape
    assert np.abs(out["merged"] - ref).max() <= 1e-5 * np.abs(ref).max()
    assert out["duration_ms"] == int(1000 * len(ref) / 24000) and out["turn_duration_ms"] == [int(1000 * len(p) / 24000) for p in ref_parts]
    assert out["turn_wavs"][1] == otts.encode_wav(turns[1], 24000) and len(out["turn_wavs"][0]) == 44 + 2 * len(turns[0])
    assert render_turns([], None, 24000)["merged"].shape == (0,)


# ---------------------------------------------------------------- realtime TTS output framing (src/realtime/server.py:238-277)
_RT_CASES = [(tag, fmt) for tag in ("full", "odd", "tiny") for fmt in ("pcm16", "g711_ulaw", "g711_alaw")]


def _rt_chunks(g, tag):
    rt = g["rt_in_24k"]
    n = {"full": len(rt), "odd": 4001, "tiny": 2}[tag]
    return [rt[:n // 2], rt[n // 2:n]]


def test_oracle_realtime_framing_matches_reference_vectors(gnext):
    for tag, fmt in _RT_CASES:
        chunks = _rt_chunks(gnext, tag)
        assert nx.realtime_response_payload(chunks, fmt) == gnext[f"rt_payload_{tag}_{fmt}"].tobytes(), (tag, fmt)
        assert "\n".join(nx.realtime_deltas(chunks, fmt)) == gnext[f"rt_deltas_{tag}_{fmt}"].tobytes().decode("ascii"), (tag, fmt)
    assert nx.realtime_response_payload([], "pcm16") == b"" and nx.realtime_deltas([], "g711_ulaw") == []


@pytest.mark.gpu
def test_realtime_framing_gpu_bit_exact(gpu, gnext):
    import base64

    from open_speech_b200.realtime import tts_out

    for tag, fmt in _RT_CASES:
        chunks = _rt_chunks(gnext, tag)
        want_payload = gnext[f"rt_payload_{tag}_{fmt}"].tobytes()
        want_deltas = gnext[f"rt_deltas_{tag}_{fmt}"].tobytes().decode("ascii")
        assert tts_out.encode_response_audio(chunks, fmt) == want_payload, (tag, fmt)
        got = tts_out.audio_deltas(chunks, fmt)
        assert "\n".join(got) == want_deltas, (tag, fmt)
        assert all(len(d) == 4000 for d in got[:-1]) and b"".join(base64.b64decode(d) for d in got) == want_payload
    # list chunks go through np.array(dtype=float32) like the reference; nothing to send -> nothing sent
    assert tts_out.audio_deltas([[0.5, -0.25, 1.0]], "pcm16") == nx.realtime_deltas([[0.5, -0.25, 1.0]], "pcm16")
    assert tts_out.encode_response_audio([], "pcm16") == b"" and tts_out.audio_deltas([], "g711_alaw") == []
    assert tts_out.audio_deltas([np.zeros(0, np.float32)], "pcm16") == []
    assert tts_out.audio_deltas([np.zeros(2, np.float32)], "g711_ulaw") == []  # int(2 / 3) == 0 output samples
    with pytest.raises(ValueError):
        tts_out.audio_deltas([np.zeros(8, np.float32)], "opus")
    # every payload length modulo 12 (vector groups, 3-byte tail, '=' padding) against the stdlib encoder
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [3000, 3001, 6007]:
        x = (rng.random(n, dtype=np.float32) * 2.2 - 1.1).astype(np.float32)
        assert "".join(tts_out.audio_deltas([x], "pcm16")) == "".join(nx.realtime_deltas([x], "pcm16")), n
    big = (rng.random(1_000_003, dtype=np.float32) * 2.0 - 1.0).astype(np.float32)  # full-size response: a checksum of the text
    for fmt in ("pcm16", "g711_ulaw"):
        assert tts_out.audio_deltas([big], fmt) == nx.realtime_deltas([big], fmt), fmt
