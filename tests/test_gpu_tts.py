"""GPU parity: TTS post-processing, effects chain, voice blend (BASELINE config 5 chain)."""
import numpy as np
import pytest

from oracle import tts as otts

pytestmark = pytest.mark.gpu
FX = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]


def _rel(got, ref):
    return float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() / max(1e-12, float(np.abs(ref).max())))


def test_postprocessing_golden(gpu, golden):
    from open_speech_b200.audio import postprocessing as post

    utt = golden["tts_in"]
    chunks = [utt[:9000], utt[9000:20000], utt[20000:]]
    out = list(post.process_tts_chunks(iter(chunks), trim=True, normalize=True))
    assert len(out) == 1 and out[0].dtype == np.float32
    assert np.array_equal(out[0], golden["tts_post_out"])  # pure elementwise f32: bit-exact
    assert np.array_equal(post.trim_silence(utt), golden["tts_trim_only"])
    assert np.array_equal(post.normalize_output(utt), golden["tts_norm_only"])
    assert list(post.process_tts_chunks(iter(()))) == []


def test_postprocessing_reference_behaviour(gpu):
    """tests/test_audio_processing.py:39-78 of the reference."""
    from open_speech_b200.audio import postprocessing as post

    x = np.concatenate([np.zeros(100), np.ones(200) * 0.5, np.zeros(100)]).astype(np.float32)
    assert len(post.trim_silence(x, threshold=0.1)) == 200
    z = np.zeros(100, dtype=np.float32)
    assert post.trim_silence(z) is z
    y = post.normalize_output(np.array([0.1, -0.2, 0.4], dtype=np.float32), peak=0.9)
    assert abs(float(np.max(np.abs(y))) - 0.9) < 1e-3
    out = list(post.process_tts_chunks(iter([np.ones(5, np.float32), np.ones(5, np.float32)]), trim=False, normalize=False))
    assert len(out) == 1 and len(out[0]) == 10
    x = np.concatenate([np.zeros(5), np.ones(5) * 0.2, np.zeros(5)]).astype(np.float32)
    o = list(post.process_tts_chunks(iter([x]), trim=True, normalize=True))[0]
    assert len(o) == 5 and float(np.max(np.abs(o))) > 0.9
    assert np.allclose(post.normalize_output(np.zeros(10, np.float32)), 0)
    e = np.zeros(0, np.float32)
    assert post.trim_silence(e) is e and post.normalize_output(e) is e


@pytest.mark.parametrize("key,fx", [
    ("fx_chain_out", FX),
    ("fx_normalize", [{"type": "normalize", "target_lufs": -20}]),
    ("fx_reverb_small", [{"type": "reverb", "room": "small"}]),
    ("fx_reverb_medium", [{"type": "reverb", "room": "medium"}]),
    ("fx_reverb_large", [{"type": "reverb", "room": "large"}]),
    ("fx_podcast_eq", [{"type": "podcast_eq"}]),
    ("fx_robot", [{"type": "robot"}]),
    ("fx_robot_then_norm", [{"type": "robot"}, {"type": "normalize", "target_lufs": -18}]),
])
def test_effects_golden(gpu, golden, key, fx):
    """Vectors produced by the reference's own src/effects/chain.py.  Tolerance 1e-4 of the peak (north_star)."""
    from open_speech_b200.effects.chain import apply_chain

    got = apply_chain(golden["tts_post_out"], 24000, fx)
    ref = golden[key]
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert _rel(got, ref) <= 1e-4, _rel(got, ref)


def test_effects_long_utterance_vs_oracle(gpu):
    """12 s utterance: many CTAs per utterance -> exercises the carry-in of the blocked recurrences."""
    from open_speech_b200 import synth
    from open_speech_b200.effects.chain import _reverb, apply_chain

    x = synth.tts_utterance(12.0, seed=3)
    for fx in (FX, [{"type": "reverb", "room": "large", "mix": 0.7}], [{"type": "podcast_eq"}, {"type": "reverb", "room": "small"}]):
        got, ref = apply_chain(x, 24000, fx), otts.apply_chain(x, 24000, fx)
        assert _rel(got, ref) <= 1e-5, (fx, _rel(got, ref))
    # reference behaviours (tests/test_effects_chain.py)
    s = np.ones(1000, dtype=np.float32) * 0.95
    assert np.max(np.abs(apply_chain(s, 24000, [{"type": "normalize", "target_lufs": -20}]))) < 0.95
    r = np.random.default_rng(0).standard_normal(1000).astype(np.float32) * 0.1
    assert np.allclose(apply_chain(r, 24000, []), r) and np.allclose(apply_chain(r, 24000, [{"type": "unknown"}]), r)
    assert np.array_equal(apply_chain(r, 24000, [{"type": "pitch", "semitones": 0}]), r)
    # tests/test_effects_chain.py:14-18: the pitch effect keeps the length and changes the audio
    tone = np.sin(2 * np.pi * 220 * np.arange(24000) / 24000).astype(np.float32)
    shifted = apply_chain(tone, 24000, [{"type": "pitch", "semitones": 4}])
    assert shifted.dtype == np.float32 and len(shifted) == len(tone) and not np.allclose(shifted, tone)
    imp = np.zeros(24000, np.float32)
    imp[0] = 1.0
    assert np.sum(np.abs(_reverb(imp, 24000, room="medium", mix=0.5)[1:])) > 0


@pytest.mark.parametrize("fx", [
    [{"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}],                                 # float32 in, float64 out of the chain
    [{"type": "normalize"}, {"type": "reverb", "room": "large", "mix": 0.6}, {"type": "podcast_eq"}, {"type": "robot"}],
    [{"type": "robot"}, {"type": "reverb", "room": "small"}, {"type": "podcast_eq"}, {"type": "normalize"}],  # float64 into the fused kernel
    [{"type": "reverb"}, {"type": "podcast_eq"}, {"type": "reverb"}, {"type": "podcast_eq"}],
])
def test_reverb_eq_one_kernel_matches_two_kernels_and_oracle(gpu, fx, monkeypatch):
    """reverb directly followed by podcast_eq runs as one kernel (k_fx_reverb_eq): same result as the two-kernel path
    (OSB_FX_UNFUSED=1) and as the oracle, on an utterance long enough for many CTAs and on short ones."""
    from open_speech_b200 import synth
    from open_speech_b200.effects.chain import apply_chain

    for x in (synth.tts_utterance(12.0, seed=5), synth.tts_utterance(0.2, seed=6), np.full(7, 0.3, np.float32)):
        monkeypatch.delenv("OSB_FX_UNFUSED", raising=False)
        fused = apply_chain(x, 24000, fx)
        monkeypatch.setenv("OSB_FX_UNFUSED", "1")
        split = apply_chain(x, 24000, fx)
        monkeypatch.delenv("OSB_FX_UNFUSED", raising=False)
        ref = otts.apply_chain(x, 24000, fx)
        assert fused.dtype == np.float32 and fused.shape == ref.shape
        assert _rel(fused, split) <= 1e-6, (fx, _rel(fused, split))
        assert _rel(fused, ref) <= 1e-5, (fx, _rel(fused, ref))


@pytest.mark.parametrize("semitones", [4, -3, 12, 0.5])
def test_pitch_shift_vs_oracle(gpu, semitones):
    """_pitch_shift (src/effects/chain.py:44-48).  PARITY UNPINNED (librosa + soxr absent): the oracle restates librosa's
    stretch-then-resample with a Kaiser-sinc resampler.  The phase accumulator is float32 by librosa's design, so a
    last-bit difference in an analysis phase can move the accumulated phase of a high bin by one float32 ulp of ~1e5 rad
    (~0.01 rad); the bar is therefore an RMS one: 3e-4 of the signal RMS (measured ~3e-5), and 1e-2 of the peak for the worst sample."""
    from open_speech_b200 import synth
    from open_speech_b200.effects.chain import _pitch_shift, apply_chain

    x = synth.tts_utterance(3.0, seed=21)
    got, ref = _pitch_shift(x, 24000, semitones), otts.pitch_shift(x, 24000, semitones)
    assert got.dtype == np.float32 and got.shape == ref.shape == x.shape
    rms = float(np.sqrt(np.mean(ref.astype(np.float64) ** 2)))
    err = got.astype(np.float64) - ref
    assert np.sqrt(np.mean(err**2)) <= 3e-4 * rms, (np.sqrt(np.mean(err**2)), rms)
    assert np.abs(err).max() <= 1e-2 * np.abs(ref).max(), (np.abs(err).max(), np.abs(ref).max())
    # the pitch really moves: dominant frequency of a tone scales by 2^(n/12)
    tone = (0.5 * np.sin(2 * np.pi * 440 * np.arange(48000) / 24000)).astype(np.float32)
    y = _pitch_shift(tone, 24000, semitones)
    f = np.fft.rfftfreq(16384, 1 / 24000)[np.argmax(np.abs(np.fft.rfft(y[8000 : 8000 + 16384] * np.hanning(16384))))]
    assert abs(f - 440 * 2 ** (semitones / 12)) <= 3.0, f
    # inside a chain: float64 state is cast to float32 first, later effects continue from the float32 result
    fx = [{"type": "podcast_eq"}, {"type": "pitch", "semitones": semitones}, {"type": "normalize", "target_lufs": -20}]
    g2, r2 = apply_chain(x, 24000, fx), otts.apply_chain(x, 24000, fx)
    assert np.sqrt(np.mean((g2.astype(np.float64) - r2) ** 2)) <= 1e-3 * np.sqrt(np.mean(r2.astype(np.float64) ** 2))


def test_pitch_shift_edge_lengths(gpu):
    """shorter than one hop / one frame, and exactly on frame boundaries."""
    from open_speech_b200.effects.chain import _pitch_shift

    rng = np.random.default_rng(5)
    for n in (1, 100, 511, 512, 2047, 2048, 2049, 5000):
        x = (0.3 * rng.standard_normal(n)).astype(np.float32)
        got, ref = _pitch_shift(x, 24000, 3), otts.pitch_shift(x, 24000, 3)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-2 * max(np.abs(ref).max(), 1e-3), (n, np.abs(got - ref).max())


def test_voice_blend_golden_bit_exact(gpu, golden):
    import torch
    from open_speech_b200.tts.blend import blend_voice_arrays, blend_voices
    from open_speech_b200.tts.voices import parse_voice_spec

    packs = [golden[f"blend_pack{i}"] for i in range(3)]
    for name, spec, k in (("a2b1", "a(2)+b(1)", 2), ("ab", "a+b", 2), ("a3b2c1", "a(3)+b(2)+c(1)", 3)):
        s = parse_voice_spec(spec)
        out = blend_voice_arrays(packs[:k], s.normalized_weights())
        assert out.shape == packs[0].shape and np.array_equal(out, golden[f"blend_{name}"]), name

    class P:  # the reference's own test (tests/test_tts_kokoro.py:128-147): 50/50 of 3.0 and 6.0 = 4.5
        def __init__(self):
            self.calls = []

        def load_voice(self, vid):
            self.calls.append(vid)
            return torch.ones(10) * (3.0 if len(self.calls) == 1 else 6.0)

    p = P()
    r = blend_voices(p, parse_voice_spec("af_bella+af_sky"))
    assert isinstance(r, torch.Tensor) and torch.allclose(r, torch.ones(10) * 4.5) and p.calls == ["af_bella", "af_sky"]
    assert parse_voice_spec("alloy").primary_id == "af_heart"
    assert parse_voice_spec("a(2)+b(1)").normalized_weights() == [2 / 3, 1 / 3]
    with pytest.raises(ValueError):
        parse_voice_spec("a(+b")


def test_tts_batch_config5_chain(gpu):
    """Ragged batch through the device entry points == per-utterance oracle chain (trim+norm -> effects -> int16)."""
    from open_speech_b200 import synth
    from open_speech_b200.batch import TtsPost

    utts = synth.tts_batch(12, seed=synth.SEED_C5, distinct=12, min_s=0.5, max_s=3.0)
    pcm, lens = TtsPost(sample_rate=24000, effects=FX).run_numpy(utts)
    for i, u in enumerate(utts):
        ref = otts.tts_chain([u], FX)
        got = pcm[i]
        assert len(got) == len(ref) == lens[i]
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 3, (i, int(d.max()))  # 1e-4 of full scale


@pytest.mark.parametrize("fx", [
    [{"type": "reverb", "room": "small"}],                                              # recurrence first: trim / peak gain while loading
    [{"type": "normalize", "target_lufs": -20}, {"type": "podcast_eq"}, {"type": "robot"}],  # normalise + recurrence: both gains deferred
    [{"type": "podcast_eq"}, {"type": "pitch", "semitones": 2}],                         # deferred head, float32 producer later in the chain
    [{"type": "robot"}, {"type": "reverb"}],                                             # other head: materialised post-processing
    [{"type": "normalize"}],                                                             # materialised, sum of squares handed over
    [],                                                                                  # no effects: post-processing + cast only
])
@pytest.mark.parametrize("trim,normalize", [(True, True), (False, True), (True, False)])
def test_tts_post_fx_one_call_matches_two_calls(gpu, fx, trim, normalize):
    """osb_tts_post_fx_dev == osb_tts_post_dev followed by osb_fx_chain_dev, for every chain head it special-cases,
    including utterances that are silent (nothing above the trim threshold, peak below 1e-8)."""
    import torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import TtsPost
    from open_speech_b200.effects.chain import encode_effects

    utts = synth.tts_batch(6, seed=77, distinct=6, min_s=0.4, max_s=1.5)
    utts.append(np.zeros(5000, np.float32))                      # silent: returned unchanged by both reference functions
    utts.append((0.004 * np.ones(3000)).astype(np.float32))      # below the trim threshold everywhere, but normalisable
    flat, offsets, lens = TtsPost.pack(utts)
    d_flat, d_off, d_len = torch.from_numpy(flat).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(lens).cuda()
    b, total, max_len = len(utts), flat.size, int(lens.max())
    types, p0, p1 = encode_effects(fx)
    stream = torch.cuda.current_stream().cuda_stream
    post = torch.empty_like(d_flat)
    lens_a, lens_b = torch.empty_like(d_len), torch.empty_like(d_len)
    out_a = torch.zeros(total, dtype=torch.int16, device="cuda")
    out_b = torch.zeros(total, dtype=torch.int16, device="cuda")
    gpu.call("osb_tts_post_dev", d_flat.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), b, max_len, int(trim), int(normalize), 0.01, 0.95,
             post.data_ptr(), lens_a.data_ptr(), stream)
    gpu.call("osb_fx_chain_dev", post.data_ptr(), d_off.data_ptr(), lens_a.data_ptr(), b, max_len, total, 24000, gpu.ptr(types), gpu.ptr(p0),
             gpu.ptr(p1), len(types), out_a.data_ptr(), 1, stream)
    gpu.call("osb_tts_post_fx_dev", d_flat.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), b, max_len, total, int(trim), int(normalize),
             0.01, 0.95, 24000, gpu.ptr(types), gpu.ptr(p0), gpu.ptr(p1), len(types), 0, lens_b.data_ptr(), out_b.data_ptr(), 1, stream)
    torch.cuda.synchronize()
    assert torch.equal(lens_a, lens_b)
    a, c = out_a.cpu().numpy(), out_b.cpu().numpy()
    for i in range(b):
        o, n = int(offsets[i]), int(lens_a[i].item())
        d = np.abs(a[o:o + n].astype(np.int32) - c[o:o + n].astype(np.int32))
        assert d.max(initial=0) <= 1, (i, int(d.max()))  # the summation order of the RMS may move a gain by one ulp
