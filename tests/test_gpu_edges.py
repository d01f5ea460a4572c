"""GPU parity at the edges of the size range: clips shorter than one FFT frame, one hop, one sample; chains of effects
in every adjacency the fused kernels special-case.  Oracle = oracle/ (see its headers for what is pinned)."""
import numpy as np
import pytest

from oracle import stt, tts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 511, 512, 1023, 1024, 1025, 4095, 30000, 30001, 59999])
def test_spectral_gate_tiny_clips(gpu, n):
    from open_speech_b200.audio import preprocessing as pre

    a = (0.1 * np.random.default_rng(n).standard_normal(n)).astype(np.float32)
    got, ref = pre.reduce_noise(a, 16000), stt.spectral_gate(a, 16000)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max(), (n, np.abs(got - ref).max())


@pytest.mark.parametrize("n", [41, 100, 400, 1000, 16000])
@pytest.mark.parametrize("noise_reduce,normalize", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_stt_frontend_short_clips(gpu, n, noise_reduce, normalize):
    """every flag combination of preprocess_stt_audio in front of the log-mel, on clips down to the reflect-pad minimum."""
    pcm = (3000 * np.random.default_rng(n).standard_normal((2, n))).astype(np.int16)
    nf = gpu.lib().osb_logmel_frames(n)
    mel = np.empty((2, 128, nf), np.float32)
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, 2, n, 16000, noise_reduce, normalize, 128, gpu.ptr(mel))
    for i in range(2):
        ref = stt.stt_frontend(pcm[i], noise_reduce=bool(noise_reduce), normalize=bool(normalize))
        err = np.abs(mel[i] - ref) / np.maximum(1.0, np.abs(ref))
        # the chain requantises to int16 before the log-mel: after a float32 gain (or the denoiser) an occasional sample
        # lands on the neighbouring LSB, which moves the near-silent cells of its frames (see test_gpu_denoise, test_gpu_logmel)
        bulk = 0.97 if noise_reduce else (0.999 if normalize else 1.0)
        assert (err <= 1e-4).mean() >= bulk and err.max() <= 5e-3, (i, float((err <= 1e-4).mean()), float(err.max()))


FX_CHAINS = [
    [{"type": "normalize"}, {"type": "podcast_eq"}],                                        # gain applied while the EQ loads
    [{"type": "normalize"}, {"type": "reverb", "room": "large"}, {"type": "robot"}],         # gain in, robot + cast out
    [{"type": "robot"}, {"type": "normalize"}, {"type": "podcast_eq"}, {"type": "robot"}],   # float64 normalise is not deferred
    [{"type": "podcast_eq"}, {"type": "reverb"}, {"type": "normalize", "target_lufs": -23}],  # recurrence -> recurrence -> stand-alone
    [{"type": "reverb"}, {"type": "unknown"}, {"type": "pitch", "semitones": 0}, {"type": "robot"}],  # no-ops between fused neighbours
]


@pytest.mark.parametrize("fx", FX_CHAINS)
@pytest.mark.parametrize("sr", [16000, 22050, 24000])
def test_effect_chain_adjacencies(gpu, fx, sr):
    from open_speech_b200 import synth
    from open_speech_b200.effects.chain import apply_chain

    x = synth.tts_utterance(1.7, seed=5)
    got, ref = apply_chain(x, sr, fx), tts.apply_chain(x, sr, fx)
    assert got.dtype == np.float32 and got.shape == ref.shape
    assert np.abs(got.astype(np.float64) - ref).max() <= 1e-5 * np.abs(ref).max(), np.abs(got.astype(np.float64) - ref).max()


@pytest.mark.parametrize("fx", [[{"type": "pitch", "semitones": -5}, {"type": "reverb"}], [{"type": "reverb"}, {"type": "robot"}, {"type": "pitch", "semitones": 2}]])
def test_effect_chain_with_pitch(gpu, fx):
    from open_speech_b200 import synth
    from open_speech_b200.effects.chain import apply_chain

    x = synth.tts_utterance(1.7, seed=5)
    got, ref = apply_chain(x, 24000, fx), tts.apply_chain(x, 24000, fx)
    e = np.sqrt(np.mean((got.astype(np.float64) - ref) ** 2)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    assert got.shape == ref.shape and e <= 3e-4, e
