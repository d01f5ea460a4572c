"""GPU parity: spectral-gating noise reduction and the composed STT front-end (BASELINE config 4 chain).

Oracle = scipy restatement of noisereduce's non-stationary gate in float64 (PARITY UNPINNED by the reference:
noisereduce is an absent optional dependency).  Tolerance: 1e-4 relative to the clip's peak for the audio,
then the per-stage tolerances of test_gpu_logmel for what follows.
"""
import io
import wave

import numpy as np
import pytest

from oracle import stt

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _noisy(seconds, seed):
    from open_speech_b200 import synth

    return synth.clip_pcm16(seconds, seed=seed, extra_noise_rms=0.01)


@pytest.mark.parametrize("seconds", [0.5, 5.0, 40.0])  # 40 s = 640,000 samples -> two 600,000-sample chunks
def test_spectral_gate_vs_oracle(gpu, seconds):
    from open_speech_b200.audio import preprocessing as pre

    a = _noisy(seconds, 9).astype(np.float32) / 32768.0
    got = pre.reduce_noise(a, 16000)
    ref = stt.spectral_gate(a, 16000)
    assert got.dtype == np.float32 and got.shape == ref.shape
    peak = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max())
    assert err <= TOL * peak, (err, peak)
    # it really denoises: the gaps between bursts lose most of their noise energy
    assert float(np.mean(got**2)) < float(np.mean(a**2))


@pytest.mark.parametrize("sr", [8000, 22050, 24000, 44100, 48000])
def test_spectral_gate_other_sample_rates(gpu, sr):
    """reduce_noise(y, sr) takes the WAV's own rate: the mask smoothing reach follows it (8 kHz: 32x1 bins, 48 kHz: 5x9),
    which runs the generic tap kernel instead of the 16 kHz running-sum instance."""
    from open_speech_b200.audio import preprocessing as pre

    a = _noisy(3.0, 17).astype(np.float32) / 32768.0
    got, ref = pre.reduce_noise(a, sr), stt.spectral_gate(a, sr)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), (sr, np.abs(got - ref).max(), np.abs(ref).max())


@pytest.mark.parametrize("n", [600001 + 255, 600000 + 4096 * 3 + 17, 1200000, 1200001])
def test_spectral_gate_short_last_chunk(gpu, n):
    """last chunk much shorter than 600,000 samples: most of its frames lie in the zero padding and are skipped."""
    from open_speech_b200.audio import preprocessing as pre

    base = _noisy(76.0, 23).astype(np.float32) / 32768.0
    a = base[:n]
    got, ref = pre.reduce_noise(a, 16000), stt.spectral_gate(a, 16000)
    assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), n


def test_spectral_gate_batch_matches_single(gpu):
    """device batch entry: ragged-free batch of clips == clip-by-clip host calls (and the fused sum of squares feeds the
    same gain as the stand-alone normalise)."""
    import torch
    from open_speech_b200 import synth

    pcm = synth.clip_batch_pcm16(4, 7.0, seed=31, distinct=4)
    n = pcm.shape[1]
    x = torch.from_numpy(pcm).cuda()
    out = torch.empty((4, n), dtype=torch.float32, device="cuda")
    gpu.call("osb_spectral_gate_dev", x.data_ptr(), 0, n, 4, n, 16000, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(4):
        one = np.empty(n, np.float32)
        gpu.call("osb_spectral_gate_host", gpu.ptr(pcm[i]), 0, gpu.ptr(one), n, 16000)
        assert np.array_equal(out[i].cpu().numpy(), one)


def test_spectral_gate_chunk_boundary_exact_multiple(gpu):
    """n == 600,000 is the single-chunk limit; n == 600,001 switches to two chunks (oracle get_traces)."""
    from open_speech_b200.audio import preprocessing as pre

    base = _noisy(37.6, 11).astype(np.float32) / 32768.0
    for n in (600000, 600001):
        a = base[:n]
        got, ref = pre.reduce_noise(a, 16000), stt.spectral_gate(a, 16000)
        assert np.abs(got - ref).max() <= TOL * np.abs(ref).max(), n


def test_preprocess_with_noise_reduce(gpu):
    from open_speech_b200.audio import preprocessing as pre

    pcm = _noisy(5.0, 13)
    b = io.BytesIO()
    with wave.open(b, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(16000); wf.writeframes(pcm.tobytes())
    out = pre.preprocess_stt_audio(b.getvalue(), noise_reduce=True, normalize=True)
    ref = stt.preprocess_stt_audio(b.getvalue(), noise_reduce=True, normalize=True)
    x, y = np.frombuffer(out[44:], np.int16).astype(np.int32), np.frombuffer(ref[44:], np.int16).astype(np.int32)
    assert out[:44] == ref[:44] and len(x) == len(y)
    # 1e-4 of full scale = 3.3 LSB
    assert np.abs(x - y).max() <= 3, int(np.abs(x - y).max())


def test_stt_frontend_config4_chain(gpu):
    """denoise -> normalise -> requantise -> log-mel on a batch; per-clip parity vs the oracle chain."""
    import torch
    from open_speech_b200 import synth

    pcm = synth.clip_batch_pcm16(3, 12.0, seed=synth.SEED_C4, distinct=3)
    n = pcm.shape[1]
    x = torch.from_numpy(pcm).cuda()
    nf = gpu.lib().osb_logmel_frames(n)
    out = torch.empty((3, 128, nf), dtype=torch.float32, device="cuda")
    gpu.call("osb_stt_frontend_dev", x.data_ptr(), n, 3, n, 16000, 1, 1, 128, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for i in range(3):
        # stage-wise: (1) the GPU's denoised+normalised int16 is within 1e-4 of full scale (3 LSB) of the oracle's,
        q = np.empty(n, np.int16)
        gpu.call("osb_preprocess_stt_host", gpu.ptr(pcm[i]), n, 1, 16000, 1, 1, -18.0, gpu.ptr(q))
        a = stt.normalize_gain(stt.spectral_gate(pcm[i].astype(np.float32) / 32768.0, 16000))
        dq = np.abs(q.astype(np.int32) - stt.quantise_pcm16(a).astype(np.int32))
        assert dq.max() <= 3 and (dq != 0).mean() <= 0.01, (int(dq.max()), float((dq != 0).mean()))
        # (2) the log-mel of the batch path equals the oracle log-mel of that int16 within 1e-4,
        got = out[i].cpu().numpy()
        ref_stage = stt.logmel(q.astype(np.float32) / 32768.0, 128)
        assert (np.abs(got - ref_stage) / np.maximum(1.0, np.abs(ref_stage))).max() <= TOL
        # (3) end to end.  The denoised audio itself is within ~5e-7 of the f64 oracle (tools/nr_err.py), but the chain
        # requantises to int16 in between: < 0.1 % of samples round to the neighbouring LSB, and in gated (near-silent)
        # stretches, where the signal is a few LSB, one flipped sample moves every mel cell of the frames that contain it
        # by more than 1e-4.  (1) and (2) pin the stages; here only the bulk and the worst cell are bounded.
        ref = stt.stt_frontend(pcm[i], noise_reduce=True, normalize=True)
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        assert (err <= TOL).mean() >= 0.97 and err.max() <= 5e-3, ((err <= TOL).mean(), err.max())
    # host-pointer entry == device entry
    mel = np.empty((3, 128, nf), np.float32)
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, 3, n, 16000, 1, 1, 128, gpu.ptr(mel))
    assert np.array_equal(mel, out.cpu().numpy())
    # no denoise, no normalise: requantise-only chain is exact up to the mel tolerance
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, 3, n, 16000, 0, 0, 128, gpu.ptr(mel))
    ref = stt.stt_frontend(pcm[0], noise_reduce=False, normalize=False)
    assert (np.abs(mel[0] - ref) / np.maximum(1.0, np.abs(ref))).max() <= TOL
