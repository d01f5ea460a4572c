"""GPU parity: Whisper log-mel front-end vs the oracle restatement (tolerance from north_star:
1e-4 relative for mel, applied as |a-b| <= 1e-4 * max(1, |b|) -- SURVEY.md hard part 5)."""
import numpy as np
import pytest

from oracle import stt

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close(got, ref):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    return float(err.max())


def test_mel_filters_match_oracle(gpu):
    from open_speech_b200.features import B200FeatureExtractor

    for n in (80, 128):
        fe = B200FeatureExtractor(feature_size=n)
        assert fe.mel_filters.shape == (n, 201)
        assert np.abs(fe.mel_filters - stt.mel_filters(16000, 400, n)).max() < 1e-7
    assert fe.nb_max_frames == 3000 and fe.n_samples == 480000 and fe.time_per_frame == 0.01


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("seconds", [0.05, 1.0, 5.0, 30.0])
def test_logmel_vs_oracle(gpu, n_mels, seconds):
    from open_speech_b200 import synth
    from open_speech_b200.features import B200FeatureExtractor

    a = synth.clip_pcm16(seconds, seed=int(seconds * 100) + n_mels).astype(np.float32) / 32768.0
    got = B200FeatureExtractor(feature_size=n_mels)(a)
    ref = stt.logmel(a, n_mels)
    assert _close(got, ref) <= TOL, _close(got, ref)


def test_logmel_config1_shape_and_values(gpu):
    """BASELINE config 1: one 30 s clip, normalise + 128-bin log-mel -> f32[128, 3001]."""
    from open_speech_b200 import synth

    pcm = synth.clip_pcm16(30.0, seed=synth.SEED_C1)
    out = np.empty((128, 3001), np.float32)
    gpu.call("osb_logmel_host", gpu.ptr(pcm), gpu.FMT_PCM16, pcm.size, 128, gpu.ptr(out), 1, -18.0)
    ref = stt.stt_frontend(pcm, noise_reduce=False, normalize=True)
    assert ref.shape == (128, 3001)
    # the fused gain may move a handful of samples by 1 LSB (see test_gpu_stt_pre); the mel tolerance absorbs it
    assert _close(out, ref) <= TOL, _close(out, ref)


def test_logmel_edge_inputs(gpu):
    from open_speech_b200.features import B200FeatureExtractor

    fe = B200FeatureExtractor(feature_size=128)
    rng = np.random.default_rng(0)
    for n in (41, 159, 160, 161, 399, 400, 401, 4799, 5120, 5121, 16000 * 3 + 77):
        a = (rng.standard_normal(n) * 0.1).astype(np.float32)
        got, ref = fe(a), stt.logmel(a, 128)
        assert _close(got, ref) <= TOL, (n, _close(got, ref))
    z = fe(np.zeros(16000, np.float32))  # silence: everything clamps to log10(1e-10)
    assert np.abs(z - stt.logmel(np.zeros(16000, np.float32), 128)).max() <= 1e-6  # log10f(1e-10f) differs by 1 ulp
    # pure tone: bins far from the tone sit > 80 dB down and are clamped by the global max
    t = np.arange(32000) / 16000.0
    a = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    assert _close(fe(a), stt.logmel(a, 128)) <= TOL
    i16 = (a * 32767).astype(np.int16)
    assert _close(fe(i16), stt.logmel(i16.astype(np.float32) / 32768.0, 128)) <= TOL


def test_logmel_batch_fused_normalise(gpu):
    """Batch path with normalize_gain + requantisation fused into the sample staging.

    Per-stage parity: (1) the GPU's normalised int16 is within 1 LSB of the oracle's (the gain can differ by
    an ulp: numpy sums the squares pairwise in f32, the GPU exactly), (2) the fused log-mel equals the oracle
    log-mel OF THE GPU's int16 within 1e-4, (3) end to end, a 1-LSB change in a near-silent frame can move a
    cell by more than 1e-4: >= 99.9 % of cells are within 1e-4 and none is off by more than 2e-3.
    """
    import torch
    from open_speech_b200 import synth

    pcm = synth.clip_batch_pcm16(6, 4.0, seed=77, distinct=6)
    n = pcm.shape[1]
    x = torch.from_numpy(pcm).cuda()
    st = torch.cuda.current_stream().cuda_stream
    nf = gpu.lib().osb_logmel_frames(n)
    out = torch.empty((6, 128, nf), dtype=torch.float32, device="cuda")
    gpu.call("osb_logmel_dev", x.data_ptr(), gpu.FMT_PCM16, n, 6, n, 128, out.data_ptr(), 1, -18.0, st)
    q = torch.empty_like(x)
    gpu.call("osb_normalize_gain_pcm16_dev", x.data_ptr(), q.data_ptr(), n, 6, n, 1, -18.0, st)
    torch.cuda.synchronize()
    q = q.cpu().numpy()
    for i in range(6):
        a = pcm[i].astype(np.float32) / 32768.0
        ref_q = stt.quantise_pcm16(stt.normalize_gain(a))
        assert np.abs(q[i].astype(np.int32) - ref_q.astype(np.int32)).max() <= 1
        got = out[i].cpu().numpy()
        assert _close(got, stt.logmel(q[i].astype(np.float32) / 32768.0, 128)) <= TOL
        ref = stt.stt_frontend(pcm[i], noise_reduce=False, normalize=True)
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        flips = float((q[i] != ref_q).mean())
        print(f"normalise-only chain clip {i}: int16 flips {flips:.5%}, cells beyond 1e-4 {(err > TOL).mean():.5%}, worst {err.max():.2e}")
        # the only difference to the oracle is the gain's last float32 bit (tree reduction vs numpy's pairwise sum): a sample whose
        # product sits within that bit of an integer lands on the neighbouring LSB: 0 to 0.11 % of the samples per clip on B200 (no flip at all
        # when the two gains agree), <= 0.05 % of the cells beyond 1e-4, worst cell 2.9e-4.
        assert flips <= 3e-3 and (err > TOL).mean() <= 1e-3 and err.max() <= 2e-3, (flips, (err > TOL).mean(), err.max())
