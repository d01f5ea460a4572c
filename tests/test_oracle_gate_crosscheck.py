"""CPU: the spectral-gate oracle (a restatement of noisereduce built from scipy primitives, PARITY UNPINNED) against an
independent implementation of the same algorithm that shares no primitive with it: torch.stft / torch.istft instead of
scipy.signal.stft / istft, explicit forward and backward one-pole loops instead of filtfilt, a direct 2-D convolution
(torch conv2d, zero padding) instead of fftconvolve.  This pins the restatement's framing, scaling, filter start-up and
'same' cropping -- not noisereduce's own semantics, which stay unpinned (no package, no reference test)."""
import numpy as np
import pytest
import torch

from oracle import stt


def _independent_gate_chunk(chunk: np.ndarray, sr: int) -> np.ndarray:
    n_fft, hop = 1024, 256
    x = torch.from_numpy(chunk.astype(np.float64))
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float64)
    # scipy.signal.stft(boundary='zeros', padded=False, scaling='spectrum'): centred frames over zero padding, / sum(window)
    S = torch.stft(x, n_fft, hop_length=hop, window=win, center=True, pad_mode="constant", return_complex=True) / win.sum()
    A = S.abs().numpy()
    t_frames = 2.0 * sr / hop
    b = (np.sqrt(1 + 4 * t_frames**2) - 1) / (2 * t_frames**2)
    # filtfilt([b], [1, b-1], padtype=None): y[n] = b x[n] + (1-b) y[n-1], started in the steady state of the first sample,
    # forward, then the same over the reversed forward output
    def one_pole(v):
        y = np.empty_like(v)
        prev = v[:, 0].copy()
        for k in range(v.shape[1]):
            prev = b * v[:, k] + (1.0 - b) * prev
            y[:, k] = prev
        return y
    fwd = one_pole(A)
    A_s = one_pole(fwd[:, ::-1])[:, ::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        M = 1.0 / (1.0 + np.exp(-((A - A_s) / A_s - 2.0) * 10.0))
    n_f, n_t = int(500 / (sr / (n_fft / 2))), int(50 / ((hop / sr) * 1000))
    tri = lambda n: np.concatenate([np.arange(1, n + 2) / (n + 1), np.arange(n, 0, -1) / (n + 1)])
    K = np.outer(tri(n_f), tri(n_t))
    K /= K.sum()
    Mt = torch.from_numpy(M)[None, None]
    Kt = torch.from_numpy(K[::-1, ::-1].copy())[None, None]  # convolution = correlation with the flipped kernel
    Ms = torch.nn.functional.conv2d(Mt, Kt, padding=(K.shape[0] // 2, K.shape[1] // 2))[0, 0]
    y = torch.istft(S * Ms * win.sum(), n_fft, hop_length=hop, window=win, center=True, length=len(chunk))
    return y.numpy()


@pytest.mark.parametrize("sr,n", [(16000, 256 * 90), (16000, 256 * 37), (8000, 256 * 64)])
def test_gate_chunk_matches_independent_implementation(sr, n):
    from open_speech_b200 import synth

    x = synth.speech_like(n, sr, seed=11, extra_noise_rms=0.01).astype(np.float64)
    ref = stt._gate_chunk(x, sr)
    got = _independent_gate_chunk(x, sr)
    assert ref.shape == got.shape
    assert np.abs(ref - got).max() <= 1e-9 * max(1.0, np.abs(x).max()), float(np.abs(ref - got).max())


def test_smoothing_filter_is_the_documented_triangle():
    f = stt._nr_smoothing_filter(16, 3)
    assert f.shape == (33, 7) and abs(f.sum() - 1.0) < 1e-15
    assert np.allclose(f[:, 3] / f[16, 3], np.concatenate([np.arange(1, 18), np.arange(16, 0, -1)]) / 17.0)
