"""GPU parity: spectral-gating noise reduction and the composed STT front-end (BASELINE config 4 chain).

Oracle = scipy restatement of noisereduce's non-stationary gate in float64 (PARITY UNPINNED by the reference:
noisereduce is an absent optional dependency).  Tolerance: 1e-4 relative to the clip's peak for the audio,
then the per-stage tolerances of test_gpu_logmel for what follows.
"""
import io
import wave

import json
import os

import numpy as np
import pytest

from oracle import stt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    records = []
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
        # (3) end to end.  The chain requantises to int16 between the denoiser and the log-mel (src/audio/preprocessing.py:23-25 truncates
        # x * 32767): a float32 denoiser lands on the other side of an integer boundary for a few samples in 10,000, and in gated
        # (near-silent) stretches, where the signal is a few LSB, one flipped sample moves every mel cell of the frames that contain it
        # by more than 1e-4.  This is MEASURED, not assumed: the same metric is taken for a CPU control -- the oracle's own recipe
        # carried out in float32 (complex64 scipy FFTs, oracle/stt.py spectral_gate(work_dtype=float32)) against the float64 oracle.
        # Measured on B200 (profiles/r02_parity_config4.json): GPU 1.1-1.6 % of the cells beyond 1e-4 (control 0.8-1.3 %), 0.05-0.08 % of
        # the int16 samples one LSB off (control 0.03-0.08 %), worst cell 1.5e-3 (control 1.4e-3): the same share, so the tolerance is a
        # property of the reference's int16 round trip, not of this implementation.  Asserted: never more than 1 LSB; per clip within 2x
        # of the measured rates; over the three clips no more than twice the control's misses and flips.
        ref = stt.stt_frontend(pcm[i], noise_reduce=True, normalize=True)
        ctl = stt.stt_frontend(pcm[i], noise_reduce=True, normalize=True, gate_dtype=np.float32)
        q_ref = stt.quantise_pcm16(a)
        q_ctl = stt.quantise_pcm16(stt.normalize_gain(stt.spectral_gate(pcm[i].astype(np.float32) / 32768.0, 16000, np.float32)))
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        err_c = np.abs(ctl - ref) / np.maximum(1.0, np.abs(ref))
        rec = {"clip": i, "gpu_cells_beyond_tol": float((err > TOL).mean()), "gpu_worst_cell": float(err.max()),
               "control_cells_beyond_tol": float((err_c > TOL).mean()), "control_worst_cell": float(err_c.max()),
               "gpu_int16_flips": float((q != q_ref).mean()), "control_int16_flips": float((q_ctl != q_ref).mean()),
               "gpu_max_lsb": int(dq.max()), "control_max_lsb": int(np.abs(q_ctl.astype(np.int32) - q_ref.astype(np.int32)).max())}
        records.append(rec)
        print("config-4 chain parity:", rec)
        assert rec["gpu_max_lsb"] <= 1 and rec["gpu_cells_beyond_tol"] <= 0.03 and rec["gpu_worst_cell"] <= 3e-3, rec
    # over the three clips together (per clip the flip counts are a few dozen samples: too few for a ratio)
    agg = {k: float(np.mean([r[k] for r in records])) for k in ("gpu_cells_beyond_tol", "control_cells_beyond_tol", "gpu_int16_flips", "control_int16_flips")}
    print("config-4 chain parity, mean of 3 clips:", agg)
    assert agg["gpu_cells_beyond_tol"] <= 2.0 * agg["control_cells_beyond_tol"], agg
    assert agg["gpu_int16_flips"] <= 2.0 * agg["control_int16_flips"], agg
    records.append(agg)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "parity_config4.json"), "w") as f:
            json.dump(records, f, indent=1)
    # host-pointer entry == device entry
    mel = np.empty((3, 128, nf), np.float32)
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, 3, n, 16000, 1, 1, 128, gpu.ptr(mel))
    assert np.array_equal(mel, out.cpu().numpy())
    # no denoise, no normalise: requantise-only chain is exact up to the mel tolerance
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, 3, n, 16000, 0, 0, 128, gpu.ptr(mel))
    ref = stt.stt_frontend(pcm[0], noise_reduce=False, normalize=False)
    assert (np.abs(mel[0] - ref) / np.maximum(1.0, np.abs(ref))).max() <= TOL
