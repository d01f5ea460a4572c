"""GPU parity: Silero-shaped VAD scoring (<=1e-3 abs on probabilities) and bit-exact segmenting."""
import ctypes

import numpy as np
import pytest

from oracle import vad as ovad

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3  # north_star: 1e-3 absolute for VAD probabilities


@pytest.fixture(scope="module")
def session(gpu):
    from open_speech_b200.vad.silero import VadSession, random_init_weights

    w = random_init_weights(1002)
    ow = ovad.make_weights(1002)
    for k in w:
        assert np.array_equal(w[k], ow[k]), k  # product and oracle draw identical seeded weights
    return VadSession(w)


def _audio(seconds, seed):
    from open_speech_b200 import synth

    return synth.clip_pcm16(seconds, seed=seed)


def test_probs_match_oracle_60s(gpu, session):
    from open_speech_b200.vad.silero import SileroVAD

    pcm = _audio(60.0, 1002)
    ref, ref_state = ovad.SileroNet().score_stream(pcm.astype(np.float32) / 32768.0)
    v = SileroVAD(session)
    probs = v._score(pcm.tobytes(), gpu.FMT_PCM16, len(pcm))
    assert probs.shape == ref.shape == (1875,)
    assert np.abs(probs - ref).max() <= PROB_TOL, float(np.abs(probs - ref).max())
    assert np.abs(v._state - ref_state).max() <= 1e-3
    assert (ref >= 0.5).mean() > 0.2 and (ref < 0.5).mean() > 0.2  # the test exercises both sides of the threshold


def test_call_semantics_and_state_carry(gpu, session):
    from open_speech_b200.vad.silero import SileroVAD

    pcm = _audio(4.0, 7)
    a = pcm.astype(np.float32) / 32768.0
    net = ovad.SileroNet()
    v = SileroVAD(session)
    assert v(np.zeros(0, np.float32)) == 0.0 and v(a[:100]) == 0.0
    # chunked calls carry the LSTM state exactly like one long call; remainders (<512) are dropped per call
    st = None
    pos = 0
    for n in (1600, 640, 512, 5000, 333, 2048):
        chunk = a[pos:pos + n]
        pos += n
        ref_probs, st = net.score_stream(chunk, st)
        got = v(chunk)
        want = float(ref_probs.max()) if len(ref_probs) else 0.0
        assert abs(got - want) <= PROB_TOL
    v.reset()
    assert np.all(v._state == 0)
    # f32 and pcm16 entry points agree
    v1, v2 = SileroVAD(session), SileroVAD(session)
    assert abs(v1(a[:5120]) - v2.score_pcm16(pcm[:5120].tobytes())) <= 1e-6
    assert isinstance(v1.is_speech(pcm[:5120].tobytes()), bool) and v1.is_speech(b"") is False


def test_segments_end_to_end_and_bit_exact_machine(gpu, session):
    from open_speech_b200.vad.silero import SileroVAD

    pcm = _audio(60.0, 1004)
    ref_probs, _ = ovad.SileroNet().score_stream(pcm.astype(np.float32) / 32768.0)
    v = SileroVAD(session)
    got = v.get_speech_segments(pcm.tobytes())
    gp = SileroVAD(session)._score(pcm.tobytes(), gpu.FMT_PCM16, len(pcm))
    # (1) the GPU segmenter is bit-exact on the GPU's own probabilities
    want = ovad.segments_from_probs(gp, len(pcm))
    assert [(s.start_ms, s.end_ms) for s in got] == [(s.start_ms, s.end_ms) for s in want]
    # (2) end to end vs the oracle network: identical unless a probability sits within tol of the threshold
    near = int((np.abs(ref_probs - 0.5) <= PROB_TOL).sum())
    ref_segs = ovad.segments_from_probs(ref_probs, len(pcm))
    if near == 0:
        assert [(s.start_ms, s.end_ms) for s in got] == [(s.start_ms, s.end_ms) for s in ref_segs]
    assert len(got) >= 2
    assert v.get_speech_segments(b"") == []


def test_segmenter_kernel_golden_scripts(gpu, golden_vad):
    """The reference's scripted-probability cases (tests/test_vad.py:126-174 pattern), bit-exact on the GPU."""
    import torch

    for c in golden_vad["segments"]:
        probs = torch.tensor(c["probs"], dtype=torch.float32, device="cuda")
        segs = torch.zeros((64, 2), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        gpu.call("osb_vad_segments_dev", probs.data_ptr(), len(c["probs"]), len(c["probs"]), 1, c["n_samples"], float(c["threshold"]),
                 c["min_speech_ms"], c["silence_ms"], segs.data_ptr(), cnt.data_ptr(), 64, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        k = int(cnt.item())
        assert segs[:k].cpu().tolist() == c["segments"]


def test_batched_streams_match_single(gpu, session):
    import torch

    pcm = np.stack([_audio(8.0, 100 + i) for i in range(5)])
    x = torch.from_numpy(pcm).cuda()
    n = pcm.shape[1]
    n_win = n // 512
    state = torch.zeros((5, 2, 128), dtype=torch.float32, device="cuda")
    probs = torch.empty((5, n_win), dtype=torch.float32, device="cuda")
    gpu.call("osb_vad_score_dev", session.handle, x.data_ptr(), gpu.FMT_PCM16, n, 5, n, state.data_ptr(), probs.data_ptr(), n_win,
             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    net = ovad.SileroNet()
    for i in range(5):
        ref, _ = net.score_stream(pcm[i].astype(np.float32) / 32768.0)
        assert np.abs(probs[i].cpu().numpy() - ref).max() <= PROB_TOL


@pytest.mark.parametrize("batch", [149, 297, 301])
def test_many_streams_share_recurrence_ctas(gpu, session, batch):
    """More than 148 streams: the recurrence kernel walks 2 (149..296) or 4 (> 296) streams per CTA in lock step, the
    last CTA partly empty.  Every stream must equal the same audio scored alone (stream-to-CTA packing is invisible),
    carried state included."""
    import torch

    base = [_audio(2.0, 400 + i) for i in range(7)]
    pcm = np.stack([base[i % 7] for i in range(batch)])
    n = pcm.shape[1]
    n_win = n // 512
    x = torch.from_numpy(pcm).cuda()
    state = torch.zeros((batch, 2, 128), dtype=torch.float32, device="cuda")
    probs = torch.empty((batch, n_win), dtype=torch.float32, device="cuda")
    for half in range(2):  # two calls: the second starts from the state the first left
        gpu.call("osb_vad_score_dev", session.handle, x.data_ptr(), gpu.FMT_PCM16, n, batch, n, state.data_ptr(), probs.data_ptr(), n_win,
                 torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got, st = probs.cpu().numpy(), state.cpu().numpy()
    one_state = torch.zeros((7, 2, 128), dtype=torch.float32, device="cuda")
    one_probs = torch.empty((7, n_win), dtype=torch.float32, device="cuda")
    x7 = torch.from_numpy(np.stack(base)).cuda()
    for half in range(2):
        gpu.call("osb_vad_score_dev", session.handle, x7.data_ptr(), gpu.FMT_PCM16, n, 7, n, one_state.data_ptr(), one_probs.data_ptr(), n_win,
                 torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref, ref_st = one_probs.cpu().numpy(), one_state.cpu().numpy()
    for i in range(batch):
        assert np.array_equal(got[i], ref[i % 7]) and np.array_equal(st[i], ref_st[i % 7]), i
    net = ovad.SileroNet()
    o1, s1 = net.score_stream(base[0].astype(np.float32) / 32768.0)
    o2, _ = net.score_stream(base[0].astype(np.float32) / 32768.0, s1)
    assert np.abs(got[0] - o2).max() <= PROB_TOL


@pytest.mark.parametrize("batch", [40, 131, 256])
def test_pipelined_chunks_match_serial(gpu, session, batch, monkeypatch):
    """More than one chunk of windows: the fused front of chunk i + 1 runs on a side stream beside the recurrence of chunk i (double-
    buffered pre-activations, front grid capped to the SMs the recurrence leaves free; 256 streams -> four streams per recurrence CTA
    instead of two).  Same probabilities and carried state, bit for bit, as the serial schedule; ragged last chunk; two calls."""
    import torch

    monkeypatch.setenv("OSB_VAD_CHUNK_WAVES", "1")  # 148 x 128 windows per chunk
    secs = {40: 50.0, 131: 16.0, 256: 9.0}[batch]  # 4-5 chunks each, the last one short
    base = [_audio(secs, 700 + i) for i in range(5)]
    pcm = np.stack([base[i % 5] for i in range(batch)])
    n = pcm.shape[1]
    n_win = n // 512
    assert n_win * batch > 3 * 148 * 128
    x = torch.from_numpy(pcm).cuda()
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("OSB_VAD_PIPELINE", mode)
        state = torch.zeros((batch, 2, 128), dtype=torch.float32, device="cuda")
        probs = torch.full((batch, n_win), -1.0, dtype=torch.float32, device="cuda")
        for _ in range(2):
            gpu.call("osb_vad_score_dev", session.handle, x.data_ptr(), gpu.FMT_PCM16, n, batch, n, state.data_ptr(), probs.data_ptr(), n_win,
                     torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res[mode] = (probs.cpu().numpy(), state.cpu().numpy())
    assert np.array_equal(res["0"][0], res["1"][0]) and np.array_equal(res["0"][1], res["1"][1])
    net = ovad.SileroNet()
    o1, s1 = net.score_stream(base[1].astype(np.float32) / 32768.0)
    o2, _ = net.score_stream(base[1].astype(np.float32) / 32768.0, s1)
    assert np.abs(res["1"][0][1] - o2).max() <= PROB_TOL


@pytest.mark.parametrize("batch", [1, 8, 11])
def test_tensor_pipe_recurrence_matches_ffma_and_oracle(gpu, session, batch):
    """The shipped recurrence (mma.sync, W_hh as fp16 fragments in registers, h as fp16 hi + lo planes, eight streams per CTA) against the
    FP32 FFMA recurrence kernels and the oracle: 60 s per stream, state carried over a second call, partly filled CTAs."""
    import torch

    pcm = np.stack([_audio(60.0, 900 + i) for i in range(min(batch, 3))])
    pcm = np.stack([pcm[i % len(pcm)] for i in range(batch)])
    n = pcm.shape[1]
    n_win = n // 512
    x = torch.from_numpy(pcm).cuda()
    res = {}
    try:
        for mode in (0, 1, 2):
            gpu.call("osb_vad_set_recurrence", session.handle, mode)
            state = torch.zeros((batch, 2, 128), dtype=torch.float32, device="cuda")
            probs = torch.full((2, batch, n_win), -1.0, dtype=torch.float32, device="cuda")
            for k in range(2):
                gpu.call("osb_vad_score_dev", session.handle, x.data_ptr(), gpu.FMT_PCM16, n, batch, n, state.data_ptr(), probs[k].data_ptr(), n_win,
                         torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            res[mode] = (probs.cpu().numpy(), state.cpu().numpy())
    finally:
        gpu.call("osb_vad_set_recurrence", session.handle, 1)
    for mode in (1, 2):  # 1: h as fp16 hi + lo (default), 2: h as one fp16 plane
        d = float(np.abs(res[0][0] - res[mode][0]).max())
        print(f"tensor-pipe (mode {mode}) vs FFMA recurrence, batch {batch}: max |dp| = {d:.2e}, max |dstate| = {float(np.abs(res[0][1] - res[mode][1]).max()):.2e}")
        assert d <= 5e-4
    net = ovad.SileroNet()
    a = pcm[0].astype(np.float32) / 32768.0
    o1, s1 = net.score_stream(a)
    o2, _ = net.score_stream(a, s1)
    assert np.abs(res[1][0][0][0] - o1).max() <= PROB_TOL and np.abs(res[1][0][1][0] - o2).max() <= PROB_TOL
    for i in range(batch):  # the column a stream sits in is invisible
        assert np.array_equal(res[1][0][:, i], res[1][0][:, i % 3 if batch >= 3 else 0])


def test_tcgen05_fronts_match_ffma_and_oracle(gpu, session):
    """The fused persistent tcgen05 front (mode 2, the default) and the per-layer tcgen05 GEMMs (mode 1), both split-bf16, against
    the FP32 FFMA kernels (mode 0) and the oracle."""
    from open_speech_b200.vad.silero import SileroVAD

    pcm = _audio(20.0, 321)
    ref, _ = ovad.SileroNet().score_stream(pcm.astype(np.float32) / 32768.0)
    out = {}
    try:
        for mode in (2, 1, 0):
            gpu.call("osb_vad_set_gemm", session.handle, mode)
            out[mode] = SileroVAD(session)._score(pcm.tobytes(), gpu.FMT_PCM16, len(pcm))
    finally:
        gpu.call("osb_vad_set_gemm", session.handle, 2)
    for mode in (0, 1, 2):
        assert np.abs(out[mode] - ref).max() <= PROB_TOL, (mode, float(np.abs(out[mode] - ref).max()))
    assert np.abs(out[1] - out[0]).max() <= 2e-4, float(np.abs(out[1] - out[0]).max())
    assert np.abs(out[2] - out[0]).max() <= 2e-4, float(np.abs(out[2] - out[0]).max())


def test_fused_front_tiles_streams_and_float_input(gpu, session):
    """Tile edges of the fused front: window counts that are not multiples of 128, tiles that span several streams, more tiles than
    SMs (a CTA walks several tiles: every barrier phase wraps), float32 input, odd strides (unaligned rows)."""
    import torch

    net = ovad.SileroNet()
    for batch, secs, pad in ((3, 1.7, 0), (37, 2.1, 3), (300, 3.0, 8), (2, 700.0, 0)):
        base = [_audio(secs, 500 + i) for i in range(min(batch, 4))]
        n = len(base[0])
        pcm = np.zeros((batch, n + pad), np.int16)
        for i in range(batch):
            pcm[i, :n] = base[i % len(base)]
        n_win = n // 512
        for as_float in (False, True):
            x = torch.from_numpy(pcm.astype(np.float32) / 32768.0 if as_float else pcm).cuda()
            state = torch.zeros((batch, 2, 128), dtype=torch.float32, device="cuda")
            probs = torch.empty((batch, n_win), dtype=torch.float32, device="cuda")
            gpu.call("osb_vad_score_dev", session.handle, x.data_ptr(), gpu.FMT_F32 if as_float else gpu.FMT_PCM16, n, batch, n + pad,
                     state.data_ptr(), probs.data_ptr(), n_win, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = probs.cpu().numpy()
            for i in range(len(base)):
                ref, _ = net.score_stream(base[i].astype(np.float32) / 32768.0)
                assert np.abs(got[i] - ref).max() <= PROB_TOL, (batch, as_float, i, float(np.abs(got[i] - ref).max()))
            for i in range(len(base), batch):
                assert np.array_equal(got[i], got[i % len(base)]), (batch, i)  # the tile a window lands in is invisible
            if secs > 100:
                break  # the long case once
