"""GPU parity at BASELINE.json's FULL sizes, through size-independent properties.

The oracle cannot process 4 hours of audio in a test, so at full size the CUDA path is checked with properties that hold
at any size -- a clip inside the full batch equals the same clip processed alone (bit for bit), chunked scoring with
carried state equals one-shot scoring, the device segmenter equals the reference's state machine run on the same
probabilities, repeated inputs give repeated outputs -- plus a bounded oracle sample of the same full-size run.
"""
import numpy as np
import pytest

from oracle import stt
from oracle import tts as otts
from oracle import vad as ovad

pytestmark = pytest.mark.gpu


def test_config4_full_batch_256x60s(gpu):
    """configs[3]: 256 x 60 s clips, denoise + normalise + log-mel.  Every clip of the full batch == that clip alone;
    one clip against the oracle chain (stage-wise bars of test_stt_frontend_config4_chain)."""
    import torch
    from open_speech_b200 import synth

    distinct = 5
    pcm = synth.clip_batch_pcm16(256, 60.0, seed=synth.SEED_C4, distinct=distinct)
    n = pcm.shape[1]
    assert n == 960_000
    nf = gpu.lib().osb_logmel_frames(n)
    x = torch.from_numpy(pcm).cuda()
    out = torch.empty((256, 128, nf), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    gpu.call("osb_stt_frontend_dev", x.data_ptr(), n, 256, n, 16000, 1, 1, 128, out.data_ptr(), stream)
    alone = torch.empty((distinct, 128, nf), dtype=torch.float32, device="cuda")
    for i in range(distinct):
        gpu.call("osb_stt_frontend_dev", x[i].data_ptr(), n, 1, n, 16000, 1, 1, 128, alone[i].data_ptr(), stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    for i in range(256):  # position in the batch, neighbours and grouping are invisible
        assert torch.equal(out[i], alone[i % distinct]), i
    # the drop-in (host) entry gives the same features for a clip taken from the middle of the batch
    mel = np.empty((128, nf), np.float32)
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm[131]), n, 1, n, 16000, 1, 1, 128, gpu.ptr(mel))
    assert np.array_equal(mel, out[131].cpu().numpy())
    # bounded oracle sample: clip 3 (two noisereduce chunks: 960,000 > 600,000 samples)
    q = np.empty(n, np.int16)
    gpu.call("osb_preprocess_stt_host", gpu.ptr(pcm[3]), n, 1, 16000, 1, 1, -18.0, gpu.ptr(q))
    a = stt.normalize_gain(stt.spectral_gate(pcm[3].astype(np.float32) / 32768.0, 16000))
    dq = np.abs(q.astype(np.int32) - stt.quantise_pcm16(a).astype(np.int32))
    assert dq.max() <= 3 and (dq != 0).mean() <= 0.01, (int(dq.max()), float((dq != 0).mean()))
    got = out[3].cpu().numpy()
    ref_stage = stt.logmel(q.astype(np.float32) / 32768.0, 128)
    assert got.shape == ref_stage.shape == (128, 6001)
    assert (np.abs(got - ref_stage) / np.maximum(1.0, np.abs(ref_stage))).max() <= 1e-4


def test_config2_one_hour_stream(gpu):
    """configs[1]: one 1 h stream, 112,500 windows.  One-shot == three chunks with carried state; the device segmenter ==
    the reference's state machine on the same probabilities; the first minute against the oracle network."""
    import torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import VadBatch
    from open_speech_b200.vad.silero import VadSession, random_init_weights

    minute = [synth.clip_pcm16(60.0, seed=synth.SEED_C1 + 50 + i) for i in range(6)]
    pcm = np.concatenate([minute[i % 6] for i in range(60)])  # 1 h: the state makes every repetition a different stretch
    assert pcm.size == 57_600_000
    vb = VadBatch(VadSession(random_init_weights(1002)))
    x = torch.from_numpy(pcm).cuda()[None, :]
    probs, state = vb.score(x)
    segs, counts = vb.segments(probs, pcm.size)
    torch.cuda.synchronize()
    p = probs[0].cpu().numpy()
    assert p.shape == (112_500,) and np.isfinite(p).all() and p.min() >= 0.0 and p.max() <= 1.0
    # carried state: 20 min + 25 min + 15 min (window-aligned cuts)
    cuts = [0, 20 * 60 * 16000, 45 * 60 * 16000, pcm.size]
    st, parts = None, []
    for a, b in zip(cuts[:-1], cuts[1:]):
        pr, st = vb.score(x[:, a:b], st)
        parts.append(pr[0].cpu().numpy())
    assert np.array_equal(np.concatenate(parts), p) and torch.equal(st, state)
    # segmenter: integer state machine, bit-exact at full length
    ref = ovad.segments_from_probs(p, pcm.size, 0.5, 250, 800)
    k = int(counts[0].item())
    got = segs[0, :k].cpu().numpy()
    assert k == len(ref) and k >= 2, (k, len(ref), float((p >= 0.5).mean()))
    assert [(int(s), int(e)) for s, e in got] == [(r.start_ms, r.end_ms) for r in ref]
    assert all(got[i][0] < got[i][1] <= got[i + 1][0] for i in range(k - 1))  # ordered, disjoint
    # bounded oracle sample: the first minute of probabilities
    o, _ = ovad.SileroNet().score_stream(pcm[:960_000].astype(np.float32) / 32768.0)
    assert np.abs(p[:1875] - o).max() <= 1e-3


def test_config5_full_batch_4096_utterances(gpu):
    """configs[4]: 4096 ragged utterances through trim + peak normalise + [normalize, reverb, podcast_eq, robot] + int16.
    Repeated utterances give identical output wherever they sit in the flat buffer; three of them against the oracle."""
    from open_speech_b200 import synth
    from open_speech_b200.batch import TtsPost

    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
    distinct = 16
    utts = synth.tts_batch(4096, seed=synth.SEED_C5, distinct=distinct)
    pcm, lens = TtsPost(sample_rate=24000, effects=fx).run_numpy(utts)
    assert len(pcm) == 4096
    exact = 0
    for i in range(4096):
        assert lens[i] == lens[i % distinct], i
        d = np.abs(pcm[i].astype(np.int32) - pcm[i % distinct].astype(np.int32))
        assert d.max(initial=0) <= 1, (i, int(d.max()))  # the order of the RMS partial sums may move a gain by one ulp
        exact += int(d.max(initial=0) == 0)
    assert exact >= 4000, exact
    for i in (0, 7, 13):
        ref = otts.tts_chain([utts[i]], fx)
        assert len(pcm[i]) == len(ref) == lens[i]
        d = np.abs(pcm[i].astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 3, (i, int(d.max()))  # 1e-4 of full scale


def test_config3_sixty_seconds_of_ticks(gpu):
    """configs[2]: 1024 mu-law streams x 3,000 ticks of 20 ms.  Every tick of every stream through the batched entry ==
    the drop-in decode of that chunk (bit-exact); a sample of chunks against the oracle."""
    import torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import RealtimeTick
    from open_speech_b200.realtime.audio_buffer import decode_audio_to_pcm16
    from oracle import codec

    ticks = synth.ulaw_streams(1024, 3000)  # uint8 [3000, 1024, 160]; 16 distinct streams tiled
    rt = RealtimeTick(1024)
    d = torch.from_numpy(ticks).cuda()
    out = torch.empty((3000, 1024, 320), dtype=torch.int16, device="cuda")
    for t in range(3000):
        rt(d[t], out[t])
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[:, 16:32], got[:, :16]) and np.array_equal(got[:, 1008:], got[:, :16])  # tiled streams
    rng = np.random.default_rng(3)
    for t, s in zip(rng.integers(0, 3000, 40), rng.integers(0, 1024, 40)):
        chunk = ticks[t, s].tobytes()
        ref = codec.decode_audio_to_pcm16(chunk, "g711_ulaw", 16000)
        assert got[t, s].tobytes() == ref == decode_audio_to_pcm16(chunk, "g711_ulaw", 16000), (t, s)
