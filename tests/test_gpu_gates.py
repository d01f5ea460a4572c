"""GPU parity: the batched realtime gate (InputAudioBuffer x S) and the streaming-session gate (StreamingSession._process_chunk x S).

State machines: bit-exact against traces of the reference's own classes (tests/golden/vad_state_machines.json,
tests/golden/stream_gate.json).  With the real network in the loop: against the oracle chain, except where a probability
sits within the 1e-3 budget of the threshold.
"""
import json
import os
from math import gcd

import numpy as np
import pytest

from oracle import codec, resample
from oracle import vad as ovad

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def session(gpu):
    from open_speech_b200.vad.silero import VadSession, random_init_weights

    return VadSession(random_init_weights(1002))


@pytest.fixture(scope="module")
def stream_golden():
    with open(os.path.join(GOLDEN, "stream_gate.json")) as f:
        return json.load(f)


def test_gate_state_machine_golden(gpu, golden_vad):
    """Scripted chunk probabilities through the device gate == the reference's InputAudioBuffer (three streams run the same script:
    the compact list must hold them in stream order)."""
    import torch

    from open_speech_b200.realtime.gate import RealtimeGate

    for c in golden_vad["input_buffer"]:
        S, n = 3, c["chunk_samples"]
        g = RealtimeGate(S, n, fmt="pcm16", from_rate=16000, threshold=c["threshold"], silence_duration_ms=c["silence_duration_ms"])
        wire = torch.zeros((S, n), dtype=torch.int16, device="cuda")
        got = []
        for i, p in enumerate(c["probs"]):
            g.tick(wire, probs=torch.full((S,), p, dtype=torch.float32, device="cuda"))
            ev = g.read_events()
            assert [e[0] for e in ev] == list(range(S))[: len(ev)] and len(ev) in (0, S)
            if ev:
                got.append([i, ev[0][1], ev[0][2]])
                assert all(e[1:] == ev[0][1:] for e in ev)
        assert got == c["events"]
        rec = g.records()
        assert int(rec["total_samples"][0]) == n * len(c["probs"])


def test_gate_many_streams_compaction_and_clear(gpu):
    """1500 streams (several compaction rounds, a partial last CTA) with independent random scripts, against the oracle's restatement
    of the machine; clear() on some streams in the middle resets their silence counters only."""
    import torch

    from open_speech_b200.realtime.gate import RealtimeGate

    S, n, ticks = 1500, 320, 60
    rng = np.random.default_rng(3)
    probs = (rng.uniform(0, 1, (ticks, S)) ** 2).astype(np.float32)
    g = RealtimeGate(S, n, fmt="pcm16", from_rate=16000, threshold=0.5, silence_duration_ms=60)
    wire = torch.zeros((S, n), dtype=torch.int16, device="cuda")
    cleared = [7, 300, 1499]
    got = [[] for _ in range(S)]
    for t in range(ticks):
        if t == 30:
            g.clear(cleared)
        g.tick(wire, probs=torch.from_numpy(probs[t]).cuda())
        ev = g.read_events()
        assert [e[0] for e in ev] == sorted(e[0] for e in ev)
        for s, kind, ms in ev:
            got[s].append((t, kind, ms))
    for s in range(S):
        if s in cleared:
            continue
        want = ovad.input_buffer_events(probs[:, s], [n] * ticks, 0.5, 60)
        assert got[s] == want, s
    for s in cleared:  # same machine with the silence counter zeroed before tick 30
        ev, total, in_speech, sil = [], 0, False, 0
        for t in range(ticks):
            if t == 30:
                sil = 0
            cur = total * 1000 // 16000
            total += n
            if probs[t, s] >= np.float32(0.5):
                sil = 0
                if not in_speech:
                    in_speech = True
                    ev.append((t, "speech_started", cur))
            elif in_speech:
                sil += n
                if sil * 1000 // 16000 >= 60:
                    in_speech, sil = False, 0
                    ev.append((t, "speech_stopped", cur))
        assert got[s] == ev, s


@pytest.mark.parametrize("poly", [False, True])
def test_gate_real_vad_ulaw_40ms(gpu, session, poly):
    """mu-law 8 kHz, 40 ms chunks (one VAD window per chunk): decode + resample bit-exact, events as the oracle chain gives them."""
    import torch

    from open_speech_b200 import synth
    from open_speech_b200.realtime.gate import RealtimeGate

    S, chunk, ticks = 5, 320, 150
    x8 = [synth.clip_pcm16(chunk * ticks / 8000.0, sr=8000, seed=40 + s) for s in range(S)]
    ul = np.stack([np.frombuffer(codec.lin2ulaw(x.tobytes()), np.uint8) for x in x8])
    g = RealtimeGate(S, chunk, fmt="g711_ulaw", session=session, threshold=0.5, silence_duration_ms=200, arena_samples=640 * ticks, poly=poly)
    assert g.n_out == 640
    got = [[] for _ in range(S)]
    for t in range(ticks):
        g.tick(torch.from_numpy(np.ascontiguousarray(ul[:, t * chunk:(t + 1) * chunk])).cuda())
        for s, kind, ms in g.read_events():
            got[s].append((t, kind, ms))
    torch.cuda.synchronize()
    arena = g.arena.cpu().numpy()
    net = ovad.SileroNet()
    for s in range(S):
        pcm, st, cp = [], None, []
        for t in range(ticks):
            b = ul[s, t * chunk:(t + 1) * chunk].tobytes()
            c16 = resample.resample_pcm16(codec.ulaw2lin(b), 8000, 16000) if poly else codec.decode_audio_to_pcm16(b, "g711_ulaw", 16000)
            pcm.append(np.frombuffer(c16, np.int16))
            p, st = net.score_stream(pcm[-1].astype(np.float32) / 32768.0, st)
            cp.append(float(p.max()) if len(p) else 0.0)
        assert np.array_equal(arena[s], np.concatenate(pcm)), s          # arena == every decoded chunk, in order
        want = ovad.input_buffer_events(cp, [640] * ticks, 0.5, 200)
        if all(abs(p - 0.5) > PROB_TOL for p in cp):
            assert got[s] == want, s
    assert any(got[s] for s in range(S))
    rec = g.records()
    assert np.all(rec["buffered_samples"] == 640 * ticks) and np.all(rec["total_samples"] == 640 * ticks)
    out = g.commit(2)
    assert out.shape[0] == 640 * ticks and np.array_equal(out.cpu().numpy(), arena[2])
    assert int(g.records()["buffered_samples"][2]) == 0


def test_gate_reference_exact_20ms_is_one_gate_launch(gpu, session):
    """The reference's 20 ms case: 320 samples hold no VAD window, the probability is 0.0, nothing ever starts; the tick is the
    resample launch plus ONE gate launch (the last CTA compacts the events)."""
    import torch

    from open_speech_b200 import synth
    from open_speech_b200.realtime.gate import RealtimeGate

    S = 1024
    data = synth.ulaw_streams(S, 8)
    g = RealtimeGate(S, 160, fmt="g711_ulaw", session=session, arena_samples=320 * 8)
    ref = [np.frombuffer(codec.decode_audio_to_pcm16(data[t, 3].tobytes(), "g711_ulaw", 16000), np.int16) for t in range(8)]
    for t in range(8):
        l0 = gpu.lib().osb_launch_count()
        g.tick(torch.from_numpy(data[t]).cuda())
        assert gpu.lib().osb_launch_count() - l0 == 2
        assert g.read_events() == []
        assert np.array_equal(g.pcm[3].cpu().numpy(), ref[t])
    rec = g.records()
    assert np.all(rec["total_samples"] == 320 * 8) and np.all(rec["in_speech"] == 0)
    assert np.array_equal(g.arena[3].cpu().numpy(), np.concatenate(ref))


def test_gate_tick_cuda_graph_matches_stream_path(gpu, session):
    """The captured tick (one graph launch: H2D, resample, gate, D2H) leaves the same pcm16, records and events as the stream path."""
    import torch

    from open_speech_b200 import synth
    from open_speech_b200.realtime.gate import RealtimeGate

    S = 64
    data = synth.ulaw_streams(S, 12)
    a = RealtimeGate(S, 160, fmt="g711_ulaw", session=session, arena_samples=320 * 12)
    b = RealtimeGate(S, 160, fmt="g711_ulaw", session=session, arena_samples=320 * 12)
    gb = b.capture()
    for t in range(12):
        a.tick(torch.from_numpy(data[t]).cuda())
        ev_a = a.read_events()
        gb.host_in.copy_(torch.from_numpy(data[t]))
        ev_b = gb.run()
        assert ev_a == ev_b
        assert torch.equal(a.pcm.cpu(), gb.host_pcm)
    assert np.array_equal(a.records(), b.records()) and torch.equal(a.arena, b.arena)
    assert int(a.records()["total_samples"][0]) == 320 * 12


def test_gate_host_io_matches_stream_path(gpu, session):
    """A host_io gate (wire bytes read from, pcm16 + events written to pinned host memory by the kernels themselves: no copies) leaves
    the same pcm16, records, arena and events as the device-buffer gate -- scripted probabilities, then the real network at 40 ms."""
    import torch

    from open_speech_b200 import synth
    from open_speech_b200.realtime.gate import RealtimeGate

    S, T = 96, 40
    data = synth.ulaw_streams(S, T)
    rng = np.random.default_rng(7)
    probs = (rng.random((T, S)) < 0.3).astype(np.float32) * 0.9
    a = RealtimeGate(S, 160, fmt="g711_ulaw", session=None, silence_duration_ms=60, arena_samples=320 * T)
    b = RealtimeGate(S, 160, fmt="g711_ulaw", session=None, silence_duration_ms=60, arena_samples=320 * T, host_io=True)
    assert b.pcm.is_pinned() and b.events.is_pinned() and not b.pcm.is_cuda
    n_ev = 0
    for t in range(T):
        p = torch.from_numpy(probs[t]).cuda()
        a.tick(torch.from_numpy(data[t]).cuda(), p)
        ev_a = a.read_events()
        l0 = gpu.lib().osb_launch_count()
        ev_b = b.tick_host(torch.from_numpy(data[t]).pin_memory(), p)
        assert gpu.lib().osb_launch_count() - l0 == 2
        assert ev_a == ev_b
        n_ev += len(ev_b)
        assert torch.equal(a.pcm.cpu(), b.pcm)
        assert np.array_equal(b.pcm[5].numpy(), np.frombuffer(codec.decode_audio_to_pcm16(data[t, 5].tobytes(), "g711_ulaw", 16000), np.int16))
    assert n_ev > S
    assert np.array_equal(a.records(), b.records()) and torch.equal(a.arena, b.arena)
    with pytest.raises(RuntimeError):
        a.tick_host(torch.from_numpy(data[0]).pin_memory())
    with pytest.raises(RuntimeError):
        b.capture()
    with pytest.raises(ValueError):
        a.tick(torch.from_numpy(data[0]).pin_memory())          # a device-buffer gate takes device tensors only
    # the network in the loop: one VAD window per 40 ms chunk, scored from the pinned pcm16
    S2, chunk, T2 = 4, 320, 60
    x8 = [synth.clip_pcm16(chunk * T2 / 8000.0, sr=8000, seed=90 + s) for s in range(S2)]
    ul = np.stack([np.frombuffer(codec.lin2ulaw(x.tobytes()), np.uint8) for x in x8])
    c = RealtimeGate(S2, chunk, fmt="g711_ulaw", session=session, silence_duration_ms=200)
    d = RealtimeGate(S2, chunk, fmt="g711_ulaw", session=session, silence_duration_ms=200, host_io=True)
    seen = 0
    for t in range(T2):
        w = np.ascontiguousarray(ul[:, t * chunk:(t + 1) * chunk])
        c.tick(torch.from_numpy(w).cuda())
        ev_c = c.read_events()
        ev_d = d.tick_host(torch.from_numpy(w).pin_memory())
        assert ev_c == ev_d
        seen += len(ev_d)
        assert torch.equal(c.pcm.cpu(), d.pcm)
    assert seen > 0 and np.array_equal(c.records(), d.records())


def test_gate_buffer_errors(gpu):
    """The two BufferError cases of InputAudioBuffer.append against the arena capacity (audio_buffer.py:118-122)."""
    import torch

    from open_speech_b200.realtime.gate import RealtimeGate

    g = RealtimeGate(2, 400, fmt="pcm16", from_rate=16000, arena_samples=1000)
    wire = torch.ones((2, 400), dtype=torch.int16, device="cuda")
    for _ in range(2):
        g.tick(wire)
        assert g.read_events() == []
    g.tick(wire)  # 800 + 400 > 1000: refused, nothing changes
    assert g.read_events() == [(0, "buffer_full", 0), (1, "buffer_full", 0)]
    rec = g.records()
    assert np.all(rec["buffered_samples"] == 800) and np.all(rec["total_samples"] == 800)
    g.clear([1])
    g.tick(wire)
    assert g.read_events() == [(0, "buffer_full", 0)]
    assert g.records()["buffered_samples"].tolist() == [800, 400]
    big = RealtimeGate(1, 400, fmt="pcm16", from_rate=16000, arena_samples=300)
    big.state[0, 2] = 100
    big.tick(wire[:1].contiguous())  # a frame larger than the whole buffer clears it
    assert big.read_events() == [(0, "frame_too_large", 0)] and int(big.records()["buffered_samples"][0]) == 0


def test_input_audio_buffer_dropin_with_real_vad(gpu, session):
    """The per-connection drop-in class (one osb_gate_append_host per append) against the oracle chain, state carried in the object."""
    from open_speech_b200 import synth
    from open_speech_b200.realtime.audio_buffer import InputAudioBuffer
    from open_speech_b200.vad.silero import SileroVAD

    pcm = synth.clip_pcm16(12.0, seed=1004)
    for chunk in (640, 1600, 320):
        buf = InputAudioBuffer(vad=SileroVAD(session), threshold=0.5, silence_duration_ms=300)
        net, st, cp, got = ovad.SileroNet(), None, [], []
        n_chunks = len(pcm) // chunk
        for i in range(n_chunks):
            c = pcm[i * chunk:(i + 1) * chunk]
            for e in buf.append(c.tobytes()):
                got.append((i, e["type"], e.get("audio_start_ms", e.get("audio_end_ms"))))
            p, st = net.score_stream(c.astype(np.float32) / 32768.0, st)
            cp.append(float(p.max()) if len(p) else 0.0)
        want = ovad.input_buffer_events(cp, [chunk] * n_chunks, 0.5, 300)
        if all(abs(p - 0.5) > PROB_TOL for p in cp):
            assert got == want, chunk
        assert buf._total_samples == n_chunks * chunk and buf.get_audio() == pcm[: n_chunks * chunk].tobytes()
        if chunk == 320:
            assert got == []  # no full window in 20 ms: the reference's VAD returns 0.0
        else:
            assert got
    with pytest.raises(ValueError):
        InputAudioBuffer(vad=SileroVAD(session)).append(b"\x00" * 641)


def _n16(n, sr):
    if sr == 16000:
        return n
    g = gcd(16000, sr)
    return (n * (16000 // g) + sr // g - 1) // (sr // g)


def test_stream_gate_golden(gpu, stream_golden):
    """Scripted probabilities through the device machine == traces of the reference's StreamingSession._process_chunk."""
    import torch

    from open_speech_b200.realtime.gate import StreamGate

    for c in stream_golden["cases"]:
        S, n, sr = 2, c["chunk_samples"], c["sample_rate"]
        g = StreamGate(S, n, sr, threshold=c["threshold"], endpointing_ms=c["endpointing_ms"])
        assert g.max_utt_bytes == stream_golden["max_utterance_bytes"] and g.n_out == _n16(n, sr)
        chunk = torch.from_numpy(np.tile((np.arange(n) % 100).astype(np.int16), (S, 1))).cuda()
        want_pcm = np.frombuffer(resample.resample_pcm16(chunk[0].cpu().numpy().tobytes(), sr, 16000), np.int16)
        cols = c["steps"]
        for i, p in enumerate(c["probs"]):
            act = g.tick(chunk, probs=torch.full((S,), p, dtype=torch.float32, device="cuda"), vad_enabled=c["vad_enabled"])
            a = act.cpu().numpy()
            rec = g.records()
            assert a[0] == a[1]
            got = (int(rec["speech_active"][0]), int(rec["silence_samples"][0]), int(rec["utterance_bytes"][0]), int(bool(a[0] & 1)),
                   int(bool(a[0] & 32)), int(bool(a[0] & 24)), int(bool(a[0] & 16)))
            want = (cols["speech_active"][i], cols["silence_samples"][i], cols["utterance_bytes"][i], cols["speech_start"][i],
                    cols["speech_end"][i], cols["transcribe_calls"][i], cols["final"][i])
            assert got == want, (c["name"], i)
        assert np.array_equal(g.pcm[1].cpu().numpy(), want_pcm), c["name"]  # the resampled chunk is resample_pcm16's


def test_session_gate_dropin_with_real_vad(gpu, session):
    """SessionGate (one osb_stream_chunk_host per chunk) on 48 kHz audio: resample bit-exact, machine == oracle on the oracle's probabilities."""
    from open_speech_b200 import synth
    from open_speech_b200.streaming import SessionGate
    from open_speech_b200.vad.silero import SileroVAD

    x16 = synth.clip_pcm16(10.0, seed=1004)
    x48 = np.frombuffer(resample.resample_pcm16(x16.tobytes(), 16000, 48000), np.int16)
    chunk = 4800  # 100 ms at 48 kHz -> 1600 samples at 16 kHz, three VAD windows
    sg = SessionGate(48000, endpointing_ms=300, vad=SileroVAD(session), threshold=0.5)
    net, st, cp, acts = ovad.SileroNet(), None, [], []
    for i in range(len(x48) // chunk):
        c = x48[i * chunk:(i + 1) * chunk].tobytes()
        out, a = sg.process_chunk(c)
        want16 = resample.resample_pcm16(c, 48000, 16000)
        assert out == want16
        p, st = net.score_stream(np.frombuffer(want16, np.int16).astype(np.float32) / 32768.0, st)
        cp.append(float(p.max()))
        acts.append(a)
    want = ovad.stream_gate_steps(cp, 1600, threshold=0.5, endpointing_samples=4800)
    if all(abs(p - 0.5) > PROB_TOL for p in cp):
        assert acts == [w[0] for w in want]
        assert (sg.speech_active, sg.silence_samples, sg.utterance_bytes) == (want[-1][1], want[-1][2], want[-1][3])
    assert any(a & 1 for a in acts) and any(a & 16 for a in acts)
    off = SessionGate(16000, endpointing_ms=300, vad=None, vad_enabled=False)
    out, a = off.process_chunk(x16[:1600].tobytes())
    assert out == x16[:1600].tobytes() and a == (2 | 4 | 8) and off.speech_active
