"""Re-entrancy: the reference calls this path from the asyncio loop thread and from 4-worker thread pools at once
(src/streaming.py:50-52, src/realtime/server.py:33-35, src/main.py:796-813; SURVEY.md 8(b) "Threading").  Every host
entry uses a thread-local stream and workspace, and one VAD session (weights) is shared by per-stream SileroVAD states.
Twelve threads hammer different entry points concurrently; every result must equal the single-threaded one bit for bit."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_host_calls_are_independent(gpu):
    from open_speech_b200 import synth
    from open_speech_b200.audio import preprocessing as pre
    from open_speech_b200.effects.chain import apply_chain
    from open_speech_b200.features import B200FeatureExtractor
    from open_speech_b200.realtime.audio_buffer import decode_audio_to_pcm16
    from open_speech_b200.realtime.tts_out import audio_deltas
    from open_speech_b200.streaming import resample_pcm16
    from open_speech_b200.vad.silero import SileroVAD, VadSession

    session = VadSession()
    clip = synth.clip_pcm16(6.0, seed=77, extra_noise_rms=0.01)
    clip_f = clip.astype(np.float32) / 32768.0
    ulaw = synth.ulaw_streams(1, 50)[0].reshape(-1).tobytes()
    utt = synth.tts_utterance(2.0, seed=78)
    fe = B200FeatureExtractor(feature_size=128)
    fx = [{"type": "normalize"}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]

    def buffer_events():
        from open_speech_b200.realtime.audio_buffer import InputAudioBuffer

        b = InputAudioBuffer(vad=SileroVAD(session), silence_duration_ms=200)
        return repr([b.append(clip[i:i + 1600].tobytes()) for i in range(0, len(clip) - 1600, 1600)]) + repr(b._total_samples)

    def session_gate():
        from open_speech_b200.streaming import SessionGate

        g = SessionGate(16000, 300, vad=SileroVAD(session))
        return repr([g.process_chunk(clip[i:i + 1600].tobytes())[1] for i in range(0, len(clip) - 1600, 1600)])

    def full_chain():
        from open_speech_b200.batch import SttFull

        out = {"probs": np.empty((2, 187), np.float32), "segments": np.empty((2, 95, 2), np.int32), "counts": np.empty(2, np.int32),
               "mel": np.empty((2, 128, 601), np.float32)}
        SttFull(session, fmt="pcm16", from_rate=16000).run_host(np.stack([clip, clip[::-1]]).copy(), out)
        return out["mel"].tobytes() + out["probs"].tobytes() + out["counts"].tobytes()

    jobs = {
        "buffer": buffer_events,
        "session_gate": session_gate,
        "full": full_chain,
        "decode": lambda: decode_audio_to_pcm16(ulaw, "g711_ulaw", 16000),
        "poly": lambda: resample_pcm16(clip.tobytes(), 16000, 8000),
        "gate": lambda: pre.reduce_noise(clip_f, 16000).tobytes(),
        "gain": lambda: pre.normalize_gain(clip_f).tobytes(),
        "mel": lambda: fe(clip_f).tobytes(),
        "vad": lambda: repr(SileroVAD(session).get_speech_segments(clip.tobytes())),
        "fx": lambda: apply_chain(utt, 24000, fx).tobytes(),
        "deltas": lambda: "\n".join(audio_deltas([utt[:20000], utt[20000:]], "g711_ulaw")),
        "pitch": lambda: apply_chain(utt, 24000, [{"type": "pitch", "semitones": 3}]).tobytes(),
    }
    expected = {k: f() for k, f in jobs.items()}
    errors, barrier = [], threading.Barrier(len(jobs))

    def worker(name, fn):
        try:
            barrier.wait()
            for _ in range(6):
                if fn() != expected[name]:
                    errors.append(name)
        except Exception as e:  # noqa: BLE001
            errors.append(f"{name}: {e!r}")

    threads = [threading.Thread(target=worker, args=kv) for kv in jobs.items()]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.gpu
def test_pageable_host_buffers_go_through_the_staging_rings(gpu):
    """numpy (pageable) buffers on the batch host entries are staged through pinned rings by copy threads (csrc/host_stage.cu): enough
    groups for both rings to wrap (four groups, two input slots, three output slots), groups big enough for the thread pool (> 4 MiB),
    and the result must be the one pinned buffers give, bit for bit -- for both batch entries."""
    import torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import SttFull
    from open_speech_b200.vad.silero import VadSession

    batch, seconds = 20, 60.0
    pcm = synth.clip_batch_pcm16(batch, seconds, seed=77, distinct=4)
    n = pcm.shape[1]
    nf = gpu.lib().osb_logmel_frames(n)
    page_out = np.full((batch, 128, nf), np.nan, np.float32)
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, batch, n, 16000, 0, 1, 128, gpu.ptr(page_out))
    pin_in = torch.from_numpy(pcm).pin_memory()
    pin_out = torch.empty((batch, 128, nf), dtype=torch.float32).pin_memory()
    gpu.call("osb_stt_frontend_host", pin_in.data_ptr(), n, batch, n, 16000, 0, 1, 128, pin_out.data_ptr())
    assert np.array_equal(page_out, pin_out.numpy())
    # mixed: pinned in, pageable out and the other way round
    mixed = np.full((batch, 128, nf), np.nan, np.float32)
    gpu.call("osb_stt_frontend_host", pin_in.data_ptr(), n, batch, n, 16000, 0, 1, 128, gpu.ptr(mixed))
    assert np.array_equal(mixed, page_out)
    pin_out.zero_()
    gpu.call("osb_stt_frontend_host", gpu.ptr(pcm), n, batch, n, 16000, 0, 1, 128, pin_out.data_ptr())
    assert np.array_equal(pin_out.numpy(), page_out)

    # composed chain: mu-law wire bytes in pageable memory, every result in pageable memory
    from oracle import codec
    wire = np.stack([np.frombuffer(codec.lin2ulaw(synth.clip_pcm16(30.0, sr=8000, seed=80 + i).tobytes()), np.uint8) for i in range(4)])
    wire = np.tile(wire, (5, 1)).copy()
    full = SttFull(VadSession(), fmt="g711_ulaw", from_rate=8000, linear_chunk=160)
    n16 = 2 * wire.shape[1]
    nf2 = gpu.lib().osb_logmel_frames(n16)

    def outs(pinned):
        mk = (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()) if pinned else (lambda shape, dt: np.empty(shape, {torch.float32: np.float32, torch.int32: np.int32}[dt]))
        return {"probs": mk((20, n16 // 512), torch.float32), "segments": mk((20, n16 // 512 // 2 + 2, 2), torch.int32),
                "counts": mk((20,), torch.int32), "mel": mk((20, 128, nf2), torch.float32)}

    a, b = outs(False), outs(True)
    full.run_host(wire, a)
    full.run_host(torch.from_numpy(wire).pin_memory(), b)
    for k in ("probs", "counts", "mel"):
        assert np.array_equal(np.asarray(a[k]), b[k].numpy()), k
