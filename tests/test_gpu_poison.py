"""No kernel may depend on what a previous kernel left in shared memory.

compute-sanitizer is not available on the GPU pool, so this is the stand-in for its initcheck: osb_debug_poison_smem fills
the shared memory of every SM with NaNs, then each entry point runs again and must return exactly what it returned on a
quiet device.  (A log-mel tile row that was only ever multiplied by zero-padded taps was found this way: 0 * NaN.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_results_do_not_depend_on_stale_shared_memory(gpu):
    from open_speech_b200 import synth
    from open_speech_b200.audio import postprocessing as post
    from open_speech_b200.audio import preprocessing as pre
    from open_speech_b200.effects.chain import apply_chain
    from open_speech_b200.features import B200FeatureExtractor
    from open_speech_b200.realtime.audio_buffer import decode_audio_to_pcm16, encode_pcm16_to_format
    from open_speech_b200.streaming import resample_pcm16
    from open_speech_b200.vad.silero import SileroVAD, VadSession

    session = VadSession()
    clip = synth.clip_pcm16(41.0, seed=91, extra_noise_rms=0.01)  # two spectral-gate chunks
    short = clip[: 16000 * 6]
    f = short.astype(np.float32) / 32768.0
    ulaw = synth.ulaw_streams(1, 50)[0].reshape(-1).tobytes()
    utt = synth.tts_utterance(2.5, seed=92)
    n = len(short)
    nf = gpu.lib().osb_logmel_frames(n)
    pcm2 = np.stack([short, short[::-1].copy()])

    def frontend(nr, norm):
        mel = np.empty((2, 128, nf), np.float32)
        gpu.call("osb_stt_frontend_host", gpu.ptr(pcm2), n, 2, n, 16000, nr, norm, 128, gpu.ptr(mel))
        return mel.tobytes()

    jobs = {
        "decode": lambda: decode_audio_to_pcm16(ulaw, "g711_ulaw", 16000),
        "encode": lambda: encode_pcm16_to_format(short.tobytes(), 16000, "g711_alaw"),
        "poly": lambda: resample_pcm16(short.tobytes(), 16000, 8000),
        "poly441": lambda: resample_pcm16(short[:20000].tobytes(), 44100, 16000),
        "gate": lambda: pre.reduce_noise(clip.astype(np.float32) / 32768.0, 16000).tobytes(),
        "gate8k": lambda: pre.reduce_noise(f, 8000).tobytes(),
        "gain": lambda: pre.normalize_gain(f).tobytes(),
        "mel128": lambda: B200FeatureExtractor(feature_size=128)(f).tobytes(),
        "mel80": lambda: B200FeatureExtractor(feature_size=80)(short).tobytes(),
        "front11": lambda: frontend(1, 1),
        "front01": lambda: frontend(0, 1),
        "vad": lambda: SileroVAD(session)._score(short.tobytes(), gpu.FMT_PCM16, n).tobytes(),
        "post": lambda: post.normalize_output(post.trim_silence(utt)).tobytes(),
        "fx": lambda: apply_chain(utt, 24000, [{"type": "normalize"}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]).tobytes(),
        "pitch": lambda: apply_chain(utt, 24000, [{"type": "pitch", "semitones": -2}]).tobytes(),
    }
    bad = []
    for name, fn in jobs.items():
        ref = fn()
        gpu.call("osb_debug_poison_smem")
        if fn() != ref:
            bad.append(name)
    assert not bad, bad
