"""Edge-size sweep of the denoise / front-end / pitch paths against the oracle (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from open_speech_b200 import synth
from open_speech_b200.audio import preprocessing as pre
from open_speech_b200.effects.chain import apply_chain
from oracle import stt, tts

rng = np.random.default_rng(3)
bad = 0
for n in (1, 2, 255, 256, 257, 511, 512, 1023, 1024, 1025, 4095, 30000, 30001, 59999):
    a = (0.1 * rng.standard_normal(n)).astype(np.float32)
    try:
        got, ref = pre.reduce_noise(a, 16000), stt.spectral_gate(a, 16000)
        ok = got.shape == ref.shape and (np.isnan(ref).any() or np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-6))
        print("gate", n, ok, float(np.abs(got - ref).max()) if n else 0, float(np.abs(ref).max()))
        bad += not ok
    except Exception as e:
        print("gate", n, "EXC", repr(e)[:120]); bad += 1
for n in (41, 100, 400, 1000, 16000):
    pcm = (3000 * rng.standard_normal((2, n))).astype(np.int16)
    from open_speech_b200 import _native as N
    nf = N.lib().osb_logmel_frames(n)
    for nr in (0, 1):
        for norm in (0, 1):
            mel = np.empty((2, 128, nf), np.float32)
            try:
                N.call("osb_stt_frontend_host", N.ptr(pcm), n, 2, n, 16000, nr, norm, 128, N.ptr(mel))
                ref = stt.stt_frontend(pcm[1], noise_reduce=bool(nr), normalize=bool(norm))
                err = np.abs(mel[1] - ref) / np.maximum(1.0, np.abs(ref))
                ok = (err <= 1e-4).mean() >= 0.97
                print("front", n, nr, norm, ok, float(err.max()))
                bad += not ok
            except Exception as e:
                print("front", n, nr, norm, "EXC", repr(e)[:160]); bad += 1
for fx in ([{"type": "pitch", "semitones": -5}, {"type": "reverb"}], [{"type": "normalize"}, {"type": "podcast_eq"}], [{"type": "normalize"}, {"type": "reverb", "room": "large"}, {"type": "robot"}],
           [{"type": "robot"}, {"type": "normalize"}, {"type": "podcast_eq"}, {"type": "robot"}], [{"type": "reverb"}, {"type": "robot"}, {"type": "pitch", "semitones": 2}]):
    for sr in (24000, 16000, 22050):
        x = synth.tts_utterance(1.7, seed=5)
        got, ref = apply_chain(x, sr, fx), tts.apply_chain(x, sr, fx)
        e = float(np.sqrt(np.mean((got.astype(np.float64) - ref) ** 2)) / np.sqrt(np.mean(ref.astype(np.float64) ** 2)))
        ok = e <= (1e-3 if any(f["type"] == "pitch" for f in fx) else 1e-5)
        print("fx", [f["type"] for f in fx], sr, ok, e)
        bad += not ok
print("BAD", bad)
