"""Error statistics of the GPU spectral gate against the float64 oracle (diagnostic, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from open_speech_b200 import synth
from open_speech_b200.audio import preprocessing as pre
from oracle import stt

for seconds, seed in ((5.0, 9), (12.0, 1004), (40.0, 9)):
    a = synth.clip_pcm16(seconds, seed=seed, extra_noise_rms=0.01).astype(np.float32) / 32768.0
    got = pre.reduce_noise(a, 16000).astype(np.float64)
    ref = stt.spectral_gate(a.astype(np.float64), 16000)
    peak = np.abs(ref).max()
    e = np.abs(got - ref)
    qg = stt.quantise_pcm16(stt.normalize_gain(got.astype(np.float32))).astype(np.int32)
    qr = stt.quantise_pcm16(stt.normalize_gain(ref.astype(np.float32))).astype(np.int32)
    print(f"{seconds}s: max {e.max()/peak:.3e} rms {np.sqrt((e**2).mean())/peak:.3e} of peak; LSB flips {np.mean(qg != qr):.4f} max {np.abs(qg-qr).max()}")
