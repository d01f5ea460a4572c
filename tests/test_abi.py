"""CPU: libosb200 builds, loads, exports every symbol include/osb200.h declares, and has no CPU path."""
import ctypes
import os

import numpy as np
import pytest

from open_speech_b200 import _native as N


def test_library_exists_and_loads():
    assert os.path.exists(N.LIB_PATH), "run `python -m open_speech_b200.build` first"
    L = N.lib()
    assert L.osb_version() == 100


def test_every_declared_symbol_is_exported():
    decl = N.parse_header()
    assert len(decl) >= 20
    L = ctypes.CDLL(N.LIB_PATH)
    missing = [name for name in decl if not hasattr(L, name)]
    assert not missing, missing


def test_no_undeclared_public_symbols():
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("osb_")}
    assert exported == set(N.parse_header()), exported ^ set(N.parse_header())


def test_poly_filter_design_matches_scipy_bit_for_bit():
    """Host-side filter design (no GPU needed) == scipy firwin -> f32 -> *up."""
    from oracle import resample as R

    L = N.lib()
    for up, down in [(2, 1), (2, 3), (1, 3), (160, 441), (1, 2), (3, 1), (320, 441), (147, 160), (16, 11)]:
        n = ctypes.c_int(0)
        N.check(L.osb_resample_poly_taps(up, down, None, 0, ctypes.byref(n)))
        t = np.zeros(n.value, np.float32)
        N.check(L.osb_resample_poly_taps(up, down, N.ptr(t), n.value, ctypes.byref(n)))
        h, _, _ = R.design(up, down)
        assert np.array_equal(h[len(h) - n.value:], t), (up, down)
    with pytest.raises(ValueError):
        N.check(L.osb_resample_poly_taps(1, 1, None, 0, ctypes.byref(n)))


@pytest.mark.skipif(N.lib().osb_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_gpu():
    """Without a device the product raises; it never computes on the CPU."""
    from open_speech_b200.realtime import audio_buffer as ab
    from open_speech_b200.streaming import resample_pcm16

    with pytest.raises(RuntimeError):
        ab.decode_audio_to_pcm16(bytes(160), "g711_ulaw", 16000)
    with pytest.raises(RuntimeError):
        resample_pcm16(np.zeros(100, np.int16).tobytes(), 8000, 16000)
    with pytest.raises(RuntimeError):
        N.require_gpu()


def test_host_logic_without_gpu():
    """Pure host logic of the drop-in modules (no compute call)."""
    from open_speech_b200.realtime import audio_buffer as ab
    from open_speech_b200.streaming import resample_pcm16

    pcm = np.array([100, -200, 300], dtype=np.int16).tobytes()
    assert resample_pcm16(pcm, 16000, 16000) is pcm
    assert resample_pcm16(b"", 16000, 48000) == b""
    assert np.frombuffer(resample_pcm16(np.array([1000], np.int16).tobytes(), 16000, 32000), np.int16).tolist() == [1000, 1000]
    assert ab._resample_linear(pcm, 8000, 8000) is pcm
    with pytest.raises(ValueError, match="Unsupported"):
        ab.decode_audio_to_pcm16(b"\x00" * 100, "mp3")
    with pytest.raises(ValueError, match="Unsupported"):
        ab.encode_pcm16_to_format(b"\x00" * 100, 16000, "mp3")
    buf = ab.InputAudioBuffer(max_buffer_bytes=1000)
    with pytest.raises(BufferError):
        buf.append(b"\x00" * 2000)
    buf.append(b"\x00" * 800)
    with pytest.raises(BufferError):
        buf.append(b"\x00" * 400)
    assert buf.commit() == b"\x00" * 800 and buf.commit() == b""


def test_input_buffer_state_machine_golden(golden_vad):
    """InputAudioBuffer gate == the reference's, driven by a scripted VAD like tests/test_realtime.py."""
    from open_speech_b200.realtime.audio_buffer import InputAudioBuffer

    for c in golden_vad["input_buffer"]:
        probs = c["probs"]

        class V:
            i = 0

            def __call__(self, audio):
                p = probs[self.i % len(probs)]
                self.i += 1
                return p

        b = InputAudioBuffer(vad=V(), threshold=c["threshold"], silence_duration_ms=c["silence_duration_ms"])
        ev = []
        for i in range(len(probs)):
            for e in b.append(np.zeros(c["chunk_samples"], np.int16).tobytes()):
                ev.append([i, e["type"], e.get("audio_start_ms", e.get("audio_end_ms"))])
        assert ev == c["events"]


def test_vad_wrapper_with_scripted_session_golden(golden_vad):
    """SileroVAD framing/max/segmenter with the reference's mock-session pattern (tests/test_vad.py)."""
    from open_speech_b200.vad.silero import SileroVAD

    class Seq:
        def __init__(self, probs):
            self.probs, self.idx = probs, 0

        def run(self, _n, inputs):
            p = self.probs[self.idx % len(self.probs)]
            self.idx += 1
            return [np.array([[p]], np.float32), inputs["state"]]

    for c in golden_vad["segments"]:
        v = SileroVAD(Seq(c["probs"]), threshold=c["threshold"])
        segs = v.get_speech_segments(np.zeros(c["n_samples"], np.int16).tobytes(), min_speech_ms=c["min_speech_ms"], silence_ms=c["silence_ms"])
        assert [[s.start_ms, s.end_ms] for s in segs] == c["segments"]
    v = SileroVAD(Seq([0.1, 0.5, 0.3]))
    assert v(np.zeros(1536, np.float32)) == pytest.approx(0.5)
    assert v(np.zeros(100, np.float32)) == 0.0 and v(np.array([], np.float32)) == 0.0
    assert SileroVAD(Seq([0.5])).is_speech(np.zeros(512, np.int16).tobytes()) is True
    assert SileroVAD(Seq([0.9])).is_speech(b"") is False
    v._state = np.ones((2, 1, 128), np.float32)
    v.reset()
    assert np.all(v._state == 0)


def test_silero_state_dict_mapping():
    """weights_from_state_dict: the silero-vad v5 key names (with and without the TorchScript "_model." prefix), the
    stored shapes ([258,1,256] basis, [1,128,1] output conv), the 8 kHz branch ignored, wrong architectures rejected."""
    import torch

    from open_speech_b200.vad import silero as S

    ref = S.random_init_weights(7)
    inv = {v: k for k, v in S._SILERO_KEYS.items()}
    sd = {}
    for name, arr in ref.items():
        t = torch.from_numpy(arr.copy())
        if name == "stft_basis":
            t = t.reshape(258, 1, 256)
        if name == "dec.weight":
            t = t.reshape(1, 128, 1)
        sd["_model." + inv[name]] = t
        sd["_model_8k." + inv[name]] = torch.zeros(3)  # the other branch must not be picked up
    got = S.weights_from_state_dict(sd)
    assert set(got) == set(ref) and all(np.array_equal(got[k], ref[k]) for k in ref)
    assert np.array_equal(S.pack_weights(got), S.pack_weights(ref))
    plain = {k[len("_model."):]: v for k, v in sd.items() if k.startswith("_model.")}
    assert all(np.array_equal(S.weights_from_state_dict(plain)[k], ref[k]) for k in ref)
    bad = dict(sd)
    bad["_model.decoder.rnn.weight_hh"] = torch.zeros(256, 64)
    with pytest.raises(ValueError):
        S.weights_from_state_dict(bad)
    del sd["_model.encoder.2.reparam_conv.bias"]
    with pytest.raises(ValueError):
        S.weights_from_state_dict(sd)
    assert S.load_installed_silero_weights() is None  # the package is not installed here: random-init path (BASELINE config 2)
