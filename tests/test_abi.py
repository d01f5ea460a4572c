"""CPU: libosb200 builds, loads, exports every symbol include/osb200.h declares, and has no CPU path."""
import ctypes
import json
import os
import sys

import numpy as np
import pytest

from open_speech_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def test_product_has_no_host_vad_or_gate():
    """The window loop, the segmenter and the start / stop gate exist only in libosb200: the drop-in classes refuse the
    mock sessions / scripted callables that the reference's own tests drive its Python implementation with."""
    from open_speech_b200.realtime.audio_buffer import InputAudioBuffer
    from open_speech_b200.vad import silero as S

    class Seq:
        def run(self, _n, inputs):
            return [np.array([[0.9]], np.float32), inputs["state"]]

    with pytest.raises(TypeError, match="VadSession"):
        S.SileroVAD(Seq())
    with pytest.raises(TypeError, match="no host implementation"):
        InputAudioBuffer(vad=lambda audio: 0.9)
    for name in ("_score_mock", "_segments_from_probs"):
        assert not hasattr(S, name) and not hasattr(S.SileroVAD, name)
    # vad=None is storage + clock only (no compute): works without a device
    b = InputAudioBuffer()
    assert b.append(b"\x00" * 640) == [] and b._total_samples == 320 and b.in_speech is False
    assert b.commit() == b"\x00" * 640 and b.get_audio() == b""


def test_get_vad_model_refuses_random_weights(monkeypatch):
    """No silero-vad package here: the server singleton must fail loudly instead of gating speech with a random network."""
    import asyncio

    from open_speech_b200.vad import silero as S

    monkeypatch.delenv(S.ALLOW_RANDOM_INIT_ENV, raising=False)
    monkeypatch.setattr(S, "_vad_model", None)
    with pytest.raises(RuntimeError, match="not available"):
        asyncio.run(S.get_vad_model())
    assert S._vad_model is None


def test_voice_spec_parser_matches_reference():
    """Own parser (hand-written scan, batched operands) == the reference's regex parser on a fuzzed alphabet."""
    import random

    from open_speech_b200.tts import voices as ours

    fixed = {"af_bella": [("af_bella", 1.0)], "alloy": [("af_heart", 1.0)], "af_bella(2)+af_sky(1)": [("af_bella", 2.0), ("af_sky", 1.0)],
             " a (2)": None, "a(2.)": None, "a(.5)": None, "a(1.5)+ b": [("a", 1.5), ("b", 1.0)], "": None, "a++b": None,
             "alloy+echo": [("alloy", 1.0), ("echo", 1.0)], "alloy(1)": [("alloy", 1.0)], "a()": None}
    for spec, want in fixed.items():
        if want is None:
            with pytest.raises(ValueError, match="Invalid voice spec component"):
                ours.parse_voice_spec(spec)
        else:
            assert [tuple(c) for c in ours.parse_voice_spec(spec).components] == want
    assert ours.parse_voice_spec("a(0)+b(0)").normalized_weights() == [0.5, 0.5]
    assert ours.parse_voice_spec("a(3)+b(1)").normalized_weights() == [0.75, 0.25]
    idx, w = ours.blend_operands(["a(2)+b(1)", "c", "a+b+c"], {"a": 0, "b": 1, "c": 2})
    assert idx.tolist() == [[0, 1, -1], [2, -1, -1], [0, 1, 2]] and w.dtype == np.float32 and abs(w[0, 0] - 2 / 3) < 1e-7
    ref_dir = "/root/reference"
    if not os.path.isdir(ref_dir):
        return  # the fuzz against the reference's module runs in the build container only
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ref_voices", os.path.join(ref_dir, "src/tts/voices.py"))
    ref = importlib.util.module_from_spec(spec)
    sys.modules["_ref_voices"] = ref
    spec.loader.exec_module(ref)
    rnd = random.Random(1)
    for _ in range(5000):
        c = "".join(rnd.choice("ab_1(2).+ )") for _ in range(rnd.randint(0, 8)))
        try:
            r = ref.parse_voice_spec(c)
            want = ([(x.voice_id, x.weight) for x in r.components], r.normalized_weights(), r.is_blend, r.primary_id)
        except ValueError as e:
            want = str(e)
        try:
            o = ours.parse_voice_spec(c)
            got = ([tuple(x) for x in o.components], o.normalized_weights(), o.is_blend, o.primary_id)
        except ValueError as e:
            got = str(e)
        assert got == want, c


def test_dropin_install_rebinds_every_name():
    """install() against the reference tree, in a fresh interpreter, with the dependent modules imported FIRST (the order that
    used to leave src.streaming / src.vad / src.realtime.server on the reference's ONNX implementation)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not mounted")
    code = r"""
import sys, types, importlib, importlib.machinery, json, logging
logging.disable(logging.CRITICAL)
sys.path.insert(0, "/root/reference"); sys.path.insert(0, %r)
m = types.ModuleType("librosa"); m.__spec__ = importlib.machinery.ModuleSpec("librosa", None); sys.modules["librosa"] = m
import src.vad, src.streaming, src.realtime.audio_buffer, src.realtime.server, src.wyoming.stt_handler, src.main
import open_speech_b200
from open_speech_b200 import dropin
rep = dropin.install()
from open_speech_b200.vad import silero as ours
from open_speech_b200.audio import preprocessing as pre, postprocessing as post
from open_speech_b200.effects import chain
from open_speech_b200.realtime import audio_buffer as ab
from open_speech_b200 import streaming as st
from open_speech_b200.tts import pipeline as pl
import src
checks = {
  "sys.modules silero": sys.modules["src.vad.silero"] is ours,
  "pkg attr silero": src.vad.silero is ours,
  "src.vad.SileroVAD": src.vad.SileroVAD is ours.SileroVAD,
  "src.vad.get_vad_model": src.vad.get_vad_model is ours.get_vad_model,
  "streaming.SileroVAD": src.streaming.SileroVAD is ours.SileroVAD,
  "streaming.get_vad_model": src.streaming.get_vad_model is ours.get_vad_model,
  "streaming.resample_pcm16": src.streaming.resample_pcm16 is st.resample_pcm16,
  "server.SileroVAD": src.realtime.server.SileroVAD is ours.SileroVAD,
  "server.get_vad_model": src.realtime.server.get_vad_model is ours.get_vad_model,
  "server.InputAudioBuffer": src.realtime.server.InputAudioBuffer is ab.InputAudioBuffer,
  "server.decode": src.realtime.server.decode_audio_to_pcm16 is ab.decode_audio_to_pcm16,
  "server.encode": src.realtime.server.encode_pcm16_to_format is ab.encode_pcm16_to_format,
  "audio_buffer.SileroVAD": src.realtime.audio_buffer.SileroVAD is ours.SileroVAD,
  "audio_buffer.InputAudioBuffer": src.realtime.audio_buffer.InputAudioBuffer is ab.InputAudioBuffer,
  "main.preprocess": src.main.preprocess_stt_audio is pre.preprocess_stt_audio,
  "main.process_tts_chunks": src.main.process_tts_chunks is post.process_tts_chunks,
  "main.apply_chain": src.main.apply_chain is chain.apply_chain,
  "pipeline.encode_wav": src.tts.pipeline.encode_wav is pl.encode_wav,
  "voices untouched": sys.modules["src.tts.voices"].__name__ == "src.tts.voices",
}
# the Wyoming handler reads the singleton through the module at call time (stt_handler.py:63-66)
ours._vad_model = "sentinel"
from src.vad.silero import _vad_model as seen
checks["wyoming sees the singleton"] = seen == "sentinel"
print(json.dumps({"checks": checks, "rebound": rep["rebound"]}))
""" % ROOT
    import subprocess

    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    bad = [k for k, v in out["checks"].items() if not v]
    assert not bad, (bad, out["rebound"])


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
    assert S.load_installed_silero_weights() is None  # the package is not installed here
