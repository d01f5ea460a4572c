"""Put this package behind the reference's module paths.

Three mechanisms, applied in an order that does not depend on what the host application imported first:

1. whole-module aliases: ``sys.modules[ref] = ours`` (and the attribute on the parent package), for modules that hold
   nothing but hot-path code.  ``src.vad.silero`` is one of them ON PURPOSE: the Wyoming handler reads the singleton
   through ``from src.vad.silero import _vad_model`` at call time (src/wyoming/stt_handler.py:63-66), so the module
   object that ``get_vad_model`` writes to must be the one registered under that name.
2. name patches for modules that also hold non-hot-path code (``src.streaming``, ``src.realtime.audio_buffer``,
   ``src.tts.pipeline``): the listed attributes are replaced.
3. an identity sweep over every already-imported ``src.*`` module: any attribute that IS one of the replaced reference
   objects (``from src.vad.silero import SileroVAD, get_vad_model`` in src/streaming.py:33, src/realtime/server.py:28,
   src/vad/__init__.py:3, ``from src.audio.preprocessing import preprocess_stt_audio`` in src/main.py:36, ...) is rebound.

``src.tts.voices`` is left alone: it is host-side text parsing, nothing in it runs on the GPU.
``install()`` returns a report {"aliased": [...], "patched": {...}, "rebound": {module: [names]}} that the tests check.
"""
from __future__ import annotations

import importlib
import sys
import types

_ALIAS = {
    "src.audio.preprocessing": "open_speech_b200.audio.preprocessing",
    "src.audio.postprocessing": "open_speech_b200.audio.postprocessing",
    "src.effects.chain": "open_speech_b200.effects.chain",
    "src.vad.silero": "open_speech_b200.vad.silero",
}
_PATCH = {
    "src.realtime.audio_buffer": ("open_speech_b200.realtime.audio_buffer",
                                  ("_resample_linear", "decode_audio_to_pcm16", "encode_pcm16_to_format", "InputAudioBuffer")),
    "src.streaming": ("open_speech_b200.streaming", ("resample_pcm16",)),
    "src.tts.pipeline": ("open_speech_b200.tts.pipeline", ("float32_to_int16", "encode_wav", "encode_pcm")),
}


def _public_callables(mod: types.ModuleType):
    for name, obj in vars(mod).items():
        if name.startswith("__"):
            continue
        if isinstance(obj, (types.FunctionType, type)) and getattr(obj, "__module__", None) == mod.__name__:
            yield name, obj


def install() -> dict:
    """Call once the reference's ``src`` package is importable; safe before or after its modules were imported."""
    replaced: dict[int, object] = {}  # id(reference object) -> our object
    keep_alive = []                   # the reference objects stay referenced while their ids are used as keys
    report = {"aliased": [], "patched": {}, "rebound": {}}

    def remember(old, new):
        if old is not new:
            replaced[id(old)] = new
            keep_alive.append(old)

    for ref, ours_name in _ALIAS.items():
        ours = importlib.import_module(ours_name)
        old = sys.modules.get(ref)
        if old is None:
            try:
                old = importlib.import_module(ref)  # so that names other modules already hold can be recognised
            except Exception:
                old = None
        if old is not None and old is not ours:
            for name, obj in _public_callables(old):
                if hasattr(ours, name):
                    remember(obj, getattr(ours, name))
        sys.modules[ref] = ours
        parent, _, leaf = ref.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, ours)
        report["aliased"].append(ref)

    for ref, (ours_name, names) in _PATCH.items():
        ours = importlib.import_module(ours_name)
        try:
            target = importlib.import_module(ref)
        except Exception:
            continue
        done = []
        for n in names:
            if hasattr(target, n):
                remember(getattr(target, n), getattr(ours, n))
            setattr(target, n, getattr(ours, n))
            done.append(n)
        report["patched"][ref] = done

    for mod_name, mod in list(sys.modules.items()):
        if mod is None or not (mod_name == "src" or mod_name.startswith("src.")):
            continue
        if getattr(mod, "__name__", "").startswith("open_speech_b200"):
            continue
        for attr, obj in list(vars(mod).items()):
            new = replaced.get(id(obj))
            if new is not None and obj is not new:
                setattr(mod, attr, new)
                report["rebound"].setdefault(mod_name, []).append(attr)
    return report
