"""Golden vectors for the SURVEY 8(f) rows, produced by the REFERENCE's own code.  Build-container only."""
from __future__ import annotations

import ast
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("OSB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _extract_function(path: str, name: str, ns: dict):
    """Run one top-level function of a reference module without importing the module (its imports are unavailable)."""
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main() -> None:
    warnings.simplefilter("ignore")
    import importlib.machinery
    import types

    import transformers  # noqa: F401  (its librosa probe must run before the stub exists)

    stub = types.ModuleType("librosa")  # src/effects/chain.py imports librosa at module level; it is not installed
    stub.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
    sys.modules.setdefault("librosa", stub)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from open_speech_b200 import synth
    from src.composer import MultiTrackComposer

    c = MultiTrackComposer()
    g = {}
    a = synth.tts_utterance(0.8, seed=71)
    g["comp_in_24k"] = a
    for dst in (16000, 48000, 22050, 8000):
        g[f"comp_resample_24k_{dst}"] = c._resample(a, 24000, dst)
    b = synth.tts_utterance(0.5, seed=72) * 1.5
    tracks = [{"samples": a, "offset_s": 0.0}, {"samples": b, "offset_s": 0.25}, {"samples": a[:5000] * 2.0, "offset_s": 0.9}]
    g["comp_track_b"] = b
    g["comp_mix"] = c._mix_prepared(tracks, 24000)
    g["comp_int16"] = c._float_to_int16(g["comp_mix"])
    f = _extract_function(os.path.join(REF, "src/wyoming/tts_handler.py"), "_resample_to_16k",
                          {"np": np, "WYOMING_RATE": 16000, "TTS_SAMPLE_RATE": 24000})
    g["wy_24k_to_16k"] = f(a, 24000)
    g["wy_22050_to_16k"] = f(a[:7777], 22050)
    g["wy_48k_to_16k"] = f(a, 48000)
    # realtime TTS output framing (src/realtime/server.py:238-277).  The handler's body is a closure inside a coroutine, so its
    # three statements are replayed here around the REFERENCE's own encode_pcm16_to_format and the stdlib base64 it calls.
    import base64

    from src.realtime.audio_buffer import encode_pcm16_to_format

    rt = np.concatenate([a * 1.7, np.array([1.0, -1.0, 1.5, -1.5, 0.99999, -0.99999, 3.0517578125e-05, -3.0517578125e-05, 0.0],
                                           dtype=np.float32)]).astype(np.float32)
    g["rt_in_24k"] = rt
    for n_take, tag in ((len(rt), "full"), (4001, "odd"), (2, "tiny")):
        for fmt in ("pcm16", "g711_ulaw", "g711_alaw"):
            combined = np.concatenate([rt[:n_take // 2], rt[n_take // 2:n_take]])
            pcm16 = (combined * 32767).clip(-32768, 32767).astype(np.int16).tobytes()
            audio_data = encode_pcm16_to_format(pcm16, 24000, fmt)
            deltas = [base64.b64encode(audio_data[i:i + 3000]).decode("ascii") for i in range(0, len(audio_data), 3000)]
            g[f"rt_payload_{tag}_{fmt}"] = np.frombuffer(audio_data, dtype=np.uint8)
            g[f"rt_deltas_{tag}_{fmt}"] = np.frombuffer("\n".join(deltas).encode("ascii"), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "reference_vectors_next.npz"), **g)
    print("wrote", len(g), "arrays")


if __name__ == "__main__":
    main()
