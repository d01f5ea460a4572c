"""Golden traces of the reference's StreamingSession._process_chunk machine (src/streaming.py:290-355, :429-498).
Build-container only: drives the REFERENCE's own class with a scripted VAD and a stub transcriber, like its test
tests/test_streaming_session_runtime.py:114-135 does, and records what the session did per chunk.

    python oracle/make_golden_stream.py   ->  tests/golden/stream_gate.json
"""
from __future__ import annotations

import asyncio
import json
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("OSB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "stream_gate.json")


class _WS:
    def __init__(self):
        self.sent = []

    async def send_text(self, text):
        self.sent.append(json.loads(text))


class _Backend:
    def __init__(self):
        self.calls = 0

    def is_model_loaded(self, _m):
        return True

    def transcribe(self, **_kw):
        self.calls += 1
        return {"text": "w%d" % self.calls}


class _ScriptVAD:
    def __init__(self, probs):
        self.probs, self.i = probs, 0

    def __call__(self, _samples):
        p = self.probs[self.i % len(self.probs)]
        self.i += 1
        return p


def trace(streaming, probs, sample_rate, chunk_ms, endpointing_ms, vad_enabled, threshold):
    streaming.settings.os_stream_chunk_ms = chunk_ms
    ws, be = _WS(), _Backend()
    streaming.backend_router = be
    s = streaming.StreamingSession(ws, model="m", language=None, sample_rate=sample_rate, interim_results=True,
                                   endpointing_ms=endpointing_ms, vad_enabled=vad_enabled)
    s.vad_state = _ScriptVAD(probs) if vad_enabled else None
    chunk = (np.arange(s.chunk_samples) % 100).astype(np.int16).tobytes()
    steps = []

    async def run():
        for _ in range(len(probs)):
            n_sent, n_calls = len(ws.sent), be.calls
            s.total_samples += s.chunk_samples
            await s._process_chunk(chunk)
            vad_events = [e["state"] for e in ws.sent[n_sent:] if e.get("type") == "vad"]
            finals = [e for e in ws.sent[n_sent:] if e.get("type") == "transcript" and e.get("speech_final")]
            steps.append({"speech_active": bool(s.speech_active), "silence_samples": int(s.silence_samples),
                          "utterance_bytes": len(s.utterance_audio), "vad_events": vad_events, "transcribe_calls": be.calls - n_calls,
                          "final": bool(finals)})

    asyncio.run(run())
    cols = {k: [st[k] for st in steps] for k in ("speech_active", "silence_samples", "utterance_bytes", "transcribe_calls", "final")}
    cols["speech_active"] = [int(v) for v in cols["speech_active"]]
    cols["final"] = [int(v) for v in cols["final"]]
    cols["speech_start"] = [int("speech_start" in st["vad_events"]) for st in steps]
    cols["speech_end"] = [int("speech_end" in st["vad_events"]) for st in steps]
    return {"probs": [round(float(p), 9) for p in probs], "sample_rate": sample_rate, "chunk_ms": chunk_ms, "chunk_samples": s.chunk_samples,
            "endpointing_ms": endpointing_ms, "vad_enabled": vad_enabled, "threshold": threshold, "steps": cols}


def main() -> None:
    warnings.simplefilter("ignore")
    import importlib.machinery
    import types

    stub = types.ModuleType("librosa")
    stub.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
    sys.modules.setdefault("librosa", stub)
    sys.path.insert(0, REF)
    import logging

    logging.disable(logging.CRITICAL)
    from src import streaming

    thr = float(streaming.settings.stt_vad_threshold)
    rng = np.random.default_rng(77)
    cases = []
    scripts = {
        "burst": [0.9] * 6 + [0.1] * 8 + [0.9] * 3 + [0.1] * 6,
        "short_blip": [0.1, 0.9, 0.1, 0.1, 0.1, 0.1, 0.9, 0.9, 0.1, 0.1, 0.1, 0.1],
        "threshold_edge": [thr, thr - 1e-6, thr, 0.0, 0.0, 0.0, 0.0, thr + 1e-6, 0.0, 0.0, 0.0, 0.0],
        "random": list(rng.uniform(0, 1, 120)),
        "sparse": list((rng.uniform(0, 1, 160) > 0.8).astype(float)),
    }
    for name, probs in scripts.items():
        for sr, chunk_ms, ep in ((16000, 100, 300), (8000, 100, 100), (48000, 50, 300), (16000, 20, 60), (44100, 100, 500)):
            c = trace(streaming, probs, sr, chunk_ms, ep, True, thr)
            c["name"] = f"{name}_{sr}_{chunk_ms}_{ep}"
            cases.append(c)
    # the 30 s force-finalise (MAX_UTTERANCE_BYTES) and the VAD-disabled door
    cases.append(dict(trace(streaming, [0.9] * 330, 16000, 100, 300, True, thr), name="max_utterance"))
    cases.append(dict(trace(streaming, [0.0] * 320, 16000, 100, 300, False, thr), name="vad_disabled"))
    cases.append(dict(trace(streaming, [0.0] * 40, 8000, 20, 300, False, thr), name="vad_disabled_short_chunks"))
    with open(OUT, "w") as f:
        json.dump({"source": "reference src/streaming.py StreamingSession._process_chunk driven by oracle/make_golden_stream.py",
                   "max_utterance_bytes": int(streaming.MAX_UTTERANCE_BYTES), "cases": cases}, f)
    print("wrote", len(cases), "cases,", sum(len(c["probs"]) for c in cases), "steps")


if __name__ == "__main__":
    main()
