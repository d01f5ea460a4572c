"""CPU oracle for the open-speech audio hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy / pure Python) of the
reference's algorithms for the hot path named in BASELINE.json.  It exists so
that the CUDA path can be checked against it.  It is NOT a product path and it
is NOT a fallback:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import it;
  * nothing under ``open_speech_b200/`` imports it, and the product raises
    ``RuntimeError`` when ``libosb200.so`` is missing or no GPU is present.

Pinning status (see DESIGN.md "Oracle"):

  pinned against the reference itself (imported from /root/reference in the
  build container by ``oracle/make_golden.py``; vectors in ``tests/golden/``):
    codec.py      G.711 tables (sha256 of SURVEY App. B), linear resample
    resample.py   polyphase resample (scipy.signal.resample_poly, bit-exact)
    stt.py        WAV<->f32, normalize_gain, requantise, preprocess driver
    tts.py        trim / peak normalise / effects chain / blend / f32->int16
    vad.py        framing + segmenter + InputAudioBuffer state machines
                  (driven through the reference's own classes with a scripted
                  session)

  PARITY UNPINNED by the reference (no reference test or runnable third-party
  package holds a value; restated from the published algorithm):
    stt.logmel            faster-whisper 1.2.1 FeatureExtractor
                          (cross-checked against transformers'
                          WhisperFeatureExtractor, which IS installed)
    stt.spectral_gate     noisereduce>=3.0 non-stationary spectral gating
    vad.SileroNet         Silero VAD v5 16 kHz network arithmetic (seeded
                          random-init weights, allowed by BASELINE config 2)
"""
