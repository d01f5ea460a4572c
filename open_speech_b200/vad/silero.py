"""Drop-in for src/vad/silero.py (reference lines 38-209), GPU-backed.

``SileroVAD(session, threshold)`` keeps the reference's attributes (.session, .sample_rate,
.threshold, ._state[2,1,128]) and methods.  ``session`` is a :class:`VadSession` (weights
resident on the GPU, shared between per-stream SileroVAD objects like the reference shares
one ONNX session).  There is no host implementation: the window loop, the network and the
segmenter all run in libosb200, and any other kind of session object is refused.
"""
from __future__ import annotations

import asyncio
import ctypes
import logging
import os
from dataclasses import dataclass

import numpy as np

from .. import _native as N

logger = logging.getLogger(__name__)

VAD_SAMPLE_RATE = 16000
WINDOW = 512
ALLOW_RANDOM_INIT_ENV = "OSB_VAD_ALLOW_RANDOM_INIT"

_vad_model: "SileroVAD | None" = None
_vad_lock = asyncio.Lock()

# flat weight layout shared with csrc/vad.cu (name, shape)
WEIGHT_LAYOUT = (
    ("stft_basis", (258, 256)),
    ("enc1.weight", (128, 129, 3)), ("enc1.bias", (128,)),
    ("enc2.weight", (64, 128, 3)), ("enc2.bias", (64,)),
    ("enc3.weight", (64, 64, 3)), ("enc3.bias", (64,)),
    ("enc4.weight", (128, 64, 3)), ("enc4.bias", (128,)),
    ("lstm.weight_ih", (512, 128)), ("lstm.weight_hh", (512, 128)),
    ("lstm.bias_ih", (512,)), ("lstm.bias_hh", (512,)),
    ("dec.weight", (128,)), ("dec.bias", (1,)),
)


@dataclass
class Segment:
    """A detected speech segment."""
    start_ms: int
    end_ms: int


def stft_basis() -> np.ndarray:
    n = np.arange(256)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / 256)
    fb = np.fft.fft(np.eye(256))
    return (np.vstack([np.real(fb[:129]), np.imag(fb[:129])]) * win[None, :]).astype(np.float32)


def random_init_weights(seed: int = 1002) -> dict[str, np.ndarray]:
    """Seeded random-init Silero-v5-shaped weights (BASELINE config 2 allows this when the
    silero-vad package / ONNX file is absent -- it is: the reference downloads it at run time)."""
    rng = np.random.default_rng(seed)
    w: dict[str, np.ndarray] = {"stft_basis": stft_basis()}

    def u(shape, bound):
        return rng.uniform(-bound, bound, size=shape).astype(np.float32)

    for name, oc, ic, k in (("enc1", 128, 129, 3), ("enc2", 64, 128, 3), ("enc3", 64, 64, 3), ("enc4", 128, 64, 3)):
        bound = 1.0 / np.sqrt(ic * k)
        w[f"{name}.weight"] = u((oc, ic, k), bound)
        if name == "enc1":
            w[f"{name}.weight"] = (w[f"{name}.weight"] * 16.0).astype(np.float32)
        w[f"{name}.bias"] = u((oc,), bound)
    b = 1.0 / np.sqrt(128)
    w["lstm.weight_ih"] = u((512, 128), b)
    w["lstm.weight_hh"] = u((512, 128), b)
    w["lstm.bias_ih"] = u((512,), b)
    w["lstm.bias_hh"] = u((512,), b)
    w["dec.weight"] = (u((128,), b) * -120.0).astype(np.float32)
    w["dec.bias"] = np.array([-2.9], dtype=np.float32)
    return w


# silero-vad v5 (16 kHz branch) state-dict names -> WEIGHT_LAYOUT names.  The packaged TorchScript model prefixes the 16 kHz
# branch with "_model." (the 8 kHz one with "_model_8k."); the plain nn.Module export has no prefix.
_SILERO_KEYS = {
    "stft.forward_basis_buffer": "stft_basis",
    "encoder.0.reparam_conv.weight": "enc1.weight", "encoder.0.reparam_conv.bias": "enc1.bias",
    "encoder.1.reparam_conv.weight": "enc2.weight", "encoder.1.reparam_conv.bias": "enc2.bias",
    "encoder.2.reparam_conv.weight": "enc3.weight", "encoder.2.reparam_conv.bias": "enc3.bias",
    "encoder.3.reparam_conv.weight": "enc4.weight", "encoder.3.reparam_conv.bias": "enc4.bias",
    "decoder.rnn.weight_ih": "lstm.weight_ih", "decoder.rnn.weight_hh": "lstm.weight_hh",
    "decoder.rnn.bias_ih": "lstm.bias_ih", "decoder.rnn.bias_hh": "lstm.bias_hh",
    "decoder.decoder.2.weight": "dec.weight", "decoder.decoder.2.bias": "dec.bias",
}


def weights_from_state_dict(sd) -> dict[str, np.ndarray]:
    """Silero-v5-shaped state dict (torch tensors or arrays) -> the weight dict of :class:`VadSession`.

    Accepts the key names of the silero-vad package's 16 kHz branch with or without the "_model." prefix; every tensor is
    reshaped to the layout shape (the STFT basis is stored as [258, 1, 256], the 1x1 output conv as [1, 128, 1]) and the
    element counts must match exactly, so a model of another architecture fails loudly instead of being mis-read."""
    shapes = dict(WEIGHT_LAYOUT)
    out: dict[str, np.ndarray] = {}
    for key, val in sd.items():
        k = key[len("_model."):] if key.startswith("_model.") else key
        name = _SILERO_KEYS.get(k)
        if name is None:
            continue
        a = val.detach().cpu().numpy() if hasattr(val, "detach") else np.asarray(val)
        if a.size != int(np.prod(shapes[name])):
            raise ValueError(f"VAD weight {key}: {a.shape} does not fit {name} {shapes[name]}")
        out[name] = np.ascontiguousarray(a, dtype=np.float32).reshape(shapes[name])
    missing = [n for n, _ in WEIGHT_LAYOUT if n not in out]
    if missing:
        raise ValueError(f"VAD state dict lacks {missing}")
    return out


def load_installed_silero_weights() -> dict[str, np.ndarray] | None:
    """Weights of the installed ``silero_vad`` package (BASELINE config 2), or None when it is absent / has another layout.
    Nothing is downloaded: the reference fetches its ONNX file at run time (vad/silero.py:28), this never does."""
    try:
        import silero_vad  # type: ignore

        model = silero_vad.load_silero_vad(onnx=False)
        return weights_from_state_dict(model.state_dict())
    except Exception as e:  # ImportError, or a package version with a different architecture
        logger.warning("silero_vad weights not available (%s)", e)
        return None


def pack_weights(w: dict[str, np.ndarray]) -> np.ndarray:
    parts = []
    for name, shape in WEIGHT_LAYOUT:
        a = np.ascontiguousarray(w[name], dtype=np.float32)
        if tuple(a.shape) != shape:
            raise ValueError(f"VAD weight {name}: shape {a.shape}, expected {shape}")
        parts.append(a.reshape(-1))
    return np.concatenate(parts)


class VadSession:
    """Silero weights resident on the current GPU (the analogue of the shared ORT session)."""

    def __init__(self, weights: dict[str, np.ndarray] | None = None):
        self.weights = weights if weights is not None else random_init_weights()
        flat = pack_weights(self.weights)
        h = ctypes.c_void_p()
        N.call("osb_vad_create", N.ptr(flat), flat.size, ctypes.byref(h))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                N.lib().osb_vad_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class SileroVAD:
    """Per-stream VAD state over a shared session (reference class at :45-177)."""

    def __init__(self, session, threshold: float = 0.5):
        if not isinstance(session, VadSession):
            raise TypeError("SileroVAD needs a VadSession (weights resident on the GPU); open_speech_b200 has no host "
                            f"implementation to drive a {type(session).__name__} with")
        self.session = session
        self.sample_rate = VAD_SAMPLE_RATE
        self.threshold = threshold
        self._state = np.zeros((2, 1, 128), dtype=np.float32)

    def reset(self):
        self._state = np.zeros((2, 1, 128), dtype=np.float32)

    # ------------------------------------------------------------------ scoring
    def _score(self, buf, fmt: int, n: int) -> np.ndarray:
        """Per-window probabilities of the full windows in buf; advances self._state."""
        n_win = n // WINDOW
        probs = np.empty(n_win, dtype=np.float32)
        if n_win == 0:
            return probs
        st = np.ascontiguousarray(self._state, dtype=np.float32)
        mx = ctypes.c_float(0.0)
        N.call("osb_vad_score_host", self.session.handle, buf, fmt, n, N.ptr(st), N.ptr(probs), ctypes.byref(mx))
        self._state = st
        return probs

    def __call__(self, audio: np.ndarray) -> float:
        """Speech probability 0-1: max over the full 512-sample windows (0.0 if none)."""
        if len(audio) == 0:
            return 0.0
        a = np.ascontiguousarray(audio, dtype=np.float32)
        probs = self._score(N.ptr(a), N.FMT_F32, len(a))
        max_prob = 0.0
        for p in probs:
            if float(p) > max_prob:
                max_prob = float(p)
        return max_prob

    def score_pcm16(self, pcm16_bytes: bytes) -> float:
        """__call__ on int16 bytes (the /32768 convert is fused into the kernel)."""
        if not pcm16_bytes:
            return 0.0
        probs = self._score(pcm16_bytes, N.FMT_PCM16, len(pcm16_bytes) // 2)
        return float(probs.max()) if len(probs) and float(probs.max()) > 0.0 else 0.0

    def is_speech(self, pcm16_bytes: bytes, threshold: float | None = None) -> bool:
        if not pcm16_bytes:
            return False
        prob = self.score_pcm16(pcm16_bytes)
        return prob >= (threshold if threshold is not None else self.threshold)

    def get_speech_segments(self, pcm16_bytes: bytes, threshold: float | None = None,
                            min_speech_ms: int = 250, silence_ms: int = 800) -> list[Segment]:
        if not pcm16_bytes:
            return []
        thresh = threshold if threshold is not None else self.threshold
        n = len(pcm16_bytes) // 2
        if len(pcm16_bytes) % 2:
            raise ValueError("buffer size must be a multiple of element size")  # np.frombuffer(..., int16) in the reference (:131)
        st = np.ascontiguousarray(self._state, dtype=np.float32)
        max_seg = n // WINDOW // 2 + 2  # a segment needs a speech window and a silence window: never more than this
        segs = np.empty((max_seg, 2), dtype=np.int32)
        cnt = ctypes.c_int(0)
        N.call("osb_vad_segments_host", self.session.handle, pcm16_bytes, N.FMT_PCM16, n, N.ptr(st),
               float(thresh), int(min_speech_ms), int(silence_ms), N.ptr(segs), max_seg, ctypes.byref(cnt))
        self._state = st
        return [Segment(int(s), int(e)) for s, e in segs[: cnt.value]]


async def get_vad_model() -> SileroVAD:
    """Lazy singleton (reference :180-209).  The weights come from the installed silero-vad package; nothing is downloaded.

    Like the reference, which raises when its model cannot be fetched or loaded (:196-206; the realtime server then disables
    server VAD, server.py:55-59), this raises when no real weights are available.  The seeded random-init network of
    BASELINE configs[1] is for benchmarks and tests only: ``VadSession()`` builds it explicitly, or set
    OSB_VAD_ALLOW_RANDOM_INIT=1 to let the server singleton use it."""
    global _vad_model
    if _vad_model is not None:
        return _vad_model
    async with _vad_lock:
        if _vad_model is None:
            weights = load_installed_silero_weights()
            if weights is None:
                if os.environ.get(ALLOW_RANDOM_INIT_ENV) != "1":
                    raise RuntimeError("Silero VAD weights are not available (pip package `silero-vad` missing or of another layout); "
                                       f"refusing to gate speech with a random network (set {ALLOW_RANDOM_INIT_ENV}=1 to allow it)")
                logger.error("Silero VAD: using the seeded RANDOM-INIT network (%s=1); its events are meaningless", ALLOW_RANDOM_INIT_ENV)
            _vad_model = SileroVAD(VadSession(weights))
            logger.info("Silero VAD weights uploaded to GPU")
        return _vad_model
