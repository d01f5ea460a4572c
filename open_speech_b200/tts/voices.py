"""Voice specifications for the GPU blend path.

Same public names and behaviour as the reference's src/tts/voices.py (``parse_voice_spec`` :58-82,
``VoiceSpec.normalized_weights`` :29-33, ``resolve_voice_name`` :50-55) so that ``KokoroBackend`` and the routers can
use either module, but written for the batched device path: a spec is an immutable pair of tuples, parsing is one
hand-written scan (no regex), and ``blend_operands`` turns MANY specs into the ``[batch][kmax]`` index / weight
arrays that ``osb_voice_blend_dev`` consumes in one launch (include/osb200.h).  ``dropin.install()`` leaves the
reference's own module in place: nothing here touches the GPU, so there is nothing to replace.
"""
from __future__ import annotations

from typing import Iterable, Mapping, NamedTuple, Sequence

import numpy as np

_ID_CHARS = frozenset("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_")

# OpenAI voice names -> Kokoro voice ids (the reference's table, src/tts/voices.py:36-43: public API data)
OPENAI_VOICE_MAP: dict[str, str] = dict(alloy="af_heart", echo="am_adam", fable="bf_emma", onyx="am_michael", nova="af_nova",
                                        shimmer="af_bella")


class VoiceComponent(NamedTuple):
    voice_id: str
    weight: float = 1.0


class VoiceSpec:
    """One or more weighted voices.  ``components`` is a tuple of :class:`VoiceComponent`."""

    __slots__ = ("components",)

    def __init__(self, components: Iterable[VoiceComponent]):
        self.components = tuple(components)

    def __repr__(self) -> str:
        return f"VoiceSpec(components={list(self.components)!r})"

    def __eq__(self, other) -> bool:
        return isinstance(other, VoiceSpec) and self.components == other.components

    @property
    def is_blend(self) -> bool:
        return len(self.components) > 1

    @property
    def primary_id(self) -> str:
        return self.components[0].voice_id

    def normalized_weights(self) -> list[float]:
        """weights / sum(weights); equal shares when the sum is zero (reference :29-33)."""
        k = len(self.components)
        total = sum(w for _, w in self.components)
        if total == 0:
            return [1.0 / k] * k
        return [w / total for _, w in self.components]


def resolve_voice_name(voice: str) -> str:
    return OPENAI_VOICE_MAP.get(voice, voice)


def _scan_component(part: str) -> VoiceComponent:
    """``name`` or ``name(weight)``: name = [A-Za-z0-9_]+, weight = digits[.digits]; anything else is a ValueError."""
    bad = ValueError(f"Invalid voice spec component: {part!r}")
    n = len(part)
    i = 0
    while i < n and part[i] in _ID_CHARS:
        i += 1
    if i == 0:
        raise bad
    name = part[:i]
    if i == n:
        return VoiceComponent(name, 1.0)
    if part[i] != "(" or part[-1] != ")":
        raise bad
    num = part[i + 1:-1]
    whole, dot, frac = num.partition(".")
    if not whole.isdecimal() or (dot and not frac.isdecimal()):  # Unicode decimals, like the reference's \d
        raise bad
    return VoiceComponent(name, float(num))


def parse_voice_spec(voice: str) -> VoiceSpec:
    """'af_bella' | 'alloy' | 'af_bella+af_sky' | 'af_bella(2)+af_sky(1)' -> VoiceSpec.

    OpenAI aliases resolve only for a single plain name (no '+', no '('), like the reference (:68-70)."""
    if "+" not in voice and "(" not in voice:
        voice = resolve_voice_name(voice)
    return VoiceSpec(_scan_component(p.strip()) for p in voice.split("+"))


def blend_operands(specs: Sequence[VoiceSpec | str], voice_index: Mapping[str, int], kmax: int | None = None):
    """Batch of specs -> (idx int32 [B, kmax] with -1 terminators, weights float32 [B, kmax]) for osb_voice_blend_dev.

    ``voice_index`` maps a voice id to its row in the resident pack table.  Weights are the normalised weights rounded to
    float32, which is what ``KokoroBackend._blend_voices`` multiplies with (kokoro.py:303-306)."""
    parsed = [parse_voice_spec(s) if isinstance(s, str) else s for s in specs]
    k = max((len(s.components) for s in parsed), default=1)
    if kmax is None:
        kmax = k
    elif k > kmax:
        raise ValueError(f"a spec has {k} components, kmax is {kmax}")
    idx = np.full((len(parsed), kmax), -1, dtype=np.int32)
    wts = np.zeros((len(parsed), kmax), dtype=np.float32)
    for b, s in enumerate(parsed):
        for j, (c, w) in enumerate(zip(s.components, s.normalized_weights())):
            if c.voice_id not in voice_index:
                raise KeyError(f"voice {c.voice_id!r} is not in the resident pack table")
            idx[b, j] = voice_index[c.voice_id]
            wts[b, j] = w
    return idx, wts
