"""Voice registry, presets and blend parsing (drop-in for src/tts/voices.py, host-side text parsing)."""
from __future__ import annotations

import re
from dataclasses import dataclass


@dataclass
class VoiceComponent:
    voice_id: str
    weight: float = 1.0


@dataclass
class VoiceSpec:
    components: list[VoiceComponent]

    @property
    def is_blend(self) -> bool:
        return len(self.components) > 1

    @property
    def primary_id(self) -> str:
        return self.components[0].voice_id

    def normalized_weights(self) -> list[float]:
        total = sum(c.weight for c in self.components)
        if total == 0:
            return [1.0 / len(self.components)] * len(self.components)
        return [c.weight / total for c in self.components]


OPENAI_VOICE_MAP: dict[str, str] = {
    "alloy": "af_heart", "echo": "am_adam", "fable": "bf_emma", "onyx": "am_michael", "nova": "af_nova", "shimmer": "af_bella",
}

_COMPONENT_RE = re.compile(r"([a-zA-Z0-9_]+)(?:\((\d+(?:\.\d+)?)\))?")


def resolve_voice_name(voice: str) -> str:
    return OPENAI_VOICE_MAP.get(voice, voice)


def parse_voice_spec(voice: str) -> VoiceSpec:
    """'af_bella(2)+af_sky(1)' -> VoiceSpec; aliases resolved only for single plain names."""
    if "+" not in voice and "(" not in voice:
        voice = resolve_voice_name(voice)
    components = []
    for part in voice.split("+"):
        part = part.strip()
        m = _COMPONENT_RE.fullmatch(part)
        if not m:
            raise ValueError(f"Invalid voice spec component: {part!r}")
        components.append(VoiceComponent(voice_id=m.group(1), weight=float(m.group(2)) if m.group(2) else 1.0))
    return VoiceSpec(components=components)
