"""Voice Activity Detection (drop-in for src/vad)."""
from .silero import Segment, SileroVAD, get_vad_model

__all__ = ["SileroVAD", "Segment", "get_vad_model"]
