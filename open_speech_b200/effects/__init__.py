"""Effects package (drop-in for src/effects)."""
from .chain import apply_chain

__all__ = ["apply_chain"]
