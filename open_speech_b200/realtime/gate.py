"""Device-resident gates for MANY concurrent streams (BASELINE configs[2]: 1024 realtime streams).

``RealtimeGate`` is S ``InputAudioBuffer`` objects (reference src/realtime/audio_buffer.py:84-166) plus the
``decode_audio_to_pcm16`` in front of them (:37-58) as ONE call per tick: wire bytes of every stream in, resampled pcm16,
per-stream records, LSTM states and the audio arena stay on the GPU, a compact ``[stream, type, ms]`` event list comes
back.  ``StreamGate`` is S ``StreamingSession._process_chunk`` machines (src/streaming.py:290-355) the same way.
PyTorch owns the buffers and the stream; every kernel is launched by libosb200 (csrc/realtime_gate.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .audio_buffer import GATE_STATE

STREAM_STATE = np.dtype([("silence_samples", "<i8"), ("utterance_bytes", "<i8"), ("speech_active", "<i4"), ("reserved", "<i4")])
EVENT_NAMES = {1: "speech_started", 2: "speech_stopped", 3: "frame_too_large", 4: "buffer_full"}
ACT_SPEECH_START, ACT_UTTERANCE_RESET, ACT_APPEND, ACT_TRANSCRIBE, ACT_FINALIZE, ACT_SPEECH_END = 1, 2, 4, 8, 16, 32
_FMT = {"pcm16": N.FMT_PCM16, "g711_ulaw": N.FMT_ULAW, "g711_alaw": N.FMT_ALAW}
_WIRE_RATE = {"pcm16": 24000, "g711_ulaw": 8000, "g711_alaw": 8000}  # audio_buffer.py:47-56


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def resampled_len(n_in: int, from_rate: int, poly: bool) -> int:
    if from_rate == 16000:
        return n_in
    if not poly:
        return int(n_in * (16000 / from_rate))  # audio_buffer.py:28
    g = int(np.gcd(16000, from_rate))
    up, down = 16000 // g, from_rate // g
    return (n_in * up + down - 1) // down       # streaming.py:80-84 (scipy resample_poly output length)


class RealtimeGate:
    """S realtime input buffers on the device."""

    def __init__(self, n_streams: int, chunk: int, fmt: str = "g711_ulaw", session=None, threshold: float = 0.5,
                 silence_duration_ms: int = 500, arena_samples: int = 0, poly: bool = False, from_rate: int | None = None,
                 max_events: int | None = None, host_io: bool = False):
        N.require_gpu()
        self.S, self.chunk, self.fmt = n_streams, chunk, _FMT[fmt]
        self.from_rate = from_rate if from_rate is not None else _WIRE_RATE[fmt]
        self.poly = bool(poly)
        self.n_out = resampled_len(chunk, self.from_rate, self.poly)
        self.session, self.threshold, self.silence_ms = session, threshold, silence_duration_ms
        self.max_events = max_events if max_events is not None else n_streams
        dev = torch.device("cuda", torch.cuda.current_device())
        self.state = torch.zeros((n_streams, 4), dtype=torch.int64, device=dev)          # osb_gate_state records
        self.vad_state = torch.zeros((n_streams, 2, 128), dtype=torch.float32, device=dev)
        # host_io: the tick's outputs (pcm16, event list) live in PINNED HOST memory and the kernels write them there themselves (unified
        # addressing: a pinned buffer is a valid device pointer); with tick_host() reading the wire bytes the same way, a tick is two
        # kernel launches and one synchronise -- no H2D / D2H copies (1024 streams: p50 49 us instead of 66 us)
        self.host_io = bool(host_io)
        if self.host_io:
            self.pcm = torch.empty((n_streams, max(self.n_out, 1)), dtype=torch.int16).pin_memory()
            self.events = torch.zeros((self.max_events + 1, 3), dtype=torch.int32).pin_memory()
        else:
            self.pcm = torch.empty((n_streams, max(self.n_out, 1)), dtype=torch.int16, device=dev)
            self.events = torch.zeros((self.max_events + 1, 3), dtype=torch.int32, device=dev)  # last row, first word: the count
        self.work = torch.zeros(N.lib().osb_gate_work_bytes(n_streams) // 4 + 4, dtype=torch.int32, device=dev)
        self.arena = torch.empty((n_streams, arena_samples), dtype=torch.int16, device=dev) if arena_samples else None
        self.in_dtype = torch.int16 if self.fmt == N.FMT_PCM16 else torch.uint8
        self._ev_host = torch.zeros((self.max_events + 1, 3), dtype=torch.int32).pin_memory()

    def tick(self, wire: torch.Tensor, probs: torch.Tensor | None = None) -> None:
        """wire: [S, chunk] device tensor of this tick's bytes / samples; probs: optional scripted chunk probabilities [S]."""
        if wire.shape != (self.S, self.chunk) or wire.dtype != self.in_dtype or not wire.is_contiguous():
            raise ValueError("wire must be a contiguous tensor [n_streams, chunk] of the wire dtype")
        if not (wire.is_cuda or (self.host_io and wire.is_pinned())):
            raise ValueError("wire must be a CUDA tensor (or, on a host_io gate, a pinned host tensor)")
        gated = self.session is not None or probs is not None
        N.call("osb_gate_tick_dev", self.session.handle if self.session is not None else None, wire.data_ptr(), self.fmt, self.chunk,
               self.from_rate, int(self.poly), self.S, self.chunk, self.pcm.data_ptr(), self.n_out, self.state.data_ptr(),
               self.vad_state.data_ptr(), probs.data_ptr() if probs is not None else None, int(gated),
               self.arena.data_ptr() if self.arena is not None else None, self.arena.shape[1] if self.arena is not None else 0,
               float(self.threshold), int(self.silence_ms), self.work.data_ptr(), self.events.data_ptr(),
               self.events[self.max_events].data_ptr(), self.max_events, _stream())

    def read_events(self) -> list[tuple[int, str, int]]:
        """The compact list of the last tick (synchronises the current stream; a D2H copy first unless the gate is host_io)."""
        ev = self.events if self.host_io else self._ev_host
        if not self.host_io:
            self._ev_host.copy_(self.events, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        k = int(ev[self.max_events, 0])
        if k == 0:
            return []
        return [(int(s), EVENT_NAMES[int(t)], int(ms)) for s, t, ms in ev[: min(k, self.max_events)].tolist()]

    def tick_host(self, wire: torch.Tensor, probs: torch.Tensor | None = None) -> list[tuple[int, str, int]]:
        """One tick with HOST buffers on a host_io gate: ``wire`` is a pinned host tensor [S, chunk]; on return ``self.pcm`` (pinned host)
        holds the tick's pcm16 and the events are returned.  Nothing is copied: the kernels read and write the pinned buffers."""
        if not self.host_io:
            raise RuntimeError("tick_host needs a gate created with host_io=True")
        self.tick(wire, probs)
        return self.read_events()

    def capture(self) -> "GateTickGraph":
        """The whole tick -- pinned wire slot -> H2D -> decode + resample + buffer + gate -> D2H of pcm16 and the event list -- as ONE
        CUDA graph launch (the tick is launch-latency bound: five driver submissions become one).  Chunks shorter than a VAD window
        only (the reference's 20 ms case): scoring allocates stream-ordered scratch, which is not captured here."""
        if self.n_out >= 512 and self.session is not None:
            raise RuntimeError("graph capture covers ticks without a full VAD window (chunk < 512 samples at 16 kHz)")
        if self.host_io:
            raise RuntimeError("a host_io gate has no copies to capture: use tick_host()")
        return GateTickGraph(self)

    def records(self) -> np.ndarray:
        return self.state.cpu().numpy().view(GATE_STATE).reshape(self.S)

    def clear(self, streams) -> None:
        ids = torch.as_tensor(list(streams), dtype=torch.int32, device=self.state.device)
        N.call("osb_gate_clear_dev", self.state.data_ptr(), ids.data_ptr(), ids.numel(), _stream())

    def commit(self, stream: int) -> torch.Tensor:
        """The stream's buffered pcm16 (device tensor, a copy) and the clear() the reference's commit() does."""
        if self.arena is None:
            raise RuntimeError("RealtimeGate was created without an arena")
        n = int(self.state[stream, 2].item())
        out = self.arena[stream, :n].clone()
        self.clear([stream])
        return out


class GateTickGraph:
    """A captured tick of a :class:`RealtimeGate`: write the tick's bytes into ``host_in`` (pinned), ``run()``, read ``host_pcm``
    and the returned events."""

    def __init__(self, gate: RealtimeGate):
        self.gate = gate
        self.host_in = torch.zeros((gate.S, gate.chunk), dtype=gate.in_dtype).pin_memory()
        self.host_pcm = torch.empty((gate.S, max(gate.n_out, 1)), dtype=torch.int16).pin_memory()
        self.dev_in = torch.empty((gate.S, gate.chunk), dtype=gate.in_dtype, device=gate.state.device)
        self.stream = torch.cuda.Stream()
        saved = (gate.state.clone(), gate.work.clone())
        with torch.cuda.stream(self.stream):
            self._body()  # warm-up outside capture (lazy initialisation inside the library, e.g. the resampler's table)
        self.stream.synchronize()
        gate.state.copy_(saved[0])
        gate.work.copy_(saved[1])
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body()

    def _body(self):
        g = self.gate
        self.dev_in.copy_(self.host_in, non_blocking=True)
        g.tick(self.dev_in)
        self.host_pcm.copy_(g.pcm, non_blocking=True)
        g._ev_host.copy_(g.events, non_blocking=True)

    def run(self) -> list[tuple[int, str, int]]:
        g = self.gate
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        self.stream.synchronize()
        k = int(g._ev_host[g.max_events, 0])
        if k == 0:
            return []
        return [(int(s), EVENT_NAMES[int(t)], int(ms)) for s, t, ms in g._ev_host[: min(k, g.max_events)].tolist()]


class StreamGate:
    """S streaming-session utterance machines on the device (client-rate pcm16 chunks in, OSB_ACT_* bits out)."""

    def __init__(self, n_streams: int, chunk: int, sample_rate: int, session=None, threshold: float = 0.5, endpointing_ms: int = 300,
                 max_utterance_seconds: int = 30):
        N.require_gpu()
        self.S, self.chunk, self.rate = n_streams, chunk, sample_rate
        self.n_out = resampled_len(chunk, sample_rate, True)
        self.session, self.threshold = session, threshold
        self.endpointing_samples = int(16000 * endpointing_ms / 1000)   # streaming.py:191
        self.max_utt_bytes = max_utterance_seconds * 16000 * 2          # streaming.py:42-43
        dev = torch.device("cuda", torch.cuda.current_device())
        self.state = torch.zeros((n_streams, 3), dtype=torch.int64, device=dev)
        self.vad_state = torch.zeros((n_streams, 2, 128), dtype=torch.float32, device=dev)
        self.pcm = torch.empty((n_streams, max(self.n_out, 1)), dtype=torch.int16, device=dev)
        self.actions = torch.zeros(n_streams, dtype=torch.int32, device=dev)

    def tick(self, chunk: torch.Tensor, probs: torch.Tensor | None = None, vad_enabled: bool = True) -> torch.Tensor:
        if chunk.shape != (self.S, self.chunk) or chunk.dtype != torch.int16 or not chunk.is_cuda or not chunk.is_contiguous():
            raise ValueError("chunk must be a contiguous CUDA int16 tensor [n_streams, chunk]")
        N.call("osb_stream_tick_dev", self.session.handle if self.session is not None else None, chunk.data_ptr(), self.chunk, self.rate,
               self.S, self.chunk, self.pcm.data_ptr(), self.n_out, self.state.data_ptr(), self.vad_state.data_ptr(),
               probs.data_ptr() if probs is not None else None, int(vad_enabled), float(self.threshold), self.endpointing_samples,
               self.max_utt_bytes, self.actions.data_ptr(), _stream())
        return self.actions

    def records(self) -> np.ndarray:
        return self.state.cpu().numpy().view(STREAM_STATE).reshape(self.S)
