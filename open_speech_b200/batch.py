"""Batched, device-resident entry points (what bench.py and the multi-GPU driver call).

PyTorch is only plumbing here: it owns device/pinned memory and the CUDA stream; every kernel is
launched through libosb200's ``*_dev`` C entry points on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class SttFrontEnd:
    """pcm16 clips -> [spectral gate] -> [normalise] -> requantise -> log-mel (BASELINE configs 1 / 4).

    The device-side equivalent of ``FeatureExtractor(decode(preprocess_stt_audio(wav)))`` for a batch of
    equal-length clips: input int16 [B, n] (device), output float32 [B, n_mels, (n+160)//160] (device).
    """

    def __init__(self, n_mels: int = 128, sample_rate: int = 16000, noise_reduce: bool = False, normalize: bool = True):
        N.require_gpu()
        self.n_mels, self.sample_rate = n_mels, sample_rate
        self.noise_reduce, self.normalize = bool(noise_reduce), bool(normalize)

    def frames(self, n: int) -> int:
        return N.lib().osb_logmel_frames(n)

    def __call__(self, pcm: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if pcm.dtype != torch.int16 or not pcm.is_cuda or pcm.dim() != 2 or not pcm.is_contiguous():
            raise ValueError("pcm must be a contiguous CUDA int16 tensor [batch, samples]")
        b, n = pcm.shape
        if out is None:
            out = torch.empty((b, self.n_mels, self.frames(n)), dtype=torch.float32, device=pcm.device)
        N.call("osb_stt_frontend_dev", pcm.data_ptr(), n, b, n, self.sample_rate, int(self.noise_reduce), int(self.normalize),
               self.n_mels, out.data_ptr(), _stream())
        return out

    def run_host(self, pcm_host: torch.Tensor, pcm_dev: torch.Tensor, mel_dev: torch.Tensor, mel_host: torch.Tensor) -> torch.Tensor:
        """End-to-end step with HOST buffers (pinned): H2D of the clips, kernels, D2H of the features."""
        pcm_dev.copy_(pcm_host, non_blocking=True)
        self(pcm_dev, mel_dev)
        mel_host.copy_(mel_dev, non_blocking=True)
        return mel_host


class SttFull:
    """The whole STT front-end as one device-resident call (north_star chain): wire audio [B, n_in] (pcm16 | G.711 at
    ``from_rate``) -> 16 kHz -> { VAD probabilities + segments | [spectral gate] -> [normalise] -> requantise -> log-mel }.

    ``linear_chunk`` = 0: whole-clip polyphase resample (resample_pcm16); k > 0: every k input samples are resampled on
    their own with np.interp arithmetic, as the realtime door does per append.  Results stay on the device.
    """

    def __init__(self, session=None, fmt: str = "g711_ulaw", from_rate: int = 8000, linear_chunk: int = 0, n_mels: int = 128,
                 noise_reduce: bool = True, normalize: bool = True, threshold: float = 0.5, min_speech_ms: int = 250, silence_ms: int = 800,
                 keep_pcm: bool = False):
        N.require_gpu()
        self.session = session
        self.fmt = {"g711_ulaw": N.FMT_ULAW, "g711_alaw": N.FMT_ALAW, "pcm16": N.FMT_PCM16}[fmt]
        self.in_dtype = torch.int16 if self.fmt == N.FMT_PCM16 else torch.uint8
        self.from_rate, self.linear_chunk, self.n_mels = from_rate, linear_chunk, n_mels
        self.noise_reduce, self.normalize = bool(noise_reduce), bool(normalize)
        self.threshold, self.min_speech_ms, self.silence_ms, self.keep_pcm = threshold, min_speech_ms, silence_ms, keep_pcm
        self._bufs = {}

    def samples_16k(self, n_in: int) -> int:
        return int(N.lib().osb_stt_full_samples(n_in, self.from_rate, self.linear_chunk))

    def buffers(self, batch: int, n_in: int, device) -> dict:
        key = (batch, n_in, str(device))
        if key not in self._bufs:
            n16 = self.samples_16k(n_in)
            n_win, max_seg = n16 // 512, n16 // 512 // 2 + 2
            b = {"n16": n16, "n_win": n_win, "max_seg": max_seg,
                 "mel": torch.empty((batch, self.n_mels, N.lib().osb_logmel_frames(n16)), dtype=torch.float32, device=device),
                 "probs": torch.empty((batch, max(n_win, 1)), dtype=torch.float32, device=device),
                 "segments": torch.empty((batch, max_seg, 2), dtype=torch.int32, device=device),
                 "counts": torch.zeros((batch,), dtype=torch.int32, device=device),
                 "pcm16k": torch.empty((batch, n16), dtype=torch.int16, device=device) if self.keep_pcm else None}
            self._bufs[key] = b
        return self._bufs[key]

    def __call__(self, wire: torch.Tensor) -> dict:
        if wire.dtype != self.in_dtype or not wire.is_cuda or wire.dim() != 2 or not wire.is_contiguous():
            raise ValueError("wire must be a contiguous CUDA tensor [batch, samples] of the wire dtype")
        batch, n_in = wire.shape
        b = self.buffers(batch, n_in, wire.device)
        vad = self.session.handle if self.session is not None else None
        N.call("osb_stt_full_dev", vad, wire.data_ptr(), self.fmt, self.from_rate, n_in, batch, n_in, self.linear_chunk, int(self.noise_reduce),
               int(self.normalize), self.n_mels, float(self.threshold), self.min_speech_ms, self.silence_ms,
               b["pcm16k"].data_ptr() if b["pcm16k"] is not None else None, b["probs"].data_ptr() if vad else None,
               b["segments"].data_ptr() if vad else None, b["counts"].data_ptr() if vad else None, b["max_seg"], b["mel"].data_ptr(), _stream())
        return b

    def run_host(self, wire_host, out: dict) -> dict:
        """End to end with HOST buffers: numpy / pinned torch rows in, results into the caller's host arrays
        (out: probs [B, n_win] f32, segments [B, max_seg, 2] i32, counts [B] i32, mel [B, n_mels, frames] f32)."""
        batch, n_in = wire_host.shape
        vad = self.session.handle if self.session is not None else None
        N.call("osb_stt_full_host", vad, N.ptr(wire_host), self.fmt, self.from_rate, n_in, batch, self.linear_chunk, int(self.noise_reduce),
               int(self.normalize), self.n_mels, float(self.threshold), self.min_speech_ms, self.silence_ms, N.ptr(out["probs"]) if vad else None,
               N.ptr(out["segments"]) if vad else None, N.ptr(out["counts"]) if vad else None, int(out["segments"].shape[1]), N.ptr(out["mel"]))
        return out


class VadBatch:
    """Silero-shaped VAD over a batch of equal-length streams resident on the device (BASELINE config 2)."""

    def __init__(self, session=None, threshold: float = 0.5, min_speech_ms: int = 250, silence_ms: int = 800, max_segments: int = 4096):
        from .vad.silero import VadSession

        N.require_gpu()
        self.session = session if session is not None else VadSession()
        self.threshold, self.min_speech_ms, self.silence_ms, self.max_segments = threshold, min_speech_ms, silence_ms, max_segments

    def score(self, pcm: torch.Tensor, state: torch.Tensor | None = None):
        b, n = pcm.shape
        n_win = n // 512
        if state is None:
            state = torch.zeros((b, 2, 128), dtype=torch.float32, device=pcm.device)
        probs = torch.empty((b, max(n_win, 1)), dtype=torch.float32, device=pcm.device)
        fmt = N.FMT_PCM16 if pcm.dtype == torch.int16 else N.FMT_F32
        N.call("osb_vad_score_dev", self.session.handle, pcm.data_ptr(), fmt, n, b, n, state.data_ptr(), probs.data_ptr(), max(n_win, 1), _stream())
        return probs[:, :n_win], state

    def segments(self, probs: torch.Tensor, n_samples: int):
        b, n_win = probs.shape
        segs = torch.empty((b, self.max_segments, 2), dtype=torch.int32, device=probs.device)
        counts = torch.zeros((b,), dtype=torch.int32, device=probs.device)
        N.call("osb_vad_segments_dev", probs.data_ptr(), probs.stride(0), n_win, b, n_samples, float(self.threshold), self.min_speech_ms,
               self.silence_ms, segs.data_ptr(), counts.data_ptr(), self.max_segments, _stream())
        return segs, counts

    def __call__(self, pcm: torch.Tensor):
        probs, state = self.score(pcm)
        segs, counts = self.segments(probs, pcm.shape[1])
        return probs, segs, counts


class RealtimeTick:
    """One 20 ms tick of BASELINE config 3: S streams x 160 G.711 bytes -> 320 pcm16 @16 kHz each, one launch."""

    def __init__(self, n_streams: int, chunk: int = 160, fmt: str = "g711_ulaw", from_rate: int = 8000, to_rate: int = 16000):
        N.require_gpu()
        self.n_streams, self.chunk = n_streams, chunk
        self.fmt = {"g711_ulaw": N.FMT_ULAW, "g711_alaw": N.FMT_ALAW, "pcm16": N.FMT_PCM16}[fmt]
        self.n_out = int(chunk * (to_rate / from_rate))

    def __call__(self, tick: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.n_streams, self.n_out), dtype=torch.int16, device=tick.device)
        N.call("osb_resample_linear_dev", tick.data_ptr(), self.fmt, out.data_ptr(), N.FMT_PCM16, self.chunk, self.n_out, self.n_streams,
               self.chunk, self.n_out, _stream())
        return out


class TtsPost:
    """Ragged batch of float32 utterances: trim + peak-normalise -> effects chain -> int16 (BASELINE config 5)."""

    def __init__(self, sample_rate: int = 24000, effects=None, trim: bool = True, normalize: bool = True):
        from .effects.chain import encode_effects

        N.require_gpu()
        self.sample_rate, self.trim, self.normalize = sample_rate, trim, normalize
        self.fx_types, self.fx_p0, self.fx_p1 = encode_effects(effects)

    def __call__(self, flat: torch.Tensor, offsets: torch.Tensor, lens: torch.Tensor, max_len: int,
                 out_pcm: torch.Tensor | None = None):
        """flat f32 [total] (device), offsets/lens int64 [B] (device) -> (int16 [total], new lens int64 [B])."""
        b, total = offsets.numel(), flat.numel()
        new_lens = torch.empty_like(lens)
        if out_pcm is None:
            out_pcm = torch.empty(total, dtype=torch.int16, device=flat.device)
        N.call("osb_tts_post_fx_dev", flat.data_ptr(), offsets.data_ptr(), lens.data_ptr(), b, int(max_len), total, int(self.trim),
               int(self.normalize), 0.01, 0.95, self.sample_rate, N.ptr(self.fx_types), N.ptr(self.fx_p0), N.ptr(self.fx_p1),
               len(self.fx_types), 0, new_lens.data_ptr(), out_pcm.data_ptr(), 1, _stream())  # 0: intermediate from the library's pool
        return out_pcm, new_lens

    @staticmethod
    def pack(utts: list[np.ndarray]):
        """list of f32 arrays -> (flat f32, offsets int64, lens int64); starts aligned to 4 samples."""
        lens = np.array([len(u) for u in utts], dtype=np.int64)
        padded = (lens + 3) // 4 * 4
        offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
        flat = np.zeros(int(padded.sum()), dtype=np.float32)
        for o, u in zip(offsets, utts):
            flat[o:o + len(u)] = u
        return flat, offsets, lens

    def run_numpy(self, utts: list[np.ndarray]):
        """Convenience: list of f32 arrays in, list of int16 arrays out (one H2D, one D2H for the whole batch)."""
        flat, offsets, lens = self.pack(utts)
        pcm, new_lens = self(torch.from_numpy(flat).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(lens).cuda(), int(lens.max()))
        torch.cuda.synchronize()
        pcm, new_lens = pcm.cpu().numpy(), new_lens.cpu().numpy()
        return [pcm[o:o + n] for o, n in zip(offsets, new_lens)], new_lens.tolist()


class RealtimeTickGraph:
    """A whole realtime tick as ONE CUDA graph launch: pinned host slot -> H2D -> decode+resample kernel -> D2H -> pinned host slot.

    The per-tick work is launch-latency bound (0.8 MB); replaying a captured graph removes two of the three
    driver submissions.  Usage: write the tick's bytes into ``host_in`` (pinned), call ``run()``, read ``host_out``.
    """

    def __init__(self, n_streams: int, chunk: int = 160, fmt: str = "g711_ulaw", from_rate: int = 8000, to_rate: int = 16000):
        self.tick = RealtimeTick(n_streams, chunk, fmt, from_rate, to_rate)
        in_dtype = torch.int16 if fmt == "pcm16" else torch.uint8
        self.host_in = torch.empty((n_streams, chunk), dtype=in_dtype).pin_memory()
        self.host_out = torch.empty((n_streams, self.tick.n_out), dtype=torch.int16).pin_memory()
        self.dev_in = torch.empty((n_streams, chunk), dtype=in_dtype, device="cuda")
        self.dev_out = torch.empty((n_streams, self.tick.n_out), dtype=torch.int16, device="cuda")
        self.stream = torch.cuda.Stream()
        self.host_in.zero_()
        with torch.cuda.stream(self.stream):
            self._body()  # warm-up outside capture (lazy init inside the library)
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self._body()

    def _body(self):
        self.dev_in.copy_(self.host_in, non_blocking=True)
        self.tick(self.dev_in, self.dev_out)
        self.host_out.copy_(self.dev_out, non_blocking=True)

    def run(self) -> torch.Tensor:
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        self.stream.synchronize()
        return self.host_out


def shard_units(n_units: int, world: int, rank: int) -> range:
    """Static sharding of independent clips / streams / utterances: contiguous, balanced to +-1 unit.
    No exchange step exists on this path, so there is no collective."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_by_length(lengths, world: int) -> list[list[int]]:
    """Length-balanced static sharding (longest-processing-time-first greedy) for ragged units."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    bins = [[] for _ in range(world)]
    load = [0] * world
    for i in order:
        k = min(range(world), key=lambda j: load[j])
        bins[k].append(i)
        load[k] += int(lengths[i])
    return [sorted(b) for b in bins]
