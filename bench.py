#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native open-speech audio hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stt_batch|stt_full|c1|vad|realtime|tts] [--impl reference] [--no-extra]

Default (headline) workload = BASELINE.json configs[3], the configuration the metric's target is quoted on: batch STT
front-end, 256 x 60 s 16 kHz pcm16 clips per GPU -> spectral-gating noise reduction -> RMS normalise -> int16 requantise
-> Whisper large-v3 128-bin log-mel.  One "step" = one pass over the batch.  The same JSON line carries, under
"configs", compact results for the other BASELINE shapes (full STT front-end with decode + resample + VAD, configs[0] as
a batch, configs[1] VAD over 256 x 1 h, configs[2] realtime ticks with events, configs[4] TTS post-processing), so that
every named shape is timed by the driver at every N.  `--workload X` prints X's own full line instead.

  value     audio-seconds per second, inputs already resident in HBM (CUDA events, barrier both sides, max over ranks)
  e2e       same metric through the C-ABI host entry point: pinned HOST buffers in, HOST results out, H2D and D2H inside
            the timed region; copy_floor_ms = the same copy schedule with no kernels; pageable_ms = numpy (pageable) buffers
  roofline  frac = the step's SURVEY 8(d) algorithmic bytes / step time / measured HBM peak; kernel_frac = the dominant
            kernel's own input + output bytes / its CUDA-event time / the same peak
  cpu_baseline / clocks: DESIGN.md "Measurement"

--impl reference times the reference's CPU implementation of the same chain (the numpy/scipy oracle port: the reference is
pure Python around numpy/scipy/noisereduce/faster-whisper/onnxruntime, four of which cannot be installed here) on all host
cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "audio-seconds/sec"
UNIT = "audio-s/s"
SR = 16000
DISTINCT = 8  # distinct synthetic clips per rank, tiled to the batch (distinct addresses: the traffic is the full batch's)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- workloads
STT = {
    # name: description, clips per GPU, seconds per clip, noise_reduce, normalize, n_mels, full chain (decode + resample + VAD in front)
    "stt_batch": ("BASELINE configs[3]: batch STT front-end, 256x60 s clips/GPU: spectral-gate denoise + normalise + 128-bin log-mel", 256, 60.0, True, True, 128, False),
    "c1": ("BASELINE configs[0] as a batch: 256x30 s clips/GPU: normalise + 128-bin log-mel (no denoise)", 256, 30.0, False, True, 128, False),
    "stt_full": ("north_star full STT front-end, 256x60 s recordings/GPU: G.711 mu-law 8 kHz -> pcm16 -> 16 kHz (np.interp per 20 ms chunk, the realtime door) "
                 "-> Silero-shaped VAD scoring + segmenting -> spectral-gate denoise + normalise + 128-bin log-mel, one call", 256, 60.0, True, True, 128, True),
}
# SURVEY.md 8(d) algorithmic bytes per audio-second (compulsory input + output, no intermediates).
#   stt_batch / c1: 32,000 (pcm16 in) + 51,200 (f32 [128][100] out)
#   stt_full: 8,000 (mu-law in) + 51,200 (features) + 125 (31.25 probabilities) -- the survey has no row for the composed chain;
#             same rule applied to its wire input and its outputs
ALG_BYTES_PER_AUDIO_S = {"stt_batch": 83200.0, "c1": 83200.0, "stt_full": 59325.0}
VAD_ALG_BYTES_PER_AUDIO_S = 32125.0
TTS_ALG_BYTES_PER_AUDIO_S = 192000.0


def kernel_io_bytes_per_audio_s(name: str, seconds: float, denoised_input: bool) -> float | None:
    """The kernel's OWN compulsory input + output bytes per audio-second in the decomposition that ships:
    what it would move if every byte crossed HBM exactly once.  Spectral-gate cells: 513 bins x the frames of noisereduce's
    600,000-sample chunks with 30,000 samples of context either side, hop 256 (frames wholly past the clip end are skipped)."""
    n = seconds * SR
    chunks = max(1, int(np.ceil(n / 600000.0)))
    frames = (n + 60000.0 * chunks) / 256.0 + chunks
    cells = 513.0 * frames / seconds
    table = {
        "k_nr_stft": 2.0 * SR + 12.0 * cells,             # pcm16 in; spectrum (8 B) + magnitude (4 B) out
        "k_nr_carry": 2.0 * cells / 16.0 * 8.0,           # per-16-frame tile aggregates in, filter states out (f64)
        "k_nr_mask": 4.0 * cells + 4.0 * cells,           # magnitude in, smoothed mask out
        "k_nr_istft": 12.0 * cells + 4.0 * SR,            # spectrum + mask in, float32 samples out
        "k_logmel": (4.0 if denoised_input else 2.0) * SR + 51200.0,
        "k_logmel_finalize": 2.0 * 51200.0,
        "k_sumsq_pcm16": 2.0 * SR,
        "k_resample_linear": 1.0 * SR / 2 + 2.0 * SR,     # mu-law 8 kHz in, pcm16 16 kHz out
        "k_vad_front_fused": 2.0 * SR + 31.25 * 2048.0,   # pcm16 in, gate pre-activations [512] f32 per window out
        "k_vad_recur": 31.25 * 2048.0 + 31.25 * 4.0,      # pre-activations in, probabilities out
    }
    bare = name.strip("()").split("::")[-1]  # OSB_LAUNCH stringifies template kernels as "(k_x<..>)"; the fused front lives in vf::
    for k, v in table.items():
        if bare.startswith(k):
            return v
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- synthetic inputs
def make_clips(workload: str, clips: int, seconds: float, rank: int, nr: bool):
    """int16 [clips, n] @16 kHz for the batch chains (the full chain takes ulaw_clips)."""
    from open_speech_b200 import synth

    return synth.clip_batch_pcm16(clips, seconds, seed=synth.SEED_C4 + 1000 * rank, extra_noise_rms=0.01 if nr else 0.0, distinct=DISTINCT)


def ulaw_clips(clips: int, seconds: float, rank: int) -> np.ndarray:
    """mu-law 8 kHz recordings: speech-like 8 kHz pcm16 encoded ON THE GPU by the library's own lin2ulaw kernel
    (bit-exact with audioop, tests/test_gpu_codec_resample.py) -- the bench's GPU arm never imports oracle/."""
    import torch

    from open_speech_b200 import _native as N
    from open_speech_b200 import synth

    n8 = int(seconds * 8000)
    base = np.stack([synth.clip_pcm16(seconds, sr=8000, seed=synth.SEED_C4 + 1000 * rank + i, extra_noise_rms=0.01) for i in range(min(DISTINCT, clips))])
    d = torch.from_numpy(base).cuda()
    out = torch.empty(d.shape, dtype=torch.uint8, device="cuda")
    N.call("osb_g711_encode_dev", d.data_ptr(), out.data_ptr(), d.numel(), N.FMT_ULAW, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ul = out.cpu().numpy()
    return np.ascontiguousarray(ul[np.arange(clips) % ul.shape[0]]).reshape(clips, n8)


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_one(args):
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import stt

    kind, data, nr, norm, n_mels = args
    t = time.perf_counter()
    if kind == "full":
        from oracle import vad as ovad

        stt.stt_full(data, "g711_ulaw", 8000, linear_chunk=160, noise_reduce=nr, normalize=norm, n_mels=n_mels, net=ovad.SileroNet())
    else:
        stt.stt_frontend(data, noise_reduce=nr, normalize=norm, n_mels=n_mels)
    return time.perf_counter() - t


def cpu_baseline(workload: str, cores: int, n_clips: int, seconds: float) -> dict:
    """Oracle port of the reference chain on `cores` host processes, bounded sample of the same workload."""
    from open_speech_b200 import synth

    _, _, _, nr, norm, n_mels, full = STT[workload]
    if full:
        from oracle import codec

        base = [codec.lin2ulaw(synth.clip_pcm16(seconds, sr=8000, seed=synth.SEED_C4 + i, extra_noise_rms=0.01).tobytes()) for i in range(min(n_clips, DISTINCT))]
        jobs = [("full", base[i % len(base)], nr, norm, n_mels) for i in range(n_clips)]
        warm = ("full", base[0][: 8000 * 2], nr, norm, n_mels)
    else:
        clips = synth.clip_batch_pcm16(n_clips, seconds, seed=synth.SEED_C4, distinct=min(n_clips, DISTINCT))
        jobs = [("batch", clips[i], nr, norm, n_mels) for i in range(n_clips)]
        warm = ("batch", clips[0][: SR * 2], nr, norm, n_mels)
    if cores <= 1:
        _cpu_one(warm)  # imports, FFT plans
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_one(j)
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_one, [warm] * cores)  # warm every worker
            t0 = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=1)
            dt = time.perf_counter() - t0
    what = ("audioop + np.interp + Silero-shaped network in numpy + noisereduce + faster-whisper FeatureExtractor restated" if full
            else "noisereduce + faster-whisper FeatureExtractor restated")
    return {"value": n_clips * seconds / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_clips} x {seconds:g} s clips of the same synthetic workload, numpy/scipy oracle port ({what}; the reference itself is "
                      f"numpy/scipy around those packages), {dt:.2f} s of wall time"}


TTS_FX = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
OTHER_CPU = {
    # workload: (jobs for one host thread on the default line, audio-seconds per job, what one job is)
    "vad": (4, 60.0, "60 s streams: Silero-shaped network restated in numpy (front batched over the windows, LSTM cell per window; the reference "
                     "calls onnxruntime once per window, src/vad/silero.py:63-91) + get_speech_segments"),
    "realtime": (256, 64 * 0.020, "streams x 64 ticks of 20 ms: per tick decode_audio_to_pcm16(g711_ulaw -> 16 kHz) + buffer append + gate machine "
                                  "(src/realtime/server.py:127-170, audio_buffer.py:37-58, :111-154; no full VAD window in 20 ms, as in the reference)"),
    "tts": (16, None, "utterances of the same synthetic batch: process_tts_chunks(trim, normalize) + apply_chain([normalize, reverb medium, podcast_eq, "
                      "robot]) + float32_to_int16 (src/audio/postprocessing.py, src/effects/chain.py, src/tts/pipeline.py)"),
}


def _cpu_other_one(args):
    """One job of the oracle port of configs[1] / [2] / [4]; returns its wall time."""
    import warnings

    warnings.filterwarnings("ignore")
    kind, data = args
    t = time.perf_counter()
    if kind == "vad":
        from oracle import vad as ovad

        if not hasattr(_cpu_other_one, "net"):
            _cpu_other_one.net = ovad.SileroNet()  # weights made once per process (the reference loads its model once, too)
        probs, _ = _cpu_other_one.net.score_stream(data.astype(np.float32) / 32768.0)
        ovad.segments_from_probs(probs, len(data))
    elif kind == "realtime":
        from oracle import codec
        from oracle import vad as ovad

        buf = bytearray()
        for k in range(data.shape[0]):
            buf.extend(codec.decode_audio_to_pcm16(data[k].tobytes(), "g711_ulaw", 16000))
        ovad.input_buffer_events([0.0] * data.shape[0], [2 * data.shape[1]] * data.shape[0], 0.5, 500)
    else:
        from oracle import tts as otts

        otts.tts_chain([data], TTS_FX)
    return time.perf_counter() - t


def cpu_baseline_other(workload: str, cores: int, n_jobs: int = 0) -> dict:
    """Oracle port of configs[1] (vad), [2] (realtime) or [4] (tts) on `cores` host processes, bounded sample of the same synthetic workload."""
    from open_speech_b200 import synth

    dflt, job_s, what = OTHER_CPU[workload]
    n_jobs = n_jobs or dflt
    if workload == "vad":
        base = [synth.clip_pcm16(job_s, seed=synth.SEED_C2 + 1 + i) for i in range(min(n_jobs, DISTINCT))]
        jobs = [("vad", base[i % len(base)]) for i in range(n_jobs)]
        warm, audio_s = ("vad", base[0][: SR * 2]), n_jobs * job_s
    elif workload == "realtime":
        data = synth.ulaw_streams(min(n_jobs, 16), 64)                       # [64 ticks, streams, 160 bytes]
        jobs = [("realtime", np.ascontiguousarray(data[:, i % data.shape[1]])) for i in range(n_jobs)]
        warm, audio_s = ("realtime", jobs[0][1][:4]), n_jobs * job_s
    else:
        utts = synth.tts_batch(n_jobs, seed=synth.SEED_C5, distinct=32)
        jobs = [("tts", u) for u in utts]
        warm, audio_s = ("tts", utts[0][:24000]), sum(len(u) for u in utts) / 24000.0
    if cores <= 1:
        _cpu_other_one(warm)
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_other_one(j)
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_other_one, [warm] * cores)
            t0 = time.perf_counter()
            pool.map(_cpu_other_one, jobs, chunksize=max(1, len(jobs) // (4 * cores)))
            dt = time.perf_counter() - t0
    return {"value": audio_s / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_jobs} {what}; numpy/scipy oracle port, {audio_s:.0f} audio-s in {dt:.2f} s of wall time"}


def run_reference(args) -> None:
    """--impl reference: the reference's CPU path (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload in OTHER_CPU:
        per_core = {"vad": 2, "realtime": 256, "tts": 8}[args.workload]
        vals, base = [], None
        for i in range(args.warmup + args.steps):
            base = cpu_baseline_other(args.workload, cores, per_core * cores)
            if i >= args.warmup:
                vals.append(base["value"])
        v = float(np.mean(vals))
        base["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
                          "config": {"workload": f"BASELINE {args.workload} config, bounded sample: " + OTHER_CPU[args.workload][2], "host_cores": cores},
                          "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    wl = args.workload if args.workload in STT else "stt_batch"
    desc, clips_per_gpu, seconds, nr, norm, n_mels, _full = STT[wl]
    n_clips = max(cores, 8)
    sample_s = 60.0 if seconds >= 60.0 else seconds
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(wl, cores, n_clips, sample_s)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * n_clips * sample_s / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32",
            "data": "synthetic", "config": {"workload": desc, "clips_per_step": n_clips, "seconds_per_clip": sample_s, "host_cores": cores,
                                            "note": f"bounded sample ({n_clips} clips per step instead of {clips_per_gpu} per GPU) of the GPU arm's workload; the metric "
                                                    "is per audio-second and CPU throughput does not depend on the batch size"},
            "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, local: int) -> dict:
    """Run this rank (and first-touch its pinned staging buffers) on the CPUs next to its GPU: with several ranks per box the
    host<->device copies of the e2e leg otherwise cross the socket interconnect.  NVML's affinity first, sysfs second; the
    outcome (or the reason nothing was bound) goes into the JSON line."""
    allowed = os.sched_getaffinity(0)
    info = {"method": None, "numa_node": None, "cpus_bound": None, "cpus_allowed": len(allowed)}
    try:
        import pynvml

        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        h = None
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0".encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & allowed
        try:
            info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            info.update(method="nvml cpu affinity", cpus_bound=len(cpus))
            return info
        info["method"] = "nvml: the GPU's affinity covers every allowed CPU (single NUMA domain visible)"
        return info
    except Exception as e:
        info["method"] = f"nvml unavailable ({type(e).__name__})"
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            info["method"] += "; sysfs numa_node = -1 (the VM exposes no NUMA topology)"
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(method="sysfs numa_node", numa_node=node, cpus_bound=len(cpus))
    except Exception as e:
        info["method"] += f"; sysfs unavailable ({type(e).__name__})"
    return info


# ----------------------------------------------------------------------------- GPU arm plumbing
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        from open_speech_b200 import _native as N

        self.torch, self.dist, self.N = torch, dist, N
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:  # the copy threads of the pageable-buffer staging (csrc/host_stage.cu) share the box's cores with the other ranks
            os.environ.setdefault("OSB_COPY_THREADS", str(max(1, min(8, (os.cpu_count() or 8) // self.world))))
        N.require_gpu()  # no device: RuntimeError here, there is nothing to fall back to
        torch.cuda.set_device(self.local)
        self.numa = bind_to_gpu_numa_node(torch, self.local)
        N.check(N.lib().osb_init(self.local))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.peak, self.peak_src = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps: int, warmup: int = 3) -> float:
        """ms per step: CUDA events on the current stream, barrier + synchronize both sides, max over ranks."""
        torch = self.torch
        self.settle(fn, warmup)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = self.max_over_ranks(ev0.elapsed_time(ev1) / steps)
        self.barrier()
        return ms

    def settle(self, fn, warmup: int, extra: int = 12) -> int:
        """`warmup` untimed calls, one at a time; then up to `extra` more until two consecutive calls agree within 3 %.

        The first calls of a process are not representative for longer than three calls: the driver maps (and scrubs) device memory for
        the stream-ordered pools lazily, call by call (measured on a fresh box: 2114, 24.4, 17.2, 14.1, 13.6, 11.1, 11.1 ... ms)."""
        torch = self.torch
        prev, n = None, 0
        for i in range(warmup + extra):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            n += 1
            if i + 1 >= warmup and prev is not None and abs(dt - prev) <= 0.03 * dt:
                break
            prev = dt
        return n

    def wall(self, fn, steps: int, warmup: int = 2) -> float:
        """ms per step of a synchronous host-to-host call, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = self.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
        self.barrier()
        return ms

    def profile(self, fn, steps: int) -> dict:
        """per-kernel CUDA-event ms per step (library hook, events on the launching stream)."""
        N = self.N
        N.check(N.lib().osb_profile_enable(1))
        for _ in range(steps):
            fn()
        buf = ctypes.create_string_buffer(1 << 16)
        N.check(N.lib().osb_profile_report(buf, len(buf)))
        N.check(N.lib().osb_profile_enable(0))
        return {k: {"ms": v["ms"] / steps, "launches": v["launches"] / steps} for k, v in json.loads(buf.value.decode()).items()}

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _vad_session():
    from open_speech_b200.vad.silero import VadSession

    return VadSession()  # seeded random-init weights (BASELINE configs[1] allows; silero-vad is not installed)


# ----------------------------------------------------------------------------- STT chains (stt_batch, c1, stt_full)
def stt_setup(ctx: Ctx, workload: str, clips: int):
    torch, N = ctx.torch, ctx.N
    from open_speech_b200.batch import SttFrontEnd, SttFull

    desc, dclips, seconds, nr, norm, n_mels, full = STT[workload]
    clips = clips or dclips
    if full:
        host_np = ulaw_clips(clips, seconds, ctx.rank)
        op = SttFull(_vad_session(), fmt="g711_ulaw", from_rate=8000, linear_chunk=160, n_mels=n_mels, noise_reduce=nr, normalize=norm)
    else:
        host_np = make_clips(workload, clips, seconds, ctx.rank, nr)
        op = SttFrontEnd(n_mels=n_mels, sample_rate=SR, noise_reduce=nr, normalize=norm)
    wire_host = torch.from_numpy(host_np).pin_memory()
    wire_dev = wire_host.cuda()
    n16 = int(seconds * SR)
    nf = N.lib().osb_logmel_frames(n16)
    if full:
        run = lambda: op(wire_dev)  # noqa: E731
    else:
        mel_dev = torch.empty((clips, n_mels, nf), dtype=torch.float32, device="cuda")
        run = lambda: op(wire_dev, mel_dev)  # noqa: E731
    return dict(desc=desc, clips=clips, seconds=seconds, nr=nr, norm=norm, n_mels=n_mels, full=full, op=op, host_np=host_np, wire_host=wire_host,
                wire_dev=wire_dev, run=run, n16=n16, nf=nf, audio_s=clips * seconds)


def stt_e2e(ctx: Ctx, s: dict, steps: int, with_floor: bool):
    """host buffers in -> host results out through the C ABI (copies inside the timed region)."""
    torch, N = ctx.torch, ctx.N
    clips, n_mels, nf, n16 = s["clips"], s["n_mels"], s["nf"], s["n16"]
    mel_host = torch.empty((clips, n_mels, nf), dtype=torch.float32).pin_memory()
    h2d = s["wire_host"].numel() * s["wire_host"].element_size()
    d2h = mel_host.numel() * 4
    if s["full"]:
        n_win, max_seg = n16 // 512, n16 // 512 // 2 + 2
        out = {"probs": torch.empty((clips, n_win), dtype=torch.float32).pin_memory(), "segments": torch.empty((clips, max_seg, 2), dtype=torch.int32).pin_memory(),
               "counts": torch.empty((clips,), dtype=torch.int32).pin_memory(), "mel": mel_host}
        d2h += sum(out[k].numel() * 4 for k in ("probs", "segments", "counts"))
        step = lambda: s["op"].run_host(s["wire_host"], out)  # noqa: E731
    else:
        n = s["wire_host"].shape[1]
        step = lambda: N.call("osb_stt_frontend_host", s["wire_host"].data_ptr(), n, clips, n, SR, int(s["nr"]), int(s["norm"]), n_mels, mel_host.data_ptr())  # noqa: E731
    ms = ctx.wall(step, steps)
    res = {"value": ctx.world * s["audio_s"] / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms,
           "checksum": float(mel_host[0, :, :16].double().sum())}  # the D2H result is really read
    if with_floor:
        row_in = s["wire_host"].shape[1] * s["wire_host"].element_size()
        floor = lambda: N.call("osb_copy_floor_host", s["wire_host"].data_ptr(), row_in, mel_host.data_ptr(), d2h // clips, clips)  # noqa: E731
        fms = ctx.wall(floor, steps)
        res["copy_floor_ms"] = fms
        res["copy_floor_gbps"] = ctx.world * (h2d + d2h) / (fms / 1e3) / 1e9
        res["over_floor"] = ms / fms
        # the drop-in as a Python caller reaches it: numpy (pageable) arrays, same C entry
        page_in = np.array(s["host_np"], copy=True)
        page_mel = np.empty((clips, n_mels, nf), dtype=np.float32)
        if s["full"]:
            pout = {"probs": np.empty((clips, n16 // 512), np.float32), "segments": np.empty((clips, n16 // 512 // 2 + 2, 2), np.int32),
                    "counts": np.empty(clips, np.int32), "mel": page_mel}
            pstep = lambda: s["op"].run_host(page_in, pout)  # noqa: E731
        else:
            n = page_in.shape[1]
            pstep = lambda: N.call("osb_stt_frontend_host", N.ptr(page_in), n, clips, n, SR, int(s["nr"]), int(s["norm"]), n_mels, N.ptr(page_mel))  # noqa: E731
        res["pageable_ms"] = ctx.wall(pstep, max(1, steps // 2), warmup=1)
    return res, step, mel_host


def stt_roofline(ctx: Ctx, workload: str, s: dict, ms_step: float, kern: dict) -> dict:
    # dominant = largest share of the step's SM-time: the VAD recurrence runs BESIDE the feature branch on one SM per eight streams
    # (csrc/vad.cu k_vad_recur_tc), so its wall time counts with the fraction of the GPU it holds
    clips = s["audio_s"] / s["seconds"]

    def sm_share(name: str) -> float:
        return min(1.0, math.ceil(clips / 8.0) / 148.0) if "k_vad_recur_tc" in name else 1.0

    dom_name, dom = max(kern.items(), key=lambda kv: kv[1]["ms"] * sm_share(kv[0]))
    dom_ms = dom["ms"] / max(1.0, dom["launches"])
    alg = ALG_BYTES_PER_AUDIO_S[workload] * s["audio_s"]
    gbs = alg / (ms_step / 1e3) / 1e9
    kio = kernel_io_bytes_per_audio_s(dom_name, s["seconds"], s["nr"])
    kbytes = kio * s["audio_s"] / max(1.0, dom["launches"]) if kio else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(workload, {})
        traffic = tj.get(dom_name, tj.get(dom_name.strip("()")))
    r = {"bound": "hbm", "achieved": gbs, "peak": ctx.peak, "unit": "GB/s", "frac": gbs / ctx.peak, "traffic": traffic, "peak_source": ctx.peak_src,
         "algorithmic_bytes_per_step": alg, "what": "whole step: SURVEY 8(d) algorithmic bytes of the chain / step time",
         "kernel": dom_name, "kernel_ms_per_launch": dom_ms, "kernel_share_of_step": dom["ms"] / ms_step,
         "kernel_io_bytes_per_launch": kbytes,
         "kernel_frac": (kbytes / (dom_ms / 1e3) / 1e9 / ctx.peak) if kbytes else None,
         "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}}
    return r


def bench_stt(ctx: Ctx, args, workload: str, extra: dict | None) -> None:
    torch, N = ctx.torch, ctx.N
    s = stt_setup(ctx, workload, args.clips)
    warm = max(args.warmup, 3)
    warm = ctx.settle(s["run"], warm)
    sampler = ClockSampler(ctx.local)
    launches0 = N.lib().osb_launch_count()
    if ctx.rank == 0:
        sampler.start()
    ms_step = ctx.timed(s["run"], args.steps, warmup=0)
    clocks = sampler.stop() if ctx.rank == 0 else None
    launches = int(N.lib().osb_launch_count() - launches0)
    kern = ctx.profile(s["run"], 2)  # per-kernel event pairs outside the timed region (the events themselves cost launch slots)
    value = ctx.world * s["audio_s"] / (ms_step / 1e3)
    e2e, e2e_step, mel_host = stt_e2e(ctx, s, max(1, min(args.steps, 5)), with_floor=True)

    single = None
    if ctx.rank == 0 and not s["full"]:  # BASELINE configs[0] is literally a single clip: latency, not throughput
        n = s["wire_host"].shape[1]
        lat = []
        for _ in range(60):
            t0 = time.perf_counter()
            N.call("osb_stt_frontend_host", s["wire_host"].data_ptr(), n, 1, n, SR, int(s["nr"]), int(s["norm"]), s["n_mels"], mel_host.data_ptr())
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[10:])
        single = {"p50_ms": lat[len(lat) // 2], "p99_ms": lat[-1], "x_realtime_p50": s["seconds"] / (lat[len(lat) // 2] / 1e3)}
    if ctx.rank != 0:
        return
    cpu = None
    if not args.no_cpu:  # bounded sample, single thread = the reference's execution model (it runs this path on the event-loop thread)
        cpu = cpu_baseline(workload, 1, 8 if s["full"] else (24 if s["seconds"] >= 60 else 48), s["seconds"])
    h2d_mb = s["wire_host"].numel() * s["wire_host"].element_size() / 1e6
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": s["desc"], "clips_per_gpu": s["clips"], "seconds_per_clip": s["seconds"], "sample_rate": SR, "n_mels": s["n_mels"],
                       "noise_reduce": s["nr"], "normalize": s["norm"], "global_clips": s["clips"] * ctx.world,
                       "parallelism": f"clip-sharded x{ctx.world}, no collective", "host_binding_rank0": ctx.numa,
                       "clips": f"{DISTINCT} distinct seeded synthetic clips per rank, tiled to {s['clips']} (distinct addresses, full traffic)",
                       "l2": f"inputs + outputs larger than L2 ({h2d_mb:.0f} MB in, {s['clips'] * s['n_mels'] * s['nf'] * 4 / 1e6:.0f} MB out per GPU per step vs 126 MB)"},
            "clocks": clocks, "e2e": e2e, "single_clip_latency": single, "gpu_launches": launches,
            "roofline": stt_roofline(ctx, workload, s, ms_step, kern), "cpu_baseline": cpu}
    if extra is not None:
        line["configs"] = extra
    print(json.dumps(line), flush=True)


def compact_stt(ctx: Ctx, workload: str, steps: int) -> dict:
    s = stt_setup(ctx, workload, 0)
    ms = ctx.timed(s["run"], steps)
    kern = ctx.profile(s["run"], 1)
    e2e, _, _ = stt_e2e(ctx, s, 3, with_floor=False)
    alg = ALG_BYTES_PER_AUDIO_S[workload] * s["audio_s"]
    top = sorted(kern.items(), key=lambda kv: -kv[1]["ms"])[:8]
    return {"workload": s["desc"], "value": ctx.world * s["audio_s"] / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "e2e_value": e2e["value"], "e2e_ms_per_step": e2e["ms_per_step"], "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
            "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "roofline_frac": alg / (ms / 1e3) / 1e9 / ctx.peak,
            "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in top}}


# ----------------------------------------------------------------------------- BASELINE configs[1]: VAD
def vad_run(ctx: Ctx, streams: int, seconds: float, steps: int) -> dict:
    """Silero-shaped VAD scoring + segmenting: (i) `streams` independent streams of `seconds` each, (ii) one 1 h stream."""
    torch = ctx.torch
    from open_speech_b200 import synth
    from open_speech_b200.batch import VadBatch

    vb = VadBatch(_vad_session())
    reps = max(1, int(round(seconds / 600.0)))
    base = torch.from_numpy(synth.clip_pcm16(min(seconds, 600.0), seed=synth.SEED_C2 + 1 + ctx.rank)).cuda()
    many = base.repeat(reps).unsqueeze(0).repeat(streams, 1).contiguous()   # device-side tiling: plumbing, not the path
    ms_many = ctx.timed(lambda: vb(many), steps, warmup=2)
    kern = ctx.profile(lambda: vb(many), 1)
    audio_s = streams * many.shape[1] / SR
    del many
    torch.cuda.empty_cache()
    one = torch.from_numpy(synth.clip_pcm16(600.0, seed=synth.SEED_C2)).cuda().repeat(6).unsqueeze(0).contiguous()  # 1 h = 57.6 M samples
    ms_one = ctx.timed(lambda: vb(one), max(1, steps // 2), warmup=1)
    alg = VAD_ALG_BYTES_PER_AUDIO_S * audio_s
    front = sum(v["ms"] for k, v in kern.items() if "gemm" in k or "front" in k)
    recur = sum(v["ms"] for k, v in kern.items() if "recur" in k)
    return {"workload": f"BASELINE configs[1]: Silero-shaped VAD 512-sample scoring + segmenting, {streams} streams x {audio_s / streams:.0f} s per GPU "
                        "(seeded random-init weights: silero-vad is not installed)",
            "value": ctx.world * audio_s / (ms_many / 1e3), "unit": UNIT, "ms_per_step": ms_many, "steps": steps,
            "front_ms": front, "recurrence_ms": recur,
            "single_stream_1h": {"ms": ms_one, "x_realtime": 3600.0 / (ms_one / 1e3), "windows": 112500},
            "roofline_frac": alg / (ms_many / 1e3) / 1e9 / ctx.peak,
            "roofline_note": "compute + serial-latency bound by arithmetic (SURVEY 8(d)); the HBM fraction is reported for completeness",
            "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}}


# ----------------------------------------------------------------------------- BASELINE configs[2]: realtime ticks
def realtime_run(ctx: Ctx, S: int, ticks: int) -> dict:
    """1024 concurrent G.711 mu-law 8 kHz streams: host wire bytes in -> host pcm16 + speech events out, per tick.

    A  reference-exact: 20 ms chunks = 320 samples < 512, the reference's VAD scores nothing and returns 0.0
    B  40 ms chunks (640 samples = one VAD window), the network runs every tick
    P  the polyphase door of north_star: audioop.ulaw2lin -> resample_pcm16(8000 -> 16000), 40 ms chunks, network runs
    """
    torch = ctx.torch
    from open_speech_b200 import synth
    from open_speech_b200.realtime.gate import RealtimeGate

    sess = _vad_session()
    data = synth.ulaw_streams(S, 64)                     # [64, S, 160], cycled
    res = {}
    for tag, chunk, poly, n_ticks in (("A_20ms_reference_exact", 160, False, ticks), ("B_40ms_vad_scored", 320, False, ticks // 2),
                                      ("P_40ms_polyphase_vad_scored", 320, True, ticks // 2)):
        per = chunk // 160
        src = np.ascontiguousarray(data.reshape(64 // per, per, S, 160).transpose(0, 2, 1, 3).reshape(64 // per, S, chunk))
        host_in = torch.from_numpy(src).pin_memory()
        g = RealtimeGate(S, chunk, fmt="g711_ulaw", session=sess, threshold=0.5, silence_duration_ms=500, poly=poly)
        dev_in = torch.empty((S, chunk), dtype=torch.uint8, device="cuda")
        host_pcm = torch.empty((S, g.n_out), dtype=torch.int16).pin_memory()
        n_ev = 0

        def one(i):
            nonlocal n_ev
            dev_in.copy_(host_in[i % host_in.shape[0]], non_blocking=True)
            g.tick(dev_in)
            host_pcm.copy_(g.pcm, non_blocking=True)
            n_ev += len(g.read_events())  # D2H of the compact list + stream synchronize

        for i in range(20):
            one(i)
        n_ev = 0
        ctx.barrier()
        lat = np.empty(n_ticks)
        for i in range(n_ticks):
            t0 = time.perf_counter()
            one(i)
            lat[i] = (time.perf_counter() - t0) * 1e3
        p50, p99 = ctx.max_over_ranks(float(np.percentile(lat, 50))), ctx.max_over_ranks(float(np.percentile(lat, 99)))
        res[tag] = {"p50_ms": p50, "p99_ms": p99, "ticks": n_ticks, "events": n_ev, "chunk_ms": chunk / 8.0,
                    "x_realtime_p50": ctx.world * S * (chunk / 8000.0) / (p50 / 1e3)}
    # variant A as one CUDA graph launch per tick (pinned slot -> H2D -> resample + gate -> D2H pcm16 + events); the host memcpy of the
    # tick's bytes into the pinned slot is inside the timed region
    g = RealtimeGate(S, 160, fmt="g711_ulaw", session=sess, threshold=0.5, silence_duration_ms=500)
    gg = g.capture()
    host_in = torch.from_numpy(np.ascontiguousarray(data)).pin_memory()
    for i in range(20):
        gg.host_in.copy_(host_in[i % 64])
        gg.run()
    ctx.barrier()
    lat = np.empty(ticks)
    for i in range(ticks):
        t0 = time.perf_counter()
        gg.host_in.copy_(host_in[i % 64])
        gg.run()
        lat[i] = (time.perf_counter() - t0) * 1e3
    p50, p99 = ctx.max_over_ranks(float(np.percentile(lat, 50))), ctx.max_over_ranks(float(np.percentile(lat, 99)))
    res["A_20ms_cuda_graph"] = {"p50_ms": p50, "p99_ms": p99, "ticks": ticks, "chunk_ms": 20.0, "x_realtime_p50": ctx.world * S * 0.020 / (p50 / 1e3)}
    # variant A with the tick's buffers in pinned host memory, read and written by the kernels themselves (no copies): wire bytes in pinned
    # host memory in -> pcm16 + events in pinned host memory out, two launches and one synchronise per tick
    g = RealtimeGate(S, 160, fmt="g711_ulaw", session=sess, threshold=0.5, silence_duration_ms=500, host_io=True)
    n_ev = 0
    for i in range(20):
        g.tick_host(host_in[i % 64])
    ctx.barrier()
    lat = np.empty(ticks)
    for i in range(ticks):
        t0 = time.perf_counter()
        n_ev += len(g.tick_host(host_in[i % 64]))
        lat[i] = (time.perf_counter() - t0) * 1e3
    p50, p99 = ctx.max_over_ranks(float(np.percentile(lat, 50))), ctx.max_over_ranks(float(np.percentile(lat, 99)))
    res["A_20ms_host_io"] = {"p50_ms": p50, "p99_ms": p99, "ticks": ticks, "events": n_ev, "chunk_ms": 20.0,
                             "x_realtime_p50": ctx.world * S * 0.020 / (p50 / 1e3)}
    a = min((res["A_20ms_reference_exact"], res["A_20ms_host_io"]), key=lambda r: r["p50_ms"])
    return {"workload": f"BASELINE configs[2]: {S} G.711 mu-law 8 kHz streams per GPU, per tick: host bytes in -> decode -> 16 kHz -> buffer + VAD gate -> "
                        "host pcm16 + compact speech events out (device-resident per-stream state)",
            "value": a["x_realtime_p50"], "unit": UNIT, "ms_per_step": a["p50_ms"], "steps": ticks, "variants": res,
            "vad_semantics": "A: reference-exact (no full window in 20 ms: probability 0.0, no events possible); B / P: one window per 40 ms chunk is scored"}


# ----------------------------------------------------------------------------- BASELINE configs[4]: TTS post-processing
def tts_run(ctx: Ctx, B: int, steps: int) -> dict:
    torch, N = ctx.torch, ctx.N
    from open_speech_b200 import synth
    from open_speech_b200.batch import TtsPost

    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
    utts = synth.tts_batch(B, seed=synth.SEED_C5 + ctx.rank, distinct=32)
    post = TtsPost(24000, fx)
    flat, offsets, lens = post.pack(utts)
    audio_s = float(lens.sum()) / 24000.0
    d_flat, d_off, d_len = torch.from_numpy(flat).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(lens).cuda()
    out = torch.empty(flat.size, dtype=torch.int16, device="cuda")
    mx = int(lens.max())
    ms = ctx.timed(lambda: post(d_flat, d_off, d_len, mx, out), steps)
    kern = ctx.profile(lambda: post(d_flat, d_off, d_len, mx, out), 1)
    # voice blends: B blends of 2-3 packs out of 3 resident packs
    packs = torch.from_numpy(np.stack([p.reshape(-1) for p in synth.voice_packs(3)])).cuda()
    n = packs.shape[1]
    idx = torch.tensor([[0, 1, -1], [0, 1, -1], [0, 1, 2]] * (B // 3 + 1), dtype=torch.int32)[:B].contiguous().cuda()
    w = torch.tensor([[2 / 3, 1 / 3, 0], [0.5, 0.5, 0], [0.5, 1 / 3, 1 / 6]] * (B // 3 + 1), dtype=torch.float32)[:B].contiguous().cuda()
    bout = torch.empty((B, n), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms_blend = ctx.timed(lambda: N.call("osb_voice_blend_dev", packs.data_ptr(), n, idx.data_ptr(), w.data_ptr(), 3, B, bout.data_ptr(), st), steps)
    alg = TTS_ALG_BYTES_PER_AUDIO_S * audio_s
    # the three packs (1.5 MB) stay in L2: the HBM bytes of a blend batch are its OUTPUT only; the survey's (K+1) x 522,240 B per
    # blend counts reads that never reach HBM here, so it is reported separately and not against the HBM peak
    out_bytes = float(B) * n * 4
    return {"workload": f"BASELINE configs[4]: {B} Kokoro-shaped 24 kHz utterances per GPU ({audio_s:.0f} audio-s): trim + peak normalise + "
                        "[normalize, reverb medium, podcast_eq, robot] + int16; plus voice-style blends",
            "value": ctx.world * audio_s / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "roofline_frac": alg / (ms / 1e3) / 1e9 / ctx.peak,
            "voice_blend": {"blends_per_s": ctx.world * B / (ms_blend / 1e3), "hbm_GBps_output_only": out_bytes / (ms_blend / 1e3) / 1e9,
                            "frac_of_hbm": out_bytes / (ms_blend / 1e3) / 1e9 / ctx.peak,
                            "survey_bytes_GBps_incl_L2_resident_reads": float((idx >= 0).sum().item() + B) * n * 4 / (ms_blend / 1e3) / 1e9},
            "kernels_ms_per_step": {k: round(v["ms"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}}


def print_sub(ctx: Ctx, args, res: dict, dtype: str) -> None:
    if ctx.rank != 0:
        return
    if ctx.world == 1 and not args.no_cpu and args.workload in OTHER_CPU:  # a larger sample than the default line's compact leg
        res["cpu_baseline"] = cpu_baseline_other(args.workload, 1, 3 * OTHER_CPU[args.workload][0])
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": res.get("steps", args.steps), "warmup": 3,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": res.pop("workload")}, "roofline": {"bound": "hbm", "frac": res.get("roofline_frac"), "peak": ctx.peak, "unit": "GB/s",
                                                                       "peak_source": ctx.peak_src, "traffic": None},
            "cpu_baseline": res.pop("cpu_baseline", None), "detail": res}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="default", choices=["default"] + sorted(STT) + ["vad", "realtime", "tts"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=0, help="clips / streams / utterances per GPU (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="default workload only: skip the compact results of the other configs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    ctx = Ctx()
    try:
        if args.workload == "vad":
            print_sub(ctx, args, vad_run(ctx, args.clips or 256, 3600.0, max(1, args.steps // 3)), "f32 (bf16x3 tcgen05 front; recurrence: fp16 W_hh x fp16 hi+lo h on mma.sync, f32 accumulate and cell)")
        elif args.workload == "realtime":
            print_sub(ctx, args, realtime_run(ctx, args.clips or 1024, 3000), "u8/f64")
        elif args.workload == "tts":
            print_sub(ctx, args, tts_run(ctx, args.clips or 4096, args.steps), "f32/f64")
        elif args.workload in STT:
            bench_stt(ctx, args, args.workload, None)
        else:
            extra = None
            if not args.no_extra:
                torch = ctx.torch
                extra = {}
                for name, fn in (("stt_full", lambda: compact_stt(ctx, "stt_full", 5)), ("c1", lambda: compact_stt(ctx, "c1", 10)),
                                 ("vad", lambda: vad_run(ctx, 256, 3600.0, 2)), ("realtime", lambda: realtime_run(ctx, 1024, 3000)),
                                 ("tts", lambda: tts_run(ctx, 4096, 5))):
                    t0 = time.perf_counter()
                    extra[name] = fn()
                    extra[name]["bench_wall_s"] = round(time.perf_counter() - t0, 1)
                    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:  # the CPU path of every named shape beside it, one host thread
                        extra[name]["cpu_baseline"] = (cpu_baseline_other(name, 1) if name in OTHER_CPU
                                                       else cpu_baseline(name, 1, 4 if name == "stt_full" else 24, STT[name][2]))
                    torch.cuda.synchronize()
                    torch.cuda.empty_cache()
            bench_stt(ctx, args, "stt_batch", extra)
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
