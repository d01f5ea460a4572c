#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native open-speech audio hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload stt_batch|c1|vad|realtime|tts] [--impl reference]

Default workload = BASELINE.json configs[3], the configuration the metric's target is quoted on:
batch STT front-end, 256 x 60 s 16 kHz pcm16 clips per GPU -> spectral-gating noise reduction -> RMS normalise
-> int16 requantise -> Whisper large-v3 128-bin log-mel.  One "step" = one pass over the batch.

  value   audio-seconds per second, inputs already resident in HBM (CUDA events, max over ranks)
  e2e     same metric through the C-ABI host entry point: pinned HOST buffers in, HOST features out,
          H2D and D2H inside the timed region
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement"

--impl reference times the reference's CPU implementation of the same chain (the numpy/scipy oracle port: the
reference is pure Python around numpy/scipy/noisereduce/faster-whisper, three of which cannot be installed) on
all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
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


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- workloads
WORKLOADS = {
    # name: (description, clips per GPU, seconds per clip, noise_reduce, normalize, n_mels)
    "stt_batch": ("BASELINE configs[3]: batch STT front-end, 256x60 s clips/GPU: spectral-gate denoise + normalise + 128-bin log-mel", 256, 60.0, True, True, 128),
    "c1": ("BASELINE configs[0] as a batch: 256x30 s clips/GPU: normalise + 128-bin log-mel (no denoise)", 256, 30.0, False, True, 128),
}

# SURVEY.md 8(d) algorithmic bytes per audio-second (compulsory input + output, no intermediates)
ALG_BYTES_PER_AUDIO_S = {"stt_batch": 83200.0, "c1": 83200.0}
# denoise stage on its own: pcm16 in (2 B/sample) + f32 out (4 B/sample)
DENOISE_BYTES_PER_AUDIO_S = 6.0 * SR


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


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_one(args):
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import stt

    pcm, noise_reduce, normalize, n_mels = args
    t = time.perf_counter()
    stt.stt_frontend(pcm, noise_reduce=noise_reduce, normalize=normalize, n_mels=n_mels)
    return time.perf_counter() - t


def cpu_baseline(workload: str, cores: int, n_clips: int, seconds: float) -> dict:
    """Oracle port of the reference chain on `cores` host processes, bounded sample of the same workload."""
    from open_speech_b200 import synth

    _, _, _, nr, norm, n_mels = WORKLOADS[workload]
    clips = synth.clip_batch_pcm16(n_clips, seconds, seed=synth.SEED_C4, distinct=min(n_clips, 8))
    jobs = [(clips[i], nr, norm, n_mels) for i in range(n_clips)]
    if cores <= 1:
        _cpu_one((clips[0][: SR * 2], nr, norm, n_mels))  # warm-up (imports, FFT plans)
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_one(j)
        dt = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_one, [(clips[0][: SR * 2], nr, norm, n_mels)] * cores)  # warm every worker
            t0 = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=1)
            dt = time.perf_counter() - t0
    return {"value": n_clips * seconds / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_clips} x {seconds:g} s clips of the same synthetic workload, numpy/scipy oracle port "
                      f"(noisereduce + faster-whisper FeatureExtractor restated; the reference itself is numpy/scipy), {dt:.2f} s of wall time"}


def run_reference(args) -> None:
    """--impl reference: the reference's CPU path (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, clips_per_gpu, seconds, nr, norm, n_mels = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_clips = max(cores, 8)
    sample_s = 60.0 if seconds >= 60.0 else seconds
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.workload, cores, n_clips, sample_s)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * n_clips * sample_s / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32",
            "data": "synthetic", "config": {"workload": desc, "clips_per_step": n_clips, "seconds_per_clip": sample_s, "host_cores": cores,
                                            "note": "bounded sample of the GPU arm's workload; CPU throughput does not depend on the batch size"},
            "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, local: int):
    """Run this rank (and first-touch its pinned staging buffers) on the NUMA node its GPU hangs off: with several ranks
    per box the host<->device copies of the e2e leg otherwise cross the socket interconnect.  Returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    from open_speech_b200 import _native as N
    from open_speech_b200 import synth
    from open_speech_b200.batch import SttFrontEnd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N.require_gpu()  # no device: RuntimeError here, there is nothing to fall back to
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)  # also at N=1: a process started on the far socket pins its staging buffers there
    N.require_gpu()
    N.check(N.lib().osb_init(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    desc, clips, seconds, nr, norm, n_mels = WORKLOADS[args.workload]
    clips = args.clips or clips
    n = int(seconds * SR)
    fe = SttFrontEnd(n_mels=n_mels, sample_rate=SR, noise_reduce=nr, normalize=norm)
    nf = fe.frames(n)

    # synthetic input of the config's shape, generated on the host, pinned; each rank its own shard (weak scaling)
    host_np = synth.clip_batch_pcm16(clips, seconds, seed=synth.SEED_C4 + 1000 * rank, extra_noise_rms=0.01 if nr else 0.0, distinct=8)
    pcm_host = torch.from_numpy(host_np).pin_memory()
    pcm_dev = torch.empty_like(pcm_host, device="cuda")
    pcm_dev.copy_(pcm_host)
    mel_dev = torch.empty((clips, n_mels, nf), dtype=torch.float32, device="cuda")
    mel_host = torch.empty((clips, n_mels, nf), dtype=torch.float32).pin_memory()
    audio_s_per_step = clips * seconds
    h2d_bytes, d2h_bytes = pcm_host.numel() * 2, mel_host.numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- warm-up (also sizes the stream-ordered scratch pool)
    for _ in range(max(args.warmup, 3)):
        fe(pcm_dev, mel_dev)
    torch.cuda.synchronize()

    # ---- resident-input throughput, per-kernel events on the launching stream, clocks sampled during the region
    sampler = ClockSampler(local)
    launches0 = N.lib().osb_launch_count()
    N.check(N.lib().osb_profile_enable(1))
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: fe(pcm_dev, mel_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    buf = ctypes.create_string_buffer(1 << 16)
    N.check(N.lib().osb_profile_report(buf, len(buf)))
    N.check(N.lib().osb_profile_enable(0))
    kern = json.loads(buf.value.decode())
    launches = int(N.lib().osb_launch_count() - launches0)
    ms_step = ms_total / args.steps
    value = world * audio_s_per_step / (ms_step / 1000.0)

    # ---- end to end: pinned host clips in, host features out, copies inside the timed region
    # (osb_stt_frontend_host: pinned host pointers; H2D / kernels / D2H pipelined over 4 clip groups inside the library)
    del pcm_dev, mel_dev
    torch.cuda.empty_cache()

    def e2e_step():
        N.call("osb_stt_frontend_host", pcm_host.data_ptr(), n, clips, n, SR, int(nr), int(norm), n_mels, mel_host.data_ptr())

    for _ in range(2):
        e2e_step()
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns when the features are in host memory
    ms_e2e = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * audio_s_per_step / (ms_e2e / 1000.0)
    checksum = float(mel_host[0, :, :16].double().sum())  # the D2H result is really read

    # ---- one clip, host bytes in -> host features out (BASELINE configs[0] is literally a single clip): latency, not throughput
    single = None
    if rank == 0:
        lat = []
        for i in range(60):
            t0 = time.perf_counter()
            N.call("osb_stt_frontend_host", pcm_host.data_ptr(), n, 1, n, SR, int(nr), int(norm), n_mels, mel_host.data_ptr())
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[10:])
        single = {"p50_ms": lat[len(lat) // 2], "p99_ms": lat[-1], "x_realtime_p50": seconds / (lat[len(lat) // 2] / 1e3)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA-event time per launch from the timed region)
    peak, peak_src = load_peaks()
    dom = max(kern.items(), key=lambda kv: kv[1]["ms"])
    dom_name, dom_ms = dom[0], dom[1]["ms"] / max(1, dom[1]["launches"])
    stage_bytes = DENOISE_BYTES_PER_AUDIO_S if dom_name.startswith("k_nr_") else ALG_BYTES_PER_AUDIO_S[args.workload]
    dom_alg = stage_bytes * audio_s_per_step  # one launch covers the whole per-GPU batch
    achieved = dom_alg / (dom_ms / 1000.0) / 1e9
    chain_alg = ALG_BYTES_PER_AUDIO_S[args.workload] * audio_s_per_step
    chain_gbs = chain_alg / (ms_step / 1000.0) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(args.workload, {}).get(dom_name)
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "ms_per_launch": dom_ms, "algorithmic_bytes_per_launch": dom_alg,
                "share_of_step": dom[1]["ms"] / ms_total,
                "chain": {"algorithmic_bytes_per_step": chain_alg, "achieved": chain_gbs, "frac": chain_gbs / peak},
                "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}}

    # ---- CPU baseline beside it (bounded sample, single thread = the reference's execution model)
    cpu = None
    if not args.no_cpu:
        cpu = cpu_baseline(args.workload, 1, 24 if seconds >= 60 else 48, seconds)  # several seconds of single-thread CPU work

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "clips_per_gpu": clips, "seconds_per_clip": seconds, "sample_rate": SR, "n_mels": n_mels,
                       "noise_reduce": nr, "normalize": norm, "global_clips": clips * world, "parallelism": f"clip-sharded x{world}, no collective", "numa_node_rank0": numa,
                       "l2": f"inputs larger than L2 ({h2d_bytes / 1e6:.0f} MB of pcm16 per GPU per step vs 126 MB)"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                                      "ms_per_step": ms_e2e, "checksum": checksum},
            "single_clip_latency": single, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------- other BASELINE configs (single GPU)
def _gpu_setup():
    import torch

    from open_speech_b200 import _native as N

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    N.require_gpu()
    N.check(N.lib().osb_init(local))
    return torch, N


def _time_ms(torch, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _profile(N, fn, steps):
    N.check(N.lib().osb_profile_enable(1))
    for _ in range(steps):
        fn()
    buf = ctypes.create_string_buffer(1 << 16)
    N.check(N.lib().osb_profile_report(buf, len(buf)))
    N.check(N.lib().osb_profile_enable(0))
    return {k: v["ms"] / steps for k, v in json.loads(buf.value.decode()).items()}


def run_vad(args):
    """BASELINE configs[1]: Silero-shaped VAD scoring + segmenting; (i) one 1 h stream, (ii) 256 streams x 10 min."""
    torch, N = _gpu_setup()
    from open_speech_b200 import synth
    from open_speech_b200.batch import VadBatch

    peak, src = load_peaks()
    vb = VadBatch()
    one = torch.from_numpy(np.tile(synth.clip_pcm16(600.0, seed=synth.SEED_C2), 6)[None, :]).cuda()      # 1 h = 57.6 M samples
    ms_one = _time_ms(torch, lambda: vb(one), max(1, args.steps // 3), warmup=1)
    streams = args.clips or 256
    many = torch.from_numpy(np.tile(synth.clip_pcm16(600.0, seed=synth.SEED_C2 + 1)[None, :], (streams, 1))).cuda()
    ms_many = _time_ms(torch, lambda: vb(many), max(1, args.steps // 3), warmup=1)
    kern = _profile(N, lambda: vb(many), 1)
    audio_s = streams * 600.0
    v = audio_s / (ms_many / 1e3)
    alg = 32125.0 * audio_s
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": max(1, args.steps // 3), "warmup": 1, "ms_per_step": ms_many,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: Silero-shaped VAD 512-sample scoring + segmenting, {streams} streams x 600 s (seeded random-init weights)",
                       "single_stream_1h": {"ms": ms_one, "x_realtime": 3600.0 / (ms_one / 1e3), "windows": 112500}},
            "roofline": {"bound": "hbm", "achieved": alg / (ms_many / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms_many / 1e3) / 1e9 / peak, "traffic": None, "peak_source": src,
                         "note": "compute + serial-latency bound by arithmetic (SURVEY 8(d)); HBM fraction reported for completeness",
                         "kernels_ms_per_step": dict(sorted(kern.items(), key=lambda kv: -kv[1]))}}
    print(json.dumps(line), flush=True)


def run_realtime(args):
    """BASELINE configs[2]: 1024 concurrent G.711 mu-law 8 kHz streams, 20 ms chunks -> pcm16 -> 16 kHz (+VAD): p50/p99 chunk latency.

    Reference-exact semantics: a 320-sample chunk holds no full 512-sample VAD window, so the reference's VAD returns
    0.0 without running; variant B feeds 40 ms chunks (640 samples = 1 window) so that the VAD runs.
    """
    torch, N = _gpu_setup()
    from open_speech_b200 import synth
    from open_speech_b200.batch import RealtimeTick, VadBatch

    S, ticks = args.clips or 1024, 1000
    data = synth.ulaw_streams(S, 64)                     # [64, S, 160], cycled
    host_in = torch.from_numpy(data).pin_memory()
    host_out = torch.empty((S, 320), dtype=torch.int16).pin_memory()
    dev_in = torch.empty((S, 160), dtype=torch.uint8, device="cuda")
    dev_out = torch.empty((S, 320), dtype=torch.int16, device="cuda")
    tick = RealtimeTick(S)

    def one(i):
        dev_in.copy_(host_in[i % 64], non_blocking=True)
        tick(dev_in, dev_out)
        host_out.copy_(dev_out, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(20):
        one(i)
    lat = []
    for i in range(ticks):
        t0 = time.perf_counter()
        one(i)
        lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat)
    # variant B: 40 ms chunks with the VAD (state carried per stream on the device)
    vb = VadBatch()
    dev_in2 = torch.empty((S, 320), dtype=torch.uint8, device="cuda")
    dev_pcm2 = torch.empty((S, 640), dtype=torch.int16, device="cuda")
    tick2 = RealtimeTick(S, chunk=320)
    state = torch.zeros((S, 2, 128), dtype=torch.float32, device="cuda")
    host_in2 = torch.from_numpy(np.ascontiguousarray(data.reshape(32, 2, S, 160).transpose(0, 2, 1, 3).reshape(32, S, 320))).pin_memory()
    host_prob = torch.empty((S, 1), dtype=torch.float32).pin_memory()

    def two(i):
        nonlocal state
        dev_in2.copy_(host_in2[i % 32], non_blocking=True)
        tick2(dev_in2, dev_pcm2)
        probs, state = vb.score(dev_pcm2, state)
        host_prob.copy_(probs, non_blocking=True)
        torch.cuda.synchronize()

    for i in range(10):
        two(i)
    lat2 = []
    for i in range(300):
        t0 = time.perf_counter()
        two(i)
        lat2.append((time.perf_counter() - t0) * 1e3)
    lat2 = np.array(lat2)
    # CUDA-graph variant: the tick's bytes are copied into a pinned slot (host memcpy included), one graph launch, sync
    from open_speech_b200.batch import RealtimeTickGraph

    rg = RealtimeTickGraph(S)
    for i in range(20):
        rg.host_in.copy_(host_in[i % 64])
        rg.run()
    lat3 = []
    for i in range(ticks):
        t0 = time.perf_counter()
        rg.host_in.copy_(host_in[i % 64])
        rg.run()
        lat3.append((time.perf_counter() - t0) * 1e3)
    lat3 = np.array(lat3)
    graph_ok = bool(torch.equal(rg.host_out, host_out)) if (ticks - 1) % 64 == (ticks - 1) % 64 else True
    ms_dev = _time_ms(torch, lambda: tick(dev_in, dev_out), 200)
    peak, src = load_peaks()
    alg = 800.0 * S
    v = S * 0.020 / (np.median(lat) / 1e3)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": ticks, "warmup": 20, "ms_per_step": float(np.median(lat)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: {S} G.711 mu-law 8 kHz streams, 20 ms chunks -> pcm16 16 kHz (host bytes in -> host bytes out per tick)",
                       "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "max": float(lat.max())},
                       "cuda_graph_latency_ms": {"p50": float(np.percentile(lat3, 50)), "p99": float(np.percentile(lat3, 99)), "matches_stream_path": graph_ok},
                       "variant_B_40ms_with_vad_latency_ms": {"p50": float(np.percentile(lat2, 50)), "p99": float(np.percentile(lat2, 99))},
                       "vad_semantics": "A: reference-exact (320 samples < 512 -> VAD scores nothing, prob 0.0); B: 40 ms chunks, 1 window scored"},
            "roofline": {"bound": "hbm", "achieved": alg / (ms_dev / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms_dev / 1e3) / 1e9 / peak,
                         "traffic": None, "peak_source": src, "ms_per_launch": ms_dev,
                         "note": "a tick is 0.8 MB: launch-latency bound, latency is the headline (SURVEY 8(d))"}}
    print(json.dumps(line), flush=True)


def run_tts(args):
    """BASELINE configs[4]: 4096 Kokoro-shaped 24 kHz utterances: trim + peak normalise + effects chain + int16, and voice blends."""
    torch, N = _gpu_setup()
    from open_speech_b200 import synth
    from open_speech_b200.batch import TtsPost

    B = args.clips or 4096
    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
    utts = synth.tts_batch(B, seed=synth.SEED_C5, distinct=32)
    post = TtsPost(24000, fx)
    flat, offsets, lens = post.pack(utts)
    audio_s = float(lens.sum()) / 24000.0
    d_flat, d_off, d_len = torch.from_numpy(flat).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(lens).cuda()
    out = torch.empty(flat.size, dtype=torch.int16, device="cuda")
    mx = int(lens.max())
    ms = _time_ms(torch, lambda: post(d_flat, d_off, d_len, mx, out), args.steps)
    kern = _profile(N, lambda: post(d_flat, d_off, d_len, mx, out), 2)
    # voice blends: B blends of 2-3 packs out of 3
    packs = torch.from_numpy(np.stack([p.reshape(-1) for p in synth.voice_packs(3)])).cuda()
    n = packs.shape[1]
    idx = torch.tensor([[0, 1, -1], [0, 1, -1], [0, 1, 2]] * (B // 3 + 1), dtype=torch.int32)[:B].contiguous().cuda()
    w = torch.tensor([[2 / 3, 1 / 3, 0], [0.5, 0.5, 0], [0.5, 1 / 3, 1 / 6]] * (B // 3 + 1), dtype=torch.float32)[:B].contiguous().cuda()
    bout = torch.empty((B, n), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms_blend = _time_ms(torch, lambda: N.call("osb_voice_blend_dev", packs.data_ptr(), n, idx.data_ptr(), w.data_ptr(), 3, B, bout.data_ptr(), st), args.steps)
    peak, src = load_peaks()
    alg = 192000.0 * audio_s
    blend_bytes = float((idx >= 0).sum().item() + B) * n * 4
    line = {"metric": METRIC, "value": audio_s / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: {B} Kokoro-shaped 24 kHz utterances ({audio_s:.0f} audio-s): trim + peak normalise + "
                                   "[normalize, reverb medium, podcast_eq, robot] + int16; plus voice-style blends",
                       "voice_blend": {"blends_per_s": B / (ms_blend / 1e3), "GBps": blend_bytes / (ms_blend / 1e3) / 1e9, "frac_of_hbm": blend_bytes / (ms_blend / 1e3) / 1e9 / peak}},
            "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms / 1e3) / 1e9 / peak,
                         "traffic": None, "peak_source": src, "kernels_ms_per_step": dict(sorted(kern.items(), key=lambda kv: -kv[1]))}}
    print(json.dumps(line), flush=True)


EXTRA = {"vad": run_vad, "realtime": run_realtime, "tts": run_tts}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="stt_batch", choices=sorted(WORKLOADS) + ["vad", "realtime", "tts"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: the config's 256)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
    if args.workload in EXTRA:
        EXTRA[args.workload](args)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
