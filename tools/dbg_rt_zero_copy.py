"""Experiment: the 1024-stream tick with its buffers in pinned host memory (kernels read / write them over PCIe: no explicit copies)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from open_speech_b200 import _native as N, synth
from open_speech_b200.realtime.gate import RealtimeGate
from oracle import codec

N.require_gpu(); N.check(N.lib().osb_init(0))
S, chunk = 1024, 160
data = synth.ulaw_streams(S, 64)
host_in = torch.from_numpy(np.ascontiguousarray(data)).pin_memory()

def run(zero_copy, ticks=3000):
    g = RealtimeGate(S, chunk, fmt="g711_ulaw", session=None, threshold=0.5, silence_duration_ms=500)
    dev_in = torch.empty((S, chunk), dtype=torch.uint8, device="cuda")
    host_pcm = torch.empty((S, g.n_out), dtype=torch.int16).pin_memory()
    if zero_copy:
        g.pcm = host_pcm                                  # kernels write the pcm16 straight into pinned host memory
        g.events = torch.zeros((g.max_events + 1, 3), dtype=torch.int32).pin_memory()
    lat = np.empty(ticks)
    for i in range(ticks + 20):
        t0 = time.perf_counter()
        if zero_copy:
            w = host_in[i % 64]
            N.call("osb_gate_tick_dev", None, w.data_ptr(), g.fmt, g.chunk, g.from_rate, 0, g.S, g.chunk, g.pcm.data_ptr(), g.n_out, g.state.data_ptr(),
                   g.vad_state.data_ptr(), None, 0, None, 0, float(g.threshold), int(g.silence_ms), g.work.data_ptr(), g.events.data_ptr(),
                   g.events[g.max_events].data_ptr(), g.max_events, torch.cuda.current_stream().cuda_stream)
            torch.cuda.current_stream().synchronize()
            k = int(g.events[g.max_events, 0])
        else:
            dev_in.copy_(host_in[i % 64], non_blocking=True)
            g.tick(dev_in)
            host_pcm.copy_(g.pcm, non_blocking=True)
            k = len(g.read_events())
        if i >= 20: lat[i - 20] = (time.perf_counter() - t0) * 1e6
    ref = codec.decode_audio_to_pcm16(host_in[(ticks + 19) % 64][5].numpy().tobytes(), "g711_ulaw", 16000)
    ok = host_pcm[5].numpy().tobytes() == ref
    print("zero_copy" if zero_copy else "copies   ", f"p50 {np.percentile(lat,50):.1f} us  p99 {np.percentile(lat,99):.1f} us  bit-exact row: {ok}")

run(False); run(True); run(False); run(True)
