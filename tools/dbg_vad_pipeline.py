"""A/B of the pipelined VAD chunks (front of chunk i + 1 beside the recurrence of chunk i) on 256 streams x 1 h.  GPU box only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_speech_b200 import _native as N, synth  # noqa: E402
from open_speech_b200.batch import VadBatch  # noqa: E402
from open_speech_b200.vad.silero import VadSession  # noqa: E402

N.require_gpu()
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
vb = VadBatch(VadSession())
base = torch.from_numpy(synth.clip_pcm16(600.0, seed=77)).cuda()
many = base.repeat(reps).unsqueeze(0).repeat(streams, 1).contiguous()
audio_s = streams * many.shape[1] / 16000


def timed(n=3):
    vb(many)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        out = vb(many)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out


ref = None
for rec, pipe, waves, width in ((0, "0", "32", "0"), (1, "0", "32", "0"), (1, "1", "32", "0"), (2, "0", "32", "0"), (2, "1", "32", "0"), (2, "1", "16", "0")):
    N.call("osb_vad_set_recurrence", vb.session.handle, rec)
    os.environ["OSB_VAD_PIPELINE"] = pipe
    os.environ["OSB_VAD_CHUNK_WAVES"] = waves
    os.environ["OSB_VAD_PIPE_WIDTH"] = width
    ms, out = timed()
    probs = out[0] if isinstance(out, (tuple, list)) else out
    p = probs.float().cpu() if hasattr(probs, "cpu") else None
    same = None
    if p is not None:
        if ref is None:
            ref = p
        else:
            same = float((ref - p).abs().max())
    print(f"recur={rec} pipeline={pipe} waves={waves} width={width}: {ms:.2f} ms  {audio_s / ms * 1e3 / 1e6:.3f} M audio-s/s  max_dp_vs_first={same}", flush=True)
