"""Short, deterministic targets for ncu (one workload, a few launches; numbers taken under ncu are never bench values).

    python tools/prof_target.py stt|full|vad|tts|rt [reps]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from open_speech_b200 import _native as N
from open_speech_b200 import synth

N.require_gpu()
N.check(N.lib().osb_init(0))
what = sys.argv[1] if len(sys.argv) > 1 else "stt"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if what == "stt":
    from open_speech_b200.batch import SttFrontEnd

    pcm = torch.from_numpy(synth.clip_batch_pcm16(256, 60.0, seed=synth.SEED_C4, extra_noise_rms=0.01, distinct=8)).cuda()
    fe = SttFrontEnd(n_mels=128, noise_reduce=True, normalize=True)
    for _ in range(reps):
        out = fe(pcm)
elif what == "full":
    sys.argv = [sys.argv[0]]
    import bench
    from open_speech_b200.batch import SttFull
    from open_speech_b200.vad.silero import VadSession

    wire = torch.from_numpy(bench.ulaw_clips(256, 60.0, 0)).cuda()
    op = SttFull(VadSession(), fmt="g711_ulaw", from_rate=8000, linear_chunk=160)
    for _ in range(reps):
        out = op(wire)
elif what == "vad":
    from open_speech_b200.batch import VadBatch

    x = torch.from_numpy(np.tile(synth.clip_pcm16(120.0, seed=synth.SEED_C2)[None, :], (256, 1))).cuda()
    vb = VadBatch()
    for _ in range(reps):
        out = vb(x)
elif what == "tts":
    from open_speech_b200.batch import TtsPost

    fx = [{"type": "normalize", "target_lufs": -16}, {"type": "reverb", "room": "medium"}, {"type": "podcast_eq"}, {"type": "robot"}]
    post = TtsPost(24000, fx)
    flat, offsets, lens = post.pack(synth.tts_batch(4096, seed=synth.SEED_C5, distinct=32))
    d = [torch.from_numpy(a).cuda() for a in (flat, offsets, lens)]
    for _ in range(reps):
        out = post(d[0], d[1], d[2], int(lens.max()))
torch.cuda.synchronize()
print("ok", what, N.lib().osb_launch_count(), "launches")
