"""Debug: one fused-front launch over 256 x 60 s with the in-kernel clock trace of CTA 0 (OSB_VF_TRACE)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
os.environ["OSB_VF_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/vf_trace.txt"
from open_speech_b200 import _native as N
from open_speech_b200 import synth
from open_speech_b200.vad.silero import VadSession, random_init_weights

N.check(N.lib().osb_init(0))
sess = VadSession(random_init_weights(1002))
batch, secs = 256, 60.0
one = synth.clip_pcm16(secs, seed=500)
pcm = np.tile(one[None, :], (batch, 1))
x = torch.from_numpy(pcm).cuda()
n = pcm.shape[1]
state = torch.zeros((batch, 2, 128), dtype=torch.float32, device="cuda")
probs = torch.empty((batch, n // 512), dtype=torch.float32, device="cuda")
for _ in range(2):
    N.call("osb_vad_score_dev", sess.handle, x.data_ptr(), N.FMT_PCM16, n, batch, n, state.data_ptr(), probs.data_ptr(), n // 512,
           torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(open(os.environ["OSB_VF_TRACE"]).read())
