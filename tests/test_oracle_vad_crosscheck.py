"""CPU: the numpy restatement of the Silero-v5-shaped network (oracle/vad.py, PARITY UNPINNED against the real ONNX file) against
the same architecture assembled from torch's own layers: F.pad(reflect) + conv1d with the STFT basis as the kernel (stride 128),
nn.Conv1d encoder blocks, nn.LSTMCell, a 1x1 Conv1d head.  Pins the restatement's arithmetic (padding, strides, gate order, head),
not the architecture itself."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import vad as ovad


def _torch_net(w):
    enc = []
    for name, oc, ic, k, s in ovad.ENCODER:
        conv = torch.nn.Conv1d(ic, oc, k, stride=s, padding=1)
        conv.weight.data = torch.from_numpy(w[f"{name}.weight"].copy())
        conv.bias.data = torch.from_numpy(w[f"{name}.bias"].copy())
        enc.append(conv)
    cell = torch.nn.LSTMCell(128, 128)
    cell.weight_ih.data = torch.from_numpy(w["lstm.weight_ih"].copy())
    cell.weight_hh.data = torch.from_numpy(w["lstm.weight_hh"].copy())
    cell.bias_ih.data = torch.from_numpy(w["lstm.bias_ih"].copy())
    cell.bias_hh.data = torch.from_numpy(w["lstm.bias_hh"].copy())
    head = torch.nn.Conv1d(128, 1, 1)
    head.weight.data = torch.from_numpy(w["dec.weight"].copy()).reshape(1, 128, 1)
    head.bias.data = torch.from_numpy(w["dec.bias"].copy())
    basis = torch.from_numpy(w["stft_basis"].copy()).reshape(258, 1, 256)

    @torch.no_grad()
    def score(audio, state=None):
        n_win = len(audio) // 512
        x = torch.from_numpy(np.asarray(audio[: n_win * 512], np.float32)).reshape(n_win, 1, 512)
        x = F.pad(x, (0, 64), mode="reflect")                      # [W,1,576]
        spec = F.conv1d(x, basis, stride=128)                      # [W,258,3]
        h = torch.sqrt(spec[:, :129] ** 2 + spec[:, 129:] ** 2)    # [W,129,3]
        for conv in enc:
            h = torch.relu(conv(h))
        feat = h[:, :, 0]                                          # [W,128]
        hs = torch.zeros(1, 128) if state is None else torch.from_numpy(state[0].copy())
        cs = torch.zeros(1, 128) if state is None else torch.from_numpy(state[1].copy())
        probs = []
        for t in range(n_win):
            hs, cs = cell(feat[t : t + 1], (hs, cs))
            probs.append(torch.sigmoid(head(torch.relu(hs)[:, :, None]))[0, 0, 0])
        return torch.stack(probs).numpy(), np.stack([hs.numpy(), cs.numpy()])

    return score


def test_numpy_net_matches_torch_layers():
    from open_speech_b200 import synth

    w = ovad.make_weights(1002)
    score = _torch_net(w)
    net = ovad.SileroNet(w)
    audio = synth.clip_pcm16(6.0, seed=21).astype(np.float32) / 32768.0
    p_ref, s_ref = net.score_stream(audio)
    p_t, s_t = score(audio)
    assert p_ref.shape == p_t.shape == (len(audio) // 512,)
    assert np.abs(p_ref - p_t).max() <= 2e-5 and np.abs(s_ref - s_t).max() <= 2e-5
    assert (p_ref >= 0.5).any() and (p_ref < 0.5).any()
    # carried state: second half from the first half's state
    half = (len(audio) // 1024) * 512
    p1, s1 = net.score_stream(audio[:half])
    p2, _ = net.score_stream(audio[half:], s1)
    q1, t1 = score(audio[:half])
    q2, _ = score(audio[half:], t1)
    assert np.abs(np.concatenate([p1, p2]) - p_ref).max() <= 1e-6
    assert np.abs(np.concatenate([q1, q2]) - np.concatenate([p1, p2])).max() <= 2e-5
    # the run() contract the reference wrapper drives (src/vad/silero.py:86-95)
    out, st = net.run(None, {"input": audio[None, :512], "state": np.zeros((2, 1, 128), np.float32), "sr": np.array(16000, np.int64)})
    assert out.shape == (1, 1) and st.shape == (2, 1, 128) and abs(float(out[0, 0]) - float(p_t[0])) <= 2e-5
