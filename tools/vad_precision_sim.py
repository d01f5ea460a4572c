"""CPU simulation of the tensor-core operand precisions considered for the fused VAD front kernel.

Products of two bf16 / fp16 values are exact in float32 and tcgen05 accumulates in float32, so a float32 matmul of
the ROUNDED operand planes reproduces a split-precision MMA up to accumulation order.  The script scores seeded
speech-like audio with the oracle network (oracle/vad.py) and with each operand mode, through the LSTM recurrence,
and prints the largest probability error: the number the 1e-3 budget of BASELINE.json's north_star applies to.

    python tools/vad_precision_sim.py [seconds]
    python tools/vad_precision_sim.py recur [seconds]     operand schemes of the tensor-pipe RECURRENCE (k_vad_recur_tc):
                                                          the front stays float32, W_hh . h goes through rounded planes
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from open_speech_b200 import synth  # noqa: E402
from oracle import vad as ovad  # noqa: E402


def rnd(x, dt):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dt).to(torch.float32).numpy()


def planes(x, dt, n):
    out, r = [], np.asarray(x, np.float32)
    for _ in range(n):
        p = rnd(r, dt)
        out.append(p)
        r = (r - p).astype(np.float32)
    return out


def mm(a, b, mode):
    """a [M,K] activations, b [N,K] weights -> a @ b.T under an operand mode (dtype, a planes, b planes, cross terms)."""
    dt, na, nb, terms = mode
    if dt is None:
        return (a @ b.T).astype(np.float32)
    pa, pb = planes(a, dt, na), planes(b, dt, nb)
    acc = np.zeros((a.shape[0], b.shape[0]), np.float32)
    for (i, j) in terms:
        acc += (pa[i] @ pb[j].T).astype(np.float32)
    return acc


def conv(x, w, b, stride, mode):
    B, C, T = x.shape
    xp = np.zeros((B, C, T + 2), np.float32)
    xp[:, :, 1:-1] = x
    t_out = (T + 2 - 3) // stride + 1
    cols = np.stack([xp[:, :, t * stride:t * stride + 3] for t in range(t_out)], axis=1)
    y = mm(cols.reshape(B * t_out, C * 3), w.reshape(w.shape[0], -1), mode) + b
    return y.reshape(B, t_out, -1).transpose(0, 2, 1).astype(np.float32)


def front(net, windows, mode0, mode):
    w = net.w
    x = np.asarray(windows, np.float32)
    frames = np.stack([x[:, 128 * f:128 * f + 256] for f in range(3)], axis=1)
    spec = mm(frames.reshape(-1, 256), w["stft_basis"], mode0).reshape(-1, 3, 258)
    h = np.sqrt(spec[..., :129] ** 2 + spec[..., 129:] ** 2).transpose(0, 2, 1).astype(np.float32)
    for name, _oc, _ic, _k, s in ovad.ENCODER:
        h = np.maximum(conv(h, w[f"{name}.weight"], w[f"{name}.bias"], s, mode), 0.0)
    return (mm(h[:, :, 0], w["lstm.weight_ih"], mode) + w["lstm.bias_ih"] + w["lstm.bias_hh"]).astype(np.float32)


BF, FP = torch.bfloat16, torch.float16
MODES = {
    "f32": (None, 0, 0, ()),
    "bf16 3-term (hi.hi+hi.lo+lo.hi)  [round 1]": (BF, 2, 2, ((0, 0), (0, 1), (1, 0))),
    "fp16 3-term": (FP, 2, 2, ((0, 0), (0, 1), (1, 0))),
    "fp16 act 1 plane, weights 2 planes (2 MMAs)": (FP, 1, 2, ((0, 0), (0, 1))),
    "fp16 act 2 planes, weights 1 plane (2 MMAs)": (FP, 2, 1, ((0, 0), (1, 0))),
    "fp16 single pass (1 MMA)": (FP, 1, 1, ((0, 0),)),
    "bf16 single pass (1 MMA)": (BF, 1, 1, ((0, 0),)),
}


RECUR_MODES = {  # (dtype, terms (W plane, h plane))
    "f32": None,
    "bf16 3-term (W_hi.h_hi + W_hi.h_lo + W_lo.h_hi)": (BF, ((0, 0), (0, 1), (1, 0))),
    "fp16 3-term": (FP, ((0, 0), (0, 1), (1, 0))),
    "fp16 W 1 plane, h 2 planes (2 MMAs)  [shipped]": (FP, ((0, 0), (0, 1))),
    "fp16 W 2 planes, h 1 plane (2 MMAs)": (FP, ((0, 0), (1, 0))),
    "bf16 W 1 plane, h 2 planes (2 MMAs)": (BF, ((0, 0), (0, 1))),
    "fp16 1 plane each (1 MMA)  [OSB_VAD_RECUR=fp16x1]": (FP, ((0, 0),)),
}


def recur_main(secs):
    net = ovad.SileroNet()
    w = net.w
    sig = ovad._sigmoid
    worst = {k: 0.0 for k in RECUR_MODES}
    for seed in (3, 11, 21):
        a = synth.clip_pcm16(secs, seed=seed).astype(np.float32) / 32768.0
        n_win = len(a) // 512
        pre = net.front(a[: n_win * 512].reshape(n_win, 512))
        ref = None
        for name, m in RECUR_MODES.items():
            h, c, p = np.zeros(128, np.float32), np.zeros(128, np.float32), np.zeros(n_win, np.float32)
            wp = planes(w["lstm.weight_hh"], m[0], 2) if m else None
            for t in range(n_win):
                if m is None:
                    g = pre[t] + w["lstm.weight_hh"] @ h
                else:
                    hp, acc = planes(h, m[0], 2), np.zeros(512, np.float32)
                    for (i, j) in m[1]:
                        acc += (wp[i] @ hp[j]).astype(np.float32)
                    g = pre[t] + acc
                c = (sig(g[128:256]) * c + sig(g[:128]) * np.tanh(g[256:384])).astype(np.float32)
                h = (sig(g[384:]) * np.tanh(c)).astype(np.float32)
                p[t] = sig(np.float32(np.dot(np.maximum(h, 0.0), w["dec.weight"]) + w["dec.bias"][0]))
            ref = p if ref is None else ref
            worst[name] = max(worst[name], float(np.abs(p - ref).max()))
    for name in RECUR_MODES:
        print(f"{name:52s} max |dprob| = {worst[name]:.2e}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "recur":
        return recur_main(float(sys.argv[2]) if len(sys.argv) > 2 else 120.0)
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    net = ovad.SileroNet()
    worst = {k: 0.0 for k in MODES}
    flips = {k: 0 for k in MODES}
    for seed in (1002, 1004, 321):
        pcm = synth.clip_pcm16(secs, seed=seed)
        a = pcm.astype(np.float32) / 32768.0
        n_win = len(a) // 512
        wins = a[: n_win * 512].reshape(n_win, 512)
        ref, _ = net.score_stream(a)
        for name, mode in MODES.items():
            # the audio operand of the DFT layer: pcm16 is exact in two bf16 / fp16 planes, so that layer always takes both
            # audio planes when the mode has one activation plane only ("act 1 plane" applies from |STFT| on)
            dt, na, nb, terms = mode
            mode0 = mode if dt is None or na == 2 else (dt, 2, nb, tuple({(0, j) for (_, j) in terms} | {(1, 0)}))
            pre = front(net, wins, mode0, mode)
            h = np.zeros(128, np.float32)
            c = np.zeros(128, np.float32)
            p = np.zeros(n_win, np.float32)
            for t in range(n_win):
                p[t], h, c = net.step(pre[t], h, c)
            worst[name] = max(worst[name], float(np.abs(p - ref).max()))
            flips[name] += int(((p >= 0.5) != (ref >= 0.5)).sum())
    for name in MODES:
        print(f"{name:52s} max |dprob| = {worst[name]:.2e}   threshold flips = {flips[name]}")


if __name__ == "__main__":
    main()
