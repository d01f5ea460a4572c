"""Per-call time of the composed chain in a fresh process: synchronised calls, then back-to-back calls timed by events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ctx = bench.Ctx()
torch = ctx.torch
s = bench.stt_setup(ctx, sys.argv[1] if len(sys.argv) > 1 else "stt_full", 0)
ts = []
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s["run"]()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("synchronised per-call ms:", " ".join(f"{t:.1f}" for t in ts))
for rep in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    host = []
    evs[0].record()
    for i in range(12):
        t0 = time.perf_counter()
        s["run"]()
        host.append((time.perf_counter() - t0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("back-to-back device ms:", " ".join(f"{evs[i].elapsed_time(evs[i + 1]):.1f}" for i in range(12)))
    print("back-to-back host   ms:", " ".join(f"{t:.1f}" for t in host))
