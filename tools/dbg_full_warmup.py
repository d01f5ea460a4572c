"""Per-call time of the composed chain in a fresh process (how many calls until the stream-ordered pool has settled)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ctx = bench.Ctx()
torch = ctx.torch
s = bench.stt_setup(ctx, sys.argv[1] if len(sys.argv) > 1 else "stt_full", 0)
ts = []
for i in range(16):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s["run"]()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("per-call ms:", " ".join(f"{t:.1f}" for t in ts))
print("timed(5):", ctx.timed(s["run"], 5, warmup=0))
