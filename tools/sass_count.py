"""Static SASS instruction mix per kernel: python tools/sass_count.py [substr ...] (kernels whose demangled name contains a substring)."""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "open_speech_b200", "libosb200.so")
out = subprocess.run(f"cuobjdump -sass {lib} | c++filt", shell=True, capture_output=True, text=True).stdout
want = sys.argv[1:] or ["k_logmel<false>", "k_nr_stft", "k_nr_istft", "k_ps_stft<float>", "k_ps_istft"]
cur, mix = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (.*)$", line)
    if m:
        cur = m.group(1).strip()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        mix[cur][m.group(1)] += 1
for k, c in mix.items():
    if not any(w in k for w in want):
        continue
    tot = sum(v for n, v in c.items() if n != "NOP")
    fp = sum(v for n, v in c.items() if n in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2"))
    print(f"{k[:70]:70s} total {tot:6d}  fp {fp:5d}  " + " ".join(f"{n}:{v}" for n, v in c.most_common(14) if n != "NOP"))
