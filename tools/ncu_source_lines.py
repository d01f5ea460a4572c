"""Instructions executed and stall samples per SOURCE line of one kernel in an ncu report (needs -lineinfo and --import-source on):
    python tools/ncu_source_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out, cur_file, cur_fn, hdr = [], None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit() and (pat in (cur_fn or "")):
        d = dict(zip(hdr, r))
        def num(k):
            v = d.get(k, "0")
            return int(v) if v.strip().lstrip("-").isdigit() else 0
        out.append((cur_file.split("/")[-1], int(r[0]), num("Instructions Executed"), num("# Samples"), r[1][:110]))
tot_i = sum(o[2] for o in out) or 1
tot_s = sum(o[3] for o in out) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
for f, ln, ins, smp, src in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"{f}:{ln:5d}  inst {100 * ins / tot_i:5.1f} %  stall samples {100 * smp / tot_s:5.1f} %  | {src.strip()}")
