"""Development aid: per-source-line warp-stall reasons of one kernel in an .ncu-rep captured with --import-source on.
   python tools/ncu_source_stalls.py report.ncu-rep <launch index> [top]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--launch-skip",
                      str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
recs = []
for r in rows:
    if r and r[0] in ("Address", "Line No", "#"):
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        recs.append(dict(zip(hdr, r)))
if not recs:
    print("no rows; header candidates:", [r[:3] for r in rows[:8]])
    sys.exit(0)
stall_cols = [c for c in hdr if c.startswith("stall_") or "Stall" in c]
print("columns:", [c for c in hdr][:40])
def num(x):
    try:
        return float(x)
    except Exception:
        return 0.0
key = "Warp Stall Sampling (All Samples)" if "Warp Stall Sampling (All Samples)" in hdr else stall_cols[0]
recs.sort(key=lambda d: -num(d.get(key, 0)))
tot = {}
for d in recs:
    for c in stall_cols:
        tot[c] = tot.get(c, 0) + num(d.get(c, 0))
print("totals:", {k: int(v) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]})
for d in recs[:top]:
    reasons = sorted(((num(d.get(c, 0)), c) for c in stall_cols if c != key), reverse=True)[:3]
    print(int(num(d.get(key, 0))), (d.get("Source") or d.get("Instruction") or "")[:70], [(int(a), b.replace("stall_", "")) for a, b in reasons if a > 0])
