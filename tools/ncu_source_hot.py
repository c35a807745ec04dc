"""Development aid: hottest CUDA source lines (warp-stall samples) of one kernel in an .ncu-rep captured with
--import-source on.   python tools/ncu_source_hot.py report.ncu-rep <launch index> [top]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip",
                      str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, lines, total = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        print(r[1])
    elif r[0] == "Line No":
        hdr = r
        ci, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        names = hdr
    elif r[0].isdigit() and hdr:
        # source text may contain unescaped quotes: index the numeric columns from the END of the row
        back = len(hdr)
        try:
            s, inst = int(r[ci - back] or 0), int(r[ii - back] or 0)
        except ValueError:
            continue
        total += s
        lines.append((s, cur, int(r[0]), inst, ",".join(r[1:len(r) - back + 2]).strip()[:110], None))
print("total samples", total)
for s, f, ln, inst, src, reasons in sorted(lines, reverse=True)[:top]:
    print("%6d %5.1f%% %9d  %s:%d  %s" % (s, 100.0 * s / max(total, 1), inst, f, ln, src))
