"""Warp-stall samples per CUDA source line of one kernel (reads `ncu -i REP --page source --csv --print-source cuda,sass`;
the report must have been taken with --import-source on and the binary built with -lineinfo)."""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 8: continue
    if r[0] == "": continue          # SASS rows under the line
    try:
        samples = float(r[6].replace(",", "")); inst = float(r[7].replace(",", ""))
    except ValueError:
        continue
    lines.append((samples, inst, cur_file, r[0], r[1].strip()))
tot = sum(l[0] for l in lines); toti = sum(l[1] for l in lines)
print("samples %d, instructions %d" % (tot, toti))
for s, i, f, ln, src in sorted(lines, key=lambda l: -l[0])[:topn]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * s / tot, 100 * i / toti, f, ln, src[:100]))
