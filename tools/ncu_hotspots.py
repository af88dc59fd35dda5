"""Opcode mix and hottest instructions of one kernel from an .ncu-rep taken with --import-source on
(reads `ncu -i REP --page source --csv`)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][0], rows[0][1][:70])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 5]
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
totinst = sum(f(r, "Instructions Executed") for r in data)
print("warp-stall samples %d, warp instructions executed %d" % (tot, totinst))
agg = collections.defaultdict(lambda: [0, 0])
for r in data:
    src = r[ix["Source"]].strip()
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    agg[op][0] += f(r, "# Samples"); agg[op][1] += f(r, "Instructions Executed")
print("opcode       samples  instructions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-12s %6.1f%%  %6.1f%%" % (k, 100 * v[0] / tot, 100 * v[1] / totinst))
print("hottest instructions (share of samples, top stall reasons)")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:12]:
    st = {k[6:]: f(r, k) for k in hdr if k.startswith("stall_") and "Not" not in k}
    top = ", ".join("%s=%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print("%5.1f%%  %-58s %s" % (100 * f(r, "# Samples") / tot, r[ix["Source"]][:58], top))
