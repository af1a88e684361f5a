"""Top SASS instructions of an .ncu-rep (captured with --set full or --section SourceCounters) by warp-stall samples:
python tools/ncu_source_top.py prof.ncu-rep out.csv [n]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if r]
hdr = None
data = []
for r in rows:
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
if not hdr:
    open(out, "w").write(raw[:200000])
    sys.exit("no source table found; raw output saved")
cs = [i for i, c in enumerate(hdr) if "Sampl" in c and "Not" not in c]
key = cs[0]
def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
data.sort(key=lambda r: -num(r[key]))
tot = sum(num(r[key]) for r in data) or 1.0
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["share_of_samples"] + hdr)
    for r in data[:top]:
        w.writerow(["%.4f" % (num(r[key]) / tot)] + r)
print(out, "lines", len(data), "total samples", tot)
