"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list.  usage: launch_summary.py <csv> [out.txt]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg, tot = None, collections.OrderedDict(), 0.0
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
    v = float(d["Metric Value"].replace(",", "")) * {"us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(d["Metric Unit"], 1.0)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
lines = [f"# per-kernel device time (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)",
         f"# source: {sys.argv[1]}"]
for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    lines.append(f"{v / 1e6:10.3f} ms {n:5d} launches {v / n / 1e3:9.1f} us/launch {100 * v / tot:5.1f}%  {k}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text, end="")
