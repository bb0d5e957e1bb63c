"""Source-level stall attribution: joins the per-instruction warp-stall samples of an `ncu --set full --import-source on`
report (SourceCounters section, `--page source --csv`) with the line table of the shipped library (nvdisasm -g), and prints
the CUDA source lines that hold the samples — which lines the warps are waiting on.
usage: ncu_by_line.py <report.ncu-rep> <kernel-regex> <cubin-stem e.g. bp_decode> [out.txt] [top-N]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kre, stem = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else None
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "kmldpc_b200", "lib", "libkmldpc_b200.so")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
# (the csv lists every instruction twice)
seen, insts = set(), []
for r in rows[2:]:
    if len(r) <= ix["# Samples"] or not r[ix["# Samples"]].isdigit():
        continue
    if r[ix["Address"]] in seen:
        continue
    seen.add(r[ix["Address"]])
    insts.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]] or 0),
                  {k[6:]: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "(" not in k}))
# mangled-name fragment to find the function in the cubin
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith(stem + ".sm_") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
# candidate functions: same instruction count and same opcode sequence as the profiled kernel
funcs, cur, name, line = {}, None, None, 0
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        name, cur = m.group(1), []
        funcs[name] = cur
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m and cur is not None:
        cur.append((m.group(1).strip(), line))
ops = lambda s: re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
want = [ops(s) for s, *_ in insts]
match = [n for n, f in funcs.items() if len(f) == len(want) and [ops(s) for s, _ in f] == want]
if not match:
    near = sorted(((abs(len(f) - len(want)), n, len(f)) for n, f in funcs.items()))[:3]
    sys.exit(f"no function of {cub} matches the {len(want)} instructions of {kname[:80]} (library rebuilt since the capture?); "
             f"closest: {[(n[-60:], l) for _, n, l in near]}")
table = funcs[match[0]]
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = sum(i[1] for i in insts) or 1
tot_e = sum(i[2] for i in insts) or 1
for (s, smp, ex, st), (_, ln) in zip(insts, table):
    a = by_line[ln]
    a[0] += smp
    a[1] += ex
    a[2].update(st)
text = open(os.path.join(root, "kmldpc_b200", "csrc", table[0][1][0])).read().splitlines() if table else []
allst = collections.Counter()
for i in insts:
    allst.update(i[3])
lines = [f"# {kname[:160]}", f"# report {rep}; line table from {cub}; {tot_s} warp-stall samples, {tot_e} executed warp instructions",
         "# all samples by reason: " + " ".join(f"{k}:{100 * v / max(sum(allst.values()), 1):.1f}%" for k, v in allst.most_common(9)),
         "# samples  share  cum   executed  share   file:line  source   [top stall reasons of the line]"]
cum = 0
for ln, (smp, ex, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    cum += smp
    try:
        txt = open(os.path.join(root, "kmldpc_b200", "csrc", ln[0])).read().splitlines()[ln[1] - 1].strip()
    except Exception:
        txt = "?"
    why = " ".join(f"{k}:{v}" for k, v in st.most_common(3) if v)
    lines.append(f"{smp:8d} {100 * smp / tot_s:5.1f}% {100 * cum / tot_s:5.1f}% {ex:10d} {100 * ex / tot_e:5.1f}%  {ln[0]}:{ln[1]:<5d} {txt[:100]}   [{why}]")
res = "\n".join(lines) + "\n"
if out:
    open(out, "w").write(res)
print(res, end="")
