"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python tools/launch_summary.py csv [out] [--all]
By default only this library's kernels (namespace aihab::) are kept: the list also holds the one-off PyTorch launches
of the benchmark set-up (text tower for the class head, weight casts, synthetic-input generators)."""
import collections, csv, re, sys
keep_all = "--all" in sys.argv
sys.argv = [a for a in sys.argv if a != "--all"]
rows = list(csv.reader(open(sys.argv[1])))
start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    # this library's kernels live in aihab::<unnamed>; depending on the ncu name base the CSV shows "aihab::..." or
    # the tail "unnamed>::..."
    if len(r) <= vi or (not keep_all and "aihab::" not in r[ki] and "unnamed>::" not in r[ki]):
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("aihab::<unnamed>::", "").replace("unnamed>::", "")[:64]
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
lines = ["# gpu__time_duration.sum per kernel (ncu, --clock-control none; cold-cache, serialised: compare SHARES)",
         f"# total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches", "kernel,launches,total_us,avg_us,share"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k},{n},{t:.1f},{t / n:.1f},{t / tot:.3f}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
