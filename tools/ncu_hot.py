"""Print the hottest SASS instructions (warp-stall samples) of an ncu report:  python tools/ncu_hot.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
si, ai, ii = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data, tot = [], 0
for n, r in enumerate(rows[start + 1:]):
    if len(r) <= ai or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name":
            break
        continue
    try:
        v = int(r[ai])
    except ValueError:
        continue
    tot += v
    data.append((v, n, r[si].strip(), r[ii]))
print("total samples", tot, "instructions", len(data))
for v, n, s, ie in sorted(data, reverse=True)[:top]:
    print(f"{v:6d} {100*v/tot:5.1f}%  #{n:4d} x{ie:>9s}  {s[:100]}")
