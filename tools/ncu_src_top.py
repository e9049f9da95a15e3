"""Summarise an ncu report's source page: top stall locations for launch index N.
usage: python tools/ncu_src_top.py <report.ncu-rep> <launch_skip> [top]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:90])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[idx['# Samples']].isdigit():
        if data:
            break
        continue
    data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
    s = int(r[idx["# Samples"]])
    st = sorted(((h, int(r[idx[h]])) for h in stalls if int(r[idx[h]]) > 0), key=lambda kv: -kv[1])[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}%  {r[idx['Source']].strip()[:64]:64s} {st} exec={r[idx['Instructions Executed']]}")
