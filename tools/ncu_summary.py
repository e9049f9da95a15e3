"""Summarise an .ncu-rep (ncu --set full) into a markdown table + JSON totals.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.md> [--json out.json]"""
import csv
import json
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct")]


def to_bytes(v, unit):
    f = float(v)
    u = unit.lower()
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, unit):
    f = float(v)
    return f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit.lower(), 1)


lines = ["| # | kernel | grid | regs | duration (us) | tensor pipe % | SM % | DRAM % | DRAM read (MB) | DRAM write (MB) | L2 hit % | warps active % |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
recs = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    g = lambda k: r[idx[k]] if k in idx else "nan"   # noqa: E731
    name = r[idx["Kernel Name"]].replace("void <unnamed>::", "").split("(")[0]
    dur = to_us(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
    rd = to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
    rec = {"i": n, "kernel": name, "grid": g("launch__grid_size"), "regs": g("launch__registers_per_thread"),
           "dur_us": dur, "tensor_pct": float(g(COLS[3][0])), "sm_pct": float(g(COLS[4][0])),
           "dram_pct": float(g(COLS[5][0])), "dram_rd": rd, "dram_wr": wr, "l2_hit": float(g(COLS[8][0])),
           "warps": float(g(COLS[9][0]))}
    recs.append(rec)
    lines.append(f"| {n} | `{name}` | {rec['grid']} | {rec['regs']} | {dur:.1f} | {rec['tensor_pct']:.1f} | "
                 f"{rec['sm_pct']:.1f} | {rec['dram_pct']:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {rec['l2_hit']:.1f} | "
                 f"{rec['warps']:.1f} |")
open(out_md, "w").write("\n".join(lines) + "\n")
if out_json:
    json.dump(recs, open(out_json, "w"), indent=1)
print("\n".join(lines[:40]))
