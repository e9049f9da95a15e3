"""Logits error of the degenerate clip shapes (T = 1 / N = 1, tiny maps) over several seeds, for the plan's
small-launch precision options (din_b200/engine.py): DIN_SMALL_EXACT (hi + lo weights for latency-bound launches) and
DIN_SMALL_EMBED_F32 (fp32 crops + fp32 fc_emb_1 for < 64 actor rows).  One process per setting (the knobs are read
at import).  usage: python tests/tools/edge_precision_study.py            (prints a table)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = [("T1_N1", dict(num_frames=1, num_boxes=1), 1), ("T1_N12", dict(num_frames=1, num_boxes=12), 2),
         ("T10_N1", dict(num_frames=10, num_boxes=1), 2), ("T3_N4_96x160", dict(num_frames=3, num_boxes=4), 2)]
SEEDS = list(range(8))


def worker():
    for p in (os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200"), os.path.join(ROOT, "oracle"),
              os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import io
    import contextlib
    import torch
    from test_e2e_gpu import _pc, _run_case
    out = {}
    for name, kw, B in CASES:
        hw = (96, 160) if "96x160" in name else (64, 96)
        errs = []
        for seed in SEEDS:
            with contextlib.redirect_stdout(io.StringIO()):
                o, r = _run_case(torch.device("cuda:0"), _pc("vgg16", hw, **kw), B, seed=seed, tol=1.0)
            errs.append(float((o.cpu() - r).abs().max() / r.abs().max()))
        out[name] = errs
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
        sys.exit(0)
    rows = []
    for exact, embed in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
        env = dict(os.environ, DIN_SMALL_EXACT=exact, DIN_SMALL_EMBED_F32=embed, DIN_OFFLINE="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=env, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        if not line:
            print(r.stderr[-2000:])
            continue
        res = json.loads(line[0][7:])
        for name, errs in res.items():
            s = sorted(errs)
            rows.append((exact, embed, name, s[len(s) // 2], s[-1], sum(e > 1e-3 for e in errs), errs))
    print("| exact weights | fp32 embed | case | median rel err | max rel err | seeds > 1e-3 (of 8) | per seed |")
    print("|---|---|---|---|---|---|---|")
    for exact, embed, name, med, mx, n_bad, errs in rows:
        print(f"| {exact} | {embed} | {name} | {med:.2e} | {mx:.2e} | {n_bad} | " + " ".join(f"{e:.1e}" for e in errs) + " |")
