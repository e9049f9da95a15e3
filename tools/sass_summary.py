"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
UTCHMMA.2CTA (cta_group::2), LDTM (tcgen05.ld), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), legacy HMMA (must be 0).
usage: python tools/sass_summary.py [libdin_sm100.so] > profiles/sass_tcgen05_rN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200", "libdin_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PATS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM\b"),
        ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"), ("UTCCP", r"\bUTCCP"),
        ("HMMA(legacy)", r"(?<![A-Z])HMMA\b"), ("SYNCS", r"\bSYNCS\b")]
counts, order, cur = collections.OrderedDict(), [], None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for name, pat in PATS:
        if re.search(pat, ln):
            counts[cur][name] += 1
demangled = {}
try:
    names = list(counts)
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(names, out))
except OSError:
    pass
print(f"# {os.path.basename(so)}: SASS mnemonic counts per kernel (cuobjdump -sass), kernels without any listed mnemonic omitted")
print("| kernel | " + " | ".join(n for n, _ in PATS) + " |")
print("|---|" + "---|" * len(PATS))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if not any(c[n] for n, _ in PATS[:8]):
        continue
    name = demangled.get(k, k)
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)          # cu++filt prints template arguments as (int)256
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).split("(")[0].replace("void ", "")
    print(f"| `{name}` | " + " | ".join(str(c[n]) for n, _ in PATS) + " |")
print("| **total** | " + " | ".join(str(tot[n]) for n, _ in PATS) + " |")
print(f"\n{len(counts)} kernels in the library; legacy HMMA instructions: {tot['HMMA(legacy)']}")
