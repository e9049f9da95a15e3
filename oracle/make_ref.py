"""make_ref.py — stage the reference's own Python sources under oracle/_ref/reference/.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it for use as a checker is a file copy.  /root/reference exists only in
the authoring container; the GPU box receives a snapshot of this repository (git-ignored files included), so a staged
copy under oracle/_ref/ (git-ignored: it never enters history; NOT gpurun-ignored: it travels like a built .so) is how
  * tests/test_dropin_gpu.py runs the reference's UNMODIFIED trainer (train_net_dynamic.train_net) against this
    package's drop-in modules on a real B200, and
  * bench.py --impl reference / cpu_baseline can time the reference's own classes (kind "reference") instead of the
    oracle port.
Nothing under the package imports from here, and nothing here is ever edited: files are copied byte for byte, the
documented patches (oracle/ref_harness.py) are applied in memory at import time.

    python oracle/make_ref.py            # no-op when /root/reference is absent
"""
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def stage(src=SRC, dst=DST):
    if not os.path.isfile(os.path.join(src, "infer_model.py")):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d != "__pycache__"]
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), src)
                os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel))
                n += 1
    with open(os.path.join(dst, "STAGED_FROM"), "w") as fh:
        fh.write(f"{src} ({n} .py files, byte-for-byte; staged by oracle/make_ref.py)\n")
    return n


if __name__ == "__main__":
    n = stage()
    print(f"staged {n} files into {DST}" if n else f"{SRC} not present: nothing staged", file=sys.stderr)
