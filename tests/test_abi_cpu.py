"""The C-ABI library loads on a CPU-only box and exports exactly the symbols include/din_sm100.h declares;
the ctypes table mirrors the header; argument validation runs before any launch (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "din_sm100.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"DIN_API\s+[\w\s\*]+?\b(din_\w+)\s*\(", src)))


def _lib():
    from din_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_declares_something():
    names = _declared()
    assert "din_conv2d_nhwc_f16" in names and "din_dynamic_infer_f32" in names and len(names) >= 12
    assert {"din_conv2d_wgrad_nhwc_f16", "din_dynamic_infer_bwd_f32", "din_stem_conv_nhwc_u8", "din_ce_metrics_f32"} <= set(names)


def test_library_exports_every_declared_symbol():
    lib = _lib()
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (din_\w+)", out)))
    assert exported == _declared()


def test_ctypes_table_matches_header():
    lib = _lib()
    assert sorted(lib.PROTOTYPES) == _declared()
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in lib.PROTOTYPES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
    assert C.sizeof(lib.DinConvDesc) == 16 * 4
    # struct layouts: field names and order as in the header
    for struct in (lib.DinConvDesc, lib.DinPackJob, lib.DinFlatJob):
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct.__name__, struct.__name__), src, flags=re.S)
        assert m, struct.__name__
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if decl:
                fields += [n.strip().lstrip("*") for n in re.sub(r"^(const\s+)?\w+\s*\*?", "", decl, count=1).split(",")]
        assert fields == [n for n, _ in struct._fields_], (struct.__name__, fields)
    assert C.sizeof(lib.DinPackJob) == 3 * 8 + 8 * 4
    assert C.sizeof(lib.DinFlatJob) == 3 * 8


def test_loads_and_validates_without_gpu():
    lib = _lib()
    h = lib.load()
    assert h.din_abi_version() == 5
    # invalid arguments are rejected before any CUDA call, with a message
    d = lib.DinConvDesc(n=1, h=8, w=8, c_in=44, x_c_stride=48, c_out=64, y_c_stride=64, kh=3, kw=3, stride=1,
                        pad_h=1, pad_w=1, relu=1, out_f32=0, pool2=0, w_split=1)
    rc = h.din_conv2d_nhwc_f16(C.byref(d), C.c_void_p(16), C.c_void_p(16), None, None, C.c_void_p(16), None)
    assert rc == -1 and b"multiple of 8" in h.din_last_error_string()
    rc = h.din_dynamic_infer_f32(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, 10, 12, 128,
                                 2, 2, 1, 1, None, 1.0, 0, None, None)
    assert rc == -1 and b"unsupported" in h.din_last_error_string()
    with pytest.raises(lib.DinError):
        lib.check(rc, "din_dynamic_infer_f32")


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "din-group-activity-recognition-benchmark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "din_oracle" not in src and "ref_harness" not in src, os.path.join(dirpath, f)
