"""CPU: the C-ABI library builds, loads, and exports every symbol include/erfnet_b200.h declares (no compute)."""
import ctypes
import os
import re

from erfnet_pytorch_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "erfnet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(erf_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/erfnet_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared, "ctypes binding table and header disagree"
    lib.erf_abi_version.restype = ctypes.c_int
    assert lib.erf_abi_version() == 1


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.lib()
    rc = L.erf_confusion_hist(None, None, 8, 10, 20, None, None)
    assert rc == -1 and b"null pointer" in L.erf_last_error_string()
    g = _lib.ConvGeom(1, 8, 8, 3, 3, 8, 8, 16, 16, 0, 3, 1, 1, 1, 0, 1, 1, 0, 16)
    e = _lib.ConvEpilogue()
    rc = L.erf_conv_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.byref(g), ctypes.byref(e), 1, None)
    assert rc == -1 and b"multiples of 4" in L.erf_last_error_string()


def test_every_pdl_launched_kernel_waits_for_its_producer():
    """Static audit of the programmatic-dependent-launch rule (csrc/common.cuh): a kernel launched through launch_pdl() may start
    before the previous kernel of the stream has finished, so its body must execute pdl_wait() (directly or via pdl_enter())."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "erfnet_pytorch_b200", "csrc")
    src = "".join(open(f).read() for f in sorted(glob.glob(os.path.join(root, "*.cu*"))))
    names = sorted(set(re.findall(r"launch_pdl\(([A-Za-z_][A-Za-z_0-9]*)", src)) - {"void"})
    assert len(names) >= 20
    for n in names:
        m = re.search(r"__global__[^;{]*?\b" + n + r"\s*\(", src, re.S)
        assert m, f"definition of {n} not found"
        i = src.index("{", m.end())
        j = src.find("__global__", i)
        body = src[i:j if j > 0 else len(src)]
        assert "pdl_enter();" in body[:400] or "pdl_wait();" in body, f"{n} is launched with PDL but never waits for its producer grid"
