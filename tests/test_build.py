"""CPU-only checks of the build artefacts: the C-ABI library loads without a GPU, exports every symbol that
include/plf_b200.h declares, fails loudly (no CPU fallback) when no device is present, and contains no VIMNMX3."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return g.LIB


def _declared():
    hdr = open(os.path.join(ROOT, "include", "plf_b200.h")).read()
    return sorted(set(re.findall(r"PLF_API\s+[\w\s\*]+?PLF_FN\((\w+)\)", hdr)))


def test_header_symbols_match_binding(plf):
    assert _declared() == sorted(plf.ABI_SYMBOLS)


def test_product_exports_every_declared_symbol(built, plf):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    for s in _declared():
        assert "plf_" + s in exported, s
    lib = plf.Library(built, "plf_")      # loads on a CPU-only box (cudart is linked statically)
    assert lib.fn("last_error") is not None


def test_oracle_exports_every_declared_symbol(oracle):
    out = subprocess.run(["nm", "-D", "--defined-only", oracle.path], capture_output=True, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    for s in _declared():
        assert "plf_cpu_" + s in exported, s


def test_no_cpu_fallback_without_device(built, plf):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = plf.Library(built, "plf_")
    p = lib.default_params()
    ctx = C.c_void_p()
    rc = lib.fn("create")(C.byref(p), 0, C.byref(ctx))
    assert rc == 4, "plf_create must fail with PLF_ERR_NO_DEVICE on a box without a GPU"
    assert b"no CPU path" in lib.fn("last_error")()


def test_sass_has_no_vimnmx3(built):
    """ptxas 12.9 fuses signed min/max chains into VIMNMX3 which returned wrong values on B200 (see orb.cu)."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass
    assert "VIMNMX3" not in sass


def test_headers_are_self_contained_and_assert_their_layouts(tmp_path):
    """A translation unit that includes ONLY the host mirror (C++) or ONLY the C header (C11) compiles; the headers carry
    static assertions on the 28 / 68 / 56 / 72-byte records that cross the boundary."""
    cpp = tmp_path / "only_hpp.cpp"
    cpp.write_text('#include "%s"\nint main() { return 0; }\n' % os.path.join(ROOT, "pli-slam_b200", "host", "plf_frontend.hpp"))
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", str(cpp)])
    c = tmp_path / "only_h.c"
    c.write_text('#include "%s"\nint main(void) { return (int)sizeof(plf_keyline) - 68; }\n' % os.path.join(ROOT, "include", "plf_b200.h"))
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-fsyntax-only", str(c)])
    hdr = open(os.path.join(ROOT, "include", "plf_b200.h")).read()
    for rec, size in (("plf_keypoint", 28), ("plf_keyline", 68), ("plf_proj_query", 56), ("plf_frame_query", 72)):
        assert "static_assert(sizeof(%s) == %d" % (rec, size) in hdr
