"""The C-ABI library loads and exports every symbol include/pfem2_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from gpupfem2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pfem2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfem2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} missing: run __graft_entry__.build()")
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"library does not export {missing}"


def test_default_options_and_version():
    L = _lib.load()
    o = _lib.Options()
    L.pfem2_default_options(ctypes.byref(o))
    assert o.struct_size == ctypes.sizeof(_lib.Options)
    assert o.subcell_mode == 0 and o.max_division_level == 4 and o.device == -1 and o.defer_correct == 1
    assert b"sm_100a" in L.pfem2_version()


def test_create_rejects_bad_arguments_without_touching_the_gpu():
    L = _lib.load()
    h = ctypes.c_void_p()
    assert L.pfem2_create(ctypes.byref(h), None, 2, None) == _lib.PFEM2_EINVAL
    view = _lib.MeshView(0, 0, None, None, None, None, None)
    assert L.pfem2_create(ctypes.byref(h), ctypes.byref(view), 2, None) == _lib.PFEM2_EINVAL
    assert b"mesh" in L.pfem2_last_error(None)
    assert L.pfem2_destroy(None) == _lib.PFEM2_OK


def test_default_options_and_removed_variants_are_refused_before_touching_the_gpu():
    """The lazy re-sort is the default; the round-1 A/B kernel variants are gone (their option fields are reserved and must be 0:
    the check precedes every CUDA call)."""
    L = _lib.load()
    o = _lib.Options()
    L.pfem2_default_options(ctypes.byref(o))
    assert o.lazy_sort == 1 and o.defer_correct == 1 and o.stable_order == 0
    assert o.struct_size == ctypes.sizeof(_lib.Options)
    buf = (ctypes.c_double * 8)()
    fake = ctypes.cast(buf, ctypes.c_void_p)  # never dereferenced: the option check fails first
    view = _lib.MeshView(1, 3, fake, fake, fake, fake, fake)
    h = ctypes.c_void_p()
    for name in ("reserved_lane_per_record", "reserved_fuse_project", "reserved_scatter_tma"):
        L.pfem2_default_options(ctypes.byref(o))
        setattr(o, name, 1)
        assert L.pfem2_create(ctypes.byref(h), ctypes.byref(view), 2, ctypes.byref(o)) == _lib.PFEM2_EINVAL, name
        assert b"removed" in L.pfem2_last_error(None)
        assert not h.value


def test_no_cpu_fallback_in_product_path():
    """The product package must not import or reference the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "gpupfem2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pfem2_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_header_is_plain_c99_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/pfem2_b200.h must compile as C99 (no C++ types), and a C program must link against the
    library and call an entry point that needs no GPU."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "cabi.c"
    src.write_text('#include <stdio.h>\n#include "pfem2_b200.h"\n'
                   "int main(void) { pfem2_options o; pfem2_default_options(&o);\n"
                   '  printf("%d %d %s\\n", o.struct_size == (int)sizeof o, pfem2_create(0, 0, 2, 0), pfem2_version()); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", f"-I{inc}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = tmp_path / "cabi"
    r = subprocess.run(["gcc", "-std=c99", f"-I{inc}", str(src), "-o", str(exe), f"-L{libdir}", "-lpfem2_b200", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[:2] == ["1", str(_lib.PFEM2_EINVAL)] and "sm_100a" in r.stdout, (r.stdout, r.stderr)
