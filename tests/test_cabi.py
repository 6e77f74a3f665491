"""The C-ABI shared library loads on a machine without a GPU and exports exactly the entry points
include/marbles_b200.h declares; the ctypes stub (marbles_b200/_lib.py) lists the same set.  No compute
calls here: without a CUDA device every entry point must fail loudly (there is no CPU path)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "marbles_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbl_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from marbles_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from marbles_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in marbles_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "marbles_b200/_lib.py:SYMBOLS and the header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mbl_[a-z_0-9]+)", out))
    assert exported == set(names), exported ^ set(names)


def test_struct_sizes_match_header(lib):
    """the ctypes mirrors have the C layout (checked against a tiny C program compiled with gcc)"""
    from marbles_b200 import _lib
    src = ('#include <stdio.h>\n#include "marbles_b200.h"\nint main(){printf("%zu %zu %zu\\n",'
           "sizeof(mbl_params),sizeof(mbl_level_geom),sizeof(mbl_layout));return 0;}\n")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.Params), C.sizeof(_lib.LevelGeom), C.sizeof(_lib.Layout)]


def test_no_device_fails_loudly(lib):
    """on the CPU-only build box mbl_create must return an error, never fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from marbles_b200 import _lib
    p = _lib.Params()
    p.nu = p.alpha = 0.1
    p.R, p.gamma, p.mesh_speed = 1.0, 5.0 / 3.0, 1.0
    for d in range(3):
        p.periodic[d] = 1
    ctx = C.c_void_p()
    assert lib.mbl_create(C.byref(p), 0, C.byref(ctx)) != 0
    assert b"no CUDA device" in lib.mbl_last_error()
    with pytest.raises(_lib.MarblesError):
        _lib.check(lib.mbl_create(C.byref(p), 0, C.byref(ctx)))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "marbles_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "marbles_oracle" not in text, f


def _run_c_caller(tmp_path):
    import shutil
    import subprocess
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    libdir = os.path.join(root, "marbles_b200")
    exe = str(tmp_path / "cabi_smoke")
    subprocess.run([gcc, "-O1", "-I", os.path.join(root, "include"), os.path.join(here, "c", "cabi_smoke.c"), "-o", exe,
                    "-L", libdir, "-lmarbles_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    print(res.stdout.strip())
    return res.stdout


def test_plain_c_caller_without_device(tmp_path):
    """tests/c/cabi_smoke.c: a C program that includes include/marbles_b200.h and links the library (no Python,
    no torch in the process).  Without a GPU mbl_create must fail loudly ("no CPU path")."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: see test_plain_c_caller_on_device")
    assert _run_c_caller(tmp_path).startswith("NO_DEVICE")


@pytest.mark.gpu
def test_plain_c_caller_on_device(tmp_path):
    """the same C program on the GPU box: defines a level, initialises a periodic box, steps it 8 times through
    mbl_step and checks mass conservation -- the C ABI driven from plain C on a device"""
    assert _run_c_caller(tmp_path).startswith("DEVICE")


def test_every_entry_point_is_documented_in_integration_md():
    """INTEGRATION.md names, for every function the header declares, the reference interface it stands for"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "marbles_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(mbl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 60
    assert [s for s in syms if s not in doc] == []
