"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and the pure host logic (size schedules, ranges, argument validation) behaves like the
reference.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "gridb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(gb):
    lib = gb._lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gridb200.h but not exported"
    assert set(syms) == set(gb._lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.gb_version() == 10100


def test_no_oracle_in_product_path():
    """The product package must not import, link or dlopen the oracle (or any CPU compute fallback)."""
    pkg = os.path.join(ROOT, "grid.jl_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|from\s+\.+oracle|liboracle|grid_oracle|oracle\.py|oracle/)")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(txt), (dirpath, f)
    out = os.popen(f"ldd {os.path.join(pkg, 'libgridb200.so')}").read()
    assert "oracle" not in out


def test_missing_library_fails_loudly(gb, monkeypatch):
    monkeypatch.setattr(gb._lib, "_lib", None)
    monkeypatch.setattr(gb._lib, "LIB_PATH", "/nonexistent/libgridb200.so")
    with pytest.raises(ImportError):
        gb._lib.lib()


def test_restrict_prolong_sizes(gb):
    lib = gb._lib.lib()
    # README.md:236-241
    pyr = gb.restrict_size((1920, 1080, 3), [[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 0, 0]])
    assert pyr.tolist() == [[961, 481, 241, 121], [541, 271, 136, 69], [3, 3, 3, 3]]
    assert gb.prolong_size((1920, 1080), [[1, 1], [1, 1]]).tolist() == [[961, 1920], [541, 1080]]
    n, chain = 1024, []
    while n > 17:
        n = lib.gb_restrict_size(n)
        chain.append(n)
    assert chain == [513, 257, 129, 65, 33, 17]
    for n in range(1, 40):
        assert lib.gb_restrict_size(n) == gb.restrict_size(n)
        for ln in (2 * n - 1, 2 * (n - 1)):
            if ln >= 1:
                out = C.c_int64()
                assert lib.gb_prolong_size(n, ln, C.byref(out)) == 0 and out.value == ln
                assert gb.prolong_size(n, ln) == ln
    assert lib.gb_prolong_size(5, 7, None) == gb._lib.GB_ERR_PROLONG_SIZE
    assert b"Cannot prolong a dimension of length 5 to 7" in lib.gb_last_error()
    with pytest.raises(gb.ErrorException):
        gb.prolong_size(5, 7)
    with pytest.raises(gb.ErrorException):
        gb.restrict_size((4, 4), [True, True, True])


def test_wrap_rules(gb, oracle):
    for bc, name in ((gb.BCreflect, "reflect"), (gb.BCperiodic, "periodic"), (gb.BCnearest, "nearest"), (gb.BCfill, "fill"), (gb.BCnan, "nan")):
        for ln in (1, 2, 5, 8):
            for pos in range(-3 * ln - 2, 3 * ln + 3):
                assert gb.wrap(bc, pos, ln) == oracle.wrap(name, pos, ln)


def test_float_range_lifting(gb):
    r = gb.frange(-1.0, 0.1, 1.0)
    assert (r.start, r.step, r.len, r.divisor) == (-10.0, 1.0, 21, 10.0)
    r = gb.frange(2.0, 0.5, 10.0)
    assert (r.start, r.step, r.len, r.divisor) == (4.0, 1.0, 17, 2.0)
    r = gb.frange(-10, 0.5, 10)
    assert (r.start, r.step, r.len, r.divisor) == (-20.0, 1.0, 41, 2.0)
    r = gb.frange(-3, 0.15, 3.5)
    assert r.len == 44
    assert gb.frange(0.0, 0.25, 1.0).values().tolist() == [0.0, 0.25, 0.5, 0.75, 1.0]
    ur = gb.api._as_range(range(3, 9))
    assert ur.kind == 2 and len(ur) == 6
    sr = gb.api._as_range(range(1, 10, 2))
    assert sr.kind == 1 and len(sr) == 5


def test_argument_validation_without_gpu(gb):
    """Errors raised before any CUDA call, in the reference's vocabulary."""
    lib = gb._lib.lib()
    h = C.c_void_p()
    dims = (C.c_int64 * 2)(4, 4)
    buf = np.zeros(16)
    rc = lib.gb_grid_create(C.byref(h), 1, 2, dims, buf.ctypes.data, 0, gb.BCreflect.code, gb.InterpCubic.code, 0.0, None)
    assert rc == gb._lib.GB_ERR_UNSUPPORTED and b"InterpCubic" in lib.gb_last_error()
    rc = lib.gb_grid_create(C.byref(h), 1, 9, (C.c_int64 * 9)(*([2] * 9)), buf.ctypes.data, 0, 0, 1, 0.0, None)
    assert rc == gb._lib.GB_ERR_UNSUPPORTED
    rc = lib.gb_grid_create(C.byref(h), 7, 2, dims, buf.ctypes.data, 0, 0, 1, 0.0, None)
    assert rc == gb._lib.GB_ERR_ARG
    with pytest.raises(gb.ErrorException):
        gb.InterpGrid(np.zeros((4, 4)), gb.BCfill, gb.InterpLinear)  # src/interp.jl:49-51
    with pytest.raises(gb.ErrorException):
        gb.pad1(np.zeros(3), gb.BCfill)  # src/interp.jl:1304
    assert gb.npoints(gb.InterpCubic) == 4 and gb.npoints(gb.InterpNearest) == 1


def _build_abi_smoke(tmp_path):
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_smoke")
    pkg = os.path.join(ROOT, "grid.jl_b200")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-O1", "-o", exe, os.path.join(ROOT, "tests", "native", "abi_smoke.c"), "-L", pkg,
                           "-l:libgridb200.so", "-Wl,-rpath," + pkg, "-lm"])
    return exe


def test_c_caller_prototypes(gb, tmp_path):
    """tests/native/abi_smoke.c: a plain-C caller with the Julia binding's argument types compiles against the header with
    -Werror, links against the library and runs its host-only part."""
    import subprocess

    exe = _build_abi_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "gridb200 10100" in out.stdout


@pytest.mark.gpu
def test_c_caller_runs_the_readme_example(gb, tmp_path):
    """the same program end to end on the GPU: gb_grid_create -> gb_eval -> gb_grid_destroy, README values bit for bit"""
    import subprocess

    exe = _build_abi_smoke(tmp_path)
    out = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "abi smoke ok" in out.stdout, out.stdout + out.stderr
