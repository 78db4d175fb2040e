"""Host-side checks of device arithmetic shortcuts that must keep the reference's bits (compiled with gcc, no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_div6_shortcut_is_the_ieee_quotient(tmp_path):
    """x / 6 of the cubic weights as mul + 2 fma (common.cuh: div6_exact) equals the hardware division on structured numerators
    (powers of two and neighbours, multiples of 3, the weights' own arguments on a 2^-20 lattice) and 5e7 random ones."""
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "div6_check")
    subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "native", "div6_check.c"), "-lm"])
    out = subprocess.run([exe, "50"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and " 0 mismatches" in out.stdout, out.stdout + out.stderr
    # the device function is the one the program checks
    src = open(os.path.join(ROOT, "grid.jl_b200", "csrc", "common.cuh")).read()
    body = src[src.index("double div6_exact(double x)"):]
    body = body[:body.index("}")]
    for line in ("const double z = 1.0 / 6.0;", "const double q = x * z;", "const double r = fma(-6.0, q, x);", "return fma(r, z, q);"):
        assert line in body
