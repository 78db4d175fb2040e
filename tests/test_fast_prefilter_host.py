"""The ARITHMETIC of the GB_FAST prefilter (grid.jl_b200/csrc/prefilter_fast.cuh: overlapped segments with the
stationary recurrences + dense end blocks), compiled for the CPU by tests/native/fast_prefilter_check.cu and compared
with the oracle's exact LU / Woodbury solve and with a long-double dense solve of the reference's matrix
(src/interp.jl:955-1014).  No GPU needed; the kernels that run the same functions are tested in test_gpu_round2.py."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMS = {"nan": 0, "reflect": 1, "periodic": 2, "nearest": 3}
SYSTEMS = [("nan", "quadratic"), ("reflect", "quadratic"), ("periodic", "quadratic"), ("nearest", "quadratic"), ("nan", "cubic")]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("native") / "fast_prefilter_check")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "native", "fast_prefilter_check.cu")], stderr=subprocess.DEVNULL)
    return exe


def dense_matrix(bc, order, n, dt):
    """The reference's system matrix, entries rounded to the element type as _interp_invert_matrix builds them."""
    T = np.dtype(dt).type
    M = np.zeros((n, n), dtype=np.longdouble)
    if order == "quadratic":
        off, dia = T(1.0 / 8), T(3.0 / 4)
        for i in range(n):
            M[i, i] = dia
            if i:
                M[i, i - 1] = off
            if i < n - 1:
                M[i, i + 1] = off
        if bc == "nan":
            M[0, :3] = [T(9.0 / 8), T(-1.0 / 4), T(1.0 / 8)]
            M[n - 1, n - 3:] = [T(1.0 / 8), T(-1.0 / 4), T(9.0 / 8)]
        elif bc == "reflect":
            M[0, 0] = M[n - 1, n - 1] = T(float(dia) + 1.0 / 8)
        elif bc == "periodic":
            M[0, n - 1] = M[n - 1, 0] = off
        else:
            M[0, 1] = M[n - 1, n - 2] = T(float(off) + 1.0 / 8)
    else:
        off, dia = T(1.0 / 6), T(4.0 / 6)
        for i in range(2, n - 2):
            M[i, i - 1:i + 2] = [off, dia, off]
        M[0, :3] = [1, -2, 1]
        M[1, 1] = 1
        M[n - 2, n - 2] = 1
        M[n - 1, n - 3:] = [1, -2, 1]
    return M


def solve_longdouble(M, f):
    A = np.concatenate([M, f.astype(np.longdouble)[:, None]], axis=1)
    n = len(f)
    for c in range(n):
        p = c + int(np.argmax(np.abs(A[c:, c])))
        if p != c:
            A[[c, p]] = A[[p, c]]
        A[c] /= A[c, c]
        rows = np.nonzero(A[:, c])[0]
        rows = rows[rows != c]
        A[rows] -= A[rows, c][:, None] * A[c][None, :]
    return A[:, n]


@pytest.mark.parametrize("dt,name,eps", [(np.float64, "f64", 2.220446049250313e-16), (np.float32, "f32", 1.1920929e-07)])
@pytest.mark.parametrize("bc,order", SYSTEMS)
def test_fast_prefilter_arithmetic(checker, oracle, dt, name, eps, bc, order):
    for n, nseg in ((256, 1), (300, 1), (512, 2), (1000, 3), (8192, 16), (4100, 7)):
        rng = np.random.default_rng(n + nseg)
        f = rng.standard_normal(n).astype(dt)
        ref = oracle.invert(f, bc, order)
        out = subprocess.run([checker, name, str(FAMS[bc]), "2" if order == "quadratic" else "3", str(n), str(nseg)], input=f.tobytes(), capture_output=True)
        assert out.returncode == 0
        x = np.frombuffer(out.stdout, dtype=dt)
        scale = float(np.abs(f).max())
        err = np.abs(x.astype(np.float64) - ref.astype(np.float64)).max()
        if dt == np.float64:
            # against the reference arithmetic (which carries its own ~2 ulp): 6 ulp of max|f|; against the truth: 4 ulp
            assert err <= 6 * eps * scale, (bc, order, n, err / (eps * scale))
            if n <= 512:
                truth = solve_longdouble(dense_matrix(bc, order, n, dt), f)
                terr = float(np.abs(x.astype(np.longdouble) - truth).max())
                rerr = float(np.abs(ref.astype(np.longdouble) - truth).max())
                assert terr <= 4 * eps * scale, (bc, order, n, terr / (eps * scale))
                assert rerr <= 4 * eps * scale  # the oracle (a17: LU + Woodbury restatement) against the same witness
        else:
            assert err <= 1e-5 * float(np.abs(ref).max()), (bc, order, n)  # north_star: 1e-5 relative for Float32
            assert err <= 4 * eps * scale
