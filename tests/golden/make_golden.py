"""Generates the committed golden vectors from the CPU oracle (run once, here, on CPU):

    python tests/golden/make_golden.py

The reference itself (Julia 0.4/0.5) cannot run in this image, so the vectors come from the oracle,
which tests/test_oracle_kats.py pins against the reference's own known-answer tests.  The GPU tests
compare the CUDA path with these files without needing the oracle at run time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

V, G, H = 1, 2, 4
CASES = [  # name, bc, order, shape, dtype, mask, fill  (miniatures of BASELINE configs C1..C4)
    ("c1_quadratic_reflect_2d", "reflect", "quadratic", (24, 20), np.float64, V, 0.0),
    ("c2_cubic_nan_3d", "nan", "cubic", (10, 9, 11), np.float64, V | G, 0.0),
    ("c3_quadratic_reflect_2d_f32_hess", "reflect", "quadratic", (32, 30), np.float32, V | G | H, 0.0),
    ("c4_linear_periodic_4d", "periodic", "linear", (6, 6, 6, 6), np.float64, V, 0.0),
    ("c4_linear_nearest_4d", "nearest", "linear", (6, 6, 6, 6), np.float64, V, 0.0),
    ("c4_linear_fill_4d", "fill", "linear", (6, 6, 6, 6), np.float64, V, -1.0),
    ("quadratic_nan_3d_hess", "nan", "quadratic", (8, 9, 7), np.float64, V | G | H, 0.0),
    ("quadratic_periodic_1d", "periodic", "quadratic", (16,), np.float64, V | G | H, 0.0),
    ("nearest_reflect_2d", "reflect", "nearest", (7, 5), np.float32, V | G, 0.0),
]


def main():
    for k, (name, bc, order, shape, dt, mask, fill) in enumerate(CASES):
        n = int(np.prod(shape))
        A = O.synth(dt, n, seed=100 + k).reshape(shape, order="F")
        coefs = O.make_coefs(A, bc, order, fill if bc == "fill" else None)
        N = len(shape)
        nq = 256
        lo = np.array([1.0 if bc == "nan" else -1.5 * s for s in shape])
        hi = np.array([float(s) if bc == "nan" else 2.5 * s for s in shape])
        u = O.synth(np.float64, nq * N, seed=500 + k).reshape(nq, N)
        X = lo + (hi - lo) * u
        X[:8] = np.round(X[:8])          # integers
        X[8:16] = np.round(X[8:16]) + 0.5  # half-integers (ties-to-even taps)
        X = np.ascontiguousarray(X.astype(dt))
        val, grad, hess, _ = O.evaluate(coefs, bc, order, X, mask, raise_invalid=False)
        out = dict(A=A, coefs=coefs, X=X, val=val, bc=bc, order=order, mask=mask, fill=fill)
        if grad is not None:
            out["grad"] = grad
        if hess is not None:
            out["hess"] = hess
        np.savez_compressed(os.path.join(HERE, f"eval_{name}.npz"), **out)
    for k, shape in enumerate([(17, 18, 19), (12, 9), (21,)]):
        A = O.synth(np.float64, int(np.prod(shape)), seed=900 + k).reshape(shape, order="F")
        R = O.restrict_flags(A, [True] * len(shape))
        P = O.prolong_sizes(R, list(shape))
        np.savez_compressed(os.path.join(HERE, f"mg_{len(shape)}d.npz"), A=A, R=R, P=P)
    # ---- widening rows: edge-corrected multigrid variants, multi-valued grids, InterpIrregular
    for k, shape in enumerate([(9, 10), (8, 7, 6)]):
        for dt in (np.float64, np.float32):
            A = O.synth(dt, int(np.prod(shape)), seed=950 + k).reshape(shape, order="F")
            out = dict(A=A)
            for d in range(len(shape)):
                out[f"restrictb_{d}"] = O.restrictb(A, d, 1.0)
                out[f"extrap_{d}"] = O.restrict_extrap(A, d, 1.0)
                out[f"prolongb_{d}"] = O.prolongb(out[f"restrictb_{d}"], d, shape[d])
            np.savez_compressed(os.path.join(HERE, f"mgb_{len(shape)}d_{np.dtype(dt).name}.npz"), **out)
    shape, nchan = (12, 10), 3
    A = O.synth(np.float64, int(np.prod(shape)) * nchan, seed=970).reshape(shape + (nchan,), order="F")
    X = 1.0 + O.synth(np.float64, 200 * 2, seed=971).reshape(200, 2) * (np.array(shape) - 1.0)
    vals, grads, coefs = [], [], []
    for c in range(nchan):
        co = O.make_coefs(np.ascontiguousarray(A[..., c]), "reflect", "quadratic")
        v, g, _, _ = O.evaluate(co, "reflect", "quadratic", X, V | G)
        coefs.append(co); vals.append(v); grads.append(g)
    np.savez_compressed(os.path.join(HERE, "channels_quadratic_reflect_2d.npz"), A=A, X=X, coefs=np.stack(coefs, axis=-1), val=np.stack(vals, axis=1),
                        grad=np.stack(grads, axis=1))
    knots = np.sort(O.synth(np.float64, 30, seed=980, lo=-2.0, hi=5.0))
    samples = O.synth(np.float64, 30, seed=981)
    xq = O.synth(np.float64, 300, seed=982, lo=-3.0, hi=6.0)
    out = dict(knots=knots, samples=samples, x=xq)
    for it in ("forward", "backward", "nearest", "linear"):
        out[f"nan_{it}"] = O.irregular(knots, samples, "nan", it, xq)[0]
        out[f"nearest_{it}"] = O.irregular(knots, samples, "nearest", it, xq)[0]
        out[f"fill_{it}"] = O.irregular(knots, samples, "fill", it, xq, fill=-7.5)[0]
    np.savez_compressed(os.path.join(HERE, "irregular_1d.npz"), **out)


if __name__ == "__main__":
    main()
