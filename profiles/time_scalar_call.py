"""Latency of the scalar API (one query per C-ABI call, host buffers): G[x], valgrad(G, x) and a 100-query batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import grid_jl_b200 as gb
A = np.random.default_rng(0).standard_normal((64, 64, 64))
G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
x = np.array([[10.3, 20.7, 30.1]])
v = np.empty(1); g = np.empty((1, 3))
def t(fn, k=2000):
    for _ in range(50): fn()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    return (time.perf_counter() - t0) / k * 1e6
print("G[x,y,z]                        %.1f us" % t(lambda: G[10.3, 20.7, 30.1]))
print("valgrad_(v, g, G, x) nq=1       %.1f us" % t(lambda: gb.valgrad_(v, g, G, x)))
X = 5 + np.random.default_rng(1).random((100, 3)) * 50; vv = np.empty(100); gg = np.empty((100, 3))
print("valgrad_ nq=100                 %.1f us" % t(lambda: gb.valgrad_(vv, gg, G, X)))
X = 5 + np.random.default_rng(1).random((10000, 3)) * 50; vv = np.empty(10000); gg = np.empty((10000, 3))
print("valgrad_ nq=10000               %.1f us" % t(lambda: gb.valgrad_(vv, gg, G, X), 500))
