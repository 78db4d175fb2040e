"""Minimal prefilter drivers for ncu captures: python profiles/run_prefilter_once.py c3 [fast] | c2
c3: interp_invert! of the 8192^2 f32 image (BCreflect, quadratic), exact or GB_FAST; c2: InterpGrid(258^3 f64 cubic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grid_jl_b200 as gb
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
fast = "fast" in sys.argv[2:]
if which == "c3":
    n = 8192
    A = gb.jl_empty((n, n), torch.float32); gb.synth_uniform(A, 3)
    for _ in range(2):
        B = A.clone()
        gb.interp_invert_(B, gb.BCreflect, gb.InterpQuadratic, fast=fast)
else:
    n = 256
    A = gb.jl_empty((n, n, n), torch.float64); gb.synth_uniform(A, 2)
    for _ in range(2):
        G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
torch.cuda.synchronize()
print("ok")
