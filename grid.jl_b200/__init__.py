"""grid.jl_b200 -- B200-native engine for Grid.jl's batched evaluation hot path.

The directory name contains a dot, so import it through the top-level shim:

    import grid_jl_b200 as gb

Everything numerical runs in libgridb200.so (hand-written sm_100a CUDA behind a C ABI,
include/gridb200.h).  This package is the host-side mirror of the reference's Julia API.
"""
from . import _lib
from ._lib import (BoundsError, CudaError, DimensionMismatch, ErrorException, GridB200Error, MethodError, LIB_PATH)
from .api import *  # noqa: F401,F403
from .api import (AbstractInterpGrid, BCfill, BCna, BCnan, BCnearest, BCnil, BCperiodic, BCreflect, BoundaryCondition,
                  CoordInterpGrid, FloatRange, InterpBackward, InterpCubic, InterpForward, InterpGrid, InterpGridChannels, InterpIrregular, InterpLinear, InterpNearest, InterpQuadratic,
                  InterpType, StepRange, UnitRange, filledges_, frange, getindex_, getindex_batch, interp_invert_, jl_empty,
                  jl_from_numpy, jl_to_numpy, launch_count, npoints, PROFILE_STAGES, profile_enable, profile_read, pad1, prolong, prolongb, prolong_size, restrict, restrictb, restrict_extrap,
                  restrict_size, synth_uniform, valgrad, valgrad_, valgrad_batch, valgrad_packed, valgrad_packed_, valgradhess, valgradhess_,
                  valgradhess_batch, wrap)

__version__ = "1.1.0"
