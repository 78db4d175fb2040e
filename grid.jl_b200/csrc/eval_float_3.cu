// eval_float_3.cu -- instantiates the point-evaluation kernels for T=float, N=3 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(float, 3)
}
