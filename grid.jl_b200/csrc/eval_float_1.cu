// eval_float_1.cu -- instantiates the point-evaluation kernels for T=float, N=1 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(float, 1)
}
