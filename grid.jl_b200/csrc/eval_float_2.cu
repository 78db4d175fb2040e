// eval_float_2.cu -- instantiates the point-evaluation kernels for T=float, N=2 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(float, 2)
}
