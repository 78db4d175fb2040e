// eval_float_4.cu -- instantiates the point-evaluation kernels for T=float, N=4 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(float, 4)
}
