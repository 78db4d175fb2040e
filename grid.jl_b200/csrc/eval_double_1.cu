// eval_double_1.cu -- instantiates the point-evaluation kernels for T=double, N=1 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(double, 1)
}
