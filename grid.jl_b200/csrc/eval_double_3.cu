// eval_double_3.cu -- instantiates the point-evaluation kernels for T=double, N=3 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(double, 3)
}
