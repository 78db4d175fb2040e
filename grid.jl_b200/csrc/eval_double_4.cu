// eval_double_4.cu -- instantiates the point-evaluation kernels for T=double, N=4 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(double, 4)
}
