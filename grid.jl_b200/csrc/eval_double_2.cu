// eval_double_2.cu -- instantiates the point-evaluation kernels for T=double, N=2 (see eval.cuh).
#include "eval.cuh"
namespace gb {
GB_DEFINE_EVAL(double, 2)
}
