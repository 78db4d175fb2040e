// eval_inst.cu -- instantiates the evaluation kernels for one (T, N, ORDER); compiled 32 times
// with -DGB_T=<float|double> -DGB_N=<1..4> -DGB_O=<0..3> (see the Makefile).
#include "eval.cuh"
namespace gb {
#define GB_CAT_(a, b, c, d) a##_##b##_##c##_##d
#define GB_CAT(a, b, c, d) GB_CAT_(a, b, c, d)
int GB_CAT(launch_eval, GB_T, GB_N, GB_O)(const EvalParams<GB_N>& p, int bcf, int outmode, int flags, cudaStream_t s) {
    return launch_order<GB_T, GB_N, GB_O>(p, bcf, outmode, flags, s);
}
int GB_CAT(launch_outer, GB_T, GB_N, GB_O)(const EvalParams<GB_N>& p, int bcf, const i64* lens, int flags, void* out, cudaStream_t s) {
    return launch_outer_order<GB_T, GB_N, GB_O>(p, bcf, lens, flags, out, s);
}
}  // namespace gb
