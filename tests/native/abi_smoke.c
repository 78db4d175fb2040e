/* abi_smoke.c -- drives the C ABI from plain C with exactly the argument types the Julia `ccall`s of
 * grid.jl_b200/julia/GridB200.jl declare (Ptr{Cvoid} = void*, Cint = int, Int64 = int64_t, Cdouble = double,
 * Ptr{Int64} = int64_t*), so the binding's signatures are checked against include/gridb200.h by a C compiler.
 *   abi_smoke          prototype / link check only (no GPU needed): prints the version
 *   abi_smoke run      gb_grid_create -> gb_eval -> gb_grid_destroy on the current device (1-D quadratic BCnil, the
 *                      README's example: yi[1.9], valgradhess(yi, 4.25))
 * Test infrastructure (built by tests/test_abi.py and __graft_entry__.build()). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gridb200.h"

int main(int argc, char** argv) {
    /* every prototype the binding uses, taken through a typed function pointer: a mismatch is a compile error */
    int (*p_create)(gb_grid**, int, int, const int64_t*, const void*, int, int, int, double, void*) = gb_grid_create;
    int (*p_create_opt)(gb_grid**, int, int, const int64_t*, const void*, int, int, int, double, int, void*) = gb_grid_create_opt;
    int (*p_eval)(const gb_grid*, int, int64_t, const void*, int, int64_t, int64_t, const double*, void*, void*, void*, int, int, int64_t*, void*) = gb_eval;
    int (*p_outer)(const gb_grid*, const int64_t*, const void* const*, int, const double*, void*, int, int, void*) = gb_eval_outer;
    int (*p_invert)(int, int, const int64_t*, void*, int, int, int, const int*, int, int, void*) = gb_invert_opt;
    int (*p_restrict)(int, int, const int64_t*, const void*, int, double, void*, int, void*) = gb_restrict;
    int (*p_prolong)(int, int, const int64_t*, const void*, int, int64_t, void*, int, void*) = gb_prolong;
    int (*p_filledges)(int, int, const int64_t*, void*, double, int, void*) = gb_filledges;
    int (*p_comm_all)(gb_comm**, int, const int*) = gb_comm_init_all;
    int (*p_r3s)(gb_comm*, int, const int64_t*, const void* const*, double, void* const*, void* const*) = gb_restrict3_sharded;
    int (*p_p3s)(gb_comm*, int, const int64_t*, const void* const*, const int64_t*, void* const*, void* const*) = gb_prolong3_sharded;
    int (*p_evs)(gb_comm*, gb_grid* const*, int, int64_t, const void*, int, const double*, void*, void*, void*, int, int64_t*) = gb_eval_sharded;
    int (*p_repl)(gb_comm*, int, gb_grid**, int, int, const int64_t*, int, int, double) = gb_grid_replicate;
    int (*p_r3l)(int, const int64_t*, int, const void*, double, void*, int, void*) = gb_restrict3_levels;
    int (*p_p3l)(int, const int64_t*, int, const int64_t*, const void*, void*, int, void*) = gb_prolong3_levels;
    int (*p_r3sl)(gb_comm*, int, const int64_t*, int, const void* const*, double, void* const*, void* const*) = gb_restrict3_sharded_levels;
    int (*p_p3sl)(gb_comm*, int, const int64_t*, int, const int64_t*, const void* const*, void* const*, void* const*) = gb_prolong3_sharded_levels;
    int (*p_peer)(gb_comm*, int, int*) = gb_comm_peer_halo;
    (void)p_create_opt; (void)p_outer; (void)p_invert; (void)p_restrict; (void)p_prolong; (void)p_filledges; (void)p_comm_all; (void)p_r3s; (void)p_p3s;
    (void)p_evs; (void)p_repl; (void)p_r3l; (void)p_p3l; (void)p_r3sl; (void)p_p3sl; (void)p_peer;
    printf("gridb200 %d\n", gb_version());
    if (gb_restrict_size(1024) != 513 || gb_restrict_size(513) != 257) return 2;
    if (argc < 2 || strcmp(argv[1], "run") != 0) return 0;

    /* README.md:36-68: y = 8.1 (x - 2.3)^2 + 1.6 on x = 1..5, InterpGrid(y, BCnil, InterpQuadratic) */
    double y[5];
    for (int i = 0; i < 5; i++) y[i] = 8.1 * ((i + 1 - 2.3) * (i + 1 - 2.3)) + 1.6; /* a*(x-c).^2+o */
    const int64_t dims[1] = {5};
    gb_grid* g = NULL;
    int rc = p_create(&g, GB_F64, 1, dims, y, GB_HOST, GB_BC_NIL, GB_QUADRATIC, NAN, NULL);
    if (rc) { fprintf(stderr, "create: %s\n", gb_last_error()); return 3; }
    const double x[2] = {3.0, 4.25};
    double v[2], gr[2], h[2];
    int64_t bad = -7;
    rc = p_eval(g, GB_OUT_VAL | GB_OUT_GRAD | GB_OUT_HESS, 2, x, GB_F64, 1, 1, NULL, v, gr, h, GB_HOST, GB_EXACT, &bad, NULL);
    if (rc) { fprintf(stderr, "eval: %s\n", gb_last_error()); return 4; }
    printf("%.17g %.17g %.17g %.17g\n", v[0], v[1], gr[1], h[1]);
    /* printed by the reference itself (README.md:44-46, 66-68) */
    if (v[0] != 5.569000000000003 || v[1] != 32.40025 || gr[1] != 31.59000000000001 || h[1] != 16.200000000000017 || bad != -1) return 5;
    const double out_of_range[1] = {7.5};
    rc = p_eval(g, GB_OUT_VAL, 1, out_of_range, GB_F64, 1, 1, NULL, v, NULL, NULL, GB_HOST, GB_EXACT, &bad, NULL);
    if (rc != GB_ERR_INVALID_LOCATION || bad != 0) return 6; /* ErrorException("Invalid location"), src/interp.jl:402-404 */
    if (gb_grid_destroy(g)) return 7;
    printf("abi smoke ok\n");
    return 0;
}
