// Host-side check of the GB_FAST prefilter ARITHMETIC (grid.jl_b200/csrc/prefilter_fast.cuh): the same segment / end-row
// functions the kernels run, compiled for the CPU, applied to one line read from stdin; the solution goes to stdout.
// tests/test_fast_prefilter_host.py compares it with the oracle's exact solve.  Test infrastructure only.
//   fast_prefilter_check <f32|f64> <bc family 0..3> <order 2|3> <n> <nseg>
#define GB_FAST_FN inline
#include "../../grid.jl_b200/csrc/prefilter_fast.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace gb;

template <typename T, int E> struct HostIO {
    const T* src;
    T* dst;
    i64 n;
    int periodic;
    void load(i64 row0, T (&v)[E]) const {
        for (int e = 0; e < E; e++) {
            i64 r = row0 + e;
            if (periodic) r = r < 0 ? r + n : (r >= n ? r - n : r);
            v[e] = (r >= 0 && r < n) ? src[r] : (T)0;
        }
    }
    void store(i64 row0, const T (&v)[E]) const {
        for (int e = 0; e < E; e++)
            if (row0 + e < n) dst[row0 + e] = v[e];
    }
};

template <typename T> int run(int fam, int order, i64 n, int nseg) {
    constexpr int E = 128 / (int)sizeof(T), KB = FAST_K / E;
    std::vector<T> f(n), x(n);
    if (fread(f.data(), sizeof(T), n, stdin) != (size_t)n) return 2;
    FastHost<T> h;
    h.setup(fam, order, n);
    const i64 boxes = (n + E - 1) / E;
    const i64 seg_rows = (boxes + nseg - 1) / nseg * E;
    for (i64 r0 = 0; r0 < n; r0 += seg_rows) {
        HostIO<T, E> io{f.data(), x.data(), n, h.periodic};
        fast_segment<T, E, KB>(io, r0, std::min<i64>(n, r0 + seg_rows), h.l, h.rd);
    }
    if (!h.periodic) {
        for (int end = 0; end < 2; end++) {
            const T* fi = f.data() + (end ? n - FAST_COLS : 0);
            T* o = x.data() + (end ? n - FAST_K : 0);
            fast_end_rows<T>(end ? h.wbot.data() : h.wtop.data(), [&](int j) { return fi[j]; }, [&](int i, T v) { o[i] = v; });
        }
    }
    fwrite(x.data(), sizeof(T), n, stdout);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 6) return 1;
    const int fam = atoi(argv[2]), order = atoi(argv[3]), nseg = atoi(argv[5]);
    const i64 n = atoll(argv[4]);
    if (n < FAST_MIN_N) return 3;
    return strcmp(argv[1], "f32") == 0 ? run<float>(fam, order, n, nseg) : run<double>(fam, order, n, nseg);
}
