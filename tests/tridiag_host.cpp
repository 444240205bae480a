// Host build of the per-thread tridiagonal routines the CUDA kernels run
// (saamge_b200/csrc/tridiag_math.cuh), for CPU unit tests.
#include "../saamge_b200/csrc/tridiag_math.cuh"

#include <vector>

extern "C" int th_sturm_count(int n, const double *d, const double *e, double x)
{
    std::vector<double> e2(n > 1 ? n - 1 : 1, 0.);
    double e2max = 0.;
    for (int i = 0; i + 1 < n; ++i)
    {
        e2[i] = e[i] * e[i];
        e2max = fmax(e2max, e2[i]);
    }
    return sa_sturm_count(n, d, e2.data(), x, DBL_MIN * fmax(1., e2max));
}

extern "C" void th_gershgorin(int n, const double *d, const double *e, double *out3)
{
    sa_gershgorin(n, d, e, &out3[0], &out3[1], &out3[2]);
}

// solves (T - shift I) x = b in place, lane-interleaved workspace like the kernel
extern "C" void th_shifted_solve(int n, const double *d, const double *e, double shift,
                                 double pivtol, double *x, int stride, int lane)
{
    std::vector<double> u0((size_t)n * stride), u1((size_t)n * stride), u2((size_t)n * stride),
        mu((size_t)n * stride);
    std::vector<int> sw((size_t)n * stride);
    sa_tridiag_lu_factor(n, d, e, shift, pivtol, u0.data() + lane, u1.data() + lane,
                         u2.data() + lane, mu.data() + lane, sw.data() + lane, stride);
    sa_tridiag_lu_solve(n, u0.data() + lane, u1.data() + lane, u2.data() + lane, mu.data() + lane,
                        sw.data() + lane, stride, x, 1);
}

extern "C" double th_hash_uniform(unsigned long long key) { return sa_hash_uniform(key); }
