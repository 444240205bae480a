// Per-thread sequential kernels of the symmetric tridiagonal eigen-stage:
// Sturm counts (the bisection of dstebz/dlaebz) and the pivoted LU solve used by
// inverse iteration (the role of dlagtf/dlagts inside dstein).  These are what
// LAPACK's dsyevx runs after dsytrd on the path the reference takes through
// dsygvx (amg/src/xpacks.cpp:260-267); LAPACK itself is not in the reference
// tree, so this follows the published algorithms.  __host__ __device__ so that the
// same code is unit-tested on the CPU (tests/test_tridiag_math.py).
#ifndef SA_TRIDIAG_MATH_CUH
#define SA_TRIDIAG_MATH_CUH

#include <cfloat>
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define SA_HD __host__ __device__ __forceinline__
#else
#define SA_HD inline
#endif

/// Number of eigenvalues of T = tridiag(e, d, e) that are <= x.
/// e2[i] = e[i]^2.  Strides allow interleaved storage.
SA_HD int sa_sturm_count(int n, const double *d, const double *e2, double x, double pivmin)
{
    int cnt = 0;
    double q = d[0] - x;
    if (fabs(q) < pivmin)
        q = -pivmin;
    cnt += (q <= 0.);
    for (int i = 1; i < n; ++i)
    {
        q = d[i] - e2[i - 1] / q - x;
        if (fabs(q) < pivmin)
            q = -pivmin;
        cnt += (q <= 0.);
    }
    return cnt;
}

/// Gershgorin interval of T and the norms used for tolerances.
SA_HD void sa_gershgorin(int n, const double *d, const double *e, double *gl, double *gu,
                         double *e2max)
{
    double lo = DBL_MAX, hi = -DBL_MAX, em = 0.;
    for (int i = 0; i < n; ++i)
    {
        const double l = (i > 0) ? fabs(e[i - 1]) : 0.;
        const double r = (i < n - 1) ? fabs(e[i]) : 0.;
        lo = fmin(lo, d[i] - l - r);
        hi = fmax(hi, d[i] + l + r);
        if (i < n - 1)
            em = fmax(em, e[i] * e[i]);
    }
    *gl = lo;
    *gu = hi;
    *e2max = em;
}

SA_HD double sa_hash_uniform(uint64_t key)
{
    // splitmix64 -> uniform in (-1, 1)
    uint64_t z = key + 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z = z ^ (z >> 31);
    return ((double)(z >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

/// LU factorisation with partial pivoting of (T - shift I), stored for repeated
/// solves.  Workspace arrays are accessed as a[i * stride]:
///   u0inv: reciprocal of the pivot, u1/u2: first/second superdiagonal of U,
///   mult: multiplier, swp: 1 if rows i and i+1 were interchanged.
SA_HD void sa_tridiag_lu_factor(int n, const double *d, const double *e, double shift,
                                double pivtol, double *u0inv, double *u1, double *u2,
                                double *mult, int *swp, int64_t stride)
{
    double a = d[0] - shift;          // current diagonal
    double b = (n > 1) ? e[0] : 0.;   // current superdiagonal
    for (int i = 0; i < n - 1; ++i)
    {
        const double l = e[i];                         // subdiagonal entry of row i+1
        const double dn = d[i + 1] - shift;            // diagonal of row i+1
        const double en = (i + 1 < n - 1) ? e[i + 1] : 0.; // superdiagonal of row i+1
        if (fabs(a) >= fabs(l))
        {
            if (fabs(a) < pivtol)
                a = (a < 0.) ? -pivtol : pivtol;
            const double inv = 1. / a;
            const double m = l * inv;
            u0inv[i * stride] = inv;
            u1[i * stride] = b;
            u2[i * stride] = 0.;
            mult[i * stride] = m;
            swp[i * stride] = 0;
            a = dn - m * b;
            b = en;
        }
        else
        {
            const double inv = 1. / l;
            const double m = a * inv;
            u0inv[i * stride] = inv;
            u1[i * stride] = dn;
            u2[i * stride] = en;
            mult[i * stride] = m;
            swp[i * stride] = 1;
            a = b - m * dn;
            b = -m * en;
        }
    }
    if (fabs(a) < pivtol)
        a = (a < 0.) ? -pivtol : pivtol;
    u0inv[(n - 1) * stride] = 1. / a;
    u1[(n - 1) * stride] = 0.;
    u2[(n - 1) * stride] = 0.;
    mult[(n - 1) * stride] = 0.;
    swp[(n - 1) * stride] = 0;
}

/// Solves (T - shift I) x = x in place with the stored factorisation; x[i * xstride].
SA_HD void sa_tridiag_lu_solve(int n, const double *u0inv, const double *u1, const double *u2,
                               const double *mult, const int *swp, int64_t stride, double *x,
                               int64_t xstride)
{
    // forward: apply the row interchanges and eliminations to the right-hand side
    double cur = x[0];
    for (int i = 0; i < n - 1; ++i)
    {
        double nxt = x[(i + 1) * xstride];
        if (swp[i * stride])
        {
            const double t = cur;
            cur = nxt;
            nxt = t;
        }
        x[i * xstride] = cur;
        cur = nxt - mult[i * stride] * cur;
    }
    x[(n - 1) * xstride] = cur;
    // backward
    double x1 = 0., x2 = 0.;
    for (int i = n - 1; i >= 0; --i)
    {
        const double xi =
            (x[i * xstride] - u1[i * stride] * x1 - u2[i * stride] * x2) * u0inv[i * stride];
        x[i * xstride] = xi;
        x2 = x1;
        x1 = xi;
    }
}

#endif
