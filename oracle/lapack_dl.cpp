// TEST INFRASTRUCTURE -- see lapack_dl.hpp.
#include "lapack_dl.hpp"

#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <string>

namespace saamge_oracle
{

static void *must_sym(void *h, const char *prefixed, const char *plain)
{
    void *s = dlsym(h, prefixed);
    if (!s)
        s = dlsym(h, plain);
    if (!s)
    {
        std::fprintf(stderr, "oracle: LAPACK symbol %s not found\n", plain);
        std::abort();
    }
    return s;
}

const lapack_t &lapack(const char *path)
{
    static lapack_t L;
    static bool loaded = false;
    if (loaded)
        return L;
    const char *p = path ? path : std::getenv("SAAMGE_ORACLE_LAPACK");
    if (!p)
    {
        std::fprintf(stderr, "oracle: set SAAMGE_ORACLE_LAPACK to scipy's "
                             "libscipy_openblas*.so (see oracle/README.md)\n");
        std::abort();
    }
    void *h = dlopen(p, RTLD_NOW | RTLD_LOCAL);
    if (!h)
    {
        std::fprintf(stderr, "oracle: dlopen(%s) failed: %s\n", p, dlerror());
        std::abort();
    }
    L.dsygvx = (dsygvx_ft)must_sym(h, "scipy_dsygvx_", "dsygvx_");
    L.dgesvd = (dgesvd_ft)must_sym(h, "scipy_dgesvd_", "dgesvd_");
    L.dgels = (dgels_ft)must_sym(h, "scipy_dgels_", "dgels_");
    L.dlamch = (dlamch_ft)must_sym(h, "scipy_dlamch_", "dlamch_");
    L.dpotrf = (dpotrf_ft)must_sym(h, "scipy_dpotrf_", "dpotrf_");
    L.dpotrs = (dpotrs_ft)must_sym(h, "scipy_dpotrs_", "dpotrs_");
    void *snt = dlsym(h, "scipy_openblas_set_num_threads");
    if (!snt)
        snt = dlsym(h, "openblas_set_num_threads");
    L.set_num_threads = (void (*)(int))snt;
    if (L.set_num_threads)
        L.set_num_threads(1);
    loaded = true;
    return L;
}

} // namespace saamge_oracle
