// TEST INFRASTRUCTURE -- part of the CPU oracle, never linked into the product.
// LAPACK entry points the reference calls (amg/src/xpacks.cpp:35-72), resolved at
// run time from scipy's bundled OpenBLAS (LP64, symbols prefixed "scipy_").
#ifndef SAAMGE_ORACLE_LAPACK_DL_HPP
#define SAAMGE_ORACLE_LAPACK_DL_HPP

namespace saamge_oracle
{

typedef void (*dsygvx_ft)(int *itype, char *jobz, char *range, char *uplo, int *n,
                          double *a, int *lda, double *b, int *ldb, double *vl,
                          double *vu, int *il, int *iu, double *abstol, int *m,
                          double *w, double *z, int *ldz, double *work, int *lwork,
                          int *iwork, int *ifail, int *info);
typedef void (*dgesvd_ft)(char *jobu, char *jobvt, int *m, int *n, double *a, int *lda,
                          double *s, double *u, int *ldu, double *vt, int *ldvt,
                          double *work, int *lwork, int *info);
typedef void (*dgels_ft)(char *trans, int *m, int *n, int *nrhs, double *a, int *lda,
                         double *b, int *ldb, double *work, int *lwork, int *info);
typedef double (*dlamch_ft)(char *cmach);
typedef void (*dpotrf_ft)(char *uplo, int *n, double *a, int *lda, int *info);
typedef void (*dpotrs_ft)(char *uplo, int *n, int *nrhs, double *a, int *lda, double *b,
                          int *ldb, int *info);

struct lapack_t
{
    dsygvx_ft dsygvx;
    dgesvd_ft dgesvd;
    dgels_ft dgels;
    dlamch_ft dlamch;
    dpotrf_ft dpotrf;
    dpotrs_ft dpotrs;
    void (*set_num_threads)(int);
};

/// Loads the library at \a path (NULL: use $SAAMGE_ORACLE_LAPACK).  Aborts with a
/// message if it cannot be loaded.  Sets OpenBLAS to one thread (the oracle
/// parallelises over AEs / MISes with OpenMP, like one MPI rank per core).
const lapack_t &lapack(const char *path = 0);

} // namespace saamge_oracle

#endif
