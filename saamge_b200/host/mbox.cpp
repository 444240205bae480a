// Binary on-disk formats of the reference (mbox_read_* / mbox_write_*, amg/src/mbox.cpp:310-483)
// on the POD containers of sa_types.hpp -- byte for byte the layouts upstream SAAMGE dumps use
// (native int32 / float64, no header beyond the sizes):
//   table          int size, int j_size, I[size + 1], J[j_size]
//   sparse matrix  int size, int width, int j_size, I[size + 1], J[j_size], data[j_size]
//   dense matrix   int height, int width, data[height * width] (column-major)
//   arrays         int n, then n matrices back to back
// and adapt_update_operators (amg/src/adapt.cpp:171-216): a new operator with the same pattern
// re-uses the spectral data -- only the weighted-l1 smoother, the smoothing of P and the Galerkin
// products are redone, on the device.
#include <fstream>

#include "saamge.hpp"

namespace saamge
{

Table *mbox_read_table(const char *filename)
{
    int size = 0, j_size = 0;
    std::ifstream itbl(filename, std::ifstream::binary);
    SA_ASSERT(itbl);
    itbl.read((char *)&size, sizeof(size));
    itbl.read((char *)&j_size, sizeof(j_size));
    SA_ASSERT(itbl && size >= 0 && j_size >= 0);
    Table *tbl = new Table;
    tbl->nrows = size;
    tbl->I.resize((size_t)size + 1);
    tbl->J.resize((size_t)j_size);
    itbl.read((char *)tbl->I.data(), sizeof(int) * ((size_t)size + 1));
    itbl.read((char *)tbl->J.data(), sizeof(int) * (size_t)j_size);
    SA_ASSERT(itbl);
    SA_ASSERT(j_size == tbl->I[size]);
    int w = 0;
    for (int j = 0; j < j_size; ++j)
        w = tbl->J[j] + 1 > w ? tbl->J[j] + 1 : w;
    tbl->ncols = w;
    return tbl;
}

void mbox_write_table(const char *filename, const Table &tbl)
{
    std::ofstream otbl(filename, std::ofstream::binary);
    SA_ASSERT(otbl);
    const int size = tbl.Size(), j_size = tbl.Size_of_connections();
    otbl.write((const char *)&size, sizeof(size));
    otbl.write((const char *)&j_size, sizeof(j_size));
    otbl.write((const char *)tbl.GetI(), sizeof(int) * ((size_t)size + 1));
    otbl.write((const char *)tbl.GetJ(), sizeof(int) * (size_t)j_size);
    SA_ASSERT(otbl);
}

SparseMatrix *mbox_read_sparse_matr(std::ifstream &ispm)
{
    int size = 0, width = 0, j_size = 0;
    SA_ASSERT(ispm);
    ispm.read((char *)&size, sizeof(size));
    ispm.read((char *)&width, sizeof(width));
    ispm.read((char *)&j_size, sizeof(j_size));
    SA_ASSERT(ispm && size >= 0 && width >= 0 && j_size >= 0);
    SparseMatrix *spm = new SparseMatrix;
    spm->h = size;
    spm->w = width;
    spm->I.resize((size_t)size + 1);
    spm->J.resize((size_t)j_size);
    spm->A.resize((size_t)j_size);
    ispm.read((char *)spm->I.data(), sizeof(int) * ((size_t)size + 1));
    ispm.read((char *)spm->J.data(), sizeof(int) * (size_t)j_size);
    ispm.read((char *)spm->A.data(), sizeof(double) * (size_t)j_size);
    SA_ASSERT(ispm);
    SA_ASSERT(j_size == spm->I[size]);
    return spm;
}

SparseMatrix *mbox_read_sparse_matr(const char *filename)
{
    std::ifstream ispm(filename, std::ifstream::binary);
    SA_ASSERT(ispm);
    return mbox_read_sparse_matr(ispm);
}

void mbox_write_sparse_matr(std::ofstream &ospm, const SparseMatrix &spm)
{
    SA_ASSERT(ospm);
    const int size = spm.Size(), width = spm.Width(), j_size = spm.NumNonZeroElems();
    ospm.write((const char *)&size, sizeof(size));
    ospm.write((const char *)&width, sizeof(width));
    ospm.write((const char *)&j_size, sizeof(j_size));
    ospm.write((const char *)spm.GetI(), sizeof(int) * ((size_t)size + 1));
    ospm.write((const char *)spm.GetJ(), sizeof(int) * (size_t)j_size);
    ospm.write((const char *)spm.GetData(), sizeof(double) * (size_t)j_size);
    SA_ASSERT(ospm);
}

void mbox_write_sparse_matr(const char *filename, const SparseMatrix &spm)
{
    std::ofstream ospm(filename, std::ofstream::binary);
    SA_ASSERT(ospm);
    mbox_write_sparse_matr(ospm, spm);
}

DenseMatrix *mbox_read_dense_matr(std::ifstream &idem)
{
    int height = 0, width = 0;
    SA_ASSERT(idem);
    idem.read((char *)&height, sizeof(height));
    idem.read((char *)&width, sizeof(width));
    SA_ASSERT(idem && height >= 0 && width >= 0);
    DenseMatrix *dem = new DenseMatrix(height, width);
    idem.read((char *)dem->Data(), sizeof(double) * (size_t)height * width);
    SA_ASSERT(idem);
    return dem;
}

DenseMatrix *mbox_read_dense_matr(const char *filename)
{
    std::ifstream idem(filename, std::ifstream::binary);
    SA_ASSERT(idem);
    return mbox_read_dense_matr(idem);
}

void mbox_write_dense_matr(std::ofstream &odem, const DenseMatrix &dem)
{
    SA_ASSERT(odem);
    const int height = dem.Height(), width = dem.Width();
    odem.write((const char *)&height, sizeof(height));
    odem.write((const char *)&width, sizeof(width));
    odem.write((const char *)dem.Data(), sizeof(double) * (size_t)height * width);
    SA_ASSERT(odem);
}

void mbox_write_dense_matr(const char *filename, const DenseMatrix &dem)
{
    std::ofstream odem(filename, std::ofstream::binary);
    SA_ASSERT(odem);
    mbox_write_dense_matr(odem, dem);
}

SparseMatrix **mbox_read_sparse_matr_arr(const char *filename, int *n)
{
    std::ifstream ispm(filename, std::ifstream::binary);
    SA_ASSERT(ispm);
    ispm.read((char *)n, sizeof(*n));
    SA_ASSERT(ispm && *n >= 0);
    SparseMatrix **arr = new SparseMatrix *[*n > 0 ? *n : 1];
    for (int i = 0; i < *n; ++i)
        arr[i] = mbox_read_sparse_matr(ispm);
    return arr;
}

void mbox_write_sparse_matr_arr(const char *filename, SparseMatrix **arr, int n)
{
    SA_ASSERT(arr);
    SA_ASSERT(n > 0);
    std::ofstream ospm(filename, std::ofstream::binary);
    SA_ASSERT(ospm);
    ospm.write((const char *)&n, sizeof(n));
    for (int i = 0; i < n; ++i)
        mbox_write_sparse_matr(ospm, *(arr[i]));
}

DenseMatrix **mbox_read_dense_matr_arr(const char *filename, int *n)
{
    std::ifstream idem(filename, std::ifstream::binary);
    SA_ASSERT(idem);
    idem.read((char *)n, sizeof(*n));
    SA_ASSERT(idem && *n >= 0);
    DenseMatrix **arr = new DenseMatrix *[*n > 0 ? *n : 1];
    for (int i = 0; i < *n; ++i)
        arr[i] = mbox_read_dense_matr(idem);
    return arr;
}

void mbox_write_dense_matr_arr(const char *filename, DenseMatrix **arr, int n)
{
    SA_ASSERT(arr);
    SA_ASSERT(n > 0);
    std::ofstream odem(filename, std::ofstream::binary);
    SA_ASSERT(odem);
    odem.write((const char *)&n, sizeof(n));
    for (int i = 0; i < n; ++i)
        mbox_write_dense_matr(odem, *(arr[i]));
}

/* ---- adapt_update_operators ---------------------------------------------------------------- */

// smpr_update_Dinv_neg (amg/inc/smpr.hpp:241): the weighted-l1 smoother of the level's operator
void smpr_update_Dinv_neg(tg_data_t &tg_data)
{
    SA_ASSERT(tg_data.gpu);
    sa_gpu_check(sa_gpu_build_Dinv_neg(tg_data.gpu), "sa_gpu_build_Dinv_neg");
}

// amg/inc/tg.hpp:678-693: re-smooth the kept tentative prolongator, drop the coarse operator
void tg_smooth_interp(tg_data_t &tg_data)
{
    SA_ASSERT(tg_data.gpu && tg_data.interp_data);
    interp_data_t &id = *tg_data.interp_data;
    sa_gpu_check(sa_gpu_smooth_P(tg_data.gpu, tg_data.smooth_interp ? id.interp_smoother_degree : 0,
                                 id.interp_smoother_roots),
                 "sa_gpu_smooth_P");
    if (tg_data.smooth_interp && id.interp_smoother_degree > 0 && id.drop_tol != 0.0)
        sa_gpu_check(sa_gpu_threshold_P(tg_data.gpu, id.drop_tol, NULL, NULL), "sa_gpu_threshold_P");
    tg_free_coarse_operator(tg_data);
}

// amg/inc/tg.hpp:735: the device Ac is overwritten by the next tg_update_coarse_operator
void tg_free_coarse_operator(tg_data_t &tg_data) { tg_data.have_Ac = false; }

// amg/src/adapt.cpp:171-187.  A == NULL: the level's operator is the finer level's Ac (already
// updated on the device); otherwise its values replace the finest operator's.
void adapt_update_operators(const SparseMatrix *A, tg_data_t &tg_data, bool resmooth_interp)
{
    SA_ASSERT(tg_data.poly_data);
    SA_ASSERT(tg_data.interp_data);
    SA_ASSERT(tg_data.gpu);
    if (A)
    {
        sa_gpu_check(sa_gpu_level_update_operator(tg_data.gpu, A->GetData()),
                     "sa_gpu_level_update_operator");
        if (tg_data.A_host && tg_data.A_host != A && !tg_data.A_host_owned)
            tg_data.A_host = A;
    }
    smpr_update_Dinv_neg(tg_data);
    if (resmooth_interp && tg_data.smooth_interp && tg_data.interp_data->interp_smoother_degree > 0 &&
        tg_data.interp_data->times_apply_smoother > 0)
        tg_smooth_interp(tg_data);
    tg_free_coarse_operator(tg_data);
}

// amg/src/adapt.cpp:189-216
void adapt_update_operators(const SparseMatrix &A, ml_data_t &ml_data, const MultilevelParameters &mlp,
                            bool resmooth_interp)
{
    SA_ASSERT(ml_data.levels_list.num_levels > 0);
    SA_ASSERT(ml_data.levels_list.finest);
    SA_ASSERT(ml_data.levels_list.finest->tg_data);
    adapt_update_operators(&A, *ml_data.levels_list.finest->tg_data, resmooth_interp);
    tg_update_coarse_operator(ml_data.levels_list.finest->tg_data,
                              NULL == ml_data.levels_list.finest->coarser, mlp.get_coarse_direct());
    for (levels_level_t *level = ml_data.levels_list.finest->coarser; level; level = level->coarser)
    {
        SA_ASSERT(level->tg_data);
        // the level's operator is the finer level's (new) Ac: a stale host copy is dropped
        delete level->tg_data->A_host_owned;
        level->tg_data->A_host_owned = NULL;
        level->tg_data->A_host = NULL;
        adapt_update_operators(NULL, *level->tg_data, resmooth_interp);
        tg_update_coarse_operator(level->tg_data, NULL == level->coarser, mlp.get_coarse_direct());
    }
    if (ml_data.correct_nullspace_level)
    {
        // the CorrectNullspace level's operator is the (new) coarsest Ac; the scaling P stays
        sa_gpu_check(sa_gpu_build_Dinv_neg(ml_data.correct_nullspace_level), "sa_gpu_build_Dinv_neg");
        sa_gpu_check(sa_gpu_rap(ml_data.correct_nullspace_level), "sa_gpu_rap");
    }
    ml_impose_cycle(ml_data, false);
}

} // namespace saamge

/* ---- C entry points for the format tests (tests/test_mbox_io.py) ---- */
using namespace saamge;

extern "C" int sa_drv_mbox_write_sparse(const char *fn, int h, int w, const int *I, const int *J,
                                        const double *A)
{
    SparseMatrix M;
    M.h = h;
    M.w = w;
    M.I.assign(I, I + h + 1);
    M.J.assign(J, J + I[h]);
    M.A.assign(A, A + I[h]);
    mbox_write_sparse_matr(fn, M);
    return 0;
}

/* I / J / A may be NULL (sizes only) */
extern "C" int sa_drv_mbox_read_sparse(const char *fn, int *h, int *w, int *nnz, int *I, int *J, double *A)
{
    SparseMatrix *M = mbox_read_sparse_matr(fn);
    *h = M->h;
    *w = M->w;
    *nnz = M->NumNonZeroElems();
    if (I)
        std::copy(M->I.begin(), M->I.end(), I);
    if (J)
        std::copy(M->J.begin(), M->J.end(), J);
    if (A)
        std::copy(M->A.begin(), M->A.end(), A);
    delete M;
    return 0;
}

extern "C" int sa_drv_mbox_write_table(const char *fn, int nrows, const int *I, const int *J)
{
    Table T;
    T.nrows = nrows;
    T.I.assign(I, I + nrows + 1);
    T.J.assign(J, J + I[nrows]);
    mbox_write_table(fn, T);
    return 0;
}

extern "C" int sa_drv_mbox_read_table(const char *fn, int *nrows, int *nconn, int *I, int *J)
{
    Table *T = mbox_read_table(fn);
    *nrows = T->Size();
    *nconn = T->Size_of_connections();
    if (I)
        std::copy(T->I.begin(), T->I.end(), I);
    if (J)
        std::copy(T->J.begin(), T->J.end(), J);
    delete T;
    return 0;
}

/* n dense matrices of sizes hs[i] x ws[i], concatenated column-major in data */
extern "C" int sa_drv_mbox_write_dense_arr(const char *fn, int n, const int *hs, const int *ws,
                                           const double *data)
{
    std::vector<DenseMatrix *> arr(n);
    size_t o = 0;
    for (int i = 0; i < n; ++i)
    {
        arr[i] = new DenseMatrix(hs[i], ws[i]);
        std::copy(data + o, data + o + (size_t)hs[i] * ws[i], arr[i]->Data());
        o += (size_t)hs[i] * ws[i];
    }
    if (n == 1)
        mbox_write_dense_matr(fn, *arr[0]);
    else
        mbox_write_dense_matr_arr(fn, arr.data(), n);
    for (int i = 0; i < n; ++i)
        delete arr[i];
    return 0;
}

/* reads an array file; returns the number of matrices, sizes into hs / ws (capacity cap), the
   concatenated data into `data` when not NULL */
extern "C" int sa_drv_mbox_read_dense_arr(const char *fn, int cap, int *hs, int *ws, double *data)
{
    int n = 0;
    DenseMatrix **arr = mbox_read_dense_matr_arr(fn, &n);
    size_t o = 0;
    for (int i = 0; i < n; ++i)
    {
        if (i < cap)
        {
            hs[i] = arr[i]->Height();
            ws[i] = arr[i]->Width();
        }
        if (data)
            std::copy(arr[i]->Data(), arr[i]->Data() + (size_t)arr[i]->Height() * arr[i]->Width(), data + o);
        o += (size_t)arr[i]->Height() * arr[i]->Width();
        delete arr[i];
    }
    delete[] arr;
    return n;
}
