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
#include <algorithm>
#include <fstream>

#include "saamge.hpp"

namespace saamge
{

namespace
{
/* raw native-endian records, the whole format: counts as int32, then the arrays */
template <class T> void get(std::istream &in, T *dst, size_t count = 1)
{
    in.read(reinterpret_cast<char *>(dst), (std::streamsize)(sizeof(T) * count));
    SA_ASSERT(in);
}
template <class T> void put(std::ostream &out, const T *src, size_t count = 1)
{
    out.write(reinterpret_cast<const char *>(src), (std::streamsize)(sizeof(T) * count));
    SA_ASSERT(out);
}
int get_count(std::istream &in)
{
    int v = 0;
    get(in, &v);
    SA_ASSERT(v >= 0);
    return v;
}
std::ifstream open_in(const char *filename)
{
    std::ifstream in(filename, std::ios::binary);
    SA_ASSERT(in);
    return in;
}
std::ofstream open_out(const char *filename)
{
    std::ofstream out(filename, std::ios::binary);
    SA_ASSERT(out);
    return out;
}
} // namespace

// amg/src/mbox.cpp:310-329: [rows][connections][I: rows + 1][J: connections]
Table *mbox_read_table(const char *filename)
{
    std::ifstream in = open_in(filename);
    Table *t = new Table;
    t->nrows = get_count(in);
    const int conn = get_count(in);
    t->I.resize((size_t)t->nrows + 1);
    t->J.resize((size_t)conn);
    get(in, t->I.data(), t->I.size());
    get(in, t->J.data(), t->J.size());
    SA_ASSERT(t->I[t->nrows] == conn);
    t->ncols = conn ? 1 + *std::max_element(t->J.begin(), t->J.end()) : 0;
    return t;
}

// amg/src/mbox.cpp:331-344
void mbox_write_table(const char *filename, const Table &tbl)
{
    std::ofstream out = open_out(filename);
    const int rows = tbl.Size(), conn = tbl.Size_of_connections();
    put(out, &rows);
    put(out, &conn);
    put(out, tbl.GetI(), (size_t)rows + 1);
    put(out, tbl.GetJ(), (size_t)conn);
}

// amg/src/mbox.cpp:355-375: [rows][cols][nnz][I][J][data]
SparseMatrix *mbox_read_sparse_matr(std::ifstream &in)
{
    SparseMatrix *m = new SparseMatrix;
    m->h = get_count(in);
    m->w = get_count(in);
    const int nnz = get_count(in);
    m->I.resize((size_t)m->h + 1);
    m->J.resize((size_t)nnz);
    m->A.resize((size_t)nnz);
    get(in, m->I.data(), m->I.size());
    get(in, m->J.data(), m->J.size());
    get(in, m->A.data(), m->A.size());
    SA_ASSERT(m->I[m->h] == nnz);
    return m;
}

SparseMatrix *mbox_read_sparse_matr(const char *filename)
{
    std::ifstream in = open_in(filename);
    return mbox_read_sparse_matr(in);
}

// amg/src/mbox.cpp:378-395
void mbox_write_sparse_matr(std::ofstream &out, const SparseMatrix &spm)
{
    const int rows = spm.Size(), cols = spm.Width(), nnz = spm.NumNonZeroElems();
    put(out, &rows);
    put(out, &cols);
    put(out, &nnz);
    put(out, spm.GetI(), (size_t)rows + 1);
    put(out, spm.GetJ(), (size_t)nnz);
    put(out, spm.GetData(), (size_t)nnz);
}

void mbox_write_sparse_matr(const char *filename, const SparseMatrix &spm)
{
    std::ofstream out = open_out(filename);
    mbox_write_sparse_matr(out, spm);
}

// amg/src/mbox.cpp:406-419: [rows][cols][column-major data]
DenseMatrix *mbox_read_dense_matr(std::ifstream &in)
{
    const int rows = get_count(in), cols = get_count(in);
    DenseMatrix *m = new DenseMatrix(rows, cols);
    get(in, m->Data(), (size_t)rows * cols);
    return m;
}

DenseMatrix *mbox_read_dense_matr(const char *filename)
{
    std::ifstream in = open_in(filename);
    return mbox_read_dense_matr(in);
}

// amg/src/mbox.cpp:438-446
void mbox_write_dense_matr(std::ofstream &out, const DenseMatrix &dem)
{
    const int rows = dem.Height(), cols = dem.Width();
    put(out, &rows);
    put(out, &cols);
    put(out, dem.Data(), (size_t)rows * cols);
}

void mbox_write_dense_matr(const char *filename, const DenseMatrix &dem)
{
    std::ofstream out = open_out(filename);
    mbox_write_dense_matr(out, dem);
}

// amg/src/mbox.cpp:448-481: [n] followed by n matrices back to back
SparseMatrix **mbox_read_sparse_matr_arr(const char *filename, int *n)
{
    std::ifstream in = open_in(filename);
    *n = get_count(in);
    SparseMatrix **arr = new SparseMatrix *[std::max(1, *n)];
    for (int k = 0; k < *n; ++k)
        arr[k] = mbox_read_sparse_matr(in);
    return arr;
}

void mbox_write_sparse_matr_arr(const char *filename, SparseMatrix **arr, int n)
{
    SA_ASSERT(arr && n > 0);
    std::ofstream out = open_out(filename);
    put(out, &n);
    for (int k = 0; k < n; ++k)
        mbox_write_sparse_matr(out, *arr[k]);
}

DenseMatrix **mbox_read_dense_matr_arr(const char *filename, int *n)
{
    std::ifstream in = open_in(filename);
    *n = get_count(in);
    DenseMatrix **arr = new DenseMatrix *[std::max(1, *n)];
    for (int k = 0; k < *n; ++k)
        arr[k] = mbox_read_dense_matr(in);
    return arr;
}

void mbox_write_dense_matr_arr(const char *filename, DenseMatrix **arr, int n)
{
    SA_ASSERT(arr && n > 0);
    std::ofstream out = open_out(filename);
    put(out, &n);
    for (int k = 0; k < n; ++k)
        mbox_write_dense_matr(out, *arr[k]);
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
