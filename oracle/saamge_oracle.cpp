// TEST INFRASTRUCTURE -- CPU oracle for the spectral-AMGe hot path (see
// saamge_oracle.hpp).  Every function cites the reference lines it follows;
// paths are relative to /root/reference/.
#include "saamge_oracle.hpp"
#include "lapack_dl.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <omp.h>

namespace saamge_oracle
{

static const double DIFF_EPS = 1e-10; // GLOBAL diff_eps, amg/inc/config.hpp:68
#define SA_REAL_ALMOST_LE(x, y) ((x) <= ((y) + DIFF_EPS)) // amg/inc/common.hpp:481-492

static double now_s()
{
    return std::chrono::duration<double>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

/* mfem::SparseMatrix in its LIL phase: Set overwrites, Add accumulates, Finalize
   produces CSR.  Columns are emitted ascending (ordering rule of this repo). */
struct LilMatrix
{
    int h, w;
    std::vector<std::vector<std::pair<int, double>>> rows;
    LilMatrix(int h_, int w_) : h(h_), w(w_), rows(h_) {}
    double *find(int i, int j)
    {
        std::vector<std::pair<int, double>> &r = rows[i];
        for (size_t k = 0; k < r.size(); ++k)
            if (r[k].first == j)
                return &r[k].second;
        return NULL;
    }
    void Set(int i, int j, double v)
    {
        double *p = find(i, j);
        if (p)
            *p = v;
        else
            rows[i].push_back(std::make_pair(j, v));
    }
    void Add(int i, int j, double v)
    {
        double *p = find(i, j);
        if (p)
            *p += v;
        else
            rows[i].push_back(std::make_pair(j, v));
    }
    SparseMatrix *Finalize()
    {
        SparseMatrix *S = new SparseMatrix;
        S->h = h;
        S->w = w;
        S->I.assign((size_t)h + 1, 0);
        for (int i = 0; i < h; ++i)
            S->I[i + 1] = S->I[i] + (int)rows[i].size();
        S->J.resize(S->I[h]);
        S->A.resize(S->I[h]);
        for (int i = 0; i < h; ++i)
        {
            std::sort(rows[i].begin(), rows[i].end());
            int q = S->I[i];
            for (size_t k = 0; k < rows[i].size(); ++k, ++q)
            {
                S->J[q] = rows[i][k].first;
                S->A[q] = rows[i][k].second;
            }
        }
        return S;
    }
};

/* ------------------------------------------------------------------ assembly */

// amg/src/aggregates.cpp:68-184
double agg_assemble_value(int di, int dj, int part,
                          const agg_partitioning_relations_t &agg_part_rels,
                          const ElementMatrixProvider *data)
{
    const int *partitioning = agg_part_rels.partitioning;
    const Table &dof_to_elem = *agg_part_rels.dof_to_elem;
    const Table &elem_to_dof = *agg_part_rels.elem_to_dof;
    double value = 0.;
    const int rsi = dof_to_elem.RowSize(di);
    const int rsj = dof_to_elem.RowSize(dj);
    // rows of a transposed Table are already ascending (rowi.Sort(), rowj.Sort())
    const int *rowi = dof_to_elem.GetRow(di);
    const int *rowj = dof_to_elem.GetRow(dj);
    bool free_matr;
    int i, k, j = 0;
    for (i = 0; i < rsi; ++i)
    {
        const int elno = rowi[i];
        if (partitioning[elno] != part)
            continue;
        if (di != dj)
        {
            while (j < rsj && rowj[j] < elno)
                ++j;
            if (j >= rsj)
                break;
            if (rowj[j] != elno)
                continue;
        }
        const int ndofs = elem_to_dof.RowSize(elno);
        const int *const dofs = elem_to_dof.GetRow(elno);
        int dii = -1, djj = -1;
        for (k = 0; k < ndofs && (dii < 0 || djj < 0); ++k)
        {
            if (dofs[k] == di)
                dii = k;
            if (dofs[k] == dj)
                djj = k;
        }
        SA_ASSERT(dii >= 0 && djj >= 0);
        const Matrix *elem_matr = data->GetMatrix(elno, free_matr);
        const DenseMatrix *elmat = dynamic_cast<const DenseMatrix *>(elem_matr);
        if (elmat)
        {
            value += (*elmat)(dii, djj);
            if (free_matr)
                delete elmat;
        }
        else
        {
            const SparseMatrix *spm = static_cast<const SparseMatrix *>(elem_matr);
            value += (*spm)(dii, djj);
            if (free_matr)
                delete spm;
        }
    }
    return value;
}

// amg/src/aggregates.cpp:855-945
SparseMatrix *agg_build_AE_stiffm_with_global(
    const SparseMatrix &A, int part, const agg_partitioning_relations_t &agg_part_rels,
    const ElementMatrixProvider *data, bool bdr_cond_imposed, bool assemble_ess_diag)
{
    const int *const row = agg_part_rels.AE_to_dof->GetRow(part);
    const int rs = agg_part_rels.AE_to_dof->RowSize(part);
    LilMatrix AE_stiffm(rs, rs);
    std::vector<char> diag(rs, 0);

    for (int i = 0; i < rs; ++i)
    {
        const int glob_dof = row[i];
        const int row_start = A.GetI()[glob_dof];
        const int *neighbours = &(A.GetJ()[row_start]);
        const double *neigh_data = &(A.GetData()[row_start]);
        const int row_size = A.RowSize(glob_dof);
        for (int j = 0; j < row_size; ++j)
        {
            const int glob_neigh = neighbours[j];
            if (agg_elem_in_col(glob_neigh, part, *agg_part_rels.dof_to_AE) < 0)
                continue;
            const int local_neigh = agg_map_id_glob_to_AE(glob_neigh, part, agg_part_rels);
            SA_ASSERT(0 <= local_neigh && local_neigh < rs);

            if (SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[glob_dof], AGG_BETWEEN_AES_FLAG) &&
                SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[glob_neigh], AGG_BETWEEN_AES_FLAG) &&
                !(bdr_cond_imposed &&
                  (SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[glob_dof],
                                    AGG_ON_ESS_DOMAIN_BORDER_FLAG) ||
                   SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[glob_neigh],
                                    AGG_ON_ESS_DOMAIN_BORDER_FLAG)) &&
                  !(assemble_ess_diag && glob_neigh == glob_dof)))
            {
                if (i < local_neigh || (i == local_neigh && !(diag[i])))
                {
                    const double value =
                        agg_assemble_value(glob_dof, glob_neigh, part, agg_part_rels, data);
                    if (0. != value)
                        AE_stiffm.Set(i, local_neigh, value);
                    if (i != local_neigh)
                    {
                        if (0. != value)
                            AE_stiffm.Set(local_neigh, i, value);
                    }
                    else
                        diag[i] = true;
                }
            }
            else
            {
                if (0. != neigh_data[j])
                    AE_stiffm.Set(i, local_neigh, neigh_data[j]);
            }
        }
    }
    return AE_stiffm.Finalize();
}

// amg/src/aggregates.cpp:959-1086
SparseMatrix *agg_build_AE_stiffm(int part,
                                  const agg_partitioning_relations_t &agg_part_rels,
                                  const ElementMatrixProvider *data)
{
    bool free_matr;
    const int *const AEelems = agg_part_rels.AE_to_elem->GetRow(part);
    const int num_AEelems = agg_part_rels.AE_to_elem->RowSize(part);
    const int num_AEdofs = agg_part_rels.AE_to_dof->RowSize(part);
    LilMatrix AE_stiffm(num_AEdofs, num_AEdofs);
    SA_ASSERT(num_AEelems > 0);
    std::vector<int> local;
    for (int i = 0; i < num_AEelems; ++i)
    {
        const int elem = AEelems[i];
        const Matrix *matr = data->GetMatrix(elem, free_matr);
        const int *const elemdofs = agg_part_rels.elem_to_dof->GetRow(elem);
        const int elem_matr_sz = agg_part_rels.elem_to_dof->RowSize(elem);
        local.resize(elem_matr_sz);
        for (int k = 0; k < elem_matr_sz; ++k)
        {
            local[k] = agg_map_id_glob_to_AE(elemdofs[k], part, agg_part_rels);
            SA_ASSERT(0 <= local[k] && local[k] < num_AEdofs);
        }
        const SparseMatrix *elem_matr = dynamic_cast<const SparseMatrix *>(matr);
        if (elem_matr)
        {
            SA_ASSERT(elem_matr->Size() == elem_matr_sz);
            const int *I = elem_matr->GetI();
            const int *J = elem_matr->GetJ();
            const double *Data = elem_matr->GetData();
            for (int k = 0; k < elem_matr_sz; ++k)
                for (int j = I[k]; j < I[k + 1]; ++j)
                {
                    const double el = Data[j];
                    if (0. != el)
                        AE_stiffm.Add(local[k], local[J[j]], el);
                }
            if (free_matr)
                delete elem_matr;
        }
        else
        {
            const DenseMatrix *elem_dmatr = static_cast<const DenseMatrix *>(matr);
            SA_ASSERT(elem_dmatr->Height() == elem_matr_sz);
            for (int k = 0; k < elem_matr_sz; ++k)
                for (int j = 0; j < elem_matr_sz; ++j)
                {
                    const double el = (*elem_dmatr)(k, j);
                    if (0. != el)
                        AE_stiffm.Add(local[k], local[j], el);
                }
            if (free_matr)
                delete elem_dmatr;
        }
    }
    return AE_stiffm.Finalize();
}

// amg/src/aggregates.cpp:1143-1179 (rows of cut_evects at the MIS's dofs)
void agg_restrict_to_agg_enforce(int part,
                                 const agg_partitioning_relations_t &agg_part_rels,
                                 int agg_size, const int *restriction,
                                 const DenseMatrix &cut_evects, DenseMatrix &restricted)
{
    SA_ASSERT(cut_evects.Height() == agg_part_rels.AE_to_dof->RowSize(part));
    const int num_vects = cut_evects.Width();
    restricted.SetSize(agg_size, num_vects);
    for (int i = 0; i < agg_size; ++i)
    {
        const int AE_dof = agg_map_id_glob_to_AE(restriction[i], part, agg_part_rels);
        SA_ASSERT(0 <= AE_dof && AE_dof < cut_evects.Height());
        for (int v = 0; v < num_vects; ++v)
            restricted(i, v) = cut_evects(AE_dof, v);
    }
}

/* ------------------------------------------------------------------- toolbox */

// amg/src/mbox.cpp:913-949
SparseMatrix *mbox_snd_D_sparse_from_sparse(const SparseMatrix &A)
{
    const int n = A.Size();
    SparseMatrix *B = new SparseMatrix;
    B->h = B->w = n;
    B->I.resize((size_t)n + 1);
    B->J.resize(n);
    B->A.resize(n);
    for (int i = 0; i < n; ++i)
    {
        B->J[i] = B->I[i] = i;
        double sum = 0.;
        const double diag = A(i, i);
        SA_ASSERT(diag > 0.);
        const int beg = A.GetI()[i];
        const int *row = A.GetJ() + beg;
        const double *a = A.GetData() + beg;
        const int a_rsz = A.RowSize(i);
        for (int j = 0; j < a_rsz; ++j)
        {
            SA_ASSERT(A(row[j], row[j]) > 0.);
            sum += fabs(a[j]) * sqrt(diag / A(row[j], row[j]));
        }
        SA_ASSERT(sum > 0.);
        B->A[i] = sum;
    }
    B->I[n] = n;
    return B;
}

// amg/src/mbox.cpp:485-505
void mbox_convert_sparse_to_dense(const SparseMatrix &Sp, DenseMatrix &De)
{
    De.SetSize(Sp.Size(), Sp.Width());
    for (int i = 0; i < Sp.Size(); ++i)
        for (int p = Sp.I[i]; p < Sp.I[i + 1]; ++p)
            De(i, Sp.J[p]) = Sp.A[p];
}

// amg/src/xpacks.cpp:222-314
int xpacks_calc_lower_eigens_dense(const DenseMatrix &Ain, Vector &evals,
                                   DenseMatrix &evects, const DenseMatrix &Bin,
                                   double upper, bool atleast_one)
{
    const lapack_t &L = lapack();
    int itype = 1;
    char jobz = 'V';
    char range = 'V';
    char uplo = 'U';
    int n = Ain.Height();
    int lda = n;
    int ldb = n;
    double vl = -1.;
    double vu = upper;
    char cmach = 'S';
    double abstol = 2. * L.dlamch(&cmach);
    int m;
    int ldz = n;
    int lwork;
    std::vector<int> iwork((size_t)5 * n), ifail(n);
    int info;
    std::vector<double> A((size_t)n * n), B((size_t)n * n), w(n), z((size_t)n * n);
    int il_dummy = 0, iu_dummy = 0;

    SA_ASSERT(n > 0);
    std::memcpy(B.data(), Bin.Data(), sizeof(double) * n * n);
    std::memcpy(A.data(), Ain.Data(), sizeof(double) * n * n);

    lwork = -1;
    double qwork;
    L.dsygvx(&itype, &jobz, &range, &uplo, &n, A.data(), &lda, B.data(), &ldb, &vl, &vu,
             &il_dummy, &iu_dummy, &abstol, &m, w.data(), z.data(), &ldz, &qwork, &lwork,
             iwork.data(), ifail.data(), &info);
    SA_ASSERT(!info);
    lwork = (int)qwork + 1;
    std::vector<double> work(lwork);

    L.dsygvx(&itype, &jobz, &range, &uplo, &n, A.data(), &lda, B.data(), &ldb, &vl, &vu,
             &il_dummy, &iu_dummy, &abstol, &m, w.data(), z.data(), &ldz, work.data(),
             &lwork, iwork.data(), ifail.data(), &info);
    SA_ASSERT(!info);

    if (atleast_one && 0 >= m)
    {
        int il = 1;
        int iu = 1;
        range = 'I';
        std::memcpy(B.data(), Bin.Data(), sizeof(double) * n * n);
        std::memcpy(A.data(), Ain.Data(), sizeof(double) * n * n);
        L.dsygvx(&itype, &jobz, &range, &uplo, &n, A.data(), &lda, B.data(), &ldb, &vl, &vu,
                 &il, &iu, &abstol, &m, w.data(), z.data(), &ldz, work.data(), &lwork,
                 iwork.data(), ifail.data(), &info);
        SA_ASSERT(!info);
        SA_ASSERT(1 == m);
    }
    evals.assign(w.begin(), w.begin() + m);
    evects.SetSize(n, m);
    std::memcpy(evects.Data(), z.data(), sizeof(double) * n * m);
    return m;
}

// amg/src/xpacks.cpp:494-589
void xpack_svd_dense_arr(const DenseMatrix *arr, int arr_size, DenseMatrix &lsvects,
                         Vector &svals)
{
    const lapack_t &L = lapack();
    char jobu = 'S';
    char jobvt = 'N';
    int m = arr[0].Height();
    int n = arr[0].Width();
    for (int i = 1; i < arr_size; ++i)
    {
        SA_ASSERT(arr[i].Height() == m);
        n += arr[i].Width();
    }
    int lda = m, ldu = m, ldvt = n;
    int lwork = -1, info;
    double qwork;
    std::vector<double> a((size_t)m * n);
    int minimal = std::min(m, n);
    SA_ASSERT(minimal > 0);

    double *ptr = a.data();
    for (int i = 0; i < arr_size; ++i)
        for (int j = 0; j < arr[i].Width(); ++j)
        {
            const double *col = arr[i].Data() + (size_t)j * m;
            double norm = 0.;
            for (int r = 0; r < m; ++r)
                norm += col[r] * col[r];
            norm = sqrt(norm); // Vector::Norml2
            if (SA_REAL_ALMOST_LE(norm, 0.))
                n = n - 1;
            else
            {
                for (int r = 0; r < m; ++r)
                    ptr[r] = col[r] / norm;
                ptr += m;
            }
        }
    minimal = std::min(m, n);
    svals.assign(minimal, 0.);
    lsvects.SetSize(m, minimal);
    if (minimal <= 0)
        return;
    double *s = svals.data();
    double *u = lsvects.Data();
    double vt_dummy = 0.;
    ldvt = std::max(1, n);
    L.dgesvd(&jobu, &jobvt, &m, &n, a.data(), &lda, s, u, &ldu, &vt_dummy, &ldvt, &qwork,
             &lwork, &info);
    SA_ASSERT(!info);
    lwork = (int)qwork + 1;
    if (lwork < std::max(3 * minimal + std::max(m, n), 5 * minimal))
        lwork = std::max(3 * minimal + std::max(m, n), 5 * minimal);
    std::vector<double> work(lwork);
    L.dgesvd(&jobu, &jobvt, &m, &n, a.data(), &lda, s, u, &ldu, &vt_dummy, &ldvt,
             work.data(), &lwork, &info);
    SA_ASSERT(!info);
}

// amg/src/xpacks.cpp:591-620
void xpack_orth_set(const DenseMatrix &lsvects, const Vector &svals,
                    DenseMatrix &orth_set, double eps)
{
    const int h = lsvects.Height();
    int i;
    SA_ASSERT(svals.size());
    eps *= svals[0];
    for (i = 0; i < (int)svals.size() && svals[i] > eps; ++i)
        ;
    SA_ASSERT(i);
    orth_set.SetSize(h, i);
    std::memcpy(orth_set.Data(), lsvects.Data(), sizeof(double) * i * h);
}

// amg/src/xpacks.cpp:627-655
void xpack_solve_lls(const DenseMatrix &A, const Vector &rhs, Vector &x)
{
    const lapack_t &L = lapack();
    char trans = 'N';
    int m = A.Height(), n = A.Width(), nrhs = 1;
    std::vector<double> a(A.d), b(rhs);
    int lda = m, ldb = m, info, lwork = 2 * (m + n);
    std::vector<double> work(lwork);
    L.dgels(&trans, &m, &n, &nrhs, a.data(), &lda, b.data(), &ldb, work.data(), &lwork,
            &info);
    SA_ASSERT(!info);
    x.assign(b.begin(), b.begin() + n);
}

// amg/src/mbox.cpp:1839-1861
Vector *mbox_build_Dinv_neg_parallel_matrix(const SparseMatrix &A)
{
    const int n = A.Size();
    Vector diag1(n), *diag2 = new Vector(n);
    for (int i = 0; i < n; ++i)
    {
        const double d = fabs(A(i, i));
        diag1[i] = 1. / sqrt(d);
        (*diag2)[i] = sqrt(d);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
    {
        double y = 0.;
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            y += fabs(A.A[p]) * diag1[A.J[p]];
        (*diag2)[i] = -1. / ((*diag2)[i] * y);
    }
    return diag2;
}

// amg/src/smpr.cpp:266-280
double *smpr_sa_poly_roots(int &nu, int *degree)
{
    SA_ASSERT(nu >= 0);
    const double denom = (double)(2 * nu + 1);
    double *roots = new double[std::max(1, *degree = nu)];
    for (int i = 1; i <= nu; ++i)
    {
        const double sin_val = sin(((double)i * M_PI) / denom);
        roots[i - 1] = sin_val * sin_val;
    }
    return roots;
}

// amg/src/smpr.cpp:282-306
double *smpr_sas_poly_roots(int &nu, int *degree)
{
    SA_ASSERT(nu > 0);
    const int twonu = 2 * nu;
    const double denom = (double)(2 * nu + 1);
    double val;
    int i;
    double *roots = new double[*degree = twonu + nu + 1];
    for (i = 0; i <= twonu; ++i)
    {
        val = cos(((double)i * M_PI) / denom);
        roots[i] = val * val;
    }
    for (i = 1; i <= nu; ++i)
    {
        val = sin(((double)i * M_PI) / denom);
        roots[i + twonu] = val * val;
    }
    return roots;
}

/* ----------------------------------------------------------------- providers */

ElementMatrixStandardGeometric::ElementMatrixStandardGeometric(
    const agg_partitioning_relations_t &rels, const SparseMatrix &A, const double *elmats,
    int ne)
    : ElementMatrixProvider(rels), A_(A), elmats_(elmats), ne_(ne)
{
    is_geometric = true;
}

// amg/src/elmat.cpp:57-62 (bdr_cond_imposed_ = assemble_ess_diag_ = true, :51-52)
SparseMatrix *ElementMatrixStandardGeometric::BuildAEStiff(int elno) const
{
    return agg_build_AE_stiffm_with_global(A_, elno, agg_part_rels, this, true, true);
}

// amg/src/elmat.cpp:68-88: a fresh dense element matrix per call, caller frees
Matrix *ElementMatrixStandardGeometric::GetMatrix(int elno, bool &free_matr) const
{
    DenseMatrix *elmat = new DenseMatrix(ne_, ne_);
    std::memcpy(elmat->Data(), elmats_ + (size_t)elno * ne_ * ne_,
                sizeof(double) * ne_ * ne_);
    free_matr = true;
    return elmat;
}

ElementMatrixParallelCoarse::ElementMatrixParallelCoarse(
    const agg_partitioning_relations_t &rels, const oracle_level_t *finer)
    : ElementMatrixProvider(rels), finer_(finer)
{
    is_geometric = false;
}

SparseMatrix *ElementMatrixParallelCoarse::BuildAEStiff(int elno) const
{
    return agg_build_AE_stiffm(elno, agg_part_rels, this);
}

// amg/src/elmat.cpp:105-195
Matrix *ElementMatrixParallelCoarse::GetMatrix(int elno, bool &free_matr) const
{
    const agg_partitioning_relations_t &frels = *finer_->agg_part_rels;
    const Table *AE_to_mis = frels.AE_to_mis;
    const Table *mis_to_dof = frels.mis_to_dof;
    const std::vector<int> &mis_numcoarsedof = finer_->mis_numcoarsedof;
    const SparseMatrix *finer_AE_stiffm = finer_->AEs_stiffm[elno];
    const int ae_finedof = finer_AE_stiffm->Size();
    int ae_coarsedof = 0;
    std::vector<int> mis_in_AE(AE_to_mis->GetRow(elno),
                               AE_to_mis->GetRow(elno) + AE_to_mis->RowSize(elno));
    std::sort(mis_in_AE.begin(), mis_in_AE.end());
    for (size_t j = 0; j < mis_in_AE.size(); ++j)
        ae_coarsedof += mis_numcoarsedof[mis_in_AE[j]];

    DenseMatrix local_interp(ae_finedof, ae_coarsedof);
    for (size_t j = 0; j < mis_in_AE.size(); ++j)
    {
        const int mis = mis_in_AE[j];
        const int num_finedof_in_mis = mis_to_dof->RowSize(mis);
        const int *finedof_in_mis = mis_to_dof->GetRow(mis);
        const DenseMatrix &Tm = *finer_->mis_tent_interps[mis];
        for (int c = 0; c < mis_numcoarsedof[mis]; ++c)
        {
            const int coarse_dof_num = agg_part_rels.mis_coarsedofoffsets[mis] + c;
            const int column_to_put =
                agg_elem_in_col(elno, coarse_dof_num, *agg_part_rels.elem_to_dof);
            SA_ASSERT(column_to_put >= 0 && column_to_put < ae_coarsedof);
            for (int i = 0; i < num_finedof_in_mis; ++i)
            {
                const int dof_in_AE = agg_map_id_glob_to_AE(finedof_in_mis[i], elno, frels);
                SA_ASSERT(dof_in_AE >= 0 && dof_in_AE < ae_finedof);
                local_interp(dof_in_AE, column_to_put) += Tm(i, c); // AddSubMatrix
            }
        }
    }
    // RAP(A, R) with R = local_interp^T: out = P^T (A P)
    DenseMatrix AP(ae_finedof, ae_coarsedof);
    for (int i = 0; i < ae_finedof; ++i)
        for (int p = finer_AE_stiffm->I[i]; p < finer_AE_stiffm->I[i + 1]; ++p)
        {
            const double a = finer_AE_stiffm->A[p];
            const int k = finer_AE_stiffm->J[p];
            for (int c = 0; c < ae_coarsedof; ++c)
                AP(i, c) += a * local_interp(k, c);
        }
    DenseMatrix *out = new DenseMatrix(ae_coarsedof, ae_coarsedof);
    for (int c2 = 0; c2 < ae_coarsedof; ++c2)
        for (int c1 = 0; c1 < ae_coarsedof; ++c1)
        {
            double s = 0.;
            for (int i = 0; i < ae_finedof; ++i)
                s += local_interp(i, c1) * AP(i, c2);
            (*out)(c1, c2) = s;
        }
    free_matr = true;
    return out;
}

/* -------------------------------------------------------------- level pieces */

oracle_level_t::~oracle_level_t()
{
    for (size_t i = 0; i < AEs_stiffm.size(); ++i)
    {
        delete AEs_stiffm[i];
        delete rhs_matrices_arr[i];
        delete cut_evects_arr[i];
    }
    for (size_t i = 0; i < mis_tent_interps.size(); ++i)
        delete mis_tent_interps[i];
    delete[] interp_smoother_roots;
    delete ltent_interp;
    delete interp;
    delete restr;
    delete Ac;
    delete[] roots;
    delete Dinv_neg;
    delete elem_data;
    if (owns_A)
        delete A;
}

oracle_ml_t::~oracle_ml_t()
{
    // a level's A is the finer level's Ac: free coarse-to-fine, clearing the alias
    for (size_t l = 0; l < levels.size(); ++l)
        levels[l]->owns_A = false;
    for (size_t l = 0; l < levels.size(); ++l)
        delete levels[l];
}

// Eigensolver::SolveDirect, amg/src/spectral.cpp:124-237 (transf = all_eigens = false)
static void eigensolver_solve_direct(const SparseMatrix &A, SparseMatrix *&B, double theta,
                                     DenseMatrix &cut_evects, Vector &evals)
{
    const double lmax = 1.;
    DenseMatrix deA, deB;
    if (!B)
        B = mbox_snd_D_sparse_from_sparse(A);
    mbox_convert_sparse_to_dense(A, deA);
    mbox_convert_sparse_to_dense(*B, deB);
    xpacks_calc_lower_eigens_dense(deA, evals, cut_evects, deB, theta * lmax, true);
}

// amg/src/interp.cpp:342-593 (build from scratch: transf = readapting = false)
void interp_compute_vectors(const agg_partitioning_relations_t &agg_part_rels,
                            oracle_level_t &lev, double &theta, bool bdr_cond_imposed)
{
    (void)bdr_cond_imposed;
    const int nparts = agg_part_rels.nparts;
    lev.AEs_stiffm.assign(nparts, NULL);
    lev.rhs_matrices_arr.assign(nparts, NULL);
    lev.cut_evects_arr.assign(nparts, NULL);
    lev.evals_arr.assign(nparts, Vector());
    double sum_skip = 0.;
    int skipctr = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : sum_skip, skipctr)
    for (int i = 0; i < nparts; ++i)
    {
        double theta_local = theta;
        lev.AEs_stiffm[i] = lev.elem_data->BuildAEStiff(i);
        lev.cut_evects_arr[i] = new DenseMatrix;
        eigensolver_solve_direct(*lev.AEs_stiffm[i], lev.rhs_matrices_arr[i], theta_local,
                                 *lev.cut_evects_arr[i], lev.evals_arr[i]);
        // mltest fixture: extra all-ones vector on AE 0 (amg/src/interp.cpp:510-524)
        if (lev.testmesh_inject && i == 0)
        {
            DenseMatrix *old = lev.cut_evects_arr[i];
            const int h = old->Height(), w = old->Width() + 1;
            DenseMatrix *nw = new DenseMatrix(h, w);
            std::memcpy(nw->Data(), old->Data(), sizeof(double) * h * (w - 1));
            for (int j = 0; j < h; ++j)
                (*nw)(j, w - 1) = 1.0;
            delete old;
            lev.cut_evects_arr[i] = nw;
        }
        sum_skip += theta_local;
        ++skipctr;
    }
    // amg/src/interp.cpp:571-589
    const double thetap = sum_skip / (double)skipctr;
    const double eta = 0.5;
    if (skipctr > 0)
        theta = (1. - eta) * theta + eta * thetap;
}

// ContribTent::contrib_filter_boundary, amg/src/contrib.cpp:102-163
static void contrib_filter_boundary(const agg_partitioning_relations_t &agg_part_rels,
                                    DenseMatrix &local, const int *restriction,
                                    bool avoid_ess_bdr_dofs)
{
    const int vects = local.Width();
    const int dim = local.Height();
    std::vector<double> newdata((size_t)vects * dim);
    int col = 0;
    for (int i = 0; i < vects; ++i)
    {
        bool atleastone = false;
        double *dst = &newdata[(size_t)col * dim];
        for (int j = 0; j < dim; ++j)
        {
            const int row = restriction[j];
            const double a = local(j, i);
            if (a == 0.0 ||
                (avoid_ess_bdr_dofs && agg_is_dof_on_essential_border(agg_part_rels, row)))
            {
                dst[j] = 0.0;
                continue;
            }
            atleastone = true;
            dst[j] = a;
        }
        if (atleastone)
            col++;
    }
    local.SetSize(dim, col);
    std::memcpy(local.Data(), newdata.data(), sizeof(double) * dim * col);
}

// contrib_mises -> CommunicateEigenvectors + SVDInsert + contrib_tent_insert_simple
// + contrib_tent_finalize, on one process (amg/src/contrib.cpp:492-714, 73-95, 170-194)
SparseMatrix *interp_sparse_tent_assemble(const agg_partitioning_relations_t &agg_part_rels,
                                          oracle_level_t &lev, bool avoid_ess_bdr_dofs)
{
    const int num_mises = agg_part_rels.num_mises;
    const double svd_eps = 1.e-10; // amg/src/contrib.cpp:61
    lev.mis_tent_interps.assign(num_mises, NULL);
    lev.mis_numcoarsedof.assign(num_mises, 0);

#pragma omp parallel for schedule(dynamic, 16)
    for (int mis = 0; mis < num_mises; ++mis)
    {
        // CommunicateEigenvectors: restrict every AE's vectors to the MIS and
        // concatenate column-wise in mis_to_AE order (amg/src/contrib.cpp:501-546)
        const int mis_size = agg_part_rels.mises_size[mis];
        const int rowsize = agg_part_rels.mis_to_AE->RowSize(mis);
        const int *row = agg_part_rels.mis_to_AE->GetRow(mis);
        const int *mis_dofs = agg_part_rels.mis_to_dof->GetRow(mis);
        int numvecs = 0;
        for (int j = 0; j < rowsize; ++j)
            numvecs += lev.cut_evects_arr[row[j]]->Width();
        DenseMatrix received(mis_size, numvecs);
        numvecs = 0;
        for (int j = 0; j < rowsize; ++j)
        {
            const int AE = row[j];
            DenseMatrix restricted;
            agg_restrict_to_agg_enforce(AE, agg_part_rels, mis_size, mis_dofs,
                                        *lev.cut_evects_arr[AE], restricted);
            std::memcpy(received.Data() + (size_t)numvecs * mis_size, restricted.Data(),
                        sizeof(double) * mis_size * restricted.Width());
            numvecs += restricted.Width();
        }

        // SVDInsert body for an owned MIS (amg/src/contrib.cpp:564-672)
        DenseMatrix *tent = new DenseMatrix;
        lev.mis_tent_interps[mis] = tent;
        const int dim = mis_size;
        if (avoid_ess_bdr_dofs)
        {
            bool interior_dofs = false;
            for (int j = 0; j < dim; ++j)
                if (!agg_is_dof_on_essential_border(agg_part_rels, mis_dofs[j]))
                {
                    interior_dofs = true;
                    break;
                }
            if (!interior_dofs)
            {
                tent->SetSize(dim, 0);
                continue;
            }
        }
        if (dim == 1)
        {
            tent->SetSize(1, 1);
            (*tent)(0, 0) = 1.0;
        }
        else
        {
            DenseMatrix lsvects;
            Vector svals;
            contrib_filter_boundary(agg_part_rels, received, mis_dofs, avoid_ess_bdr_dofs);
            if (received.Width() == 0)
                svals.clear();
            else
                xpack_svd_dense_arr(&received, 1, lsvects, svals);
            if (svals.size() == 0)
            {
                tent->SetSize(dim, 0);
                continue;
            }
            xpack_orth_set(lsvects, svals, *tent, svd_eps);
        }
    }

    // contrib_tent_insert_simple in MIS order (amg/src/contrib.cpp:170-194, 646-670)
    LilMatrix tent_interp(agg_part_rels.ND, 0);
    int filled_cols = 0;
    for (int mis = 0; mis < num_mises; ++mis)
    {
        const DenseMatrix &local = *lev.mis_tent_interps[mis];
        const int *restriction = agg_part_rels.mis_to_dof->GetRow(mis);
        const int vects = local.Width();
        const int dim = local.Height();
        int col = filled_cols;
        for (int i = 0; i < vects; ++i)
        {
            for (int j = 0; j < dim; ++j)
                if (fabs(local(j, i)) > 0.0) // threshold_ = 0.0
                    tent_interp.Set(restriction[j], col, local(j, i));
            ++col;
        }
        lev.mis_numcoarsedof[mis] = col - filled_cols;
        filled_cols = col;
    }
    tent_interp.w = filled_cols;
    return tent_interp.Finalize();
}

// AltThresholdLocal (amg/src/interp.cpp:86-126): keeps the entries with fabs(val) > threshold
static SparseMatrix *AltThreshold(const SparseMatrix &mat, double val)
{
    SparseMatrix *out = new SparseMatrix;
    out->h = mat.h;
    out->w = mat.w;
    out->I.assign((size_t)mat.h + 1, 0);
    for (int i = 0; i < mat.h; ++i)
    {
        int thisrowcount = 0;
        for (int jp = mat.I[i]; jp < mat.I[i + 1]; ++jp)
            if (fabs(mat.A[jp]) > val)
            {
                ++thisrowcount;
                out->J.push_back(mat.J[jp]);
                out->A.push_back(mat.A[jp]);
            }
        out->I[i + 1] = out->I[i] + thisrowcount;
    }
    return out;
}

// amg/src/interp.cpp:64-82, 172-229 (times_apply_smoother = 1)
SparseMatrix *interp_smooth(int degree, const double *roots, const SparseMatrix &A,
                            const SparseMatrix &tent, const Vector &Dinv_neg, double drop_tol)
{
    // smoother_matr = diag(Dinv_neg) * A
    SparseMatrix smoother_matr = A;
    for (int i = 0; i < A.h; ++i)
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            smoother_matr.A[p] = Dinv_neg[i] * A.A[p];
    SparseMatrix *interp = new SparseMatrix(tent);
    for (int k = 0; k < degree; ++k)
    {
        SparseMatrix iter_matr = smoother_matr;
        const double scale = 1. / roots[k];
        for (size_t p = 0; p < iter_matr.A.size(); ++p)
            iter_matr.A[p] *= scale; // mbox_scale_clone_parallel_matrix
        for (int i = 0; i < iter_matr.h; ++i) // mbox_add_diag_parallel_matrix(., 1.)
            for (int p = iter_matr.I[i]; p < iter_matr.I[i + 1]; ++p)
                if (iter_matr.J[p] == i)
                    iter_matr.A[p] += 1.;
        SparseMatrix *new_interp = new SparseMatrix;
        SpMultMat(iter_matr, *interp, *new_interp); // ParMult
        delete interp;
        interp = new_interp;
    }
    if (drop_tol == 0.0)
        return interp;
    SparseMatrix *new_interp = AltThreshold(*interp, drop_tol);
    delete interp;
    return new_interp;
}

/* ---------------------------------------------------------------------- solve */

// amg/inc/smpr.hpp:319-339
void smpr_compute_poly(const SparseMatrix &A, const Vector &b, Vector &x, int degree,
                       const double *roots, const Vector &Dinv_neg)
{
    const int n = (int)b.size();
    Vector tmp(n), tmp1(n);
    for (int i = 0; i < degree; ++i)
    {
        const double mult = 1. / roots[i];
#pragma omp parallel for schedule(static)
        for (int r = 0; r < n; ++r)
        {
            double s = 0.;
            for (int p = A.I[r]; p < A.I[r + 1]; ++p)
                s += A.A[p] * x[A.J[p]];
            tmp1[r] = s;
        }
#pragma omp parallel for schedule(static)
        for (int r = 0; r < n; ++r)
        {
            double t = -1. * b[r]; // tmp.Set(-1., b)
            t += tmp1[r];          // tmp += tmp1
            t *= Dinv_neg[r];      // mbox_entry_mult_vector
            x[r] += mult * t;      // x.Add(mult, tmp)
        }
    }
}

static void par_spmult(const SparseMatrix &A, const Vector &x, Vector &y)
{
    const int n = A.h;
    y.resize(n);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r)
    {
        double s = 0.;
        for (int p = A.I[r]; p < A.I[r + 1]; ++p)
            s += A.A[p] * x[A.J[p]];
        y[r] = s;
    }
}

// VCycleSolver::Mult (amg/src/solve.cpp:309-323) + tg_cycle_atb (amg/src/tg.cpp:91-132)
void vcycle_mult(const oracle_ml_t &ml, int level, const Vector &b, Vector &x)
{
    const oracle_level_t &L = *ml.levels[level];
    const SparseMatrix &A = *L.A;
    x.assign(b.size(), 0.0); // iterative_mode == false
    Vector res(b.size()), resc, xc;
    // pre_smoother = smpr_sym_poly (roots2 == NULL)
    smpr_compute_poly(A, b, x, L.degree, L.roots, *L.Dinv_neg);
    par_spmult(A, x, res);
    for (size_t i = 0; i < b.size(); ++i)
        res[i] = b[i] - res[i];
    par_spmult(*L.restr, res, resc);
    xc.assign(resc.size(), 0.0);
    if (level + 1 < (int)ml.levels.size())
        vcycle_mult(ml, level + 1, resc, xc); // ml_impose_cycle, amg/src/ml.cpp:361-377
    else
    {
        // exact coarsest solve (HypreDirect/UMFPACK in the reference,
        // amg/src/tg.cpp:991-998); here dense Cholesky
        const lapack_t &LP = lapack();
        char uplo = 'L';
        int n = (int)resc.size(), nrhs = 1, info;
        xc = resc;
        LP.dpotrs(&uplo, &n, &nrhs, const_cast<double *>(L.Ac_chol.data()), &n, xc.data(),
                  &n, &info);
        SA_ASSERT(!info);
    }
    Vector Pxc;
    par_spmult(*L.interp, xc, Pxc);
    for (size_t i = 0; i < x.size(); ++i)
        x[i] += Pxc[i];
    smpr_compute_poly(A, b, x, L.degree, L.roots, *L.Dinv_neg);
}

static double inner(const Vector &a, const Vector &b)
{
    double s = 0.;
    for (size_t i = 0; i < a.size(); ++i)
        s += a[i] * b[i];
    return s;
}

// amg/src/mfem_addons.cpp:106-248 (zero_rhs = false)
int kalchev_pcg(const SparseMatrix &A, const oracle_ml_t &ml, const Vector &b, Vector &x,
                int max_num_iter, double RTOLERANCE, double ATOLERANCE,
                std::vector<double> *brr)
{
    const int dim = (int)x.size();
    int i, iters = 0;
    double r0, den, nom, betanom = 0., alpha, beta;
    Vector r(dim), d(dim), z(dim);

    par_spmult(A, x, r);
    for (int k = 0; k < dim; ++k)
        r[k] = b[k] - r[k];
    vcycle_mult(ml, 0, r, z);
    d = z;
    nom = inner(z, r);
    if (brr)
        brr->push_back(nom);
    if ((r0 = nom * RTOLERANCE) < ATOLERANCE)
        r0 = ATOLERANCE;
    if (nom < r0)
        return -1;
    par_spmult(A, d, z);
    den = inner(z, d);
    if (0. == den)
        return -1;
    for (i = 1; i <= max_num_iter; i++)
    {
        alpha = nom / den;
        for (int k = 0; k < dim; ++k)
        {
            x[k] = x[k] + alpha * d[k];
            r[k] = r[k] - alpha * z[k];
        }
        vcycle_mult(ml, 0, r, z);
        betanom = inner(r, z);
        if (brr)
            brr->push_back(betanom);
        if (betanom < 0.0)
        {
            iters = -i;
            break;
        }
        if (betanom < r0)
        {
            iters = i;
            break;
        }
        beta = betanom / nom;
        for (int k = 0; k < dim; ++k)
            d[k] = z[k] + beta * d[k];
        par_spmult(A, d, z);
        den = inner(d, z);
        nom = betanom;
    }
    if (i > max_num_iter)
        iters = -(i - 1);
    return iters;
}

/* ------------------------------------------------------------- slow MIS scan */

// amg/src/aggregates.cpp:541-607 on one process: for every not yet distributed
// dof i, count for every dof k the AEs shared with i; k joins the MIS of i iff
// count[k] == rowsum[k] == rowsum[i].
void agg_construct_mises_local_scan(const Table &dof_to_AE, std::vector<int> &mises,
                                    Table &mis_to_dof)
{
    const int nd = dof_to_AE.Size();
    std::vector<int> count(nd), distributed(nd, 0);
    mises.assign(nd, -1);
    std::vector<std::vector<int>> rows;
    for (int i = 0; i < nd; ++i)
    {
        if (distributed[i])
            continue;
        std::fill(count.begin(), count.end(), 0);
        for (int j = dof_to_AE.I[i]; j < dof_to_AE.I[i + 1]; ++j)
            for (int k = 0; k < nd; ++k)
                for (int kj = dof_to_AE.I[k]; kj < dof_to_AE.I[k + 1]; ++kj)
                    if (dof_to_AE.J[kj] == dof_to_AE.J[j])
                        count[k]++;
        std::vector<int> newrow;
        for (int k = 0; k < nd; ++k)
            if (count[k] == dof_to_AE.RowSize(k) && count[k] == dof_to_AE.RowSize(i))
                newrow.push_back(k);
        for (size_t k = 0; k < newrow.size(); ++k)
        {
            distributed[newrow[k]] = 1;
            mises[newrow[k]] = (int)rows.size();
        }
        rows.push_back(newrow);
    }
    mis_to_dof.nrows = (int)rows.size();
    mis_to_dof.ncols = nd;
    mis_to_dof.I.assign(rows.size() + 1, 0);
    mis_to_dof.J.clear();
    for (size_t m = 0; m < rows.size(); ++m)
    {
        mis_to_dof.J.insert(mis_to_dof.J.end(), rows[m].begin(), rows[m].end());
        mis_to_dof.I[m + 1] = (int)mis_to_dof.J.size();
    }
}

/* ------------------------------------------------------- hierarchy (ml_/tg_) */

static void export_level(const oracle_level_t &L, sa_level_results_t &R,
                         const ElementMatrixProvider *coarse_provider)
{
    const agg_partitioning_relations_t &rels = *L.agg_part_rels;
    R.nparts = rels.nparts;
    R.num_mises = rels.num_mises;
    R.ND = rels.ND;
    R.ae_m.resize(rels.nparts);
    R.ae_eval_off.assign((size_t)rels.nparts + 1, 0);
    R.ae_evect_off.assign((size_t)rels.nparts + 1, 0);
    R.ae_D.clear();
    for (int i = 0; i < rels.nparts; ++i)
    {
        const int m = L.cut_evects_arr[i]->Width();
        const int n = L.cut_evects_arr[i]->Height();
        R.ae_m[i] = m;
        R.ae_eval_off[i + 1] = R.ae_eval_off[i] + (int64_t)L.evals_arr[i].size();
        R.ae_evect_off[i + 1] = R.ae_evect_off[i] + (int64_t)n * m;
        R.evals.insert(R.evals.end(), L.evals_arr[i].begin(), L.evals_arr[i].end());
        R.evects.insert(R.evects.end(), L.cut_evects_arr[i]->d.begin(),
                        L.cut_evects_arr[i]->d.end());
        R.ae_D.insert(R.ae_D.end(), L.rhs_matrices_arr[i]->A.begin(),
                      L.rhs_matrices_arr[i]->A.end());
    }
    R.mis_numcoarsedof = L.mis_numcoarsedof;
    R.mis_off.assign((size_t)rels.num_mises + 1, 0);
    for (int mis = 0; mis < rels.num_mises; ++mis)
    {
        const DenseMatrix &T = *L.mis_tent_interps[mis];
        R.mis_off[mis + 1] = R.mis_off[mis] + (int64_t)T.h * T.w;
        R.mis_tent.insert(R.mis_tent.end(), T.d.begin(), T.d.end());
    }
    R.tent_interp = *L.ltent_interp;
    R.interp = *L.interp;
    R.Ac = *L.Ac;
    R.Dinv_neg = *L.Dinv_neg;
    R.NDc = L.Ac->h;
    if (coarse_provider)
    {
        R.celmat_off.assign((size_t)rels.nparts + 1, 0);
        for (int e = 0; e < rels.nparts; ++e)
        {
            bool fr;
            DenseMatrix *M = static_cast<DenseMatrix *>(coarse_provider->GetMatrix(e, fr));
            R.celmat_off[e + 1] = R.celmat_off[e] + (int64_t)M->h * M->w;
            R.celmat.insert(R.celmat.end(), M->d.begin(), M->d.end());
            delete M;
        }
    }
}

// tg_init_data + tg_build_hierarchy + tg_update_coarse_operator
// (amg/src/tg.cpp:402-430, 502-540, 432-473, 979-1014; amg/inc/tg.hpp:678-709)
static void tg_build_level(oracle_level_t &L, int nu_pro, int nu_relax, double theta,
                           bool avoid_ess_bdr_dofs, std::map<std::string, double> &times,
                           int levelno)
{
    const agg_partitioning_relations_t &rels = *L.agg_part_rels;
    char key[64];
    double t0 = now_s();
    // tg_init_data
    L.theta = theta;
    L.nu_pro = nu_pro;
    L.interp_smoother_roots = smpr_sa_poly_roots(L.nu_pro, &L.interp_smoother_degree);
    L.nu = nu_relax;
    L.Dinv_neg = mbox_build_Dinv_neg_parallel_matrix(*L.A);
    L.roots = smpr_sas_poly_roots(L.nu, &L.degree);
    std::snprintf(key, sizeof key, "l%d.init", levelno);
    times[key] = now_s() - t0;
    // tg_build_hierarchy -> interp_sparse_tent_build
    t0 = now_s();
    interp_compute_vectors(rels, L, L.theta, avoid_ess_bdr_dofs);
    std::snprintf(key, sizeof key, "l%d.local_spectral", levelno);
    times[key] = now_s() - t0;
    t0 = now_s();
    L.ltent_interp = interp_sparse_tent_assemble(rels, L, avoid_ess_bdr_dofs);
    std::snprintf(key, sizeof key, "l%d.tentative", levelno);
    times[key] = now_s() - t0;
    // tg_assemble_and_smooth: interp_global_tent_assemble is the identity on one process
    t0 = now_s();
    if (nu_pro > 0)
        L.interp = interp_smooth(L.interp_smoother_degree, L.interp_smoother_roots, *L.A,
                                 *L.ltent_interp, *L.Dinv_neg, L.drop_tol);
    else
        L.interp = new SparseMatrix(*L.ltent_interp);
    L.restr = new SparseMatrix;
    SpTranspose(*L.interp, *L.restr);
    std::snprintf(key, sizeof key, "l%d.smooth_P", levelno);
    times[key] = now_s() - t0;
    // tg_update_coarse_operator: Ac = RAP(A, interp)
    t0 = now_s();
    SparseMatrix AP;
    SpMultMat(*L.A, *L.interp, AP);
    L.Ac = new SparseMatrix;
    SpMultMat(*L.restr, AP, *L.Ac);
    std::snprintf(key, sizeof key, "l%d.rap", levelno);
    times[key] = now_s() - t0;
}

/* CorrectNullspace (amg/src/solve.cpp:52-110) below the last spectral level: scaling P from
   local_coarse_one_representation (amg/src/contrib.cpp:655-668: per MIS with coarse dofs,
   xpack_solve_lls of mis_tent_interps[mis] x = 1, normalised) assembled as in
   interp_scaling_P_assemble (amg/src/interp.cpp:842-909); Ac = RAP(A, scaling P); smoother
   smpr_init_poly_data(A, 3, 0.0) (SAS, nu = 3); its Mult is one tg_cycle_atb, i.e. one more level
   of vcycle_mult.  The reference solves the level below with BoomerAMG; here exactly. */
static oracle_level_t *build_correct_nullspace(oracle_level_t &last)
{
    oracle_level_t *N = new oracle_level_t;
    N->A = last.Ac;
    N->owns_A = false;
    SparseMatrix *SP = new SparseMatrix;
    SP->I.push_back(0);
    int rows = 0, col = 0;
    for (size_t mis = 0; mis < last.mis_tent_interps.size(); ++mis)
    {
        const DenseMatrix &V = *last.mis_tent_interps[mis];
        if (last.mis_numcoarsedof[mis] <= 0)
            continue;
        SA_ASSERT(V.Width() == last.mis_numcoarsedof[mis]);
        Vector b((size_t)V.Height(), 1.0), x;
        xpack_solve_lls(V, b, x);
        double norm = 0.;
        for (size_t k = 0; k < x.size(); ++k)
            norm += x[k] * x[k];
        norm = sqrt(norm);
        for (size_t k = 0; k < x.size(); ++k)
        {
            SP->J.push_back(col);
            SP->A.push_back(x[k] / norm);
            SP->I.push_back((int)SP->J.size());
            ++rows;
        }
        ++col;
    }
    SP->h = rows;
    SP->w = col;
    SA_ASSERT(rows == N->A->h);
    N->interp = SP;
    N->restr = new SparseMatrix;
    SpTranspose(*N->interp, *N->restr);
    N->nu = 3;
    N->Dinv_neg = mbox_build_Dinv_neg_parallel_matrix(*N->A);
    N->roots = smpr_sas_poly_roots(N->nu, &N->degree);
    SparseMatrix AP;
    SpMultMat(*N->A, *N->interp, AP);
    N->Ac = new SparseMatrix;
    SpMultMat(*N->restr, AP, *N->Ac);
    return N;
}

static void factor_coarsest(oracle_level_t &L)
{
    const lapack_t &LP = lapack();
    const int n = L.Ac->h;
    L.Ac_chol.assign((size_t)n * n, 0.);
    for (int i = 0; i < n; ++i)
        for (int p = L.Ac->I[i]; p < L.Ac->I[i + 1]; ++p)
            L.Ac_chol[(size_t)L.Ac->J[p] * n + i] = L.Ac->A[p];
    char uplo = 'L';
    int nn = n, info;
    LP.dpotrf(&uplo, &nn, L.Ac_chol.data(), &nn, &info);
    SA_ASSERT(!info);
}

} // namespace saamge_oracle

using namespace saamge_oracle;

static int g_threads = 1;

extern "C" void sa_orc_init(const char *lapack_path, int num_threads)
{
    lapack(lapack_path && lapack_path[0] ? lapack_path : NULL);
    if (num_threads > 0)
    {
        omp_set_num_threads(num_threads);
        g_threads = num_threads;
    }
    else
        g_threads = omp_get_max_threads();
}

extern "C" int sa_orc_num_threads(void) { return g_threads; }

static void oracle_impl_free(void *p) { delete (oracle_ml_t *)p; }

// ml_produce_data + ml_produce_hierarchy_from_level (amg/src/ml.cpp:379-472, 111-236)
extern "C" void *sa_orc_ml_build(void *prob_, const sa_drv_params_t *p)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    sa_hierarchy_t *H = new sa_hierarchy_t;
    H->prob = prob;
    H->params = *p;
    oracle_ml_t *ml = new oracle_ml_t;
    H->impl = ml;
    H->impl_free = oracle_impl_free;
    const int coarsenings = p->num_levels - 1;
    std::vector<int> nparts_arr = sa_target_nparts(prob->fem->NE, *p);
    const bool avoid = p->avoid_ess_bdr_dofs != 0;
    const double tstart = now_s();

    H->rels.push_back(prob->rels);
    for (int i = 0; i < coarsenings; ++i)
    {
        oracle_level_t *L = new oracle_level_t;
        ml->levels.push_back(L);
        if (i == 0)
        {
            L->agg_part_rels = prob->rels;
            L->A = &prob->fem->A;
            L->elem_data = new ElementMatrixStandardGeometric(
                *prob->rels, prob->fem->A, prob->fem->elmat.data(), prob->fem->ne);
            L->testmesh_inject = p->testmesh_inject != 0;
        }
        else
        {
            oracle_level_t *F = ml->levels[i - 1];
            const double t0 = now_s();
            int nparts = nparts_arr[i];
            int *partitioning = sa_prescribed_coarse_partitioning(
                *prob, *p, i, F->agg_part_rels->nparts, &nparts);
            // the oracle's own restatement (orc_topology.cpp), not the product's construction
            agg_partitioning_relations_t *rels = orc_create_partitioning_coarse(
                *F->agg_part_rels, *F->ltent_interp, F->mis_numcoarsedof.data(), &nparts, partitioning);
            H->rels.push_back(rels);
            char key[64];
            std::snprintf(key, sizeof key, "l%d.topology", i);
            H->times[key] = now_s() - t0;
            L->agg_part_rels = rels;
            L->A = F->Ac;
            L->elem_data = new ElementMatrixParallelCoarse(*rels, F);
        }
        L->drop_tol = p->smooth_drop_tol;
        tg_build_level(*L, i == 0 ? p->first_nu_pro : p->nu_pro, p->nu_relax,
                       i == 0 ? p->first_theta : p->theta, avoid, H->times, i);
    }
    if (p->correct_nullspace)
    {
        ml->levels.push_back(build_correct_nullspace(*ml->levels.back()));
        H->cn_P = *ml->levels.back()->interp;
        H->cn_Ac = *ml->levels.back()->Ac;
    }
    factor_coarsest(*ml->levels.back());
    H->times["setup"] = now_s() - tstart;

    H->levels.resize(coarsenings);
    for (int i = 0; i < coarsenings; ++i)
        export_level(*ml->levels[i], H->levels[i],
                     i + 1 < coarsenings ? ml->levels[i + 1]->elem_data : NULL);
    return H;
}

/* ---- algebraic entry: the oracle's own restatement of ExtractSubMatrices and of the
   ElementMatrixArray provider (amg/src/tg.cpp:580-668, amg/src/elmat.cpp:197-225); the relations
   are the caller's (cells = dofs), as in tg_produce_data_algebraic (amg/src/tg.cpp:862-886) ---- */
namespace saamge_oracle
{
static void orc_extract_submatrices(const SparseMatrix &A, const agg_partitioning_relations_t &rels,
                                    std::vector<SparseMatrix *> &out)
{
    const int nparts = rels.nparts;
    out.assign(nparts, (SparseMatrix *)NULL);
    for (int part = 0; part < nparts; ++part)
    {
        const int localsize = rels.AE_to_dof->RowSize(part);
        const int *row = rels.AE_to_dof->GetRow(part);
        // LIL build with Set(), rows finalised with ascending columns
        std::vector<std::map<int, double>> lil(localsize);
        for (int i = 0; i < localsize; ++i)
        {
            const int glob_dof = row[i];
            for (int j = A.I[glob_dof]; j < A.I[glob_dof + 1]; ++j)
            {
                const int glob_neigh = A.J[j];
                if (agg_elem_in_col(glob_neigh, part, *rels.dof_to_AE) < 0)
                    continue;
                const int local_neigh = agg_map_id_glob_to_AE(glob_neigh, part, rels);
                if (0. != A.A[j])
                    lil[i][local_neigh] = A.A[j];
            }
        }
        if (localsize > 1)
        {
            for (int i = 0; i < localsize; ++i)
            {
                double rowsum = 0.;
                for (std::map<int, double>::const_iterator it = lil[i].begin(); it != lil[i].end(); ++it)
                    rowsum += it->second;
                if (lil[i].size() > 1)
                    lil[i][i] += -rowsum;
                if (lil[i][i] <= 0.0)
                    lil[i][i] = 1.0;
            }
        }
        else
            lil[0][0] = 1.0;
        SparseMatrix *S = new SparseMatrix;
        S->h = S->w = localsize;
        S->I.assign((size_t)localsize + 1, 0);
        for (int i = 0; i < localsize; ++i)
        {
            for (std::map<int, double>::const_iterator it = lil[i].begin(); it != lil[i].end(); ++it)
            {
                S->J.push_back(it->first);
                S->A.push_back(it->second);
            }
            S->I[i + 1] = (int)S->J.size();
        }
        out[part] = S;
    }
}

class OrcElementMatrixArray : public ElementMatrixProvider
{
public:
    OrcElementMatrixArray(const agg_partitioning_relations_t &rels, const std::vector<SparseMatrix *> &m)
        : ElementMatrixProvider(rels), m_(m)
    {
    }
    virtual ~OrcElementMatrixArray()
    {
        for (size_t i = 0; i < m_.size(); ++i)
            delete m_[i];
    }
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const
    {
        free_matr = true;
        return BuildAEStiff(elno);
    }
    // (a copy: the oracle's level owns and frees what BuildAEStiff returns)
    virtual SparseMatrix *BuildAEStiff(int elno) const { return new SparseMatrix(*m_[elno]); }

private:
    std::vector<SparseMatrix *> m_;
};
} // namespace saamge_oracle

extern "C" void *sa_orc_ml_build_algebraic(void *prob_, const sa_drv_params_t *p)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    sa_hierarchy_t *H = new sa_hierarchy_t;
    H->prob = prob;
    H->params = *p;
    oracle_ml_t *ml = new oracle_ml_t;
    H->impl = ml;
    H->impl_free = oracle_impl_free;
    const int coarsenings = p->num_levels - 1;
    std::vector<int> nparts_arr = sa_target_nparts(prob->fem->NE, *p);
    const bool avoid = p->avoid_ess_bdr_dofs != 0;
    const double tstart = now_s();
    H->rels.push_back(prob->rels);
    for (int i = 0; i < coarsenings; ++i)
    {
        oracle_level_t *L = new oracle_level_t;
        ml->levels.push_back(L);
        if (i == 0)
        {
            L->agg_part_rels = prob->rels;
            L->A = &prob->fem->A;
            std::vector<SparseMatrix *> mats;
            orc_extract_submatrices(prob->fem->A, *prob->rels, mats);
            L->elem_data = new OrcElementMatrixArray(*prob->rels, mats);
        }
        else
        {
            oracle_level_t *F = ml->levels[i - 1];
            int nparts = nparts_arr[i];
            int *partitioning = sa_prescribed_coarse_partitioning(*prob, *p, i, F->agg_part_rels->nparts, &nparts);
            agg_partitioning_relations_t *rels = orc_create_partitioning_coarse(
                *F->agg_part_rels, *F->ltent_interp, F->mis_numcoarsedof.data(), &nparts, partitioning);
            H->rels.push_back(rels);
            L->agg_part_rels = rels;
            L->A = F->Ac;
            L->elem_data = new ElementMatrixParallelCoarse(*rels, F);
        }
        L->drop_tol = p->smooth_drop_tol;
        tg_build_level(*L, i == 0 ? p->first_nu_pro : p->nu_pro, p->nu_relax,
                       i == 0 ? p->first_theta : p->theta, avoid, H->times, i);
    }
    if (p->correct_nullspace)
    {
        ml->levels.push_back(build_correct_nullspace(*ml->levels.back()));
        H->cn_P = *ml->levels.back()->interp;
        H->cn_Ac = *ml->levels.back()->Ac;
    }
    factor_coarsest(*ml->levels.back());
    H->times["setup"] = now_s() - tstart;
    H->levels.resize(coarsenings);
    for (int i = 0; i < coarsenings; ++i)
        export_level(*ml->levels[i], H->levels[i], i + 1 < coarsenings ? ml->levels[i + 1]->elem_data : NULL);
    return H;
}

/* adapt_update_operators (amg/src/adapt.cpp:171-216) on the oracle's hierarchy, with the problem's
   current operator values: smpr_update_Dinv_neg, tg_smooth_interp (amg/inc/tg.hpp:678-693) from
   the kept tentative prolongator when resmooth_interp, tg_update_coarse_operator (fresh RAP), level
   by level; then the coarsest factorisation (ml_impose_cycle). */
extern "C" int sa_orc_ml_update_operators(void *hier, int resmooth_interp)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    oracle_ml_t *ml = (oracle_ml_t *)H->impl;
    for (size_t i = 0; i < ml->levels.size(); ++i)
    {
        oracle_level_t &L = *ml->levels[i];
        // (a CorrectNullspace level at the end has no tentative prolongator -- nu_pro == 0 --: only
        // its smoother and its Galerkin operator follow the new coarsest Ac)
        if (i > 0)
            L.A = ml->levels[i - 1]->Ac; // Af = finer Ac
        delete L.Dinv_neg;
        L.Dinv_neg = mbox_build_Dinv_neg_parallel_matrix(*L.A);
        if (resmooth_interp && L.nu_pro > 0 && L.interp_smoother_degree > 0)
        {
            delete L.interp;
            delete L.restr;
            L.interp = interp_smooth(L.interp_smoother_degree, L.interp_smoother_roots, *L.A,
                                     *L.ltent_interp, *L.Dinv_neg, L.drop_tol);
            L.restr = new SparseMatrix;
            SpTranspose(*L.interp, *L.restr);
        }
        delete L.Ac; // tg_free_coarse_operator
        SparseMatrix AP;
        SpMultMat(*L.A, *L.interp, AP);
        L.Ac = new SparseMatrix;
        SpMultMat(*L.restr, AP, *L.Ac);
    }
    factor_coarsest(*ml->levels.back());
    const int coarsenings = (int)H->rels.size();
    if ((int)ml->levels.size() > coarsenings)
        H->cn_Ac = *ml->levels.back()->Ac;
    H->levels.clear();
    H->levels.resize(coarsenings);
    for (int i = 0; i < coarsenings; ++i)
        export_level(*ml->levels[i], H->levels[i], i + 1 < coarsenings ? ml->levels[i + 1]->elem_data : NULL);
    return 0;
}

extern "C" int sa_orc_ml_pcg(void *hier, int maxiter, double rtol, double atol)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    oracle_ml_t *ml = (oracle_ml_t *)H->impl;
    const SparseMatrix &A = H->prob->fem->A;
    const Vector &b = H->prob->fem->b;
    H->pcg.x.assign(b.size(), 0.);
    H->pcg.brr.clear();
    const double t0 = now_s();
    H->pcg.iterations = kalchev_pcg(A, *ml, b, H->pcg.x, maxiter, rtol, atol, &H->pcg.brr);
    H->times["pcg"] = now_s() - t0;
    Vector r(b.size());
    SpMult(A, H->pcg.x.data(), r.data());
    double s = 0.;
    for (size_t i = 0; i < b.size(); ++i)
        s += (b[i] - r[i]) * (b[i] - r[i]);
    H->pcg.final_res_norm = sqrt(s);
    return H->pcg.iterations;
}

extern "C" double sa_orc_time_local_spectral(void *prob_, const sa_drv_params_t *p,
                                             int ae_begin, int ae_end)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    const agg_partitioning_relations_t &rels = *prob->rels;
    ElementMatrixStandardGeometric emp(rels, prob->fem->A, prob->fem->elmat.data(),
                                       prob->fem->ne);
    ae_end = std::min(ae_end, rels.nparts);
    const double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = ae_begin; i < ae_end; ++i)
    {
        SparseMatrix *AE = emp.BuildAEStiff(i);
        SparseMatrix *B = NULL;
        DenseMatrix cut;
        Vector evals;
        DenseMatrix deA, deB;
        B = mbox_snd_D_sparse_from_sparse(*AE);
        mbox_convert_sparse_to_dense(*AE, deA);
        mbox_convert_sparse_to_dense(*B, deB);
        xpacks_calc_lower_eigens_dense(deA, evals, cut, deB, p->first_theta, true);
        delete AE;
        delete B;
    }
    return now_s() - t0;
}

/* Local spectral stage (a2-a7) of the finest level for AEs [ae_begin, ae_end), results returned:
   m_out[i - ae_begin] accepted vectors, evals_out[(i - ae_begin) * eval_cap + j] (first
   min(m, eval_cap) eigenvalues), D_out at the AE's offset in AE_to_dof (caller passes the full
   array).  Single threaded: callers that want all cores fork processes over AE ranges (the
   LAPACK in this image serialises concurrent calls from one process). */
extern "C" int sa_orc_local_spectral(void *prob_, const sa_drv_params_t *p, int ae_begin, int ae_end,
                                     int *m_out, double *evals_out, int eval_cap, double *D_out)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    const agg_partitioning_relations_t &rels = *prob->rels;
    ElementMatrixStandardGeometric emp(rels, prob->fem->A, prob->fem->elmat.data(), prob->fem->ne);
    ae_end = std::min(ae_end, rels.nparts);
    for (int i = ae_begin; i < ae_end; ++i)
    {
        SparseMatrix *AE = emp.BuildAEStiff(i);
        SparseMatrix *B = mbox_snd_D_sparse_from_sparse(*AE);
        DenseMatrix cut, deA, deB;
        Vector evals;
        mbox_convert_sparse_to_dense(*AE, deA);
        mbox_convert_sparse_to_dense(*B, deB);
        xpacks_calc_lower_eigens_dense(deA, evals, cut, deB, p->first_theta, true);
        m_out[i - ae_begin] = cut.Width();
        for (int j = 0; j < eval_cap; ++j)
            evals_out[(size_t)(i - ae_begin) * eval_cap + j] = j < (int)evals.size() ? evals[j] : 0.;
        const int off = rels.AE_to_dof->GetI()[i];
        for (int k = 0; k < AE->h; ++k)
            D_out[off + k] = B->A[k];
        delete AE;
        delete B;
    }
    return 0;
}

extern "C" int sa_orc_check_mises(void *prob_)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    const agg_partitioning_relations_t &rels = *prob->rels;
    std::vector<int> mises;
    Table mis_to_dof;
    agg_construct_mises_local_scan(*rels.dof_to_AE, mises, mis_to_dof);
    if (mis_to_dof.nrows != rels.num_mises)
        return 1;
    for (int d = 0; d < rels.ND; ++d)
        if (mises[d] != rels.mises[d])
            return 2;
    if (mis_to_dof.I != rels.mis_to_dof->I || mis_to_dof.J != rels.mis_to_dof->J)
        return 3;
    return 0;
}

extern "C" int sa_orc_check_relations(void *prob_)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    return orc_check_fine_relations(*prob->rels, prob->fem->bdr_dofs.data(), prob->fem->NE);
}

/* Builds the relations of coarse level `level` (>= 1) of an ORACLE hierarchy a second time with
   the PRODUCT's construction (saamge_b200/host/aggregates.cpp) from the same inputs and compares
   every table: 0 = identical, else the index of the first differing table (1 elem_to_dof,
   2 dof_to_elem, 3 AE_to_elem, 4 AE_to_dof, 5 dof_to_AE, 6 dof_id_inAE, 7 sizes, 8 mises,
   9 mis_to_dof, 10 mis_to_AE, 11 AE_to_mis, 12 agg_flags). */
extern "C" int sa_orc_check_coarse_relations(void *hier, int level)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    oracle_ml_t *ml = (oracle_ml_t *)H->impl;
    if (level < 1 || level >= (int)H->rels.size())
        return -1;
    const agg_partitioning_relations_t &fine = *H->rels[level - 1];
    const agg_partitioning_relations_t &o = *H->rels[level];
    int nparts = o.nparts;
    int *part = new int[fine.nparts];
    std::memcpy(part, o.partitioning, sizeof(int) * fine.nparts);
    agg_partitioning_relations_t *p = agg_create_partitioning_coarse(
        fine, ml->levels[level - 1]->mis_numcoarsedof.data(), &nparts,
        H->params.avoid_ess_bdr_dofs != 0, part);
    int bad = 0, idx = 0;
    auto cmp_t = [&](const Table *a, const Table *b) {
        ++idx;
        if (!bad && (a->nrows != b->nrows || a->I != b->I || a->J != b->J))
            bad = idx;
    };
    cmp_t(p->elem_to_dof, o.elem_to_dof);
    cmp_t(p->dof_to_elem, o.dof_to_elem);
    cmp_t(p->AE_to_elem, o.AE_to_elem);
    cmp_t(p->AE_to_dof, o.AE_to_dof);
    cmp_t(p->dof_to_AE, o.dof_to_AE);
    ++idx;
    if (!bad && std::memcmp(p->dof_id_inAE, o.dof_id_inAE, sizeof(int) * o.dof_to_AE->Size_of_connections()))
        bad = idx;
    ++idx;
    if (!bad && (p->ND != o.ND || p->num_mises != o.num_mises))
        bad = idx;
    ++idx;
    if (!bad && std::memcmp(p->mises, o.mises, sizeof(int) * o.ND))
        bad = idx;
    if (!bad)
    {
        cmp_t(p->mis_to_dof, o.mis_to_dof);
        cmp_t(p->mis_to_AE, o.mis_to_AE);
        cmp_t(p->AE_to_mis, o.AE_to_mis);
        ++idx;
        if (!bad && std::memcmp(p->agg_flags, o.agg_flags, (size_t)o.ND))
            bad = idx;
    }
    agg_free_partitioning(p);
    return bad;
}

/* ---- CPU baselines on inputs taken from a hierarchy the GPU built (bench.py) ---- */

/* Eigen stage of one AE given its assembled dense matrix (n x n column-major): weighted-l1 D
   (mbox_snd_D_sparse_from_sparse), densification of A and B, dsygvx
   (xpacks_calc_lower_eigens_dense) -- what Eigensolver::SolveDirect does after BuildAEStiff
   (amg/src/spectral.cpp:124-237).  Returns seconds; *m_out = accepted vectors. */
extern "C" double sa_orc_time_dense_AE(int n, const double *A, double theta, int *m_out)
{
    SparseMatrix Asp;
    Asp.h = Asp.w = n;
    Asp.I.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j)
            if (A[(size_t)j * n + i] != 0. || i == j)
            {
                Asp.J.push_back(j);
                Asp.A.push_back(A[(size_t)j * n + i]);
            }
        Asp.I[i + 1] = (int)Asp.J.size();
    }
    const double t0 = now_s();
    SparseMatrix *B = mbox_snd_D_sparse_from_sparse(Asp);
    DenseMatrix deA, deB, cut;
    Vector evals;
    mbox_convert_sparse_to_dense(Asp, deA);
    mbox_convert_sparse_to_dense(*B, deB);
    xpacks_calc_lower_eigens_dense(deA, evals, cut, deB, theta, true);
    const double t = now_s() - t0;
    if (m_out)
        *m_out = cut.Width();
    delete B;
    return t;
}

/* The reference's CPU solve (kalchev_pcg + V-cycle with the SAS polynomial smoother, OpenMP over
   rows) on the operators of a DOWNLOADED hierarchy (sa_drv_ml_download): same P, Ac, Dinv_neg as
   the GPU solve uses.  Runs at most run_iters iterations (a bounded sample of a long solve);
   returns seconds, *iters_out = iterations done (negative = not converged within run_iters). */
extern "C" double sa_orc_time_pcg_on_hierarchy(void *hier, int run_iters, double rtol, double atol,
                                               int *iters_out)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    const int nl = (int)H->levels.size();
    oracle_ml_t ml;
    for (int i = 0; i < nl; ++i)
    {
        oracle_level_t *L = new oracle_level_t;
        ml.levels.push_back(L);
        const sa_level_results_t &R = H->levels[i];
        L->A = (i == 0) ? &H->prob->fem->A : ml.levels[i - 1]->Ac;
        L->interp = new SparseMatrix(R.interp);
        L->restr = new SparseMatrix;
        SpTranspose(*L->interp, *L->restr);
        L->Ac = new SparseMatrix(R.Ac);
        L->Dinv_neg = new Vector(R.Dinv_neg);
        L->nu = H->params.nu_relax;
        L->roots = smpr_sas_poly_roots(L->nu, &L->degree);
    }
    factor_coarsest(*ml.levels.back());
    const SparseMatrix &A = H->prob->fem->A;
    const Vector &b = H->prob->fem->b;
    Vector x(b.size(), 0.);
    const double t0 = now_s();
    const int it = kalchev_pcg(A, ml, b, x, run_iters, rtol, atol, NULL);
    const double t = now_s() - t0;
    if (iters_out)
        *iters_out = it;
    return t;
}
