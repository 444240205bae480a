// Algebraic (matrix-only) entry of the hierarchy builder:
//   fem_create_partitioning_from_matrix   amg/src/fem.cpp:720-762 (+ the isolate variant,
//                                         amg/src/aggregates.cpp:1250-1314)
//   ExtractSubMatrices                    amg/src/tg.cpp:580-668
//   ElementMatrixArray                    amg/inc/elmat.hpp:142-151, amg/src/elmat.cpp:197-225
//   tg_produce_data_algebraic             amg/src/tg.cpp:862-886
// The user has no element matrices: the "cells" are the dofs, the agglomerates are parts of the
// matrix graph, and the local matrix of an agglomerate is its principal submatrix with the row
// sums moved to the diagonal (so that the constant vector stays in the local near-null space).
// The device path is the ordinary one: the agglomerate matrices are handed to it as ONE dense
// element block per agglomerate (relations whose elements are the agglomerates), assembled
// without the global operator -- exactly what ElementMatrixArray::BuildAEStiff returns in the
// reference.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>

#include "hierarchy.hpp"
#include "part.hpp"
#include "saamge.hpp"

namespace saamge
{

/* amg/src/tg.cpp:580-668 */
void ExtractSubMatrices(const SparseMatrix &A, const agg_partitioning_relations_t &agg_part_rels,
                        std::vector<SparseMatrix *> &agglomerate_element_matrices)
{
    const int nparts = agg_part_rels.nparts;
    agglomerate_element_matrices.assign(nparts, (SparseMatrix *)NULL);
    std::vector<int> loc(A.Height(), -1);
    for (int part = 0; part < nparts; ++part)
    {
        const int n = agg_part_rels.AE_to_dof->RowSize(part);
        const int *row = agg_part_rels.AE_to_dof->GetRow(part);
        for (int i = 0; i < n; ++i)
            loc[row[i]] = i;
        SparseMatrix *S = new SparseMatrix;
        S->h = S->w = n;
        S->I.assign((size_t)n + 1, 0);
        // principal submatrix, explicit zeros dropped; columns ascending in the LOCAL numbering
        // (what Finalize() of the reference's LIL matrix produces)
        std::vector<std::pair<int, double>> ent;
        for (int i = 0; i < n; ++i)
        {
            const int g = row[i];
            ent.clear();
            for (int p = A.I[g]; p < A.I[g + 1]; ++p)
            {
                const int lj = loc[A.J[p]];
                if (lj >= 0 && 0. != A.A[p])
                    ent.push_back(std::make_pair(lj, A.A[p]));
            }
            std::sort(ent.begin(), ent.end());
            for (size_t q = 0; q < ent.size(); ++q)
            {
                S->J.push_back(ent[q].first);
                S->A.push_back(ent[q].second);
            }
            S->I[i + 1] = (int)S->J.size();
        }
        for (int i = 0; i < n; ++i)
            loc[row[i]] = -1;
        // row sums to the diagonal; single cells get a 1
        if (n > 1)
        {
            for (int i = 0; i < n; ++i)
            {
                double rowsum = 0.;
                int dpos = -1;
                for (int p = S->I[i]; p < S->I[i + 1]; ++p)
                {
                    rowsum += S->A[p];
                    if (S->J[p] == i)
                        dpos = p;
                }
                SA_ASSERT(dpos >= 0);
                if (S->RowSize(i) > 1)
                    S->A[dpos] += -rowsum;
                if (S->A[dpos] <= 0.0)
                    S->A[dpos] = 1.0;
            }
        }
        else
        {
            S->I[1] = 1;
            S->J.assign(1, 0);
            S->A.assign(1, 1.0);
        }
        agglomerate_element_matrices[part] = S;
    }
}

/* amg/inc/elmat.hpp:142-151.  Takes ownership of the matrices.  Besides the reference contract
   (both virtuals return the agglomerate matrix) it carries the batched view of the device path:
   relations whose elements are the agglomerates and one dense block per agglomerate. */
ElementMatrixArray::ElementMatrixArray(const agg_partitioning_relations_t &rels,
                                       const std::vector<SparseMatrix *> &elem_matrs)
    : ElementMatrixProvider(rels), elem_matrs_(elem_matrs), brels_(NULL)
{
    const int nparts = rels.nparts;
    SA_ASSERT((int)elem_matrs.size() == nparts);
    offsets_.assign((size_t)nparts + 1, 0);
    for (int e = 0; e < nparts; ++e)
    {
        const int64_t n = rels.AE_to_dof->RowSize(e);
        SA_ASSERT(elem_matrs[e] && elem_matrs[e]->Height() == n);
        offsets_[e + 1] = offsets_[e] + n * n;
    }
    blocks_.assign((size_t)offsets_[nparts], 0.);
    for (int e = 0; e < nparts; ++e)
    {
        const SparseMatrix &S = *elem_matrs[e];
        const int n = S.Height();
        double *dst = blocks_.data() + offsets_[e];
        for (int i = 0; i < n; ++i)
            for (int p = S.I[i]; p < S.I[i + 1]; ++p)
                dst[(size_t)S.J[p] * n + i] += S.A[p];
    }
    // elements = agglomerates: elem_to_dof = AE_to_dof, partitioning = identity; every other
    // table (AE_to_dof order, dof_to_AE, MISes, flags) comes out as in the caller's relations
    Table *e2d = new Table(*rels.AE_to_dof);
    Table *e2e = new Table;
    e2e->nrows = e2e->ncols = nparts;
    e2e->I.assign((size_t)nparts + 1, 0);
    int *partitioning = new int[nparts];
    for (int e = 0; e < nparts; ++e)
        partitioning[e] = e;
    std::vector<agg_dof_status_t> bdr(rels.ND, 0);
    for (int d = 0; d < rels.ND; ++d)
        if (agg_is_dof_on_essential_border(rels, d))
            bdr[d] = AGG_ON_ESS_DOMAIN_BORDER_FLAG;
    int np = nparts;
    brels_ = agg_create_partitioning_fine(nparts, e2d, e2e, partitioning, bdr.data(), &np, false);
    SA_ASSERT(np == nparts && brels_->ND == rels.ND);
}

ElementMatrixArray::~ElementMatrixArray()
{
    for (size_t i = 0; i < elem_matrs_.size(); ++i)
        delete elem_matrs_[i];
    if (brels_)
        agg_free_partitioning(brels_);
}

Matrix *ElementMatrixArray::GetMatrix(int elno, bool &free_matr) const
{
    free_matr = false;
    return BuildAEStiff(elno);
}

SparseMatrix *ElementMatrixArray::BuildAEStiff(int elno) const
{
    SA_ASSERT(0 <= elno && elno < (int)elem_matrs_.size() && elem_matrs_[elno]);
    return elem_matrs_[elno];
}

/* amg/src/fem.cpp:720-762 with isolated cells as in agg_create_partitioning_fine_isolate
   (amg/src/aggregates.cpp:1250-1314): cells = dofs, elem_to_elem = graph of A (no self loops),
   METIS k-way on that graph, isolated cells moved to parts of their own, connected components. */
agg_partitioning_relations_t *fem_create_partitioning_from_matrix(const SparseMatrix &A, int *nparts,
                                                                   const std::vector<int> &isolated_cells)
{
    const int n = A.Height();
    Table *elem_to_elem = new Table;
    elem_to_elem->nrows = elem_to_elem->ncols = n;
    elem_to_elem->I.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            if (A.J[p] != i)
                elem_to_elem->J.push_back(A.J[p]);
        elem_to_elem->I[i + 1] = (int)elem_to_elem->J.size();
    }
    Table *elem_to_dof = new Table;
    elem_to_dof->nrows = elem_to_dof->ncols = n;
    elem_to_dof->I.resize((size_t)n + 1);
    elem_to_dof->J.resize(n);
    for (int i = 0; i <= n; ++i)
        elem_to_dof->I[i] = i;
    for (int i = 0; i < n; ++i)
        elem_to_dof->J[i] = i;
    std::vector<agg_dof_status_t> bdr_dofs(n, 0);
    int *partitioning = NULL;
    if (!isolated_cells.empty())
    {
        // The isolated cells are taken out of the graph before METIS sees it (the reference's
        // driver eliminates them from the matrix first, amg/test/algebraic/algebraic.cpp:224-238)
        std::vector<int> keep(n, 1), newid(n, -1);
        for (size_t q = 0; q < isolated_cells.size(); ++q)
            keep[isolated_cells[q]] = 0;
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (keep[i])
                newid[i] = m++;
        Table g;
        g.nrows = g.ncols = m;
        g.I.assign((size_t)m + 1, 0);
        for (int i = 0; i < n; ++i)
        {
            if (!keep[i])
                continue;
            for (int p = elem_to_elem->I[i]; p < elem_to_elem->I[i + 1]; ++p)
                if (keep[elem_to_elem->J[p]])
                    g.J.push_back(newid[elem_to_elem->J[p]]);
            g.I[newid[i] + 1] = (int)g.J.size();
        }
        int *sub = part_generate_partitioning_unweighted(g, nparts);
        partitioning = new int[n];
        int c_elem = *nparts;
        for (int i = 0; i < n; ++i)
            partitioning[i] = keep[i] ? sub[newid[i]] : -1;
        for (size_t q = 0; q < isolated_cells.size(); ++q)
            partitioning[isolated_cells[q]] = c_elem++;
        *nparts = c_elem;
        delete[] sub;
        std::vector<int> pv(partitioning, partitioning + n);
        *nparts = connectedComponents(pv, *elem_to_elem);
        std::copy(pv.begin(), pv.end(), partitioning);
    }
    return agg_create_partitioning_fine(n, elem_to_dof, elem_to_elem, partitioning, bdr_dofs.data(),
                                        nparts, false);
}

/* amg/src/tg.cpp:862-886.  \a Alocal and \a Ag are the same matrix on one process (the
   reference asserts PROC_NUM == 1 and says so); the window variant (WindowSubMatrices) is not
   implemented.  The returned tg_data owns the provider. */
tg_data_t *tg_produce_data_algebraic(const SparseMatrix &Alocal, const SparseMatrix &Ag,
                                     const agg_partitioning_relations_t &agg_part_rels, int nu_pro,
                                     int nu_relax, double spectral_tol, bool smooth_interp,
                                     int polynomial_coarse_arg, bool use_window, bool use_arpack,
                                     bool avoid_ess_bdr_dofs)
{
    SA_ASSERT(!use_window);
    std::vector<SparseMatrix *> mats;
    ExtractSubMatrices(Alocal, agg_part_rels, mats);
    ElementMatrixArray *emp = new ElementMatrixArray(agg_part_rels, mats);
    return tg_produce_data(Ag, emp->BatchedRelations(), nu_pro, nu_relax, emp, spectral_tol,
                           smooth_interp, polynomial_coarse_arg, use_arpack, avoid_ess_bdr_dofs);
}

/* multilevel form of the same (the reference's driver is two-level; coarser levels are built
   from the device-resident coarse element matrices like in the geometric case) */
ml_data_t *ml_produce_data_algebraic(const SparseMatrix &Ag, const agg_partitioning_relations_t &agg_part_rels,
                                     const MultilevelParameters &mlp)
{
    std::vector<SparseMatrix *> mats;
    ExtractSubMatrices(Ag, agg_part_rels, mats);
    ElementMatrixArray *emp = new ElementMatrixArray(agg_part_rels, mats);
    return ml_produce_data(Ag, const_cast<agg_partitioning_relations_t *>(&emp->BatchedRelations()), emp,
                           mlp);
}

/* ReadHypreMat of the reference's driver (amg/test/algebraic/algebraic.cpp:63-85): text file
   "row0 row1 col0 col1" then "i j value" triplets; duplicates are added. */
SparseMatrix *ReadHypreMat(const char *filename)
{
    std::ifstream in(filename);
    SA_ASSERT(in.good());
    int row0, row1, col0, col1;
    in >> row0 >> row1 >> col0 >> col1;
    SA_ASSERT(row0 == 0 && col0 == 0);
    const int h = row1 + 1, w = col1 + 1;
    std::vector<std::vector<std::pair<int, double>>> rows(h);
    int i, j;
    double x;
    while (in >> i >> j >> x)
    {
        SA_ASSERT(0 <= i && i < h && 0 <= j && j < w);
        rows[i].push_back(std::make_pair(j, x));
    }
    SparseMatrix *out = new SparseMatrix;
    out->h = h;
    out->w = w;
    out->I.assign((size_t)h + 1, 0);
    for (int r = 0; r < h; ++r)
    {
        std::stable_sort(rows[r].begin(), rows[r].end(),
                         [](const std::pair<int, double> &a, const std::pair<int, double> &b) {
                             return a.first < b.first;
                         });
        for (size_t q = 0; q < rows[r].size(); ++q)
        {
            if (!out->J.empty() && (int)out->J.size() > out->I[r] && out->J.back() == rows[r][q].first)
                out->A.back() += rows[r][q].second;
            else
            {
                out->J.push_back(rows[r][q].first);
                out->A.push_back(rows[r][q].second);
            }
        }
        out->I[r + 1] = (int)out->J.size();
    }
    return out;
}

} // namespace saamge

using namespace saamge;

/* ---- driver API (ctypes): a problem handle from a matrix, partitioned like the reference's
   algebraic driver does it (amg/test/algebraic/algebraic.cpp:219-262) ---- */

/* CSR matrix (n x n), rhs b (NULL: all ones, the driver's bg = 1.0).  isolated[0..n_isolated):
   dofs that become agglomerates of their own (the driver's eliminated dof 0: decoupled rows). */
extern "C" void *sa_drv_problem_from_matrix(int n, const int *I, const int *J, const double *A,
                                            const double *b, int elems_per_agg, const int *isolated,
                                            int n_isolated)
{
    sa_problem_t *prob = new sa_problem_t;
    fem_problem_t *f = new fem_problem_t;
    prob->fem = f;
    f->dim = 0;
    f->NE = n;
    f->ND = n;
    f->ne = 1;
    f->A.h = f->A.w = n;
    f->A.I.assign(I, I + n + 1);
    f->A.J.assign(J, J + I[n]);
    f->A.A.assign(A, A + I[n]);
    if (b)
        f->b.assign(b, b + n);
    else
        f->b.assign(n, 1.0);
    f->bdr_dofs.assign(n, 0);
    f->elmat.assign(n, 0.); // (1 x 1 "cell matrices": the diagonal; not used by the algebraic path)
    for (int i = 0; i < n; ++i)
        f->elmat[i] = f->A(i, i);
    std::vector<int> iso(isolated, isolated + n_isolated);
    // nparts from the number of NON-isolated cells (the driver partitions the eliminated matrix)
    int nparts = std::max(1, (n - n_isolated) / std::max(1, elems_per_agg));
    prob->target_nparts0 = nparts;
    if (iso.empty())
        prob->rels = fem_create_partitioning_from_matrix(f->A, &nparts, iso);
    else
        prob->rels = fem_create_partitioning_from_matrix(f->A, &nparts, iso);
    f->elem_to_dof = *prob->rels->elem_to_dof;
    f->elem_to_elem = *prob->rels->elem_to_elem;
    return prob;
}

extern "C" void *sa_drv_problem_from_hypre_file(const char *path, int elems_per_agg, int isolate_first)
{
    SparseMatrix *M = ReadHypreMat(path);
    const int iso0 = 0;
    void *p = sa_drv_problem_from_matrix(M->Height(), M->GetI(), M->GetJ(), M->GetData(), NULL,
                                         elems_per_agg, &iso0, isolate_first ? 1 : 0);
    delete M;
    return p;
}
