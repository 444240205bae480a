// TEST INFRASTRUCTURE -- independent restatement of the reference's topology construction
// (amg/src/aggregates.cpp), used by the oracle so that the "integer maps are bit-exact"
// check of tests/parity.py compares two separately written implementations (the product's
// hash / sort construction in saamge_b200/host/aggregates.cpp against this line-faithful one)
// instead of one function with itself.  Nothing here calls the product's table routines:
// mfem::Table's Mult / Transpose are restated below from MFEM's documented behaviour
// (Transpose: counting sort, rows ascending; Mult: marker array, first-encounter order).
#include <cmath>
#include <cstring>
#include <map>

#include "saamge_oracle.hpp"
#include "../saamge_b200/host/part.hpp"

namespace saamge_oracle
{

// mfem::Transpose(const Table &A, Table &At, int ncols_A)
void orc_table_transpose(const Table &A, Table &At, int ncols)
{
    At.nrows = ncols;
    At.ncols = A.nrows;
    At.I.assign((size_t)ncols + 1, 0);
    for (int i = 0; i < A.nrows; ++i)
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            At.I[A.J[p] + 1]++;
    for (int c = 0; c < ncols; ++c)
        At.I[c + 1] += At.I[c];
    At.J.assign(A.nrows ? A.I[A.nrows] : 0, 0);
    std::vector<int> fill(At.I.begin(), At.I.end() - 1);
    for (int i = 0; i < A.nrows; ++i)
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            At.J[fill[A.J[p]]++] = i;
}

// mfem::Mult(const Table &A, const Table &B, Table &C)
void orc_table_mult(const Table &A, const Table &B, Table &C)
{
    std::vector<int> B_marker(B.ncols, -1);
    C.nrows = A.nrows;
    C.ncols = B.ncols;
    C.I.assign((size_t)A.nrows + 1, 0);
    C.J.clear();
    for (int i = 0; i < A.nrows; ++i)
    {
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
        {
            const int j = A.J[p];
            for (int q = B.I[j]; q < B.I[j + 1]; ++q)
            {
                const int k = B.J[q];
                if (B_marker[k] != i)
                {
                    C.J.push_back(k);
                    B_marker[k] = i;
                }
            }
        }
        C.I[i + 1] = (int)C.J.size();
    }
}

// agg_construct_tables_from_arr (amg/src/aggregates.cpp:218-236)
static void orc_tables_from_arr(const int *arr, int n, int nparts, Table *&elem_to_AE,
                                Table *&AE_to_elem)
{
    elem_to_AE = new Table;
    elem_to_AE->nrows = n;
    elem_to_AE->ncols = nparts;
    elem_to_AE->I.resize((size_t)n + 1);
    elem_to_AE->J.assign(arr, arr + n);
    for (int i = 0; i <= n; ++i)
        elem_to_AE->I[i] = i;
    AE_to_elem = new Table;
    orc_table_transpose(*elem_to_AE, *AE_to_elem, nparts);
}

// agg_build_glob_to_AE_id_map (amg/src/aggregates.cpp:1202-1244): for every AE and every local
// position j of a dof in AE_to_dof, store j at the dof's slot of that AE in dof_to_AE
static void orc_build_glob_to_AE_id_map(agg_partitioning_relations_t &r)
{
    const Table &AE_to_dof = *r.AE_to_dof, &dof_to_AE = *r.dof_to_AE;
    r.dof_id_inAE = new int[std::max(1, dof_to_AE.Size_of_connections())];
    for (int i = 0; i < r.nparts; ++i)
        for (int j = 0; j < AE_to_dof.RowSize(i); ++j)
        {
            const int dof = AE_to_dof.GetRow(i)[j];
            int pos = -1;
            for (int q = 0; q < dof_to_AE.RowSize(dof); ++q)
                if (dof_to_AE.GetRow(dof)[q] == i)
                    pos = q;
            SA_ASSERT(pos >= 0);
            r.dof_id_inAE[dof_to_AE.I[dof] + pos] = j;
        }
}

// agg_construct_mises_local + agg_produce_mises (amg/src/aggregates.cpp:501-653, 712-853) on
// one process.  The reference scans, for every not yet distributed dof i, ALL dofs k and puts
// k into the MIS of i iff both belong to exactly the same AEs (count[k] == rowsum[k] ==
// rowsum[i]); MISes are therefore numbered by their first dof and list their dofs ascending.
// The same classes and the same numbering follow from keying the dofs by their (ascending) AE
// list -- rows of dof_to_AE are ascending because the table is a Transpose.
static void orc_produce_mises(agg_partitioning_relations_t &r)
{
    const Table &dof_to_AE = *r.dof_to_AE;
    const int nd = dof_to_AE.Size();
    std::map<std::vector<int>, int> classes;
    std::vector<std::vector<int>> rows;
    r.mises = new int[std::max(1, nd)];
    for (int i = 0; i < nd; ++i)
    {
        std::vector<int> key(dof_to_AE.GetRow(i), dof_to_AE.GetRow(i) + dof_to_AE.RowSize(i));
        std::map<std::vector<int>, int>::iterator it = classes.find(key);
        int mis;
        if (it == classes.end())
        {
            mis = (int)rows.size();
            classes[key] = mis;
            rows.push_back(std::vector<int>());
        }
        else
            mis = it->second;
        rows[mis].push_back(i);
        r.mises[i] = mis;
    }
    r.num_mises = r.num_owned_mises = (int)rows.size();
    Table *m2d = new Table;
    m2d->nrows = r.num_mises;
    m2d->ncols = nd;
    m2d->I.assign((size_t)r.num_mises + 1, 0);
    for (int m = 0; m < r.num_mises; ++m)
    {
        m2d->J.insert(m2d->J.end(), rows[m].begin(), rows[m].end());
        m2d->I[m + 1] = (int)m2d->J.size();
    }
    r.mis_to_dof = m2d;
    r.truemis_to_dof = new Table(*m2d);
    r.mises_size = new int[std::max(1, r.num_mises)];
    r.mis_master = new int[std::max(1, r.num_mises)];
    for (int m = 0; m < r.num_mises; ++m)
    {
        r.mises_size[m] = m2d->RowSize(m);
        r.mis_master[m] = 0;
    }
    // amg/src/aggregates.cpp:773-774
    r.mis_to_AE = new Table;
    orc_table_mult(*r.mis_to_dof, dof_to_AE, *r.mis_to_AE);
    r.AE_to_mis = new Table;
    orc_table_transpose(*r.mis_to_AE, *r.AE_to_mis, r.nparts);
}

// agg_construct_agg_flags (amg/src/aggregates.cpp:198-216)
static void orc_construct_agg_flags(agg_partitioning_relations_t &r, const agg_dof_status_t *bdr_dofs)
{
    const int ND = r.dof_to_AE->Size();
    r.agg_flags = new agg_dof_status_t[std::max(1, ND)];
    for (int i = 0; i < ND; ++i)
    {
        agg_dof_status_t f = bdr_dofs ? bdr_dofs[i] : 0;
        if ((f & AGG_ON_PROC_IFACE_FLAG) || r.dof_to_AE->RowSize(i) > 1)
            f |= AGG_BETWEEN_AES_FLAG;
        r.agg_flags[i] = f;
    }
}

// the tables every level derives from (partitioning, elem_to_dof):
// amg/src/aggregates.cpp:1357-1443 (fine), :1481-1602 (coarse)
static void orc_create_tables(agg_partitioning_relations_t &r, int NE, Table *elem_to_dof,
                              const agg_dof_status_t *bdr_dofs)
{
    r.elem_to_dof = elem_to_dof;
    r.dof_to_elem = new Table;
    orc_table_transpose(*elem_to_dof, *r.dof_to_elem, elem_to_dof->ncols);
    r.ND = r.dof_to_elem->Size();
    orc_tables_from_arr(r.partitioning, NE, r.nparts, r.elem_to_AE, r.AE_to_elem);
    r.AE_to_dof = new Table;
    orc_table_mult(*r.AE_to_elem, *r.elem_to_dof, *r.AE_to_dof);
    r.dof_to_AE = new Table;
    orc_table_transpose(*r.AE_to_dof, *r.dof_to_AE, r.ND);
    orc_build_glob_to_AE_id_map(r);
    orc_produce_mises(r);
    orc_construct_agg_flags(r, bdr_dofs);
}

// agg_create_partitioning_coarse + agg_create_rels_except_elem_coarse +
// agg_build_coarse_Dof_TrueDof (amg/src/aggregates.cpp:1735-1832, 1481-1602, 1611-1730) on one
// process.  finedof_to_dof is the pattern of the tentative prolongator (amg/src/ml.cpp:150-154
// passes tg_data->tent_interp; amg/src/aggregates.cpp:1445-1479), elem_to_dof = fine AE_to_dof *
// finedof_to_dof.  \a partitioning: given (fixtures / shared METIS input) or NULL => METIS on
// the AE graph with AE-dof-count weights (:1797-1804).
agg_partitioning_relations_t *
orc_create_partitioning_coarse(const agg_partitioning_relations_t &fine, const SparseMatrix &tent_interp,
                               const int *mis_numcoarsedof, int *nparts, int *partitioning)
{
    agg_partitioning_relations_t *rels = new agg_partitioning_relations_t;
    std::memset(rels, 0, sizeof(*rels));
    rels->testmesh = fine.testmesh;
    // coarse dofs are numbered MIS-major (amg/src/aggregates.cpp:1687-1695)
    rels->mis_coarsedofoffsets = new int[(size_t)fine.num_mises + 1];
    int off = 0;
    for (int mis = 0; mis < fine.num_mises; ++mis)
    {
        rels->mis_coarsedofoffsets[mis] = off;
        off += mis_numcoarsedof[mis];
    }
    rels->mis_coarsedofoffsets[fine.num_mises] = off;
    rels->dof_masterproc = new int[std::max(1, off)];
    std::memset(rels->dof_masterproc, 0, sizeof(int) * std::max(1, off));
    SA_ASSERT(tent_interp.w == off);

    // elem_to_elem = AE_to_elem * elem_to_elem * elem_to_AE (:1768-1771)
    Table tmptbl;
    rels->elem_to_elem = new Table;
    orc_table_mult(*fine.AE_to_elem, *fine.elem_to_elem, tmptbl);
    orc_table_mult(tmptbl, *fine.elem_to_AE, *rels->elem_to_elem);
    SA_ASSERT(rels->elem_to_elem->Size() == fine.nparts);

    const int num_elem = fine.nparts;
    if (partitioning)
        rels->partitioning = partitioning;
    else
    {
        // METIS is a shared input producer (partitions are inputs of both sides); it
        // does not accept the self loops the product has
        std::vector<int> weights(num_elem);
        for (int i = 0; i < num_elem; ++i)
            weights[i] = fine.AE_to_dof->RowSize(i);
        Table graph;
        graph.nrows = graph.ncols = num_elem;
        graph.I.assign((size_t)num_elem + 1, 0);
        for (int i = 0; i < num_elem; ++i)
        {
            for (int k = 0; k < rels->elem_to_elem->RowSize(i); ++k)
                if (rels->elem_to_elem->GetRow(i)[k] != i)
                    graph.J.push_back(rels->elem_to_elem->GetRow(i)[k]);
            graph.I[i + 1] = (int)graph.J.size();
        }
        rels->partitioning = part_generate_partitioning(graph, weights.data(), nparts);
    }
    rels->nparts = *nparts;

    // finedof_to_dof: the diag block of Dof_TrueDof * interp * TrueDof_Dof^T is interp's own
    // pattern on one process (:1445-1479).
    // DEVIATION (documented in DESIGN.md): upstream takes the raw pattern, whose exact zeros are
    // an accident of LAPACK round-off: rows of essential boundary dofs are zeroed BEFORE the SVD
    // (contrib_filter_boundary, amg/src/contrib.cpp:102-163) and dgesvd returns either exact zeros
    // or dust (~1e-16) there; symmetric agglomerates give singular vectors with entries that are
    // zero up to round-off; contrib_tent_insert_simple keeps every entry with abs(v) > 0 (:188).
    // Whether such an entry exists decides the first-encounter ORDER of the coarse dofs inside
    // coarse elements / AEs, so upstream's local numbering is not reproducible across LAPACK
    // builds.  The rule used on both sides here: a fine dof whose prolongator row has at least one
    // entry above 1e-10 of the largest entry connects to ALL coarse dofs of its MIS (in order);
    // a row without such an entry (essential boundary) connects to none.
    double pmax = 0.;
    for (size_t q = 0; q < tent_interp.A.size(); ++q)
        pmax = std::max(pmax, fabs(tent_interp.A[q]));
    Table finedof_to_dof;
    finedof_to_dof.nrows = tent_interp.h;
    finedof_to_dof.ncols = tent_interp.w;
    finedof_to_dof.I.assign((size_t)tent_interp.h + 1, 0);
    for (int d = 0; d < tent_interp.h; ++d)
    {
        bool live = false;
        for (int q = tent_interp.I[d]; q < tent_interp.I[d + 1]; ++q)
            if (fabs(tent_interp.A[q]) > 1e-10 * pmax)
            {
                live = true;
                // every entry of the row lies in the block of the dof's own MIS
                SA_ASSERT(tent_interp.J[q] >= rels->mis_coarsedofoffsets[fine.mises[d]] &&
                          tent_interp.J[q] < rels->mis_coarsedofoffsets[fine.mises[d] + 1]);
            }
        if (live)
            for (int c = rels->mis_coarsedofoffsets[fine.mises[d]];
                 c < rels->mis_coarsedofoffsets[fine.mises[d] + 1]; ++c)
                finedof_to_dof.J.push_back(c);
        finedof_to_dof.I[d + 1] = (int)finedof_to_dof.J.size();
    }
    Table *elem_to_dof = new Table;
    orc_table_mult(*fine.AE_to_dof, finedof_to_dof, *elem_to_dof);
    elem_to_dof->ncols = off;
    orc_create_tables(*rels, num_elem, elem_to_dof, NULL);
    SA_ASSERT(rels->ND == off);
    return rels;
}

// Rebuilds the fine-level tables of a problem from (elem_to_dof, partitioning, boundary flags)
// and compares every table of the product's construction with them.  Returns 0 when all agree,
// else the (1-based) index of the first differing table.
int orc_check_fine_relations(const agg_partitioning_relations_t &p, const agg_dof_status_t *bdr_dofs, int NE)
{
    agg_partitioning_relations_t r;
    std::memset(&r, 0, sizeof r);
    r.nparts = p.nparts;
    r.partitioning = p.partitioning;
    Table *e2d = new Table(*p.elem_to_dof);
    orc_create_tables(r, NE, e2d, bdr_dofs);
    int bad = 0, idx = 0;
    auto cmp_t = [&](const Table *a, const Table *b) {
        ++idx;
        if (!bad && (a->nrows != b->nrows || a->I != b->I || a->J != b->J))
            bad = idx;
    };
    auto cmp_a = [&](const int *a, const int *b, size_t n) {
        ++idx;
        if (!bad && n && std::memcmp(a, b, n * sizeof(int)) != 0)
            bad = idx;
    };
    cmp_t(r.dof_to_elem, p.dof_to_elem);   // 1
    cmp_t(r.elem_to_AE, p.elem_to_AE);     // 2
    cmp_t(r.AE_to_elem, p.AE_to_elem);     // 3
    cmp_t(r.AE_to_dof, p.AE_to_dof);       // 4
    cmp_t(r.dof_to_AE, p.dof_to_AE);       // 5
    cmp_a(r.dof_id_inAE, p.dof_id_inAE, (size_t)r.dof_to_AE->Size_of_connections()); // 6
    ++idx;                                 // 7
    if (!bad && (r.num_mises != p.num_mises || r.ND != p.ND))
        bad = idx;
    cmp_a(r.mises, p.mises, (size_t)r.ND); // 8
    if (!bad)
    {
        cmp_t(r.mis_to_dof, p.mis_to_dof); // 9
        cmp_t(r.mis_to_AE, p.mis_to_AE);   // 10
        cmp_t(r.AE_to_mis, p.AE_to_mis);   // 11
        cmp_a(r.mises_size, p.mises_size, (size_t)r.num_mises); // 12
        ++idx;                             // 13
        if (!bad && std::memcmp(r.agg_flags, p.agg_flags, (size_t)r.ND * sizeof(agg_dof_status_t)) != 0)
            bad = idx;
    }
    delete r.dof_to_elem;
    delete r.elem_to_dof;
    delete r.elem_to_AE;
    delete r.AE_to_elem;
    delete r.AE_to_dof;
    delete r.dof_to_AE;
    delete[] r.dof_id_inAE;
    delete[] r.mises;
    delete r.mis_to_dof;
    delete r.truemis_to_dof;
    delete r.mis_to_AE;
    delete r.AE_to_mis;
    delete[] r.mises_size;
    delete[] r.mis_master;
    delete[] r.agg_flags;
    return bad;
}

} // namespace saamge_oracle
