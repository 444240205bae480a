// Partitioning relations on one process (see aggregates.hpp).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "aggregates.hpp"
#include "part.hpp"

#include <algorithm>
#include <cstring>
#include <unordered_map>

namespace saamge
{

namespace
{
// SA_HOST_PROFILE=1: wall-clock laps of the host-side topology construction on stderr
struct Lap
{
    bool on;
    std::chrono::steady_clock::time_point t;
    Lap() : on(getenv("SA_HOST_PROFILE") != NULL), t(std::chrono::steady_clock::now()) {}
    void operator()(const char *what)
    {
        if (!on)
            return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[host] %-28s %8.1f ms\n", what,
                std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};
} // namespace


void agg_construct_agg_flags(agg_partitioning_relations_t &agg_part_rels,
                             const agg_dof_status_t *bdr_dofs)
{
    // amg/src/aggregates.cpp:198-216
    const int ND = agg_part_rels.dof_to_AE->Size();
    SA_ASSERT(!agg_part_rels.agg_flags);
    agg_part_rels.agg_flags = new agg_dof_status_t[ND];
    for (int i = 0; i < ND; ++i)
    {
        agg_part_rels.agg_flags[i] = bdr_dofs ? bdr_dofs[i] : 0;
        if (SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[i], AGG_ON_PROC_IFACE_FLAG) ||
            agg_part_rels.dof_to_AE->RowSize(i) > 1)
            agg_part_rels.agg_flags[i] |= AGG_BETWEEN_AES_FLAG;
    }
}

void agg_build_glob_to_AE_id_map(agg_partitioning_relations_t &agg_part_rels)
{
    // amg/src/aggregates.cpp:1202-1244
    const Table &AE_to_dof = *agg_part_rels.AE_to_dof;
    const Table &dof_to_AE = *agg_part_rels.dof_to_AE;
    agg_part_rels.dof_id_inAE = new int[std::max(1, dof_to_AE.Size_of_connections())];
    const int *I = dof_to_AE.GetI();
    // (every (AE, dof) pair writes its own slot: AEs are independent)
#pragma omp parallel for num_threads(sa_host_threads()) schedule(dynamic, 64)
    for (int i = 0; i < agg_part_rels.nparts; ++i)
    {
        const int *row = AE_to_dof.GetRow(i);
        const int rs = AE_to_dof.RowSize(i);
        for (int j = 0; j < rs; ++j)
        {
            const int dof = row[j];
            const int pos = agg_elem_in_col(dof, i, dof_to_AE);
            SA_ASSERT(pos >= 0);
            agg_part_rels.dof_id_inAE[pos + I[dof]] = j;
        }
    }
}

void agg_produce_mises(agg_partitioning_relations_t &agg_part_rels)
{
    const Table &dof_to_AE = *agg_part_rels.dof_to_AE;
    const int ND = dof_to_AE.Size();
    agg_part_rels.mises = new int[std::max(1, ND)];
    std::unordered_map<uint64_t, std::vector<int>> buckets; // hash -> MIS ids
    buckets.reserve((size_t)ND / 4 + 16);
    std::vector<int> mis_first; // first dof of each MIS
    std::vector<int> mis_count;
    for (int d = 0; d < ND; ++d)
    {
        const int *row = dof_to_AE.GetRow(d);
        const int rs = dof_to_AE.RowSize(d);
        uint64_t h = 1469598103934665603ull ^ (uint64_t)rs;
        for (int k = 0; k < rs; ++k)
        {
            h ^= (uint64_t)(uint32_t)row[k] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
            h *= 1099511628211ull;
        }
        std::vector<int> &cand = buckets[h];
        int found = -1;
        for (size_t c = 0; c < cand.size() && found < 0; ++c)
        {
            const int f = mis_first[cand[c]];
            if (dof_to_AE.RowSize(f) == rs &&
                0 == std::memcmp(dof_to_AE.GetRow(f), row, sizeof(int) * rs))
                found = cand[c];
        }
        if (found < 0)
        {
            found = (int)mis_first.size();
            mis_first.push_back(d);
            mis_count.push_back(0);
            cand.push_back(found);
        }
        agg_part_rels.mises[d] = found;
        mis_count[found]++;
    }
    const int num_mises = (int)mis_first.size();
    agg_part_rels.num_mises = num_mises;
    agg_part_rels.num_owned_mises = num_mises;

    Table *mis_to_dof = new Table;
    mis_to_dof->nrows = num_mises;
    mis_to_dof->ncols = ND;
    mis_to_dof->I.assign((size_t)num_mises + 1, 0);
    for (int m = 0; m < num_mises; ++m)
        mis_to_dof->I[m + 1] = mis_to_dof->I[m] + mis_count[m];
    mis_to_dof->J.resize(ND);
    std::vector<int> pos(mis_to_dof->I.begin(), mis_to_dof->I.end() - 1);
    for (int d = 0; d < ND; ++d)
        mis_to_dof->J[pos[agg_part_rels.mises[d]]++] = d;
    agg_part_rels.mis_to_dof = mis_to_dof;
    agg_part_rels.truemis_to_dof = new Table(*mis_to_dof);

    agg_part_rels.mises_size = new int[std::max(1, num_mises)];
    agg_part_rels.mis_master = new int[std::max(1, num_mises)];
    for (int m = 0; m < num_mises; ++m)
    {
        agg_part_rels.mises_size[m] = mis_to_dof->RowSize(m);
        agg_part_rels.mis_master[m] = 0;
    }
    // amg/src/aggregates.cpp:773-774
    agg_part_rels.mis_to_AE = new Table;
    Mult(*agg_part_rels.mis_to_dof, dof_to_AE, *agg_part_rels.mis_to_AE);
    agg_part_rels.AE_to_mis = new Table;
    Transpose(*agg_part_rels.mis_to_AE, *agg_part_rels.AE_to_mis, agg_part_rels.nparts);
}

void agg_create_partitioning_tables(agg_partitioning_relations_t *agg_part_rels,
                                    int NE, Table *elem_to_dof,
                                    const agg_dof_status_t *bdr_dofs)
{
    // amg/src/aggregates.cpp:1357-1443
    Lap lap;
    agg_part_rels->elem_to_dof = elem_to_dof;
    agg_part_rels->dof_to_elem = new Table;
    Transpose(*elem_to_dof, *agg_part_rels->dof_to_elem, elem_to_dof->ncols);
    agg_part_rels->ND = agg_part_rels->dof_to_elem->Size();
    lap("tables: dof_to_elem");

    agg_part_rels->elem_to_AE = new Table;
    TableFromArray(agg_part_rels->partitioning, NE, agg_part_rels->nparts,
                   *agg_part_rels->elem_to_AE);
    agg_part_rels->AE_to_elem = new Table;
    Transpose(*agg_part_rels->elem_to_AE, *agg_part_rels->AE_to_elem,
              agg_part_rels->nparts);
    agg_part_rels->AE_to_dof = new Table;
    Mult(*agg_part_rels->AE_to_elem, *agg_part_rels->elem_to_dof,
         *agg_part_rels->AE_to_dof);
    agg_part_rels->dof_to_AE = new Table;
    Transpose(*agg_part_rels->AE_to_dof, *agg_part_rels->dof_to_AE,
              agg_part_rels->ND);
    lap("tables: AE tables");
    agg_build_glob_to_AE_id_map(*agg_part_rels);
    lap("tables: glob_to_AE_id_map");
    agg_produce_mises(*agg_part_rels);
    lap("tables: produce_mises");
    agg_construct_agg_flags(*agg_part_rels, bdr_dofs);
    lap("tables: agg_flags");
}

agg_partitioning_relations_t *
agg_create_partitioning_fine(int NE, Table *elem_to_dof, Table *elem_to_elem,
                             int *partitioning, const agg_dof_status_t *bdr_dofs,
                             int *nparts, bool testmesh)
{
    agg_partitioning_relations_t *agg_part_rels = new agg_partitioning_relations_t;
    std::memset(agg_part_rels, 0, sizeof(*agg_part_rels));
    agg_part_rels->testmesh = testmesh;
    SA_ASSERT(0 < NE);
    SA_ASSERT(elem_to_elem);
    agg_part_rels->elem_to_elem = elem_to_elem;
    if (partitioning)
        agg_part_rels->partitioning = partitioning;
    else
        agg_part_rels->partitioning =
            part_generate_partitioning_unweighted(*elem_to_elem, nparts);
    agg_part_rels->nparts = *nparts;
    agg_create_partitioning_tables(agg_part_rels, NE, elem_to_dof, bdr_dofs);
    return agg_part_rels;
}

static Table *coarse_elem_to_elem_product(const agg_partitioning_relations_t &fine)
{
    // elem_to_elem = AE_to_elem * elem_to_elem * elem_to_AE (amg/src/aggregates.cpp:1768-1771)
    Table tmptbl;
    Table *e2e = new Table;
    Mult(*fine.AE_to_elem, *fine.elem_to_elem, tmptbl);
    Mult(tmptbl, *fine.elem_to_AE, *e2e);
    return e2e;
}

static int *coarse_metis_partitioning(const agg_partitioning_relations_t &fine,
                                      const Table &elem_to_elem, int *nparts)
{
    const int num_elem = fine.nparts;
    std::vector<int> weights(num_elem);
    for (int i = 0; i < num_elem; ++i)
        weights[i] = fine.AE_to_dof->RowSize(i);
    // METIS does not accept self loops; the product has them.
    Table graph;
    graph.nrows = graph.ncols = num_elem;
    graph.I.assign((size_t)num_elem + 1, 0);
    for (int i = 0; i < num_elem; ++i)
    {
        const int *row = elem_to_elem.GetRow(i);
        for (int k = 0; k < elem_to_elem.RowSize(i); ++k)
            if (row[k] != i)
                graph.J.push_back(row[k]);
        graph.I[i + 1] = (int)graph.J.size();
    }
    return part_generate_partitioning(graph, weights.data(), nparts);
}

agg_coarse_topology_t agg_coarse_topology(const agg_partitioning_relations_t &fine,
                                          int nparts_target)
{
    agg_coarse_topology_t t;
    t.elem_to_elem = coarse_elem_to_elem_product(fine);
    t.nparts = nparts_target;
    t.partitioning = coarse_metis_partitioning(fine, *t.elem_to_elem, &t.nparts);
    return t;
}

agg_partitioning_relations_t *
agg_create_partitioning_coarse(const agg_partitioning_relations_t &fine,
                               const int *mis_numcoarsedof, int *nparts,
                               bool avoid_ess_bdr_dofs, int *partitioning,
                               Table *coarse_elem_to_elem)
{
    agg_partitioning_relations_t *rels = new agg_partitioning_relations_t;
    std::memset(rels, 0, sizeof(*rels));
    rels->testmesh = fine.testmesh;

    // coarse dof numbering: MIS-major (amg/src/aggregates.cpp:1687-1695)
    const int num_mises = fine.num_mises;
    rels->mis_coarsedofoffsets = new int[(size_t)num_mises + 1];
    int off = 0;
    for (int mis = 0; mis < num_mises; ++mis)
    {
        rels->mis_coarsedofoffsets[mis] = off;
        off += mis_numcoarsedof[mis];
    }
    rels->mis_coarsedofoffsets[num_mises] = off;
    const int NDc = off;
    rels->dof_masterproc = new int[std::max(1, NDc)];
    std::memset(rels->dof_masterproc, 0, sizeof(int) * std::max(1, NDc));

    Lap lap;
    rels->elem_to_elem = coarse_elem_to_elem ? coarse_elem_to_elem : coarse_elem_to_elem_product(fine);
    lap("coarse: elem_to_elem product");

    const int num_elem = fine.nparts;
    if (partitioning)
        rels->partitioning = partitioning;
    else
        rels->partitioning = coarse_metis_partitioning(fine, *rels->elem_to_elem, nparts);
    rels->nparts = *nparts;
    lap("coarse: partitioning");

    // finedof_to_dof = pattern of the tentative prolongator
    // (amg/src/aggregates.cpp:1445-1479 on one process)
    Table finedof_to_dof;
    finedof_to_dof.nrows = fine.ND;
    finedof_to_dof.ncols = NDc;
    finedof_to_dof.I.assign((size_t)fine.ND + 1, 0);
    for (int d = 0; d < fine.ND; ++d)
    {
        const bool ess = avoid_ess_bdr_dofs && agg_is_dof_on_essential_border(fine, d);
        finedof_to_dof.I[d + 1] =
            finedof_to_dof.I[d] + (ess ? 0 : mis_numcoarsedof[fine.mises[d]]);
    }
    finedof_to_dof.J.resize(finedof_to_dof.I[fine.ND]);
#pragma omp parallel for num_threads(sa_host_threads()) schedule(static)
    for (int d = 0; d < fine.ND; ++d)
    {
        int q = finedof_to_dof.I[d];
        const int cnt = finedof_to_dof.I[d + 1] - q;
        const int base = rels->mis_coarsedofoffsets[fine.mises[d]];
        for (int k = 0; k < cnt; ++k)
            finedof_to_dof.J[q + k] = base + k;
    }
    // elem_to_dof = fine AE_to_dof * finedof_to_dof (amg/src/aggregates.cpp:1510-1514)
    Table *elem_to_dof = new Table;
    Mult(*fine.AE_to_dof, finedof_to_dof, *elem_to_dof);
    elem_to_dof->ncols = NDc;
    lap("coarse: elem_to_dof");

    agg_create_partitioning_tables(rels, num_elem, elem_to_dof, NULL);
    SA_ASSERT(rels->ND == NDc);
    return rels;
}

void agg_free_partitioning(agg_partitioning_relations_t *r)
{
    if (!r)
        return;
    delete r->dof_to_elem;
    delete r->dof_to_dof;
    delete r->elem_to_dof;
    delete r->AE_to_dof;
    delete r->dof_to_AE;
    delete r->truemis_to_dof;
    delete r->mis_to_dof;
    delete r->mis_to_AE;
    delete r->AE_to_mis;
    delete[] r->mis_master;
    delete[] r->mises_size;
    delete[] r->mises;
    delete[] r->dof_id_inAE;
    delete[] r->agg_flags;
    delete[] r->partitioning;
    delete r->AE_to_elem;
    delete r->elem_to_AE;
    delete r->elem_to_elem;
    delete[] r->mis_coarsedofoffsets;
    delete[] r->dof_masterproc;
    delete r;
}

} // namespace saamge
