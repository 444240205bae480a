// Driver C API: problem creation, partitioning and generic read access
// (include/saamge_b200_driver.h).  The hierarchy-building entry points live in
// ml.cpp (product) and in oracle/ (checker).
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <type_traits>

#include <cstdlib>

#include "hierarchy.hpp"
#include "part.hpp"

namespace saamge
{
/* coarse-topology prefetch (ml.cpp) */
void sa_topology_prefetch_start(const agg_partitioning_relations_t *rels, int nparts_target);
void sa_topology_prefetch_drop(const agg_partitioning_relations_t *rels);
} // namespace saamge

using namespace saamge;

static double now_s()
{
    return std::chrono::duration<double>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

namespace saamge
{

static void grid_dims_at_level(const sa_problem_t &prob, const sa_drv_params_t &p, int level,
                               int g[3])
{
    // level 0 elements: nx x ny x nz cells; level 1 elements: blocks; ...
    g[0] = prob.fem->nx;
    g[1] = prob.fem->ny;
    g[2] = prob.fem->nz;
    for (int l = 0; l < level; ++l)
        for (int d = 0; d < 3; ++d)
        {
            const int b = (l == 0) ? std::max(1, p.block[d]) : std::max(1, p.coarse_block);
            g[d] = (g[d] + b - 1) / b;
        }
    if (prob.fem->dim == 2)
        g[2] = 1;
}

int *sa_block_coarse_partitioning(const sa_problem_t &prob, const sa_drv_params_t &p,
                                  int level, int num_elem, int *nparts)
{
    int g[3];
    grid_dims_at_level(prob, p, level, g);
    SA_ASSERT(g[0] * g[1] * g[2] == num_elem);
    const int cb = std::max(1, p.coarse_block);
    return part_generate_partitioning_blocks(prob.fem->dim, g[0], g[1], g[2], cb, cb, cb,
                                             nparts);
}

int *sa_prescribed_coarse_partitioning(const sa_problem_t &prob, const sa_drv_params_t &p,
                                       int level, int num_elem, int *nparts)
{
    std::map<int, std::vector<int>>::const_iterator it = prob.coarse_partitions.find(level);
    if (it != prob.coarse_partitions.end())
    {
        SA_ASSERT((int)it->second.size() == num_elem);
        int *out = new int[num_elem];
        int mx = 0;
        for (int i = 0; i < num_elem; ++i)
        {
            out[i] = it->second[i];
            mx = std::max(mx, out[i] + 1);
        }
        *nparts = mx;
        return out;
    }
    if (p.partition_kind == 1)
        return sa_block_coarse_partitioning(prob, p, level, num_elem, nparts);
    return NULL;
}

} // namespace saamge

extern "C" void sa_drv_default_params(sa_drv_params_t *p)
{
    // defaults of the mltest driver (amg/test/mltest/mltest.cpp:332-404)
    std::memset(p, 0, sizeof(*p));
    p->num_levels = 2;
    p->first_elems_per_agg = 256;
    p->elems_per_agg = 256;
    p->first_nu_pro = 0;
    p->nu_pro = 0;
    p->nu_relax = 3;
    p->first_theta = 0.003;
    p->theta = 0.003;
    p->avoid_ess_bdr_dofs = 1;
    p->partition_kind = 0;
    p->block[0] = p->block[1] = p->block[2] = 4;
    p->coarse_block = 2;
    p->testmesh_inject = 0;
    p->smooth_drop_tol = 0.0;
    p->correct_nullspace = 0;
}

extern "C" void *sa_drv_problem_create(int dim, int nx, int ny, int nz, int order,
                                       int coef_kind, double contrast, uint64_t seed)
{
    sa_problem_t *prob = new sa_problem_t;
    const double t0 = now_s();
    prob->fem = fem_generate_structured(dim, nx, ny, nz, order, coef_kind, contrast, seed);
    prob->times["fem"] = now_s() - t0;
    return prob;
}

extern "C" void *sa_drv_problem_create_ex(int dim, int nx, int ny, int nz, int order,
                                          int coef_kind, double contrast, uint64_t seed,
                                          int ess_mask)
{
    sa_problem_t *prob = new sa_problem_t;
    prob->fem =
        fem_generate_structured_ex(dim, nx, ny, nz, order, coef_kind, contrast, seed, ess_mask);
    return prob;
}

extern "C" int sa_drv_problem_partition_array(void *prob_, const int *part, int nparts)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->fem && !prob->rels);
    const fem_problem_t &f = *prob->fem;
    prob->target_nparts0 = nparts;
    int *partitioning = new int[f.NE];
    std::memcpy(partitioning, part, sizeof(int) * f.NE);
    Table *elem_to_dof = new Table(f.elem_to_dof);
    Table *elem_to_elem = new Table(f.elem_to_elem);
    prob->rels = agg_create_partitioning_fine(f.NE, elem_to_dof, elem_to_elem, partitioning,
                                              f.bdr_dofs.data(), &nparts, true);
    return nparts;
}

extern "C" void sa_drv_problem_set_coarse_partition(void *prob_, int level, const int *part, int n)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    prob->coarse_partitions[level].assign(part, part + n);
}

extern "C" void sa_drv_problem_destroy(void *prob_)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    if (!prob)
        return;
    if (prob->rels)
    {
        sa_topology_prefetch_drop(prob->rels);
        // elem_to_dof / elem_to_elem were copies owned by the relations
        agg_free_partitioning(prob->rels);
    }
    delete prob->fem;
    delete prob;
}

extern "C" int sa_drv_problem_partition(void *prob_, const sa_drv_params_t *p)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->fem && !prob->rels);
    const fem_problem_t &f = *prob->fem;
    const double t0 = now_s();
    int nparts = f.NE / p->first_elems_per_agg;
    if (nparts == 0)
        nparts = 1;
    prob->target_nparts0 = nparts;
    int *partitioning = NULL;
    if (p->partition_kind == 1)
        partitioning = part_generate_partitioning_blocks(f.dim, f.nx, f.ny, f.nz, p->block[0],
                                                         p->block[1], p->block[2], &nparts);
    else if (p->partition_kind == 2)
    {
        // METIS on one tile of block[0]^dim elements, replicated over the grid (the
        // toolkit's METIS needs minutes for 128^3 elements / 40k parts; a tile keeps the
        // irregular METIS agglomerate shapes at a fraction of the cost)
        const int t = std::max(1, p->block[0]);
        SA_ASSERT(f.nx % t == 0 && f.ny % t == 0 && (f.dim == 2 || f.nz % t == 0));
        const int tz = (f.dim == 3) ? t : 1;
        const int tn = t * t * tz;
        Table g;
        g.nrows = g.ncols = tn;
        g.I.assign((size_t)tn + 1, 0);
        for (int e = 0; e < tn; ++e)
        {
            const int ex = e % t, ey = (e / t) % t, ez = e / (t * t);
            if (f.dim == 3 && ez > 0) g.J.push_back(e - t * t);
            if (ey > 0) g.J.push_back(e - t);
            if (ex > 0) g.J.push_back(e - 1);
            if (ex < t - 1) g.J.push_back(e + 1);
            if (ey < t - 1) g.J.push_back(e + t);
            if (f.dim == 3 && ez < tz - 1) g.J.push_back(e + t * t);
            g.I[e + 1] = (int)g.J.size();
        }
        int tparts = std::max(1, tn / p->first_elems_per_agg);
        const double tm = now_s();
        int *tp = part_generate_partitioning_unweighted(g, &tparts);
        prob->times["metis"] = now_s() - tm;
        partitioning = new int[f.NE];
        const int gx = f.nx / t, gy = f.ny / t;
        for (int e = 0; e < f.NE; ++e)
        {
            const int ex = e % f.nx, ey = (e / f.nx) % f.ny, ez = e / (f.nx * f.ny);
            const int tile = (ex / t) + gx * ((ey / t) + gy * (ez / t));
            const int loc = (ex % t) + t * ((ey % t) + t * (ez % t));
            partitioning[e] = tile * tparts + tp[loc];
        }
        delete[] tp;
        nparts = tparts * gx * gy * ((f.dim == 3) ? f.nz / t : 1);
    }
    // the relations take ownership of the tables (amg/inc/aggregates.hpp:353-369)
    Table *elem_to_dof = new Table(f.elem_to_dof);
    Table *elem_to_elem = new Table(f.elem_to_elem);
    const double t1 = now_s();
    if (!partitioning)
    {
        partitioning = part_generate_partitioning_unweighted(*elem_to_elem, &nparts);
        prob->times["metis"] = now_s() - t1;
    }
    prob->rels = agg_create_partitioning_fine(f.NE, elem_to_dof, elem_to_elem, partitioning,
                                              f.bdr_dofs.data(), &nparts, false);
    prob->times["partition"] = now_s() - t0;
    // the METIS agglomeration of the first coarse level is a host input as well: start it now
    if (p->num_levels > 2 && p->partition_kind != 1 && !getenv("SA_NO_TOPOLOGY_PREFETCH"))
        sa_topology_prefetch_start(prob->rels, sa_target_nparts(f.NE, *p)[1]);
    return nparts;
}

extern "C" void sa_drv_hier_destroy(void *hier)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    if (!H)
        return;
    if (H->impl && H->impl_free)
        H->impl_free(H->impl);
    if (H->owns_rels)
        for (size_t l = 1; l < H->rels.size(); ++l)
            agg_free_partitioning(H->rels[l]);
    delete H;
}

/* ------------------------------------------------------------------ getters */

template <class T> struct dtype_of;
template <> struct dtype_of<int> { enum { v = 0 }; };
template <> struct dtype_of<int64_t> { enum { v = 1 }; };
template <> struct dtype_of<double> { enum { v = 2 }; };
template <> struct dtype_of<char> { enum { v = 3 }; };

#define RET_VEC(vec)                                                                   \
    do {                                                                               \
        *ptr = (vec).data();                                                           \
        *count = (int64_t)(vec).size();                                                \
        *dtype = dtype_of<typename std::remove_cv<typename std::remove_reference<decltype((vec)[0])>::type>::type>::v; \
        return 0;                                                                      \
    } while (0)
#define RET_ARR(p, n, T)                                                               \
    do {                                                                               \
        *ptr = (p);                                                                    \
        *count = (int64_t)(n);                                                         \
        *dtype = dtype_of<T>::v;                                                       \
        return 0;                                                                      \
    } while (0)

static int get_table(const Table *T, const std::string &f, const void **ptr, int64_t *count,
                     int *dtype)
{
    if (!T)
        return 2;
    if (f == "I")
        RET_VEC(T->I);
    if (f == "J")
        RET_VEC(T->J);
    return 1;
}

static int get_sparse(const SparseMatrix &S, const std::string &f, const void **ptr,
                      int64_t *count, int *dtype)
{
    if (f == "I")
        RET_VEC(S.I);
    if (f == "J")
        RET_VEC(S.J);
    if (f == "A")
        RET_VEC(S.A);
    return 1;
}

static int get_rels(const agg_partitioning_relations_t *r, const std::string &name,
                    const void **ptr, int64_t *count, int *dtype)
{
    if (!r)
        return 2;
    const size_t dot = name.find('.');
    const std::string base = name.substr(0, dot);
    const std::string field = dot == std::string::npos ? "" : name.substr(dot + 1);
    if (base == "elem_to_dof") return get_table(r->elem_to_dof, field, ptr, count, dtype);
    if (base == "dof_to_elem") return get_table(r->dof_to_elem, field, ptr, count, dtype);
    if (base == "AE_to_elem") return get_table(r->AE_to_elem, field, ptr, count, dtype);
    if (base == "elem_to_AE") return get_table(r->elem_to_AE, field, ptr, count, dtype);
    if (base == "elem_to_elem") return get_table(r->elem_to_elem, field, ptr, count, dtype);
    if (base == "AE_to_dof") return get_table(r->AE_to_dof, field, ptr, count, dtype);
    if (base == "dof_to_AE") return get_table(r->dof_to_AE, field, ptr, count, dtype);
    if (base == "mis_to_dof") return get_table(r->mis_to_dof, field, ptr, count, dtype);
    if (base == "mis_to_AE") return get_table(r->mis_to_AE, field, ptr, count, dtype);
    if (base == "AE_to_mis") return get_table(r->AE_to_mis, field, ptr, count, dtype);
    if (name == "partitioning")
        RET_ARR(r->partitioning, r->elem_to_dof->Size(), int);
    if (name == "dof_id_inAE")
        RET_ARR(r->dof_id_inAE, r->dof_to_AE->Size_of_connections(), int);
    if (name == "agg_flags")
        RET_ARR(r->agg_flags, r->ND, char);
    if (name == "mises")
        RET_ARR(r->mises, r->ND, int);
    if (name == "mises_size")
        RET_ARR(r->mises_size, r->num_mises, int);
    if (name == "mis_master")
        RET_ARR(r->mis_master, r->num_mises, int);
    return 1;
}

extern "C" int sa_drv_get(void *obj, const char *name_, int level, const void **ptr,
                          int64_t *count, int *dtype)
{
    const std::string name(name_);
    const int magic = *(const int *)obj;
    if (magic == 0x50524f42)
    {
        sa_problem_t *prob = (sa_problem_t *)obj;
        const fem_problem_t &f = *prob->fem;
        if (name.compare(0, 2, "A.") == 0)
            return get_sparse(f.A, name.substr(2), ptr, count, dtype);
        if (name == "b") RET_VEC(f.b);
        if (name == "coef") RET_VEC(f.coef);
        if (name == "elmat") RET_VEC(f.elmat);
        if (name == "bdr_dofs") RET_VEC(f.bdr_dofs);
        return get_rels(prob->rels, name, ptr, count, dtype);
    }
    if (magic != 0x48494552)
        return 3;
    sa_hierarchy_t *H = (sa_hierarchy_t *)obj;
    if (name == "pcg.brr") RET_VEC(H->pcg.brr);
    if (name == "time_keys")
    {
        // newline separated list of timer names
        static thread_local std::string keys;
        keys.clear();
        for (std::map<std::string, double>::const_iterator it = H->times.begin();
             it != H->times.end(); ++it)
            keys += it->first + "\n";
        RET_ARR(keys.data(), keys.size(), char);
    }
    if (name == "pcg.x") RET_VEC(H->pcg.x);
    if (name.compare(0, 5, "cn_P.") == 0)
        return get_sparse(H->cn_P, name.substr(5), ptr, count, dtype);
    if (name.compare(0, 6, "cn_Ac.") == 0)
        return get_sparse(H->cn_Ac, name.substr(6), ptr, count, dtype);
    if (level < 0)
        return 4;
    if (name == "mis_coarsedofoffsets")
    {
        if (level < 1 || level >= (int)H->rels.size())
            return 4;
        RET_ARR(H->rels[level]->mis_coarsedofoffsets, H->rels[level - 1]->num_mises + 1, int);
    }
    if (level < (int)H->rels.size() && 0 == get_rels(H->rels[level], name, ptr, count, dtype))
        return 0;
    if (level >= (int)H->levels.size())
        return 4;
    const sa_level_results_t &R = H->levels[level];
    if (name == "ae_m") RET_VEC(R.ae_m);
    if (name == "ae_eval_off") RET_VEC(R.ae_eval_off);
    if (name == "evals") RET_VEC(R.evals);
    if (name == "ae_evect_off") RET_VEC(R.ae_evect_off);
    if (name == "evects") RET_VEC(R.evects);
    if (name == "ae_D") RET_VEC(R.ae_D);
    if (name == "mis_numcoarsedof") RET_VEC(R.mis_numcoarsedof);
    if (name == "mis_off") RET_VEC(R.mis_off);
    if (name == "mis_tent") RET_VEC(R.mis_tent);
    if (name == "Dinv_neg") RET_VEC(R.Dinv_neg);
    if (name == "celmat_off") RET_VEC(R.celmat_off);
    if (name == "celmat") RET_VEC(R.celmat);
    if (name.compare(0, 12, "tent_interp.") == 0)
        return get_sparse(R.tent_interp, name.substr(12), ptr, count, dtype);
    if (name.compare(0, 7, "interp.") == 0)
        return get_sparse(R.interp, name.substr(7), ptr, count, dtype);
    if (name.compare(0, 3, "Ac.") == 0)
        return get_sparse(R.Ac, name.substr(3), ptr, count, dtype);
    return 1;
}

extern "C" double sa_drv_get_scalar(void *obj, const char *name_, int level)
{
    const std::string name(name_);
    const int magic = *(const int *)obj;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const std::map<std::string, double> *times = NULL;
    const agg_partitioning_relations_t *rels = NULL;
    if (magic == 0x50524f42)
    {
        sa_problem_t *prob = (sa_problem_t *)obj;
        times = &prob->times;
        rels = prob->rels;
        if (name == "NE") return prob->fem->NE;
        if (name == "ne") return prob->fem->ne;
        if (name == "nnz") return prob->fem->A.NumNonZeroElems();
        if (name == "ND" && !rels) return prob->fem->ND;
    }
    else if (magic == 0x48494552)
    {
        sa_hierarchy_t *H = (sa_hierarchy_t *)obj;
        times = &H->times;
        if (name == "num_levels") return (double)H->rels.size() + 1. - (H->rels.size() > H->levels.size() ? 1. : 0.);
        if (name == "num_coarsenings") return (double)H->levels.size();
        if (name == "num_rels") return (double)H->rels.size();
        if (name == "pcg.iterations") return H->pcg.iterations;
        if (name == "pcg.final_res_norm") return H->pcg.final_res_norm;
        if (level >= 0 && level < (int)H->rels.size())
            rels = H->rels[level];
        if (level >= 0 && level < (int)H->levels.size())
        {
            const sa_level_results_t &R = H->levels[level];
            if (name == "NDc") return R.NDc;
            if (name == "nnz_interp") return R.interp.NumNonZeroElems();
            if (name == "nnz_Ac") return R.Ac.NumNonZeroElems();
        }
    }
    else
        return nan;
    if (name.compare(0, 5, "time.") == 0)
    {
        std::map<std::string, double>::const_iterator it = times->find(name.substr(5));
        return it == times->end() ? nan : it->second;
    }
    if (rels)
    {
        if (name == "ND") return rels->ND;
        if (name == "nparts") return rels->nparts;
        if (name == "num_mises") return rels->num_mises;
    }
    return nan;
}
