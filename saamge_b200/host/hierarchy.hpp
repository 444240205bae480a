// Handle types behind the driver C API (include/saamge_b200_driver.h).
#ifndef SAAMGE_B200_HIERARCHY_HPP
#define SAAMGE_B200_HIERARCHY_HPP

#include <map>
#include <string>
#include <vector>

#include "../../include/saamge_b200_driver.h"
#include "aggregates.hpp"
#include "fem.hpp"
#include "level_results.hpp"

namespace saamge
{

struct sa_problem_t
{
    int magic = 0x50524f42; // 'PROB'
    fem_problem_t *fem = NULL;
    agg_partitioning_relations_t *rels = NULL; // finest relations (owned)
    int target_nparts0 = 0;                    // requested number of AEs on the finest level
    // explicit coarse partitions (fixtures): coarsening index -> element -> AE
    std::map<int, std::vector<int>> coarse_partitions;
    std::map<std::string, double> times;
    std::vector<const void *> pinned; // host arrays registered by sa_drv_problem_pin
};

struct sa_hierarchy_t
{
    int magic = 0x48494552; // 'HIER'
    sa_problem_t *prob = NULL;
    sa_drv_params_t params;
    std::vector<agg_partitioning_relations_t *> rels; // rels[0] borrowed from prob
    bool owns_rels = true;                            // rels[1..] freed with the handle
    std::vector<sa_level_results_t> levels;           // one per coarsening
    sa_pcg_results_t pcg;
    // CorrectNullspace level (when built): scaling P and its Galerkin operator ("cn_P.*", "cn_Ac.*")
    SparseMatrix cn_P, cn_Ac;
    std::map<std::string, double> times;
    void *impl = NULL;
    void (*impl_free)(void *) = NULL;
};

/// Target AE counts per coarsening (amg/test/mltest/mltest.cpp:698-722).
static inline std::vector<int> sa_target_nparts(int NE, const sa_drv_params_t &p)
{
    std::vector<int> nparts(p.num_levels - 1);
    nparts[0] = NE / p.first_elems_per_agg;
    if (nparts[0] == 0)
        nparts[0] = 1;
    for (int i = 1; i < p.num_levels - 1; ++i)
    {
        nparts[i] = (int)((double)nparts[i - 1] / (double)p.elems_per_agg + 0.5);
        if (nparts[i] < 1)
            nparts[i] = 1;
    }
    return nparts;
}

/// Coarse partition for partition_kind == 1: groups of coarse_block^dim fine AE
/// blocks (the fine AEs themselves form a regular grid).  Returns new int[].
int *sa_block_coarse_partitioning(const sa_problem_t &prob, const sa_drv_params_t &p,
                                  int level, int num_elem, int *nparts);
/// Fixture / block partition of coarsening \a level if one is prescribed, else NULL (METIS).
int *sa_prescribed_coarse_partitioning(const sa_problem_t &prob, const sa_drv_params_t &p,
                                       int level, int num_elem, int *nparts);

} // namespace saamge

#endif
