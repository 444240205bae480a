// Partitioning relations: the integer data contract between the topology
// stage (host input) and the hot path.  Field names, meaning and ownership
// follow agg_partitioning_relations_t (amg/inc/aggregates.hpp:120-179); the
// construction follows agg_create_partitioning_fine / _tables / _coarse
// (amg/src/aggregates.cpp:1316-1443, 1481-1602, 1735-1832) restricted to one
// process (Dof_TrueDof == identity), with MIS construction done by hashing the
// dof->AE rows instead of the O(#MIS * ND) scan of
// agg_construct_mises_local (amg/src/aggregates.cpp:501-653) -- same numbering.
#ifndef SAAMGE_B200_AGGREGATES_HPP
#define SAAMGE_B200_AGGREGATES_HPP

#include "fem.hpp"
#include "sa_types.hpp"

namespace saamge
{

#define AGG_BETWEEN_AES_FLAG 0x01
#define AGG_ON_ESS_DOMAIN_BORDER_FLAG 0x02
#define AGG_ON_PROC_IFACE_FLAG 0x04
#define AGG_OWNED_FLAG 0x08
#define AGG_ALL_FLAGS                                                          \
    (AGG_BETWEEN_AES_FLAG | AGG_ON_ESS_DOMAIN_BORDER_FLAG |                    \
     AGG_ON_PROC_IFACE_FLAG | AGG_OWNED_FLAG)

#define SA_IS_SET_A_FLAG(var, flag) (((var) & (flag)) != 0)

typedef struct
{
    int ND;            /*!< number of DoFs on this level */
    int nparts;        /*!< number of AEs */
    int *partitioning; /*!< element -> AE */
    Table *dof_to_elem;
    Table *dof_to_dof; /*!< unused (debug only in the reference) */
    Table *elem_to_dof;
    Table *AE_to_elem;
    Table *elem_to_AE;
    Table *elem_to_elem;
    Table *AE_to_dof;
    Table *dof_to_AE;
    int *dof_id_inAE; /*!< local id of a dof inside each AE of its dof_to_AE row */
    agg_dof_status_t *agg_flags;

    Table *truemis_to_dof;
    Table *mis_to_dof;
    int *mis_master;
    Table *mis_to_AE;
    Table *AE_to_mis;
    void *mis_truemis; /*!< (parallel relation; identity on one process) */
    int num_owned_mises;
    int num_mises;
    int *mises;      /*!< dof -> MIS */
    int *mises_size; /*!< MIS -> number of dofs */

    int *mis_coarsedofoffsets; /*!< (coarse rels only) finer MIS -> first coarse dof; size finer num_mises+1 */
    int *dof_masterproc;

    void *Dof_TrueDof; /*!< identity on one process; kept for layout fidelity */
    bool owns_Dof_TrueDof;
    bool testmesh;
} agg_partitioning_relations_t;

static inline bool
agg_is_dof_on_essential_border(const agg_partitioning_relations_t &agg_part_rels,
                               int dof_id)
{
    return SA_IS_SET_A_FLAG(agg_part_rels.agg_flags[dof_id],
                            AGG_ON_ESS_DOMAIN_BORDER_FLAG);
}

/// Position of \a col in row \a elem of \a elem_to_col, or -1 (amg/inc/aggregates.hpp:633-651).
static inline int agg_elem_in_col(int elem, int col, const Table &elem_to_col)
{
    const int *row = elem_to_col.GetRow(elem);
    const int rowsz = elem_to_col.RowSize(elem);
    for (int i = 0; i < rowsz; ++i)
        if (row[i] == col)
            return i;
    return -1;
}

/// Global dof id -> local id inside AE \a part, or negative (amg/inc/aggregates.hpp:653-673).
static inline int
agg_map_id_glob_to_AE(int glob_id, int part,
                      const agg_partitioning_relations_t &agg_part_rels)
{
    const Table &dof_to_AE = *agg_part_rels.dof_to_AE;
    const int order = agg_elem_in_col(glob_id, part, dof_to_AE);
    return (0 > order ? order
                      : agg_part_rels.dof_id_inAE[order + dof_to_AE.GetI()[glob_id]]);
}

void agg_construct_agg_flags(agg_partitioning_relations_t &agg_part_rels,
                             const agg_dof_status_t *bdr_dofs);
void agg_build_glob_to_AE_id_map(agg_partitioning_relations_t &agg_part_rels);

/// MISes by hashing dof->AE rows; numbering = order of the first dof, members
/// ascending (what amg/src/aggregates.cpp:541-607 produces on one process).
void agg_produce_mises(agg_partitioning_relations_t &agg_part_rels);

/*! Fine-level relations (amg/src/aggregates.cpp:1316-1355).  \a elem_to_dof,
    \a elem_to_elem and \a partitioning become owned by the returned struct
    (amg/inc/aggregates.hpp:353-369).  If \a partitioning is NULL, METIS is called
    with *nparts target parts; *nparts returns the actual number. */
agg_partitioning_relations_t *
agg_create_partitioning_fine(int NE, Table *elem_to_dof, Table *elem_to_elem,
                             int *partitioning, const agg_dof_status_t *bdr_dofs,
                             int *nparts, bool testmesh = false);

void agg_create_partitioning_tables(agg_partitioning_relations_t *agg_part_rels,
                                    int NE, Table *elem_to_dof,
                                    const agg_dof_status_t *bdr_dofs);

/*! Coarse-level relations (amg/src/aggregates.cpp:1735-1832).  Coarse elements
    are the fine AEs; coarse dofs are numbered MIS-major from \a mis_numcoarsedof
    (amg/src/aggregates.cpp:1651-1695).  \a avoid_ess_bdr_dofs tells whether
    essential fine dofs carry no coarse dofs (tentative-P rows empty).
    \a partitioning may be given (fixtures) or NULL (METIS with AE-dof-count
    weights, amg/src/aggregates.cpp:1797-1804). */
agg_partitioning_relations_t *
agg_create_partitioning_coarse(const agg_partitioning_relations_t &agg_part_rels_fine,
                               const int *mis_numcoarsedof, int *nparts,
                               bool avoid_ess_bdr_dofs, int *partitioning = NULL,
                               Table *coarse_elem_to_elem = NULL);

/*! The part of agg_create_partitioning_coarse that depends on the fine relations only (not
    on the spectral results): coarse elem_to_elem and, unless given, the METIS partitioning
    of the fine AEs.  It can therefore run on a host thread while the GPU works on the fine
    level; the results are handed to agg_create_partitioning_coarse (which takes ownership). */
struct agg_coarse_topology_t
{
    Table *elem_to_elem;
    int *partitioning;
    int nparts;
};
agg_coarse_topology_t agg_coarse_topology(const agg_partitioning_relations_t &agg_part_rels_fine,
                                          int nparts_target);

void agg_free_partitioning(agg_partitioning_relations_t *agg_part_rels);

} // namespace saamge

#endif
