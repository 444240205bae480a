// Host-side mirror of the reference's hierarchy-building interface for the hot
// path: same names, argument meaning and ownership rules as the tg_ / ml_ / interp_
// entry points of SAAMGE, on the MFEM-free types of sa_types.hpp.  Every function
// hands its work to the CUDA library through the C ABI of include/saamge_b200.h;
// there is no CPU implementation behind these calls (they fail if no GPU is present).
//
//   reference header            what is mirrored here
//   amg/inc/elmat.hpp:53-170    ElementMatrixProvider + StandardGeometric / DenseArray / ParallelCoarse
//   amg/inc/interp.hpp:54-100   interp_data_t
//   amg/inc/smpr.hpp:89-108     smpr_poly_data_t, smpr_sa_poly_roots, smpr_sas_poly_roots
//   amg/inc/tg_data.hpp:47-83   tg_data_t
//   amg/inc/tg.hpp:428-610      tg_init_data, tg_build_hierarchy, tg_produce_data,
//                               tg_update_coarse_operator, tg_free_data, tg_pcg_run
//   amg/inc/levels.hpp:47-64    levels_level_t, levels_list_t
//   amg/inc/ml.hpp:59-196       MultilevelParameters, ml_data_t, ml_produce_data,
//                               ml_produce_hierarchy_from_level, ml_impose_cycle, ml_free_data
//   amg/inc/solve.hpp:129-181   VCycleSolver
#ifndef SAAMGE_B200_SAAMGE_HPP
#define SAAMGE_B200_SAAMGE_HPP

#include "../../include/saamge_b200.h"
#include "aggregates.hpp"
#include "elmat.hpp"
#include "level_results.hpp"
#include <fstream>

#include "sa_types.hpp"

namespace saamge
{

/* ---- device context (plays the role of proc_init / PROC_COMM, amg/inc/process.hpp:94-98) ---- */
sa_gpu_ctx *proc_gpu_init(int device);
sa_gpu_ctx *proc_gpu_ctx();
void proc_gpu_finalize();
/// SA_ASSERT-style abort with the CUDA library's message when rc != 0
void sa_gpu_check(int rc, const char *what);

/* ---- multi-GPU: AE sharding of the local spectral stage ---- */
/// Called after this rank computed AEs [begin, end): must leave the complete set of
/// per-AE results on the level (sa_gpu_get_spectral / all-gather / sa_gpu_set_spectral).
typedef void (*sa_spectral_exchange_ft)(sa_gpu_level *level, int begin, int end, int nparts);
void sa_set_sharding(int rank, int world, sa_spectral_exchange_ft exchange);
void sa_shard_range(const agg_partitioning_relations_t &rels, int rank, int world, int *begin,
                    int *end);
/// Owner-sharded setup: with a library communicator (sa_gpu_comm_create) installed, the tentative
/// prolongator is built by the MIS owners (sa_gpu_dist_tentative_P), coarse element matrices by
/// the AE owners, smoothing / RAP by row blocks; NULL restores the replicated stages.
void sa_set_sharding_comm(sa_gpu_comm *comm, int rank, int world);
const double *sa_sharding_stats();

/* ---- element-matrix providers ---- */
class ElementMatrixStandardGeometric : public ElementMatrixProvider
{
public:
    /// \a assembled_processor_matrix: BC-eliminated matrix; \a blocks: dense element
    /// matrices (what bf->ComputeElementMatrix returns), block e at blocks[offsets[e]]
    ElementMatrixStandardGeometric(const agg_partitioning_relations_t &agg_part_rels,
                                   const SparseMatrix &assembled_processor_matrix,
                                   const double *blocks, const int64_t *offsets);
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const;
    virtual SparseMatrix *BuildAEStiff(int elno) const;
    virtual const double *DenseBlocks() const { return blocks_; }
    virtual const int64_t *DenseBlockOffsets() const { return offsets_; }
    virtual const SparseMatrix *AssembledMatrix() const { return &A_; }

private:
    const SparseMatrix &A_;
    const double *blocks_;
    const int64_t *offsets_;
};

class ElementMatrixDenseArray : public ElementMatrixProvider
{
public:
    ElementMatrixDenseArray(const agg_partitioning_relations_t &agg_part_rels,
                            const double *blocks, const int64_t *offsets);
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const;
    virtual SparseMatrix *BuildAEStiff(int elno) const;
    virtual const double *DenseBlocks() const { return blocks_; }
    virtual const int64_t *DenseBlockOffsets() const { return offsets_; }

private:
    const double *blocks_;
    const int64_t *offsets_;
};

/// Agglomerate matrices given directly (amg/inc/elmat.hpp:142-151): GetMatrix and BuildAEStiff
/// both return the matrix of agglomerate \a elno (not a copy; the provider owns the matrices).
/// BatchedRelations(): relations whose elements are the agglomerates -- the form in which the
/// device path takes them (one dense element block per agglomerate).
class ElementMatrixArray : public ElementMatrixProvider
{
public:
    ElementMatrixArray(const agg_partitioning_relations_t &agg_part_rels,
                       const std::vector<SparseMatrix *> &elem_matrs);
    virtual ~ElementMatrixArray();
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const;
    virtual SparseMatrix *BuildAEStiff(int elno) const;
    virtual const double *DenseBlocks() const { return blocks_.data(); }
    virtual const int64_t *DenseBlockOffsets() const { return offsets_.data(); }
    const agg_partitioning_relations_t &BatchedRelations() const { return *brels_; }

private:
    std::vector<SparseMatrix *> elem_matrs_;
    std::vector<double> blocks_;
    std::vector<int64_t> offsets_;
    agg_partitioning_relations_t *brels_;
};

struct levels_level_struct;

/// Coarse "element" matrices P_e^T A_AE(e) P_e of the finer level's AEs; they are
/// computed and kept on the device (sa_gpu_coarse_elmats).
class ElementMatrixParallelCoarse : public ElementMatrixProvider
{
public:
    ElementMatrixParallelCoarse(const agg_partitioning_relations_t &agg_part_rels,
                                struct levels_level_struct *level);
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const;
    virtual SparseMatrix *BuildAEStiff(int elno) const;
    struct levels_level_struct *finer_level() const { return level; }

private:
    struct levels_level_struct *level;
    mutable std::vector<double> cache_; // read-back of the device blocks, on demand
};

/* ---- data ---- */
typedef struct
{
    int nparts;
    int nu_pro;
    int interp_smoother_degree;
    double *interp_smoother_roots;
    int times_apply_smoother;
    bool use_arpack; /*!< accepted for API compatibility; the device path is always direct */
    bool scaling_P;  /*!< not supported on the device path (CorrectNullspace is out of scope) */
    int *mis_numcoarsedof;
    int coarse_truedof_offset;
    int num_mises;
    double drop_tol;
    bool testmesh_inject;
    // cut_evects_arr / rhs_matrices_arr / AEs_stiffm / mis_tent_interps of the reference
    // live on the device; interp_download() copies them into a flat host record
} interp_data_t;

typedef struct
{
    int nu;
    int degree;
    const double *roots;
    double weightfirst;
    int degree2;
    const double *roots2;
    double param;
} smpr_poly_data_t;

double *smpr_sa_poly_roots(int &nu, int *degree);
double *smpr_sas_poly_roots(int &nu, int *degree);

class VCycleSolver;

/*! Smoother plug (amg/inc/smpr.hpp:59-60): x <- relax(A, b, x); \a data is whatever the caller
    stored in tg_data_t::smoother_data (the reference passes tg_data->poly_data).  \a A is the
    operator of the level (read back from the device on coarse levels, cached). */
typedef void (*smpr_ft)(const SparseMatrix &A, const Vector &b, Vector &x, void *data);

typedef struct
{
    interp_data_t *interp_data;
    sa_gpu_level *gpu; /*!< device-resident Ac, interp, restr, tent_interp, Dinv_neg */
    bool smooth_interp;
    double theta;
    VCycleSolver *coarse_solver;
    smpr_poly_data_t *poly_data;
    bool use_w_cycle;
    int polynomial_coarse_space;
    bool doing_spectral;
    int tag;
    ElementMatrixProvider *elem_data;
    bool have_Ac;
    /*! tg_data_t::pre_smoother / post_smoother (amg/inc/tg_data.hpp:68-69).  NULL (default) =
        smpr_sym_poly, the SAS polynomial smoother, fused on the device; anything else is called
        on the host by the V-cycle (ml_impose_cycle installs it). */
    smpr_ft pre_smoother, post_smoother;
    void *smoother_data;
    const SparseMatrix *A_host; /*!< operator of the level for user smoothers (finest: the caller's) */
    SparseMatrix *A_host_owned;
} tg_data_t;

typedef struct levels_level_struct
{
    struct levels_level_struct *finer;
    agg_partitioning_relations_t *agg_part_rels;
    tg_data_t *tg_data;
    struct levels_level_struct *coarser;
} levels_level_t;

typedef struct
{
    int num_levels;
    levels_level_t *finest;
    levels_level_t *coarsest;
} levels_list_t;

class MultilevelParameters
{
public:
    MultilevelParameters(int coarsenings, int *nparts_arr, int first_nu_pro, int nu_pro,
                         int nu_relax, double first_theta, double theta,
                         int polynomial_coarse_space, bool use_correct_nullspace,
                         bool use_arpack, bool do_aggregates);
    ~MultilevelParameters();
    int get_num_coarsenings() const { return num_coarsenings; }
    int get_nu_pro(int j) const { return nu_pro[j]; }
    int get_nu_relax(int j) const { return nu_relax[j]; }
    double get_theta(int j) const { return theta[j]; }
    bool get_smooth_interp(int j) const { return (nu_pro[j] > 0); }
    int get_polynomial_coarse_space(int j) const { return polynomial_coarse_space[j]; }
    bool get_use_correct_nullspace() const { return use_correct_nullspace; }
    bool get_use_arpack() const { return use_arpack; }
    bool get_do_aggregates() const { return do_aggregates; }
    int get_nparts(int j) const { return nparts_arr[j]; }
    bool get_avoid_ess_bdr_dofs() const { return avoid_ess_bdr_dofs; }
    double get_smooth_drop_tol() const { return smooth_drop_tol; }
    bool get_coarse_direct() const { return coarse_direct; }
    void set_coarse_direct(bool cd) { coarse_direct = cd; }
    void set_smooth_drop_tol(double tol) { smooth_drop_tol = tol; }
    /* fixtures: fixed coarse partitions instead of METIS (cf. the hard-coded
       partitions of the mltest fixture, amg/src/aggregates.cpp:1777-1794) */
    typedef int *(*coarse_partition_ft)(int level, int num_elem, int *nparts, void *data);
    void set_coarse_partitioner(coarse_partition_ft f, void *data)
    {
        coarse_partitioner = f;
        coarse_partitioner_data = data;
    }
    coarse_partition_ft coarse_partitioner;
    void *coarse_partitioner_data;
    bool testmesh_inject;

private:
    int num_coarsenings;
    int *nparts_arr;
    int *nu_pro;
    int *nu_relax;
    double *theta;
    int *polynomial_coarse_space;
    bool use_correct_nullspace;
    bool use_arpack;
    bool do_aggregates;
    bool avoid_ess_bdr_dofs;
    bool coarse_direct;
    double smooth_drop_tol;
};

typedef struct
{
    levels_list_t levels_list;
    sa_gpu_solver *gpu_solver; /*!< V-cycle chain on the device (ml_impose_cycle) */
    int nu_relax;
    /*! CorrectNullspace (amg/src/solve.cpp:52-164; ml_produce_hierarchy_from_level,
        amg/src/ml.cpp:225-235): the level below the coarsest spectral one, prolongator =
        scaling P (tg_data_t::scaling_P of the reference), NULL when not used. */
    sa_gpu_level *correct_nullspace_level;
    SparseMatrix *scaling_P; /*!< host copy of the scaling P (interp_scaling_P_assemble) */
} ml_data_t;

/* ---- the local eigensolver as a class (amg/inc/spectral.hpp:91-224) ---- */
/*! Same shape as the reference's Eigensolver: Solve() computes, for the given AE matrix \a A, the
    weighted-l1 diagonal \a B (allocated when NULL; the caller frees, amg/inc/spectral.hpp:147-150)
    and the eigenvectors of A z = lambda B z with lambda <= theta (at least one), appended as the
    columns of \a cut_evects, normalised z^T B z = 1.  The hierarchy builder does not call it AE by
    AE (it batches all AEs of a level in sa_gpu_local_spectral); Solve() runs that same device
    path on a level of one AE whose single element matrix is \a A.  Returns true when a vector
    was added. */
class Eigensolver
{
public:
    Eigensolver(const int *aggregates, const agg_partitioning_relations_t &agg_part_rels,
                int threshold = 0x7fffffff);
    virtual ~Eigensolver() {}
    virtual bool Solve(const SparseMatrix &A, SparseMatrix *&B, int part, int agg_id,
                       int aggregate_size, double &theta, DenseMatrix &cut_evects);
    void GetStatistics(int &o_count_solves, int &o_count_direct_solves, int &o_count_max_used,
                       double &o_smallest_eigenvalue_skipped);
    /// eigenvalues of the last Solve (ascending)
    const Vector &LastEigenvalues() const { return last_evals; }

private:
    const agg_partitioning_relations_t &agg_part_rels;
    int threshold;
    int count_solves, count_direct_solves, count_max_used;
    double smallest_eigenvalue_skipped;
    Vector last_evals;
};

/* ---- two-grid entry points ---- */
tg_data_t *tg_init_data(const SparseMatrix *A, const agg_partitioning_relations_t &agg_part_rels,
                        int nu_pro, int nu_relax, double theta, bool smooth_interp,
                        double smooth_drop_tol, bool use_arpack);
/// \a Ag == NULL: the operator is the coarse operator the finer level left on the
/// device.  \a finer: finer level's tg_data (NULL on the finest level).
void tg_build_hierarchy(const SparseMatrix *Ag, tg_data_t &tg_data,
                        const agg_partitioning_relations_t &agg_part_rels,
                        ElementMatrixProvider *elem_data, bool avoid_ess_bdr_dofs,
                        tg_data_t *finer = NULL);
tg_data_t *tg_produce_data(const SparseMatrix &Ag,
                           const agg_partitioning_relations_t &agg_part_rels, int nu_pro,
                           int nu_relax, ElementMatrixProvider *elem_data, double theta,
                           bool smooth_interp, int polynomial_coarse_arg, bool use_arpack,
                           bool avoid_ess_bdr_dofs);
void tg_update_coarse_operator(tg_data_t *tg_data, bool perform_solve_init, bool coarse_direct);
void tg_free_data(tg_data_t *tg_data);

/* the stages tg_build_hierarchy runs, exposed like in the reference */
void interp_compute_vectors(const agg_partitioning_relations_t &agg_part_rels,
                            const interp_data_t &interp_data, tg_data_t &tg_data, double &theta);
void interp_sparse_tent_assemble(const agg_partitioning_relations_t &agg_part_rels,
                                 interp_data_t &interp_data, tg_data_t &tg_data,
                                 bool avoid_ess_bdr_dofs);

/* ---- algebraic (matrix-only) entry: algebraic.cpp; amg/src/tg.cpp:580-668, 862-886,
        amg/src/fem.cpp:720-762 ---- */
void ExtractSubMatrices(const SparseMatrix &A, const agg_partitioning_relations_t &agg_part_rels,
                        std::vector<SparseMatrix *> &agglomerate_element_matrices);
agg_partitioning_relations_t *fem_create_partitioning_from_matrix(const SparseMatrix &A, int *nparts,
                                                                   const std::vector<int> &isolated_cells);
tg_data_t *tg_produce_data_algebraic(const SparseMatrix &Alocal, const SparseMatrix &Ag,
                                     const agg_partitioning_relations_t &agg_part_rels, int nu_pro,
                                     int nu_relax, double spectral_tol, bool smooth_interp,
                                     int polynomial_coarse_arg, bool use_window, bool use_arpack,
                                     bool avoid_ess_bdr_dofs);
SparseMatrix *ReadHypreMat(const char *filename);

/* ---- multilevel entry points ---- */
ml_data_t *ml_produce_data(const SparseMatrix &Ag, agg_partitioning_relations_t *agg_part_rels,
                           ElementMatrixProvider *elem_data_finest,
                           const MultilevelParameters &mlp);
ml_data_t *ml_produce_data_algebraic(const SparseMatrix &Ag, const agg_partitioning_relations_t &agg_part_rels,
                                     const MultilevelParameters &mlp);
void ml_produce_hierarchy_from_level(int coarsenings, int starting_level, ml_data_t &ml_data,
                                     const MultilevelParameters &mlp);
void ml_impose_cycle(ml_data_t &ml_data, bool Wcycle);
/// interp_scaling_P_assemble (amg/src/interp.cpp:842-909) from the MIS bases of the coarsest
/// spectral level (local_coarse_one_representation, amg/src/contrib.cpp:655-668: per MIS the
/// least-squares representation of the constant vector in the basis, normalised) and the
/// CorrectNullspace level built on it (amg/src/solve.cpp:52-110).  Called by ml_produce_data when
/// MultilevelParameters::use_correct_nullspace is set; ml_impose_cycle then appends the level.
void ml_build_correct_nullspace(ml_data_t &ml_data);

/* ---- operator update without new eigensolves (amg/src/adapt.cpp:171-216, amg/inc/tg.hpp:678-693,
   735; amg/inc/smpr.hpp:241) ---- */
void smpr_update_Dinv_neg(tg_data_t &tg_data);
void tg_smooth_interp(tg_data_t &tg_data);
void tg_free_coarse_operator(tg_data_t &tg_data);
/// \a A: the new finest operator (same pattern), or NULL on a coarse level whose operator is the
/// finer level's Ac.
void adapt_update_operators(const SparseMatrix *A, tg_data_t &tg_data, bool resmooth_interp);
void adapt_update_operators(const SparseMatrix &A, ml_data_t &ml_data, const MultilevelParameters &mlp,
                            bool resmooth_interp);

/* ---- binary formats of the reference's dumps (amg/src/mbox.cpp:310-483; amg/inc/mbox.hpp:344-516):
   caller owns what the read functions return ---- */
Table *mbox_read_table(const char *filename);
void mbox_write_table(const char *filename, const Table &tbl);
SparseMatrix *mbox_read_sparse_matr(const char *filename);
SparseMatrix *mbox_read_sparse_matr(std::ifstream &ispm);
void mbox_write_sparse_matr(const char *filename, const SparseMatrix &spm);
void mbox_write_sparse_matr(std::ofstream &ospm, const SparseMatrix &spm);
DenseMatrix *mbox_read_dense_matr(const char *filename);
DenseMatrix *mbox_read_dense_matr(std::ifstream &idem);
void mbox_write_dense_matr(const char *filename, const DenseMatrix &dem);
void mbox_write_dense_matr(std::ofstream &odem, const DenseMatrix &dem);
SparseMatrix **mbox_read_sparse_matr_arr(const char *filename, int *n);
void mbox_write_sparse_matr_arr(const char *filename, SparseMatrix **arr, int n);
DenseMatrix **mbox_read_dense_matr_arr(const char *filename, int *n);
void mbox_write_dense_matr_arr(const char *filename, DenseMatrix **arr, int n);
void ml_free_data(ml_data_t *ml_data);
levels_level_t *levels_list_get_level(const levels_list_t &list, int i);

/* ---- solve ---- */
class VCycleSolver
{
public:
    VCycleSolver(ml_data_t *ml_data) : ml_data(ml_data) {}
    /// one V-cycle from x = 0 (amg/src/solve.cpp:309-323)
    void Mult(const Vector &b, Vector &x) const;

private:
    ml_data_t *ml_data;
};

/// kalchev_pcg (amg/src/mfem_addons.cpp:106-248) with the hierarchy's V-cycle as B.
/// Returns the iteration count (negative on failure).
int kalchev_pcg(ml_data_t &ml_data, const Vector &b, Vector &x, int print_iter,
                int max_num_iter, double RTOLERANCE, double ATOLERANCE,
                std::vector<double> *brr_history = NULL);

/// Copies one level's device results into the flat host record (tests, file dumps).
void tg_download_results(const tg_data_t &tg_data, const agg_partitioning_relations_t &rels,
                         tg_data_t *coarser, sa_level_results_t &R);

/* pipelined upload of the finest level (see ml.cpp) */
void sa_set_async_finest_upload(bool on);

/* coarse-topology prefetch (ml.cpp) */
void sa_topology_prefetch_start(const agg_partitioning_relations_t *rels, int nparts_target);
void sa_topology_prefetch_drop(const agg_partitioning_relations_t *rels);

} // namespace saamge

#endif
