/* Host-driver C API (ctypes-facing) used by tests/ and bench.py.
 *
 * This is NOT the drop-in boundary (that is include/saamge_b200.h, the C ABI of
 * the CUDA library).  It is a thin handle-based wrapper over the C++ host
 * mirror of the reference's tg_ / ml_ entry points (saamge_b200/host), playing the
 * role of the reference's test drivers (amg/test/mltest/mltest.cpp): generate a
 * problem, partition it, build the hierarchy, run PCG, and read back any
 * intermediate object for comparison with the oracle.
 */
#ifndef SAAMGE_B200_DRIVER_H
#define SAAMGE_B200_DRIVER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Options of the mltest driver that matter on the hot path
 * (amg/test/mltest/mltest.cpp:332-404) + MultilevelParameters (amg/inc/ml.hpp:59-114). */
typedef struct
{
    int num_levels;          /* total levels, >= 2 (--num-levels) */
    int first_elems_per_agg; /* --first-elems-per-agg */
    int elems_per_agg;       /* --elems-per-agg (coarser levels) */
    int first_nu_pro;        /* --first-nu-pro */
    int nu_pro;              /* --nu-pro */
    int nu_relax;            /* --nu-relax, SAS polynomial degree 3*nu+1 */
    double first_theta;      /* --first-theta */
    double theta;            /* --theta */
    int avoid_ess_bdr_dofs;  /* MultilevelParameters::avoid_ess_bdr_dofs (always 1 upstream) */
    int partition_kind;      /* 0 = METIS k-way (reference), 1 = regular blocks (fixtures) */
    int block[3];            /* block edge in elements per direction, finest level (partition_kind 1) */
    int coarse_block;        /* fine AEs per coarse AE per direction (partition_kind 1) */
    int testmesh_inject;     /* 1 = add the all-ones vector on AE 0 (amg/src/interp.cpp:510-524) */
    double smooth_drop_tol;  /* MultilevelParameters::smooth_drop_tol: entries of the smoothed P with
                                |p| <= tol are dropped (AltThreshold, amg/src/interp.cpp:134-170); 0 = off */
    int correct_nullspace;   /* --correct-nulspace: CorrectNullspace coarsest solver (amg/src/solve.cpp:52-164) */
} sa_drv_params_t;

void sa_drv_default_params(sa_drv_params_t *p);

/* ---- problem (input producer) ---- */
void *sa_drv_problem_create(int dim, int nx, int ny, int nz, int order, int coef_kind,
                            double contrast, uint64_t seed);
/* ess_mask: essential sides, bit 0: x=0, 1: x=1, 2: y=0, 3: y=1, 4: z=0, 5: z=1 */
void *sa_drv_problem_create_ex(int dim, int nx, int ny, int nz, int order, int coef_kind,
                               double contrast, uint64_t seed, int ess_mask);
/* fixtures: explicit partitions (finest level / coarsening `level` >= 1) */
int sa_drv_problem_partition_array(void *prob, const int *part, int nparts);
void sa_drv_problem_set_coarse_partition(void *prob, int level, const int *part, int n);
void sa_drv_problem_destroy(void *prob);
/* Fine-level partition + relations; returns number of AEs. */
int sa_drv_problem_partition(void *prob, const sa_drv_params_t *p);

/* ---- hierarchy through the B200 path (sa_gpu_* C ABI underneath) ---- */
/* device: CUDA device ordinal.  ae_shard / ae_nshards: this process handles AEs
 * of shard ae_shard out of ae_nshards in the local spectral stage of the finest
 * level (0,1 = all). */
void *sa_drv_ml_build(void *prob, const sa_drv_params_t *p, int device);
/* Runs PCG with the V-cycle preconditioner on the device; returns iterations
 * (negative on failure, amg/src/mfem_addons.cpp:201,232). */
int sa_drv_ml_pcg(void *hier, int maxiter, double rtol, double atol);
/* Operator update without new eigensolves (adapt_update_operators, amg/src/adapt.cpp:171-216):
   sa_drv_problem_set_A_values replaces the values of the problem's operator (same pattern),
   sa_drv_ml_update_operators redoes the smoothers, (resmooth_interp) the smoothing of the kept
   tentative prolongators, the Galerkin products and the coarsest solver on the device. */
int sa_drv_problem_set_A_values(void *prob, const double *vals, int64_t n);
int sa_drv_ml_update_operators(void *hier, int resmooth_interp);
/* Pull every level's device results into the host record (for comparisons). */
int sa_drv_ml_download(void *hier);
void sa_drv_hier_destroy(void *hier);

/* ---- multi-GPU: shard the AE loop of every level over `world` ranks; `exchange` is
 * called with the device level after the rank computed AEs [begin, end) and must
 * all-gather the per-AE results (sa_gpu_get_spectral -> collective -> sa_gpu_set_spectral) */
typedef void (*sa_drv_exchange_ft)(void *gpu_level, int begin, int end, int nparts);
void sa_drv_set_sharding(int rank, int world, sa_drv_exchange_ft exchange);

/* ---- generic read access (problem, relations, hierarchy records) ----
 * obj   : problem or hierarchy handle
 * name  : e.g. "A.I", "elem_to_dof.J", "AE_to_dof.I", "mises", "evals", "evects",
 *         "tent_interp.I", "Ac.A", "pcg.brr", "timing.<stage>" ...
 * level : hierarchy level (0 = finest); ignored for problem-only arrays
 * dtype : 0 = int32, 1 = int64, 2 = float64, 3 = int8
 * Returns 0 on success, nonzero if the name is unknown. */
int sa_drv_get(void *obj, const char *name, int level, const void **ptr, int64_t *count,
               int *dtype);
/* Scalars by name ("ND", "nparts", "num_mises", "num_levels", "pcg.iterations",
 * "time.<stage>" in seconds, ...).  Returns NaN for unknown names. */
double sa_drv_get_scalar(void *obj, const char *name, int level);

/* ---- bench driver (bench.py): local spectral stage of the finest level ---- */
void *sa_drv_bench_create(void *prob, const sa_drv_params_t *p, int device);
void sa_drv_bench_destroy(void *bench);
/* mode 0: inputs resident on the device, returns device milliseconds (CUDA events);
   mode 1: end to end from host buffers (H2D of all inputs + compute + D2H of results)
   mode 2: as 1 for one rank of a sharded stage (only the inputs / results of its AE range move) */
double sa_drv_bench_step(void *bench, int mode, int ae_begin, int ae_end);
/* the resident sa_gpu_level of the bench (for the sharded stage + device-side exchange) */
void *sa_drv_bench_level(void *bench);
/* "h2d_bytes", "d2h_bytes", "launches", "flops", "bytes", "sum_m", "pinned", "phase.N" */
double sa_drv_bench_scalar(void *bench, const char *name);
/* device stage profile of the process-wide context (see sa_gpu_ctx_profile) */
int sa_drv_gpu_profile(int enable, char *buf, int buflen);
/* the process-wide sa_gpu_ctx */
void *sa_drv_ctx(void);
/* device handles of a hierarchy built by sa_drv_ml_build (sa_gpu_level*, sa_gpu_solver*) */
void *sa_drv_ml_gpu_level(void *hier, int level);
void *sa_drv_ml_gpu_solver(void *hier);
/* kind 0: SpMV with the finest operator, 1: fused smoother step; ms per call */
double sa_drv_ml_spmv_bench(void *hier, int kind, int reps);

#ifdef __cplusplus
}
#endif

#endif
