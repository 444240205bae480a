/* saamge_b200 -- C ABI of the B200 (sm_100a) implementation of SAAMGE's spectral
 * AMGe hot path.  This is the drop-in boundary: plain pointers and sizes, no C++
 * or torch types.  A maintainer of the reference binds these entry points from
 * the reference's own C++ at the call sites cited with each function (see
 * INTEGRATION.md for the stubs).  All file:line citations are relative to the
 * reference tree (amg/...).
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on failure; sa_gpu_last_error()
 *     gives the message.  The reference has no error codes (SA_ASSERT aborts,
 *     amg/inc/common.hpp:635-647); the C++ shim maps nonzero to that behaviour.
 *   - host pointers in, host pointers out, caller allocated; nothing passed in is
 *     retained after the call returns (data is copied to the device).
 *   - index data int32, floating point FP64, dense blocks column-major
 *     (mfem::Table / SparseMatrix / DenseMatrix layouts).
 *   - one context per GPU; a context and the levels created from it are not
 *     re-entrant (the reference is single threaded per rank as well).
 */
#ifndef SAAMGE_B200_H
#define SAAMGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sa_gpu_ctx sa_gpu_ctx;
typedef struct sa_gpu_level sa_gpu_level;
typedef struct sa_gpu_solver sa_gpu_solver;

/* ---- context ---- */
int sa_gpu_ctx_create(int device, sa_gpu_ctx **ctx);
void sa_gpu_ctx_destroy(sa_gpu_ctx *ctx);
const char *sa_gpu_last_error(void);
/* cudaStream_t of the context as an opaque pointer (for event timing in bench.py) */
/* Hands the unused blocks of the stream-ordered device memory pool back to the driver.  Worth
   calling between unrelated large jobs in one process: a pool fragmented by many blocks of
   other sizes can take seconds to serve a fresh multi-GB request. */
int sa_gpu_ctx_trim_pool(sa_gpu_ctx *ctx);
void *sa_gpu_ctx_stream(sa_gpu_ctx *ctx);
int sa_gpu_ctx_sync(sa_gpu_ctx *ctx);
/* number of kernel launches issued through this context so far */
int64_t sa_gpu_ctx_launch_count(sa_gpu_ctx *ctx);
/* ms between two points: call with begin!=0 to record the start event, then with
   begin==0 to record the end event, synchronise and return elapsed ms (CUDA events
   on the context's stream) */
double sa_gpu_ctx_timer(sa_gpu_ctx *ctx, int begin);

/* per-stage device timings (CUDA events around each kernel group).  enable: 1/0/-1
   (-1 = unchanged; also switched on by the environment variable SA_GPU_PROFILE=1);
   buf receives "name milliseconds" lines accumulated so far and the table is reset */
int sa_gpu_ctx_profile(sa_gpu_ctx *ctx, int enable, char *buf, int buflen);

/* ---- one level: the integer data contract agg_partitioning_relations_t
 *      (amg/inc/aggregates.hpp:120-179) + operator + element matrices ---- */
typedef struct
{
    int ND;        /* dofs */
    int NE;        /* elements (fine elements, or finer AEs on coarse levels) */
    int nparts;    /* AEs */
    int num_mises; /* MISes */
    const int *elem_to_dof_I, *elem_to_dof_J;
    const int *dof_to_elem_I, *dof_to_elem_J; /* rows ascending */
    const int *AE_to_elem_I, *AE_to_elem_J;   /* rows ascending */
    const int *AE_to_dof_I, *AE_to_dof_J;     /* local dof order inside each AE */
    const int *dof_to_AE_I, *dof_to_AE_J;
    const int *dof_id_inAE;                   /* parallel to dof_to_AE_J */
    const int *partitioning;                  /* element -> AE */
    const char *agg_flags;                    /* AGG_* bit flags per dof */
    const int *mis_to_dof_I, *mis_to_dof_J;   /* rows ascending */
    const int *mis_to_AE_I, *mis_to_AE_J;
    const int *AE_to_mis_I, *AE_to_mis_J;     /* rows ascending */
    const int *mises;                         /* dof -> MIS */
    /* operator of this level (CSR).  NULL => use the coarse operator Ac that
       sa_gpu_rap left on the device in the finer level. */
    const int *A_I, *A_J;
    const double *A_data;
    /* dense element blocks (ElementMatrixDenseArray / what
       ElementMatrixStandardGeometric::GetMatrix returns, amg/src/elmat.cpp:68-88,
       242-253), element e at elmat[elmat_off[e]], size RowSize(e)^2, column-major.
       NULL => use the blocks sa_gpu_coarse_elmats left on the device in the finer
       level (ElementMatrixParallelCoarse, amg/src/elmat.cpp:105-195). */
    const double *elmat;
    const int64_t *elmat_off; /* NE+1 */
    /* 1: agg_build_AE_stiffm_with_global (amg/src/aggregates.cpp:855-945,
          bdr_cond_imposed = assemble_ess_diag = true, amg/src/elmat.cpp:51-52)
       0: agg_build_AE_stiffm (amg/src/aggregates.cpp:959-1086) */
    int assemble_with_global;
    /* (coarse levels) finer MIS -> first coarse dof, finer num_mises+1 entries;
       needed by sa_gpu_coarse_elmats */
    const int *mis_coarsedofoffsets;
    /* 1: pipelined upload.  sa_gpu_level_create returns after queueing the tables; the large
          arrays (operator, element blocks) are copied on a separate copy stream in the order
          sa_gpu_local_spectral needs them, which starts on the first agglomerates as soon as
          the rows / blocks they read have arrived.  The caller keeps every host array of this description valid
          and unmodified until the first call that uses the level returns (or until
          sa_gpu_level_upload_wait).  Host arrays should be pinned (sa_gpu_host_register);
          pageable memory works but does not overlap.
       2: as 1, and the operator rows / element blocks that sa_gpu_local_spectral(level, theta,
          ae_begin, ae_end) does not read stay on the host until another entry point needs them
          (a rank of a sharded stage uploads the inputs of its own AEs only).
       0: all copies are complete when sa_gpu_level_create returns. */
    int async_upload;
} sa_gpu_level_desc;

/* sizeof(sa_gpu_level_desc) as compiled into the library (lets FFI bindings that mirror
   the struct by hand check their layout). */
size_t sa_gpu_level_desc_size(void);

/* Copies the description to the device.  \a finer may be NULL (finest level). */
int sa_gpu_level_create(sa_gpu_ctx *ctx, const sa_gpu_level_desc *desc, sa_gpu_level *finer,
                        sa_gpu_level **level);
void sa_gpu_level_destroy(sa_gpu_level *level);
/* Returns the cached work arrays of the local spectral stage (reflector block, inverse-
   iteration workspace, ...; they belong to the level's context and are shared by its
   levels) to the device memory pool.  Results are kept; a later sa_gpu_local_spectral
   simply allocates them again. */
int sa_gpu_level_trim(sa_gpu_level *level);
/* operator + element-block bytes a pipelined upload has queued so far */
double sa_gpu_level_uploaded_bytes(sa_gpu_level *level);
/* Blocks until a pipelined upload (desc.async_upload) has finished; no-op otherwise. */
int sa_gpu_level_upload_wait(sa_gpu_level *level);

/* ---- setup: local spectral stage ---- */

/* Replaces the loop of interp_compute_vectors (amg/src/interp.cpp:387-556):
 * for every AE in [ae_begin, ae_end): BuildAEStiff (a2/a3), weighted-l1 D
 * (mbox_snd_D_sparse_from_sparse, amg/src/mbox.cpp:913-949), and the eigenpairs of
 * A z = lambda D z with lambda in (-1, theta] that the reference obtains from
 * dsygvx (xpacks_calc_lower_eigens_dense, amg/src/xpacks.cpp:222-314; at least one
 * pair is always returned).  inject_ones_ae0 reproduces the mltest fixture
 * (amg/src/interp.cpp:510-524).  Results stay on the device. */
int sa_gpu_local_spectral(sa_gpu_level *level, double theta, int ae_begin, int ae_end,
                          int inject_ones_ae0);
/* sizes (dofs) of the AEs of the level, nparts ints */
int sa_gpu_get_AE_sizes(sa_gpu_level *level, int *ae_n);
/* number of accepted vectors per AE (cut_evects_arr[i]->Width()), nparts ints */
int sa_gpu_get_spectral_counts(sa_gpu_level *level, int *ae_m);
/* evals: sum(m_i) doubles (AE-major, ascending per AE; an injected vector has no
 * eigenvalue), evects: sum(n_i*m_i) doubles (cut_evects_arr[i], column-major),
 * D: sum(n_i) doubles (rhs_matrices_arr[i] diagonal).  Any pointer may be NULL. */
int sa_gpu_get_spectral(sa_gpu_level *level, double *evals, double *evects, double *D);
/* Overwrites the device-resident eigenvectors of AEs [ae_begin, ae_end) (used when
 * the AE loop was sharded over several GPUs and the pieces were exchanged by the
 * caller).  Layout as returned by sa_gpu_get_spectral. */
int sa_gpu_set_spectral(sa_gpu_level *level, int ae_begin, int ae_end, const int *ae_m,
                        const double *evals, const double *evects, const double *D);
/* Guard-band report of the two threshold decisions of the setup (SURVEY hard part 2): number of
 * AEs (of the last sa_gpu_local_spectral range) with an eigenvalue within 1e-12 of theta, i.e.
 * whose m = #{lambda <= theta} (amg/src/xpacks.cpp:233-234) another equally accurate eigensolver
 * could count differently, and number of MISes (last sa_gpu_tentative_P) with a singular value
 * within a factor 10 of the rank cut 1e-10 sigma_0 (amg/src/xpacks.cpp:609-610).  The decisions
 * themselves are taken exactly as the reference takes them; this only counts the fragile ones. */
int sa_gpu_get_borderline(sa_gpu_level *level, int *theta_borderline, int *rank_borderline);
/* Device-side exchange for a local spectral stage sharded over several GPUs (the reference
 * distributes the AEs over MPI ranks, amg/src/interp.cpp:387): after
 * sa_gpu_local_spectral(level, theta, ae_begin, ae_end) the caller announces the counts of ALL
 * AEs (ae_m_full, nparts ints, gathered from the ranks); the level's result arrays are laid out
 * for the full set with this rank's slice already in place, and their device addresses are
 * returned so that the caller's collective (NCCL broadcast / all-gather over NVLink) can write
 * the other ranks' slices straight into them -- no host bounce.  Layout as sa_gpu_get_spectral:
 * evals offsets = prefix sums of ae_m, evects offsets = prefix sums of n_i * m_i, D offsets =
 * AE_to_dof.I.  The arrays are complete once every slice has arrived (caller synchronises). */
int sa_gpu_spectral_gather_begin(sa_gpu_level *level, int ae_begin, int ae_end, const int *ae_m_full,
                                 double **d_evals, double **d_evects, double **d_D);
/* assembled dense AE matrix (n x n column-major) of one AE -- the value of
 * ElementMatrixProvider::BuildAEStiff(part) (amg/inc/elmat.hpp:71) */
int sa_gpu_build_AE_stiff(sa_gpu_level *level, int part, double *dense_out);

/* ---- setup: tentative prolongator ---- */

/* Replaces ContribTent::contrib_mises (amg/src/contrib.cpp:699-714) +
 * contrib_tent_finalize (:73-95): per MIS restrict (agg_restrict_to_agg_enforce,
 * amg/src/aggregates.cpp:1143-1179), boundary filter (:102-163), column
 * normalisation + thin SVD + rank cut sigma_i > 1e-10 sigma_0 (xpack_svd_dense_arr,
 * xpack_orth_set, amg/src/xpacks.cpp:494-620), insertion (:170-194).
 * mis_numcoarsedof: num_mises ints out.  *NDc: number of coarse dofs. */
int sa_gpu_tentative_P(sa_gpu_level *level, int avoid_ess_bdr_dofs, int *mis_numcoarsedof,
                       int *NDc);
/* mis_tent_interps (amg/inc/interp.hpp:93): block of MIS i is s_i x k_i
 * column-major; total size sum(s_i*k_i). */
int sa_gpu_get_mis_tent(sa_gpu_level *level, double *mis_tent);

/* ---- setup: operators ---- */

/* mbox_build_Dinv_neg_parallel_matrix (amg/src/mbox.cpp:1839-1861) */
int sa_gpu_build_Dinv_neg(sa_gpu_level *level);
/* interp_smooth (amg/src/interp.cpp:172-229): P = prod_k (I + roots[k]^-1 Dinv_neg A) P_tent,
 * then restr = P^T (tg_smooth_interp, amg/inc/tg.hpp:678-693).  degree 0 => P = P_tent. */
int sa_gpu_smooth_P(sa_gpu_level *level, int degree, const double *roots);
/* tg_coarse_matr (amg/inc/tg.hpp:695-709): Ac = P^T A P */
/* CorrectNullspace (amg/src/solve.cpp:52-164): a level whose prolongator is given -- the
   "scaling P" (amg/src/interp.cpp:842-909; one column per MIS of the last spectral level: the
   coarse representation of the constant vector, amg/src/contrib.cpp:655-668).  Its operator is
   finer's Ac; complete it with sa_gpu_build_Dinv_neg and sa_gpu_rap and pass it to
   sa_gpu_solver_create as the last level (destroy with sa_gpu_level_destroy). */
int sa_gpu_level_create_from_P(sa_gpu_ctx *ctx, sa_gpu_level *finer, int cols, const int *P_I,
                               const int *P_J, const double *P_A, sa_gpu_level **out);
/* adapt_update_operators (amg/src/adapt.cpp:171-216): new operator values (same pattern) for a
   level that owns its operator; afterwards sa_gpu_build_Dinv_neg (smpr_update_Dinv_neg),
   optionally sa_gpu_smooth_P (tg_smooth_interp, from the kept tentative P) and sa_gpu_rap. */
int sa_gpu_level_update_operator(sa_gpu_level *level, const double *A_data);
/* AltThreshold (amg/src/interp.cpp:86-170), the drop_tol branch of interp_smooth
   (amg/src/interp.cpp:219-228): keeps the entries of the smoothed prolongator with
   |p_ij| > drop_tol and rebuilds R = P^T.  Call between sa_gpu_smooth_P and sa_gpu_rap. */
int sa_gpu_threshold_P(sa_gpu_level *level, double drop_tol, int *nnz_before, int *nnz_after);
int sa_gpu_rap(sa_gpu_level *level);
/* ElementMatrixParallelCoarse::GetMatrix for every finer AE (amg/src/elmat.cpp:105-195):
 * coarse element matrices P_e^T A_AE(e) P_e, kept on the device of \a coarse. */
int sa_gpu_coarse_elmats(sa_gpu_level *finer, sa_gpu_level *coarse);
/* ElementMatrixProvider::GetMatrix(elno) of the level's provider (amg/inc/elmat.hpp:62): the
 * dense block of one element (ne x ne column-major; on coarse levels the block
 * sa_gpu_coarse_elmats produced).  out may be NULL to query *ne_out only. */
int sa_gpu_get_element_matrix(sa_gpu_level *level, int elno, double *out, int *ne_out);
/* read back: nc_e^2 doubles per finer AE, column-major, concatenated */
int sa_gpu_get_coarse_elmats(sa_gpu_level *coarse, double *celmat);

/* CSR read-back.  which: 0 = A, 1 = tentative P, 2 = P, 3 = restr (P^T), 4 = Ac */
enum { SA_GPU_MAT_A = 0, SA_GPU_MAT_PTENT = 1, SA_GPU_MAT_P = 2, SA_GPU_MAT_R = 3, SA_GPU_MAT_AC = 4 };
int sa_gpu_get_csr_sizes(sa_gpu_level *level, int which, int *rows, int *cols, int *nnz);
int sa_gpu_get_csr(sa_gpu_level *level, int which, int *I, int *J, double *data);
int sa_gpu_get_Dinv_neg(sa_gpu_level *level, double *dinv_neg);

/* ---- solve ---- */

/* y = M x with M one of the level's matrices (HypreParMatrix::Mult); host buffers */
int sa_gpu_spmv(sa_gpu_level *level, int which, const double *x, double *y);
/* smpr_compute_poly (amg/inc/smpr.hpp:319-339): for i < degree:
 *   x += roots[i]^-1 * Dinv_neg .* (A x - b);  host buffers, x in/out */
int sa_gpu_poly_smooth(sa_gpu_level *level, const double *b, double *x, int degree,
                       const double *roots);

/* Chains the levels into a V-cycle (ml_impose_cycle, amg/src/ml.cpp:361-377) with the
 * SAS polynomial smoother of degree 3*nu_relax+1 (smpr_sas_poly_roots,
 * amg/src/smpr.cpp:282-306; sa_gpu_build_Dinv_neg must have been called on every
 * level) and an exact coarsest solve (dense Cholesky of the last Ac; the
 * reference's --coarse-direct option, amg/src/tg.cpp:991-998). */
int sa_gpu_solver_create(sa_gpu_ctx *ctx, sa_gpu_level **levels, int nlevels, int nu_relax,
                         sa_gpu_solver **solver);
void sa_gpu_solver_destroy(sa_gpu_solver *solver);
/* The smoother plug smpr_ft (amg/inc/smpr.hpp:59-60; tg_data_t::pre_smoother / post_smoother,
 * amg/inc/tg_data.hpp:68-69, amg/src/tg.cpp:48-57, 411-414): user relaxation of level `level`
 * of the V-cycle, x <- relax(A_level, b, x), called with HOST vectors (n = dofs of the level).
 * NULL keeps the device SAS polynomial smoother.  A user smoother makes the cycle leave the
 * device twice per level and cycle; it exists for interface completeness. */
typedef void (*sa_gpu_smoother_ft)(int level, int n, const double *b, double *x, void *data);
int sa_gpu_solver_set_smoothers(sa_gpu_solver *solver, int level, sa_gpu_smoother_ft pre,
                                sa_gpu_smoother_ft post, void *data);
/* VCycleSolver::Mult (amg/src/solve.cpp:309-323) = tg_cycle_atb from x = 0
 * (amg/src/tg.cpp:91-132); host buffers */
int sa_gpu_vcycle(sa_gpu_solver *solver, const double *b, double *x);
/* kalchev_pcg (amg/src/mfem_addons.cpp:106-248, zero_rhs = false) with the V-cycle as
 * preconditioner.  x in/out (initial guess).  *iters follows the reference's
 * convention (negative = no convergence / SPD breakdown).  brr_hist (optional)
 * receives (B r, r) per iteration, at most hist_cap entries; *hist_len entries written. */
int sa_gpu_pcg(sa_gpu_solver *solver, const double *b, double *x, int maxiter, double rtol,
               double atol, int *iters, double *brr_hist, int hist_cap, int *hist_len);
/* device-resident variant used for kernel-only timing: b and x already on the
 * device (allocated by sa_gpu_solver_upload) */
int sa_gpu_solver_upload(sa_gpu_solver *solver, const double *b, const double *x0);
int sa_gpu_pcg_resident(sa_gpu_solver *solver, int maxiter, double rtol, double atol,
                        int *iters);
int sa_gpu_solver_download(sa_gpu_solver *solver, double *x);

/* page-lock / unlock a caller-owned host buffer so that the copies made by
   sa_gpu_level_create run at full PCIe speed (optional) */
int sa_gpu_host_register(const void *p, size_t bytes);
int sa_gpu_host_unregister(const void *p);
/* cycle counters of the four phases of the assemble+tridiagonalise kernel summed over
   thread blocks since the last call (diagnostics) */
int sa_gpu_debug_phase_clocks(double *out8);

/* Diagnostics of the large-AE eigensolver (two-stage tridiagonalisation, twostage.cu): reduces
   the dense symmetric n x n matrix A (host, column-major) to tridiagonal form (d_out, e_out;
   T_out: band + reflectors, tau1_out) and keeps the reflectors on the device;
   sa_gpu_debug_twostage_back then maps nvec eigenvectors of that tridiagonal matrix (Y, n x nvec,
   in/out) back to eigenvectors of A.  Any output pointer may be NULL. */
int sa_gpu_debug_twostage(sa_gpu_ctx *ctx, int n, const double *A, double *T_out, double *tau1_out,
                          double *d_out, double *e_out);
int sa_gpu_debug_twostage_back(sa_gpu_ctx *ctx, int n, int nvec, double *Y);
/* cycle counters of the phases of the two-stage kernels since the last call (diagnostics) */
int sa_gpu_debug_ts_clocks(double *out16);
/* cholsi.cu on `nmats` dense symmetric n x n matrices (column-major, spectrum in [0, 1]): per matrix
   info2 = {number of eigenvalues <= theta or a negative failure code, iterations}, 8 Ritz values
   (ascending), the 8 Ritz vectors n x 8 (tests/test_cholsi.py) */
int sa_gpu_debug_cholsi(sa_gpu_ctx *ctx, int nmats, int n, const double *A, double theta, int *info2,
                        double *lam, double *X);

/* ---- device-pointer entry points (row-partitioned multi-GPU solve: the caller owns the
 *      vectors on the device, e.g. torch tensors, and drives the halo exchange) ----
 * Device addresses of a level's CSR matrix (valid while the level lives). */
int sa_gpu_level_dev_csr(sa_gpu_level *level, int which, const int **I, const int **J,
                         const double **A, int *rows, int *cols, int *nnz);
int sa_gpu_level_dev_dinv(sa_gpu_level *level, const double **dinv_neg);
/* Rows [row0, row0+nrows) of a device CSR matrix applied to x (indexed by global column) on
 * the context's stream.  I_row0 = I + row0; b, dinv, y, xrow are already offset to row0.
 * mode 0: y = A x, 1: y = b - A x, 2: y += A x,
 *      3: y = xrow + mult * dinv .* (A x - b)  (one smpr_compute_poly step),
 *      4: y = mult * dinv .* (-b)              (first step from x = 0) */
int sa_gpu_dev_spmv(sa_gpu_ctx *ctx, int mode, int nrows, double avg_nnz_per_row,
                    const int *I_row0, const int *J, const double *A, const double *x,
                    const double *xrow, const double *b, const double *dinv, double mult,
                    double *y);
/* dense exact coarsest solve of a solver on device vectors: x = Ac^-1 b */
int sa_gpu_solver_dev_coarse(sa_gpu_solver *solver, const double *b, double *x);
/* polynomial degree, roots (at most cap entries) and coarsest size of a solver */
int sa_gpu_solver_info(sa_gpu_solver *solver, int *degree, double *roots, int cap, int *nc);

/* ---- row-partitioned multi-GPU solve (dist.cu): one process per GPU, NCCL over NVLink ----
 * Replaces the reference's hypre ParCSR matvecs with their MPI halo exchange inside
 * tg_cycle_atb (amg/src/tg.cpp:91-132), smpr_compute_poly (amg/inc/smpr.hpp:319-339) and
 * kalchev_pcg (amg/src/mfem_addons.cpp:106-248).  Every rank holds the hierarchy and owns the
 * rows [n q / N, n (q + 1) / N) of every level. */
typedef struct sa_gpu_comm sa_gpu_comm;
typedef struct sa_gpu_dist_solver sa_gpu_dist_solver;
/* rank 0: 128-byte NCCL unique id, to be handed to every rank by the caller's launcher */
int sa_gpu_nccl_unique_id(void *id128);
/* collective over the nranks processes (nranks == 1: no NCCL communicator is created) */
int sa_gpu_comm_create(sa_gpu_ctx *ctx, const void *id128, int nranks, int rank, sa_gpu_comm **out);
void sa_gpu_comm_destroy(sa_gpu_comm *comm);
/* ---- sharded setup (SURVEY.md section 8e rows "Tentative P a8-a9" and "Smoothed P / RAP
   a13-a14"): one process per GPU, every rank holds the level's tables and ran
   sa_gpu_local_spectral on its own AE range ae_part[rank] .. ae_part[rank + 1].

   sa_gpu_dist_tentative_P: mirror of the reduce-to-owner exchange of
   ContribTent::CommunicateEigenvectors (amg/src/contrib.cpp:492-549) -- the MIS owner is the rank
   of the lowest-numbered AE containing it (amg/src/aggregates.cpp:583-593); ONE grouped
   ncclSend / ncclRecv all-to-all-v moves the MIS-restricted eigenvector blocks to the owners, the
   owners run the batched SVD (SVDInsert, amg/src/contrib.cpp:551-687), mis_numcoarsedof is
   all-reduced (the MPI_Scan of :684) and the bases go back to every rank (sec.Broadcast,
   amg/src/aggregates.cpp:1618).  stats4 (optional): owned MISes, bytes sent, bytes received,
   bytes of the gathered bases.
   sa_gpu_dist_coarse_elmats: ElementMatrixParallelCoarse::GetMatrix (amg/src/elmat.cpp:105-195)
   for the rank's own finer AEs + all-gather-v (ae_part NULL: the ranges of the tentative stage).
   sa_gpu_dist_smooth_P / sa_gpu_dist_rap: interp_smooth (amg/src/interp.cpp:172-229) and
   tg_coarse_matr (amg/inc/tg.hpp:695-709) as row-partitioned SpGEMM (hypre ParMult / RAP own a
   row block per rank); every rank ends with the complete P, R, Ac. */
/* host only (no GPU): the exchange plan of sa_gpu_dist_tentative_P for one rank -- owner[mis]
   (rank of the lowest-numbered AE containing the MIS) and the doubles sent to / received from
   every rank; ae_m = accepted vectors of every AE */
int sa_gpu_mis_exchange_plan(int nmis, const int *mis_to_AE_I, const int *mis_to_AE_J,
                             const int *mis_to_dof_I, int nparts, const int *ae_m, int nranks, int rank,
                             const int *ae_part, int *owner, int64_t *send_doubles, int64_t *recv_doubles);
int sa_gpu_dist_tentative_P(sa_gpu_level *level, sa_gpu_comm *comm, const int *ae_part,
                            int avoid_ess_bdr_dofs, int *mis_numcoarsedof, int *NDc_out,
                            double *stats4);
int sa_gpu_dist_coarse_elmats(sa_gpu_level *finer, sa_gpu_level *coarse, sa_gpu_comm *comm,
                              const int *ae_part);
int sa_gpu_dist_smooth_P(sa_gpu_level *level, sa_gpu_comm *comm, int degree, const double *roots);
int sa_gpu_dist_rap(sa_gpu_level *level, sa_gpu_comm *comm, double *bytes_moved);

/* halo plans (packed boundary lists per matrix and peer), vectors; the solver must outlive it */
int sa_gpu_dist_solver_create(sa_gpu_solver *solver, sa_gpu_comm *comm, sa_gpu_dist_solver **out);
void sa_gpu_dist_solver_destroy(sa_gpu_dist_solver *d);
/* PCG from x0 = 0 with the V-cycle preconditioner; b, x: full-length host vectors (x: the rank's
 * rows, or the whole solution on every rank when gather != 0); iters as kalchev_pcg returns it
 * (negative: not converged / SPD breakdown); brr_hist: (Br, r) per iteration; solve_seconds:
 * device time of the iteration (CUDA events on the library's stream) */
int sa_gpu_dist_pcg(sa_gpu_dist_solver *d, const double *b, double *x, int maxiter, double rtol,
                    double atol, int gather, int *iters, double *brr_hist, int hist_cap,
                    int *hist_len, double *solve_seconds);
int sa_gpu_dist_solver_stats(sa_gpu_dist_solver *d, int *row_begin, int *row_end, long *halo_calls,
                             long *halo_doubles);
/* rows, nnz(A), nnz(P) of one level (for the bytes-per-iteration roofline of the solve) */
int sa_gpu_dist_solver_level_info(sa_gpu_dist_solver *d, int level, int *rows, long *nnz_A, long *nnz_P);
/* host only (no GPU needed): the exchange lists of `rank` for one CSR pattern partitioned by
 * rows (row_part) and columns (col_part), nranks + 1 entries each.  send_cnt / recv_cnt: per
 * peer; send_idx / recv_idx: concatenated, peer-major, ascending.  Returns 2 (and the needed
 * sizes in n_send / n_recv) when a capacity is too small. */
int sa_gpu_halo_plan(int rows, const int *I, const int *J, int nranks, int rank, const int *row_part,
                     const int *col_part, int *send_cnt, int *recv_cnt, int *send_idx, int send_cap,
                     int *recv_idx, int recv_cap, int *n_send, int *n_recv);

/* ---- micro-benchmarks used by bench.py for the roofline denominators ---- */
/* runs `reps` SpMVs y = A x on device-resident vectors; returns ms per SpMV */
double sa_gpu_bench_spmv(sa_gpu_level *level, int which, int reps);
/* runs `reps` fused smoother steps; returns ms per step */
double sa_gpu_bench_smoother(sa_gpu_level *level, int reps);
/* measured FP64 FMA peak of this GPU in TFLOP/s (dependent-free DFMA loop) */
double sa_gpu_bench_fp64_peak(sa_gpu_ctx *ctx);

#ifdef __cplusplus
}
#endif

#endif
