// Hierarchy builders (tg_ / ml_ / interp_) on top of the CUDA library's C ABI
// (see saamge.hpp).  Control flow follows amg/src/ml.cpp:111-236, 361-472 and
// amg/src/tg.cpp:402-540, 979-1014; all arithmetic happens in sa_gpu_* calls.
#include <chrono>
#include <cstdlib>
#include <future>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "hierarchy.hpp"
#include "saamge.hpp"

namespace saamge
{

static double now_s()
{
    return std::chrono::duration<double>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

static sa_gpu_ctx *g_ctx = NULL;

// wall-clock log of the device stages (every sa_gpu_* stage call returns synchronised)
static std::vector<std::pair<std::string, double>> g_stage_log;
static int g_stage_level = 0;
struct StageTimer
{
    std::string name;
    double t0;
    StageTimer(const char *n) : name(n), t0(now_s())
    {
        if (getenv("SA_GPU_ALLOC_DEBUG")) // lets the allocation log be read stage by stage
            std::fprintf(stderr, "[stage] l%d.%s begins\n", g_stage_level, n);
    }
    ~StageTimer()
    {
        char key[96];
        std::snprintf(key, sizeof key, "l%d.%s", g_stage_level, name.c_str());
        g_stage_log.push_back(std::make_pair(std::string(key), now_s() - t0));
    }
};

sa_gpu_ctx *proc_gpu_init(int device)
{
    if (!g_ctx)
        sa_gpu_check(sa_gpu_ctx_create(device, &g_ctx), "sa_gpu_ctx_create");
    return g_ctx;
}

sa_gpu_ctx *proc_gpu_ctx()
{
    if (!g_ctx)
        proc_gpu_init(0);
    return g_ctx;
}

void proc_gpu_finalize()
{
    if (g_ctx)
        sa_gpu_ctx_destroy(g_ctx);
    g_ctx = NULL;
}

void sa_gpu_check(int rc, const char *what)
{
    if (rc)
    {
        // the reference's failure mode: message + abort (amg/inc/common.hpp:635-647)
        std::fprintf(stderr, "ASSERT: %s failed: %s\n", what, sa_gpu_last_error());
        std::abort();
    }
}

/* ------------------------------------------------------------------ providers */

ElementMatrixStandardGeometric::ElementMatrixStandardGeometric(
    const agg_partitioning_relations_t &agg_part_rels,
    const SparseMatrix &assembled_processor_matrix, const double *blocks,
    const int64_t *offsets)
    : ElementMatrixProvider(agg_part_rels), A_(assembled_processor_matrix), blocks_(blocks),
      offsets_(offsets)
{
    is_geometric = true;
}

Matrix *ElementMatrixStandardGeometric::GetMatrix(int elno, bool &free_matr) const
{
    const int ne = agg_part_rels.elem_to_dof->RowSize(elno);
    DenseMatrix *elmat = new DenseMatrix(ne, ne);
    std::memcpy(elmat->Data(), blocks_ + offsets_[elno], sizeof(double) * ne * ne);
    free_matr = true;
    return elmat;
}

// single-AE assembly through the same device code path the batched stage uses
static SparseMatrix *build_AE_stiff_on_device(const agg_partitioning_relations_t &rels,
                                              const ElementMatrixProvider *emp, int elno);

SparseMatrix *ElementMatrixStandardGeometric::BuildAEStiff(int elno) const
{
    return build_AE_stiff_on_device(agg_part_rels, this, elno);
}

ElementMatrixDenseArray::ElementMatrixDenseArray(const agg_partitioning_relations_t &agg_part_rels,
                                                 const double *blocks, const int64_t *offsets)
    : ElementMatrixProvider(agg_part_rels), blocks_(blocks), offsets_(offsets)
{
    is_geometric = false;
}

Matrix *ElementMatrixDenseArray::GetMatrix(int elno, bool &free_matr) const
{
    const int ne = agg_part_rels.elem_to_dof->RowSize(elno);
    DenseMatrix *elmat = new DenseMatrix(ne, ne);
    std::memcpy(elmat->Data(), blocks_ + offsets_[elno], sizeof(double) * ne * ne);
    free_matr = true;
    return elmat;
}

SparseMatrix *ElementMatrixDenseArray::BuildAEStiff(int elno) const
{
    return build_AE_stiff_on_device(agg_part_rels, this, elno);
}

ElementMatrixParallelCoarse::ElementMatrixParallelCoarse(
    const agg_partitioning_relations_t &agg_part_rels, levels_level_t *level)
    : ElementMatrixProvider(agg_part_rels), level(level)
{
    is_geometric = false;
}

/* The level whose ELEMENTS these matrices are is the coarser neighbour of `level` (the finer
   level the provider was built from, amg/src/elmat.cpp:105-130); its device handle holds the
   blocks sa_gpu_coarse_elmats produced. */
static sa_gpu_level *coarse_gpu_level(const levels_level_t *finer_level)
{
    SA_ASSERT(finer_level && finer_level->coarser && finer_level->coarser->tg_data &&
              finer_level->coarser->tg_data->gpu);
    return finer_level->coarser->tg_data->gpu;
}

// amg/src/elmat.cpp:105-195: the coarse element matrix P_e^T A_AE(e) P_e of finer AE elno, read
// back from the device (caller frees: free_matr = true, amg/src/elmat.cpp:177)
Matrix *ElementMatrixParallelCoarse::GetMatrix(int elno, bool &free_matr) const
{
    sa_gpu_level *g = coarse_gpu_level(level);
    int ne = 0;
    sa_gpu_check(sa_gpu_get_element_matrix(g, elno, NULL, &ne), "sa_gpu_get_element_matrix");
    DenseMatrix *M = new DenseMatrix(ne, ne);
    sa_gpu_check(sa_gpu_get_element_matrix(g, elno, M->Data(), &ne), "sa_gpu_get_element_matrix");
    free_matr = true;
    return M;
}

// agg_build_AE_stiffm of coarse AE elno (amg/src/aggregates.cpp:959-1086), assembled on the device
SparseMatrix *ElementMatrixParallelCoarse::BuildAEStiff(int elno) const
{
    sa_gpu_level *g = coarse_gpu_level(level);
    const int n = agg_part_rels.AE_to_dof->RowSize(elno);
    std::vector<double> dense((size_t)n * n);
    sa_gpu_check(sa_gpu_build_AE_stiff(g, elno, dense.data()), "sa_gpu_build_AE_stiff");
    SparseMatrix *S = new SparseMatrix;
    S->h = S->w = n;
    S->I.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j)
            if (dense[(size_t)j * n + i] != 0.)
            {
                S->J.push_back(j);
                S->A.push_back(dense[(size_t)j * n + i]);
            }
        S->I[i + 1] = (int)S->J.size();
    }
    return S;
}

static void fill_desc(sa_gpu_level_desc &d, const agg_partitioning_relations_t &r)
{
    std::memset(&d, 0, sizeof d);
    d.ND = r.ND;
    d.NE = r.elem_to_dof->Size();
    d.nparts = r.nparts;
    d.num_mises = r.num_mises;
    d.elem_to_dof_I = r.elem_to_dof->GetI();
    d.elem_to_dof_J = r.elem_to_dof->GetJ();
    d.dof_to_elem_I = r.dof_to_elem->GetI();
    d.dof_to_elem_J = r.dof_to_elem->GetJ();
    d.AE_to_elem_I = r.AE_to_elem->GetI();
    d.AE_to_elem_J = r.AE_to_elem->GetJ();
    d.AE_to_dof_I = r.AE_to_dof->GetI();
    d.AE_to_dof_J = r.AE_to_dof->GetJ();
    d.dof_to_AE_I = r.dof_to_AE->GetI();
    d.dof_to_AE_J = r.dof_to_AE->GetJ();
    d.dof_id_inAE = r.dof_id_inAE;
    d.partitioning = r.partitioning;
    d.agg_flags = r.agg_flags;
    d.mis_to_dof_I = r.mis_to_dof->GetI();
    d.mis_to_dof_J = r.mis_to_dof->GetJ();
    d.mis_to_AE_I = r.mis_to_AE->GetI();
    d.mis_to_AE_J = r.mis_to_AE->GetJ();
    d.AE_to_mis_I = r.AE_to_mis->GetI();
    d.AE_to_mis_J = r.AE_to_mis->GetJ();
    d.mises = r.mises;
    d.mis_coarsedofoffsets = r.mis_coarsedofoffsets;
}

static SparseMatrix *build_AE_stiff_on_device(const agg_partitioning_relations_t &rels,
                                              const ElementMatrixProvider *emp, int elno)
{
    sa_gpu_level_desc d;
    fill_desc(d, rels);
    const SparseMatrix *A = emp->AssembledMatrix();
    if (A)
    {
        d.A_I = A->GetI();
        d.A_J = A->GetJ();
        d.A_data = A->GetData();
        d.assemble_with_global = 1;
    }
    else
    {
        // the stage needs some operator only in with_global mode; give it an empty one
        static const int zeroI[1] = {0};
        (void)zeroI;
    }
    d.elmat = emp->DenseBlocks();
    d.elmat_off = emp->DenseBlockOffsets();
    sa_gpu_level *lev = NULL;
    std::vector<int> emptyI;
    if (!A)
    {
        emptyI.assign((size_t)rels.ND + 1, 0);
        d.A_I = emptyI.data();
        d.A_J = emptyI.data();
        static const double zero = 0.;
        d.A_data = &zero;
    }
    sa_gpu_check(sa_gpu_level_create(proc_gpu_ctx(), &d, NULL, &lev), "sa_gpu_level_create");
    const int n = rels.AE_to_dof->RowSize(elno);
    std::vector<double> dense((size_t)n * n);
    sa_gpu_check(sa_gpu_build_AE_stiff(lev, elno, dense.data()), "sa_gpu_build_AE_stiff");
    sa_gpu_level_destroy(lev);
    SparseMatrix *S = new SparseMatrix;
    S->h = S->w = n;
    S->I.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j)
            if (dense[(size_t)j * n + i] != 0.)
            {
                S->J.push_back(j);
                S->A.push_back(dense[(size_t)j * n + i]);
            }
        S->I[i + 1] = (int)S->J.size();
    }
    return S;
}

/* ------------------------------------------------------------------ Eigensolver */

Eigensolver::Eigensolver(const int *aggregates, const agg_partitioning_relations_t &agg_part_rels,
                         int threshold)
    : agg_part_rels(agg_part_rels), threshold(threshold), count_solves(0), count_direct_solves(0),
      count_max_used(0), smallest_eigenvalue_skipped(1.)
{
    (void)aggregates;
}

void Eigensolver::GetStatistics(int &o_count_solves, int &o_count_direct_solves, int &o_count_max_used,
                                double &o_smallest_eigenvalue_skipped)
{
    o_count_solves = count_solves;
    o_count_direct_solves = count_direct_solves;
    o_count_max_used = count_max_used;
    o_smallest_eigenvalue_skipped = smallest_eigenvalue_skipped;
}

bool Eigensolver::Solve(const SparseMatrix &A, SparseMatrix *&B, int part, int agg_id, int aggregate_size,
                        double &theta, DenseMatrix &cut_evects)
{
    (void)part;
    (void)agg_id;
    (void)aggregate_size;
    const int n = A.Width();
    SA_ASSERT(A.Width() == A.Size());
    SA_ASSERT(theta >= 0. && theta <= 1. + 1e-12);
    count_solves++;
    count_direct_solves++; // the device path is always direct (no ARPACK)
    // a level of one AE whose single element is the whole matrix
    std::vector<int> I01(2), iota(n), zerosI((size_t)n + 1), zeros(std::max(1, n), 0), onesI((size_t)n + 1);
    I01[0] = 0;
    I01[1] = n;
    for (int i = 0; i < n; ++i)
        iota[i] = i;
    for (int i = 0; i <= n; ++i)
        onesI[i] = i;
    std::vector<char> flags(std::max(1, n), 0);
    std::vector<double> dense((size_t)n * n, 0.);
    for (int i = 0; i < n; ++i)
        for (int q = A.I[i]; q < A.I[i + 1]; ++q)
            dense[(size_t)A.J[q] * n + i] = A.A[q];
    const int64_t eoff[2] = {0, (int64_t)n * n};
    const int one01[2] = {0, 1};
    const int zero1[1] = {0};
    sa_gpu_level_desc d;
    std::memset(&d, 0, sizeof d);
    d.ND = n;
    d.NE = 1;
    d.nparts = 1;
    d.num_mises = 1;
    d.elem_to_dof_I = I01.data();
    d.elem_to_dof_J = iota.data();
    d.dof_to_elem_I = onesI.data();
    d.dof_to_elem_J = zeros.data();
    d.AE_to_elem_I = one01;
    d.AE_to_elem_J = zero1;
    d.AE_to_dof_I = I01.data();
    d.AE_to_dof_J = iota.data();
    d.dof_to_AE_I = onesI.data();
    d.dof_to_AE_J = zeros.data();
    d.dof_id_inAE = iota.data();
    d.partitioning = zero1;
    d.agg_flags = flags.data();
    d.mis_to_dof_I = I01.data();
    d.mis_to_dof_J = iota.data();
    d.mis_to_AE_I = one01;
    d.mis_to_AE_J = zero1;
    d.AE_to_mis_I = one01;
    d.AE_to_mis_J = zero1;
    d.mises = zeros.data();
    // (an operator is only read in with_global mode; give the level an empty one)
    d.A_I = zerosI.data();
    d.A_J = zeros.data();
    static const double zero = 0.;
    d.A_data = &zero;
    d.elmat = dense.data();
    d.elmat_off = eoff;
    d.assemble_with_global = 0;
    sa_gpu_level *lev = NULL;
    sa_gpu_check(sa_gpu_level_create(proc_gpu_ctx(), &d, NULL, &lev), "sa_gpu_level_create");
    sa_gpu_check(sa_gpu_local_spectral(lev, theta, 0, 1, 0), "sa_gpu_local_spectral");
    int m = 0;
    sa_gpu_check(sa_gpu_get_spectral_counts(lev, &m), "sa_gpu_get_spectral_counts");
    last_evals.assign(m, 0.);
    std::vector<double> Z((size_t)n * m), D(n);
    sa_gpu_check(sa_gpu_get_spectral(lev, last_evals.data(), Z.data(), D.data()), "sa_gpu_get_spectral");
    sa_gpu_level_destroy(lev);
    if (!B)
    {
        B = new SparseMatrix;
        B->h = B->w = n;
        B->I.resize((size_t)n + 1);
        B->J.resize(n);
        B->A.resize(n);
        for (int i = 0; i < n; ++i)
        {
            B->I[i] = B->J[i] = i;
            B->A[i] = D[i];
        }
        B->I[n] = n;
    }
    // append to cut_evects (amg/src/spectral.cpp:199-222)
    const int beg = cut_evects.Width();
    DenseMatrix out(n, beg + m);
    if (beg)
    {
        SA_ASSERT(cut_evects.Height() == n);
        std::memcpy(out.Data(), cut_evects.Data(), sizeof(double) * (size_t)n * beg);
    }
    std::memcpy(out.Data() + (size_t)n * beg, Z.data(), sizeof(double) * (size_t)n * m);
    cut_evects = out;
    count_max_used = std::max(count_max_used, m);
    if (theta < 0.)
        theta = 0.;
    return m > 0;
}

/* ---------------------------------------------------------------- parameters */

MultilevelParameters::MultilevelParameters(int coarsenings, int *nparts_arr_arg,
                                           int first_nu_pro, int nu_pro_arg, int nu_relax_arg,
                                           double first_theta, double theta_arg,
                                           int polynomial_coarse_space_arg,
                                           bool use_correct_nullspace, bool use_arpack,
                                           bool do_aggregates)
    : coarse_partitioner(NULL), coarse_partitioner_data(NULL), testmesh_inject(false),
      num_coarsenings(coarsenings), use_correct_nullspace(use_correct_nullspace),
      use_arpack(use_arpack), do_aggregates(do_aggregates), avoid_ess_bdr_dofs(true),
      coarse_direct(false), smooth_drop_tol(0.0)
{
    // amg/src/ml.cpp:54-89
    nparts_arr = new int[num_coarsenings];
    nu_pro = new int[num_coarsenings];
    nu_relax = new int[num_coarsenings];
    theta = new double[num_coarsenings];
    polynomial_coarse_space = new int[num_coarsenings];
    nparts_arr[0] = nparts_arr_arg[0];
    nu_pro[0] = first_nu_pro;
    nu_relax[0] = nu_relax_arg;
    theta[0] = first_theta;
    polynomial_coarse_space[0] = polynomial_coarse_space_arg;
    for (int i = 1; i < num_coarsenings; ++i)
    {
        nparts_arr[i] = nparts_arr_arg[i];
        nu_pro[i] = nu_pro_arg;
        nu_relax[i] = nu_relax_arg;
        theta[i] = theta_arg;
        polynomial_coarse_space[i] = polynomial_coarse_space_arg;
    }
}

MultilevelParameters::~MultilevelParameters()
{
    delete[] nparts_arr;
    delete[] nu_pro;
    delete[] nu_relax;
    delete[] theta;
    delete[] polynomial_coarse_space;
}

// amg/src/smpr.cpp:266-280
double *smpr_sa_poly_roots(int &nu, int *degree)
{
    SA_ASSERT(nu >= 0);
    const double denom = (double)(2 * nu + 1);
    double *roots = new double[(*degree = nu) > 0 ? nu : 1];
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
    double *roots = new double[*degree = twonu + nu + 1];
    for (int i = 0; i <= twonu; ++i)
    {
        const double val = cos(((double)i * M_PI) / denom);
        roots[i] = val * val;
    }
    for (int i = 1; i <= nu; ++i)
    {
        const double val = sin(((double)i * M_PI) / denom);
        roots[i + twonu] = val * val;
    }
    return roots;
}

/* -------------------------------------------------------------------- two-grid */

// amg/src/tg.cpp:402-430 + interp_init_data (amg/src/interp.cpp:231-277) +
// smpr_init_poly_data (amg/src/smpr.cpp:359-423)
tg_data_t *tg_init_data(const SparseMatrix *A, const agg_partitioning_relations_t &agg_part_rels,
                        int nu_pro, int nu_relax, double theta, bool smooth_interp,
                        double smooth_drop_tol, bool use_arpack)
{
    tg_data_t *tg_data = new tg_data_t;
    std::memset(tg_data, 0, sizeof(*tg_data));
    tg_data->A_host = A; // (NULL on coarse levels: read back on demand)
    tg_data->theta = theta;
    interp_data_t *id = new interp_data_t;
    std::memset(id, 0, sizeof(*id));
    id->nparts = agg_part_rels.nparts;
    SA_ASSERT(nu_pro >= 0);
    id->nu_pro = nu_pro;
    id->interp_smoother_roots = smpr_sa_poly_roots(id->nu_pro, &id->interp_smoother_degree);
    id->times_apply_smoother = 1;
    id->use_arpack = use_arpack;
    id->scaling_P = false;
    id->drop_tol = smooth_drop_tol;
    tg_data->interp_data = id;
    smpr_poly_data_t *pd = new smpr_poly_data_t;
    std::memset(pd, 0, sizeof(*pd));
    SA_ASSERT(nu_relax > 0);
    pd->nu = nu_relax;
    pd->roots = smpr_sas_poly_roots(pd->nu, &pd->degree);
    pd->weightfirst = 1.;
    tg_data->poly_data = pd;
    tg_data->smooth_interp = smooth_interp;
    tg_data->tag = -1;
    tg_data->doing_spectral = false;
    tg_data->polynomial_coarse_space = -1;
    return tg_data;
}

// the AE loop of amg/src/interp.cpp:387-556 as one batched device stage
/* AE sharding over ranks (one process per GPU): the rank computes a contiguous AE range
   balanced on n^3 and an exchange callback (NCCL all-gather, installed by the launcher)
   completes the set on every rank -- how the reference spreads its AE loop over MPI ranks
   (amg/src/interp.cpp:387). */
static int g_shard_rank = 0, g_shard_world = 1;
static sa_spectral_exchange_ft g_shard_exchange = NULL;

void sa_set_sharding(int rank, int world, sa_spectral_exchange_ft exchange)
{
    g_shard_rank = rank;
    g_shard_world = world;
    g_shard_exchange = exchange;
}

/* Owner-sharded setup (SURVEY section 8e rows "Tentative P" and "Smoothed P / RAP"): with a
   library communicator installed the eigenvectors stay on the rank that computed them; the
   tentative prolongator is built by the MIS owners after one all-to-all-v of MIS-restricted
   blocks (sa_gpu_dist_tentative_P), the coarse element matrices by the AE owners, and the
   smoothing / RAP products by row blocks (sa_gpu_dist_smooth_P, sa_gpu_dist_rap). */
static sa_gpu_comm *g_shard_comm = NULL;
static double g_shard_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};

void sa_set_sharding_comm(sa_gpu_comm *comm, int rank, int world)
{
    g_shard_comm = comm;
    if (comm)
    {
        g_shard_rank = rank;
        g_shard_world = world;
    }
    for (int i = 0; i < 8; ++i)
        g_shard_stats[i] = 0.;
}

const double *sa_sharding_stats() { return g_shard_stats; }

static bool owner_sharded() { return g_shard_comm != NULL && g_shard_world > 1; }

static std::vector<int> shard_part(const agg_partitioning_relations_t &rels)
{
    std::vector<int> part((size_t)g_shard_world + 1, 0);
    for (int q = 0; q < g_shard_world; ++q)
    {
        int a = 0, b = 0;
        sa_shard_range(rels, q, g_shard_world, &a, &b);
        part[q] = a;
        part[q + 1] = b;
    }
    return part;
}

void sa_shard_range(const agg_partitioning_relations_t &rels, int rank, int world, int *begin,
                    int *end)
{
    const int nparts = rels.nparts;
    std::vector<double> cost((size_t)nparts + 1, 0.);
    for (int i = 0; i < nparts; ++i)
    {
        const double n = rels.AE_to_dof->RowSize(i);
        cost[i + 1] = cost[i] + n * n * n;
    }
    auto bound = [&](int r) {
        if (r <= 0)
            return 0;
        if (r >= world)
            return nparts;
        const double target = cost[nparts] * r / world;
        int b = (int)(std::lower_bound(cost.begin() + 1, cost.end(), target) - (cost.begin() + 1)) + 1;
        return std::min(b, nparts);
    };
    *begin = bound(rank);
    *end = std::max(*begin, bound(rank + 1));
}

void interp_compute_vectors(const agg_partitioning_relations_t &agg_part_rels,
                            const interp_data_t &interp_data, tg_data_t &tg_data, double &theta)
{
    StageTimer tm("local_spectral");
    if (owner_sharded() && !interp_data.testmesh_inject)
    {
        // this rank's AEs only; nothing is exchanged here (see interp_sparse_tent_assemble)
        int a = 0, b = 0;
        sa_shard_range(agg_part_rels, g_shard_rank, g_shard_world, &a, &b);
        sa_gpu_check(sa_gpu_local_spectral(tg_data.gpu, theta, a, b, 0), "sa_gpu_local_spectral");
        return;
    }
    if (g_shard_world > 1 && g_shard_exchange)
    {
        int a = 0, b = 0;
        sa_shard_range(agg_part_rels, g_shard_rank, g_shard_world, &a, &b);
        sa_gpu_check(sa_gpu_local_spectral(tg_data.gpu, theta, a, b,
                                           (interp_data.testmesh_inject && a == 0) ? 1 : 0),
                     "sa_gpu_local_spectral");
        g_shard_exchange(tg_data.gpu, a, b, agg_part_rels.nparts);
        return;
    }
    sa_gpu_check(sa_gpu_local_spectral(tg_data.gpu, theta, 0, agg_part_rels.nparts,
                                       interp_data.testmesh_inject ? 1 : 0),
                 "sa_gpu_local_spectral");
    // theta averaging of amg/src/interp.cpp:571-589 is the identity when all_eigens == false
}

// amg/src/interp.cpp:728-759
void interp_sparse_tent_assemble(const agg_partitioning_relations_t &agg_part_rels,
                                 interp_data_t &interp_data, tg_data_t &tg_data,
                                 bool avoid_ess_bdr_dofs)
{
    delete[] interp_data.mis_numcoarsedof;
    interp_data.mis_numcoarsedof = new int[agg_part_rels.num_mises > 0 ? agg_part_rels.num_mises : 1];
    int NDc = 0;
    StageTimer tm("tentative");
    if (owner_sharded() && !interp_data.testmesh_inject)
    {
        const std::vector<int> part = shard_part(agg_part_rels);
        double st4[4] = {0, 0, 0, 0};
        sa_gpu_check(sa_gpu_dist_tentative_P(tg_data.gpu, g_shard_comm, part.data(),
                                             avoid_ess_bdr_dofs ? 1 : 0,
                                             interp_data.mis_numcoarsedof, &NDc, st4),
                     "sa_gpu_dist_tentative_P");
        g_shard_stats[0] += st4[1]; // bytes of MIS blocks sent to their owners
        g_shard_stats[1] += st4[2]; // ... received as owner
        g_shard_stats[2] += st4[3]; // bytes of the all-reduced MIS bases
    }
    else
        sa_gpu_check(sa_gpu_tentative_P(tg_data.gpu, avoid_ess_bdr_dofs ? 1 : 0,
                                        interp_data.mis_numcoarsedof, &NDc),
                     "sa_gpu_tentative_P");
    interp_data.num_mises = agg_part_rels.num_mises;
    interp_data.coarse_truedof_offset = 0;
}

// amg/src/tg.cpp:502-540 (spectral branch) + tg_assemble_and_smooth (:432-473)
/* Pipelined upload of the finest level (desc.async_upload = 1): allowed when the caller has
   promised that the host arrays of the level stay valid and are page-locked (sa_drv_problem_pin);
   the eigen stage then starts while the operator / element blocks are still in flight. */
static bool g_async_finest_upload = false;
void sa_set_async_finest_upload(bool on) { g_async_finest_upload = on; }

void tg_build_hierarchy(const SparseMatrix *Ag, tg_data_t &tg_data,
                        const agg_partitioning_relations_t &agg_part_rels,
                        ElementMatrixProvider *elem_data, bool avoid_ess_bdr_dofs,
                        tg_data_t *finer)
{
    SA_ASSERT(tg_data.polynomial_coarse_space == -1 && tg_data.theta > 0.0);
    tg_data.elem_data = elem_data;
    tg_data.doing_spectral = true;

    sa_gpu_level_desc d;
    fill_desc(d, agg_part_rels);
    if (Ag)
    {
        d.A_I = Ag->GetI();
        d.A_J = Ag->GetJ();
        d.A_data = Ag->GetData();
    }
    d.elmat = elem_data->DenseBlocks();
    d.elmat_off = elem_data->DenseBlockOffsets();
    d.assemble_with_global = elem_data->AssembledMatrix() ? 1 : 0;
    // A provider that only implements the reference contract (GetMatrix / BuildAEStiff,
    // amg/inc/elmat.hpp:53-77) has no batched view: pack its element matrices into one
    // contiguous array by calling GetMatrix(e, free_matr) for every element -- the only place
    // the callback runs, on the host, before any kernel launch (amg/inc/elmat.hpp:32-35).
    std::vector<double> packed;
    std::vector<int64_t> packed_off;
    if (!d.elmat && !finer)
    {
        const int NE = d.NE;
        packed_off.assign((size_t)NE + 1, 0);
        for (int e = 0; e < NE; ++e)
        {
            const int ne = agg_part_rels.elem_to_dof->RowSize(e);
            packed_off[e + 1] = packed_off[e] + (int64_t)ne * ne;
        }
        packed.assign((size_t)packed_off[NE], 0.);
        for (int e = 0; e < NE; ++e)
        {
            bool free_matr = false;
            Matrix *M = elem_data->GetMatrix(e, free_matr);
            SA_ASSERT(M);
            const int ne = agg_part_rels.elem_to_dof->RowSize(e);
            double *dst = packed.data() + packed_off[e];
            if (const DenseMatrix *D = dynamic_cast<const DenseMatrix *>(M))
            {
                SA_ASSERT(D->Height() == ne && D->Width() == ne);
                std::memcpy(dst, D->Data(), sizeof(double) * ne * ne);
            }
            else if (const SparseMatrix *S = dynamic_cast<const SparseMatrix *>(M))
            {
                SA_ASSERT(S->Height() == ne && S->Width() == ne);
                for (int i = 0; i < ne; ++i)
                    for (int q = S->I[i]; q < S->I[i + 1]; ++q)
                        dst[(size_t)S->J[q] * ne + i] += S->A[q];
            }
            else
                SA_ASSERT(!"ElementMatrixProvider::GetMatrix returned an unknown matrix type");
            if (free_matr)
                delete M;
        }
        d.elmat = packed.data();
        d.elmat_off = packed_off.data();
    }
    if (!finer && g_async_finest_upload && packed.empty() && d.elmat)
        d.async_upload = 1;
    if (tg_data.gpu)
        sa_gpu_level_destroy(tg_data.gpu);
    tg_data.gpu = NULL;
    {
        StageTimer tm("upload");
        sa_gpu_check(sa_gpu_level_create(proc_gpu_ctx(), &d, finer ? finer->gpu : NULL,
                                         &tg_data.gpu),
                     "sa_gpu_level_create");
    }
    if (!d.elmat)
    {
        // ElementMatrixParallelCoarse: P_e^T A_AE P_e of every finer AE, on the device
        SA_ASSERT(finer && finer->gpu);
        StageTimer tm("coarse_elmats");
        if (owner_sharded() && finer->interp_data && !finer->interp_data->testmesh_inject)
            // (NULL: the AE ranges sa_gpu_dist_tentative_P used on the finer level)
            sa_gpu_check(sa_gpu_dist_coarse_elmats(finer->gpu, tg_data.gpu, g_shard_comm, NULL),
                         "sa_gpu_dist_coarse_elmats");
        else
            sa_gpu_check(sa_gpu_coarse_elmats(finer->gpu, tg_data.gpu), "sa_gpu_coarse_elmats");
    }
    // interp_sparse_tent_build (amg/src/interp.cpp:694-726)
    interp_compute_vectors(agg_part_rels, *tg_data.interp_data, tg_data, tg_data.theta);
    {
        // smpr_update_Dinv_neg (tg_init_data -> smpr_init_poly_data in the reference).  After the
        // eigen stage: it is the first stage that needs ALL rows of the operator, and with a
        // pipelined upload of the finest level the eigen stage runs while they are in flight.
        StageTimer tm("Dinv_neg");
        sa_gpu_check(sa_gpu_build_Dinv_neg(tg_data.gpu), "sa_gpu_build_Dinv_neg");
    }
    interp_sparse_tent_assemble(agg_part_rels, *tg_data.interp_data, tg_data, avoid_ess_bdr_dofs);
    // tg_smooth_interp (amg/inc/tg.hpp:678-693)
    interp_data_t &id = *tg_data.interp_data;
    StageTimer tm("smooth_P");
    if (owner_sharded())
        sa_gpu_check(sa_gpu_dist_smooth_P(tg_data.gpu, g_shard_comm,
                                          tg_data.smooth_interp ? id.interp_smoother_degree : 0,
                                          id.interp_smoother_roots),
                     "sa_gpu_dist_smooth_P");
    else
        sa_gpu_check(sa_gpu_smooth_P(tg_data.gpu, tg_data.smooth_interp ? id.interp_smoother_degree : 0,
                                     id.interp_smoother_roots),
                     "sa_gpu_smooth_P");
    // interp_smooth's drop_tol branch (AltThreshold, amg/src/interp.cpp:219-228)
    if (tg_data.smooth_interp && id.interp_smoother_degree > 0 && id.drop_tol != 0.0)
        sa_gpu_check(sa_gpu_threshold_P(tg_data.gpu, id.drop_tol, NULL, NULL), "sa_gpu_threshold_P");
    tg_data.have_Ac = false;
}

// amg/src/tg.cpp:917-932
tg_data_t *tg_produce_data(const SparseMatrix &Ag,
                           const agg_partitioning_relations_t &agg_part_rels, int nu_pro,
                           int nu_relax, ElementMatrixProvider *elem_data, double theta,
                           bool smooth_interp, int polynomial_coarse_arg, bool use_arpack,
                           bool avoid_ess_bdr_dofs)
{
    tg_data_t *tg_data =
        tg_init_data(&Ag, agg_part_rels, nu_pro, nu_relax, theta, smooth_interp, 0.0, use_arpack);
    tg_data->polynomial_coarse_space = polynomial_coarse_arg;
    tg_build_hierarchy(&Ag, *tg_data, agg_part_rels, elem_data, avoid_ess_bdr_dofs);
    return tg_data;
}

// amg/src/tg.cpp:979-1014; the coarsest solver is created by ml_impose_cycle
void tg_update_coarse_operator(tg_data_t *tg_data, bool perform_solve_init, bool coarse_direct)
{
    (void)perform_solve_init;
    (void)coarse_direct;
    SA_ASSERT(tg_data && tg_data->gpu);
    StageTimer tm("rap");
    if (owner_sharded())
    {
        double moved = 0.;
        sa_gpu_check(sa_gpu_dist_rap(tg_data->gpu, g_shard_comm, &moved), "sa_gpu_dist_rap");
        g_shard_stats[3] += moved; // bytes of product rows received from the other ranks
    }
    else
        sa_gpu_check(sa_gpu_rap(tg_data->gpu), "sa_gpu_rap");
    tg_data->have_Ac = true;
}

void tg_free_data(tg_data_t *tg_data)
{
    if (!tg_data)
        return;
    if (tg_data->poly_data)
    {
        delete[] tg_data->poly_data->roots;
        delete tg_data->poly_data;
    }
    if (tg_data->interp_data)
    {
        delete[] tg_data->interp_data->interp_smoother_roots;
        delete[] tg_data->interp_data->mis_numcoarsedof;
        delete tg_data->interp_data;
    }
    sa_gpu_level_destroy(tg_data->gpu);
    delete tg_data->A_host_owned;
    delete tg_data->elem_data; // tg_data takes ownership (amg/src/tg.cpp:518,946)
    delete tg_data;
}

/* ------------------------------------------------------------------ multilevel */

// Host work of the NEXT level that needs the current relations only (coarse elem_to_elem,
// METIS on the AE graph: 0.9 s for 40k AEs) runs on a thread while the GPU builds the current
// level.  SA_NO_TOPOLOGY_PREFETCH=1 turns it off.
namespace
{
struct TopologyPrefetch
{
    const agg_partitioning_relations_t *src = NULL;
    std::future<agg_coarse_topology_t> fut;
    void start(const agg_partitioning_relations_t *rels, int nparts_target)
    {
        if (src == rels && fut.valid())
            return; // already under way (started when the problem was partitioned)
        drop();
        if (getenv("SA_NO_TOPOLOGY_PREFETCH"))
            return;
        src = rels;
        fut = std::async(std::launch::async, [rels, nparts_target] {
            sa_host_threads_serial_here(true); // hidden behind the GPU stages: no thread team of its own
            return agg_coarse_topology(*rels, nparts_target);
        });
    }
    bool take(const agg_partitioning_relations_t *rels, agg_coarse_topology_t &out)
    {
        if (!src || !fut.valid())
            return false;
        out = fut.get();
        const bool match = src == rels;
        src = NULL;
        if (!match)
        {
            delete out.elem_to_elem;
            delete[] out.partitioning;
        }
        return match;
    }
    void drop()
    {
        agg_coarse_topology_t t;
        if (src && fut.valid())
        {
            t = fut.get();
            delete t.elem_to_elem;
            delete[] t.partitioning;
        }
        src = NULL;
    }
} g_topo_prefetch;
} // namespace

/* Look-ahead of the NEXT level's relations (agg_create_partitioning_coarse): they need the finer
   relations and mis_numcoarsedof only -- host data that exist once the tentative prolongator is
   built -- so a helper thread builds them while the GPU smooths P and forms the Galerkin product
   of the current level (128^3: 0.075 s of host work behind 0.04 s of RAP).  METIS path only (a user
   coarse_partitioner callback is not assumed to be thread safe).  SA_NO_TOPOLOGY_LOOKAHEAD=1: off. */
namespace
{
struct TopologyAhead
{
    const agg_partitioning_relations_t *src = NULL;
    std::future<agg_partitioning_relations_t *> fut;
    void start(const agg_partitioning_relations_t *rels, const int *mis_numcoarsedof, int nparts_target,
               bool avoid_ess_bdr_dofs)
    {
        drop();
        if (getenv("SA_NO_TOPOLOGY_LOOKAHEAD"))
            return;
        src = rels;
        fut = std::async(std::launch::async, [rels, mis_numcoarsedof, nparts_target, avoid_ess_bdr_dofs] {
            int nparts = nparts_target;
            int *partitioning = NULL;
            Table *pre_e2e = NULL;
            agg_coarse_topology_t topo;
            if (g_topo_prefetch.take(rels, topo))
            {
                pre_e2e = topo.elem_to_elem;
                partitioning = topo.partitioning;
                nparts = topo.nparts;
            }
            return agg_create_partitioning_coarse(*rels, mis_numcoarsedof, &nparts, avoid_ess_bdr_dofs,
                                                  partitioning, pre_e2e);
        });
    }
    agg_partitioning_relations_t *take(const agg_partitioning_relations_t *rels)
    {
        if (!src || !fut.valid())
            return NULL;
        agg_partitioning_relations_t *r = fut.get();
        const bool match = src == rels;
        src = NULL;
        if (!match)
        {
            agg_free_partitioning(r);
            return NULL;
        }
        return r;
    }
    void drop()
    {
        if (src && fut.valid())
            agg_free_partitioning(fut.get());
        src = NULL;
    }
} g_topo_ahead;
} // namespace

/* The coarse agglomeration of the first coarse level only needs the fine relations: drivers
   may start it as soon as those exist (METIS agglomeration is a host-side input of the path). */
void sa_topology_prefetch_start(const agg_partitioning_relations_t *rels, int nparts_target)
{
    g_topo_prefetch.start(rels, nparts_target);
}

void sa_topology_prefetch_drop(const agg_partitioning_relations_t *rels)
{
    if (g_topo_prefetch.src == rels)
        g_topo_prefetch.drop();
}

static void levels_list_push_coarse_data(levels_list_t &list,
                                         agg_partitioning_relations_t *agg_part_rels,
                                         tg_data_t *tg_data)
{
    levels_level_t *lev = new levels_level_t;
    lev->finer = list.coarsest;
    lev->coarser = NULL;
    lev->agg_part_rels = agg_part_rels;
    lev->tg_data = tg_data;
    if (list.coarsest)
        list.coarsest->coarser = lev;
    else
        list.finest = lev;
    list.coarsest = lev;
    list.num_levels++;
}

levels_level_t *levels_list_get_level(const levels_list_t &list, int i)
{
    levels_level_t *l = list.finest;
    while (l && i-- > 0)
        l = l->coarser;
    return l;
}

// amg/src/ml.cpp:111-236
void ml_produce_hierarchy_from_level(int coarsenings, int starting_level, ml_data_t &ml_data,
                                     const MultilevelParameters &mlp)
{
    SA_ASSERT(1 <= ml_data.levels_list.num_levels);
    agg_partitioning_relations_t *agg_part_rels = ml_data.levels_list.coarsest->agg_part_rels;
    tg_data_t *tg_data = ml_data.levels_list.coarsest->tg_data;
    for (int i = starting_level; i < coarsenings; ++i)
    {
        SA_ASSERT(tg_data->have_Ac);
        g_stage_level = i;
        int nparts = mlp.get_nparts(i);
        int *partitioning = NULL;
        StageTimer *ttopo = new StageTimer("topology");
        Table *pre_e2e = NULL;
        if (agg_partitioning_relations_t *ahead = g_topo_ahead.take(agg_part_rels))
        {
            // built by the helper thread while the GPU formed the finer level's P and Ac
            agg_part_rels = ahead;
            nparts = ahead->nparts;
            goto topology_done;
        }
        if (mlp.coarse_partitioner)
        {
            g_topo_prefetch.drop();
            partitioning = mlp.coarse_partitioner(i, agg_part_rels->nparts, &nparts,
                                                  mlp.coarse_partitioner_data);
        }
        else
        {
            agg_coarse_topology_t topo;
            if (g_topo_prefetch.take(agg_part_rels, topo))
            {
                pre_e2e = topo.elem_to_elem;
                partitioning = topo.partitioning;
                nparts = topo.nparts;
            }
        }
        agg_part_rels = agg_create_partitioning_coarse(
            *agg_part_rels, tg_data->interp_data->mis_numcoarsedof, &nparts,
            mlp.get_avoid_ess_bdr_dofs(), partitioning, pre_e2e);
    topology_done:
        delete ttopo;
        if (i + 1 < coarsenings && !mlp.coarse_partitioner)
            g_topo_prefetch.start(agg_part_rels, mlp.get_nparts(i + 1));
        tg_data_t *finer_tg = tg_data;
        tg_data = tg_init_data(NULL, *agg_part_rels, mlp.get_nu_pro(i), mlp.get_nu_relax(i),
                               mlp.get_theta(i), mlp.get_smooth_interp(i),
                               mlp.get_smooth_drop_tol(), mlp.get_use_arpack());
        tg_data->use_w_cycle = false;
        tg_data->polynomial_coarse_space = mlp.get_polynomial_coarse_space(i);
        ElementMatrixProvider *emp =
            new ElementMatrixParallelCoarse(*agg_part_rels, ml_data.levels_list.coarsest);
        tg_build_hierarchy(NULL, *tg_data, *agg_part_rels, emp, mlp.get_avoid_ess_bdr_dofs(),
                           finer_tg);
        if (i + 1 < coarsenings && !mlp.coarse_partitioner)
            g_topo_ahead.start(agg_part_rels, tg_data->interp_data->mis_numcoarsedof, mlp.get_nparts(i + 1),
                               mlp.get_avoid_ess_bdr_dofs());
        tg_update_coarse_operator(tg_data, i + 1 == coarsenings, mlp.get_coarse_direct());
        levels_list_push_coarse_data(ml_data.levels_list, agg_part_rels, tg_data);
    }
    // amg/src/ml.cpp:225-235: the coarsest solver becomes a CorrectNullspace two-grid cycle
    if (mlp.get_use_correct_nullspace())
        ml_build_correct_nullspace(ml_data);
    ml_impose_cycle(ml_data, false);
}

void ml_build_correct_nullspace(ml_data_t &ml_data)
{
    levels_level_t *last = ml_data.levels_list.coarsest;
    SA_ASSERT(last && last->tg_data && last->tg_data->gpu && last->tg_data->have_Ac);
    // the reference fixes the smoother of this cycle to smpr_init_poly_data(A, 3, 0.0)
    // (amg/src/solve.cpp:75, amg/src/ml.cpp:233-235); the device cycle has one degree for all levels
    SA_ASSERT(ml_data.nu_relax == 3);
    StageTimer tm("correct_nullspace");
    const agg_partitioning_relations_t &rels = *last->agg_part_rels;
    const interp_data_t &id = *last->tg_data->interp_data;
    const int nmis = rels.num_mises;
    // MIS bases of the level (mis_tent_interps): s x k column-major blocks, MIS-major
    std::vector<int64_t> off((size_t)nmis + 1, 0);
    for (int mis = 0; mis < nmis; ++mis)
        off[mis + 1] = off[mis] + (int64_t)rels.mis_to_dof->RowSize(mis) * id.mis_numcoarsedof[mis];
    std::vector<double> tent((size_t)std::max<int64_t>(1, off[nmis]));
    sa_gpu_check(sa_gpu_get_mis_tent(last->tg_data->gpu, tent.data()), "sa_gpu_get_mis_tent");
    // local_coarse_one_representation (amg/src/contrib.cpp:655-668): x = argmin ||V x - 1||,
    // normalised; V has orthonormal columns, so x = V^T 1 (what dgels returns to round-off).
    // interp_scaling_P_assemble (amg/src/interp.cpp:842-909): coarse dof (row) -> its MIS (column,
    // counting the MISes with coarse dofs), value = that representation
    SparseMatrix *SP = new SparseMatrix;
    int rows = 0, col = 0;
    SP->I.push_back(0);
    for (int mis = 0; mis < nmis; ++mis)
    {
        const int k = id.mis_numcoarsedof[mis], sz = rels.mis_to_dof->RowSize(mis);
        if (k <= 0)
            continue;
        std::vector<double> x(k, 0.);
        double norm = 0.;
        for (int c = 0; c < k; ++c)
        {
            const double *v = tent.data() + off[mis] + (int64_t)sz * c;
            for (int r = 0; r < sz; ++r)
                x[c] += v[r];
            norm += x[c] * x[c];
        }
        norm = std::sqrt(norm);
        for (int c = 0; c < k; ++c)
        {
            SP->J.push_back(col);
            SP->A.push_back(x[c] / norm);
            SP->I.push_back((int)SP->J.size());
            ++rows;
        }
        ++col;
    }
    SP->h = rows;
    SP->w = col;
    delete ml_data.scaling_P;
    ml_data.scaling_P = SP;
    if (ml_data.correct_nullspace_level)
        sa_gpu_level_destroy(ml_data.correct_nullspace_level);
    ml_data.correct_nullspace_level = NULL;
    sa_gpu_check(sa_gpu_level_create_from_P(proc_gpu_ctx(), last->tg_data->gpu, SP->w, SP->I.data(),
                                            SP->J.data(), SP->A.data(), &ml_data.correct_nullspace_level),
                 "sa_gpu_level_create_from_P");
    sa_gpu_check(sa_gpu_build_Dinv_neg(ml_data.correct_nullspace_level), "sa_gpu_build_Dinv_neg");
    sa_gpu_check(sa_gpu_rap(ml_data.correct_nullspace_level), "sa_gpu_rap");
}

// amg/src/ml.cpp:361-377: chain the levels; coarsest level gets the exact solver
void ml_impose_cycle(ml_data_t &ml_data, bool Wcycle)
{
    SA_ASSERT(!Wcycle);
    std::vector<sa_gpu_level *> levels;
    int i = 0;
    for (levels_level_t *level = ml_data.levels_list.finest; level; level = level->coarser)
    {
        level->tg_data->tag = i++;
        levels.push_back(level->tg_data->gpu);
    }
    // CorrectNullspace::Mult (amg/src/solve.cpp:137-164) is one tg_cycle_atb on the coarsest
    // operator: one more level of the device cycle
    if (ml_data.correct_nullspace_level)
        levels.push_back(ml_data.correct_nullspace_level);
    if (ml_data.gpu_solver)
        sa_gpu_solver_destroy(ml_data.gpu_solver);
    ml_data.gpu_solver = NULL;
    StageTimer tm("solver_create");
    sa_gpu_check(sa_gpu_solver_create(proc_gpu_ctx(), levels.data(), (int)levels.size(),
                                      ml_data.nu_relax, &ml_data.gpu_solver),
                 "sa_gpu_solver_create");
    // user smoothers (tg_data_t::pre_smoother / post_smoother): host callbacks through a trampoline
    i = 0;
    for (levels_level_t *level = ml_data.levels_list.finest; level; level = level->coarser, ++i)
    {
        tg_data_t *tg = level->tg_data;
        if (!tg->pre_smoother && !tg->post_smoother)
            continue;
        if (!tg->A_host)
        {
            // the operator of a coarse level lives on the device: read it back once
            int rows = 0, cols = 0, nnz = 0;
            sa_gpu_check(sa_gpu_get_csr_sizes(tg->gpu, SA_GPU_MAT_A, &rows, &cols, &nnz), "sa_gpu_get_csr_sizes");
            SparseMatrix *A = new SparseMatrix;
            A->h = rows;
            A->w = cols;
            A->I.resize((size_t)rows + 1);
            A->J.resize(std::max(1, nnz));
            A->A.resize(std::max(1, nnz));
            sa_gpu_check(sa_gpu_get_csr(tg->gpu, SA_GPU_MAT_A, A->I.data(), A->J.data(), A->A.data()), "sa_gpu_get_csr");
            A->J.resize(nnz);
            A->A.resize(nnz);
            tg->A_host_owned = A;
            tg->A_host = A;
        }
        struct tramp
        {
            static void pre(int, int n, const double *b, double *x, void *data)
            {
                tg_data_t *t = (tg_data_t *)data;
                Vector vb(b, b + n), vx(x, x + n);
                t->pre_smoother(*t->A_host, vb, vx, t->smoother_data);
                std::copy(vx.begin(), vx.end(), x);
            }
            static void post(int, int n, const double *b, double *x, void *data)
            {
                tg_data_t *t = (tg_data_t *)data;
                Vector vb(b, b + n), vx(x, x + n);
                t->post_smoother(*t->A_host, vb, vx, t->smoother_data);
                std::copy(vx.begin(), vx.end(), x);
            }
        };
        sa_gpu_check(sa_gpu_solver_set_smoothers(ml_data.gpu_solver, i, tg->pre_smoother ? tramp::pre : NULL,
                                                 tg->post_smoother ? tramp::post : NULL, tg),
                     "sa_gpu_solver_set_smoothers");
    }
}

// amg/src/ml.cpp:379-472
ml_data_t *ml_produce_data(const SparseMatrix &Ag, agg_partitioning_relations_t *agg_part_rels,
                           ElementMatrixProvider *elem_data_finest,
                           const MultilevelParameters &mlp)
{
    SA_ASSERT(elem_data_finest);
    ml_data_t *ml_data = new ml_data_t;
    std::memset(ml_data, 0, sizeof(*ml_data));
    SA_ASSERT(mlp.get_num_coarsenings() > 0);
    g_stage_level = 0;
    g_stage_log.clear();
    ml_data->nu_relax = mlp.get_nu_relax(0);
    tg_data_t *tg_data =
        tg_init_data(&Ag, *agg_part_rels, mlp.get_nu_pro(0), mlp.get_nu_relax(0), mlp.get_theta(0),
                     mlp.get_smooth_interp(0), mlp.get_smooth_drop_tol(), mlp.get_use_arpack());
    tg_data->use_w_cycle = false;
    tg_data->polynomial_coarse_space = mlp.get_polynomial_coarse_space(0);
    tg_data->interp_data->testmesh_inject = mlp.testmesh_inject;
    if (mlp.get_num_coarsenings() > 1 && !mlp.coarse_partitioner)
        g_topo_prefetch.start(agg_part_rels, mlp.get_nparts(1));
    tg_build_hierarchy(&Ag, *tg_data, *agg_part_rels, elem_data_finest,
                       mlp.get_avoid_ess_bdr_dofs());
    if (mlp.get_num_coarsenings() > 1 && !mlp.coarse_partitioner)
        g_topo_ahead.start(agg_part_rels, tg_data->interp_data->mis_numcoarsedof, mlp.get_nparts(1),
                           mlp.get_avoid_ess_bdr_dofs());
    tg_update_coarse_operator(tg_data, 1 >= mlp.get_num_coarsenings(), mlp.get_coarse_direct());
    levels_list_push_coarse_data(ml_data->levels_list, agg_part_rels, tg_data);
    ml_produce_hierarchy_from_level(mlp.get_num_coarsenings(), 1, *ml_data, mlp);
    return ml_data;
}

// amg/src/ml.cpp:474-486: the finest agg_part_rels belongs to the caller
void ml_free_data(ml_data_t *ml_data)
{
    if (!ml_data)
        return;
    if (ml_data->gpu_solver)
        sa_gpu_solver_destroy(ml_data->gpu_solver);
    if (ml_data->correct_nullspace_level)
        sa_gpu_level_destroy(ml_data->correct_nullspace_level);
    delete ml_data->scaling_P;
    // free coarse to fine: a level's operator aliases the finer level's Ac
    levels_level_t *level = ml_data->levels_list.coarsest;
    while (level)
    {
        levels_level_t *finer = level->finer;
        tg_free_data(level->tg_data);
        if (finer)
            agg_free_partitioning(level->agg_part_rels);
        delete level;
        level = finer;
    }
    delete ml_data;
}

void VCycleSolver::Mult(const Vector &b, Vector &x) const
{
    x.resize(b.size());
    sa_gpu_check(sa_gpu_vcycle(ml_data->gpu_solver, b.data(), x.data()), "sa_gpu_vcycle");
}

int kalchev_pcg(ml_data_t &ml_data, const Vector &b, Vector &x, int print_iter,
                int max_num_iter, double RTOLERANCE, double ATOLERANCE,
                std::vector<double> *brr_history)
{
    int iters = 0, hl = 0;
    std::vector<double> hist((size_t)max_num_iter + 2);
    sa_gpu_check(sa_gpu_pcg(ml_data.gpu_solver, b.data(), x.data(), max_num_iter, RTOLERANCE,
                            ATOLERANCE, &iters, hist.data(), (int)hist.size(), &hl),
                 "sa_gpu_pcg");
    if (print_iter)
        for (int i = 0; i < hl; ++i)
            std::printf("PCG Iteration: %d, (B r, r) = %g\n", i, hist[i]);
    if (brr_history)
        brr_history->assign(hist.begin(), hist.begin() + hl);
    return iters;
}

void tg_download_results(const tg_data_t &tg_data, const agg_partitioning_relations_t &rels,
                         tg_data_t *coarser, sa_level_results_t &R)
{
    sa_gpu_level *g = tg_data.gpu;
    R.nparts = rels.nparts;
    R.num_mises = rels.num_mises;
    R.ND = rels.ND;
    R.ae_m.resize(rels.nparts);
    sa_gpu_check(sa_gpu_get_spectral_counts(g, R.ae_m.data()), "sa_gpu_get_spectral_counts");
    R.ae_eval_off.assign((size_t)rels.nparts + 1, 0);
    R.ae_evect_off.assign((size_t)rels.nparts + 1, 0);
    const bool inject = tg_data.interp_data->testmesh_inject;
    for (int i = 0; i < rels.nparts; ++i)
    {
        const int n = rels.AE_to_dof->RowSize(i);
        const int nev = R.ae_m[i] - ((inject && i == 0) ? 1 : 0);
        R.ae_eval_off[i + 1] = R.ae_eval_off[i] + nev;
        R.ae_evect_off[i + 1] = R.ae_evect_off[i] + (int64_t)n * R.ae_m[i];
    }
    R.evals.resize(R.ae_eval_off[rels.nparts]);
    R.evects.resize(R.ae_evect_off[rels.nparts]);
    R.ae_D.resize(rels.AE_to_dof->Size_of_connections());
    sa_gpu_check(sa_gpu_get_spectral(g, R.evals.data(), R.evects.data(), R.ae_D.data()),
                 "sa_gpu_get_spectral");
    R.mis_numcoarsedof.assign(tg_data.interp_data->mis_numcoarsedof,
                              tg_data.interp_data->mis_numcoarsedof + rels.num_mises);
    R.mis_off.assign((size_t)rels.num_mises + 1, 0);
    for (int mis = 0; mis < rels.num_mises; ++mis)
        R.mis_off[mis + 1] =
            R.mis_off[mis] + (int64_t)rels.mises_size[mis] * R.mis_numcoarsedof[mis];
    R.mis_tent.resize(R.mis_off[rels.num_mises]);
    sa_gpu_check(sa_gpu_get_mis_tent(g, R.mis_tent.data()), "sa_gpu_get_mis_tent");
    struct
    {
        int which;
        SparseMatrix *M;
    } mats[3] = {{SA_GPU_MAT_PTENT, &R.tent_interp}, {SA_GPU_MAT_P, &R.interp},
                 {SA_GPU_MAT_AC, &R.Ac}};
    for (int k = 0; k < 3; ++k)
    {
        int rows, cols, nnz;
        sa_gpu_check(sa_gpu_get_csr_sizes(g, mats[k].which, &rows, &cols, &nnz),
                     "sa_gpu_get_csr_sizes");
        SparseMatrix &M = *mats[k].M;
        M.h = rows;
        M.w = cols;
        M.I.resize((size_t)rows + 1);
        M.J.resize(nnz);
        M.A.resize(nnz);
        sa_gpu_check(sa_gpu_get_csr(g, mats[k].which, M.I.data(), M.J.data(), M.A.data()),
                     "sa_gpu_get_csr");
    }
    R.NDc = R.Ac.h;
    R.Dinv_neg.resize(rels.ND);
    sa_gpu_check(sa_gpu_get_Dinv_neg(g, R.Dinv_neg.data()), "sa_gpu_get_Dinv_neg");
    R.celmat_off.clear();
    R.celmat.clear();
    if (coarser && coarser->gpu)
    {
        // sizes follow the coarse elem_to_dof rows; the coarse level knows them
        int rows, cols, nnz;
        (void)cols;
        (void)nnz;
        (void)rows;
    }
}

} // namespace saamge

extern "C" void sa_drv_set_sharding(int rank, int world, sa_drv_exchange_ft cb)
{
    saamge::sa_set_sharding(rank, world, (saamge::sa_spectral_exchange_ft)cb);
}

extern "C" void sa_drv_set_sharding_comm(void *comm, int rank, int world)
{
    saamge::sa_set_sharding_comm((sa_gpu_comm *)comm, rank, world);
}

extern "C" void sa_drv_sharding_stats(double *out8)
{
    for (int i = 0; i < 8; ++i)
        out8[i] = saamge::sa_sharding_stats()[i];
}

/* ------------------------------------------------------ driver entry points */

using namespace saamge;

struct product_impl_t
{
    ml_data_t *ml = NULL;
    int64_t launches0 = 0;
};

static void product_impl_free(void *p)
{
    product_impl_t *pi = (product_impl_t *)p;
    ml_free_data(pi->ml);
    delete pi;
}

struct block_partitioner_data_t
{
    const sa_problem_t *prob;
    const sa_drv_params_t *p;
};

static int *block_coarse_partitioner(int level, int num_elem, int *nparts, void *data)
{
    block_partitioner_data_t *d = (block_partitioner_data_t *)data;
    return sa_prescribed_coarse_partitioning(*d->prob, *d->p, level, num_elem, nparts);
}

/* ---- user-style plug-ins for the boundary test (what a reference user would write) ---- */
namespace
{
// implements ONLY the reference contract (amg/inc/elmat.hpp:53-77): no batched view
class UserGetMatrixProvider : public ElementMatrixProvider
{
public:
    UserGetMatrixProvider(const agg_partitioning_relations_t &rels, const fem_problem_t &f)
        : ElementMatrixProvider(rels), f_(f)
    {
    }
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const
    {
        DenseMatrix *M = new DenseMatrix(f_.ne, f_.ne);
        std::memcpy(M->Data(), f_.elmat.data() + (size_t)elno * f_.ne * f_.ne, sizeof(double) * f_.ne * f_.ne);
        free_matr = true;
        return M;
    }
    virtual SparseMatrix *BuildAEStiff(int) const
    {
        SA_ASSERT(!"not used: the hierarchy builder batches the AEs");
        return NULL;
    }

private:
    const fem_problem_t &f_;
};

// a user smoother with the smpr_ft signature: the SAS polynomial smoother written out on the host
// (amg/inc/smpr.hpp:319-339), so the V-cycle must give the same iteration count as the device one
struct user_smoother_data_t
{
    int degree;
    const double *roots;
    Vector dinv_neg;
    int calls;
};
void user_host_smoother(const SparseMatrix &A, const Vector &b, Vector &x, void *data)
{
    user_smoother_data_t *d = (user_smoother_data_t *)data;
    d->calls++;
    const int n = (int)b.size();
    Vector t(n);
    for (int k = 0; k < d->degree; ++k)
    {
        SpMult(A, x.data(), t.data());
        for (int i = 0; i < n; ++i)
            x[i] += (1. / d->roots[k]) * d->dinv_neg[i] * (t[i] - b[i]);
    }
}
user_smoother_data_t g_user_smoother;
} // namespace

/* flags bit 0: element matrices through a GetMatrix-only provider; bit 1: user smoother
   (host callback) on the finest level.  Returns a hierarchy handle like sa_drv_ml_build. */
extern "C" void *sa_drv_ml_build_user(void *prob_, const sa_drv_params_t *p, int device, int flags)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    proc_gpu_init(device);
    sa_hierarchy_t *H = new sa_hierarchy_t;
    H->prob = prob;
    H->params = *p;
    product_impl_t *pi = new product_impl_t;
    H->impl = pi;
    H->impl_free = product_impl_free;
    H->owns_rels = false;
    const fem_problem_t &f = *prob->fem;
    std::vector<int> nparts_arr = sa_target_nparts(f.NE, *p);
    std::vector<int64_t> offsets((size_t)f.NE + 1);
    for (int e = 0; e <= f.NE; ++e)
        offsets[e] = (int64_t)e * f.ne * f.ne;
    ElementMatrixProvider *emp;
    if (flags & 1)
        emp = new UserGetMatrixProvider(*prob->rels, f);
    else
        emp = new ElementMatrixDenseArray(*prob->rels, f.elmat.data(), offsets.data());
    MultilevelParameters mlp(p->num_levels - 1, nparts_arr.data(), p->first_nu_pro, p->nu_pro,
                             p->nu_relax, p->first_theta, p->theta, -1, p->correct_nullspace != 0, false, false);
    mlp.set_coarse_direct(true);
    mlp.set_smooth_drop_tol(p->smooth_drop_tol);
    block_partitioner_data_t bpd = {prob, p};
    if (p->partition_kind == 1 || !prob->coarse_partitions.empty())
        mlp.set_coarse_partitioner(block_coarse_partitioner, &bpd);
    const double t0 = now_s();
    pi->ml = ml_produce_data(f.A, prob->rels, emp, mlp);
    if (flags & 2)
    {
        tg_data_t *tg = pi->ml->levels_list.finest->tg_data;
        g_user_smoother.degree = tg->poly_data->degree;
        g_user_smoother.roots = tg->poly_data->roots;
        g_user_smoother.dinv_neg.resize(f.A.Height());
        g_user_smoother.calls = 0;
        sa_gpu_check(sa_gpu_get_Dinv_neg(tg->gpu, g_user_smoother.dinv_neg.data()), "sa_gpu_get_Dinv_neg");
        tg->pre_smoother = user_host_smoother;
        tg->post_smoother = user_host_smoother;
        tg->smoother_data = &g_user_smoother;
        ml_impose_cycle(*pi->ml, false);
    }
    sa_gpu_ctx_sync(proc_gpu_ctx());
    H->times["setup"] = now_s() - t0;
    for (levels_level_t *l = pi->ml->levels_list.finest; l; l = l->coarser)
        H->rels.push_back(l->agg_part_rels);
    return H;
}

/* Page-locks the large host arrays of the problem (element blocks, operator, relation tables)
   and lets sa_drv_ml_build upload the finest level through the pipelined path.  Part of the host
   input producer (like the pinning of sa_drv_bench_create); returns seconds spent. */
extern "C" double sa_drv_problem_pin(void *prob_, int device)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    proc_gpu_init(device);
    const double t0 = now_s();
    if (prob->pinned.empty())
    {
        const fem_problem_t &f = *prob->fem;
        const agg_partitioning_relations_t &r = *prob->rels;
        auto pin = [&](const void *p, size_t bytes) {
            if (p && bytes && sa_gpu_host_register(p, bytes) == 0)
                prob->pinned.push_back(p);
        };
        pin(f.elmat.data(), f.elmat.size() * sizeof(double));
        pin(f.A.GetData(), f.A.A.size() * sizeof(double));
        pin(f.A.GetJ(), f.A.J.size() * sizeof(int));
        pin(f.A.GetI(), f.A.I.size() * sizeof(int));
        const Table *tabs[] = {r.elem_to_dof, r.dof_to_elem, r.AE_to_elem, r.AE_to_dof,
                               r.dof_to_AE,   r.mis_to_dof,  r.mis_to_AE,  r.AE_to_mis};
        for (const Table *t : tabs)
        {
            pin(t->GetI(), t->I.size() * sizeof(int));
            pin(t->GetJ(), t->J.size() * sizeof(int));
        }
        pin(r.dof_id_inAE, (size_t)r.dof_to_AE->Size_of_connections() * sizeof(int));
        pin(r.partitioning, (size_t)f.NE * sizeof(int));
        pin(r.agg_flags, (size_t)r.ND);
        pin(r.mises, (size_t)r.ND * sizeof(int));
    }
    sa_set_async_finest_upload(true);
    return now_s() - t0;
}

extern "C" void sa_drv_problem_unpin(void *prob_)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    for (size_t i = 0; i < prob->pinned.size(); ++i)
        sa_gpu_host_unregister(prob->pinned[i]);
    prob->pinned.clear();
    sa_set_async_finest_upload(false);
}

extern "C" int sa_drv_user_smoother_calls(void) { return g_user_smoother.calls; }

/* Eigensolver::Solve through the C++ mirror on a dense symmetric matrix (n x n column-major):
   returns m, evals (cap entries), vectors (n x m) and the diagonal B. */
extern "C" int sa_drv_eigensolver_solve(int device, int n, const double *A, double theta, int cap,
                                        double *evals, double *evects, double *Bdiag)
{
    proc_gpu_init(device);
    SparseMatrix S;
    S.h = S.w = n;
    S.I.assign((size_t)n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
        for (int j = 0; j < n; ++j)
            if (A[(size_t)j * n + i] != 0. || i == j)
            {
                S.J.push_back(j);
                S.A.push_back(A[(size_t)j * n + i]);
            }
        S.I[i + 1] = (int)S.J.size();
    }
    agg_partitioning_relations_t dummy;
    std::memset(&dummy, 0, sizeof dummy);
    Eigensolver es(NULL, dummy);
    SparseMatrix *B = NULL;
    DenseMatrix cut;
    double th = theta;
    es.Solve(S, B, 0, 0, n, th, cut);
    const int m = cut.Width();
    for (int j = 0; j < std::min(m, cap); ++j)
    {
        evals[j] = es.LastEigenvalues()[j];
        std::memcpy(evects + (size_t)j * n, cut.Data() + (size_t)j * n, sizeof(double) * n);
    }
    for (int i = 0; i < n; ++i)
        Bdiag[i] = B->A[i];
    delete B;
    return m;
}

/* ElementMatrixParallelCoarse::GetMatrix / BuildAEStiff of the provider of coarse level `level`
   (>= 1): reads element elno / assembles AE ae through the C++ mirror; dense column-major out. */
extern "C" int sa_drv_coarse_provider_probe(void *hier, int level, int elno, double *elmat_out, int *ne_out,
                                            int ae, double *AE_out, int *n_out)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    product_impl_t *pi = (product_impl_t *)H->impl;
    levels_level_t *l = levels_list_get_level(pi->ml->levels_list, level);
    if (!l || !l->tg_data || !l->tg_data->elem_data)
        return 1;
    const ElementMatrixProvider *emp = l->tg_data->elem_data;
    bool fr = false;
    Matrix *M = emp->GetMatrix(elno, fr);
    const DenseMatrix *D = dynamic_cast<const DenseMatrix *>(M);
    if (!D)
        return 2;
    *ne_out = D->Height();
    if (elmat_out)
        std::memcpy(elmat_out, D->Data(), sizeof(double) * D->Height() * D->Width());
    if (fr)
        delete M;
    SparseMatrix *S = emp->BuildAEStiff(ae);
    *n_out = S->Height();
    if (AE_out)
    {
        std::fill(AE_out, AE_out + (size_t)S->h * S->h, 0.);
        for (int i = 0; i < S->h; ++i)
            for (int q = S->I[i]; q < S->I[i + 1]; ++q)
                AE_out[(size_t)S->J[q] * S->h + i] = S->A[q];
    }
    delete S;
    return 0;
}

extern "C" void *sa_drv_ml_build(void *prob_, const sa_drv_params_t *p, int device)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    sa_gpu_ctx *gctx = proc_gpu_init(device);
    // (Measured in bench.py's process, twice each: trimming the device memory pool before the
    // build costs 0.6-3.4 s -- the trim itself plus fresh memory for the 2 GB of inputs --
    // against 7.6-7.7 s without: the pool is left alone; sa_gpu_ctx_trim_pool stays available.)
    (void)gctx;
    sa_hierarchy_t *H = new sa_hierarchy_t;
    H->prob = prob;
    H->params = *p;
    product_impl_t *pi = new product_impl_t;
    H->impl = pi;
    H->impl_free = product_impl_free;
    H->owns_rels = false; // coarse relations belong to ml_data (ml_free_data)
    const fem_problem_t &f = *prob->fem;
    std::vector<int> nparts_arr = sa_target_nparts(f.NE, *p);
    const double t0 = now_s();
    std::vector<int64_t> offsets((size_t)f.NE + 1);
    for (int e = 0; e <= f.NE; ++e)
        offsets[e] = (int64_t)e * f.ne * f.ne;
    // the provider is owned by tg_data (amg/src/tg.cpp:518,946); offsets must outlive
    // tg_build_hierarchy only (data is copied to the device there)
    ElementMatrixProvider *emp =
        new ElementMatrixStandardGeometric(*prob->rels, f.A, f.elmat.data(), offsets.data());
    MultilevelParameters mlp(p->num_levels - 1, nparts_arr.data(), p->first_nu_pro, p->nu_pro,
                             p->nu_relax, p->first_theta, p->theta, -1, p->correct_nullspace != 0, false, false);
    mlp.set_coarse_direct(true);
    mlp.set_smooth_drop_tol(p->smooth_drop_tol);
    mlp.testmesh_inject = p->testmesh_inject != 0;
    block_partitioner_data_t bpd = {prob, p};
    if (p->partition_kind == 1 || !prob->coarse_partitions.empty())
        mlp.set_coarse_partitioner(block_coarse_partitioner, &bpd);
    pi->ml = ml_produce_data(f.A, prob->rels, emp, mlp);
    sa_gpu_ctx_sync(proc_gpu_ctx());
    H->times["setup"] = now_s() - t0;
    for (size_t i = 0; i < g_stage_log.size(); ++i)
        H->times[g_stage_log[i].first] += g_stage_log[i].second;
    for (levels_level_t *l = pi->ml->levels_list.finest; l; l = l->coarser)
        H->rels.push_back(l->agg_part_rels);
    return H;
}

/* Algebraic entry (algebraic.cpp): the problem comes from sa_drv_problem_from_matrix; local
   matrices by ExtractSubMatrices, then the ordinary multilevel build.  Level 0 of the handle
   reports the caller's relations (cells = dofs). */
extern "C" void *sa_drv_ml_build_algebraic(void *prob_, const sa_drv_params_t *p, int device)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    proc_gpu_init(device);
    sa_hierarchy_t *H = new sa_hierarchy_t;
    H->prob = prob;
    H->params = *p;
    product_impl_t *pi = new product_impl_t;
    H->impl = pi;
    H->impl_free = product_impl_free;
    H->owns_rels = false;
    const fem_problem_t &f = *prob->fem;
    std::vector<int> nparts_arr = sa_target_nparts(f.NE, *p);
    nparts_arr[0] = prob->rels->nparts;
    const double t0 = now_s();
    MultilevelParameters mlp(p->num_levels - 1, nparts_arr.data(), p->first_nu_pro, p->nu_pro,
                             p->nu_relax, p->first_theta, p->theta, -1, p->correct_nullspace != 0, false, false);
    mlp.set_coarse_direct(true);
    mlp.set_smooth_drop_tol(p->smooth_drop_tol);
    pi->ml = ml_produce_data_algebraic(f.A, *prob->rels, mlp);
    sa_gpu_ctx_sync(proc_gpu_ctx());
    H->times["setup"] = now_s() - t0;
    for (size_t i = 0; i < g_stage_log.size(); ++i)
        H->times[g_stage_log[i].first] += g_stage_log[i].second;
    for (levels_level_t *l = pi->ml->levels_list.finest; l; l = l->coarser)
        H->rels.push_back(l == pi->ml->levels_list.finest ? prob->rels : l->agg_part_rels);
    return H;
}

extern "C" int sa_drv_ml_download(void *hier)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    product_impl_t *pi = (product_impl_t *)H->impl;
    H->levels.clear();
    H->levels.resize(pi->ml->levels_list.num_levels);
    int i = 0;
    for (levels_level_t *l = pi->ml->levels_list.finest; l; l = l->coarser, ++i)
    {
        tg_download_results(*l->tg_data, *l->agg_part_rels, l->coarser ? l->coarser->tg_data : NULL,
                            H->levels[i]);
        if (l->coarser)
        {
            // coarse element matrices of this level's AEs live on the coarser level
            const agg_partitioning_relations_t &cr = *l->coarser->agg_part_rels;
            sa_level_results_t &R = H->levels[i];
            R.celmat_off.assign((size_t)l->agg_part_rels->nparts + 1, 0);
            for (int e = 0; e < l->agg_part_rels->nparts; ++e)
            {
                const int64_t nc = cr.elem_to_dof->RowSize(e);
                R.celmat_off[e + 1] = R.celmat_off[e] + nc * nc;
            }
            R.celmat.resize(R.celmat_off[l->agg_part_rels->nparts]);
            sa_gpu_check(sa_gpu_get_coarse_elmats(l->coarser->tg_data->gpu, R.celmat.data()),
                         "sa_gpu_get_coarse_elmats");
        }
    }
    if (pi->ml->correct_nullspace_level)
    {
        H->cn_P = *pi->ml->scaling_P;
        sa_gpu_level *g = pi->ml->correct_nullspace_level;
        int rows = 0, cols = 0, nnz = 0;
        sa_gpu_check(sa_gpu_get_csr_sizes(g, SA_GPU_MAT_AC, &rows, &cols, &nnz), "sa_gpu_get_csr_sizes");
        H->cn_Ac.h = rows;
        H->cn_Ac.w = cols;
        H->cn_Ac.I.assign((size_t)rows + 1, 0);
        H->cn_Ac.J.assign((size_t)std::max(1, nnz), 0);
        H->cn_Ac.A.assign((size_t)std::max(1, nnz), 0.);
        sa_gpu_check(sa_gpu_get_csr(g, SA_GPU_MAT_AC, H->cn_Ac.I.data(), H->cn_Ac.J.data(), H->cn_Ac.A.data()),
                     "sa_gpu_get_csr");
        H->cn_Ac.J.resize(nnz);
        H->cn_Ac.A.resize(nnz);
    }
    return 0;
}

/* New values (same pattern) for the problem's operator: the input of an operator update
   (time-dependent coefficient); the element matrices and every relation stay. */
extern "C" int sa_drv_problem_set_A_values(void *prob_, const double *vals, int64_t n)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    if ((int64_t)prob->fem->A.A.size() != n)
        return 1;
    std::copy(vals, vals + n, prob->fem->A.A.begin());
    return 0;
}

/* adapt_update_operators (amg/src/adapt.cpp:189-216) with the problem's CURRENT operator values:
   new weighted-l1 smoothers, (optionally) re-smoothed prolongators from the kept tentative ones,
   fresh Galerkin products on every level, new coarsest solver -- no eigensolves. */
extern "C" int sa_drv_ml_update_operators(void *hier, int resmooth_interp)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    product_impl_t *pi = (product_impl_t *)H->impl;
    int one = 1;
    MultilevelParameters mlp(1, &one, 0, 0, pi->ml->nu_relax, 0.003, 0.003, -1, false, false, false);
    mlp.set_coarse_direct(true);
    const double t0 = now_s();
    adapt_update_operators(H->prob->fem->A, *pi->ml, mlp, resmooth_interp != 0);
    sa_gpu_ctx_sync(proc_gpu_ctx());
    H->times["update_operators"] = now_s() - t0;
    return 0;
}

extern "C" int sa_drv_ml_pcg(void *hier, int maxiter, double rtol, double atol)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    product_impl_t *pi = (product_impl_t *)H->impl;
    const fem_problem_t &f = *H->prob->fem;
    H->pcg.x.assign(f.b.size(), 0.);
    H->pcg.brr.clear();
    const double t0 = now_s();
    H->pcg.iterations = kalchev_pcg(*pi->ml, f.b, H->pcg.x, 0, maxiter, rtol, atol, &H->pcg.brr);
    H->times["pcg"] = now_s() - t0;
    Vector r(f.b.size());
    SpMult(f.A, H->pcg.x.data(), r.data());
    double s = 0.;
    for (size_t i = 0; i < f.b.size(); ++i)
        s += (f.b[i] - r[i]) * (f.b[i] - r[i]);
    H->pcg.final_res_norm = std::sqrt(s);
    return H->pcg.iterations;
}
