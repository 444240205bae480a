// Internal declarations shared by the CUDA translation units.
#ifndef SA_GPU_INTERNAL_CUH
#define SA_GPU_INTERNAL_CUH

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/saamge_b200.h"

void sa_gpu_set_error(const char *fmt, ...);

#define SA_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            sa_gpu_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,         \
                             cudaGetErrorString(e__));                             \
            throw std::runtime_error("cuda");                                      \
        }                                                                          \
    } while (0)

#define SA_FAIL(...)                                                               \
    do {                                                                           \
        sa_gpu_set_error(__VA_ARGS__);                                             \
        throw std::runtime_error("sa_gpu");                                        \
    } while (0)

// every extern "C" body is wrapped so no exception crosses the C ABI
#define SA_API_BEGIN try {
#define SA_API_END                                                                 \
    }                                                                              \
    catch (const std::exception &ex__)                                             \
    {                                                                              \
        if (std::strcmp(ex__.what(), "cuda") != 0 &&                               \
            std::strcmp(ex__.what(), "sa_gpu") != 0)                               \
            sa_gpu_set_error("exception: %s", ex__.what());                        \
        return 1;                                                                  \
    }                                                                              \
    return 0;

/* stream used for stream-ordered allocation (set by sa_gpu_ctx_create; one context per
   process is the expected use).  With a pool release threshold of "never", cudaMallocAsync /
   cudaFreeAsync recycle device memory without the cost of cudaMalloc / cudaFree. */
extern cudaStream_t g_sa_alloc_stream; // main stream of the live context(s); nullptr when none
extern bool g_sa_alloc_async;          // stream-ordered pool in use
extern int g_sa_alloc_device, g_sa_alloc_refs;

/* Device arena (capi.cu): one slab reserved when the first context is created; DevBuf takes its
   memory from it with a first-fit free list.  The stream-ordered pool of the driver was measured
   to take 0.1 - 1 s for single multi-GB requests at unpredictable points of a hierarchy build
   (growing / re-mapping the pool); sub-allocating a slab costs microseconds.  Safe for the same
   reason the pool was: every kernel that touches a DevBuf is ordered on the one main stream, so a
   block handed out again is only written by work queued after the work that used it before.
   Requests the arena cannot serve fall back to the pool.  SA_GPU_ARENA_GB=0 disables it. */
void *sa_arena_alloc(size_t bytes);       // nullptr: not served
bool sa_arena_free(void *p);              // false: not an arena pointer

template <class T> struct DevBuf
{
    T *p = nullptr;
    size_t n = 0;   // elements in use
    size_t cap = 0; // elements allocated (grow-only until release)
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    bool async_owned = false;
    bool is_view = false; // points into memory owned by someone else (never freed here)
    void view(T *ptr, size_t count)
    {
        release();
        p = ptr;
        n = cap = count;
        is_view = true;
    }
    void release()
    {
        if (p && is_view)
        {
            p = nullptr;
            n = cap = 0;
            is_view = false;
            return;
        }
        if (p)
        {
            if (arena_owned)
                sa_arena_free(p);
            // (no live context: plain cudaFree, which is valid for pool memory and synchronises)
            else if (async_owned && g_sa_alloc_stream)
                cudaFreeAsync(p, g_sa_alloc_stream);
            else
                cudaFree(p);
        }
        p = nullptr;
        n = 0;
        cap = 0;
        arena_owned = false;
    }
    bool arena_owned = false;
    void alloc(size_t count)
    {
        release();
        n = count;
        cap = count;
        if ((p = (T *)sa_arena_alloc((count ? count : 1) * sizeof(T))) != nullptr)
        {
            arena_owned = true;
            async_owned = false;
            return;
        }
        if (g_sa_alloc_async && g_sa_alloc_stream)
        {
            const size_t bytes = (count ? count : 1) * sizeof(T);
            static const bool dbg = getenv("SA_GPU_ALLOC_DEBUG") != NULL;
            const auto t0 = std::chrono::steady_clock::now();
            SA_CUDA(cudaMallocAsync((void **)&p, bytes, g_sa_alloc_stream));
            if (dbg && bytes >= ((size_t)64 << 20))
                fprintf(stderr, "[alloc] %8.1f MB in %7.2f ms\n", bytes / 1048576.,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0)
                            .count());
            async_owned = true;
        }
        else
        {
            SA_CUDA(cudaMalloc((void **)&p, (count ? count : 1) * sizeof(T)));
            async_owned = false;
        }
    }
    void ensure(size_t count)
    {
        if (count > cap || !p)
            alloc(count);
        else if (count > n)
            n = count;
    }
    void upload(const T *h, size_t count, cudaStream_t s)
    {
        ensure(count);
        n = count;
        if (count)
            SA_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count, cudaStream_t s) const
    {
        if (count)
            SA_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void zero(cudaStream_t s)
    {
        if (n)
            SA_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void swap(DevBuf &o)
    {
        std::swap(is_view, o.is_view);
        std::swap(p, o.p);
        std::swap(n, o.n);
        std::swap(cap, o.cap);
        std::swap(async_owned, o.async_owned);
        std::swap(arena_owned, o.arena_owned);
    }
};

struct DevCsr
{
    int rows = 0, cols = 0, nnz = 0;
    DevBuf<int> I, J;
    DevBuf<double> A;
    void swap(DevCsr &o)
    {
        std::swap(rows, o.rows);
        std::swap(cols, o.cols);
        std::swap(nnz, o.nnz);
        I.swap(o.I);
        J.swap(o.J);
        A.swap(o.A);
    }
};

/* Pinned, device-visible host staging area for the small per-chunk index arrays of the
   local spectral stage.  A kernel on the main stream pulls the data over PCIe, so these
   uploads do not queue behind a pipelined bulk upload in the copy engine's FIFO. */
struct HostStage
{
    char *p = nullptr;
    size_t cap = 0, used = 0;
    HostStage() {}
    HostStage(const HostStage &) = delete;
    HostStage &operator=(const HostStage &) = delete;
    ~HostStage()
    {
        if (p)
            cudaFreeHost(p);
    }
};

/* Grow-only work arrays of the local spectral stage, owned by the CONTEXT and shared by all
   its levels (calls are serial per context).  The stream-ordered pool recycles blocks of a
   size it has seen quickly, but carving a request out of larger free blocks can take seconds:
   one persistent set of arrays avoids both the per-level allocations and that carving. */
struct SpectralWs
{
    DevBuf<int> ae, doff, status, order, nev, mtot, ev_slot, ev_idx, ws_i, gpind;
    size_t invit_NB = 0; // eigenvectors per inverse-iteration batch the workspace holds
    DevBuf<int64_t> voff, eval_off, evect_off;
    DevBuf<double> V, d, e, tau, sinv, glo, ghi, tn, ws_d;
    DevBuf<double> G; // distributed copies of the matrices reduced by k_tridiag_reg
    // large-matrix (cooperative) path
    DevBuf<double> Twork, pbuf;
    DevBuf<unsigned int> counters;
    DevBuf<char> coopmats;
    // large-matrix two-stage path (twostage.cu)
    DevBuf<double> ts_band, ts_tau1, ts_wsV, ts_wsW;
    DevBuf<int> ts_of_slot;
    DevBuf<int64_t> ts_toff, ts_mats;
    DevBuf<unsigned long long> ts_slots;
    // large-matrix Cholesky + subspace-iteration path (cholsi.cu)
    DevBuf<double> cs_X, cs_X2, cs_Z, cs_lam, cs_small;
    DevBuf<int> cs_info;
    DevBuf<int64_t> cs_mats;
};

/* one matrix of the two-stage tridiagonalisation (twostage.cu) */
struct sa_ts_mat
{
    int n;
    double *T;    // n x n column-major, lower triangle valid on entry; reflectors on exit
    double *band; // 64 x n band + bulge storage of stage 2
    double *d, *e; // tridiagonal matrix (out)
    double *tau1; // n: tau of the stage-1 reflectors, by column (out)
    double *tauz; // n or NULL: zeroed (the one-stage reflector array of the slot)
};

struct sa_gpu_ctx
{
    SpectralWs sws;
    int device = 0;
    int num_sms = 148;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // pipelined uploads (desc.async_upload)
    // side streams + fork/join events: independent launches of one stage (the occupancy
    // classes of the eigen stage) run side by side so their tails overlap
    static const int NAUX = 3;
    HostStage stage; // see HostStage
    cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[NAUX] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    // optional per-stage device timing (SA_GPU_PROFILE=1): name -> accumulated ms
    bool profile = false;
    cudaEvent_t pev0 = nullptr, pev1 = nullptr;
    std::vector<std::pair<std::string, double>> prof;
    void prof_add(const std::string &name, double ms)
    {
        for (size_t i = 0; i < prof.size(); ++i)
            if (prof[i].first == name)
            {
                prof[i].second += ms;
                return;
            }
        prof.push_back(std::make_pair(name, ms));
    }
};

/* scope timer: CUDA events on the context's stream, active only when profiling */
struct ProfScope
{
    sa_gpu_ctx *ctx;
    std::string name;
    ProfScope(sa_gpu_ctx *c, const char *n) : ctx(c), name(n)
    {
        if (ctx->profile)
            cudaEventRecord(ctx->pev0, ctx->stream);
    }
    ~ProfScope()
    {
        if (ctx->profile)
        {
            cudaEventRecord(ctx->pev1, ctx->stream);
            cudaEventSynchronize(ctx->pev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->pev0, ctx->pev1);
            ctx->prof_add(name, ms);
        }
    }
};

#define SA_LAUNCH(ctx, kernel, grid, block, smem, ...)                             \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);           \
        (ctx)->launches++;                                                         \
        SA_CUDA(cudaGetLastError());                                               \
    } while (0)

/* device view of the integer tables of a level */
struct LevelTables
{
    int ND, NE, nparts, num_mises;
    const int *e2d_I, *e2d_J;
    const int *d2e_I, *d2e_J;
    const int *AE2e_I, *AE2e_J;
    const int *AE2d_I, *AE2d_J;
    const int *d2AE_I, *d2AE_J, *dof_id_inAE;
    const int *partitioning;
    const char *agg_flags;
    const int *mis2d_I, *mis2d_J;
    const int *mis2AE_I, *mis2AE_J;
    const int *AE2mis_I, *AE2mis_J;
    const int *mises;
    const int *A_I, *A_J;
    const double *A_data;
    const double *elmat;
    const int64_t *elmat_off;
    int with_global;
};


/* Pipelined (demand-driven) upload of a level.  sa_gpu_level_create queues the tables on the
   context's copy stream and returns; the operator rows and the element blocks follow in the
   order the local spectral stage asks for them: sa_level_queue_upload(a0, a1) queues exactly
   the rows / blocks AEs [a0, a1) read and that are not queued yet (merged into at most a few
   hundred contiguous runs) and returns the index of the event that fires when they have
   arrived.  This works for any AE numbering; sa_level_queue_rest queues what is left. */
/* Events of the upload requests of a level: appended by the helper thread of the eigen stage
   while the main thread waits on earlier entries.  Fixed storage and an atomic count, so that
   size() / operator[] on the reader's side never race with an append (entries below size() are
   published by the release store). */
struct EventList
{
    static const int CAP = 256;
    cudaEvent_t e[CAP];
    std::atomic<int> n{0};
    size_t size() const { return (size_t)n.load(std::memory_order_acquire); }
    bool empty() const { return size() == 0; }
    cudaEvent_t operator[](size_t i) const { return e[i]; }
    cudaEvent_t back() const { return e[size() - 1]; }
    void reserve(size_t) {}
    void clear() { n.store(0, std::memory_order_release); }
    void push_back(cudaEvent_t ev)
    {
        const int k = n.load(std::memory_order_relaxed);
        if (k >= CAP)
            throw std::runtime_error("EventList: too many upload requests");
        e[k] = ev;
        n.store(k + 1, std::memory_order_release);
    }
};

struct PendingUpload
{
    bool active = false;
    bool complete = false; // every row / element block has been queued
    EventList ev;
    // the caller keeps the host arrays valid while the upload is pending (API contract), so
    // the level's host-side copies of the index arrays are deferred until the GPU is busy
    sa_gpu_level_desc desc;
    bool host_copies_done = true;
    std::vector<unsigned char> elem_mark, row_mark; // 0 not queued, 1 queued, 2 being queued
    bool timing = false;
    // desc.async_upload == 2: what the local spectral stage does not read stays on the host until
    // another entry point needs it (sharded stage: a rank uploads its own AEs' inputs only)
    bool lazy_rest = false;
    double bytes_queued = 0.; // operator + element block bytes queued so far
};

struct sa_gpu_level
{
    PendingUpload pending;
    sa_gpu_ctx *ctx = nullptr;
    sa_gpu_level *finer = nullptr;
    int ND = 0, NE = 0, nparts = 0, num_mises = 0;
    int with_global = 0;
    // host copies of the small index arrays needed for bucketing
    std::vector<int> h_AE2d_I, h_mis2d_I, h_mis2AE_I, h_mis2AE_J, h_e2d_I;
    std::vector<int> h_mis_coarsedofoffsets; // finer MIS -> coarse dof (coarse levels)
    // device tables
    DevBuf<int> e2d_I, e2d_J, d2e_I, d2e_J, AE2e_I, AE2e_J, AE2d_I, AE2d_J, d2AE_I, d2AE_J,
        dof_id_inAE, partitioning, mis2d_I, mis2d_J, mis2AE_I, mis2AE_J, AE2mis_I, AE2mis_J,
        mises, mis_cdof_off;
    DevBuf<char> agg_flags;
    // operator: own copy (finest) or alias of finer->Ac
    DevCsr A_own;
    DevCsr *A = nullptr;
    // element matrices: own (finest) or produced by sa_gpu_coarse_elmats
    DevBuf<double> elmat;
    DevBuf<int64_t> elmat_off;
    std::vector<int64_t> h_elmat_off;
    bool have_elmat = false;
    // local spectral results
    bool have_spectral = false;
    std::vector<int> h_ae_m;             // accepted vectors per AE (incl. injected)
    std::vector<int> h_ae_nev;           // eigenvalues per AE
    std::vector<int64_t> h_evect_off;    // nparts+1
    std::vector<int64_t> h_eval_off;     // nparts+1
    DevBuf<int> ae_m;
    DevBuf<int64_t> evect_off, eval_off;
    DevBuf<double> evals, evects, ae_D;  // ae_D offsets = AE2d_I
    double max_residual = 0.;
    // [0]: AEs with an eigenvalue within 1e-12 of theta, [1]: MISes with a singular value within
    // 10x of the rank cut (decisions made on round-off; reported, not altered)
    DevBuf<int> borderline;
    int h_theta_borderline = 0;
    // tentative P
    bool have_tent = false;
    int avoid_ess = 1;
    std::vector<int> h_mis_ncd;          // mis_numcoarsedof
    std::vector<int64_t> h_mis_off;      // num_mises+1 (s*k)
    DevBuf<int> mis_ncd, mis_cd_off;     // coarse dof offsets per MIS (num_mises+1)
    DevBuf<int64_t> mis_off;
    DevBuf<double> mis_tent;
    std::vector<int> h_ae_part;          // AE ranges of the ranks of a sharded setup (nranks + 1)
    int NDc = 0;
    DevCsr Ptent, P, R, Ac;
    bool have_P = false, have_Ac = false;
    DevBuf<double> Dinv_neg;
    bool have_Dinv = false;
    // scratch vectors for host-buffer SpMV / smoother calls and micro-benchmarks
    DevBuf<double> vx, vy, vb;

    LevelTables tables() const;
};

/* capi.cu: make the context's stream wait for a pending pipelined upload (all of it) */
void sa_level_ready(sa_gpu_level *lev);
/* demand-driven parts (see PendingUpload) */
int sa_level_queue_upload(sa_gpu_level *lev, int a0, int a1);
int sa_level_queue_rest(sa_gpu_level *lev);
void sa_level_wait_event(sa_gpu_level *lev, int idx);
/* deferred host-side copies of a pipelined level (no-op when done) */
void sa_level_host_copies(sa_gpu_level *lev);

/* ---- sparse.cu ---- */
void dev_csr_transpose(sa_gpu_ctx *ctx, const DevCsr &A, DevCsr &At);
void dev_spgemm(sa_gpu_ctx *ctx, const DevCsr &A, const DevCsr &B, DevCsr &C);
void dev_spmv(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, double *y);
/* y = b - A x */
void dev_residual(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, const double *b, double *y);
/* y += A x */
void dev_spmv_add(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, double *y);
/* xout = xin + mult * dinv_neg .* (A xin - b);  first==1: xin == 0 (skips the SpMV) */
void dev_smoother_step(sa_gpu_ctx *ctx, const DevCsr &A, const double *dinv_neg, const double *b,
                       const double *xin, double *xout, double mult, int xin_is_zero);
void dev_spmv_rows(sa_gpu_ctx *ctx, int mode, int nrows, double avg, const int *I_row0,
                   const int *J, const double *A, const double *x, const double *xrow,
                   const double *b, const double *dinv, double mult, double *y);
void dev_exclusive_scan_i32(sa_gpu_ctx *ctx, const int *in, int *out, int n); // out has n+1
void dev_exclusive_scan_i64(sa_gpu_ctx *ctx, const int64_t *in, int64_t *out, int n);

/* ---- twostage.cu ---- */
/* dense -> band -> tridiagonal for nmats matrices (sorted largest first; nmax = largest n) */
void sa_ts_reduce(sa_gpu_ctx *ctx, const sa_ts_mat *d_mats, int nmats, int nmax, cudaStream_t st);
/* z <- Q1 Q2 z for every eigenvector of the chunk whose slot went through sa_ts_reduce */
void sa_ts_back(sa_gpu_ctx *ctx, const sa_ts_mat *d_mats, const int *d_ts_of_slot, const int *d_ev_slot,
                const int *d_ev_idx, int nev_total, const int64_t *d_evect_off_slot, double *d_evects,
                int nmax, cudaStream_t st);
size_t sa_ts_s1_smem_bytes(int nmax, int *w_in_out);

#endif
