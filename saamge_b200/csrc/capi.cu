// Context / level lifetime, read-back and micro-benchmark entry points of the C ABI
// (include/saamge_b200.h).
#include <algorithm>
#include <chrono>
#include <cmath>

#include "sa_gpu_internal.cuh"

static thread_local char g_err[1024] = "";
cudaStream_t g_sa_alloc_stream = nullptr;
bool g_sa_alloc_async = false;
int g_sa_alloc_device = -1, g_sa_alloc_refs = 0;

// ---- device arena (see sa_gpu_internal.cuh)
#include <map>
namespace
{
struct Arena
{
    char *base = nullptr;
    size_t size = 0;
    std::map<size_t, size_t> free_blocks; // offset -> size
    std::map<size_t, size_t> used;        // offset -> size
} g_arena;
const size_t ARENA_ALIGN = 512;
} // namespace

void *sa_arena_alloc(size_t bytes)
{
    if (!g_arena.base)
        return nullptr;
    bytes = (bytes + ARENA_ALIGN - 1) & ~(ARENA_ALIGN - 1);
    for (std::map<size_t, size_t>::iterator it = g_arena.free_blocks.begin(); it != g_arena.free_blocks.end(); ++it)
        if (it->second >= bytes)
        {
            const size_t off = it->first, sz = it->second;
            g_arena.free_blocks.erase(it);
            if (sz > bytes)
                g_arena.free_blocks[off + bytes] = sz - bytes;
            g_arena.used[off] = bytes;
            return g_arena.base + off;
        }
    return nullptr;
}

bool sa_arena_free(void *p)
{
    if (!g_arena.base || (char *)p < g_arena.base || (char *)p >= g_arena.base + g_arena.size)
        return false;
    size_t off = (size_t)((char *)p - g_arena.base);
    std::map<size_t, size_t>::iterator u = g_arena.used.find(off);
    if (u == g_arena.used.end())
        return false;
    size_t sz = u->second;
    g_arena.used.erase(u);
    // coalesce with the neighbours
    std::map<size_t, size_t>::iterator nx = g_arena.free_blocks.lower_bound(off);
    if (nx != g_arena.free_blocks.end() && off + sz == nx->first)
    {
        sz += nx->second;
        g_arena.free_blocks.erase(nx);
    }
    nx = g_arena.free_blocks.lower_bound(off);
    if (nx != g_arena.free_blocks.begin())
    {
        std::map<size_t, size_t>::iterator pv = nx;
        --pv;
        if (pv->first + pv->second == off)
        {
            off = pv->first;
            sz += pv->second;
            g_arena.free_blocks.erase(pv);
        }
    }
    g_arena.free_blocks[off] = sz;
    return true;
}

static void arena_create(int device)
{
    if (g_arena.base)
        return;
    const char *e = getenv("SA_GPU_ARENA_GB");
    const char *na = getenv("SA_GPU_NO_ASYNC_ALLOC");
    if (na && na[0] == '1')
        return; // (the plain cudaMalloc test path)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
        return;
    double gb = e ? atof(e) : std::min(96., 0.6 * (double)free_b / 1073741824.);
    if (gb <= 0.)
        return;
    size_t bytes = (size_t)(gb * 1073741824.) & ~(size_t)0xfffff;
    static const bool dbg = getenv("SA_GPU_ALLOC_DEBUG") != NULL;
    const auto t0 = std::chrono::steady_clock::now();
    void *p = nullptr;
    while (bytes >= ((size_t)1 << 30) && cudaMalloc(&p, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        p = nullptr;
        bytes /= 2;
    }
    if (!p)
        return;
    if (dbg)
        fprintf(stderr, "[arena] %.1f GB reserved in %.1f ms\n", bytes / 1073741824.,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    g_arena.base = (char *)p;
    g_arena.size = bytes;
    g_arena.free_blocks.clear();
    g_arena.used.clear();
    g_arena.free_blocks[0] = bytes;
    (void)device;
}

static void arena_destroy()
{
    // only when nothing lives in it any more (buffers may outlive the last context)
    if (g_arena.base && g_arena.used.empty())
    {
        cudaFree(g_arena.base);
        g_arena.base = nullptr;
        g_arena.size = 0;
        g_arena.free_blocks.clear();
    }
}

void sa_gpu_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *sa_gpu_last_error(void) { return g_err; }

extern "C" size_t sa_gpu_level_desc_size(void) { return sizeof(sa_gpu_level_desc); }

extern "C" int sa_gpu_ctx_create(int device, sa_gpu_ctx **out)
{
    SA_API_BEGIN
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        SA_FAIL("sa_gpu_ctx_create: no CUDA device available (%s); this library has no CPU "
                "fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev)
        SA_FAIL("sa_gpu_ctx_create: device %d out of range (%d devices)", device, ndev);
    SA_CUDA(cudaSetDevice(device));
    sa_gpu_ctx *ctx = new sa_gpu_ctx;
    ctx->device = device;
    cudaDeviceProp prop;
    SA_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    // Device memory is allocated and freed in stream order on the context's main stream, so a
    // buffer is never recycled while a kernel still uses it.  The allocation stream is a process
    // global (DevBuf has no context pointer): contexts created while another one is alive share
    // its main stream (same device only), and the stream goes away with the last of them.
    if (g_sa_alloc_refs > 0)
    {
        if (g_sa_alloc_device != device)
        {
            delete ctx;
            SA_FAIL("sa_gpu_ctx_create: a context on device %d is alive; this process cannot open "
                    "another one on device %d (one device per process)", g_sa_alloc_device, device);
        }
        ctx->stream = g_sa_alloc_stream;
    }
    else
    {
        SA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        g_sa_alloc_stream = ctx->stream;
        g_sa_alloc_device = device;
        arena_create(device);
    }
    ++g_sa_alloc_refs;
    SA_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    for (int i = 0; i < sa_gpu_ctx::NAUX; ++i)
    {
        SA_CUDA(cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
        SA_CUDA(cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming));
    }
    SA_CUDA(cudaEventCreate(&ctx->ev0));
    SA_CUDA(cudaEventCreate(&ctx->ev1));
    SA_CUDA(cudaEventCreate(&ctx->pev0));
    SA_CUDA(cudaEventCreate(&ctx->pev1));
    // stream-ordered allocator that never returns memory to the driver between calls
    {
        int supported = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device);
        const char *na = getenv("SA_GPU_NO_ASYNC_ALLOC");
        g_sa_alloc_async = false;
        if (supported && !(na && na[0] == '1'))
        {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
                g_sa_alloc_async = true;
            }
        }
    }
    const char *pe = getenv("SA_GPU_PROFILE");
    ctx->profile = pe && pe[0] == '1';
    *out = ctx;
    SA_API_END
}

extern "C" void sa_gpu_ctx_destroy(sa_gpu_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->ev0)
        cudaEventDestroy(ctx->ev0);
    if (ctx->ev1)
        cudaEventDestroy(ctx->ev1);
    if (ctx->pev0)
        cudaEventDestroy(ctx->pev0);
    if (ctx->pev1)
        cudaEventDestroy(ctx->pev1);
    // work arrays first (stream-ordered frees need the stream)
    ctx->sws.~SpectralWs();
    new (&ctx->sws) SpectralWs();
    bool last = false;
    if (ctx->stream)
    {
        cudaStreamSynchronize(ctx->stream);
        if (--g_sa_alloc_refs <= 0)
        {
            // buffers that outlive the last context are freed with cudaFree (DevBuf::release)
            g_sa_alloc_refs = 0;
            g_sa_alloc_stream = nullptr;
            g_sa_alloc_async = false;
            last = true;
            arena_destroy();
        }
    }
    if (ctx->copy_stream)
        cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < sa_gpu_ctx::NAUX; ++i)
    {
        if (ctx->aux[i])
            cudaStreamDestroy(ctx->aux[i]);
        if (ctx->join_ev[i])
            cudaEventDestroy(ctx->join_ev[i]);
    }
    if (ctx->fork_ev)
        cudaEventDestroy(ctx->fork_ev);
    if (ctx->stream && last)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int sa_gpu_ctx_profile(sa_gpu_ctx *ctx, int enable, char *buf, int buflen)
{
    // enable: 1/0 switches profiling, -1 leaves it; the accumulated "name ms" lines are
    // written to buf (if given) and the table is cleared
    if (buf && buflen > 0)
    {
        std::string out;
        char line[160];
        for (size_t i = 0; i < ctx->prof.size(); ++i)
        {
            snprintf(line, sizeof line, "%s %.6f\n", ctx->prof[i].first.c_str(),
                     ctx->prof[i].second);
            out += line;
        }
        snprintf(buf, buflen, "%s", out.c_str());
        ctx->prof.clear();
    }
    if (enable >= 0)
        ctx->profile = enable != 0;
    return 0;
}

extern "C" int sa_gpu_ctx_trim_pool(sa_gpu_ctx *ctx)
{
    SA_API_BEGIN
    SA_CUDA(cudaSetDevice(ctx->device));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaMemPool_t pool;
    if (g_sa_alloc_async && cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess)
        SA_CUDA(cudaMemPoolTrimTo(pool, 0));
    SA_API_END
}

extern "C" void *sa_gpu_ctx_stream(sa_gpu_ctx *ctx) { return (void *)ctx->stream; }

extern "C" int sa_gpu_ctx_sync(sa_gpu_ctx *ctx)
{
    SA_API_BEGIN
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    SA_API_END
}

extern "C" int64_t sa_gpu_ctx_launch_count(sa_gpu_ctx *ctx) { return ctx->launches; }

extern "C" double sa_gpu_ctx_timer(sa_gpu_ctx *ctx, int begin)
{
    if (begin)
    {
        cudaEventRecord(ctx->ev0, ctx->stream);
        return 0.;
    }
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaEventSynchronize(ctx->ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    return (double)ms;
}

static void upload_table(DevBuf<int> &dI, DevBuf<int> &dJ, const int *I, const int *J, int rows,
                         cudaStream_t st)
{
    dI.upload(I, (size_t)rows + 1, st);
    dJ.upload(J, (size_t)I[rows], st);
}

void sa_level_wait_event(sa_gpu_level *lev, int idx)
{
    PendingUpload &P = lev->pending;
    if (!P.active || idx < 0 || idx >= (int)P.ev.size())
        return;
    SA_CUDA(cudaStreamWaitEvent(lev->ctx->stream, P.ev[idx], 0));
}

namespace
{
/* runs of entries marked 2 in [lo, hi]; unmarked gaps of at most `gap` entries are swallowed
   (and marked) so that the number of copies stays small */
void collect_runs(std::vector<unsigned char> &mark, int lo, int hi, int gap,
                  std::vector<std::pair<int, int>> &runs)
{
    runs.clear();
    int i = lo;
    while (i <= hi)
    {
        if (mark[i] != 2)
        {
            ++i;
            continue;
        }
        int r0 = i, r1 = i + 1; // [r0, r1)
        int j = r1;
        while (j <= hi)
        {
            if (mark[j] == 2)
            {
                r1 = ++j;
                continue;
            }
            // look ahead: is there another entry to queue within the gap?
            int k = j;
            while (k <= hi && k - r1 < gap && mark[k] != 2)
                ++k;
            if (k <= hi && k - r1 < gap && mark[k] == 2)
                j = k;
            else
                break;
        }
        runs.push_back(std::make_pair(r0, r1));
        i = r1;
    }
}

int queue_marked(sa_gpu_level *L, int elo, int ehi, int rlo, int rhi)
{
    PendingUpload &P = L->pending;
    const sa_gpu_level_desc &d = P.desc;
    cudaStream_t cs = L->ctx->copy_stream;
    std::vector<std::pair<int, int>> runs;
    if (d.elmat && elo <= ehi)
    {
        int gap = 8;
        do
        {
            collect_runs(P.elem_mark, elo, ehi, gap, runs);
            gap *= 4;
        } while (runs.size() > 48);
        for (size_t q = 0; q < runs.size(); ++q)
        {
            const int e0 = runs[q].first, e1 = runs[q].second;
            for (int e = e0; e < e1; ++e)
                P.elem_mark[e] = 1;
            const size_t k0 = d.elmat_off[e0], k1 = d.elmat_off[e1];
            if (k1 > k0)
            {
                SA_CUDA(cudaMemcpyAsync(L->elmat.p + k0, d.elmat + k0, (k1 - k0) * sizeof(double),
                                        cudaMemcpyHostToDevice, cs));
                P.bytes_queued += (double)(k1 - k0) * sizeof(double);
            }
        }
    }
    if (d.A_I && rlo <= rhi)
    {
        int gap = 32;
        do
        {
            collect_runs(P.row_mark, rlo, rhi, gap, runs);
            gap *= 4;
        } while (runs.size() > 48);
        for (size_t q = 0; q < runs.size(); ++q)
        {
            const int r0 = runs[q].first, r1 = runs[q].second;
            for (int r = r0; r < r1; ++r)
                P.row_mark[r] = 1;
            const size_t k0 = d.A_I[r0], k1 = d.A_I[r1];
            if (k1 > k0)
            {
                SA_CUDA(cudaMemcpyAsync(L->A_own.J.p + k0, d.A_J + k0, (k1 - k0) * sizeof(int),
                                        cudaMemcpyHostToDevice, cs));
                SA_CUDA(cudaMemcpyAsync(L->A_own.A.p + k0, d.A_data + k0, (k1 - k0) * sizeof(double),
                                        cudaMemcpyHostToDevice, cs));
                P.bytes_queued += (double)(k1 - k0) * (sizeof(double) + sizeof(int));
            }
        }
    }
    cudaEvent_t ev;
    SA_CUDA(cudaEventCreateWithFlags(&ev, P.timing ? cudaEventDefault : cudaEventDisableTiming));
    SA_CUDA(cudaEventRecord(ev, cs));
    P.ev.push_back(ev);
    return (int)P.ev.size() - 1;
}
} // namespace

int sa_level_queue_upload(sa_gpu_level *L, int a0, int a1)
{
    PendingUpload &P = L->pending;
    if (!P.active)
        return -1;
    const sa_gpu_level_desc &d = P.desc;
    int elo = 1 << 30, ehi = -1, rlo = 1 << 30, rhi = -1;
    if (d.elmat)
        for (int k = d.AE_to_elem_I[a0]; k < d.AE_to_elem_I[a1]; ++k)
        {
            const int e = d.AE_to_elem_J[k];
            if (!P.elem_mark[e])
            {
                P.elem_mark[e] = 2;
                elo = std::min(elo, e);
                ehi = std::max(ehi, e);
            }
        }
    if (d.A_I)
        for (int k = d.AE_to_dof_I[a0]; k < d.AE_to_dof_I[a1]; ++k)
        {
            const int r = d.AE_to_dof_J[k];
            if (!P.row_mark[r])
            {
                P.row_mark[r] = 2;
                rlo = std::min(rlo, r);
                rhi = std::max(rhi, r);
            }
        }
    return queue_marked(L, elo, ehi, rlo, rhi);
}

int sa_level_queue_rest(sa_gpu_level *L)
{
    PendingUpload &P = L->pending;
    if (!P.active)
        return -1;
    if (P.complete)
        return (int)P.ev.size() - 1;
    for (size_t e = 0; e < P.elem_mark.size(); ++e)
        if (!P.elem_mark[e])
            P.elem_mark[e] = 2;
    for (size_t r = 0; r < P.row_mark.size(); ++r)
        if (!P.row_mark[r])
            P.row_mark[r] = 2;
    P.complete = true;
    return queue_marked(L, 0, (int)P.elem_mark.size() - 1, 0, (int)P.row_mark.size() - 1);
}

static void level_host_copies_from(sa_gpu_level *L, const sa_gpu_level_desc *d)
{
    L->h_mis2d_I.assign(d->mis_to_dof_I, d->mis_to_dof_I + d->num_mises + 1);
    L->h_mis2AE_I.assign(d->mis_to_AE_I, d->mis_to_AE_I + d->num_mises + 1);
    L->h_mis2AE_J.assign(d->mis_to_AE_J, d->mis_to_AE_J + d->mis_to_AE_I[d->num_mises]);
    L->h_e2d_I.assign(d->elem_to_dof_I, d->elem_to_dof_I + d->NE + 1);
    if (d->elmat)
        L->h_elmat_off.assign(d->elmat_off, d->elmat_off + d->NE + 1);
}

void sa_level_host_copies(sa_gpu_level *lev)
{
    PendingUpload &P = lev->pending;
    if (P.host_copies_done)
        return;
    level_host_copies_from(lev, &P.desc);
    P.host_copies_done = true;
}

void sa_level_ready(sa_gpu_level *lev)
{
    PendingUpload &P = lev->pending;
    sa_level_host_copies(lev);
    if (!P.active)
        return;
    sa_level_queue_rest(lev);
    // the host arrays are the caller's again once this returns: block the host, not just
    // the stream
    if (!P.ev.empty())
    {
        cudaStreamWaitEvent(lev->ctx->stream, P.ev.back(), 0);
        cudaEventSynchronize(P.ev.back());
    }
    for (size_t i = 0; i < P.ev.size(); ++i)
        cudaEventDestroy(P.ev[i]);
    P.ev.clear();
    P.elem_mark.clear();
    P.row_mark.clear();
    P.active = false;
}

extern "C" double sa_gpu_level_uploaded_bytes(sa_gpu_level *lev)
{
    return lev->pending.bytes_queued;
}

extern "C" int sa_gpu_level_trim(sa_gpu_level *lev)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    SA_CUDA(cudaStreamSynchronize(lev->ctx->stream));
    SpectralWs &W = lev->ctx->sws;
    W.V.release();
    W.Twork.release();
    W.ws_d.release();
    W.ws_i.release();
    W.pbuf.release();
    W.invit_NB = 0;
    SA_API_END
}

extern "C" int sa_gpu_level_upload_wait(sa_gpu_level *lev)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    SA_API_END
}

extern "C" int sa_gpu_level_create(sa_gpu_ctx *ctx, const sa_gpu_level_desc *d,
                                   sa_gpu_level *finer, sa_gpu_level **out)
{
    SA_API_BEGIN
    SA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    sa_gpu_level *L = new sa_gpu_level;
    struct Guard
    {
        sa_gpu_level *l;
        ~Guard()
        {
            if (l)
            {
                sa_level_ready(l);
                delete l;
            }
        }
    } g{L};
    L->ctx = ctx;
    L->finer = finer;
    L->ND = d->ND;
    L->NE = d->NE;
    L->nparts = d->nparts;
    L->num_mises = d->num_mises;
    L->with_global = d->assemble_with_global;
    const bool dbg = getenv("SA_GPU_PIPE_DEBUG") != NULL;
    const auto t_in = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (dbg)
            fprintf(stderr, "[level_create] %s at %.2f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_in)
                        .count());
    };
    const bool async = d->async_upload != 0 && (d->A_I || d->elmat);
    if (!d->A_I)
    {
        if (!finer || !finer->have_Ac)
            SA_FAIL("sa_gpu_level_create: no operator given and the finer level has no Ac");
        if (finer->Ac.rows != d->ND)
            SA_FAIL("sa_gpu_level_create: finer Ac has %d rows, level has %d dofs",
                    finer->Ac.rows, d->ND);
        L->A = &finer->Ac;
    }
    else
    {
        L->A_own.rows = L->A_own.cols = d->ND;
        L->A_own.nnz = d->A_I[d->ND];
        L->A = &L->A_own;
    }
    const size_t nnz = d->A_I ? (size_t)d->A_I[d->ND] : 0;
    const size_t nel = d->elmat ? (size_t)d->elmat_off[d->NE] : 0;
    L->have_elmat = d->elmat != NULL;

    // every host -> device copy of the level: {device buffer, source, bytes}; the buffers
    // are allocated first (main-stream order), the copies follow on one stream
    struct Copy
    {
        void *dst;
        const void *src;
        size_t bytes;
    };
    std::vector<Copy> small, big;
    auto reg_i = [&](DevBuf<int> &b, const int *src, size_t count, std::vector<Copy> &list) {
        b.ensure(count);
        b.n = count;
        if (count)
            list.push_back(Copy{b.p, src, count * sizeof(int)});
    };
    auto reg_table = [&](DevBuf<int> &dI, DevBuf<int> &dJ, const int *I, const int *J, int rows) {
        reg_i(dI, I, (size_t)rows + 1, small);
        reg_i(dJ, J, (size_t)I[rows], small);
    };
    reg_table(L->e2d_I, L->e2d_J, d->elem_to_dof_I, d->elem_to_dof_J, d->NE);
    reg_table(L->d2e_I, L->d2e_J, d->dof_to_elem_I, d->dof_to_elem_J, d->ND);
    reg_table(L->AE2e_I, L->AE2e_J, d->AE_to_elem_I, d->AE_to_elem_J, d->nparts);
    reg_table(L->AE2d_I, L->AE2d_J, d->AE_to_dof_I, d->AE_to_dof_J, d->nparts);
    reg_table(L->d2AE_I, L->d2AE_J, d->dof_to_AE_I, d->dof_to_AE_J, d->ND);
    reg_i(L->dof_id_inAE, d->dof_id_inAE, (size_t)d->dof_to_AE_I[d->ND], small);
    reg_i(L->partitioning, d->partitioning, d->NE, small);
    L->agg_flags.ensure(d->ND);
    L->agg_flags.n = d->ND;
    if (d->ND)
        small.push_back(Copy{L->agg_flags.p, d->agg_flags, (size_t)d->ND});
    reg_table(L->mis2d_I, L->mis2d_J, d->mis_to_dof_I, d->mis_to_dof_J, d->num_mises);
    reg_table(L->mis2AE_I, L->mis2AE_J, d->mis_to_AE_I, d->mis_to_AE_J, d->num_mises);
    reg_table(L->AE2mis_I, L->AE2mis_J, d->AE_to_mis_I, d->AE_to_mis_J, d->nparts);
    reg_i(L->mises, d->mises, d->ND, small);
    if (d->A_I)
    {
        reg_i(L->A_own.I, d->A_I, (size_t)d->ND + 1, small);
        reg_i(L->A_own.J, d->A_J, nnz, big);
        L->A_own.A.ensure(nnz);
        L->A_own.A.n = nnz;
        if (nnz)
            big.push_back(Copy{L->A_own.A.p, d->A_data, nnz * sizeof(double)});
    }
    if (d->elmat)
    {
        L->elmat_off.ensure((size_t)d->NE + 1);
        L->elmat_off.n = (size_t)d->NE + 1;
        small.push_back(Copy{L->elmat_off.p, d->elmat_off, ((size_t)d->NE + 1) * sizeof(int64_t)});
        L->elmat.ensure(nel);
        L->elmat.n = nel;
        if (nel)
            big.push_back(Copy{L->elmat.p, d->elmat, nel * sizeof(double)});
    }
    auto issue = [&](const std::vector<Copy> &list, cudaStream_t s) {
        for (size_t i = 0; i < list.size(); ++i)
            SA_CUDA(cudaMemcpyAsync(list[i].dst, list[i].src, list[i].bytes, cudaMemcpyHostToDevice,
                                    s));
    };
    if (!async)
    {
        issue(small, st);
        issue(big, st);
    }
    else
    {
        // Pipelined upload: the tables go out on the copy stream now; the operator rows and
        // the element blocks follow on demand (PendingUpload).
        cudaStream_t cs = ctx->copy_stream;
        cudaEvent_t ea; // the buffers were allocated in main-stream order
        SA_CUDA(cudaEventCreateWithFlags(&ea, cudaEventDisableTiming));
        SA_CUDA(cudaEventRecord(ea, st));
        SA_CUDA(cudaStreamWaitEvent(cs, ea, 0));
        issue(small, cs);
        SA_CUDA(cudaEventRecord(ea, cs));
        SA_CUDA(cudaStreamWaitEvent(st, ea, 0));
        SA_CUDA(cudaEventDestroy(ea));
        PendingUpload &P = L->pending;
        P.active = true;
        P.complete = false;
        P.timing = dbg;
        P.lazy_rest = d->async_upload == 2;
        P.bytes_queued = 0.;
        P.ev.reserve(64); // (a helper thread appends while the main thread reads earlier entries)
        P.elem_mark.assign(d->elmat ? (size_t)d->NE : 0, 0);
        P.row_mark.assign(d->A_I ? (size_t)d->ND : 0, 0);
    }
    lap("copies issued");
    // host copies of the small index arrays
    L->h_AE2d_I.assign(d->AE_to_dof_I, d->AE_to_dof_I + d->nparts + 1);
    if (async)
    {
        PendingUpload &P = L->pending;
        P.desc = *d;
        P.host_copies_done = false; // deferred: sa_level_host_copies
    }
    else
        level_host_copies_from(L, d);
    if (d->mis_coarsedofoffsets && finer)
        L->h_mis_coarsedofoffsets.assign(d->mis_coarsedofoffsets,
                                         d->mis_coarsedofoffsets + finer->num_mises + 1);
    // synchronous mode: the host arrays are the caller's again on return
    if (!async)
        SA_CUDA(cudaStreamSynchronize(st));
    lap("return");
    g.l = nullptr;
    *out = L;
    SA_API_END
}

extern "C" void sa_gpu_level_destroy(sa_gpu_level *level)
{
    if (!level)
        return;
    cudaSetDevice(level->ctx->device);
    PendingUpload &P = level->pending;
    if (P.active && P.lazy_rest && !P.complete)
    {
        // a lazily uploaded level that never needed the rest: wait for what is in flight only
        cudaStreamSynchronize(level->ctx->copy_stream);
        for (size_t i = 0; i < P.ev.size(); ++i)
            cudaEventDestroy(P.ev[i]);
        P.ev.clear();
        P.active = false;
    }
    else
        sa_level_ready(level);
    delete level;
}

static DevCsr *pick(sa_gpu_level *lev, int which)
{
    switch (which)
    {
    case SA_GPU_MAT_A:
        return lev->A;
    case SA_GPU_MAT_PTENT:
        return lev->have_tent ? &lev->Ptent : nullptr;
    case SA_GPU_MAT_P:
        return lev->have_P ? &lev->P : nullptr;
    case SA_GPU_MAT_R:
        return lev->have_P ? &lev->R : nullptr;
    case SA_GPU_MAT_AC:
        return lev->have_Ac ? &lev->Ac : nullptr;
    }
    return nullptr;
}

extern "C" int sa_gpu_get_csr_sizes(sa_gpu_level *lev, int which, int *rows, int *cols, int *nnz)
{
    SA_API_BEGIN
    DevCsr *M = pick(lev, which);
    if (!M)
        SA_FAIL("sa_gpu_get_csr_sizes: matrix %d not available", which);
    *rows = M->rows;
    *cols = M->cols;
    *nnz = M->nnz;
    SA_API_END
}

extern "C" int sa_gpu_get_csr(sa_gpu_level *lev, int which, int *I, int *J, double *data)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    DevCsr *M = pick(lev, which);
    if (!M)
        SA_FAIL("sa_gpu_get_csr: matrix %d not available", which);
    cudaStream_t st = lev->ctx->stream;
    if (I)
        M->I.download(I, (size_t)M->rows + 1, st);
    if (J)
        M->J.download(J, M->nnz, st);
    if (data)
        M->A.download(data, M->nnz, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

extern "C" int sa_gpu_get_Dinv_neg(sa_gpu_level *lev, double *dinv_neg)
{
    SA_API_BEGIN
    if (!lev->have_Dinv)
        SA_FAIL("sa_gpu_get_Dinv_neg: not built");
    lev->Dinv_neg.download(dinv_neg, lev->ND, lev->ctx->stream);
    SA_CUDA(cudaStreamSynchronize(lev->ctx->stream));
    SA_API_END
}

extern "C" int sa_gpu_spmv(sa_gpu_level *lev, int which, const double *x, double *y)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    DevCsr *M = pick(lev, which);
    if (!M)
        SA_FAIL("sa_gpu_spmv: matrix %d not available", which);
    cudaStream_t st = lev->ctx->stream;
    lev->vx.upload(x, M->cols, st);
    lev->vy.ensure(M->rows);
    dev_spmv(lev->ctx, *M, lev->vx.p, lev->vy.p);
    lev->vy.download(y, M->rows, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

extern "C" int sa_gpu_poly_smooth(sa_gpu_level *lev, const double *b, double *x, int degree,
                                  const double *roots)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    if (!lev->have_Dinv)
        SA_FAIL("sa_gpu_poly_smooth: sa_gpu_build_Dinv_neg has not been called");
    cudaStream_t st = lev->ctx->stream;
    const int n = lev->ND;
    lev->vb.upload(b, n, st);
    lev->vx.upload(x, n, st);
    lev->vy.ensure(n);
    double *cur = lev->vx.p, *alt = lev->vy.p;
    for (int i = 0; i < degree; ++i)
    {
        dev_smoother_step(lev->ctx, *lev->A, lev->Dinv_neg.p, lev->vb.p, cur, alt, 1. / roots[i], 0);
        std::swap(cur, alt);
    }
    SA_CUDA(cudaMemcpyAsync(x, cur, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

/* ------------------------------------------------------- micro-benchmarks */

__global__ void k_fill(double *x, int n, double v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        x[i] = v + 1e-3 * (i & 1023);
}

extern "C" double sa_gpu_bench_spmv(sa_gpu_level *lev, int which, int reps)
{
    sa_level_ready(lev);
    try
    {
        DevCsr *M = pick(lev, which);
        if (!M)
            return -1.;
        sa_gpu_ctx *ctx = lev->ctx;
        lev->vx.ensure(std::max(M->cols, M->rows));
        lev->vy.ensure(std::max(M->cols, M->rows));
        SA_LAUNCH(ctx, k_fill, (M->cols + 255) / 256, 256, 0, lev->vx.p, M->cols, 1.0);
        for (int i = 0; i < 3; ++i)
            dev_spmv(ctx, *M, lev->vx.p, lev->vy.p);
        SA_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int i = 0; i < reps; ++i)
            dev_spmv(ctx, *M, lev->vx.p, lev->vy.p);
        SA_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        SA_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        SA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        return (double)ms / reps;
    }
    catch (...)
    {
        return -1.;
    }
}

extern "C" double sa_gpu_bench_smoother(sa_gpu_level *lev, int reps)
{
    sa_level_ready(lev);
    try
    {
        if (!lev->have_Dinv || !lev->A)
            return -1.;
        sa_gpu_ctx *ctx = lev->ctx;
        const int n = lev->ND;
        lev->vx.ensure(n);
        lev->vy.ensure(n);
        lev->vb.ensure(n);
        SA_LAUNCH(ctx, k_fill, (n + 255) / 256, 256, 0, lev->vx.p, n, 0.0);
        SA_LAUNCH(ctx, k_fill, (n + 255) / 256, 256, 0, lev->vb.p, n, 1.0);
        double *cur = lev->vx.p, *alt = lev->vy.p;
        for (int i = 0; i < 3; ++i)
        {
            dev_smoother_step(ctx, *lev->A, lev->Dinv_neg.p, lev->vb.p, cur, alt, 0.5, 0);
            std::swap(cur, alt);
        }
        SA_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int i = 0; i < reps; ++i)
        {
            dev_smoother_step(ctx, *lev->A, lev->Dinv_neg.p, lev->vb.p, cur, alt, 0.5, 0);
            std::swap(cur, alt);
        }
        SA_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        SA_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        SA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        return (double)ms / reps;
    }
    catch (...)
    {
        return -1.;
    }
}

// 8 independent DFMA chains per thread
__global__ void k_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4.,
           a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i)
    {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" double sa_gpu_bench_fp64_peak(sa_gpu_ctx *ctx)
{
    try
    {
        const int blocks = ctx->num_sms * 8, threads = 256, iters = 1 << 14;
        DevBuf<double> out;
        out.alloc((size_t)blocks * threads);
        SA_LAUNCH(ctx, k_fp64_peak, blocks, threads, 0, out.p, iters);
        double best = 0.;
        for (int rep = 0; rep < 5; ++rep)
        {
            SA_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
            SA_LAUNCH(ctx, k_fp64_peak, blocks, threads, 0, out.p, iters);
            SA_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
            SA_CUDA(cudaEventSynchronize(ctx->ev1));
            float ms = 0.f;
            SA_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            const double flops = 2. * 8. * iters * (double)blocks * threads;
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        return best;
    }
    catch (...)
    {
        return -1.;
    }
}

/* pinned host memory helpers for callers that stage large inputs (bench e2e path) */
extern "C" int sa_gpu_host_register(const void *p, size_t bytes)
{
    if (!p || !bytes)
        return 1;
    cudaError_t e = cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        return 1;
    }
    return 0;
}

extern "C" int sa_gpu_host_unregister(const void *p)
{
    if (!p)
        return 1;
    cudaError_t e = cudaHostUnregister(const_cast<void *>(p));
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        return 1;
    }
    return 0;
}

/* ---- device-pointer entry points for row-partitioned (multi-GPU) solves ---- */

extern "C" int sa_gpu_level_dev_csr(sa_gpu_level *lev, int which, const int **I, const int **J,
                                    const double **A, int *rows, int *cols, int *nnz)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    DevCsr *M = pick(lev, which);
    if (!M)
        SA_FAIL("sa_gpu_level_dev_csr: matrix %d not available", which);
    *I = M->I.p;
    *J = M->J.p;
    *A = M->A.p;
    *rows = M->rows;
    *cols = M->cols;
    *nnz = M->nnz;
    SA_API_END
}

extern "C" int sa_gpu_level_dev_dinv(sa_gpu_level *lev, const double **dinv)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    if (!lev->have_Dinv)
        SA_FAIL("sa_gpu_level_dev_dinv: not built");
    *dinv = lev->Dinv_neg.p;
    SA_API_END
}

extern "C" int sa_gpu_dev_spmv(sa_gpu_ctx *ctx, int mode, int nrows, double avg_nnz_per_row,
                               const int *I_row0, const int *J, const double *A, const double *x,
                               const double *xrow, const double *b, const double *dinv,
                               double mult, double *y)
{
    SA_API_BEGIN
    dev_spmv_rows(ctx, mode, nrows, avg_nnz_per_row, I_row0, J, A, x, xrow, b, dinv, mult, y);
    SA_API_END
}
