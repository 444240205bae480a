// Row-partitioned multi-GPU solve in the library (SURVEY.md section 8e, row "Solve a15-a17"):
// one process per GPU, every rank holds the hierarchy and owns a contiguous row range of every
// level.  Each SpMV / smoother step / restriction / prolongation (tg_cycle_atb,
// amg/src/tg.cpp:91-132; smpr_compute_poly, amg/inc/smpr.hpp:319-339) runs the CSR kernels on the
// rank's rows after a halo exchange of exactly the off-rank entries those rows reference:
// packed boundary lists, one grouped ncclSend / ncclRecv per exchange over NVLink (contiguous
// pieces -- the planes of a slab partition -- go straight from / into the vector, no pack).  PCG
// (kalchev_pcg, amg/src/mfem_addons.cpp:106-248) keeps alpha and beta on the device: the dots
// are reduced into device scalars by ncclAllReduce, the update kernels read them there, and the
// host reads ONE scalar per iteration (the convergence test).  The coarsest system is solved
// replicated after an all-reduce of the restricted residual.
//
// NCCL is resolved at run time (the copy already loaded by the process -- torch's -- else
// SA_NCCL_LIB, else libnccl.so.2): the library has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>

#include "nccl_dl.cuh"
#include "solve_internal.cuh"

static NcclApi g_nccl;

const NcclApi &sa_nccl()
{
    if (g_nccl.lib)
        return g_nccl;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h && getenv("SA_NCCL_LIB"))
        h = dlopen(getenv("SA_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
    if (!h)
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
        SA_FAIL("NCCL not found (libnccl.so.2; set SA_NCCL_LIB): %s", dlerror());
#define SA_NCCL_SYM(field, name)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                    \
    if (!g_nccl.field)                                                             \
        SA_FAIL("NCCL symbol %s missing", name);
    SA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    SA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    SA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SA_NCCL_SYM(Send, "ncclSend")
    SA_NCCL_SYM(Recv, "ncclRecv")
    SA_NCCL_SYM(GroupStart, "ncclGroupStart")
    SA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    SA_NCCL_SYM(AllReduce, "ncclAllReduce")
    SA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SA_NCCL_SYM
    g_nccl.lib = h;
    return g_nccl;
}

static inline const NcclApi &nccl() { return sa_nccl(); }

/* ---- halo plan (host; no GPU needed: tests/test_dist_plan.py) ---------------------------- */

/* Rows [row_part[rank], row_part[rank + 1]) of the CSR pattern (I, J) reference columns owned by
   other ranks (col_part).  need[q] = sorted unique columns of rank q that `rank` reads. */
static void halo_need(const int *I, const int *J, int rank, int nranks, const int *row_part,
                      const int *col_part, std::vector<std::vector<int>> &need)
{
    need.assign(nranks, std::vector<int>());
    const int r0 = row_part[rank], r1 = row_part[rank + 1];
    const int c0 = col_part[rank], c1 = col_part[rank + 1];
    std::vector<int> cols;
    cols.reserve((size_t)(I[r1] - I[r0]) / 4 + 16);
    for (int p = I[r0]; p < I[r1]; ++p)
        if (J[p] < c0 || J[p] >= c1)
            cols.push_back(J[p]);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    int q = 0;
    for (size_t t = 0; t < cols.size(); ++t)
    {
        while (cols[t] >= col_part[q + 1])
            ++q;
        need[q].push_back(cols[t]);
    }
}

/* C ABI: the exchange lists of `rank` for one matrix.  Every rank holds the whole pattern, so
   both directions are computed locally: recv = what my rows need, send[q] = what rank q's rows
   need from my range.  Output: counts per peer and the concatenated index lists (peer-major,
   ascending); capacities are checked.  Returns 0, or 2 when a capacity is too small (the needed
   sizes are still written to n_send / n_recv). */
extern "C" int sa_gpu_halo_plan(int rows, const int *I, const int *J, int nranks, int rank,
                                const int *row_part, const int *col_part, int *send_cnt,
                                int *recv_cnt, int *send_idx, int send_cap, int *recv_idx,
                                int recv_cap, int *n_send, int *n_recv)
{
    SA_API_BEGIN
    (void)rows;
    std::vector<std::vector<int>> mine, other;
    halo_need(I, J, rank, nranks, row_part, col_part, mine);
    int ns = 0, nr = 0;
    std::vector<std::vector<int>> sends(nranks);
    for (int q = 0; q < nranks; ++q)
    {
        recv_cnt[q] = (int)mine[q].size();
        nr += recv_cnt[q];
        if (q == rank)
        {
            send_cnt[q] = 0;
            continue;
        }
        halo_need(I, J, q, nranks, row_part, col_part, other);
        sends[q].swap(other[rank]);
        send_cnt[q] = (int)sends[q].size();
        ns += send_cnt[q];
    }
    *n_send = ns;
    *n_recv = nr;
    if (ns > send_cap || nr > recv_cap)
        return 2;
    int so = 0, ro = 0;
    for (int q = 0; q < nranks; ++q)
    {
        std::copy(sends[q].begin(), sends[q].end(), send_idx + so);
        so += send_cnt[q];
        std::copy(mine[q].begin(), mine[q].end(), recv_idx + ro);
        ro += recv_cnt[q];
    }
    SA_API_END
}

/* ---- device side ------------------------------------------------------------------------- */
namespace
{
struct DistMat
{
    const DevCsr *M = nullptr;
    int r0 = 0, r1 = 0; // my rows
    double avg = 1.;    // nonzeros per row of my rows
    std::vector<int> send_cnt, recv_cnt, send_off, recv_off;
    std::vector<char> send_contig, recv_contig; // the piece is one contiguous index range
    std::vector<int> send_lo, recv_lo;
    DevBuf<int> send_idx, recv_idx;
    int nsend = 0, nrecv = 0;
    bool any = false;
};

struct DistLevel
{
    sa_gpu_level *lev = nullptr;
    int n = 0;
    DistMat A, P, R;
    DevBuf<double> b, xa, xb, r;
};

__global__ void k_pack(int n, const int *idx, const double *x, double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        buf[i] = x[idx[i]];
}
__global__ void k_unpack(int n, const int *idx, const double *buf, double *x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        x[idx[i]] = buf[i];
}

// per-block partial sums of a_i b_i over [0, n) (fixed order: deterministic)
__global__ void k_ddot(int n, const double *a, const double *b, double *partials)
{
    __shared__ double sh[32];
    double s = 0.;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        s += a[i] * b[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.;
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0)
            partials[blockIdx.x] = t;
    }
}
__global__ void k_ddot_final(int nblocks, const double *partials, double *out)
{
    double s = 0.;
    for (int i = threadIdx.x; i < nblocks; i += 32)
        s += partials[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0)
        *out = s;
}
// x += alpha d, r -= alpha z with alpha = *nom / *den (device scalars); rb = r (next rhs)
__global__ void k_dpcg_update(int n, const double *nom, const double *den, const double *d,
                              const double *z, double *x, double *r, double *rb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double alpha = *nom / *den;
    x[i] = x[i] + alpha * d[i];
    const double ri = r[i] - alpha * z[i];
    r[i] = ri;
    rb[i] = ri;
}
// d = z + beta d with beta = *betanom / *nom
__global__ void k_dpcg_dir(int n, const double *betanom, const double *nom, const double *z,
                           double *d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double beta = *betanom / *nom;
    d[i] = z[i] + beta * d[i];
}
} // namespace

struct sa_gpu_dist_solver
{
    sa_gpu_solver *S = nullptr;
    sa_gpu_comm *C = nullptr;
    sa_gpu_ctx *ctx = nullptr;
    std::vector<DistLevel *> L;
    std::vector<std::vector<int>> part; // per level (+ coarsest): row ranges of the ranks
    DevBuf<double> sendbuf, recvbuf;
    DevBuf<double> bc, xc;
    DevBuf<double> pb, px, pr, pd, pz, dots, partials;
    long halo_calls = 0, halo_doubles = 0;
    ~sa_gpu_dist_solver()
    {
        for (size_t i = 0; i < L.size(); ++i)
            delete L[i];
    }
};

namespace
{
void make_plan(sa_gpu_dist_solver *D, DistMat &m, const DevCsr *M, const std::vector<int> &rp,
               const std::vector<int> &cp)
{
    sa_gpu_ctx *ctx = D->ctx;
    const int nr = D->C->nranks, me = D->C->rank;
    m.M = M;
    m.r0 = rp[me];
    m.r1 = rp[me + 1];
    std::vector<int> hI((size_t)M->rows + 1), hJ((size_t)std::max(1, M->nnz));
    M->I.download(hI.data(), (size_t)M->rows + 1, ctx->stream);
    M->J.download(hJ.data(), (size_t)M->nnz, ctx->stream);
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    m.avg = (double)(hI[m.r1] - hI[m.r0]) / std::max(1, m.r1 - m.r0);
    m.send_cnt.assign(nr, 0);
    m.recv_cnt.assign(nr, 0);
    int ns = 0, nrv = 0;
    std::vector<int> sidx(1), ridx(1);
    int rc = sa_gpu_halo_plan(M->rows, hI.data(), hJ.data(), nr, me, rp.data(), cp.data(),
                              m.send_cnt.data(), m.recv_cnt.data(), sidx.data(), 0, ridx.data(), 0,
                              &ns, &nrv);
    if (rc == 1)
        throw std::runtime_error("sa_gpu");
    sidx.resize((size_t)std::max(1, ns));
    ridx.resize((size_t)std::max(1, nrv));
    rc = sa_gpu_halo_plan(M->rows, hI.data(), hJ.data(), nr, me, rp.data(), cp.data(),
                          m.send_cnt.data(), m.recv_cnt.data(), sidx.data(), ns, ridx.data(), nrv, &ns,
                          &nrv);
    if (rc != 0)
        SA_FAIL("halo plan failed");
    m.nsend = ns;
    m.nrecv = nrv;
    m.any = ns > 0 || nrv > 0;
    m.send_off.assign(nr + 1, 0);
    m.recv_off.assign(nr + 1, 0);
    m.send_contig.assign(nr, 0);
    m.recv_contig.assign(nr, 0);
    m.send_lo.assign(nr, 0);
    m.recv_lo.assign(nr, 0);
    for (int q = 0; q < nr; ++q)
    {
        m.send_off[q + 1] = m.send_off[q] + m.send_cnt[q];
        m.recv_off[q + 1] = m.recv_off[q] + m.recv_cnt[q];
        if (m.send_cnt[q])
        {
            const int *s = sidx.data() + m.send_off[q];
            m.send_lo[q] = s[0];
            m.send_contig[q] = (s[m.send_cnt[q] - 1] - s[0] + 1 == m.send_cnt[q]);
        }
        if (m.recv_cnt[q])
        {
            const int *s = ridx.data() + m.recv_off[q];
            m.recv_lo[q] = s[0];
            m.recv_contig[q] = (s[m.recv_cnt[q] - 1] - s[0] + 1 == m.recv_cnt[q]);
        }
    }
    m.send_idx.upload(sidx.data(), (size_t)std::max(1, ns), ctx->stream);
    m.recv_idx.upload(ridx.data(), (size_t)std::max(1, nrv), ctx->stream);
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    D->sendbuf.ensure((size_t)std::max(1, ns));
    D->recvbuf.ensure((size_t)std::max(1, nrv));
}

/* brings the off-rank entries of x that the rows of m reference up to date */
void halo(sa_gpu_dist_solver *D, const DistMat &m, double *x)
{
    if (!m.any || D->C->nranks == 1)
        return;
    sa_gpu_ctx *ctx = D->ctx;
    const int nr = D->C->nranks;
    const NcclApi &N = nccl();
    // pack the non-contiguous pieces (one launch each: there are few)
    for (int q = 0; q < nr; ++q)
        if (m.send_cnt[q] && !m.send_contig[q])
            SA_LAUNCH(ctx, k_pack, (m.send_cnt[q] + 255) / 256, 256, 0, m.send_cnt[q],
                      m.send_idx.p + m.send_off[q], x, D->sendbuf.p + m.send_off[q]);
    SA_NCCL(N.GroupStart());
    for (int q = 0; q < nr; ++q)
    {
        if (m.send_cnt[q])
            SA_NCCL(N.Send(m.send_contig[q] ? (const void *)(x + m.send_lo[q])
                                            : (const void *)(D->sendbuf.p + m.send_off[q]),
                           (size_t)m.send_cnt[q], ncclDouble, q, D->C->comm, ctx->stream));
        if (m.recv_cnt[q])
            SA_NCCL(N.Recv(m.recv_contig[q] ? (void *)(x + m.recv_lo[q])
                                            : (void *)(D->recvbuf.p + m.recv_off[q]),
                           (size_t)m.recv_cnt[q], ncclDouble, q, D->C->comm, ctx->stream));
    }
    SA_NCCL(N.GroupEnd());
    for (int q = 0; q < nr; ++q)
        if (m.recv_cnt[q] && !m.recv_contig[q])
            SA_LAUNCH(ctx, k_unpack, (m.recv_cnt[q] + 255) / 256, 256, 0, m.recv_cnt[q],
                      m.recv_idx.p + m.recv_off[q], D->recvbuf.p + m.recv_off[q], x);
    D->halo_calls++;
    D->halo_doubles += m.nsend;
}

/* mode (dev_spmv_rows): 0 y = M x, 1 y = b - M x, 2 y += M x, 3 smoother step, 4 smoother step
   from x = 0; on the rank's rows */
void rows_op(sa_gpu_dist_solver *D, const DistMat &m, int mode, const double *x, double *y,
             const double *xrow, const double *b, const double *dinv, double mult)
{
    const int n = m.r1 - m.r0;
    if (n <= 0)
        return;
    dev_spmv_rows(D->ctx, mode, n, m.avg, m.M->I.p + m.r0, m.M->J.p, m.M->A.p, x,
                  xrow ? xrow + m.r0 : nullptr, b ? b + m.r0 : nullptr, dinv ? dinv + m.r0 : nullptr,
                  mult, y + m.r0);
}

void dist_smooth(sa_gpu_dist_solver *D, DistLevel &L, double **xcur, double **xalt, bool x_is_zero)
{
    const sa_gpu_solver *S = D->S;
    for (int i = 0; i < S->degree; ++i)
    {
        const double mult = 1. / S->roots[i];
        if (x_is_zero && i == 0)
            rows_op(D, L.A, 4, *xcur, *xalt, *xcur, L.b.p, L.lev->Dinv_neg.p, mult);
        else
        {
            halo(D, L.A, *xcur);
            rows_op(D, L.A, 3, *xcur, *xalt, *xcur, L.b.p, L.lev->Dinv_neg.p, mult);
        }
        std::swap(*xcur, *xalt);
    }
}

/* tg_cycle_atb on level l: rhs L.b (own rows valid); returns the iterate (own rows valid) */
double *dist_vcycle(sa_gpu_dist_solver *D, int l)
{
    sa_gpu_ctx *ctx = D->ctx;
    DistLevel &L = *D->L[l];
    double *xcur = L.xa.p, *xalt = L.xb.p;
    dist_smooth(D, L, &xcur, &xalt, true);
    halo(D, L.A, xcur);
    rows_op(D, L.A, 1, xcur, L.r.p, nullptr, L.b.p, nullptr, 0.);
    halo(D, L.R, L.r.p);
    double *xc;
    if (l + 1 < (int)D->L.size())
    {
        DistLevel &Lc = *D->L[l + 1];
        rows_op(D, L.R, 0, L.r.p, Lc.b.p, nullptr, nullptr, nullptr, 0.);
        xc = dist_vcycle(D, l + 1);
    }
    else
    {
        const int nc = D->S->nc;
        if (D->C->nranks > 1)
            SA_CUDA(cudaMemsetAsync(D->bc.p, 0, (size_t)nc * sizeof(double), ctx->stream));
        rows_op(D, L.R, 0, L.r.p, D->bc.p, nullptr, nullptr, nullptr, 0.);
        if (D->C->nranks > 1)
            SA_NCCL(nccl().AllReduce(D->bc.p, D->bc.p, (size_t)nc, ncclDouble, ncclSum, D->C->comm,
                                     ctx->stream));
        if (sa_gpu_solver_dev_coarse(D->S, D->bc.p, D->xc.p))
            throw std::runtime_error("sa_gpu");
        xc = D->xc.p;
    }
    if (l + 1 < (int)D->L.size()) // (the coarsest solution is replicated)
        halo(D, L.P, xc);
    rows_op(D, L.P, 2, xc, xcur, nullptr, nullptr, nullptr, 0.);
    dist_smooth(D, L, &xcur, &xalt, false);
    return xcur;
}

/* (a, b) over the rank's rows of level 0, all-reduced into dots[slot] */
void dist_dot(sa_gpu_dist_solver *D, const double *a, const double *b, int slot)
{
    sa_gpu_ctx *ctx = D->ctx;
    const int r0 = D->part[0][D->C->rank], n = D->part[0][D->C->rank + 1] - r0;
    const int blocks = std::max(1, std::min(ctx->num_sms * 4, (n + 255) / 256));
    D->partials.ensure(blocks);
    SA_LAUNCH(ctx, k_ddot, blocks, 256, 0, n, a + r0, b + r0, D->partials.p);
    SA_LAUNCH(ctx, k_ddot_final, 1, 32, 0, blocks, D->partials.p, D->dots.p + slot);
    if (D->C->nranks > 1)
        SA_NCCL(nccl().AllReduce(D->dots.p + slot, D->dots.p + slot, 1, ncclDouble, ncclSum,
                                 D->C->comm, ctx->stream));
}
double read_scalar(sa_gpu_dist_solver *D, int slot)
{
    double h = 0.;
    SA_CUDA(cudaMemcpyAsync(&h, D->dots.p + slot, sizeof(double), cudaMemcpyDeviceToHost,
                            D->ctx->stream));
    SA_CUDA(cudaStreamSynchronize(D->ctx->stream));
    return h;
}
} // namespace

extern "C" int sa_gpu_nccl_unique_id(void *id128)
{
    SA_API_BEGIN
    ncclUniqueId id;
    SA_NCCL(nccl().GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id128, &id, 128);
    SA_API_END
}

extern "C" int sa_gpu_comm_create(sa_gpu_ctx *ctx, const void *id128, int nranks, int rank,
                                  sa_gpu_comm **out)
{
    SA_API_BEGIN
    *out = nullptr;
    sa_gpu_comm *C = new sa_gpu_comm;
    C->ctx = ctx;
    C->rank = rank;
    C->nranks = nranks;
    if (nranks > 1)
    {
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        SA_CUDA(cudaSetDevice(ctx->device));
        ncclResult_t r = nccl().CommInitRank(&C->comm, nranks, id, rank);
        if (r != ncclSuccess)
        {
            delete C;
            SA_FAIL("ncclCommInitRank: %s", nccl().GetErrorString(r));
        }
        // NCCL connects channels lazily (first collective / first send-recv per peer: 0.5 - 2 s
        // measured): do it here, not inside the first timed stage
        // (a few message sizes: the protocol / algorithm NCCL picks, and with it the connections
        // it sets up on first use, depend on the size -- 8-GPU run: 40 ms inside the first small
        // all-reduce of a level that followed a warm-up with one size only)
        const size_t sizes[3] = {16, (size_t)1 << 15, (size_t)1 << 21}; // doubles
        DevBuf<double> w;
        w.alloc(sizes[2] * 2 + (size_t)nranks * sizes[1]);
        w.zero(ctx->stream);
        for (int si = 0; si < 3; ++si)
        {
            SA_NCCL(nccl().AllReduce(w.p, w.p, sizes[si], ncclDouble, ncclSum, C->comm, ctx->stream));
            SA_NCCL(nccl().AllReduce(w.p, w.p, sizes[si], ncclInt32, ncclSum, C->comm, ctx->stream));
        }
        for (int si = 0; si < 2; ++si)
        {
            SA_NCCL(nccl().GroupStart());
            for (int q = 0; q < nranks; ++q)
                if (q != rank)
                {
                    SA_NCCL(nccl().Send(w.p, sizes[si], ncclDouble, q, C->comm, ctx->stream));
                    SA_NCCL(nccl().Recv(w.p + sizes[2] * 2 + (size_t)q * sizes[1], sizes[si], ncclDouble, q,
                                        C->comm, ctx->stream));
                }
            SA_NCCL(nccl().GroupEnd());
        }
        SA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *out = C;
    SA_API_END
}

extern "C" void sa_gpu_comm_destroy(sa_gpu_comm *C)
{
    if (!C)
        return;
    if (C->comm)
        g_nccl.CommDestroy(C->comm);
    delete C;
}

extern "C" int sa_gpu_dist_solver_create(sa_gpu_solver *S, sa_gpu_comm *C, sa_gpu_dist_solver **out)
{
    SA_API_BEGIN
    *out = nullptr;
    sa_gpu_ctx *ctx = S->ctx;
    cudaStream_t st = ctx->stream;
    for (size_t l = 0; l < S->L.size(); ++l)
        if (S->L[l]->pre || S->L[l]->post)
            SA_FAIL("sa_gpu_dist_solver_create: user smoothers are not supported by the distributed solve");
    sa_gpu_dist_solver *D = new sa_gpu_dist_solver;
    struct Guard
    {
        sa_gpu_dist_solver *p;
        ~Guard() { delete p; }
    } guard{D};
    D->S = S;
    D->C = C;
    D->ctx = ctx;
    const int nl = (int)S->L.size(), nr = C->nranks;
    std::vector<int> sizes(nl + 1);
    for (int l = 0; l < nl; ++l)
        sizes[l] = S->L[l]->lev->ND;
    sizes[nl] = S->nc;
    D->part.resize(nl + 1);
    for (int l = 0; l <= nl; ++l)
    {
        D->part[l].resize(nr + 1);
        for (int q = 0; q <= nr; ++q)
            D->part[l][q] = (int)(((int64_t)sizes[l] * q) / nr);
    }
    for (int l = 0; l < nl; ++l)
    {
        DistLevel *L = new DistLevel;
        D->L.push_back(L);
        L->lev = S->L[l]->lev;
        L->n = sizes[l];
        sa_level_ready(L->lev);
        if (!L->lev->have_Dinv || !L->lev->have_P)
            SA_FAIL("sa_gpu_dist_solver_create: level %d has no smoother diagonal / prolongator", l);
        make_plan(D, L->A, L->lev->A, D->part[l], D->part[l]);
        make_plan(D, L->P, &L->lev->P, D->part[l], D->part[l + 1]);
        make_plan(D, L->R, &L->lev->R, D->part[l + 1], D->part[l]);
        L->b.alloc(sizes[l]);
        L->xa.alloc(sizes[l]);
        L->xb.alloc(sizes[l]);
        L->r.alloc(sizes[l]);
        L->b.zero(st);
        L->xa.zero(st);
        L->xb.zero(st);
        L->r.zero(st);
    }
    D->bc.alloc(std::max(1, S->nc));
    D->xc.alloc(std::max(1, S->nc));
    D->bc.zero(st);
    D->xc.zero(st);
    const int n0 = sizes[0];
    D->pb.alloc(n0);
    D->px.alloc(n0);
    D->pr.alloc(n0);
    D->pd.alloc(n0);
    D->pz.alloc(n0);
    D->dots.alloc(8);
    D->dots.zero(st);
    SA_CUDA(cudaStreamSynchronize(st));
    guard.p = nullptr;
    *out = D;
    SA_API_END
}

extern "C" void sa_gpu_dist_solver_destroy(sa_gpu_dist_solver *D) { delete D; }

/* kalchev_pcg (amg/src/mfem_addons.cpp:106-248), x0 = 0, b and x full-length host vectors (x:
   the rank's rows are written, the rest left untouched unless gather != 0, which all-reduces
   the solution so that every rank returns the whole vector).  iters < 0: not converged / SPD
   breakdown, as the reference returns it. */
extern "C" int sa_gpu_dist_pcg(sa_gpu_dist_solver *D, const double *b, double *x, int maxiter,
                               double rtol, double atol, int gather, int *iters, double *brr_hist,
                               int hist_cap, int *hist_len, double *solve_seconds)
{
    SA_API_BEGIN
    sa_gpu_ctx *ctx = D->ctx;
    cudaStream_t st = ctx->stream;
    const int me = D->C->rank;
    const int n = D->L[0]->n, r0 = D->part[0][me], r1 = D->part[0][me + 1], nloc = r1 - r0;
    const int tb = 256, gb = std::max(1, (nloc + tb - 1) / tb);
    DistLevel &L0 = *D->L[0];
    double *xv = D->px.p, *r = D->pr.p, *d = D->pd.p, *z = D->pz.p;
    D->pb.upload(b, n, st);
    D->px.zero(st);
    cudaEvent_t e0, e1;
    SA_CUDA(cudaEventCreate(&e0));
    SA_CUDA(cudaEventCreate(&e1));
    SA_CUDA(cudaEventRecord(e0, st));
    int hl = 0, it = 0;
    enum { NOM = 0, DEN = 1, BETANOM = 2 };
    int s_nom = NOM, s_beta = BETANOM;
    auto precond = [&](const double *rhs_is_in_L0b) {
        (void)rhs_is_in_L0b;
        double *res = dist_vcycle(D, 0);
        if (nloc)
            SA_CUDA(cudaMemcpyAsync(z + r0, res + r0, (size_t)nloc * sizeof(double),
                                    cudaMemcpyDeviceToDevice, st));
    };
    // r = b (x0 = 0); z = B r; d = z
    if (nloc)
    {
        SA_CUDA(cudaMemcpyAsync(r + r0, D->pb.p + r0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, st));
        SA_CUDA(cudaMemcpyAsync(L0.b.p + r0, D->pb.p + r0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    precond(L0.b.p);
    if (nloc)
        SA_CUDA(cudaMemcpyAsync(d + r0, z + r0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, st));
    dist_dot(D, z, r, s_nom);
    double nom = read_scalar(D, s_nom);
    if (brr_hist && hl < hist_cap)
        brr_hist[hl++] = nom;
    double r0tol = nom * rtol;
    if (r0tol < atol)
        r0tol = atol;
    if (nom < r0tol)
        it = -1;
    else
    {
        halo(D, L0.A, d);
        rows_op(D, L0.A, 0, d, z, nullptr, nullptr, nullptr, 0.);
        dist_dot(D, z, d, DEN);
        const double den = read_scalar(D, DEN);
        if (den == 0.)
            it = -1;
        else
        {
            int i;
            for (i = 1; i <= maxiter; ++i)
            {
                // x += alpha d; r -= alpha z; rhs of the preconditioner = r
                SA_LAUNCH(ctx, k_dpcg_update, gb, tb, 0, nloc, D->dots.p + s_nom, D->dots.p + DEN,
                          d + r0, z + r0, xv + r0, r + r0, L0.b.p + r0);
                precond(L0.b.p);
                dist_dot(D, r, z, s_beta);
                const double betanom = read_scalar(D, s_beta); // the one host read per iteration
                if (brr_hist && hl < hist_cap)
                    brr_hist[hl++] = betanom;
                if (betanom < 0.)
                {
                    it = -i;
                    break;
                }
                if (betanom < r0tol)
                {
                    it = i;
                    break;
                }
                SA_LAUNCH(ctx, k_dpcg_dir, gb, tb, 0, nloc, D->dots.p + s_beta, D->dots.p + s_nom,
                          z + r0, d + r0);
                halo(D, L0.A, d);
                rows_op(D, L0.A, 0, d, z, nullptr, nullptr, nullptr, 0.);
                dist_dot(D, d, z, DEN);
                std::swap(s_nom, s_beta);
            }
            if (i > maxiter)
                it = -(i - 1);
        }
    }
    SA_CUDA(cudaEventRecord(e1, st));
    SA_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    SA_CUDA(cudaEventDestroy(e0));
    SA_CUDA(cudaEventDestroy(e1));
    if (solve_seconds)
        *solve_seconds = ms * 1e-3;
    if (gather && D->C->nranks > 1)
    {
        // rows of other ranks are still zero (x0 = 0 and only own rows are updated)
        SA_NCCL(nccl().AllReduce(xv, xv, (size_t)n, ncclDouble, ncclSum, D->C->comm, st));
        D->px.download(x, n, st);
    }
    else if (nloc)
        SA_CUDA(cudaMemcpyAsync(x + r0, xv + r0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost, st));
    SA_CUDA(cudaStreamSynchronize(st));
    *iters = it;
    if (hist_len)
        *hist_len = hl;
    SA_API_END
}

extern "C" int sa_gpu_dist_solver_stats(sa_gpu_dist_solver *D, int *row_begin, int *row_end,
                                        long *halo_calls, long *halo_doubles)
{
    SA_API_BEGIN
    *row_begin = D->part[0][D->C->rank];
    *row_end = D->part[0][D->C->rank + 1];
    *halo_calls = D->halo_calls;
    *halo_doubles = D->halo_doubles;
    SA_API_END
}

/* sizes of one level of the distributed solver: rows, nnz(A), nnz(P); returns 1 past the last */
extern "C" int sa_gpu_dist_solver_level_info(sa_gpu_dist_solver *D, int level, int *rows, long *nnz_A,
                                             long *nnz_P)
{
    SA_API_BEGIN
    if (level < 0 || level >= (int)D->L.size())
        SA_FAIL("sa_gpu_dist_solver_level_info: no level %d", level);
    *rows = D->L[level]->n;
    *nnz_A = D->L[level]->A.M->nnz;
    *nnz_P = D->L[level]->P.M->nnz;
    SA_API_END
}
