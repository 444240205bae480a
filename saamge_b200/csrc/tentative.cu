// Tentative prolongator on the device:
//   k_mis_svd       : one warp per MIS -- restrict the eigenvectors of every AE that
//                     contains the MIS to the MIS's dofs (agg_restrict_to_agg_enforce,
//                     amg/src/aggregates.cpp:1143-1179; concatenation order of
//                     CommunicateEigenvectors, amg/src/contrib.cpp:501-546), boundary
//                     filter (contrib_filter_boundary, amg/src/contrib.cpp:102-163),
//                     column normalisation (xpack_svd_dense_arr, amg/src/xpacks.cpp:533-559),
//                     thin SVD by one-sided Jacobi (the reference calls dgesvd 'S','N';
//                     only U and sigma are needed), rank cut sigma_i > 1e-10 sigma_0
//                     (xpack_orth_set, amg/src/xpacks.cpp:591-620)
//   k_mis_finalize  : sort by sigma, write mis_tent_interps blocks
//   k_ptent_count / k_ptent_fill : contrib_tent_insert_simple + contrib_tent_finalize
//                     (amg/src/contrib.cpp:170-194, 73-95) -> CSR tentative P
//   k_coarse_elmat  : ElementMatrixParallelCoarse::GetMatrix (amg/src/elmat.cpp:105-195)
#include <algorithm>
#include <cfloat>

#include "assemble.cuh"
#include "nccl_dl.cuh"
#include "sa_gpu_internal.cuh"

namespace
{

__device__ __forceinline__ double wsum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct MisWork
{
    const int64_t *xoff; // MIS -> offset of its s x c work matrix
    const int *coff;     // MIS -> offset of its c singular values
    double *X;
    double *sig;
    int *ncols; // columns kept after filtering (c')
    int *ncd;   // numcoarsedof
};

/* Sharded form (sa_gpu_dist_tentative_P): the warp works on MIS mis_list[w] (the MISes this rank
   owns; W.xoff / W.coff are indexed by w), and a (MIS, AE) pair whose AE lives on another rank
   reads its s x m block from the receive buffer at pair_off[p] (>= 0) instead of this rank's
   eigenvector array. */
struct MisRemote
{
    const int *mis_list = nullptr;
    const int64_t *pair_off = nullptr;
    const double *recv = nullptr;
};

__global__ void k_mis_svd(LevelTables L, MisWork W, const int *ae_m, const int64_t *evect_off,
                          const double *evects, int avoid_ess, int nmis, int *borderline,
                          MisRemote RM)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nmis)
        return;
    const int mis = RM.mis_list ? RM.mis_list[w] : w;
    const int db = L.mis2d_I[mis];
    const int s = L.mis2d_I[mis + 1] - db;
    const int *mdofs = L.mis2d_J + db;
    double *X = W.X + W.xoff[w];
    double *sig = W.sig + W.coff[w];

    // MIS entirely on the essential boundary -> no coarse dofs (amg/src/contrib.cpp:578-605)
    if (avoid_ess)
    {
        int interior = 0;
        for (int r = lane; r < s; r += 32)
            if (!(L.agg_flags[mdofs[r]] & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG))
                interior = 1;
        if (!__any_sync(0xffffffffu, interior))
        {
            if (lane == 0)
            {
                W.ncols[mis] = 0;
                W.ncd[mis] = 0;
            }
            return;
        }
    }
    if (s == 1)
    {
        // amg/src/contrib.cpp:607-612
        if (lane == 0)
        {
            W.ncols[mis] = 1;
            W.ncd[mis] = 1;
            X[0] = 1.0;
            sig[0] = 1.0;
        }
        return;
    }
    // gather + filter + normalise, column by column
    int c = 0;
    for (int p = L.mis2AE_I[mis]; p < L.mis2AE_I[mis + 1]; ++p)
    {
        const int ae = L.mis2AE_J[p];
        const int n = L.AE2d_I[ae + 1] - L.AE2d_I[ae];
        const int m = ae_m[ae];
        const int64_t roff = RM.pair_off ? RM.pair_off[p] : -1;
        const double *E = (roff >= 0) ? RM.recv + roff : evects + evect_off[ae];
        for (int v = 0; v < m; ++v)
        {
            double *xc = X + (int64_t)s * c;
            double nrm2 = 0.;
            int nz = 0;
            for (int r = lane; r < s; r += 32)
            {
                const int dof = mdofs[r];
                double a = 0.;
                if (!(avoid_ess && (L.agg_flags[dof] & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG)))
                    a = (roff >= 0) ? E[r + (int64_t)s * v]
                                    : E[sa_dev_map_id_glob_to_AE(L, dof, ae) + (int64_t)n * v];
                xc[r] = a;
                nrm2 += a * a;
                nz |= (a != 0.);
            }
            nrm2 = wsum(nrm2);
            nz = __any_sync(0xffffffffu, nz);
            const double norm = sqrt(nrm2);
            // all-zero columns are dropped by the filter, (near) zero norms by the SVD wrapper
            if (!nz || norm <= 0. + 1e-10)
                continue;
            const double inv = 1. / norm;
            for (int r = lane; r < s; r += 32)
                xc[r] = xc[r] / norm;
            (void)inv;
            ++c;
        }
    }
    __syncwarp();
    if (c == 0)
    {
        if (lane == 0)
        {
            W.ncols[mis] = 0;
            W.ncd[mis] = 0;
        }
        return;
    }
    // one-sided Jacobi (Hestenes): rotate column pairs until mutually orthogonal
    const double tol = 4. * DBL_EPSILON;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        int rotated = 0;
        for (int p = 0; p < c - 1; ++p)
            for (int q = p + 1; q < c; ++q)
            {
                double *xp = X + (int64_t)s * p;
                double *xq = X + (int64_t)s * q;
                double a = 0., b = 0., g = 0.;
                for (int r = lane; r < s; r += 32)
                {
                    const double up = xp[r], uq = xq[r];
                    a += up * up;
                    b += uq * uq;
                    g += up * uq;
                }
                a = wsum(a);
                b = wsum(b);
                g = wsum(g);
                if (g == 0. || fabs(g) <= tol * sqrt(a * b))
                    continue;
                rotated = 1;
                const double zeta = (b - a) / (2. * g);
                const double t = copysign(1., zeta) / (fabs(zeta) + sqrt(1. + zeta * zeta));
                const double cs = 1. / sqrt(1. + t * t);
                const double sn = cs * t;
                for (int r = lane; r < s; r += 32)
                {
                    const double up = xp[r], uq = xq[r];
                    xp[r] = cs * up - sn * uq;
                    xq[r] = sn * up + cs * uq;
                }
                __syncwarp();
            }
        if (!rotated)
            break;
    }
    // singular values = column norms
    double smax = 0.;
    for (int q = 0; q < c; ++q)
    {
        const double *xq = X + (int64_t)s * q;
        double a = 0.;
        for (int r = lane; r < s; r += 32)
            a += xq[r] * xq[r];
        a = sqrt(wsum(a));
        if (lane == 0)
            sig[q] = a;
        smax = fmax(smax, a);
    }
    // xpack_orth_set: keep sigma_i > eps * sigma_0; at most min(s, c) singular values exist
    const double cut = 1.e-10 * smax;
    int k = 0, near = 0;
    for (int q = 0; q < c; ++q)
    {
        const double *xq = X + (int64_t)s * q;
        double a = 0.;
        for (int r = lane; r < s; r += 32)
            a += xq[r] * xq[r];
        a = sqrt(wsum(a));
        if (a > cut)
            ++k;
        // guard band: a singular value within a factor 10 of the rank cut decides the number of
        // coarse dofs on round-off (reported through sa_gpu_get_borderline, not altered)
        if (a > 0.1 * cut && a <= 10. * cut)
            near = 1;
    }
    if (near && lane == 0)
        atomicAdd(borderline, 1);
    k = min(k, min(s, c));
    if (lane == 0)
    {
        W.ncols[mis] = c;
        W.ncd[mis] = k;
    }
}

// one warp per MIS: place the k columns with the largest singular values, in
// descending order, normalised, into the compact mis_tent array
__global__ void k_mis_finalize(LevelTables L, MisWork W, const int64_t *mis_off, double *mis_tent,
                               int nmis, const int *mis_list)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nmis)
        return;
    const int mis = mis_list ? mis_list[w] : w;
    const int s = L.mis2d_I[mis + 1] - L.mis2d_I[mis];
    const int c = W.ncols[mis];
    const int k = W.ncd[mis];
    if (k == 0)
        return;
    const double *X = W.X + W.xoff[w];
    const double *sig = W.sig + W.coff[w];
    double *U = mis_tent + mis_off[mis];
    for (int q = 0; q < c; ++q)
    {
        const double sq = sig[q];
        int rank = 0;
        for (int o = lane; o < c; o += 32)
        {
            const double so = sig[o];
            rank += (so > sq) || (so == sq && o < q);
        }
        for (int o = 16; o > 0; o >>= 1)
            rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (rank >= k)
            continue;
        const double inv = 1. / sq;
        for (int r = lane; r < s; r += 32)
            U[r + (int64_t)s * rank] = X[r + (int64_t)s * q] * inv;
    }
}

__global__ void k_ptent_count(LevelTables L, const int *ncd, int avoid_ess, int *rowcnt)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= L.ND)
        return;
    const bool ess = avoid_ess && (L.agg_flags[d] & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG);
    rowcnt[d] = ess ? 0 : ncd[L.mises[d]];
}

__global__ void k_ptent_fill(LevelTables L, const int *ncd, const int *cdoff,
                             const int64_t *mis_off, const double *mis_tent, const int *PI, int *PJ,
                             double *PA)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= L.ND)
        return;
    const int b = PI[d], cnt = PI[d + 1] - b;
    if (!cnt)
        return;
    const int mis = L.mises[d];
    const int db = L.mis2d_I[mis];
    const int s = L.mis2d_I[mis + 1] - db;
    // position of d inside the (ascending) MIS row
    int lo = 0, hi = s - 1;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (L.mis2d_J[db + mid] < d)
            lo = mid + 1;
        else
            hi = mid;
    }
    const double *U = mis_tent + mis_off[mis];
    for (int c = 0; c < cnt; ++c)
    {
        PJ[b + c] = cdoff[mis] + c;
        PA[b + c] = U[lo + (int64_t)s * c];
    }
}

/* one warp per outgoing (MIS, AE) pair of a sharded tentative-P stage: the rows of the AE's
   eigenvectors at the MIS's dofs (agg_restrict_to_agg_enforce, amg/src/aggregates.cpp:1143-1179),
   s x m column-major, raw values -- the owner filters and normalises */
__global__ void k_mis_pack(LevelTables L, const int *pair_mis, const int *pair_ae, const int64_t *pair_off,
                           int npairs, const int *ae_m, const int64_t *evect_off, const double *evects,
                           double *out)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= npairs)
        return;
    const int mis = pair_mis[w], ae = pair_ae[w];
    const int db = L.mis2d_I[mis];
    const int s = L.mis2d_I[mis + 1] - db;
    const int n = L.AE2d_I[ae + 1] - L.AE2d_I[ae];
    const int m = ae_m[ae];
    const double *E = evects + evect_off[ae];
    double *o = out + pair_off[w];
    for (int r = lane; r < s; r += 32)
    {
        const int lid = sa_dev_map_id_glob_to_AE(L, L.mis2d_J[db + r], ae);
        for (int v = 0; v < m; ++v)
            o[r + (int64_t)s * v] = E[lid + (int64_t)n * v];
    }
}

/* coarse element matrices ------------------------------------------------- */

struct CoarseElmatArgs
{
    const int *ce2d_I, *ce2d_J; // coarse elem_to_dof (rows = finer AEs)
    const int *cdof_mis;        // coarse dof -> finer MIS
    const int *cdoff;           // finer MIS -> first coarse dof
    const int64_t *mis_off;
    const double *mis_tent;
    const int64_t *out_off;
    double *out;
    double *scratch;            // per block: tile n*n + W n*nc
    int64_t scratch_stride;
    int max_mis;                // largest MIS size (shared buffers)
};

__global__ void k_coarse_elmat(LevelTables L, CoarseElmatArgs C, int e_begin, int e_end,
                               int preassembled)
{
    extern __shared__ double sm[];
    double *ubuf = sm;                              // max_mis
    int *lidbuf = (int *)(sm + C.max_mis);          // max_mis
    // gridDim.y blocks share one AE: block (x, y) owns the coarse columns lc == y (mod gridDim.y)
    // in both products (W = A P_e and out = P_e^T W are independent column by column)
    const int split = blockIdx.y, nsplit = gridDim.y;
    for (int e = e_begin + blockIdx.x; e < e_end; e += gridDim.x)
    {
        const int n = L.AE2d_I[e + 1] - L.AE2d_I[e];
        const int cb = C.ce2d_I[e];
        const int nc = C.ce2d_I[e + 1] - cb;
        if (nc == 0)
            continue;
        double *T = C.scratch + (int64_t)blockIdx.x * C.scratch_stride;
        double *Wm = T + (int64_t)n * n;
        if (!preassembled) // (large AEs: k_assemble_large has filled the tile of this block)
            sa_dev_assemble_AE(L, e, T, n);
        // W = A_AE * P_e, column by column
        for (int lc = split; lc < nc; lc += nsplit)
        {
            const int cd = C.ce2d_J[cb + lc];
            const int mis = C.cdof_mis[cd];
            const int idx = cd - C.cdoff[mis];
            const int db = L.mis2d_I[mis];
            const int s = L.mis2d_I[mis + 1] - db;
            const double *U = C.mis_tent + C.mis_off[mis] + (int64_t)s * idx;
            for (int r = threadIdx.x; r < s; r += blockDim.x)
            {
                lidbuf[r] = sa_dev_map_id_glob_to_AE(L, L.mis2d_J[db + r], e);
                ubuf[r] = U[r];
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x)
            {
                double acc = 0.;
                for (int r = 0; r < s; ++r)
                    acc += T[i + (int64_t)n * lidbuf[r]] * ubuf[r];
                Wm[i + (int64_t)n * lc] = acc;
            }
            __syncthreads();
        }
        // out = P_e^T W
        double *out = C.out + C.out_off[e];
        for (int lc1 = 0; lc1 < nc; ++lc1)
        {
            const int cd = C.ce2d_J[cb + lc1];
            const int mis = C.cdof_mis[cd];
            const int idx = cd - C.cdoff[mis];
            const int db = L.mis2d_I[mis];
            const int s = L.mis2d_I[mis + 1] - db;
            const double *U = C.mis_tent + C.mis_off[mis] + (int64_t)s * idx;
            for (int r = threadIdx.x; r < s; r += blockDim.x)
            {
                lidbuf[r] = sa_dev_map_id_glob_to_AE(L, L.mis2d_J[db + r], e);
                ubuf[r] = U[r];
            }
            __syncthreads();
            for (int lc2 = split + nsplit * (int)threadIdx.x; lc2 < nc; lc2 += nsplit * (int)blockDim.x)
            {
                double acc = 0.;
                for (int r = 0; r < s; ++r)
                    acc += ubuf[r] * Wm[lidbuf[r] + (int64_t)n * lc2];
                out[lc1 + (int64_t)nc * lc2] = acc;
            }
            __syncthreads();
        }
    }
}

} // namespace


// contrib_tent_insert_simple + contrib_tent_finalize (amg/src/contrib.cpp:170-194, 73-95): CSR
// tentative P from the per-MIS blocks (every rank of a sharded setup emits all rows)
static void sa_build_ptent_csr(sa_gpu_level *lev)
{
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    LevelTables L = lev->tables();
    const int avoid_ess_bdr_dofs = lev->avoid_ess;
    DevCsr &P = lev->Ptent;
    P.rows = lev->ND;
    P.cols = lev->NDc;
    DevBuf<int> rowcnt;
    rowcnt.alloc(lev->ND);
    P.I.alloc((size_t)lev->ND + 1);
    const int tb = 256;
    SA_LAUNCH(ctx, k_ptent_count, (lev->ND + tb - 1) / tb, tb, 0, L, lev->mis_ncd.p,
              avoid_ess_bdr_dofs, rowcnt.p);
    dev_exclusive_scan_i32(ctx, rowcnt.p, P.I.p, lev->ND);
    int nnz = 0;
    SA_CUDA(cudaMemcpyAsync(&nnz, P.I.p + lev->ND, sizeof(int), cudaMemcpyDeviceToHost, st));
    SA_CUDA(cudaStreamSynchronize(st));
    P.nnz = nnz;
    P.J.alloc(nnz);
    P.A.alloc(nnz);
    SA_LAUNCH(ctx, k_ptent_fill, (lev->ND + tb - 1) / tb, tb, 0, L, lev->mis_ncd.p,
              lev->mis_cd_off.p, lev->mis_off.p, lev->mis_tent.p, P.I.p, P.J.p, P.A.p);
    SA_CUDA(cudaStreamSynchronize(st));
}

extern "C" int sa_gpu_tentative_P(sa_gpu_level *lev, int avoid_ess_bdr_dofs,
                                  int *mis_numcoarsedof, int *NDc_out)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    if (!lev->have_spectral)
        SA_FAIL("sa_gpu_tentative_P: run sa_gpu_local_spectral first");
    const int nmis = lev->num_mises;
    LevelTables L = lev->tables();
    lev->avoid_ess = avoid_ess_bdr_dofs;

    // work-matrix offsets: s x c per MIS, c = sum of m over the AEs containing it
    std::vector<int64_t> xoff(nmis + 1, 0);
    std::vector<int> coff(nmis + 1, 0);
    for (int mis = 0; mis < nmis; ++mis)
    {
        const int s = lev->h_mis2d_I[mis + 1] - lev->h_mis2d_I[mis];
        int c = 0;
        for (int p = lev->h_mis2AE_I[mis]; p < lev->h_mis2AE_I[mis + 1]; ++p)
            c += lev->h_ae_m[lev->h_mis2AE_J[p]];
        c = std::max(c, 1);
        xoff[mis + 1] = xoff[mis] + (int64_t)s * c;
        coff[mis + 1] = coff[mis] + c;
    }
    DevBuf<int64_t> d_xoff;
    DevBuf<int> d_coff, d_ncols;
    DevBuf<double> d_X, d_sig;
    d_xoff.upload(xoff.data(), nmis + 1, st);
    d_coff.upload(coff.data(), nmis + 1, st);
    d_X.alloc(xoff[nmis]);
    d_sig.alloc(coff[nmis]);
    d_ncols.alloc(nmis);
    lev->mis_ncd.alloc(nmis);
    MisWork W;
    W.xoff = d_xoff.p;
    W.coff = d_coff.p;
    W.X = d_X.p;
    W.sig = d_sig.p;
    W.ncols = d_ncols.p;
    W.ncd = lev->mis_ncd.p;
    const int wpb = 4;
    lev->borderline.ensure(2);
    SA_CUDA(cudaMemsetAsync(lev->borderline.p + 1, 0, sizeof(int), st));
    SA_LAUNCH(ctx, k_mis_svd, (nmis + wpb - 1) / wpb, wpb * 32, 0, L, W, lev->ae_m.p,
              lev->evect_off.p, lev->evects.p, avoid_ess_bdr_dofs, nmis, lev->borderline.p + 1,
              MisRemote());
    lev->h_mis_ncd.resize(nmis);
    lev->mis_ncd.download(lev->h_mis_ncd.data(), nmis, st);
    SA_CUDA(cudaStreamSynchronize(st));

    lev->h_mis_off.assign(nmis + 1, 0);
    std::vector<int> cdoff(nmis + 1, 0);
    for (int mis = 0; mis < nmis; ++mis)
    {
        const int s = lev->h_mis2d_I[mis + 1] - lev->h_mis2d_I[mis];
        lev->h_mis_off[mis + 1] = lev->h_mis_off[mis] + (int64_t)s * lev->h_mis_ncd[mis];
        cdoff[mis + 1] = cdoff[mis] + lev->h_mis_ncd[mis];
    }
    lev->NDc = cdoff[nmis];
    lev->mis_off.upload(lev->h_mis_off.data(), nmis + 1, st);
    lev->mis_cd_off.upload(cdoff.data(), nmis + 1, st);
    lev->mis_tent.alloc(lev->h_mis_off[nmis]);
    SA_LAUNCH(ctx, k_mis_finalize, (nmis + wpb - 1) / wpb, wpb * 32, 0, L, W, lev->mis_off.p,
              lev->mis_tent.p, nmis, (const int *)nullptr);
    sa_build_ptent_csr(lev);
    lev->have_tent = true;
    lev->have_P = false;
    lev->have_Ac = false;
    if (mis_numcoarsedof)
        std::copy(lev->h_mis_ncd.begin(), lev->h_mis_ncd.end(), mis_numcoarsedof);
    if (NDc_out)
        *NDc_out = lev->NDc;
    SA_API_END
}


/* The exchange plan of the sharded tentative prolongator for rank `me` (host; identical inputs on
   every rank): owner of a MIS = rank of the lowest-numbered AE containing it
   (amg/src/aggregates.cpp:583-593); a (MIS, AE) pair whose AE lives on rank src != owner moves
   s x m_AE doubles from src to the owner (pairs of a MIS of one dof, or of an AE without vectors,
   move nothing).  owned: the MISes of `me` ascending; pair_off[p] >= 0: offset of pair p inside the
   piece received from its source rank; send_mis / send_ae[q]: the pairs sent to rank q, in the
   order the receiver expects them (ascending pair index). */
static void mis_exchange_plan(int nmis, const int *MI, const int *MJ, const int *DI, int nparts,
                              const int *m_full, int nr, int me, const int *ae_part,
                              std::vector<int> &owned, std::vector<int64_t> &pair_off,
                              std::vector<std::vector<int>> &send_mis,
                              std::vector<std::vector<int>> &send_ae, std::vector<int64_t> &send_cnt,
                              std::vector<int64_t> &recv_cnt)
{
    for (int mis = 0; mis < nmis; ++mis)
    {
        int lowest = nparts;
        for (int p = MI[mis]; p < MI[mis + 1]; ++p)
            lowest = std::min(lowest, MJ[p]);
        const int owner = (lowest < nparts) ? sa_rank_of(ae_part, nr, lowest) : 0;
        if (owner == me)
            owned.push_back(mis);
        const int s = DI[mis + 1] - DI[mis];
        if (s == 1)
            continue; // the owner sets the single entry to 1 without looking at the vectors
        for (int p = MI[mis]; p < MI[mis + 1]; ++p)
        {
            const int ae = MJ[p];
            const int src = sa_rank_of(ae_part, nr, ae);
            if (src == owner || m_full[ae] == 0)
                continue;
            const int64_t sz = (int64_t)s * m_full[ae];
            if (src == me)
            {
                send_mis[owner].push_back(mis);
                send_ae[owner].push_back(ae);
                send_cnt[owner] += sz;
            }
            else if (owner == me)
            {
                pair_off[p] = recv_cnt[src]; // relative to the source's piece, fixed up by the caller
                recv_cnt[src] += sz;
            }
        }
    }
}

/* C ABI, host only (no GPU needed: tests/test_dist_plan.py): the plan above for one rank --
   owner[mis] for every MIS, doubles sent to / received from every rank. */
extern "C" int sa_gpu_mis_exchange_plan(int nmis, const int *mis_to_AE_I, const int *mis_to_AE_J,
                                        const int *mis_to_dof_I, int nparts, const int *ae_m, int nranks,
                                        int rank, const int *ae_part, int *owner, int64_t *send_doubles,
                                        int64_t *recv_doubles)
{
    SA_API_BEGIN
    if (ae_part[0] != 0 || ae_part[nranks] != nparts)
        SA_FAIL("sa_gpu_mis_exchange_plan: ae_part must cover [0, nparts)");
    std::vector<int> owned;
    std::vector<int64_t> pair_off((size_t)std::max(1, mis_to_AE_I[nmis]), -1);
    std::vector<std::vector<int>> send_mis(nranks), send_ae(nranks);
    std::vector<int64_t> sc(nranks, 0), rc(nranks, 0);
    mis_exchange_plan(nmis, mis_to_AE_I, mis_to_AE_J, mis_to_dof_I, nparts, ae_m, nranks, rank, ae_part,
                      owned, pair_off, send_mis, send_ae, sc, rc);
    for (int mis = 0; mis < nmis; ++mis)
    {
        int lowest = nparts;
        for (int p = mis_to_AE_I[mis]; p < mis_to_AE_I[mis + 1]; ++p)
            lowest = std::min(lowest, mis_to_AE_J[p]);
        owner[mis] = (lowest < nparts) ? sa_rank_of(ae_part, nranks, lowest) : 0;
    }
    for (int q = 0; q < nranks; ++q)
    {
        send_doubles[q] = sc[q];
        recv_doubles[q] = rc[q];
    }
    SA_API_END
}

/* Sharded tentative prolongator (SURVEY.md section 8e row "Tentative P a8-a9"; mirror of the
   reduce-to-owner exchange of ContribTent::CommunicateEigenvectors, amg/src/contrib.cpp:492-549,
   and of SharedEntityCommunication's "lowest rank owns", amg/src/aggregates.cpp:583-593):
     1. all-reduce of the per-AE vector counts (every rank ran the eigen stage on its own AEs),
     2. every (MIS, AE) pair whose AE is here and whose MIS owner -- the rank of the lowest-numbered
        AE containing the MIS -- is elsewhere is restricted to the MIS and packed; ONE grouped
        ncclSend / ncclRecv all-to-all-v moves the blocks to the owners,
     3. the owner runs the batched SVD on its MISes,
     4. all-reduce of mis_numcoarsedof (replaces the MPI_Scan of amg/src/contrib.cpp:684),
     5. the owners' blocks go back to everyone (mirror of sec.Broadcast,
        amg/src/aggregates.cpp:1618) as a zero-padded in-switch all-reduce of mis_tent,
     6. every rank emits the rows of the tentative P.
   The eigenvectors themselves never travel. */
extern "C" int sa_gpu_dist_tentative_P(sa_gpu_level *lev, sa_gpu_comm *C, const int *ae_part,
                                       int avoid_ess_bdr_dofs, int *mis_numcoarsedof, int *NDc_out,
                                       double *stats4)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    if (!lev->have_spectral)
        SA_FAIL("sa_gpu_dist_tentative_P: run sa_gpu_local_spectral first");
    if (!C || C->nranks < 2 || !C->comm)
        SA_FAIL("sa_gpu_dist_tentative_P: needs a communicator of at least two ranks");
    const NcclApi &N = sa_nccl();
    const int nr = C->nranks, me = C->rank;
    const int nmis = lev->num_mises, nparts = lev->nparts;
    if (ae_part[0] != 0 || ae_part[nr] != nparts)
        SA_FAIL("sa_gpu_dist_tentative_P: ae_part must cover [0, nparts)");
    LevelTables L = lev->tables();
    lev->avoid_ess = avoid_ess_bdr_dofs;
    lev->h_ae_part.assign(ae_part, ae_part + nr + 1);

    // 1. counts of every AE (this rank holds its own range; zeros elsewhere)
    for (int i = 0; i < nparts; ++i)
        if ((i < ae_part[me] || i >= ae_part[me + 1]) && lev->h_ae_m[i] != 0)
            SA_FAIL("sa_gpu_dist_tentative_P: AE %d outside this rank's range has vectors", i);
    DevBuf<int> d_m_full;
    d_m_full.alloc(nparts);
    SA_NCCL(N.AllReduce(lev->ae_m.p, d_m_full.p, nparts, ncclInt32, ncclSum, C->comm, st));
    std::vector<int> m_full(nparts);
    d_m_full.download(m_full.data(), nparts, st);
    SA_CUDA(cudaStreamSynchronize(st));

    // 2. ownership, pair lists (computed identically on every rank from the replicated tables)
    const std::vector<int> &MI = lev->h_mis2AE_I, &MJ = lev->h_mis2AE_J, &DI = lev->h_mis2d_I;
    std::vector<int> owned; // MISes of this rank, ascending
    std::vector<int64_t> pair_off((size_t)std::max(1, MI[nmis]), -1);
    std::vector<std::vector<int>> send_mis(nr), send_ae(nr);
    std::vector<int64_t> send_cnt(nr, 0), recv_cnt(nr, 0);
    mis_exchange_plan(nmis, MI.data(), MJ.data(), DI.data(), nparts, m_full.data(), nr, me, ae_part, owned,
                      pair_off, send_mis, send_ae, send_cnt, recv_cnt);
    std::vector<int64_t> send_base(nr + 1, 0), recv_base(nr + 1, 0);
    for (int q = 0; q < nr; ++q)
    {
        send_base[q + 1] = send_base[q] + send_cnt[q];
        recv_base[q + 1] = recv_base[q] + recv_cnt[q];
    }
    for (int mis : owned)
        for (int p = MI[mis]; p < MI[mis + 1]; ++p)
            if (pair_off[p] >= 0)
                pair_off[p] += recv_base[sa_rank_of(ae_part, nr, MJ[p])];
    std::vector<int> h_pm, h_pa;
    std::vector<int64_t> h_po;
    for (int q = 0; q < nr; ++q)
    {
        int64_t o = send_base[q];
        for (size_t t = 0; t < send_mis[q].size(); ++t)
        {
            h_pm.push_back(send_mis[q][t]);
            h_pa.push_back(send_ae[q][t]);
            h_po.push_back(o);
            o += (int64_t)(DI[send_mis[q][t] + 1] - DI[send_mis[q][t]]) * m_full[send_ae[q][t]];
        }
    }
    const int nsp = (int)h_pm.size();
    DevBuf<int> d_pm, d_pa, d_owned;
    DevBuf<int64_t> d_po, d_pair_off;
    DevBuf<double> sendbuf, recvbuf;
    sendbuf.alloc((size_t)std::max<int64_t>(1, send_base[nr]));
    recvbuf.alloc((size_t)std::max<int64_t>(1, recv_base[nr]));
    d_pair_off.upload(pair_off.data(), pair_off.size(), st);
    d_owned.upload(owned.data(), std::max<size_t>(1, owned.size()), st);
    const int wpb = 4;
    if (nsp)
    {
        d_pm.upload(h_pm.data(), nsp, st);
        d_pa.upload(h_pa.data(), nsp, st);
        d_po.upload(h_po.data(), nsp, st);
        SA_LAUNCH(ctx, k_mis_pack, (nsp + wpb - 1) / wpb, wpb * 32, 0, L, d_pm.p, d_pa.p, d_po.p, nsp,
                  lev->ae_m.p, lev->evect_off.p, lev->evects.p, sendbuf.p);
    }
    SA_NCCL(N.GroupStart());
    for (int q = 0; q < nr; ++q)
    {
        if (q == me)
            continue;
        if (send_cnt[q])
            SA_NCCL(N.Send(sendbuf.p + send_base[q], (size_t)send_cnt[q], ncclDouble, q, C->comm, st));
        if (recv_cnt[q])
            SA_NCCL(N.Recv(recvbuf.p + recv_base[q], (size_t)recv_cnt[q], ncclDouble, q, C->comm, st));
    }
    SA_NCCL(N.GroupEnd());

    // 3. batched SVD of the owned MISes
    const int nown = (int)owned.size();
    std::vector<int64_t> xoff((size_t)nown + 1, 0);
    std::vector<int> coff((size_t)nown + 1, 0);
    for (int w = 0; w < nown; ++w)
    {
        const int mis = owned[w];
        const int s = DI[mis + 1] - DI[mis];
        int c = 0;
        for (int p = MI[mis]; p < MI[mis + 1]; ++p)
            c += m_full[MJ[p]];
        c = std::max(c, 1);
        xoff[w + 1] = xoff[w] + (int64_t)s * c;
        coff[w + 1] = coff[w] + c;
    }
    DevBuf<int64_t> d_xoff;
    DevBuf<int> d_coff, d_ncols, d_ncd_own;
    DevBuf<double> d_X, d_sig;
    d_xoff.upload(xoff.data(), (size_t)nown + 1, st);
    d_coff.upload(coff.data(), (size_t)nown + 1, st);
    d_X.alloc((size_t)std::max<int64_t>(1, xoff[nown]));
    d_sig.alloc((size_t)std::max(1, coff[nown]));
    d_ncols.alloc(nmis);
    d_ncd_own.alloc(nmis);
    d_ncd_own.zero(st);
    lev->mis_ncd.alloc(nmis);
    MisWork W;
    W.xoff = d_xoff.p;
    W.coff = d_coff.p;
    W.X = d_X.p;
    W.sig = d_sig.p;
    W.ncols = d_ncols.p;
    W.ncd = d_ncd_own.p;
    MisRemote RM;
    RM.mis_list = d_owned.p;
    RM.pair_off = d_pair_off.p;
    RM.recv = recvbuf.p;
    lev->borderline.ensure(2);
    SA_CUDA(cudaMemsetAsync(lev->borderline.p + 1, 0, sizeof(int), st));
    if (nown)
        SA_LAUNCH(ctx, k_mis_svd, (nown + wpb - 1) / wpb, wpb * 32, 0, L, W, d_m_full.p,
                  lev->evect_off.p, lev->evects.p, avoid_ess_bdr_dofs, nown, lev->borderline.p + 1, RM);
    // 4. numcoarsedof of every MIS
    SA_NCCL(N.AllReduce(d_ncd_own.p, lev->mis_ncd.p, nmis, ncclInt32, ncclSum, C->comm, st));
    lev->h_mis_ncd.resize(nmis);
    lev->mis_ncd.download(lev->h_mis_ncd.data(), nmis, st);
    SA_CUDA(cudaStreamSynchronize(st));
    lev->h_mis_off.assign(nmis + 1, 0);
    std::vector<int> cdoff(nmis + 1, 0);
    for (int mis = 0; mis < nmis; ++mis)
    {
        const int s = DI[mis + 1] - DI[mis];
        lev->h_mis_off[mis + 1] = lev->h_mis_off[mis] + (int64_t)s * lev->h_mis_ncd[mis];
        cdoff[mis + 1] = cdoff[mis] + lev->h_mis_ncd[mis];
    }
    lev->NDc = cdoff[nmis];
    lev->mis_off.upload(lev->h_mis_off.data(), nmis + 1, st);
    lev->mis_cd_off.upload(cdoff.data(), nmis + 1, st);
    // 5. blocks of the owned MISes, then everyone's
    const int64_t ntent = lev->h_mis_off[nmis];
    lev->mis_tent.alloc((size_t)std::max<int64_t>(1, ntent));
    lev->mis_tent.zero(st);
    W.ncd = lev->mis_ncd.p;
    if (nown)
        SA_LAUNCH(ctx, k_mis_finalize, (nown + wpb - 1) / wpb, wpb * 32, 0, L, W, lev->mis_off.p,
                  lev->mis_tent.p, nown, d_owned.p);
    if (ntent)
        SA_NCCL(N.AllReduce(lev->mis_tent.p, lev->mis_tent.p, (size_t)ntent, ncclDouble, ncclSum,
                            C->comm, st));
    // 6. rows of P
    sa_build_ptent_csr(lev);
    lev->have_tent = true;
    lev->have_P = false;
    lev->have_Ac = false;
    if (mis_numcoarsedof)
        std::copy(lev->h_mis_ncd.begin(), lev->h_mis_ncd.end(), mis_numcoarsedof);
    if (NDc_out)
        *NDc_out = lev->NDc;
    if (stats4)
    {
        stats4[0] = (double)nown;
        stats4[1] = (double)send_base[nr] * sizeof(double);
        stats4[2] = (double)recv_base[nr] * sizeof(double);
        stats4[3] = (double)ntent * sizeof(double);
    }
    SA_API_END
}

extern "C" int sa_gpu_get_mis_tent(sa_gpu_level *lev, double *mis_tent)
{
    SA_API_BEGIN
    if (!lev->have_tent)
        SA_FAIL("sa_gpu_get_mis_tent: no tentative prolongator");
    lev->mis_tent.download(mis_tent, lev->h_mis_off[lev->num_mises], lev->ctx->stream);
    SA_CUDA(cudaStreamSynchronize(lev->ctx->stream));
    SA_API_END
}

/* coarse element matrices of the finer AEs [ae_a, ae_b) (offsets / storage for all of them) */
static void sa_coarse_elmats_range(sa_gpu_level *finer, sa_gpu_level *coarse, int ae_a, int ae_b)
{
    sa_level_ready(finer);
    sa_gpu_ctx *ctx = finer->ctx;
    cudaStream_t st = ctx->stream;
    if (!finer->have_tent)
        SA_FAIL("sa_gpu_coarse_elmats: finer level has no tentative prolongator");
    if (coarse->NE != finer->nparts)
        SA_FAIL("sa_gpu_coarse_elmats: coarse elements must be the finer AEs");
    const int nparts = finer->nparts;
    // coarse dof -> finer MIS
    std::vector<int> cdof_mis(std::max(1, finer->NDc));
    {
        int cd = 0;
        for (int mis = 0; mis < finer->num_mises; ++mis)
            for (int k = 0; k < finer->h_mis_ncd[mis]; ++k)
                cdof_mis[cd++] = mis;
    }
    DevBuf<int> d_cdof_mis;
    d_cdof_mis.upload(cdof_mis.data(), cdof_mis.size(), st);
    coarse->h_elmat_off.assign((size_t)nparts + 1, 0);
    int64_t max_scratch = 1;
    for (int e = 0; e < nparts; ++e)
    {
        const int64_t nc = coarse->h_e2d_I[e + 1] - coarse->h_e2d_I[e];
        const int64_t n = finer->h_AE2d_I[e + 1] - finer->h_AE2d_I[e];
        coarse->h_elmat_off[e + 1] = coarse->h_elmat_off[e] + nc * nc;
        max_scratch = std::max(max_scratch, n * n + n * nc);
    }
    int max_mis = 1;
    for (int mis = 0; mis < finer->num_mises; ++mis)
        max_mis = std::max(max_mis, finer->h_mis2d_I[mis + 1] - finer->h_mis2d_I[mis]);
    coarse->elmat_off.upload(coarse->h_elmat_off.data(), (size_t)nparts + 1, st);
    coarse->elmat.alloc(coarse->h_elmat_off[nparts]);
    // bounded scratch: persistent blocks
    const size_t scratch_budget = (size_t)1 << 29; // 4 GB of doubles
    int blocks = std::min(std::max(1, ae_b - ae_a), ctx->num_sms * 4);
    blocks = (int)std::max<size_t>(1, std::min<size_t>(blocks, scratch_budget / (size_t)max_scratch));
    // scratch: the finer level's (idle) reflector block is reused when it exists -- a fresh
    // multi-GB request can cost the stream-ordered pool over a second when it is fragmented
    DevBuf<double> scratch_own;
    DevBuf<double> &scratch = ctx->sws.V.p ? ctx->sws.V : scratch_own;
    scratch.ensure((size_t)blocks * max_scratch);
    CoarseElmatArgs C;
    C.ce2d_I = coarse->e2d_I.p;
    C.ce2d_J = coarse->e2d_J.p;
    C.cdof_mis = d_cdof_mis.p;
    C.cdoff = finer->mis_cd_off.p;
    C.mis_off = finer->mis_off.p;
    C.mis_tent = finer->mis_tent.p;
    C.out_off = coarse->elmat_off.p;
    C.out = coarse->elmat.p;
    C.scratch = scratch.p;
    C.scratch_stride = max_scratch;
    C.max_mis = max_mis;
    LevelTables L = finer->tables();
    const size_t smem = (size_t)max_mis * (sizeof(double) + sizeof(int)) + 16;
    if (smem > ctx->smem_optin)
        SA_FAIL("sa_gpu_coarse_elmats: MIS of %d dofs exceeds the shared buffers", max_mis);
    SA_CUDA(cudaFuncSetAttribute(k_coarse_elmat, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ctx->smem_optin));
    int nmax = 1;
    for (int e = 0; e < nparts; ++e)
        nmax = std::max(nmax, finer->h_AE2d_I[e + 1] - finer->h_AE2d_I[e]);
    bool rounds = false;
    if (!L.with_global && nmax > 256)
    {
        // large AEs: every round assembles `blocks` AE matrices with all SMs (by columns),
        // then one block per AE forms P_e^T A_AE P_e
        std::vector<int> iota(nparts);
        for (int e = 0; e < nparts; ++e)
            iota[e] = e;
        DevBuf<int> d_parts;
        d_parts.upload(iota.data(), nparts, st);
        rounds = true;
        for (int e0 = ae_a; e0 < ae_b && rounds; e0 += blocks)
        {
            const int cnt = std::min(blocks, ae_b - e0);
            if (!sa_launch_assemble_large(ctx, L, d_parts.p + e0, nullptr, nullptr, cnt, nmax,
                                          scratch.p, max_scratch, nullptr, st))
            {
                if (e0 != ae_a)
                    SA_FAIL("sa_gpu_coarse_elmats: large assembly became unavailable");
                rounds = false;
                break;
            }
            // few AEs per round: several blocks per AE (columns dealt to them)
            const int nsplit = std::max(1, std::min(8, (3 * ctx->num_sms) / std::max(1, cnt)));
            SA_LAUNCH(ctx, k_coarse_elmat, dim3(cnt, nsplit), 256, smem, L, C, e0, e0 + cnt, 1);
        }
        SA_CUDA(cudaStreamSynchronize(st)); // iota / d_parts go out of scope
    }
    if (!rounds && ae_b > ae_a)
        SA_LAUNCH(ctx, k_coarse_elmat, blocks, 256, smem, L, C, ae_a, ae_b, 0);
    SA_CUDA(cudaStreamSynchronize(st));
    coarse->have_elmat = true;
}

extern "C" int sa_gpu_coarse_elmats(sa_gpu_level *finer, sa_gpu_level *coarse)
{
    SA_API_BEGIN
    sa_coarse_elmats_range(finer, coarse, 0, finer->nparts);
    SA_API_END
}

/* Sharded form ("for a10 the owner broadcasts MIS blocks back"): every rank
   forms P_e^T A_AE P_e for the finer AEs of its own range -- the AEs whose inputs it holds -- and
   one in-place all-gather-v completes the array on every rank. */
extern "C" int sa_gpu_dist_coarse_elmats(sa_gpu_level *finer, sa_gpu_level *coarse, sa_gpu_comm *C,
                                         const int *ae_part)
{
    SA_API_BEGIN
    if (!C || C->nranks < 2 || !C->comm)
        SA_FAIL("sa_gpu_dist_coarse_elmats: needs a communicator of at least two ranks");
    if (!ae_part)
    {
        if ((int)finer->h_ae_part.size() != C->nranks + 1)
            SA_FAIL("sa_gpu_dist_coarse_elmats: no AE ranges (pass ae_part or run "
                    "sa_gpu_dist_tentative_P on the finer level)");
        ae_part = finer->h_ae_part.data();
    }
    if (ae_part[0] != 0 || ae_part[C->nranks] != finer->nparts)
        SA_FAIL("sa_gpu_dist_coarse_elmats: ae_part must cover [0, nparts)");
    sa_coarse_elmats_range(finer, coarse, ae_part[C->rank], ae_part[C->rank + 1]);
    std::vector<int64_t> offs((size_t)C->nranks + 1);
    for (int q = 0; q <= C->nranks; ++q)
        offs[q] = coarse->h_elmat_off[ae_part[q]];
    sa_dev_allgatherv(C, coarse->elmat.p, offs.data(), sizeof(double));
    SA_CUDA(cudaStreamSynchronize(finer->ctx->stream));
    SA_API_END
}

extern "C" int sa_gpu_get_element_matrix(sa_gpu_level *lev, int elno, double *out, int *ne_out)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    if (!lev->have_elmat)
        SA_FAIL("sa_gpu_get_element_matrix: the level has no element matrices");
    if (elno < 0 || elno >= lev->NE)
        SA_FAIL("sa_gpu_get_element_matrix: bad element index %d", elno);
    const int ne = lev->h_e2d_I[elno + 1] - lev->h_e2d_I[elno];
    if (ne_out)
        *ne_out = ne;
    if (out)
    {
        SA_CUDA(cudaMemcpyAsync(out, lev->elmat.p + lev->h_elmat_off[elno], (size_t)ne * ne * sizeof(double),
                                cudaMemcpyDeviceToHost, lev->ctx->stream));
        SA_CUDA(cudaStreamSynchronize(lev->ctx->stream));
    }
    SA_API_END
}

extern "C" int sa_gpu_get_coarse_elmats(sa_gpu_level *coarse, double *celmat)
{
    SA_API_BEGIN
    if (!coarse->have_elmat)
        SA_FAIL("sa_gpu_get_coarse_elmats: no element matrices");
    coarse->elmat.download(celmat, coarse->h_elmat_off[coarse->NE], coarse->ctx->stream);
    SA_CUDA(cudaStreamSynchronize(coarse->ctx->stream));
    SA_API_END
}
