// Sparse kernels of the path:
//   CSR SpMV (HypreParMatrix::Mult), residual, fused polynomial-smoother step
//   (smpr_compute_poly, amg/inc/smpr.hpp:319-339), weighted-l1 -D^-1
//   (mbox_build_Dinv_neg_parallel_matrix, amg/src/mbox.cpp:1839-1861), CSR transpose
//   (interp->Transpose(), amg/inc/tg.hpp:692), SpGEMM (hypre ParMult / RAP call sites
//   amg/src/interp.cpp:77,207 and amg/inc/tg.hpp:700), prefix sums.
#include <algorithm>
#include <climits>

#include "sa_gpu_internal.cuh"

namespace
{

/* ------------------------------------------------------------------ scans */

template <class T> __global__ void k_scan_block_sums(const T *in, T *bsum, int n, int per_block)
{
    __shared__ T sh[32];
    const int b0 = blockIdx.x * per_block;
    T s = 0;
    for (int i = b0 + threadIdx.x; i < min(n, b0 + per_block); i += blockDim.x)
        s += in[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        T t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1)
            t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0)
            bsum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of bsum (nb entries) in place, total at bsum[nb]
template <class T> __global__ void k_scan_small(T *bsum, int nb)
{
    __shared__ T carry;
    __shared__ T sh[1024];
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024)
    {
        const int i = base + threadIdx.x;
        const T v = (i < nb) ? bsum[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1)
        {
            T t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb)
            bsum[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        bsum[nb] = carry;
}

// out[i] = exclusive prefix; each block rescans its segment sequentially by chunks of blockDim
template <class T>
__global__ void k_scan_apply(const T *in, const T *bsum, T *out, int n, int per_block, int nb)
{
    __shared__ T sh[256];
    __shared__ T carry;
    const int b0 = blockIdx.x * per_block;
    if (threadIdx.x == 0)
        carry = bsum[blockIdx.x];
    __syncthreads();
    for (int base = b0; base < min(n, b0 + per_block); base += 256)
    {
        const int i = base + threadIdx.x;
        const T v = (i < n && i < b0 + per_block) ? in[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1)
        {
            T t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n && i < b0 + per_block)
            out[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += sh[255];
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        out[n] = bsum[nb];
}

template <class T> void exclusive_scan(sa_gpu_ctx *ctx, const T *in, T *out, int n)
{
    const int per_block = 4096;
    const int nb = std::max(1, (n + per_block - 1) / per_block);
    DevBuf<T> bsum;
    bsum.alloc((size_t)nb + 1);
    auto k1 = k_scan_block_sums<T>;
    auto k2 = k_scan_small<T>;
    auto k3 = k_scan_apply<T>;
    SA_LAUNCH(ctx, k1, nb, 256, 0, in, bsum.p, n, per_block);
    SA_LAUNCH(ctx, k2, 1, 1024, 0, bsum.p, nb);
    SA_LAUNCH(ctx, k3, nb, 256, 0, in, bsum.p, out, n, per_block, nb);
    SA_CUDA(cudaStreamSynchronize(ctx->stream)); // bsum freed on return
}

/* ------------------------------------------------------------------- SpMV */

// MODE 0: y = A x ; 1: y = b - A x ; 2: y += A x ;
// 3: y = xin + mult * dinv .* (A xin - b)   (x = xin)
// 4: y = mult * dinv .* (-b)                (xin == 0: first smoother step)
template <int TPR, int MODE>
__global__ void k_spmv(int rows, const int *__restrict__ I, const int *__restrict__ J,
                       const double *__restrict__ A, const double *__restrict__ x,
                       const double *__restrict__ b, const double *__restrict__ dinv, double mult,
                       double *__restrict__ y, const double *__restrict__ xrow)
{
    // I, b, dinv, y and xrow are indexed by the LOCAL row (they may point into the middle
    // of the global arrays: row-partitioned use); x is indexed by the global column
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = t / TPR;
    const int sub = t % TPR;
    const bool valid = row < rows;
    double s = 0.;
    // epilogue operands are fetched up front (by the lane that writes the row) so their
    // latency overlaps the matrix stream
    double br = 0., dr = 0., xr = 0.;
    if (valid && sub == 0)
    {
        // read-once operands are streamed (evict-first): only the gathered x should stay in L2.
        // ncu showed 100 MB of extra DRAM reads per fused smoother step when b, dinv and the
        // output competed with x for the L2 of one die.
        if (MODE == 1 || MODE == 3 || MODE == 4)
            br = __ldcs(b + row);
        if (MODE == 3 || MODE == 4)
            dr = __ldcs(dinv + row);
        if (MODE == 3)
            xr = __ldcs(xrow + row);
        if (MODE == 2)
            xr = __ldcs(y + row);
    }
    if (MODE != 4)
    {
        if (valid)
        {
            // matrix entries are streamed (evict-first) so that x stays cached
            const int e = I[row + 1];
            int p = I[row] + sub;
            double s1 = 0.;
            for (; p + TPR < e; p += 2 * TPR)
            {
                const int j0 = __ldcs(J + p), j1 = __ldcs(J + p + TPR);
                const double a0 = __ldcs(A + p), a1 = __ldcs(A + p + TPR);
                s += a0 * x[j0];
                s1 += a1 * x[j1];
            }
            if (p < e)
                s += __ldcs(A + p) * x[__ldcs(J + p)];
            s += s1;
        }
#pragma unroll
        for (int o = TPR >> 1; o > 0; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o, TPR);
    }
    if (valid && sub == 0)
    {
        double out;
        if (MODE == 0)
            out = s;
        else if (MODE == 1)
            out = br - s;
        else if (MODE == 2)
            out = xr + s;
        else if (MODE == 3)
        {
            double tmp = -1. * br;
            tmp += s;
            tmp *= dr;
            out = xr + mult * tmp;
        }
        else
        {
            double tmp = -1. * br;
            tmp *= dr;
            out = mult * tmp;
        }
        __stcs(y + row, out);
    }
}

template <int MODE>
void launch_spmv_raw(sa_gpu_ctx *ctx, int rows, double avg, const int *I, const int *J,
                     const double *Aval, const double *x, const double *xrow, const double *b,
                     const double *dinv, double mult, double *y)
{
    if (rows == 0)
        return;
    const int tb = 256;
#define SA_SPMV_CASE(TPR)                                                              \
    {                                                                                  \
        const long long threads = (long long)rows * TPR;                               \
        auto kern = k_spmv<TPR, MODE>;                                                 \
        SA_LAUNCH(ctx, kern, (unsigned)((threads + tb - 1) / tb), tb, 0, rows, I, J,   \
                  Aval, x, b, dinv, mult, y, xrow);                                    \
    }
    static int tpr_env = getenv("SA_GPU_SPMV_TPR") ? atoi(getenv("SA_GPU_SPMV_TPR")) : 0;
    if (tpr_env == 2)
        SA_SPMV_CASE(2)
    else if (tpr_env == 4)
        SA_SPMV_CASE(4)
    else if (tpr_env == 8)
        SA_SPMV_CASE(8)
    else if (tpr_env == 16)
        SA_SPMV_CASE(16)
    else if (tpr_env == 32)
        SA_SPMV_CASE(32)
    else if (avg <= 4.)
        SA_SPMV_CASE(2)
    else if (avg <= 48.)
        SA_SPMV_CASE(4)
    else if (avg <= 128.)
        SA_SPMV_CASE(8)
    else if (avg <= 384.)
        SA_SPMV_CASE(16)
    else
        SA_SPMV_CASE(32)
#undef SA_SPMV_CASE
}

template <int MODE>
void launch_spmv(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, const double *b,
                 const double *dinv, double mult, double *y)
{
    launch_spmv_raw<MODE>(ctx, A.rows, (double)A.nnz / std::max(1, A.rows), A.I.p, A.J.p, A.A.p,
                          x, x, b, dinv, mult, y);
}

/* device-pointer entry: rows [row0, row0 + nrows) of a CSR matrix whose arrays live on the
   device; I_row0 = I + row0 (absolute offsets), b / dinv / y / xrow already offset */
void dev_spmv_rows_impl(sa_gpu_ctx *ctx, int mode, int nrows, double avg, const int *I_row0,
                        const int *J, const double *A, const double *x, const double *xrow,
                        const double *b, const double *dinv, double mult, double *y)
{
    switch (mode)
    {
    case 0:
        launch_spmv_raw<0>(ctx, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
        break;
    case 1:
        launch_spmv_raw<1>(ctx, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
        break;
    case 2:
        launch_spmv_raw<2>(ctx, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
        break;
    case 3:
        launch_spmv_raw<3>(ctx, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
        break;
    default:
        launch_spmv_raw<4>(ctx, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
        break;
    }
}

/* -------------------------------------------------------------- Dinv_neg */

__global__ void k_diag_isqrt(int rows, const int *I, const int *J, const double *A, double *d1,
                             int *bad)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    double dg = 0.;
    for (int p = I[r]; p < I[r + 1]; ++p)
        if (J[p] == r)
            dg = fabs(A[p]);
    if (!(dg > 0.))
        *bad = 1;
    d1[r] = 1. / sqrt(dg);
}

template <int TPR>
__global__ void k_dinv_neg(int rows, const int *I, const int *J, const double *A, const double *d1,
                           double *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = t / TPR, sub = t % TPR;
    const bool valid = row < rows;
    double s = 0.;
    if (valid)
        for (int p = I[row] + sub; p < I[row + 1]; p += TPR)
            s += fabs(A[p]) * d1[J[p]];
    for (int o = TPR >> 1; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o, TPR);
    if (valid && sub == 0)
        out[row] = -1. / ((1. / d1[row]) * s); // -1 / (sqrt|a_ii| * y_i)
}

/* ------------------------------------------------------------ transpose */

__global__ void k_count_cols(int nnz, const int *J, int *cnt)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x)
        atomicAdd(&cnt[J[p]], 1);
}

__global__ void k_transpose_fill(int rows, const int *I, const int *J, const double *A,
                                 const int *tI, int *fill, int *tJ, double *tA)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    for (int p = I[r]; p < I[r + 1]; ++p)
    {
        const int c = J[p];
        const int q = tI[c] + atomicAdd(&fill[c], 1);
        tJ[q] = r;
        tA[q] = A[p];
    }
}

// warp per row: rank sort (keys within a row are distinct) from (J,A) into (oJ,oA)
__global__ void k_sort_rows(int rows, const int *I, const int *J, const double *A, int *oJ,
                            double *oA)
{
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows)
        return;
    const int b = I[row], e = I[row + 1];
    for (int p = b + lane; p < e; p += 32)
    {
        const int key = J[p];
        int rank = 0;
        for (int q = b; q < e; ++q)
            rank += (J[q] < key);
        oJ[b + rank] = key;
        oA[b + rank] = A[p];
    }
}

/* --------------------------------------------------------------- SpGEMM */

__global__ void k_spgemm_ub(int rows, const int *AI, const int *AJ, const int *BI, int *ub)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    long long s = 0;
    for (int p = AI[r]; p < AI[r + 1]; ++p)
        s += BI[AJ[p] + 1] - BI[AJ[p]];
    ub[r] = (int)min(s, (long long)0x7fffffff);
}

__device__ __forceinline__ unsigned sa_hash(int key) { return (unsigned)key * 2654435761u; }

// One block per listed row.  Hash table of `tsize` (power of two) int keys [+ double
// values] in shared memory (table_global == 0) or in global scratch.
// NUMERIC == 0: count distinct columns -> rowcnt[row]
// NUMERIC == 1: accumulate, compact, sort ascending, write at CI[row]
template <int NUMERIC>
__global__ void k_spgemm_rows(const int *row_list, int nrows, const int *AI, const int *AJ,
                              const double *AA, const int *BI, const int *BJ, const double *BA,
                              int tsize_fixed, const int64_t *tab_off, const int *tab_size,
                              int *gkeys, double *gvals, int *rowcnt, const int *CI, int *CJ,
                              double *CA)
{
    extern __shared__ unsigned char smraw[];
    __shared__ int s_cnt;
    const int li = blockIdx.x;
    if (li >= nrows)
        return;
    const int row = row_list[li];
    int tsize;
    int *keys;
    double *vals;
    if (tab_off)
    {
        tsize = tab_size[li];
        keys = gkeys + tab_off[li];
        vals = gvals + tab_off[li];
    }
    else
    {
        tsize = tsize_fixed;
        vals = (double *)smraw;
        keys = (int *)(smraw + (NUMERIC ? (size_t)tsize * sizeof(double) : 0));
    }
    const unsigned mask = (unsigned)tsize - 1u;
    for (int t = threadIdx.x; t < tsize; t += blockDim.x)
    {
        keys[t] = -1;
        if (NUMERIC)
            vals[t] = 0.;
    }
    if (threadIdx.x == 0)
        s_cnt = 0;
    __syncthreads();
    // lanes of a warp walk one B row together; warps take different entries of the A row
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int local_new = 0;
    for (int p = AI[row] + w; p < AI[row + 1]; p += nw)
    {
        const int k = AJ[p];
        const double a = NUMERIC ? AA[p] : 0.;
        for (int q = BI[k] + lane; q < BI[k + 1]; q += 32)
        {
            const int col = BJ[q];
            unsigned h = sa_hash(col) & mask;
            for (;;)
            {
                const int old = atomicCAS(&keys[h], -1, col);
                if (old == -1)
                {
                    ++local_new;
                    break;
                }
                if (old == col)
                    break;
                h = (h + 1) & mask;
            }
            if (NUMERIC)
                atomicAdd(&vals[h], a * BA[q]);
        }
    }
    if (!NUMERIC)
    {
        for (int o = 16; o > 0; o >>= 1)
            local_new += __shfl_xor_sync(0xffffffffu, local_new, o);
        if (lane == 0 && local_new)
            atomicAdd(&s_cnt, local_new);
        __syncthreads();
        if (threadIdx.x == 0)
            rowcnt[row] = s_cnt;
        return;
    }
    __syncthreads();
    // compact (arbitrary order) into the output segment
    const int cb = CI[row];
    for (int t = threadIdx.x; t < tsize; t += blockDim.x)
        if (keys[t] >= 0)
        {
            const int q = atomicAdd(&s_cnt, 1);
            CJ[cb + q] = keys[t];
            CA[cb + q] = vals[t];
        }
    __syncthreads();
    const int cnt = s_cnt;
    // rank sort through the (now free) table
    for (int t = threadIdx.x; t < cnt; t += blockDim.x)
    {
        const int key = CJ[cb + t];
        int rank = 0;
        for (int q = 0; q < cnt; ++q)
            rank += (CJ[cb + q] < key);
        keys[rank] = key;
        vals[rank] = CA[cb + t];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += blockDim.x)
    {
        CJ[cb + t] = keys[t];
        CA[cb + t] = vals[t];
    }
}

__global__ void k_scale_rows_add_identity(int rows, const int *I, const int *J, double *A,
                                          const double *dinv_neg, double scale)
{
    // iter_matr = I + scale * diag(dinv_neg) * A   (amg/src/interp.cpp:64-82, 200-202)
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    const double dr = dinv_neg[r] * scale;
    for (int p = I[r]; p < I[r + 1]; ++p)
    {
        double v = dr * A[p];
        if (J[p] == r)
            v += 1.;
        A[p] = v;
    }
}

} // namespace

void dev_spmv_rows(sa_gpu_ctx *ctx, int mode, int nrows, double avg, const int *I_row0,
                   const int *J, const double *A, const double *x, const double *xrow,
                   const double *b, const double *dinv, double mult, double *y)
{
    dev_spmv_rows_impl(ctx, mode, nrows, avg, I_row0, J, A, x, xrow, b, dinv, mult, y);
}

void dev_exclusive_scan_i32(sa_gpu_ctx *ctx, const int *in, int *out, int n)
{
    exclusive_scan<int>(ctx, in, out, n);
}
void dev_exclusive_scan_i64(sa_gpu_ctx *ctx, const int64_t *in, int64_t *out, int n)
{
    exclusive_scan<int64_t>(ctx, in, out, n);
}

void dev_spmv(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, double *y)
{
    launch_spmv<0>(ctx, A, x, nullptr, nullptr, 0., y);
}
void dev_residual(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, const double *b, double *y)
{
    launch_spmv<1>(ctx, A, x, b, nullptr, 0., y);
}
void dev_spmv_add(sa_gpu_ctx *ctx, const DevCsr &A, const double *x, double *y)
{
    launch_spmv<2>(ctx, A, x, nullptr, nullptr, 0., y);
}
void dev_smoother_step(sa_gpu_ctx *ctx, const DevCsr &A, const double *dinv_neg, const double *b,
                       const double *xin, double *xout, double mult, int xin_is_zero)
{
    if (xin_is_zero)
        launch_spmv<4>(ctx, A, xin, b, dinv_neg, mult, xout);
    else
        launch_spmv<3>(ctx, A, xin, b, dinv_neg, mult, xout);
}

void dev_csr_transpose(sa_gpu_ctx *ctx, const DevCsr &A, DevCsr &At)
{
    cudaStream_t st = ctx->stream;
    At.rows = A.cols;
    At.cols = A.rows;
    At.nnz = A.nnz;
    At.I.alloc((size_t)A.cols + 1);
    At.J.alloc(A.nnz);
    At.A.alloc(A.nnz);
    DevBuf<int> cnt, fill, tJ;
    DevBuf<double> tA;
    cnt.alloc(A.cols);
    cnt.zero(st);
    fill.alloc(A.cols);
    fill.zero(st);
    tJ.alloc(A.nnz);
    tA.alloc(A.nnz);
    if (A.nnz)
        SA_LAUNCH(ctx, k_count_cols, std::min(4096, (A.nnz + 255) / 256), 256, 0, A.nnz, A.J.p,
                  cnt.p);
    dev_exclusive_scan_i32(ctx, cnt.p, At.I.p, A.cols);
    if (A.rows)
        SA_LAUNCH(ctx, k_transpose_fill, (A.rows + 255) / 256, 256, 0, A.rows, A.I.p, A.J.p, A.A.p,
                  At.I.p, fill.p, tJ.p, tA.p);
    if (At.rows)
        SA_LAUNCH(ctx, k_sort_rows, (At.rows + 7) / 8, 256, 0, At.rows, At.I.p, tJ.p, tA.p, At.J.p,
                  At.A.p);
    SA_CUDA(cudaStreamSynchronize(st));
}

/* total of the row counts in 64 bits (the row pointers are int32: products beyond 2^31 - 1
   entries must fail, not wrap) */
__global__ void k_total_i64(int n, const int *cnt, unsigned long long *total)
{
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        s += (unsigned long long)cnt[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s)
        atomicAdd(total, s);
}

static int pow2ceil(long long v)
{
    long long p = 1;
    while (p < v)
        p <<= 1;
    return (int)std::min<long long>(p, 1ll << 30);
}

void dev_spgemm(sa_gpu_ctx *ctx, const DevCsr &A, const DevCsr &B, DevCsr &C)
{
    cudaStream_t st = ctx->stream;
    if (A.cols != B.rows)
        SA_FAIL("dev_spgemm: dimension mismatch %d vs %d", A.cols, B.rows);
    const int rows = A.rows;
    C.rows = rows;
    C.cols = B.cols;
    C.I.alloc((size_t)rows + 1);
    DevBuf<int> ub, rowcnt;
    ub.alloc(rows);
    rowcnt.alloc(rows);
    rowcnt.zero(st);
    if (rows)
        SA_LAUNCH(ctx, k_spgemm_ub, (rows + 255) / 256, 256, 0, rows, A.I.p, A.J.p, B.I.p, ub.p);
    std::vector<int> h_ub(rows);
    ub.download(h_ub.data(), rows, st);
    SA_CUDA(cudaStreamSynchronize(st));

    // bins: shared-memory tables of 512 / 4096 / 16384 slots, else global tables
    const int nbins = 4;
    const int smem_slots[3] = {512, 4096, 16384};
    std::vector<int> lists[nbins];
    for (int r = 0; r < rows; ++r)
    {
        const long long need = 2ll * std::min<long long>(h_ub[r], B.cols);
        if (h_ub[r] == 0)
            continue; // empty row
        int b = 3;
        for (int k = 0; k < 3; ++k)
            if (need <= smem_slots[k])
            {
                b = k;
                break;
            }
        lists[b].push_back(r);
    }
    // static shared (s_cnt) counts against the opt-in limit
    SA_CUDA(cudaFuncSetAttribute(k_spgemm_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ctx->smem_optin - 1024));
    SA_CUDA(cudaFuncSetAttribute(k_spgemm_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)ctx->smem_optin - 1024));
    DevBuf<int> d_list[nbins];
    for (int b = 0; b < nbins; ++b)
        d_list[b].upload(lists[b].data(), lists[b].size(), st);
    const int threads_of_bin[nbins] = {32, 128, 256, 256};

    // global-table batches for the large bin
    struct Batch
    {
        int begin, end;
    };
    std::vector<Batch> batches;
    std::vector<int64_t> h_taboff(lists[3].size() + 1, 0);
    std::vector<int> h_tabsize(lists[3].size());
    const int64_t tab_budget = (int64_t)1 << 28; // slots per batch (3 GB of keys+values)
    {
        int bb = 0;
        int64_t acc = 0;
        for (size_t i = 0; i < lists[3].size(); ++i)
        {
            const int r = lists[3][i];
            h_tabsize[i] = pow2ceil(2ll * std::min<long long>(h_ub[r], B.cols));
            if (acc + h_tabsize[i] > tab_budget && (int)i > bb)
            {
                batches.push_back({bb, (int)i});
                bb = (int)i;
                acc = 0;
            }
            h_taboff[i] = acc;
            acc += h_tabsize[i];
        }
        if (lists[3].size())
            batches.push_back({bb, (int)lists[3].size()});
    }
    DevBuf<int64_t> d_taboff;
    DevBuf<int> d_tabsize, gkeys;
    DevBuf<double> gvals;
    if (lists[3].size())
    {
        d_taboff.upload(h_taboff.data(), lists[3].size(), st);
        d_tabsize.upload(h_tabsize.data(), lists[3].size(), st);
        int64_t maxslots = 0;
        for (size_t b = 0; b < batches.size(); ++b)
        {
            int64_t acc = 0;
            for (int i = batches[b].begin; i < batches[b].end; ++i)
                acc += h_tabsize[i];
            maxslots = std::max(maxslots, acc);
        }
        gkeys.alloc(maxslots);
        gvals.alloc(maxslots);
    }

    auto kspg0 = k_spgemm_rows<0>;
    auto kspg1 = k_spgemm_rows<1>;
    for (int pass = 0; pass < 2; ++pass)
    {
        if (pass == 1)
        {
            DevBuf<unsigned long long> d_total;
            d_total.alloc(1);
            d_total.zero(st);
            if (rows)
                SA_LAUNCH(ctx, k_total_i64, std::min((rows + 255) / 256, ctx->num_sms * 8), 256, 0, rows,
                          rowcnt.p, d_total.p);
            dev_exclusive_scan_i32(ctx, rowcnt.p, C.I.p, rows);
            int nnz = 0;
            unsigned long long total = 0;
            SA_CUDA(cudaMemcpyAsync(&nnz, C.I.p + rows, sizeof(int), cudaMemcpyDeviceToHost, st));
            d_total.download(&total, 1, st);
            SA_CUDA(cudaStreamSynchronize(st));
            if (total > (unsigned long long)INT_MAX)
                SA_FAIL("dev_spgemm: the product has %llu entries, more than the int32 row pointers hold "
                        "(%d x %d times %d x %d)", total, A.rows, A.cols, B.rows, B.cols);
            C.nnz = nnz;
            C.J.alloc(nnz);
            C.A.alloc(nnz);
        }
        for (int b = 0; b < 3; ++b)
        {
            const int cnt = (int)lists[b].size();
            if (!cnt)
                continue;
            const size_t smem = (size_t)smem_slots[b] * (pass ? sizeof(double) + sizeof(int) : sizeof(int));
            if (pass == 0)
                SA_LAUNCH(ctx, kspg0, cnt, threads_of_bin[b], smem, d_list[b].p, cnt,
                          A.I.p, A.J.p, A.A.p, B.I.p, B.J.p, B.A.p, smem_slots[b],
                          (const int64_t *)nullptr, (const int *)nullptr, (int *)nullptr,
                          (double *)nullptr, rowcnt.p, (const int *)nullptr, (int *)nullptr,
                          (double *)nullptr);
            else
                SA_LAUNCH(ctx, kspg1, cnt, threads_of_bin[b], smem, d_list[b].p, cnt,
                          A.I.p, A.J.p, A.A.p, B.I.p, B.J.p, B.A.p, smem_slots[b],
                          (const int64_t *)nullptr, (const int *)nullptr, (int *)nullptr,
                          (double *)nullptr, rowcnt.p, C.I.p, C.J.p, C.A.p);
        }
        for (size_t bt = 0; bt < batches.size(); ++bt)
        {
            const int cnt = batches[bt].end - batches[bt].begin;
            const int off = batches[bt].begin;
            if (pass == 0)
                SA_LAUNCH(ctx, kspg0, cnt, threads_of_bin[3], 0, d_list[3].p + off, cnt,
                          A.I.p, A.J.p, A.A.p, B.I.p, B.J.p, B.A.p, 0, d_taboff.p + off,
                          d_tabsize.p + off, gkeys.p, gvals.p, rowcnt.p, (const int *)nullptr,
                          (int *)nullptr, (double *)nullptr);
            else
                SA_LAUNCH(ctx, kspg1, cnt, threads_of_bin[3], 0, d_list[3].p + off, cnt,
                          A.I.p, A.J.p, A.A.p, B.I.p, B.J.p, B.A.p, 0, d_taboff.p + off,
                          d_tabsize.p + off, gkeys.p, gvals.p, rowcnt.p, C.I.p, C.J.p, C.A.p);
        }
    }
    SA_CUDA(cudaStreamSynchronize(st));
}

extern "C" int sa_gpu_build_Dinv_neg(sa_gpu_level *lev)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    if (!lev->A)
        SA_FAIL("sa_gpu_build_Dinv_neg: level has no operator");
    const DevCsr &A = *lev->A;
    DevBuf<double> d1;
    DevBuf<int> bad;
    d1.alloc(A.rows);
    bad.alloc(1);
    bad.zero(ctx->stream);
    lev->Dinv_neg.alloc(A.rows);
    SA_LAUNCH(ctx, k_diag_isqrt, (A.rows + 255) / 256, 256, 0, A.rows, A.I.p, A.J.p, A.A.p, d1.p,
              bad.p);
    const long long threads = (long long)A.rows * 8;
    auto kdn = k_dinv_neg<8>;
    SA_LAUNCH(ctx, kdn, (unsigned)((threads + 255) / 256), 256, 0, A.rows, A.I.p, A.J.p, A.A.p,
              d1.p, lev->Dinv_neg.p);
    int h = 0;
    bad.download(&h, 1, ctx->stream);
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h)
        SA_FAIL("sa_gpu_build_Dinv_neg: zero diagonal entry");
    lev->have_Dinv = true;
    SA_API_END
}

extern "C" int sa_gpu_smooth_P(sa_gpu_level *lev, int degree, const double *roots)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    if (!lev->have_tent)
        SA_FAIL("sa_gpu_smooth_P: no tentative prolongator");
    if (degree > 0 && !lev->have_Dinv)
        SA_FAIL("sa_gpu_smooth_P: sa_gpu_build_Dinv_neg has not been called");
    // interp = clone(tent)
    DevCsr &P = lev->P;
    P.rows = lev->Ptent.rows;
    P.cols = lev->Ptent.cols;
    P.nnz = lev->Ptent.nnz;
    P.I.alloc((size_t)P.rows + 1);
    P.J.alloc(P.nnz);
    P.A.alloc(P.nnz);
    SA_CUDA(cudaMemcpyAsync(P.I.p, lev->Ptent.I.p, ((size_t)P.rows + 1) * sizeof(int),
                            cudaMemcpyDeviceToDevice, st));
    SA_CUDA(cudaMemcpyAsync(P.J.p, lev->Ptent.J.p, (size_t)P.nnz * sizeof(int),
                            cudaMemcpyDeviceToDevice, st));
    SA_CUDA(cudaMemcpyAsync(P.A.p, lev->Ptent.A.p, (size_t)P.nnz * sizeof(double),
                            cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < degree; ++k)
    {
        const DevCsr &A = *lev->A;
        DevCsr iter;
        iter.rows = A.rows;
        iter.cols = A.cols;
        iter.nnz = A.nnz;
        iter.I.alloc((size_t)A.rows + 1);
        iter.J.alloc(A.nnz);
        iter.A.alloc(A.nnz);
        SA_CUDA(cudaMemcpyAsync(iter.I.p, A.I.p, ((size_t)A.rows + 1) * sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
        SA_CUDA(cudaMemcpyAsync(iter.J.p, A.J.p, (size_t)A.nnz * sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
        SA_CUDA(cudaMemcpyAsync(iter.A.p, A.A.p, (size_t)A.nnz * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
        SA_LAUNCH(ctx, k_scale_rows_add_identity, (A.rows + 255) / 256, 256, 0, A.rows, iter.I.p,
                  iter.J.p, iter.A.p, lev->Dinv_neg.p, 1. / roots[k]);
        DevCsr newP;
        dev_spgemm(ctx, iter, P, newP);
        P.swap(newP);
    }
    dev_csr_transpose(ctx, P, lev->R);
    SA_CUDA(cudaStreamSynchronize(st));
    lev->have_P = true;
    lev->have_Ac = false;
    SA_API_END
}

extern "C" int sa_gpu_rap(sa_gpu_level *lev)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    if (!lev->have_P)
        SA_FAIL("sa_gpu_rap: no prolongator (call sa_gpu_smooth_P)");
    DevCsr AP;
    dev_spgemm(ctx, *lev->A, lev->P, AP);
    dev_spgemm(ctx, lev->R, AP, lev->Ac);
    lev->have_Ac = true;
    SA_API_END
}
