// Sharded setup stages over the ranks of a communicator, next to the
// sharded tentative prolongator / coarse element matrices of tentative.cu:
//   sa_dev_allgatherv      in-place all-gather-v on device arrays (grouped ncclSend / ncclRecv)
//   sa_gpu_dist_smooth_P   interp_smooth (amg/src/interp.cpp:172-229) and
//   sa_gpu_dist_rap        tg_coarse_matr (amg/inc/tg.hpp:695-709) as ROW-PARTITIONED SpGEMM: the
//                          reference's hypre ParMult / RAP own a row block per MPI rank; here every
//                          rank forms its row block of each product with the single-GPU SpGEMM
//                          kernels and the blocks are all-gathered (row counts, then column indices
//                          and values straight into the full CSR arrays), because the next stage of
//                          the replicated hierarchy reads all rows.
#include <algorithm>

#include "nccl_dl.cuh"

void sa_dev_allgatherv(sa_gpu_comm *C, void *buf, const int64_t *offs, size_t elem_bytes)
{
    const NcclApi &N = sa_nccl();
    cudaStream_t st = C->ctx->stream;
    const int nr = C->nranks, me = C->rank;
    char *b = (char *)buf;
    const size_t mine = (size_t)(offs[me + 1] - offs[me]) * elem_bytes;
    SA_NCCL(N.GroupStart());
    for (int q = 0; q < nr; ++q)
    {
        if (q == me)
            continue;
        if (mine)
            SA_NCCL(N.Send(b + (size_t)offs[me] * elem_bytes, mine, ncclChar, q, C->comm, st));
        const size_t theirs = (size_t)(offs[q + 1] - offs[q]) * elem_bytes;
        if (theirs)
            SA_NCCL(N.Recv(b + (size_t)offs[q] * elem_bytes, theirs, ncclChar, q, C->comm, st));
    }
    SA_NCCL(N.GroupEnd());
}

namespace
{
__global__ void k_row_counts(int r0, int r1, const int *I_local, int *cnt_full)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < r1 - r0)
        cnt_full[r0 + i] = I_local[i + 1] - I_local[i];
}

/* out = A * B with the rows of A dealt to the ranks in equal contiguous blocks; every rank ends
   with the complete product.  bytes_moved (optional) += what this rank received. */
void dist_spgemm(sa_gpu_comm *C, const DevCsr &A, const DevCsr &B, DevCsr &out, double *bytes_moved)
{
    sa_gpu_ctx *ctx = C->ctx;
    cudaStream_t st = ctx->stream;
    const int nr = C->nranks, me = C->rank;
    const int rows = A.rows;
    std::vector<int64_t> rp((size_t)nr + 1);
    for (int q = 0; q <= nr; ++q)
        rp[q] = ((int64_t)rows * q) / nr;
    const int r0 = (int)rp[me], r1 = (int)rp[me + 1];
    // my row block (a view of A's rows: the row pointers keep their absolute offsets)
    DevCsr Av, loc;
    Av.rows = r1 - r0;
    Av.cols = A.cols;
    Av.nnz = A.nnz;
    Av.I.view(A.I.p + r0, (size_t)(r1 - r0) + 1);
    Av.J.view(A.J.p, (size_t)A.nnz);
    Av.A.view(A.A.p, (size_t)A.nnz);
    dev_spgemm(ctx, Av, B, loc);
    // row counts of everyone -> row pointers of the full product
    DevBuf<int> cnt;
    cnt.alloc((size_t)rows);
    if (r1 > r0)
        SA_LAUNCH(ctx, k_row_counts, (r1 - r0 + 255) / 256, 256, 0, r0, r1, loc.I.p, cnt.p);
    sa_dev_allgatherv(C, cnt.p, rp.data(), sizeof(int));
    out.rows = rows;
    out.cols = B.cols;
    out.I.alloc((size_t)rows + 1);
    dev_exclusive_scan_i32(ctx, cnt.p, out.I.p, rows);
    std::vector<int> hI((size_t)rows + 1);
    out.I.download(hI.data(), (size_t)rows + 1, st);
    SA_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < rows; ++r)
        if (hI[r + 1] < hI[r]) // (a wrapped int32 prefix sum is not monotone)
            SA_FAIL("dist_spgemm: the gathered product exceeds the int32 row pointers (%d rows)", rows);
    out.nnz = hI[rows];
    out.J.alloc((size_t)out.nnz);
    out.A.alloc((size_t)out.nnz);
    std::vector<int64_t> np((size_t)nr + 1);
    for (int q = 0; q <= nr; ++q)
        np[q] = hI[rp[q]];
    if (np[me + 1] - np[me] != loc.nnz)
        SA_FAIL("dist_spgemm: row block has %d entries, the gathered row pointers say %lld", loc.nnz,
                (long long)(np[me + 1] - np[me]));
    if (loc.nnz)
    {
        SA_CUDA(cudaMemcpyAsync(out.J.p + np[me], loc.J.p, (size_t)loc.nnz * sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
        SA_CUDA(cudaMemcpyAsync(out.A.p + np[me], loc.A.p, (size_t)loc.nnz * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
    }
    sa_dev_allgatherv(C, out.J.p, np.data(), sizeof(int));
    sa_dev_allgatherv(C, out.A.p, np.data(), sizeof(double));
    SA_CUDA(cudaStreamSynchronize(st));
    if (bytes_moved)
        *bytes_moved += (double)(out.nnz - loc.nnz) * 12. + (double)(rows - (r1 - r0)) * 4.;
}

/* row pointers of a rows-by-* matrix whose rows [c0, c1) are those of a local block (row pointers
   Iloc, starting at 0) and whose other rows are empty */
__global__ void k_place_row_ptr(int rows, int c0, int c1, const int *Iloc, int *I)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > rows)
        return;
    I[r] = (r <= c0) ? 0 : (r >= c1 ? Iloc[c1 - c0] : Iloc[r - c0]);
}

__global__ void k_scale_rows_add_identity_d(int rows, const int *I, const int *J, double *A,
                                            const double *dinv_neg, double mult)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    const double s = dinv_neg[r] * mult;
    for (int p = I[r]; p < I[r + 1]; ++p)
    {
        double v = A[p] * s;
        if (J[p] == r)
            v += 1.;
        A[p] = v;
    }
}
} // namespace

namespace
{
__global__ void k_thr_count(int rows, const int *I, const double *A, double tol, int *cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    int c = 0;
    for (int p = I[r]; p < I[r + 1]; ++p)
        c += fabs(A[p]) > tol;
    cnt[r] = c;
}
__global__ void k_thr_fill(int rows, const int *I, const int *J, const double *A, double tol,
                           const int *In, int *Jn, double *An)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    int o = In[r];
    for (int p = I[r]; p < I[r + 1]; ++p)
        if (fabs(A[p]) > tol)
        {
            Jn[o] = J[p];
            An[o] = A[p];
            ++o;
        }
}
} // namespace

/* AltThreshold (amg/src/interp.cpp:86-170) on the smoothed prolongator: keeps the entries with
   |p_ij| > drop_tol (row order preserved), R = P^T is rebuilt.  interp_smooth applies it when
   drop_tol != 0 (amg/src/interp.cpp:219-228). */
extern "C" int sa_gpu_threshold_P(sa_gpu_level *lev, double drop_tol, int *nnz_before, int *nnz_after)
{
    SA_API_BEGIN
    if (!lev->have_P)
        SA_FAIL("sa_gpu_threshold_P: no prolongator (call sa_gpu_smooth_P)");
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    DevCsr &P = lev->P;
    if (nnz_before)
        *nnz_before = P.nnz;
    DevBuf<int> cnt;
    cnt.alloc((size_t)P.rows);
    DevCsr Pn;
    Pn.rows = P.rows;
    Pn.cols = P.cols;
    Pn.I.alloc((size_t)P.rows + 1);
    if (P.rows)
        SA_LAUNCH(ctx, k_thr_count, (P.rows + 255) / 256, 256, 0, P.rows, P.I.p, P.A.p, drop_tol, cnt.p);
    dev_exclusive_scan_i32(ctx, cnt.p, Pn.I.p, P.rows);
    int nnz = 0;
    SA_CUDA(cudaMemcpyAsync(&nnz, Pn.I.p + P.rows, sizeof(int), cudaMemcpyDeviceToHost, st));
    SA_CUDA(cudaStreamSynchronize(st));
    Pn.nnz = nnz;
    Pn.J.alloc((size_t)nnz);
    Pn.A.alloc((size_t)nnz);
    if (P.rows)
        SA_LAUNCH(ctx, k_thr_fill, (P.rows + 255) / 256, 256, 0, P.rows, P.I.p, P.J.p, P.A.p, drop_tol,
                  Pn.I.p, Pn.J.p, Pn.A.p);
    P.swap(Pn);
    dev_csr_transpose(ctx, P, lev->R);
    SA_CUDA(cudaStreamSynchronize(st));
    lev->have_Ac = false;
    if (nnz_after)
        *nnz_after = nnz;
    SA_API_END
}

/* A level below the last spectral one whose prolongator is GIVEN (CorrectNullspace,
   amg/src/solve.cpp:52-164: the "scaling P" of interp_scaling_P_assemble, amg/src/interp.cpp:842-909
   -- one column per MIS, the coarse representation of the constant vector): operator = the finer
   level's Ac (alias), P from the host CSR, R = P^T.  sa_gpu_build_Dinv_neg and sa_gpu_rap complete
   it; sa_gpu_solver_create takes it as one more level of the cycle (tg_cycle_atb with the same SAS
   smoother, exact solve on its Ac instead of BoomerAMG). */
extern "C" int sa_gpu_level_create_from_P(sa_gpu_ctx *ctx, sa_gpu_level *finer, int cols, const int *P_I,
                                          const int *P_J, const double *P_A, sa_gpu_level **out)
{
    SA_API_BEGIN
    *out = nullptr;
    if (!finer || !finer->have_Ac)
        SA_FAIL("sa_gpu_level_create_from_P: the finer level has no Ac");
    SA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int rows = finer->Ac.rows;
    sa_gpu_level *L = new sa_gpu_level;
    L->ctx = ctx;
    L->finer = finer;
    L->ND = rows;
    L->A = &finer->Ac;
    L->NDc = cols;
    const int nnz = P_I[rows];
    L->P.rows = rows;
    L->P.cols = cols;
    L->P.nnz = nnz;
    try
    {
        L->P.I.upload(P_I, (size_t)rows + 1, st);
        L->P.J.upload(P_J, (size_t)nnz, st);
        L->P.A.upload(P_A, (size_t)nnz, st);
        dev_csr_transpose(ctx, L->P, L->R);
        SA_CUDA(cudaStreamSynchronize(st));
    }
    catch (...)
    {
        delete L;
        throw;
    }
    L->have_P = true;
    *out = L;
    SA_API_END
}

/* New values for the operator of a level that owns it (the finest), same sparsity pattern:
   smpr_update_Dinv_neg / tg_smooth_interp / tg_update_coarse_operator are then re-run by the
   caller (adapt_update_operators, amg/src/adapt.cpp:171-216) -- the spectral data, the tentative
   prolongator and every table stay. */
extern "C" int sa_gpu_level_update_operator(sa_gpu_level *lev, const double *A_data)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    if (lev->A != &lev->A_own || !lev->A_own.nnz)
        SA_FAIL("sa_gpu_level_update_operator: the level does not own its operator (coarse levels "
                "alias the finer level's Ac: redo sa_gpu_rap there)");
    cudaStream_t st = lev->ctx->stream;
    SA_CUDA(cudaMemcpyAsync(lev->A_own.A.p, A_data, (size_t)lev->A_own.nnz * sizeof(double),
                            cudaMemcpyHostToDevice, st));
    SA_CUDA(cudaStreamSynchronize(st));
    lev->have_Dinv = false;
    lev->have_Ac = false;
    SA_API_END
}

extern "C" int sa_gpu_dist_rap(sa_gpu_level *lev, sa_gpu_comm *C, double *bytes_moved)
{
    SA_API_BEGIN
    sa_level_ready(lev);
    if (!C || C->nranks < 2 || !C->comm)
        SA_FAIL("sa_gpu_dist_rap: needs a communicator of at least two ranks");
    if (!lev->have_P)
        SA_FAIL("sa_gpu_dist_rap: no prolongator (call sa_gpu_smooth_P)");
    if (bytes_moved)
        *bytes_moved = 0.;
    // Ac by blocks of COARSE rows: rank q forms (R[q-th block, :] A) P -- both products local, the
    // intermediate R_q A has nnz(AP) / N entries and never leaves the rank -- and only the rows of
    // Ac are all-gathered.  (Gathering A P itself, as the first version did, moves ~40x more and
    // overflows the int32 row pointers at 256^3: nnz(A P) ~ 2.4e9.)
    {
        sa_gpu_ctx *ctx = lev->ctx;
        const DevCsr &R = lev->R;
        const int nr = C->nranks, me = C->rank;
        const int c0 = (int)(((int64_t)R.rows * me) / nr), c1 = (int)(((int64_t)R.rows * (me + 1)) / nr);
        DevCsr Rv, RA;
        Rv.rows = c1 - c0;
        Rv.cols = R.cols;
        Rv.nnz = R.nnz;
        Rv.I.view(R.I.p + c0, (size_t)(c1 - c0) + 1);
        Rv.J.view(R.J.p, (size_t)R.nnz);
        Rv.A.view(R.A.p, (size_t)R.nnz);
        dev_spgemm(ctx, Rv, *lev->A, RA);
        // rows [c0, c1) of Ac = RA P: dist_spgemm deals equal row blocks of its first factor, so
        // hand it a first factor that has rows only in this rank's block
        DevCsr RAfull;
        RAfull.rows = R.rows;
        RAfull.cols = RA.cols;
        RAfull.nnz = RA.nnz;
        RAfull.I.alloc((size_t)R.rows + 1);
        SA_LAUNCH(ctx, k_place_row_ptr, (R.rows + 1 + 255) / 256, 256, 0, R.rows, c0, c1, RA.I.p, RAfull.I.p);
        RAfull.J.view(RA.J.p, (size_t)RA.nnz);
        RAfull.A.view(RA.A.p, (size_t)RA.nnz);
        dist_spgemm(C, RAfull, lev->P, lev->Ac, bytes_moved);
    }
    lev->have_Ac = true;
    SA_API_END
}

extern "C" int sa_gpu_dist_smooth_P(sa_gpu_level *lev, sa_gpu_comm *C, int degree, const double *roots)
{
    SA_API_BEGIN
    if (!C || C->nranks < 2 || !C->comm)
        SA_FAIL("sa_gpu_dist_smooth_P: needs a communicator of at least two ranks");
    if (degree <= 0) // P = clone(tent), R = P^T: nothing to partition
        return sa_gpu_smooth_P(lev, degree, roots);
    sa_level_ready(lev);
    sa_gpu_ctx *ctx = lev->ctx;
    cudaStream_t st = ctx->stream;
    if (!lev->have_tent)
        SA_FAIL("sa_gpu_dist_smooth_P: no tentative prolongator");
    if (!lev->have_Dinv)
        SA_FAIL("sa_gpu_dist_smooth_P: sa_gpu_build_Dinv_neg has not been called");
    // the factors I - tau_k^-1 D^-1 A share A's pattern; P_(k+1) = factor * P_k by row blocks
    const DevCsr &A = *lev->A;
    DevCsr cur;
    const DevCsr *src = &lev->Ptent;
    for (int k = 0; k < degree; ++k)
    {
        DevCsr iter;
        iter.rows = A.rows;
        iter.cols = A.cols;
        iter.nnz = A.nnz;
        iter.I.view(A.I.p, (size_t)A.rows + 1);
        iter.J.view(A.J.p, (size_t)A.nnz);
        iter.A.alloc((size_t)A.nnz);
        SA_CUDA(cudaMemcpyAsync(iter.A.p, A.A.p, (size_t)A.nnz * sizeof(double), cudaMemcpyDeviceToDevice, st));
        SA_LAUNCH(ctx, k_scale_rows_add_identity_d, (A.rows + 255) / 256, 256, 0, A.rows, iter.I.p,
                  iter.J.p, iter.A.p, lev->Dinv_neg.p, 1. / roots[k]);
        DevCsr next;
        dist_spgemm(C, iter, *src, next, nullptr);
        cur.swap(next);
        src = &cur;
    }
    lev->P.swap(cur);
    dev_csr_transpose(ctx, lev->P, lev->R);
    SA_CUDA(cudaStreamSynchronize(st));
    lev->have_P = true;
    lev->have_Ac = false;
    SA_API_END
}
