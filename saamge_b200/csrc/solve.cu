// Solve phase on the device: SAS polynomial
// smoother (smpr_sym_poly / smpr_compute_poly, amg/src/smpr.cpp:213-234,
// amg/inc/smpr.hpp:319-339), V-cycle (tg_cycle_atb, amg/src/tg.cpp:91-132;
// VCycleSolver::Mult, amg/src/solve.cpp:309-323; ml_impose_cycle, amg/src/ml.cpp:361-377),
// PCG (kalchev_pcg, amg/src/mfem_addons.cpp:106-248) and the exact coarsest solve that
// stands in for the reference's --coarse-direct UMFPACK solver (amg/src/tg.cpp:991-998).
#include <algorithm>
#include <cmath>

#include "sa_gpu_internal.cuh"

#include "solve_internal.cuh"

// ---- cuBLAS (plain library TRSM / SYRK of the blocked dense coarsest factorisation), resolved
// at run time so the library has no link-time dependency on it
#include <dlfcn.h>
namespace
{
typedef void *cublasHandle_p;
struct CublasApi
{
    void *dl = nullptr;
    cublasHandle_p h = nullptr;
    int (*create)(cublasHandle_p *) = nullptr;
    int (*destroy)(cublasHandle_p) = nullptr;
    int (*set_stream)(cublasHandle_p, cudaStream_t) = nullptr;
    // cublasDtrsm_v2(handle, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb)
    int (*dtrsm)(cublasHandle_p, int, int, int, int, int, int, const double *, const double *, int,
                 double *, int) = nullptr;
    // cublasDsyrk_v2(handle, uplo, trans, n, k, alpha, A, lda, beta, C, ldc)
    int (*dsyrk)(cublasHandle_p, int, int, int, int, const double *, const double *, int,
                 const double *, double *, int) = nullptr;
    bool ok = false;
    bool load()
    {
        if (ok)
            return true;
        const char *names[] = {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12"};
        for (int i = 0; i < 3 && !dl; ++i)
            dl = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        if (!dl)
            return false;
        create = (int (*)(cublasHandle_p *))dlsym(dl, "cublasCreate_v2");
        destroy = (int (*)(cublasHandle_p))dlsym(dl, "cublasDestroy_v2");
        set_stream = (int (*)(cublasHandle_p, cudaStream_t))dlsym(dl, "cublasSetStream_v2");
        dtrsm = (decltype(dtrsm))dlsym(dl, "cublasDtrsm_v2");
        dsyrk = (decltype(dsyrk))dlsym(dl, "cublasDsyrk_v2");
        if (!create || !destroy || !set_stream || !dtrsm || !dsyrk)
            return false;
        if (create(&h) != 0)
            return false;
        ok = true;
        return true;
    }
};
CublasApi g_cublas;
// cuBLAS enum values (cublas_api.h): CUBLAS_SIDE_LEFT 0 / RIGHT 1, FILL_MODE_LOWER 0,
// OP_N 0 / OP_T 1, DIAG_NON_UNIT 0
enum { CB_LEFT = 0, CB_RIGHT = 1, CB_LOWER = 0, CB_N = 0, CB_T = 1, CB_NON_UNIT = 0 };

// Cholesky of the nb x nb diagonal block at (k0, k0) of the column-major n x n matrix, in
// shared memory, one block (the panel step of the blocked factorisation)
__global__ void k_potf2_block(int n, double *A, int k0, int nb, int *info)
{
    extern __shared__ double sb[]; // nb x nb, column-major
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x)
        sb[t] = A[(k0 + t % nb) + (int64_t)n * (k0 + t / nb)];
    __syncthreads();
    for (int k = 0; k < nb; ++k)
    {
        const double a = sb[k + nb * k];
        if (!(a > 0.))
        {
            if (threadIdx.x == 0)
                atomicCAS(info, 0, k0 + k + 1);
            return; // uniform: every thread reads the same pivot
        }
        const double pv = sqrt(a), inv = 1. / pv;
        __syncthreads();
        for (int i = k + threadIdx.x; i < nb; i += blockDim.x)
            sb[i + nb * k] = (i == k) ? pv : sb[i + nb * k] * inv;
        __syncthreads();
        const int rem = nb - k - 1;
        for (int t = threadIdx.x; t < rem * rem; t += blockDim.x)
        {
            const int i = k + 1 + t % rem, j = k + 1 + t / rem;
            if (i >= j)
                sb[i + nb * j] -= sb[i + nb * k] * sb[j + nb * k];
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x)
        if (t % nb >= t / nb)
            A[(k0 + t % nb) + (int64_t)n * (k0 + t / nb)] = sb[t];
}

__global__ void k_set_identity(int n, double *X)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < (int64_t)n * n)
        X[t] = (t % n == t / n) ? 1. : 0.;
}

__global__ void k_densify(int n, const int *I, const int *J, const double *A, double *D)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
        return;
    for (int p = I[r]; p < I[r + 1]; ++p)
        D[r + (int64_t)n * J[p]] = A[p];
}

// single block, right-looking Cholesky of the lower triangle (column-major)
__global__ void k_cholesky(int n, double *A, int *info)
{
    __shared__ double piv;
    for (int k = 0; k < n; ++k)
    {
        if (threadIdx.x == 0)
        {
            const double a = A[k + (int64_t)n * k];
            if (!(a > 0.))
            {
                *info = k + 1;
                piv = 0.;
            }
            else
            {
                piv = sqrt(a);
                A[k + (int64_t)n * k] = piv;
            }
        }
        __syncthreads();
        const double pv = piv;
        if (pv == 0.)
            return;
        const double inv = 1. / pv;
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x)
            A[i + (int64_t)n * k] *= inv;
        __syncthreads();
        // trailing update: column j gets -= L[j,k] * L[j:n,k]
        const int64_t rem = n - k - 1;
        for (int64_t t = threadIdx.x; t < rem * rem; t += blockDim.x)
        {
            const int i = k + 1 + (int)(t % rem);
            const int j = k + 1 + (int)(t / rem);
            if (i >= j)
                A[i + (int64_t)n * j] -= A[i + (int64_t)n * k] * A[j + (int64_t)n * k];
        }
        __syncthreads();
    }
}

// thread per column j of the inverse: solve L y = e_j, L^T x = y; X stored row-major
// (X[j + n*i] = x_i) so that writes coalesce
__global__ void k_chol_inverse(int n, const double *Lm, double *X)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    for (int i = 0; i < n; ++i)
    {
        double s = (i == j) ? 1. : 0.;
        if (i >= j)
        {
            for (int k = j; k < i; ++k)
                s -= Lm[i + (int64_t)n * k] * X[j + (int64_t)n * k];
            s /= Lm[i + (int64_t)n * i];
        }
        else
            s = 0.;
        X[j + (int64_t)n * i] = s;
    }
    for (int i = n - 1; i >= 0; --i)
    {
        double s = X[j + (int64_t)n * i];
        for (int k = i + 1; k < n; ++k)
            s -= Lm[k + (int64_t)n * i] * X[j + (int64_t)n * k];
        X[j + (int64_t)n * i] = s / Lm[i + (int64_t)n * i];
    }
}

__global__ void k_symmetrize(int n, const double *X, double *S)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * n)
        return;
    const int i = (int)(t % n), j = (int)(t / n);
    S[t] = 0.5 * (X[i + (int64_t)n * j] + X[j + (int64_t)n * i]);
}

// y = S x, S dense symmetric n x n; one warp per row (rows read as columns: coalesced)
__global__ void k_dense_symv(int n, const double *S, const double *x, double *y)
{
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n)
        return;
    const double *col = S + (int64_t)n * row;
    double s = 0.;
    for (int i = lane; i < n; i += 32)
        s += col[i] * x[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0)
        y[row] = s;
}

// out[slot] = sum a_i b_i (two-stage: per-block partials then atomicAdd)
__global__ void k_dot(int n, const double *a, const double *b, double *out)
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
            out[blockIdx.x] = t; // per-block partial; summed in block order by k_dot_final
    }
}

// deterministic second pass: one warp adds the block partials in a fixed order
__global__ void k_dot_final(int nblocks, const double *partials, double *out)
{
    double s = 0.;
    for (int i = threadIdx.x; i < nblocks; i += 32)
        s += partials[i];
    for (int o = 16; o > 0; o >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0)
        *out = s;
}

// x += alpha d ; r -= alpha z
__global__ void k_pcg_update(int n, double alpha, const double *d, const double *z, double *x,
                             double *r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    x[i] = x[i] + alpha * d[i];
    r[i] = r[i] - alpha * z[i];
}

// d = z + beta d
__global__ void k_pcg_dir(int n, double beta, const double *z, double *d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    d[i] = z[i] + beta * d[i];
}

// the same updates with alpha = *nom / *den and beta = *betanom / *nom read from device scalars
// (the host reads one scalar per iteration, for the convergence test)
__global__ void k_pcg_update_dev(int n, const double *nom, const double *den, const double *d,
                                 const double *z, double *x, double *r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double alpha = *nom / *den;
    x[i] = x[i] + alpha * d[i];
    r[i] = r[i] - alpha * z[i];
}
__global__ void k_pcg_dir_dev(int n, const double *betanom, const double *nom, const double *z,
                              double *d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double beta = *betanom / *nom;
    d[i] = z[i] + beta * d[i];
}

// (a, b) into the device scalar S->dots[slot]; no host synchronisation
void dev_dot_async(sa_gpu_solver *S, int n, const double *a, const double *b, int slot)
{
    sa_gpu_ctx *ctx = S->ctx;
    const int blocks = std::max(1, std::min(ctx->num_sms * 4, (n + 255) / 256));
    S->dot_partials.ensure(blocks);
    SA_LAUNCH(ctx, k_dot, blocks, 256, 0, n, a, b, S->dot_partials.p);
    SA_LAUNCH(ctx, k_dot_final, 1, 32, 0, blocks, S->dot_partials.p, S->dots.p + slot);
}
double dev_read_scalar(sa_gpu_solver *S, int slot)
{
    double h = 0.;
    SA_CUDA(cudaMemcpyAsync(&h, S->dots.p + slot, sizeof(double), cudaMemcpyDeviceToHost,
                            S->ctx->stream));
    SA_CUDA(cudaStreamSynchronize(S->ctx->stream));
    return h;
}

double dev_dot(sa_gpu_solver *S, int n, const double *a, const double *b)
{
    sa_gpu_ctx *ctx = S->ctx;
    const int blocks = std::max(1, std::min(ctx->num_sms * 4, (n + 255) / 256));
    S->dot_partials.ensure(blocks);
    SA_LAUNCH(ctx, k_dot, blocks, 256, 0, n, a, b, S->dot_partials.p);
    SA_LAUNCH(ctx, k_dot_final, 1, 32, 0, blocks, S->dot_partials.p, S->dots.p);
    double h = 0.;
    SA_CUDA(cudaMemcpyAsync(&h, S->dots.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(cudaStreamSynchronize(ctx->stream));
    return h;
}

// smpr_compute_poly on device buffers; result ends in *xcur (pointer swap)
void dev_poly_smooth(sa_gpu_ctx *ctx, const DevCsr &A, const double *dinv, const double *b,
                     double **xcur, double **xalt, int degree, const double *roots,
                     bool x_is_zero)
{
    for (int i = 0; i < degree; ++i)
    {
        const double mult = 1. / roots[i];
        dev_smoother_step(ctx, A, dinv, b, *xcur, *xalt, mult, (x_is_zero && i == 0) ? 1 : 0);
        std::swap(*xcur, *xalt);
    }
}

// tg_cycle_atb on level l; rhs in SL.b; result pointer returned (xa or xb)
double *dev_vcycle(sa_gpu_solver *S, int l)
{
    sa_gpu_ctx *ctx = S->ctx;
    SolverLevel &SL = *S->L[l];
    sa_gpu_level *lev = SL.lev;
    const DevCsr &A = *lev->A;
    double *xcur = SL.xa.p, *xalt = SL.xb.p;
    const int nl = lev->ND;
    // a user smoother runs on the host: vectors go down and the iterate comes back (slow path,
    // there for API completeness; the default never leaves the device)
    auto host_smooth = [&](sa_gpu_smoother_ft f, bool x_is_zero) {
        cudaStream_t st = ctx->stream;
        SL.hb.resize(nl);
        SL.hx.assign(nl, 0.);
        SL.b.download(SL.hb.data(), nl, st);
        if (!x_is_zero)
            SA_CUDA(cudaMemcpyAsync(SL.hx.data(), xcur, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, st));
        SA_CUDA(cudaStreamSynchronize(st));
        f(l, nl, SL.hb.data(), SL.hx.data(), SL.smoother_data);
        SA_CUDA(cudaMemcpyAsync(xcur, SL.hx.data(), (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, st));
    };
    // x = 0 (iterative_mode == false); pre-smoother
    if (SL.pre)
        host_smooth(SL.pre, true);
    else
        dev_poly_smooth(ctx, A, lev->Dinv_neg.p, SL.b.p, &xcur, &xalt, S->degree, S->roots.data(),
                        true);
    // res = b - A x ; resc = restr * res
    dev_residual(ctx, A, xcur, SL.b.p, SL.r.p);
    double *xc = nullptr;
    if (l + 1 < (int)S->L.size())
    {
        SolverLevel &SC = *S->L[l + 1];
        dev_spmv(ctx, lev->R, SL.r.p, SC.b.p);
        xc = dev_vcycle(S, l + 1);
    }
    else
    {
        dev_spmv(ctx, lev->R, SL.r.p, S->bc.p);
        SA_LAUNCH(ctx, k_dense_symv, (S->nc + 7) / 8, 256, 0, S->nc, S->Ainv.p, S->bc.p, S->xc.p);
        xc = S->xc.p;
    }
    // x += interp * xc
    dev_spmv_add(ctx, lev->P, xc, xcur);
    // post-smoother
    if (SL.post)
        host_smooth(SL.post, false);
    else
        dev_poly_smooth(ctx, A, lev->Dinv_neg.p, SL.b.p, &xcur, &xalt, S->degree, S->roots.data(),
                        false);
    return xcur;
}

} // namespace

extern "C" int sa_gpu_solver_create(sa_gpu_ctx *ctx, sa_gpu_level **levels, int nlevels,
                                    int nu_relax, sa_gpu_solver **out)
{
    SA_API_BEGIN
    if (nlevels < 1 || nu_relax < 1)
        SA_FAIL("sa_gpu_solver_create: bad arguments");
    for (int l = 0; l < nlevels; ++l)
        sa_level_ready(levels[l]);
    sa_gpu_solver *S = new sa_gpu_solver;
    S->ctx = ctx;
    struct Guard
    {
        sa_gpu_solver *s;
        ~Guard() { delete s; }
    } g{S};
    cudaStream_t st = ctx->stream;
    // smpr_sas_poly_roots (amg/src/smpr.cpp:282-306)
    {
        const int nu = nu_relax, twonu = 2 * nu;
        const double denom = (double)(2 * nu + 1);
        S->degree = twonu + nu + 1;
        S->roots.resize(S->degree);
        for (int i = 0; i <= twonu; ++i)
        {
            const double val = cos(((double)i * M_PI) / denom);
            S->roots[i] = val * val;
        }
        for (int i = 1; i <= nu; ++i)
        {
            const double val = sin(((double)i * M_PI) / denom);
            S->roots[i + twonu] = val * val;
        }
    }
    for (int l = 0; l < nlevels; ++l)
    {
        sa_gpu_level *lev = levels[l];
        if (!lev->have_Ac || !lev->have_P || !lev->have_Dinv)
            SA_FAIL("sa_gpu_solver_create: level %d is not fully set up", l);
        SolverLevel *SL = new SolverLevel;
        S->L.push_back(SL);
        SL->lev = lev;
        SL->b.alloc(lev->ND);
        SL->xa.alloc(lev->ND);
        SL->xb.alloc(lev->ND);
        SL->r.alloc(lev->ND);
    }
    // exact coarsest solve: dense inverse of the last Ac through Cholesky
    {
        const DevCsr &Ac = levels[nlevels - 1]->Ac;
        const int n = Ac.rows;
        if (n > 20000)
            SA_FAIL("sa_gpu_solver_create: coarsest operator has %d rows; the exact dense "
                    "coarsest solve supports at most 20000 -- add levels", n);
        S->nc = n;
        DevBuf<double> Lm, X;
        DevBuf<int> info;
        Lm.alloc((size_t)n * n);
        Lm.zero(st);
        X.alloc((size_t)n * n);
        info.alloc(1);
        info.zero(st);
        S->Ainv.alloc((size_t)n * n);
        S->bc.alloc(n);
        S->xc.alloc(n);
        if (n)
        {
            SA_LAUNCH(ctx, k_densify, (n + 255) / 256, 256, 0, n, Ac.I.p, Ac.J.p, Ac.A.p, Lm.p);
            const int64_t nn = (int64_t)n * n;
            static const int blocked_min =
                getenv("SA_GPU_COARSE_BLOCKED_MIN") ? atoi(getenv("SA_GPU_COARSE_BLOCKED_MIN")) : 2500;
            if (n >= blocked_min && g_cublas.load())
            {
                // blocked right-looking Cholesky: own kernel for the diagonal block, library
                // TRSM / SYRK for the panel and the trailing update; then A^-1 = L^-T L^-1 by
                // two triangular solves with the identity
                const int NB = 64;
                const double one = 1., minus1 = -1.;
                g_cublas.set_stream(g_cublas.h, st);
                for (int k0 = 0; k0 < n; k0 += NB)
                {
                    const int nb = std::min(NB, n - k0), rem = n - k0 - nb;
                    SA_LAUNCH(ctx, k_potf2_block, 1, 256, (size_t)nb * nb * sizeof(double), n, Lm.p,
                              k0, nb, info.p);
                    if (rem > 0)
                    {
                        double *L11 = Lm.p + k0 + (int64_t)n * k0;
                        double *L21 = Lm.p + (k0 + nb) + (int64_t)n * k0;
                        double *A22 = Lm.p + (k0 + nb) + (int64_t)n * (k0 + nb);
                        if (g_cublas.dtrsm(g_cublas.h, CB_RIGHT, CB_LOWER, CB_T, CB_NON_UNIT, rem, nb,
                                           &one, L11, n, L21, n) != 0 ||
                            g_cublas.dsyrk(g_cublas.h, CB_LOWER, CB_N, rem, nb, &minus1, L21, n, &one,
                                           A22, n) != 0)
                            SA_FAIL("sa_gpu_solver_create: cuBLAS TRSM/SYRK failed");
                    }
                }
                SA_LAUNCH(ctx, k_set_identity, (unsigned)((nn + 255) / 256), 256, 0, n, X.p);
                if (g_cublas.dtrsm(g_cublas.h, CB_LEFT, CB_LOWER, CB_N, CB_NON_UNIT, n, n, &one, Lm.p, n,
                                   X.p, n) != 0 ||
                    g_cublas.dtrsm(g_cublas.h, CB_LEFT, CB_LOWER, CB_T, CB_NON_UNIT, n, n, &one, Lm.p, n,
                                   X.p, n) != 0)
                    SA_FAIL("sa_gpu_solver_create: cuBLAS TRSM failed");
            }
            else
            {
                SA_LAUNCH(ctx, k_cholesky, 1, 1024, 0, n, Lm.p, info.p);
                SA_LAUNCH(ctx, k_chol_inverse, (n + 63) / 64, 64, 0, n, Lm.p, X.p);
            }
            SA_LAUNCH(ctx, k_symmetrize, (unsigned)((nn + 255) / 256), 256, 0, n, X.p, S->Ainv.p);
        }
        int h = 0;
        info.download(&h, 1, st);
        SA_CUDA(cudaStreamSynchronize(st));
        if (h)
            SA_FAIL("sa_gpu_solver_create: coarsest operator is not positive definite "
                    "(pivot %d)", h);
    }
    const int n0 = levels[0]->ND;
    S->pb.alloc(n0);
    S->px.alloc(n0);
    S->pr.alloc(n0);
    S->pd.alloc(n0);
    S->pz.alloc(n0);
    S->dots.alloc(4);
    g.s = nullptr;
    *out = S;
    SA_API_END
}

extern "C" void sa_gpu_solver_destroy(sa_gpu_solver *S) { delete S; }

extern "C" int sa_gpu_solver_set_smoothers(sa_gpu_solver *S, int level, sa_gpu_smoother_ft pre,
                                           sa_gpu_smoother_ft post, void *data)
{
    SA_API_BEGIN
    if (level < 0 || level >= (int)S->L.size())
        SA_FAIL("sa_gpu_solver_set_smoothers: bad level %d", level);
    S->L[level]->pre = pre;
    S->L[level]->post = post;
    S->L[level]->smoother_data = data;
    SA_API_END
}

extern "C" int sa_gpu_vcycle(sa_gpu_solver *S, const double *b, double *x)
{
    SA_API_BEGIN
    cudaStream_t st = S->ctx->stream;
    const int n = S->L[0]->lev->ND;
    S->L[0]->b.upload(b, n, st);
    double *res = dev_vcycle(S, 0);
    SA_CUDA(cudaMemcpyAsync(x, res, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

// B.Mult(r, z) with device vectors of level 0
static void precond(sa_gpu_solver *S, const double *r, double *z)
{
    cudaStream_t st = S->ctx->stream;
    const int n = S->L[0]->lev->ND;
    SA_CUDA(cudaMemcpyAsync(S->L[0]->b.p, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice,
                            st));
    double *res = dev_vcycle(S, 0);
    SA_CUDA(cudaMemcpyAsync(z, res, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
}

// kalchev_pcg on device-resident S->pb (b) and S->px (x)
static int pcg_resident(sa_gpu_solver *S, int max_num_iter, double RTOLERANCE, double ATOLERANCE,
                        double *brr_hist, int hist_cap, int *hist_len)
{
    sa_gpu_ctx *ctx = S->ctx;
    const DevCsr &A = *S->L[0]->lev->A;
    const int dim = A.rows;
    const int tb = 256, gb = (dim + tb - 1) / tb;
    double *x = S->px.p, *b = S->pb.p, *r = S->pr.p, *d = S->pd.p, *z = S->pz.p;
    int i, iters = 0, hl = 0;
    double r0, den, nom, betanom = 0.;

    dev_residual(ctx, A, x, b, r); // r = b - A x
    precond(S, r, z);
    SA_CUDA(cudaMemcpyAsync(d, z, (size_t)dim * sizeof(double), cudaMemcpyDeviceToDevice,
                            ctx->stream));
    // device scalars: dots[1 + s_nom] = nom, dots[3] = den, dots[1 + s_beta] = betanom
    int s_nom = 0, s_beta = 1;
    dev_dot_async(S, dim, z, r, 1 + s_nom);
    nom = dev_read_scalar(S, 1 + s_nom);
    if (brr_hist && hl < hist_cap)
        brr_hist[hl++] = nom;
    if (hist_len)
        *hist_len = hl;
    if ((r0 = nom * RTOLERANCE) < ATOLERANCE)
        r0 = ATOLERANCE;
    if (nom < r0)
        return -1;
    dev_spmv(ctx, A, d, z);
    dev_dot_async(S, dim, z, d, 3);
    den = dev_read_scalar(S, 3);
    if (0. == den)
        return -1;
    for (i = 1; i <= max_num_iter; i++)
    {
        SA_LAUNCH(ctx, k_pcg_update_dev, gb, tb, 0, dim, S->dots.p + 1 + s_nom, S->dots.p + 3, d, z, x, r);
        precond(S, r, z);
        dev_dot_async(S, dim, r, z, 1 + s_beta);
        betanom = dev_read_scalar(S, 1 + s_beta); // the one host read per iteration
        if (brr_hist && hl < hist_cap)
            brr_hist[hl++] = betanom;
        if (betanom < 0.0)
        {
            iters = -i;
            break;
        }
        if (betanom < r0)
        {
            iters = i;
            break;
        }
        SA_LAUNCH(ctx, k_pcg_dir_dev, gb, tb, 0, dim, S->dots.p + 1 + s_beta, S->dots.p + 1 + s_nom, z, d);
        dev_spmv(ctx, A, d, z);
        dev_dot_async(S, dim, d, z, 3);
        std::swap(s_nom, s_beta);
    }
    if (i > max_num_iter)
        iters = -(i - 1);
    if (hist_len)
        *hist_len = hl;
    return iters;
}

extern "C" int sa_gpu_pcg(sa_gpu_solver *S, const double *b, double *x, int maxiter, double rtol,
                          double atol, int *iters, double *brr_hist, int hist_cap, int *hist_len)
{
    SA_API_BEGIN
    cudaStream_t st = S->ctx->stream;
    const int n = S->L[0]->lev->ND;
    S->pb.upload(b, n, st);
    S->px.upload(x, n, st);
    *iters = pcg_resident(S, maxiter, rtol, atol, brr_hist, hist_cap, hist_len);
    S->px.download(x, n, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

extern "C" int sa_gpu_solver_upload(sa_gpu_solver *S, const double *b, const double *x0)
{
    SA_API_BEGIN
    cudaStream_t st = S->ctx->stream;
    const int n = S->L[0]->lev->ND;
    S->pb.upload(b, n, st);
    if (x0)
        S->px.upload(x0, n, st);
    else
        S->px.zero(st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

extern "C" int sa_gpu_pcg_resident(sa_gpu_solver *S, int maxiter, double rtol, double atol,
                                   int *iters)
{
    SA_API_BEGIN
    *iters = pcg_resident(S, maxiter, rtol, atol, nullptr, 0, nullptr);
    SA_CUDA(cudaStreamSynchronize(S->ctx->stream));
    SA_API_END
}

extern "C" int sa_gpu_solver_download(sa_gpu_solver *S, double *x)
{
    SA_API_BEGIN
    S->px.download(x, S->L[0]->lev->ND, S->ctx->stream);
    SA_CUDA(cudaStreamSynchronize(S->ctx->stream));
    SA_API_END
}

extern "C" int sa_gpu_solver_dev_coarse(sa_gpu_solver *S, const double *b, double *x)
{
    SA_API_BEGIN
    if (S->nc)
        SA_LAUNCH(S->ctx, k_dense_symv, (S->nc + 7) / 8, 256, 0, S->nc, S->Ainv.p, b, x);
    SA_API_END
}

extern "C" int sa_gpu_solver_info(sa_gpu_solver *S, int *degree, double *roots, int cap, int *nc)
{
    SA_API_BEGIN
    *degree = S->degree;
    for (int i = 0; i < S->degree && i < cap; ++i)
        roots[i] = S->roots[i];
    *nc = S->nc;
    SA_API_END
}
