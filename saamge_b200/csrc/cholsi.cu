// Lower end of the spectrum of LARGE AE matrices without a tridiagonalisation (coarse levels,
// n ~ 10^3, and order-2 fine levels, n ~ 700): what xpacks_calc_lower_eigens_dense
// (amg/src/xpacks.cpp:222-314) asks of dsygvx is "every eigenpair with lambda in (-1, theta]" --
// a handful of pairs (128^3: 774 vectors for the 630 level-1 AEs) of a matrix whose spectrum lies
// in [0, 1].  dsytrd + dstebz + dstein spend 4/3 n^3 flops (half of them BLAS-2) to get them; here
//
//   k_cs_chol     M = A^ - sigma I = L L^T, sigma = -max(theta / 10, 1e-6) < 0 <= lambda_min:
//                 blocked right-looking Cholesky, n^3 / 3 flops, all of the O(n^3) part a rank-32
//                 update of the trailing lower triangle on the FP64 tensor pipe (DMMA m8n8k4,
//                 operands straight from the panel, one warp per 32 x 32 tile).  One thread block
//                 per matrix, or -- a handful of huge matrices (128^3 level 2: 10 AEs of n ~ 4000)
//                 -- a group of G co-resident blocks per matrix (cooperative launch, two group
//                 barriers per panel, tiles and panel rows dealt to the blocks).
//   k_cs_iterate  subspace iteration with M^-1 on K = 8 vectors: Z = L^-T L^-1 X (blocked
//                 triangular solves, two passes over L per iteration), Rayleigh-Ritz on span(Z)
//                 WITHOUT another pass over the matrix (M Z = X, so Z^T M Z = Z^T X and the
//                 residual of a Ritz pair is X q - mu Z q), until every pair with lambda <= theta
//                 has residual <= 1e-13 (||M|| ~ 1) and the first Ritz value above theta exceeds
//                 it by more than its residual.
//
// The count m = #{lambda <= theta} is the number of converged Ritz values <= theta; a matrix with
// all K Ritz values <= theta, a non-positive pivot (A^ - sigma I not SPD) or no convergence is
// reported and the caller sends the chunk through the two-stage tridiagonalisation (twostage.cu)
// instead, so the decisions are those of the reference in every case.
#include <algorithm>

#include <cooperative_groups.h>

#include "cholsi.cuh"

namespace
{
constexpr int CS_K = SA_CS_K;
constexpr int CS_NB = 32;
constexpr int CS_NT = 512;
constexpr int CS_NW = CS_NT / 32;
constexpr int CS_LD = 33;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool MULTI> __device__ __forceinline__ double ldt(const double *p)
{
    // (several blocks per matrix: data written by another SM must not come from this SM's L1)
    return MULTI ? __ldcg(p) : *p;
}

/* barrier of the G blocks of one matrix group (all co-resident: cooperative launch) */
__device__ __forceinline__ void group_barrier(unsigned int *counter, int G, unsigned int &target)
{
    __syncthreads();
    if (G > 1)
    {
        if (threadIdx.x == 0)
        {
            __threadfence();
            atomicAdd(counter, 1u);
            target += (unsigned int)G;
            while (*(volatile unsigned int *)counter < target)
                ;
            __threadfence();
        }
        __syncthreads();
    }
}

/* M = T - sigma I = L L^T in the lower triangle of T (column-major, ld = n).  Returns false on a
   non-positive pivot (decided identically by every block of the group). */
template <bool MULTI>
__device__ bool chol_one(double *__restrict__ T, int n, double sigma, int G, int gr,
                         unsigned int *counter, unsigned int &target, double *Ls, double *rdiag,
                         int *flag)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const size_t ld = (size_t)n;
    // shift (rows dealt to the blocks)
    for (int i = gr * CS_NT + tid; i < n; i += G * CS_NT)
        T[i + ld * i] -= sigma;
    group_barrier(counter, G, target);
    for (int j0 = 0; j0 < n; j0 += CS_NB)
    {
        const int jb = min(CS_NB, n - j0);
        // 1. diagonal block (every block of the group factors its own copy)
        for (int idx = tid; idx < CS_NB * CS_NB; idx += CS_NT)
        {
            const int i = idx & 31, c = idx >> 5;
            double v = (i == c) ? 1. : 0.;
            if (i < jb && c < jb && i >= c)
                v = ldt<MULTI>(T + (j0 + i) + ld * (j0 + c));
            Ls[i * CS_LD + c] = v;
        }
        if (tid == 0)
            *flag = 0;
        __syncthreads();
        if (wid == 0)
        {
            for (int k = 0; k < CS_NB; ++k)
            {
                const double p = Ls[k * CS_LD + k];
                if (!(p > 0.))
                {
                    if (lane == 0)
                        *flag = 1;
                    break;
                }
                const double s = sqrt(p);
                double lik = 0.;
                if (lane > k)
                {
                    lik = Ls[lane * CS_LD + k] / s;
                    Ls[lane * CS_LD + k] = lik;
                }
                else if (lane == k)
                {
                    Ls[k * CS_LD + k] = s;
                    rdiag[k] = 1. / s;
                }
                __syncwarp();
                for (int j = k + 1; j < CS_NB; ++j)
                    if (lane >= j)
                        Ls[lane * CS_LD + j] -= lik * Ls[j * CS_LD + k];
                __syncwarp();
            }
        }
        __syncthreads();
        if (*flag)
            return false;
        const int r0 = j0 + jb;
        // 2. panel below the diagonal block: X L^T = B, one row per thread
        for (int r = r0 + gr * CS_NT + tid; r < n; r += G * CS_NT)
        {
            double x[CS_NB];
#pragma unroll
            for (int c = 0; c < CS_NB; ++c)
                x[c] = (c < jb) ? ldt<MULTI>(T + r + ld * (j0 + c)) : 0.;
#pragma unroll
            for (int c = 0; c < CS_NB; ++c)
            {
                double s = x[c];
#pragma unroll
                for (int q = 0; q < c; ++q)
                    s -= x[q] * Ls[c * CS_LD + q];
                x[c] = s * rdiag[c];
            }
#pragma unroll
            for (int c = 0; c < CS_NB; ++c)
                if (c < jb)
                    T[r + ld * (j0 + c)] = x[c];
        }
        group_barrier(counter, G, target);
        // (every block of the group has read the diagonal block by now: the factor replaces it)
        if (gr == 0)
            for (int idx = tid; idx < CS_NB * CS_NB; idx += CS_NT)
            {
                const int i = idx & 31, c = idx >> 5;
                if (i < jb && c < jb && i >= c)
                    T[(j0 + i) + ld * (j0 + c)] = Ls[i * CS_LD + c];
            }
        if (r0 >= n)
            break;
        // 3. trailing lower triangle -= P P^T, one warp per 32 x 32 tile (DMMA)
        const int r = n - r0;
        const int nt = (r + 31) >> 5;
        const int ntiles = nt * (nt + 1) / 2;
        const double *P = T + ld * j0; // panel columns
        for (int tl = gr * CS_NW + wid; tl < ntiles; tl += G * CS_NW)
        {
            int ti = (int)((sqrt(8. * (double)tl + 1.) - 1.) * 0.5);
            while ((ti + 1) * (ti + 2) / 2 <= tl)
                ++ti;
            while (ti * (ti + 1) / 2 > tl)
                --ti;
            const int tj = tl - ti * (ti + 1) / 2;
            const int i0 = r0 + ti * 32, c0 = r0 + tj * 32;
            double c[4][4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
            {
                const int row = i0 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    const int col = c0 + ni * 8 + 2 * t;
                    c[mi][ni][0] = (row < n && col <= row) ? ldt<MULTI>(T + row + ld * col) : 0.;
                    c[mi][ni][1] = (row < n && col + 1 <= row) ? ldt<MULTI>(T + row + ld * (col + 1)) : 0.;
                }
            }
#pragma unroll 2
            for (int kk = 0; kk < 8; ++kk)
            {
                const size_t ko = ld * (size_t)(kk * 4 + t);
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
                {
                    const int row = i0 + mi * 8 + g;
                    a[mi] = (row < n) ? -ldt<MULTI>(P + row + ko) : 0.;
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    const int row = c0 + ni * 8 + g;
                    b[ni] = (row < n) ? ldt<MULTI>(P + row + ko) : 0.;
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                        dmma884(c[mi][ni][0], c[mi][ni][1], a[mi], b[ni]);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
            {
                const int row = i0 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    const int col = c0 + ni * 8 + 2 * t;
                    if (row < n && col <= row)
                        T[row + ld * col] = c[mi][ni][0];
                    if (row < n && col + 1 <= row)
                        T[row + ld * (col + 1)] = c[mi][ni][1];
                }
            }
        }
        group_barrier(counter, G, target);
    }
    return true;
}

template <bool MULTI>
__global__ void __launch_bounds__(CS_NT, 1)
k_cs_chol(const sa_cs_mat *mats, int nmats, double sigma, int G, unsigned int *counters,
          unsigned int *queue)
{
    __shared__ double Ls[CS_NB * CS_LD];
    __shared__ double rdiag[CS_NB];
    __shared__ int flag;
    __shared__ int next;
    if (!MULTI)
    {
        // one block per matrix, work queue (largest first)
        unsigned int target = 0;
        while (true)
        {
            if (threadIdx.x == 0)
                next = (int)atomicAdd(queue, 1u);
            __syncthreads();
            const int b = next;
            __syncthreads();
            if (b >= nmats)
                return;
            const sa_cs_mat M = mats[b];
            const bool ok = chol_one<false>(M.T, M.n, sigma, 1, 0, nullptr, target, Ls, rdiag, &flag);
            if (threadIdx.x == 0)
                M.info[0] = ok ? 0 : -1;
            __syncthreads();
        }
    }
    else
    {
        const int ngroups = gridDim.x / G;
        const int gi = blockIdx.x / G, gr = blockIdx.x % G;
        if (gi >= ngroups)
            return;
        unsigned int target = 0;
        for (int b = gi; b < nmats; b += ngroups)
        {
            const sa_cs_mat M = mats[b];
            const bool ok = chol_one<true>(M.T, M.n, sigma, G, gr, counters + gi, target, Ls, rdiag, &flag);
            if (threadIdx.x == 0 && gr == 0)
                M.info[0] = ok ? 0 : -1;
            // (a failed factorisation leaves the group at the same point in every block)
            group_barrier(counters + gi, G, target);
        }
    }
}

/* ---- subspace iteration ------------------------------------------------------------------ */

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* sums NV per-thread values over the block; the totals land in out[0..NV) (shared) */
template <int NV> __device__ void block_sum(double (&v)[NV], double *red, double *out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i)
    {
        const double s = warp_sum(v[i]);
        if (lane == 0)
            red[wid * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV)
    {
        double s = 0.;
        for (int w = 0; w < CS_NW; ++w)
            s += red[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

/* symmetric K x K eigenproblem by cyclic Jacobi (one thread): C -> eigenvalues on the diagonal,
   V = eigenvectors (columns) */
__device__ void jacobi_small(double (*C)[CS_K], double (*V)[CS_K])
{
    for (int i = 0; i < CS_K; ++i)
        for (int j = 0; j < CS_K; ++j)
            V[i][j] = (i == j) ? 1. : 0.;
    for (int sweep = 0; sweep < 40; ++sweep)
    {
        double off = 0., dia = 0.;
        for (int i = 0; i < CS_K; ++i)
        {
            dia += C[i][i] * C[i][i];
            for (int j = 0; j < i; ++j)
                off += C[i][j] * C[i][j];
        }
        if (off <= 1e-34 * dia || off == 0.)
            break;
        for (int p = 0; p < CS_K - 1; ++p)
            for (int q = p + 1; q < CS_K; ++q)
            {
                const double apq = C[p][q];
                if (apq == 0.)
                    continue;
                const double zeta = (C[q][q] - C[p][p]) / (2. * apq);
                const double tt = copysign(1., zeta) / (fabs(zeta) + sqrt(1. + zeta * zeta));
                const double cs = 1. / sqrt(1. + tt * tt), sn = cs * tt;
                for (int k = 0; k < CS_K; ++k)
                {
                    const double kp = C[k][p], kq = C[k][q];
                    C[k][p] = cs * kp - sn * kq;
                    C[k][q] = sn * kp + cs * kq;
                }
                for (int k = 0; k < CS_K; ++k)
                {
                    const double pk = C[p][k], qk = C[q][k];
                    C[p][k] = cs * pk - sn * qk;
                    C[q][k] = sn * pk + cs * qk;
                }
                for (int k = 0; k < CS_K; ++k)
                {
                    const double kp = V[k][p], kq = V[k][q];
                    V[k][p] = cs * kp - sn * kq;
                    V[k][q] = sn * kp + cs * kq;
                }
            }
    }
}

struct CsSmall
{
    double Gm[CS_K][CS_K], Hm[CS_K][CS_K], Q[CS_K][CS_K], Cm[CS_K][CS_K], V[CS_K][CS_K];
    double mu[CS_K];
    int perm[CS_K];
    int bad;
};

__device__ __forceinline__ double cs_start_value(int r, int k, int salt)
{
    // deterministic pseudo-random start vectors (the first one is constant: the scaled kernel
    // vector of an interior AE is close to it)
    if (k == 0)
        return 1.;
    unsigned int h = (unsigned int)r * 2654435761u ^ ((unsigned int)(k + 8 * salt) * 40503u + 0x9e3779b9u);
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    h *= 3266489917u;
    h ^= h >> 16;
    return ((double)(h & 0xffffffu) / 8388608.) - 1.;
}

__global__ void __launch_bounds__(CS_NT, 1)
k_cs_iterate(const sa_cs_mat *mats, int nmats, double sigma, double theta, int max_its, double tol)
{
    __shared__ double Ls[CS_NB * CS_LD];
    __shared__ double rinv[CS_NB];
    __shared__ double Ys[CS_NB][CS_K];
    __shared__ double red[CS_NW * 2 * CS_K];
    __shared__ double tot[2 * CS_K];
    __shared__ CsSmall S;
    __shared__ int done;
    const int b = blockIdx.x;
    if (b >= nmats)
        return;
    const sa_cs_mat M = mats[b];
    const int n = M.n;
    const size_t ld = (size_t)n;
    const double *__restrict__ T = M.T;
    double *__restrict__ X = M.X, *__restrict__ Z = M.Z;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (M.info[0] != 0)
        return; // factorisation failed
    for (int r = tid; r < n; r += CS_NT)
#pragma unroll
        for (int k = 0; k < CS_K; ++k)
            X[(size_t)k * n + r] = cs_start_value(r, k, b);
    __syncthreads();
    int its = 0, result = -3; // -3: no convergence
    for (its = 1; its <= max_its; ++its)
    {
        // Z = X
        for (int idx = tid; idx < CS_K * n; idx += CS_NT)
            Z[idx] = X[idx];
        __syncthreads();
        // forward: L Y = Z
        for (int jb = 0; jb < n; jb += CS_NB)
        {
            const int w = min(CS_NB, n - jb);
            for (int idx = tid; idx < CS_NB * CS_NB; idx += CS_NT)
            {
                const int i = idx & 31, c = idx >> 5;
                double v = (i == c) ? 1. : 0.;
                if (i < w && c < w && i >= c)
                    v = T[(jb + i) + ld * (jb + c)];
                Ls[i * CS_LD + c] = v;
                if (i == c)
                    rinv[i] = 1. / v;
            }
            __syncthreads();
            if (wid < CS_K)
            {
                double bi = (lane < w) ? Z[(size_t)wid * n + jb + lane] : 0.;
                for (int c = 0; c < CS_NB; ++c)
                {
                    const double tc = __shfl_sync(0xffffffffu, bi, c) * rinv[c];
                    if (lane == c)
                        bi = tc;
                    else if (lane > c)
                        bi -= Ls[lane * CS_LD + c] * tc;
                }
                Ys[lane][wid] = bi;
                if (lane < w)
                    Z[(size_t)wid * n + jb + lane] = bi;
            }
            __syncthreads();
            for (int r = jb + CS_NB + tid; r < n; r += CS_NT)
            {
                double acc[CS_K];
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                    acc[k] = Z[(size_t)k * n + r];
#pragma unroll 8
                for (int c = 0; c < CS_NB; ++c)
                {
                    const double l = T[r + ld * (jb + c)];
#pragma unroll
                    for (int k = 0; k < CS_K; ++k)
                        acc[k] -= l * Ys[c][k];
                }
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                    Z[(size_t)k * n + r] = acc[k];
            }
            __syncthreads();
        }
        // backward: L^T Z = Y
        for (int jb = ((n - 1) / CS_NB) * CS_NB; jb >= 0; jb -= CS_NB)
        {
            const int w = min(CS_NB, n - jb);
            for (int idx = tid; idx < CS_NB * CS_NB; idx += CS_NT)
            {
                const int i = idx & 31, c = idx >> 5;
                double v = (i == c) ? 1. : 0.;
                if (i < w && c < w && i >= c)
                    v = T[(jb + i) + ld * (jb + c)];
                Ls[i * CS_LD + c] = v;
                if (i == c)
                    rinv[i] = 1. / v;
            }
            __syncthreads();
            if (wid < CS_K)
            {
                double bi = (lane < w) ? Z[(size_t)wid * n + jb + lane] : 0.;
                for (int c = CS_NB - 1; c >= 0; --c)
                {
                    const double tc = __shfl_sync(0xffffffffu, bi, c) * rinv[c];
                    if (lane == c)
                        bi = tc;
                    else if (lane < c)
                        bi -= Ls[c * CS_LD + lane] * tc;
                }
                Ys[lane][wid] = bi;
                if (lane < w)
                    Z[(size_t)wid * n + jb + lane] = bi;
            }
            __syncthreads();
            for (int cc = tid; cc < jb; cc += CS_NT)
            {
                double acc[CS_K];
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                    acc[k] = Z[(size_t)k * n + cc];
                const double *col = T + jb + ld * cc;
#pragma unroll 8
                for (int q = 0; q < CS_NB; ++q)
                {
                    const double l = (q < w) ? col[q] : 0.;
#pragma unroll
                    for (int k = 0; k < CS_K; ++k)
                        acc[k] -= l * Ys[q][k];
                }
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                    Z[(size_t)k * n + cc] = acc[k];
            }
            __syncthreads();
        }
        // Rayleigh-Ritz on span(Z): G = Z^T Z, H = Z^T X (= Z^T M Z)
        for (int a = 0; a < CS_K; ++a)
        {
            double v[2 * CS_K];
#pragma unroll
            for (int i = 0; i < 2 * CS_K; ++i)
                v[i] = 0.;
            for (int r = tid; r < n; r += CS_NT)
            {
                const double za = Z[(size_t)a * n + r];
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                {
                    v[k] += za * Z[(size_t)k * n + r];
                    v[CS_K + k] += za * X[(size_t)k * n + r];
                }
            }
            block_sum<2 * CS_K>(v, red, tot);
            if (tid < CS_K)
            {
                S.Gm[a][tid] = tot[tid];
                S.Hm[a][tid] = tot[CS_K + tid];
            }
            __syncthreads();
        }
        if (tid == 0)
        {
            S.bad = 0;
            // G = R^T R (upper R stored in Gm), C = R^-T H R^-1
            double (*R)[CS_K] = S.Gm;
            for (int j = 0; j < CS_K && !S.bad; ++j)
            {
                double d = R[j][j];
                for (int k = 0; k < j; ++k)
                    d -= R[k][j] * R[k][j];
                if (!(d > 0.))
                {
                    S.bad = 1;
                    break;
                }
                d = sqrt(d);
                R[j][j] = d;
                for (int i = j + 1; i < CS_K; ++i)
                {
                    double s = R[j][i];
                    for (int k = 0; k < j; ++k)
                        s -= R[k][j] * R[k][i];
                    R[j][i] = s / d;
                }
            }
            if (!S.bad)
            {
                // Hs = symmetrised H; W = R^-T Hs (solve R^T W = Hs), C = W R^-1
                for (int i = 0; i < CS_K; ++i)
                    for (int j = 0; j < CS_K; ++j)
                        S.Cm[i][j] = 0.5 * (S.Hm[i][j] + S.Hm[j][i]);
                for (int col = 0; col < CS_K; ++col) // R^T W = Hs, column by column
                    for (int i = 0; i < CS_K; ++i)
                    {
                        double s = S.Cm[i][col];
                        for (int k = 0; k < i; ++k)
                            s -= R[k][i] * S.Cm[k][col];
                        S.Cm[i][col] = s / R[i][i];
                    }
                for (int row = 0; row < CS_K; ++row) // C R = W  =>  C = W R^-1, row by row
                    for (int j = 0; j < CS_K; ++j)
                    {
                        double s = S.Cm[row][j];
                        for (int k = 0; k < j; ++k)
                            s -= S.Cm[row][k] * R[k][j];
                        S.Cm[row][j] = s / R[j][j];
                    }
                for (int i = 0; i < CS_K; ++i)
                    for (int j = 0; j < i; ++j)
                    {
                        const double m = 0.5 * (S.Cm[i][j] + S.Cm[j][i]);
                        S.Cm[i][j] = S.Cm[j][i] = m;
                    }
                jacobi_small(S.Cm, S.V);
                for (int i = 0; i < CS_K; ++i)
                    S.perm[i] = i;
                for (int i = 1; i < CS_K; ++i) // ascending mu
                {
                    const int p = S.perm[i];
                    int j = i - 1;
                    while (j >= 0 && S.Cm[S.perm[j]][S.perm[j]] > S.Cm[p][p])
                    {
                        S.perm[j + 1] = S.perm[j];
                        --j;
                    }
                    S.perm[j + 1] = p;
                }
                // Q = R^-1 V (columns permuted)
                for (int c = 0; c < CS_K; ++c)
                {
                    const int pc = S.perm[c];
                    S.mu[c] = S.Cm[pc][pc];
                    for (int i = CS_K - 1; i >= 0; --i)
                    {
                        double s = S.V[i][pc];
                        for (int k = i + 1; k < CS_K; ++k)
                            s -= R[i][k] * S.Q[k][c];
                        S.Q[i][c] = s / R[i][i];
                    }
                }
            }
        }
        __syncthreads();
        if (S.bad)
        {
            result = -4; // Gram matrix not positive definite (should not happen)
            break;
        }
        // X <- Z Q (unit vectors), residuals || X_old q - mu Z q ||
        double rs[CS_K];
#pragma unroll
        for (int k = 0; k < CS_K; ++k)
            rs[k] = 0.;
        for (int r = tid; r < n; r += CS_NT)
        {
            double zr[CS_K], xr[CS_K];
#pragma unroll
            for (int k = 0; k < CS_K; ++k)
            {
                zr[k] = Z[(size_t)k * n + r];
                xr[k] = X[(size_t)k * n + r];
            }
#pragma unroll
            for (int c = 0; c < CS_K; ++c)
            {
                double zn = 0., xo = 0.;
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                {
                    zn += zr[k] * S.Q[k][c];
                    xo += xr[k] * S.Q[k][c];
                }
                const double d = xo - S.mu[c] * zn;
                rs[c] += d * d;
                X[(size_t)c * n + r] = zn;
            }
        }
        block_sum<CS_K>(rs, red, tot);
        if (tid == 0)
        {
            // wanted: every pair with lambda <= theta and the first one above it
            int m = 0;
            while (m < CS_K && S.mu[m] + sigma <= theta)
                ++m;
            int ok = 1;
            for (int i = 0; i < max(1, m); ++i)
                if (!(sqrt(tot[i]) <= tol))
                    ok = 0;
            // the first Ritz value above theta only has to be above it for certain: Ritz values
            // bound the eigenvalues from above and an eigenvalue lies within the residual of it
            if (m >= 1 && m < CS_K && !(S.mu[m] + sigma - sqrt(tot[m]) > theta))
                ok = 0;
            // Ritz values bound the K lowest eigenvalues from above: all of them <= theta means
            // the block is too small whatever the residuals are
            done = (m == CS_K) ? 2 : ok;
        }
        __syncthreads();
        if (done)
        {
            result = (done == 2) ? -2 : 0;
            break;
        }
    }
    if (tid == 0)
    {
        int m = 0;
        if (result == 0)
        {
            while (m < CS_K && S.mu[m] + sigma <= theta)
                ++m;
            if (m == CS_K)
                result = -2; // the block is too small for this matrix
        }
        M.info[0] = (result == 0) ? m : result;
        M.info[1] = min(its, max_its);
        for (int i = 0; i < CS_K; ++i)
            M.lam[i] = (result == 0 || result == -2) ? S.mu[i] + sigma : 0.;
    }
}

/* the accepted vectors of every matrix, un-scaled (z = D^-1/2 y, z^T D z = 1), and their
   eigenvalues into the chunk's flat arrays */
__global__ void k_cs_gather(const sa_cs_mat *mats, int nmats, const int *nev, const int64_t *eval_off,
                            const int64_t *evect_off, const double *sinv, const int *doff,
                            double *evals, double *evects, double theta, int *borderline)
{
    const int b = blockIdx.x;
    if (b >= nmats)
        return;
    const sa_cs_mat M = mats[b];
    const int n = M.n, m = nev[b];
    const double *si = sinv + doff[b];
    for (int idx = threadIdx.x; idx < m * n; idx += blockDim.x)
    {
        const int k = idx / n, r = idx - k * n;
        evects[evect_off[b] + (int64_t)k * n + r] = M.X[(size_t)k * n + r] * si[r];
    }
    if (threadIdx.x == 0)
    {
        int near = 0;
        for (int k = 0; k < m; ++k)
            evals[eval_off[b] + k] = M.lam[k];
        for (int k = 0; k < CS_K; ++k)
            if (fabs(M.lam[k] - theta) <= 1e-12)
                near = 1;
        if (near)
            atomicAdd(borderline, 1);
    }
}
} // namespace

double sa_cs_sigma(double theta) { return -std::max(0.1 * theta, 1e-6); }

void sa_cs_factor_iterate(sa_gpu_ctx *ctx, const sa_cs_mat *d_mats, int nmats, int nmax, double theta,
                          cudaStream_t st)
{
    (void)nmax;
    if (nmats <= 0)
        return;
    SpectralWs &WS = ctx->sws;
    const double sigma = sa_cs_sigma(theta);
    static const int force_g = getenv("SA_GPU_CS_GROUP") ? atoi(getenv("SA_GPU_CS_GROUP")) : 0;
    int G = 1;
    if (nmats * 2 <= ctx->num_sms)
        G = std::max(1, std::min(16, ctx->num_sms / nmats));
    if (force_g > 0)
        G = force_g;
    WS.counters.ensure((size_t)ctx->num_sms + 8);
    SA_CUDA(cudaMemsetAsync(WS.counters.p, 0, ((size_t)ctx->num_sms + 8) * sizeof(unsigned int), st));
    unsigned int *queue = WS.counters.p + ctx->num_sms + 4;
    ProfScope *pc = (st == ctx->stream) ? new ProfScope(ctx, "eig.cs_chol") : nullptr;
    if (G == 1)
    {
        const int grid = std::min(nmats, ctx->num_sms);
        k_cs_chol<false><<<grid, CS_NT, 0, st>>>(d_mats, nmats, sigma, 1, WS.counters.p, queue);
        SA_CUDA(cudaGetLastError());
    }
    else
    {
        int ngroups = std::min(nmats, ctx->num_sms / G);
        int grid = ngroups * G;
        void *args[] = {(void *)&d_mats, (void *)&nmats, (void *)&sigma, (void *)&G,
                        (void *)&WS.counters.p, (void *)&queue};
        SA_CUDA(cudaLaunchCooperativeKernel((const void *)k_cs_chol<true>, dim3(grid), dim3(CS_NT), args, 0, st));
    }
    ctx->launches++;
    delete pc;
    ProfScope pi(ctx, "eig.cs_iterate");
    static const int max_its = getenv("SA_GPU_CS_MAXIT") ? atoi(getenv("SA_GPU_CS_MAXIT")) : 120;
    k_cs_iterate<<<nmats, CS_NT, 0, st>>>(d_mats, nmats, sigma, theta, max_its, 1e-13);
    SA_CUDA(cudaGetLastError());
    ctx->launches++;
}

void sa_cs_gather(sa_gpu_ctx *ctx, const sa_cs_mat *d_mats, int nmats, const int *d_nev,
                  const int64_t *d_eval_off, const int64_t *d_evect_off, const double *d_sinv,
                  const int *d_doff, double *evals, double *evects, double theta, int *borderline,
                  cudaStream_t st)
{
    if (nmats <= 0)
        return;
    k_cs_gather<<<nmats, 256, 0, st>>>(d_mats, nmats, d_nev, d_eval_off, d_evect_off, d_sinv, d_doff,
                                        evals, evects, theta, borderline);
    SA_CUDA(cudaGetLastError());
    ctx->launches++;
}

/* Development / test entry: the lower eigenpairs of `nmats` dense symmetric n x n matrices
   (column-major, spectrum in [0, 1]) -- tests/test_cholsi.py compares with numpy.linalg.eigh. */
extern "C" int sa_gpu_debug_cholsi(sa_gpu_ctx *ctx, int nmats, int n, const double *A, double theta,
                                   int *info2, double *lam, double *X)
{
    SA_API_BEGIN
    cudaStream_t st = ctx->stream;
    DevBuf<double> T, Xd, Zd, ld;
    DevBuf<int> inf;
    const size_t nn = (size_t)n * n;
    T.upload(A, nn * nmats, st);
    Xd.alloc((size_t)nmats * n * CS_K);
    Zd.alloc((size_t)nmats * n * CS_K);
    ld.alloc((size_t)nmats * CS_K);
    inf.alloc((size_t)nmats * 2);
    inf.zero(st);
    std::vector<sa_cs_mat> hm(nmats);
    for (int b = 0; b < nmats; ++b)
    {
        hm[b].n = n;
        hm[b].slot = b;
        hm[b].T = T.p + nn * b;
        hm[b].X = Xd.p + (size_t)b * n * CS_K;
        hm[b].Z = Zd.p + (size_t)b * n * CS_K;
        hm[b].lam = ld.p + (size_t)b * CS_K;
        hm[b].info = inf.p + 2 * b;
    }
    DevBuf<int64_t> dm;
    static_assert(sizeof(sa_cs_mat) % sizeof(int64_t) == 0, "descriptor upload");
    dm.upload((const int64_t *)hm.data(), hm.size() * sizeof(sa_cs_mat) / sizeof(int64_t), st);
    sa_cs_factor_iterate(ctx, (const sa_cs_mat *)dm.p, nmats, n, theta, st);
    inf.download(info2, (size_t)nmats * 2, st);
    ld.download(lam, (size_t)nmats * CS_K, st);
    Xd.download(X, (size_t)nmats * n * CS_K, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}
