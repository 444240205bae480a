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
constexpr int CS_NB = 32;  // inner panel width
constexpr int CS_OB = 128; // outer panel width (rank of the trailing update)
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

/* one 32 x 32 tile of a trailing update: C(i0.., c0..) -= P(i0.., 0:kdim) P(c0.., 0:kdim)^T, lower
   part only (one warp; 16 DMMA accumulators of 8 x 8; kdim a multiple of 4) */
template <bool MULTI>
__device__ __noinline__ void syrk_tile(double *__restrict__ T, const double *P, size_t ld, int n,
                                       int i0, int c0, int kdim)
{
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    double c[4][4][2];
    double *Cp = T + (i0 + g) + ld * (size_t)(c0 + 2 * t);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
    {
        const int row = i0 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
        {
            const int col = c0 + ni * 8 + 2 * t;
            const double *q = Cp + mi * 8 + ld * (size_t)(ni * 8);
            c[mi][ni][0] = (row < n && col <= row) ? (MULTI ? __ldcg(q) : __ldcs(q)) : 0.;
            c[mi][ni][1] = (row < n && col + 1 <= row) ? (MULTI ? __ldcg(q + ld) : __ldcs(q + ld)) : 0.;
        }
    }
    // operand rows clamped into the matrix (rows >= n only feed accumulators that are not stored)
    const double *Ap = P + ld * (size_t)t;
    int ra[4], rb[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
    {
        ra[mi] = min(i0 + mi * 8 + g, n - 1);
        rb[mi] = min(c0 + mi * 8 + g, n - 1);
    }
#pragma unroll 2
    for (int kk = 0; kk < kdim; kk += 4)
    {
        const double *col = Ap + ld * (size_t)kk;
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
        {
            a[mi] = -ldt<MULTI>(col + ra[mi]);
            b[mi] = ldt<MULTI>(col + rb[mi]);
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
            double *q = Cp + mi * 8 + ld * (size_t)(ni * 8);
            if (row < n && col <= row)
                *q = c[mi][ni][0];
            if (row < n && col + 1 <= row)
                q[ld] = c[mi][ni][1];
        }
    }
}

/* rows below the diagonal block: X = B Linv^T (Linv = inverse of the 32 x 32 diagonal factor, in
   shared memory, zero above the diagonal) as DMMA products, one warp per 32 rows, in place: the
   warp reads its 32 x 32 piece of B completely before it writes X. */
template <bool MULTI>
__device__ __noinline__ void trsm_tiles(double *__restrict__ T, size_t ld, int n, int j0, int jb,
                                        int first_tile, int tile_stride, const double *Li)
{
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = j0 + jb;
    for (int i0 = r0 + 32 * first_tile; i0 < n; i0 += 32 * tile_stride)
    {
        double a[4][8], c[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
        {
            const int row = min(i0 + mi * 8 + g, n - 1);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
            {
                const int col = kk * 4 + t;
                a[mi][kk] = (col < jb) ? ldt<MULTI>(T + row + ld * (size_t)(j0 + col)) : 0.;
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                c[mi][ni][0] = c[mi][ni][1] = 0.;
        }
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
        {
            // B operand: (Linv^T)(k, col) = Linv(col, k)
            double bq[4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                bq[ni] = Li[(ni * 8 + g) * CS_LD + kk * 4 + t];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    dmma884(c[mi][ni][0], c[mi][ni][1], a[mi][kk], bq[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
        {
            const int row = i0 + mi * 8 + g;
            if (row < n)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    const int col = ni * 8 + 2 * t;
                    if (col < jb)
                        T[row + ld * (size_t)(j0 + col)] = c[mi][ni][0];
                    if (col + 1 < jb)
                        T[row + ld * (size_t)(j0 + col + 1)] = c[mi][ni][1];
                }
        }
    }
}

/* trailing update by the panel columns [pc0, pc0 + kdim): the tiles of the block columns
   [0, ntc) of the lower triangle that starts at row / column r0 (ntc >= nt: the whole triangle).
   The 16 warps of a block work on one SUPER-TILE of 4 x 4 tiles at a time (warp w: tile row w / 4,
   tile column w % 4), so that the block touches 4 + 4 operand strips of the panel for 16 tiles
   (L1 shares them) instead of 16 + 1 -- the panels of the ~148 matrices in flight do not fit L2,
   so every operand strip that is not shared comes from HBM (ncu: 353 GB read for 132 GB of C with
   tiles dealt one by one).  Super-tiles are dealt to the blocks of the group, column-major. */
template <bool MULTI>
__device__ __forceinline__ void syrk_region(double *__restrict__ T, size_t ld, int n, int pc0, int kdim,
                                            int r0, int ntc, int G, int gr)
{
    const int wid = threadIdx.x >> 5;
    const int nt = (n - r0 + 31) >> 5;
    ntc = min(ntc, nt);
    const int nst = (nt + 3) >> 2, nstc = (ntc + 3) >> 2;
    // column-major enumeration of the super-tiles: super column sj holds nst - sj of them
    const int nsuper = nstc * nst - nstc * (nstc - 1) / 2;
    const double *P = T + ld * (size_t)pc0;
    for (int sl = gr; sl < nsuper; sl += G)
    {
        const double bq = 2. * nst + 1.;
        int sj = (int)((bq - sqrt(fmax(0., bq * bq - 8. * (double)sl))) * 0.5);
        sj = max(0, min(sj, nstc - 1));
        while (sj > 0 && sj * nst - sj * (sj - 1) / 2 > sl)
            --sj;
        while (sj + 1 < nstc && (sj + 1) * nst - (sj + 1) * sj / 2 <= sl)
            ++sj;
        const int si = sj + (sl - (sj * nst - sj * (sj - 1) / 2));
        const int ti = 4 * si + (wid >> 2), tj = 4 * sj + (wid & 3);
        if (ti < nt && tj < ntc && ti >= tj)
            syrk_tile<MULTI>(T, P, ld, n, r0 + ti * 32, r0 + tj * 32, kdim);
    }
}

/* M = T - sigma I = L L^T in the lower triangle of T (column-major, ld = n).  Two-level blocking:
   outer panels of CS_OB = 128 columns, factored by inner panels of 32 columns (diagonal block,
   rows below, rank-32 update of the REST OF THE OUTER PANEL only); then ONE rank-128 update of
   the trailing triangle per outer panel, so that the trailing matrix crosses HBM n / 128 times
   (n^3 / 48 bytes per matrix) and every C tile that is loaded gets 128 DMMAs x 4.  Returns false
   on a non-positive pivot (decided identically by every block of the group). */
template <bool MULTI>
__device__ bool chol_one(double *__restrict__ T, int n, double sigma, int G, int gr,
                         unsigned int *counter, unsigned int &target, double *Ls, double *Li,
                         double *rdiag, int *flag)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t ld = (size_t)n;
    // shift (rows dealt to the blocks)
    for (int i = gr * CS_NT + tid; i < n; i += G * CS_NT)
        T[i + ld * i] -= sigma;
    group_barrier(counter, G, target);
    for (int J0 = 0; J0 < n; J0 += CS_OB)
    {
        const int Jend = min(n, J0 + CS_OB);
        for (int j0 = J0; j0 < Jend; j0 += CS_NB)
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
            // inverse of the diagonal factor, column by column (lane c: Linv e_c by forward
            // substitution); the solves of k_cs_iterate and the rows below only ever need L11^-1
            if (wid == 0)
            {
                double x[CS_NB];
#pragma unroll
                for (int i = 0; i < CS_NB; ++i)
                {
                    double sacc = (i == lane) ? 1. : 0.;
#pragma unroll
                    for (int q = 0; q < i; ++q)
                        sacc -= Ls[i * CS_LD + q] * x[q];
                    x[i] = (i >= lane) ? sacc * rdiag[i] : 0.;
                }
#pragma unroll
                for (int i = 0; i < CS_NB; ++i)
                    Li[i * CS_LD + lane] = x[i];
            }
            __syncthreads();
            const int r0 = j0 + jb;
            // 2. rows below the diagonal block: X = B Linv^T
            trsm_tiles<MULTI>(T, ld, n, j0, jb, gr * CS_NW + wid, G * CS_NW, Li);
            group_barrier(counter, G, target);
            // (every block of the group has read the diagonal block by now: the INVERSE of its
            // factor replaces it)
            if (gr == 0)
                for (int idx = tid; idx < CS_NB * CS_NB; idx += CS_NT)
                {
                    const int i = idx & 31, c = idx >> 5;
                    if (i < jb && c < jb && i >= c)
                        T[(j0 + i) + ld * (j0 + c)] = Li[i * CS_LD + c];
                }
            if (r0 >= Jend)
                break;
            // 3. rank-32 update of the remaining columns of the outer panel (all rows below)
            syrk_region<MULTI>(T, ld, n, j0, CS_NB, r0, (Jend - r0 + 31) >> 5, G, gr);
            group_barrier(counter, G, target);
        }
        if (Jend >= n)
            break;
        // 4. rank-128 update of the trailing triangle
        syrk_region<MULTI>(T, ld, n, J0, CS_OB, Jend, 1 << 28, G, gr);
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
    __shared__ double Li[CS_NB * CS_LD];
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
            const bool ok = chol_one<false>(M.T, M.n, sigma, 1, 0, nullptr, target, Ls, Li, rdiag, &flag);
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
            const bool ok = chol_one<true>(M.T, M.n, sigma, G, gr, counters + gi, target, Ls, Li, rdiag, &flag);
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

/* Z = M^-1 X = L^-T L^-1 X with the K right-hand sides of every row in registers: thread t owns
   the rows t, t + 512, ... (QMAX of them).  Per block of 32 columns the owning warp publishes its
   entries, warps 0..K-1 solve the 32 x 32 triangle (lanes = rows, shuffles), everyone updates the
   rows (forward) / columns (backward) it owns: L is read exactly once per solve, the vectors
   never leave the register file in between. */
/* generalised K x K eigenproblem H q = mu G q of the Rayleigh-Ritz step (one thread): on exit
   S.mu ascending, S.Q the G-orthonormal eigenvectors; S.bad when G is not positive definite */
__device__ void cs_small_solve(CsSmall &S)
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

/* 32 x 32 diagonal block at (jb, jb) -- the inverse of the diagonal factor, written by k_cs_chol --
   into shared memory (zero above the diagonal, identity beyond the matrix) */
__device__ __forceinline__ void cs_stage_diag(const double *__restrict__ T, size_t ld, int n, int jb,
                                              double *Ls, double *rinv)
{
    const int w = min(CS_NB, n - jb);
    for (int idx = threadIdx.x; idx < CS_NB * CS_NB; idx += CS_NT)
    {
        const int i = idx & 31, c = idx >> 5;
        double v = (i == c) ? 1. : 0.;
        if (i < w && c < w && i >= c)
            v = T[(jb + i) + ld * (jb + c)];
        Ls[i * CS_LD + c] = v;
        if (i == c)
            rinv[i] = 1. / v;
    }
}

/* Subspace iteration with M^-1 = L^-T L^-1 for one matrix by a group of G = K / KL blocks:
   block gr solves for the KL vectors gr KL .. gr KL + KL - 1 (the triangular solves are independent
   per right-hand side), forms its rows of G = Z^T Z and H = Z^T X, and after a group barrier every
   block solves the same K x K problem and rotates its own vectors (ping-pong buffers X / X2: the
   other blocks still read the old ones).  Three group barriers per iteration; G = 1 is the
   one-block-per-matrix form used when there are enough matrices to fill the GPU. */
template <int KL, bool MULTI>
__global__ void __launch_bounds__(CS_NT, 1)
k_cs_iterate(const sa_cs_mat *mats, int nmats, double sigma, double theta, int max_its, double tol,
             unsigned int *counters)
{
    constexpr int G = CS_K / KL;
    __shared__ double Lsb[2][CS_NB * CS_LD];
    __shared__ double rinvb[2][CS_NB];
    __shared__ double Ys[CS_NB][CS_K]; // solved block, columns >= KL stay zero
    __shared__ double red[CS_NW * 2 * CS_K];
    __shared__ double tot[2 * CS_K];
    __shared__ CsSmall S;
    __shared__ int done;
    const int b = blockIdx.x / G, gr = blockIdx.x % G;
    if (b >= nmats)
        return;
    const sa_cs_mat M = mats[b];
    const int n = M.n;
    const size_t ld = (size_t)n;
    const double *__restrict__ T = M.T;
    double *Xc = M.X, *Xn = M.X2, *Z = M.Z;
    double *small = M.small; // [0,64) G, [64,128) H, [128,136) residuals^2
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k0 = gr * KL;
    unsigned int target = 0;
    unsigned int *counter = counters + b;
    if (M.info[0] != 0)
        return; // factorisation failed (every block of the group sees the same word)
    for (int idx = tid; idx < CS_NB * CS_K; idx += CS_NT)
        Ys[idx / CS_K][idx % CS_K] = 0.;
    for (int r = tid; r < n; r += CS_NT)
#pragma unroll
        for (int k = 0; k < KL; ++k)
            Xc[(size_t)(k0 + k) * n + r] = cs_start_value(r, k0 + k, M.slot);
    group_barrier(counter, G, target);
    int its = 0, result = -3; // -3: no convergence
    int buf = 0;
    for (its = 1; its <= max_its; ++its)
    {
        // own vectors: Z = X, then L Y = Z and L^T Z = Y in place
        for (int r = tid; r < n; r += CS_NT)
#pragma unroll
            for (int k = 0; k < KL; ++k)
                Z[(size_t)(k0 + k) * n + r] = Xc[(size_t)(k0 + k) * n + r];
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass)
        {
            const int nblk = (n + CS_NB - 1) / CS_NB;
            for (int bi_ = 0; bi_ < nblk; ++bi_)
            {
                const int jb = (pass == 0 ? bi_ : nblk - 1 - bi_) * CS_NB;
                const int w = min(CS_NB, n - jb);
                // the diagonal block of this step was staged during the previous step's update
                // (double buffer); the first one of a pass is staged here
                if (bi_ == 0)
                {
                    cs_stage_diag(T, ld, n, jb, Lsb[buf], rinvb[buf]);
                    __syncthreads();
                }
                double *Ls = Lsb[buf], *rinv = rinvb[buf];
                if (wid < KL)
                {
                    // the diagonal block stores the INVERSE of its factor: y = Linv b (forward),
                    // z = Linv^T y (backward) -- 32 independent products per lane, no substitution chain
                    const double bl = (lane < w) ? Z[(size_t)(k0 + wid) * n + jb + lane] : 0.;
                    double bi = 0.;
                    if (pass == 0)
                    {
#pragma unroll 8
                        for (int c = 0; c < CS_NB; ++c)
                            bi += Ls[lane * CS_LD + c] * __shfl_sync(0xffffffffu, bl, c);
                    }
                    else
                    {
#pragma unroll 8
                        for (int c = 0; c < CS_NB; ++c)
                            bi += Ls[c * CS_LD + lane] * __shfl_sync(0xffffffffu, bl, c);
                    }
                    Ys[lane][wid] = bi;
                    if (lane < w)
                        Z[(size_t)(k0 + wid) * n + jb + lane] = bi;
                }
                __syncthreads();
                // forward: the rows below the block; backward: the columns left of it --
                // Z(i, :) -= sum_q L(i, q) Y(q, :) as DMMA m8n8k4 (M = 32 rows / columns per warp
                // pass, N = the right-hand sides padded to 8, K = the 32 entries of the block);
                // forward L(i, q) = T[i + ld (jb + q)], backward L(i, q) = T[jb + q + ld i].  All 32
                // operand loads of a lane are in flight before the first DMMA.
                {
                    if (bi_ + 1 < nblk)
                        cs_stage_diag(T, ld, n, (pass == 0 ? bi_ + 1 : nblk - 2 - bi_) * CS_NB, Lsb[buf ^ 1],
                                      rinvb[buf ^ 1]);
                    const int lo = (pass == 0) ? jb + CS_NB : 0, hi = (pass == 0) ? n : jb;
                    const int g = lane >> 2, t = lane & 3;
                    const double *base = (pass == 0) ? T + ld * (size_t)jb : T + jb;
                    const size_t si = (pass == 0) ? 1 : ld, sq = (pass == 0) ? ld : 1;
                    for (int i0 = lo + wid * 32; i0 < hi; i0 += CS_NW * 32)
                    {
                        double a[4][8], c[4][2];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                        {
                            const int ri = min(i0 + mi * 8 + g, hi - 1);
                            const double *row = base + si * (size_t)ri + sq * (size_t)t;
#pragma unroll
                            for (int kk = 0; kk < 8; ++kk)
                                a[mi][kk] = row[sq * (size_t)(kk * 4)];
                            c[mi][0] = (2 * t < KL) ? Z[(size_t)(k0 + 2 * t) * n + ri] : 0.;
                            c[mi][1] = (2 * t + 1 < KL) ? Z[(size_t)(k0 + 2 * t + 1) * n + ri] : 0.;
                        }
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                        {
                            const double bq = Ys[kk * 4 + t][g];
#pragma unroll
                            for (int mi = 0; mi < 4; ++mi)
                                dmma884(c[mi][0], c[mi][1], -a[mi][kk], bq);
                        }
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                        {
                            const int ri = i0 + mi * 8 + g;
                            if (ri < hi)
                            {
                                if (2 * t < KL)
                                    Z[(size_t)(k0 + 2 * t) * n + ri] = c[mi][0];
                                if (2 * t + 1 < KL)
                                    Z[(size_t)(k0 + 2 * t + 1) * n + ri] = c[mi][1];
                            }
                        }
                    }
                }
                __syncthreads();
                buf ^= 1;
            }
        }
        group_barrier(counter, G, target);
        // Rayleigh-Ritz on span(Z): rows k0 .. k0 + KL - 1 of G = Z^T Z and H = Z^T X (= Z^T M Z)
        for (int a = k0; a < k0 + KL; ++a)
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
                    v[k] += za * ldt<MULTI>(Z + (size_t)k * n + r);
                    v[CS_K + k] += za * ldt<MULTI>(Xc + (size_t)k * n + r);
                }
            }
            block_sum<2 * CS_K>(v, red, tot);
            if (tid < CS_K)
            {
                small[a * CS_K + tid] = tot[tid];
                small[64 + a * CS_K + tid] = tot[CS_K + tid];
            }
            __syncthreads();
        }
        group_barrier(counter, G, target);
        if (tid < 64)
        {
            S.Gm[tid >> 3][tid & 7] = ldt<MULTI>(small + tid);
            S.Hm[tid >> 3][tid & 7] = ldt<MULTI>(small + 64 + tid);
        }
        __syncthreads();
        if (tid == 0)
            cs_small_solve(S);
        __syncthreads();
        if (S.bad)
        {
            result = -4; // Gram matrix not positive definite (should not happen)
            break;
        }
        // own vectors: X_new = Z q (unit vectors), residuals || X_old q - mu Z q ||
        double rs[KL];
#pragma unroll
        for (int k = 0; k < KL; ++k)
            rs[k] = 0.;
        for (int r = tid; r < n; r += CS_NT)
        {
            double zr[CS_K], xr[CS_K];
#pragma unroll
            for (int k = 0; k < CS_K; ++k)
            {
                zr[k] = ldt<MULTI>(Z + (size_t)k * n + r);
                xr[k] = ldt<MULTI>(Xc + (size_t)k * n + r);
            }
#pragma unroll
            for (int c = 0; c < KL; ++c)
            {
                double zn = 0., xo = 0.;
#pragma unroll
                for (int k = 0; k < CS_K; ++k)
                {
                    zn += zr[k] * S.Q[k][k0 + c];
                    xo += xr[k] * S.Q[k][k0 + c];
                }
                const double d = xo - S.mu[k0 + c] * zn;
                rs[c] += d * d;
                Xn[(size_t)(k0 + c) * n + r] = zn;
            }
        }
        block_sum<KL>(rs, red, tot);
        if (tid < KL)
            small[128 + k0 + tid] = tot[tid];
        group_barrier(counter, G, target);
        {
            double *tmp = Xc;
            Xc = Xn;
            Xn = tmp;
        }
        if (tid == 0)
        {
            double res2[CS_K];
            for (int i = 0; i < CS_K; ++i)
                res2[i] = ldt<MULTI>(small + 128 + i);
            // wanted: every pair with lambda <= theta ...
            int m = 0;
            while (m < CS_K && S.mu[m] + sigma <= theta)
                ++m;
            int ok = 1;
            for (int i = 0; i < max(1, m); ++i)
                if (!(sqrt(res2[i]) <= tol))
                    ok = 0;
            // ... and the first Ritz value above theta only has to be above it for certain: Ritz
            // values bound the eigenvalues from above and an eigenvalue lies within the residual
            if (m >= 1 && m < CS_K && !(S.mu[m] + sigma - sqrt(res2[m]) > theta))
                ok = 0;
            // all K Ritz values <= theta: the block is too small whatever the residuals are
            done = (m == CS_K) ? 2 : ok;
        }
        __syncthreads();
        if (done)
        {
            result = (done == 2) ? -2 : 0;
            break;
        }
        // (the next iteration's first group barrier separates these reads of `small` from the
        // writes of the next Rayleigh-Ritz step)
    }
    // the final vectors live in Xc; the caller reads M.X
    if (result == 0 && Xc != M.X)
    {
        for (int r = tid; r < n; r += CS_NT)
#pragma unroll
            for (int k = 0; k < KL; ++k)
                M.X[(size_t)(k0 + k) * n + r] = Xc[(size_t)(k0 + k) * n + r];
    }
    if (tid == 0 && gr == 0)
    {
        int m = 0;
        if (result == 0)
            while (m < CS_K && S.mu[m] + sigma <= theta)
                ++m;
        M.info[1] = min(its, max_its);
        for (int i = 0; i < CS_K; ++i)
            M.lam[i] = (result == 0 || result == -2) ? S.mu[i] + sigma : 0.;
        __threadfence();
        M.info[0] = (result == 0) ? m : result;
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
    static int max_its = getenv("SA_GPU_CS_MAXIT") ? atoi(getenv("SA_GPU_CS_MAXIT")) : 120;
    // few matrices: G blocks per matrix, each with K / G of the vectors (co-resident: cooperative)
    static const int force_gi = getenv("SA_GPU_CS_ITER_GROUP") ? atoi(getenv("SA_GPU_CS_ITER_GROUP")) : 0;
    int Gi = 1;
    while (Gi < CS_K && nmats * Gi * 2 <= ctx->num_sms)
        Gi *= 2;
    if (force_gi > 0)
        Gi = force_gi;
    const double tol = 1e-13;
    unsigned int *cnt2 = WS.counters.p; // (k_cs_chol has finished with them: same stream)
    SA_CUDA(cudaMemsetAsync(cnt2, 0, ((size_t)ctx->num_sms + 8) * sizeof(unsigned int), st));
    if (Gi == 1)
        k_cs_iterate<CS_K, false><<<nmats, CS_NT, 0, st>>>(d_mats, nmats, sigma, theta, max_its, tol, cnt2);
    else
    {
        if (nmats > ctx->num_sms + 8)
            SA_FAIL("sa_cs_factor_iterate: group counters");
        int grid = nmats * Gi;
        void *args[] = {(void *)&d_mats, (void *)&nmats, (void *)&sigma, (void *)&theta,
                        (void *)&max_its, (void *)&tol, (void *)&cnt2};
        const void *fn = (Gi == 2)   ? (const void *)k_cs_iterate<4, true>
                         : (Gi == 4) ? (const void *)k_cs_iterate<2, true>
                                     : (const void *)k_cs_iterate<1, true>;
        SA_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(CS_NT), args, 0, st));
    }
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
    DevBuf<double> T, Xd, X2d, Zd, ld, sm;
    DevBuf<int> inf;
    const size_t nn = (size_t)n * n;
    T.upload(A, nn * nmats, st);
    Xd.alloc((size_t)nmats * n * CS_K);
    Zd.alloc((size_t)nmats * n * CS_K);
    X2d.alloc((size_t)nmats * n * CS_K);
    sm.alloc((size_t)nmats * 160);
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
        hm[b].X2 = X2d.p + (size_t)b * n * CS_K;
        hm[b].small = sm.p + (size_t)b * 160;
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
