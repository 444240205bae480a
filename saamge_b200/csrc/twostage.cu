// Two-stage tridiagonalisation of LARGE AE matrices (n above the shared-memory kernel's
// limit: coarse levels, n ~ 10^3, and order-2 fine levels, n ~ 700) -- the part of
// xpacks_calc_lower_eigens_dense (amg/src/xpacks.cpp:222-314) that LAPACK's dsygvx spends in
// dsytrd, reorganised so that (almost) all flops are FP64 tensor-core GEMMs:
//
//   stage 1, k_sy2sb: dense -> band of width TS_B = 32.  One thread block per matrix (work
//     queue, largest first).  Per panel of 32 columns: Householder QR of the block below the
//     band (sub-panels of 8 columns in shared memory, block reflector applied to the rest of
//     the panel), compact WY factor Tf from the Gram matrix V^T V (DMMA), then the two-sided
//     update of the trailing block  A22 <- (I - V Tf V^T)^T A22 (I - V Tf V^T):
//        W = A22 V            symmetric matrix x 32 columns, DMMA m8n8k4, A fragments loaded
//                             straight from the lower triangle (max/min indexing), V through
//                             shared memory
//        X = W Tf, S = V^T X (DMMA), Z = X - 1/2 V Tf^T S
//        A22 -= Z V^T + V Z^T rank-64 update of the lower triangle, DMMA
//     HBM traffic per matrix: 16 r^2 bytes per panel => 16 n^3 / (3*32) bytes, i.e. n/6 passes
//     over the matrix instead of the n passes of an unblocked reduction.
//   stage 2, k_sb2st: band -> tridiagonal by bulge chasing (Householder, 32 x 32 blocks).  One
//     block per matrix, one warp per sweep, sweeps pipelined two steps apart, lanes = rows.
//   back-transformation, k_ts_back: z = Q1 Q2 y, one block per eigenvector, vector in shared
//     memory: stage-2 reflectors sweep by sweep in reverse, then the stage-1 reflectors.
//
// Storage: T (n x n, column-major) holds on exit the stage-1 reflectors below the band
// (reflector of column k: unit entry in row k+32, tail below) and the stage-2 reflectors in
// its UPPER triangle (sweep s, row i: T[s + n i]; the first entry of every reflector holds
// tau).  tests/twostage_ref.py is the numpy statement of exactly this algorithm and layout.
#include <algorithm>

#include <cooperative_groups.h>

#include "sa_gpu_internal.cuh"

namespace cg = cooperative_groups;

namespace
{

constexpr int TS_B = 32;    // band width = panel width
constexpr int TS_NT = 512;  // threads of the stage-1 block
constexpr int TS_NW = 16;   // warps of the stage-1 block
constexpr int TS_LD = 36;   // leading dimension of the 32-wide shared operand tiles (== 4 mod 16)
constexpr int TS_LDB = 64;  // rows of the band storage (band + bulge)
constexpr int TS_S2_NW = 8; // warps (concurrent sweeps) of the stage-2 block
constexpr int TS_S2_PER_WARP = 32 * 33 + 32 * 17 + 96; // doubles of shared memory per stage-2 warp

// cycle counters of the phases of the two kernels summed over blocks (diagnostics):
// 0 panel QR, 1 Gram + Tf, 2 symm, 3 X = W Tf, 4 S + M, 5 Z, 6 syr2k, 7 band copy,
// 8 stage-2 ticks, 9 stage-2 cycles, 10 stage-2 steps
__device__ unsigned long long g_ts_clk[16];
#define TS_CLK(idx)                                                            \
    do {                                                                       \
        if (threadIdx.x == 0)                                                  \
        {                                                                      \
            const long long now__ = clock64();                                 \
            atomicAdd(&g_ts_clk[idx], (unsigned long long)(now__ - tclk));     \
            tclk = now__;                                                      \
        }                                                                      \
    } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* Sums 32 values per lane over the 32 lanes of a warp with 31 shuffles: lane l returns the
   total of v[l]. */
__device__ __forceinline__ double warp_transpose_reduce32(double (&v)[32])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        const bool up = lane & 16;
        const double send = up ? v[i] : v[i + 16];
        const double keep = up ? v[i + 16] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const bool up = lane & 8;
        const double send = up ? v[i] : v[i + 8];
        const double keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const bool up = lane & 4;
        const double send = up ? v[i] : v[i + 4];
        const double keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
    {
        const bool up = lane & 2;
        const double send = up ? v[i] : v[i + 2];
        const double keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const bool up = lane & 1;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    return v[0];
}

/* Loads of data another thread block of the cluster may have written (cluster mode: a matrix is
   shared by the blocks of a thread-block cluster): L2 only, never a stale L1 line. */
template <bool CL> __device__ __forceinline__ double ldx(const double *p)
{
    if (CL)
        return __ldcg(p);
    return *p;
}

/* which block of the cluster owns pass q (256 rows) of the row-distributed phases */
__device__ __forceinline__ bool s1_owns(int q, int rank, int CS) { return CS == 1 || (q % CS) == rank; }
/* ... and of the triangular update (cost grows with q): dealt in snake order */
__device__ __forceinline__ bool s1_owns_tri(int q, int rank, int CS)
{
    if (CS == 1)
        return true;
    const int round = q / CS, pos = q % CS;
    return ((round & 1) ? CS - 1 - pos : pos) == rank;
}

/* dlarfg on (alpha, |x|^2): beta, tau and the scale 1 / (alpha - beta) of the tail */
__device__ __forceinline__ void ts_larfg(double alpha, double xn2, double &beta, double &tau,
                                         double &scal)
{
    beta = alpha;
    tau = 0.;
    scal = 0.;
    if (xn2 > 0.)
    {
        beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        tau = (beta - alpha) / beta;
        scal = 1. / (alpha - beta);
    }
}

// ------------------------------------------------------------------------------- stage 1

/* shared-memory carve-up of k_sy2sb (doubles) */
struct S1Smem
{
    double *red;  // TS_NW * 32 reduction scratch
    double *sval; // 64 reduced values / scalars
    double *Tf;   // 32 x 33 compact WY factor of the panel (upper triangular)
    double *G;    // 32 x 33 Gram / S / M
    double *Tin;  // 8 x 9 factor of the current sub-panel
    double *Gin;  // 8 x 9
    double *taus; // 32
    double *Y;    // 8 x 4 (+ pad)
    double *U;    // big region: sub-panel / operand tiles / Gram partials
};

__device__ __forceinline__ S1Smem s1_carve(double *sm)
{
    S1Smem S;
    S.red = sm;
    S.sval = S.red + TS_NW * 32;
    S.Tf = S.sval + 64;
    S.G = S.Tf + 32 * 33;
    S.Tin = S.G + 32 * 33;
    S.Gin = S.Tin + 72;
    S.taus = S.Gin + 72;
    S.Y = S.taus + 32;
    S.U = S.Y + 64;
    return S;
}
constexpr int TS_S1_FIXED = TS_NW * 32 + 64 + 2 * 32 * 33 + 72 + 72 + 32 + 64; // doubles before U
constexpr int TS_S1_UMIN = 2 * 64 * 36 + 16 * 64 * 16;                         // syr2k operand tiles (>= Gram partials)

/* Sums W (<= 8) values per thread over the block; totals land in S.sval[0..W-1] (visible to all
   threads on return).  Deterministic: warp shuffles, then warps in order. */
template <int W>
__device__ __forceinline__ void s1_block_reduce(const S1Smem S, double (&v)[W])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < W; ++q)
        v[q] = warp_sum(v[q]);
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < W; ++q)
            S.red[wid * 8 + q] = v[q];
    __syncthreads();
    if (threadIdx.x < W)
    {
        double s = 0.;
        for (int w = 0; w < TS_NW; ++w)
            s += S.red[w * 8 + threadIdx.x];
        S.sval[threadIdx.x] = s;
    }
    __syncthreads();
}

/* out (32 x 33 in shared memory) = A^T B over r rows; A, B: r x 32 column-major with leading
   dimension ld in global memory.  DMMA, rows dealt to the warps in slabs of 32, partial
   products reduced through S.U in two rounds of 8 warps (fixed order). */
template <bool CL>
__device__ __noinline__ void s1_gram(const S1Smem S, const double *__restrict__ A, const double *__restrict__ B,
                                     int ld, int r, double *out)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            acc[mi][ni][0] = acc[mi][ni][1] = 0.;
    for (int slab = wid; slab * 32 < r; slab += TS_NW)
    {
#pragma unroll 2
        for (int kk = 0; kk < 8; ++kk)
        {
            const int row = slab * 32 + kk * 4 + t;
            const bool ok = row < r;
            double a[4], b[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                a[q] = ok ? ldx<CL>(A + row + (size_t)ld * (q * 8 + g)) : 0.;
                b[q] = ok ? ldx<CL>(B + row + (size_t)ld * (q * 8 + g)) : 0.;
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    for (int round = 0; round < 2; ++round)
    {
        __syncthreads();
        if ((wid >> 3) == round)
        {
            double *u = S.U + (size_t)(wid & 7) * 1024;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    u[(mi * 8 + g) * 32 + ni * 8 + 2 * t] = acc[mi][ni][0];
                    u[(mi * 8 + g) * 32 + ni * 8 + 2 * t + 1] = acc[mi][ni][1];
                }
        }
        __syncthreads();
        for (int e = tid; e < 1024; e += TS_NT)
        {
            double s = round ? out[(e >> 5) * 33 + (e & 31)] : 0.;
            for (int w = 0; w < 8; ++w)
                s += S.U[(size_t)w * 1024 + e];
            out[(e >> 5) * 33 + (e & 31)] = s;
        }
    }
    __syncthreads();
}

/* Householder QR of the r x 32 panel P (leading dimension ld, global memory), nr reflectors.
   On exit: P holds R on / above its diagonal and the reflector tails below; Vc (r x 32, leading
   dimension ldv) holds the reflectors with explicit unit diagonal and zeros above (columns
   >= nr are zero); S.taus / tau_out hold tau. */
template <bool CL>
__device__ __noinline__ void s1_panel_qr(const S1Smem S, double *__restrict__ P, int ld, int r, int nr, int w_in,
                                         double *__restrict__ Vc, int ldv, double *__restrict__ tau_out)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *Ps = S.U;
    for (int c0 = 0; c0 < TS_B; c0 += w_in)
    {
        const int w = w_in; // TS_B is a multiple of every w_in
        const int rows = r - c0;
        if (c0 >= nr || rows <= 0)
        {
            // no reflectors left: zero columns of V
            for (int idx = tid; idx < (TS_B - c0) * r; idx += TS_NT)
                Vc[(idx % r) + (size_t)ldv * (c0 + idx / r)] = 0.;
            if (tid < TS_B - c0)
            {
                S.taus[c0 + tid] = 0.;
                tau_out[c0 + tid] = 0.;
            }
            break;
        }
        const int ldp = rows | 1; // odd: columns land in different banks
        for (int idx = tid; idx < rows * w; idx += TS_NT)
        {
            const int cc = idx / rows, i = idx - cc * rows;
            Ps[cc * ldp + i] = ldx<CL>(P + (c0 + i) + (size_t)ld * (c0 + cc));
        }
        if (tid < 72)
        {
            S.Tin[tid] = 0.;
            S.Gin[tid] = 0.;
        }
        __syncthreads();
        for (int cc = 0; cc < w; ++cc)
        {
            const int c = c0 + cc;
            if (c >= nr)
            {
                if (tid == 0)
                {
                    S.taus[c] = 0.;
                    tau_out[c] = 0.;
                }
                continue; // (uniform)
            }
            // one pass: |x|^2 and the products of x with every other column of the sub-panel
            double acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                acc[q] = 0.;
            for (int i = cc + 1 + tid; i < rows; i += TS_NT)
            {
                const double x = Ps[cc * ldp + i];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < w)
                        acc[q] += x * Ps[q * ldp + i];
            }
            s1_block_reduce<8>(S, acc);
            const double alpha = Ps[cc * ldp + cc];
            double beta, tau, scal;
            ts_larfg(alpha, S.sval[cc], beta, tau, scal);
            // s_q = tau (P[cc][q] + scal * x.P[:, q]) for the columns to the right
            double sq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                sq[q] = (q > cc && q < w) ? tau * (Ps[q * ldp + cc] + scal * S.sval[q]) : 0.;
            // Gram entries with the earlier reflectors of the sub-panel (for Tin)
            double gq = 0.;
            if (tid < cc)
                gq = Ps[tid * ldp + cc] + scal * S.sval[tid];
            __syncthreads();
            for (int i = cc + 1 + tid; i < rows; i += TS_NT)
            {
                const double v = Ps[cc * ldp + i] * scal;
                Ps[cc * ldp + i] = v;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q > cc && q < w)
                        Ps[q * ldp + i] -= sq[q] * v;
            }
            if (tid > cc && tid < w)
                Ps[tid * ldp + cc] -= tau * (Ps[tid * ldp + cc] + scal * S.sval[tid]);
            if (tid < cc)
                S.Gin[tid * 9 + cc] = gq;
            if (tid == 0)
            {
                Ps[cc * ldp + cc] = beta;
                S.taus[c] = tau;
                tau_out[c] = tau;
            }
            __syncthreads();
        }
        // dlarft of the sub-panel: Tin[:cc, cc] = -tau Tin[:cc, :cc] Gin[:cc, cc]
        if (tid == 0)
        {
            for (int cc = 0; cc < w; ++cc)
            {
                const double tau = S.taus[c0 + cc];
                for (int p = 0; p < cc; ++p)
                {
                    double s = 0.;
                    for (int q = p; q < cc; ++q)
                        s += S.Tin[p * 9 + q] * S.Gin[q * 9 + cc];
                    S.Tin[p * 9 + cc] = -tau * s;
                }
                S.Tin[cc * 9 + cc] = tau;
            }
        }
        // write the sub-panel back and its clean reflector columns to Vc
        for (int idx = tid; idx < rows * w; idx += TS_NT)
        {
            const int cc = idx / rows, i = idx - cc * rows;
            const double val = Ps[cc * ldp + i];
            P[(c0 + i) + (size_t)ld * (c0 + cc)] = val;
            const bool live = (c0 + cc) < nr;
            Vc[(c0 + i) + (size_t)ldv * (c0 + cc)] = !live ? 0. : (i < cc ? 0. : (i == cc ? 1. : val));
        }
        for (int idx = tid; idx < c0 * w; idx += TS_NT)
            Vc[(idx % c0) + (size_t)ldv * (c0 + idx / c0)] = 0.;
        __syncthreads();
        // block reflector on the rest of the panel, four columns at a time:
        //   P_rest -= V Tin^T (V^T P_rest)
        for (int q0 = c0 + w; q0 < TS_B; q0 += 4)
        {
            double part[32];
#pragma unroll
            for (int q = 0; q < 32; ++q)
                part[q] = 0.;
            for (int i = tid; i < rows; i += TS_NT)
            {
                double p[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    p[j] = ldx<CL>(P + (c0 + i) + (size_t)ld * (q0 + j));
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                    if (cc < w)
                    {
                        const double v = (i < cc) ? 0. : (i == cc ? 1. : Ps[cc * ldp + i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            part[cc * 4 + j] += v * p[j];
                    }
            }
            const double tot = warp_transpose_reduce32(part);
            S.red[wid * 32 + lane] = tot;
            __syncthreads();
            if (tid < 32)
            {
                double s = 0.;
                for (int ww = 0; ww < TS_NW; ++ww)
                    s += S.red[ww * 32 + tid];
                S.sval[tid] = s; // Wt[cc][j] at cc * 4 + j
            }
            __syncthreads();
            if (tid < 32)
            {
                const int cc = tid >> 2, j = tid & 3;
                double s = 0.;
                if (cc < w && (c0 + cc) < nr)
                    for (int c2 = 0; c2 <= cc; ++c2)
                        s += S.Tin[c2 * 9 + cc] * S.sval[c2 * 4 + j];
                S.Y[tid] = s;
            }
            __syncthreads();
            for (int i = tid; i < rows; i += TS_NT)
            {
                double d[4] = {0., 0., 0., 0.};
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                    if (cc < w)
                    {
                        const double v = (i < cc) ? 0. : (i == cc ? 1. : Ps[cc * ldp + i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            d[j] += v * S.Y[cc * 4 + j];
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    P[(c0 + i) + (size_t)ld * (q0 + j)] = ldx<CL>(P + (c0 + i) + (size_t)ld * (q0 + j)) - d[j];
            }
            __syncthreads();
        }
        __syncthreads();
    }
    __syncthreads();
}

/* W = A22 V.  A22: r x r symmetric, LOWER triangle valid, leading dimension ld; V, W: r x 32,
   leading dimension ldv.  Every warp owns a strip of 16 rows per pass (256 rows per pass);
   A fragments come straight from global memory (element (i, k) lives at max(i,k) + ld min(i,k);
   L2 only, they are used once), prefetched one 32-column chunk ahead into a SECOND register set
   (the chunk loop is unrolled by two: ncu showed the loads of a single set waiting for the DMMAs
   that still read it); the V chunk is shared through S.U: fetched into registers before the MMAs
   of the current chunk, parked in shared memory after them (double buffered, one barrier per
   chunk).  (Measured alternative, slower by 25 %: no barriers, V fragments re-read through L1.) */
template <bool CL>
__device__ __noinline__ void s1_symm(const S1Smem S, const double *__restrict__ A2, int ld, int r,
                                     const double *__restrict__ V, double *__restrict__ W, int ldv, int rank,
                                     int CS)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *Vs = S.U; // 2 x [32 columns][TS_LD]
    for (int pass0 = 0; pass0 < r; pass0 += TS_NW * 16)
    {
        if (!s1_owns(pass0 / (TS_NW * 16), rank, CS))
            continue;
        const int i0 = pass0 + wid * 16;
        const bool active = i0 < r;
        const int ra = i0 + g, rb = i0 + 8 + g;
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
                acc[mi][ni][0] = acc[mi][ni][1] = 0.;
        double A0[2][8], A1[2][8];
        auto load_a = [&](double (&A)[2][8], int k0) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
            {
                const int k = k0 + kk * 4 + t;
                const bool kok = active && k < r;
                const int lo_a = min(ra, k), hi_a = max(ra, k);
                const int lo_b = min(rb, k), hi_b = max(rb, k);
                A[0][kk] = (kok && ra < r) ? __ldcg(A2 + hi_a + (size_t)ld * lo_a) : 0.;
                A[1][kk] = (kok && rb < r) ? __ldcg(A2 + hi_b + (size_t)ld * lo_b) : 0.;
            }
        };
        const int vc0 = tid >> 5, vk = tid & 31; // entries (vc0, vk) and (vc0 + 16, vk) of a chunk
        double vreg[2];
        auto fetch_v = [&](int k0) {
            const bool ok = k0 + vk < r;
            vreg[0] = ok ? ldx<CL>(V + (k0 + vk) + (size_t)ldv * vc0) : 0.;
            vreg[1] = ok ? ldx<CL>(V + (k0 + vk) + (size_t)ldv * (vc0 + 16)) : 0.;
        };
        auto park_v = [&](int buf) {
            Vs[buf * 32 * TS_LD + vc0 * TS_LD + vk] = vreg[0];
            Vs[buf * 32 * TS_LD + (vc0 + 16) * TS_LD + vk] = vreg[1];
        };
        auto compute = [&](const double (&A)[2][8], int buf) {
            if (!active)
                return;
            const double *vb = Vs + buf * 32 * TS_LD;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
            {
                double b[4];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    b[ni] = vb[(ni * 8 + g) * TS_LD + kk * 4 + t];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    dmma884(acc[0][ni][0], acc[0][ni][1], A[0][kk], b[ni]);
                    dmma884(acc[1][ni][0], acc[1][ni][1], A[1][kk], b[ni]);
                }
            }
        };
        load_a(A0, 0);
        fetch_v(0);
        park_v(0);
        for (int k0 = 0; k0 < r; k0 += 64)
        {
            __syncthreads(); // Vs[0] is complete; everyone is done with Vs[1]
            const bool more1 = k0 + 32 < r;
            if (more1)
            {
                load_a(A1, k0 + 32);
                fetch_v(k0 + 32);
            }
            compute(A0, 0);
            if (!more1)
                break;
            park_v(1);
            __syncthreads(); // Vs[1] is complete; everyone is done with Vs[0]
            const bool more2 = k0 + 64 < r;
            if (more2)
            {
                load_a(A0, k0 + 64);
                fetch_v(k0 + 64);
            }
            compute(A1, 1);
            if (more2)
                park_v(0);
        }
        if (active)
        {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
            {
                const int c = ni * 8 + 2 * t;
                if (ra < r)
                {
                    W[ra + (size_t)ldv * c] = acc[0][ni][0];
                    W[ra + (size_t)ldv * (c + 1)] = acc[0][ni][1];
                }
                if (rb < r)
                {
                    W[rb + (size_t)ldv * c] = acc[1][ni][0];
                    W[rb + (size_t)ldv * (c + 1)] = acc[1][ni][1];
                }
            }
        }
        __syncthreads();
    }
}

/* A22 -= Z V^T + V Z^T on the lower triangle (rank-64 update), DMMA.  A warp owns a strip of 16
   rows per pass; its [-Z | -V] operand (16 x 64) sits in a per-warp, XOR-swizzled shared tile
   (registers spilled when it lived there: ncu), the [V | Z]^T operand of a block of 32 columns is
   shared by all warps through S.U (registers -> shared memory, double buffered); the C fragments
   of the NEXT column block are prefetched (L2 only) while the current one is updated. */
template <bool CL>
__device__ __noinline__ void s1_syr2k(const S1Smem S, double *__restrict__ A2, int ld, int r,
                                      const double *__restrict__ V, const double *__restrict__ Z, int ldv,
                                      int rank, int CS)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    double *Bs = S.U;                                         // 2 x [64 k][TS_LD]
    double *As = S.U + 2 * 64 * TS_LD + (size_t)wid * 64 * 16; // per warp: [64 k][16 rows], swizzled
    for (int pass0 = 0; pass0 < r; pass0 += TS_NW * 16)
    {
        if (!s1_owns_tri(pass0 / (TS_NW * 16), rank, CS))
            continue;
        const int i0 = pass0 + wid * 16;
        const bool active = i0 < r;
        const int ra = i0 + g, rb = i0 + 8 + g;
        // strip operand: element (k, row) at k * 16 + (row ^ ((k & 3) << 2)); k < 32: -Z, else -V
        __syncwarp();
        for (int idx = lane; idx < 64 * 16; idx += 32)
        {
            const int k = idx >> 4, row = idx & 15;
            const double *src = (k < 32) ? Z + (size_t)ldv * k : V + (size_t)ldv * (k - 32);
            As[k * 16 + (row ^ ((k & 3) << 2))] = (active && i0 + row < r) ? -ldx<CL>(src + i0 + row) : 0.;
        }
        __syncwarp();
        const int jend = min(r, pass0 + TS_NW * 16); // columns needed by this pass
        // Bs[k][j] = (k < 32) ? V[j0 + j][k] : Z[j0 + j][k - 32]; four entries per thread
        const int bj = tid & 31, bk0 = tid >> 5; // entries k = bk0 + 16 q
        double breg[4];
        auto fetch_b = [&](int j0) {
            const bool ok = j0 + bj < r;
#pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                const int k = bk0 + 16 * q;
                const double *src = (k < 32) ? V + (size_t)ldv * k : Z + (size_t)ldv * (k - 32);
                breg[q] = ok ? ldx<CL>(src + j0 + bj) : 0.;
            }
        };
        auto park_b = [&](int buf) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                Bs[buf * 64 * TS_LD + (bk0 + 16 * q) * TS_LD + bj] = breg[q];
        };
        double cn[2][4][2]; // C fragments of the next column block
        auto load_c = [&](int j0) {
            const bool in = active && j0 < jend && j0 <= i0 + 15;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
            {
                const int j = j0 + ni * 8 + 2 * t;
                cn[0][ni][0] = (in && ra < r && j <= ra) ? __ldcg(A2 + ra + (size_t)ld * j) : 0.;
                cn[0][ni][1] = (in && ra < r && j + 1 <= ra) ? __ldcg(A2 + ra + (size_t)ld * (j + 1)) : 0.;
                cn[1][ni][0] = (in && rb < r && j <= rb) ? __ldcg(A2 + rb + (size_t)ld * j) : 0.;
                cn[1][ni][1] = (in && rb < r && j + 1 <= rb) ? __ldcg(A2 + rb + (size_t)ld * (j + 1)) : 0.;
            }
        };
        fetch_b(0);
        park_b(0);
        load_c(0);
        int buf = 0;
        for (int j0 = 0; j0 < jend; j0 += 32, buf ^= 1)
        {
            double c[2][4][2];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
            {
                c[0][ni][0] = cn[0][ni][0];
                c[0][ni][1] = cn[0][ni][1];
                c[1][ni][0] = cn[1][ni][0];
                c[1][ni][1] = cn[1][ni][1];
            }
            __syncthreads();
            const bool more = j0 + 32 < jend;
            if (more)
            {
                fetch_b(j0 + 32);
                load_c(j0 + 32);
            }
            if (active && j0 <= i0 + 15)
            {
                const double *bb = Bs + buf * 64 * TS_LD;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk)
                {
                    const int k = kk * 4 + t;
                    const double a0 = As[k * 16 + (g ^ (t << 2))];
                    const double a1 = As[k * 16 + ((8 + g) ^ (t << 2))];
                    double b[4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                        b[ni] = bb[k * TS_LD + ni * 8 + g];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
                    {
                        dmma884(c[0][ni][0], c[0][ni][1], a0, b[ni]);
                        dmma884(c[1][ni][0], c[1][ni][1], a1, b[ni]);
                    }
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                {
                    const int j = j0 + ni * 8 + 2 * t;
                    if (ra < r && j <= ra)
                        A2[ra + (size_t)ld * j] = c[0][ni][0];
                    if (ra < r && j + 1 <= ra)
                        A2[ra + (size_t)ld * (j + 1)] = c[0][ni][1];
                    if (rb < r && j <= rb)
                        A2[rb + (size_t)ld * j] = c[1][ni][0];
                    if (rb < r && j + 1 <= rb)
                        A2[rb + (size_t)ld * (j + 1)] = c[1][ni][1];
                }
            }
            if (more)
                park_b(buf ^ 1);
        }
        __syncthreads();
    }
}

/* rows a block owns in the row-distributed phases: all of them (CS == 1), else the passes of
   256 rows dealt by s1_owns; calls f(i) for every owned row i < r */
template <class F> __device__ __forceinline__ void s1_for_rows(int r, int rank, int CS, F f)
{
    if (CS == 1)
    {
        for (int i = threadIdx.x; i < r; i += TS_NT)
            f(i);
        return;
    }
    for (int q = rank; q * (TS_NW * 16) < r; q += CS)
    {
        const int i = q * (TS_NW * 16) + threadIdx.x;
        if (threadIdx.x < TS_NW * 16 && i < r)
            f(i);
    }
}

/* X = W Tf in place, row by row: x[c] = sum_{c2 <= c} w[c2] Tf[c2][c], computed from the last
   column down so that it can overwrite w (volatile: the compiler must not hoist the 528 factor
   entries out of the row loop) */
__device__ __noinline__ void s1_apply_tf(const S1Smem S, double *__restrict__ Wb, int ldws, int r, int rank,
                                         int CS)
{
    const volatile double *Tf = S.Tf;
    s1_for_rows(r, rank, CS, [&](int i) {
        double x[32];
#pragma unroll
        for (int c = 0; c < 32; ++c)
            x[c] = Wb[i + (size_t)ldws * c]; // (written by this block's own symm pass)
#pragma unroll
        for (int c = 31; c >= 0; --c)
        {
            double s = 0.;
#pragma unroll
            for (int c2 = 0; c2 <= c; ++c2)
                s += x[c2] * Tf[c2 * 33 + c];
            x[c] = s;
        }
#pragma unroll
        for (int c = 0; c < 32; ++c)
            Wb[i + (size_t)ldws * c] = x[c];
    });
}

/* Z = X - V M in place in Wb (M = S.G) */
template <bool CL>
__device__ __noinline__ void s1_form_z(const S1Smem S, const double *__restrict__ Vc, double *__restrict__ Wb,
                                       int ldws, int r, int rank, int CS)
{
    const volatile double *Mm = S.G;
    s1_for_rows(r, rank, CS, [&](int i) {
        double v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c)
            v[c] = ldx<CL>(Vc + i + (size_t)ldws * c);
#pragma unroll 1
        for (int c0 = 0; c0 < 32; c0 += 8)
        {
            double z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                z[j] = Wb[i + (size_t)ldws * (c0 + j)];
#pragma unroll
            for (int c2 = 0; c2 < 32; ++c2)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    z[j] -= v[c2] * Mm[c2 * 33 + c0 + j];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                Wb[i + (size_t)ldws * (c0 + j)] = z[j];
        }
    });
}

/* CL = false: one block per matrix, matrices from a work queue (largest first).
   CL = true:  launched with thread-block clusters; the CS blocks of a cluster share one matrix
   (a handful of very large matrices would otherwise occupy a handful of SMs): the panel
   factorisation runs on block 0, the small replicated pieces (Gram matrices, Tf, M) on every
   block, the row / tile passes of W = A22 V, X, Z and the rank-64 update are dealt to the
   blocks; four cluster barriers per panel. */
template <bool CL>
__global__ void __launch_bounds__(TS_NT, 1)
k_sy2sb(const sa_ts_mat *__restrict__ mats, int nmats, unsigned int *queue, double *wsV, double *wsW,
        int ldws, int w_in)
{
    extern __shared__ double sm_s1[];
    const S1Smem S = s1_carve(sm_s1);
    __shared__ int s_next;
    const int tid = threadIdx.x;
    int rank = 0, CS = 1;
    if (CL)
    {
        cg::cluster_group cl = cg::this_cluster();
        rank = (int)cl.block_rank();
        CS = (int)cl.num_blocks();
    }
    const int unit = blockIdx.x / CS, nunits = gridDim.x / CS; // cluster (or block) index
    auto group_sync = [&]() {
        if (CL)
            cg::this_cluster().sync();
        else
            __syncthreads();
    };
    double *Vc = wsV + (size_t)unit * ldws * TS_B;
    double *Wb = wsW + (size_t)unit * ldws * TS_B;
    int mi = CL ? unit - nunits : 0;
    while (true)
    {
        if (CL)
            mi += nunits; // static deal: every block of the cluster sees the same sequence
        else
        {
            __syncthreads();
            if (tid == 0)
                s_next = (int)atomicAdd(queue, 1u);
            __syncthreads();
            mi = s_next;
        }
        if (mi >= nmats)
            break;
        const sa_ts_mat M = mats[mi];
        const int n = M.n;
        double *T = M.T;
        if (rank == 0)
            for (int i = tid; i < n; i += TS_NT)
            {
                M.tau1[i] = 0.;
                if (M.tauz)
                    M.tauz[i] = 0.; // the one-stage back-transformation sees no reflectors
            }
        long long tclk = clock64();
        for (int j0 = 0; n - j0 - TS_B >= 2; j0 += TS_B)
        {
            const int r = n - j0 - TS_B;
            const int nr = min(TS_B, r - 1);
            double *P = T + (j0 + TS_B) + (size_t)n * j0;
            double *A2 = T + (j0 + TS_B) + (size_t)n * (j0 + TS_B);
            if (rank == 0)
                s1_panel_qr<CL>(S, P, n, r, nr, w_in, Vc, ldws, M.tau1 + j0);
            group_sync();
            TS_CLK(0);
            // compact WY factor: G = V^T V, Tf[:c, c] = -tau_c Tf[:c, :c] G[:c, c]
            s1_gram<CL>(S, Vc, Vc, ldws, r, S.G);
            if (tid < 32)
            {
                for (int c = 0; c < TS_B; ++c)
                {
                    const double tau = ldx<CL>(M.tau1 + j0 + c);
                    double s = 0.;
                    if (tid < c)
                        for (int q = tid; q < c; ++q)
                            s += S.Tf[tid * 33 + q] * S.G[q * 33 + c];
                    __syncwarp();
                    S.Tf[tid * 33 + c] = (tid < c) ? -tau * s : (tid == c ? tau : 0.);
                    __syncwarp();
                }
            }
            __syncthreads();
            TS_CLK(1);
            s1_symm<CL>(S, A2, n, r, Vc, Wb, ldws, rank, CS);
            TS_CLK(2);
            s1_apply_tf(S, Wb, ldws, r, rank, CS);
            group_sync();
            TS_CLK(3);
            // S = V^T X, M = 1/2 Tf^T S
            s1_gram<CL>(S, Vc, Wb, ldws, r, S.G);
            double mreg[2];
#pragma unroll
            for (int q = 0; q < 2; ++q)
            {
                const int e = tid + q * TS_NT, p = e >> 5, c = e & 31;
                double s = 0.;
                for (int c2 = 0; c2 <= p; ++c2)
                    s += S.Tf[c2 * 33 + p] * S.G[c2 * 33 + c];
                mreg[q] = 0.5 * s;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 2; ++q)
            {
                const int e = tid + q * TS_NT;
                S.G[(e >> 5) * 33 + (e & 31)] = mreg[q];
            }
            // (cluster mode: every block must have read X for its Gram matrix before any block
            // overwrites rows with Z)
            group_sync();
            TS_CLK(4);
            s1_form_z<CL>(S, Vc, Wb, ldws, r, rank, CS);
            group_sync();
            TS_CLK(5);
            s1_syr2k<CL>(S, A2, n, r, Vc, Wb, ldws, rank, CS);
            group_sync();
            TS_CLK(6);
        }
        group_sync();
        // band (+ zeroed bulge rows) for stage 2: band[k + 64 j] = A[j + k][j]
        for (size_t idx = (size_t)rank * TS_NT + tid; idx < (size_t)n * TS_LDB; idx += (size_t)CS * TS_NT)
        {
            const int j = (int)(idx / TS_LDB), k = (int)(idx % TS_LDB);
            M.band[idx] = (k <= TS_B && j + k < n) ? ldx<CL>(T + (j + k) + (size_t)n * j) : 0.;
        }
        TS_CLK(7);
    }
}

// ------------------------------------------------------------------------------- stage 2

/* Stage 2: band -> tridiagonal by bulge chasing.
   A TEAM of G thread blocks (8 warps each) works on one matrix; teams take the matrices
   team, team + nteams, ... of the list.  Warp gw of the team runs the sweeps gw, gw + NWT, ...
   (NWT = 8 G).  Sweep s may run its step t once sweep s-1 has completed step t+1 (the blocks the
   two touch overlap up to there), so consecutive sweeps follow each other two steps apart -- a
   dataflow pipeline without block-wide barriers: every warp publishes (matrix, sweep, completed
   steps) in a global slot after a step (behind a __threadfence) and polls the slot of the warp
   that owns the previous sweep before the next one.  Band data is read with ld.global.cg (the
   other warps of the team may sit on other SMs).
   Lane a of a warp owns row a of the 32 x 32 blocks:
     O (off-diagonal block, rows i0.., columns st..): right application of the previous
       reflector, new reflector from its first column, left application to the other columns
       (column sums through a 32 x 17 shared half tile);
     D (diagonal block at i0): H D H with the symmetric rank-2 formula; the lower triangle lives
       in registers, the transposed part is read from a 32 x 33 shared tile. */
__device__ __forceinline__ unsigned long long s2_enc(int mi, int sweep, int steps)
{
    return ((unsigned long long)(mi + 1) << 40) | ((unsigned long long)sweep << 16) |
           (unsigned long long)steps;
}

__device__ __forceinline__ void s2_two_sided(double *__restrict__ Bd, int i0, int L, double v, double tau,
                                             double *Ds, double *vs, double *qs)
{
    const int a = threadIdx.x & 31;
    double Dl[32]; // row a of the lower triangle
#pragma unroll
    for (int c = 0; c < 32; ++c)
        Dl[c] = (a >= c && a < L) ? __ldcg(Bd + (a - c) + (size_t)TS_LDB * (i0 + c)) : 0.;
#pragma unroll
    for (int c = 0; c < 32; ++c)
        Ds[a * 33 + c] = Dl[c];
    vs[a] = v;
    __syncwarp();
    double p = 0.;
#pragma unroll
    for (int c = 0; c < 32; ++c)
    {
        // D(a, c): own row for c <= a, column a of row c otherwise (zero outside the block)
        const double dac = (c <= a) ? Dl[c] : Ds[c * 33 + a];
        p += dac * vs[c];
    }
    p *= tau;
    const double pv = warp_sum(p * v);
    const double q = p - 0.5 * tau * pv * v;
    qs[a] = q;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 32; ++c)
        if (a >= c && a < L)
            Bd[(a - c) + (size_t)TS_LDB * (i0 + c)] = Dl[c] - v * qs[c] - q * vs[c];
    __syncwarp();
}

__global__ void __launch_bounds__(TS_S2_NW * 32, 2)
k_sb2st(const sa_ts_mat *__restrict__ mats, int nmats, int G, unsigned long long *slots)
{
    extern __shared__ double sm_s2[]; // per warp: tile 32 x 33, half tile 32 x 17, v, q, w
    const int tid = threadIdx.x, a = tid & 31, wid = tid >> 5;
    double *Ds = sm_s2 + (size_t)wid * TS_S2_PER_WARP, *Os = Ds + 32 * 33, *vs = Os + 32 * 17,
           *qs = vs + 32, *ws = qs + 32;
    const int team = blockIdx.x / G, nteams = gridDim.x / G;
    const int NWT = G * TS_S2_NW;
    const int gw = (blockIdx.x % G) * TS_S2_NW + wid;
    volatile unsigned long long *tslots = slots + (size_t)team * NWT;
    volatile unsigned long long *myslot = tslots + gw;
    volatile unsigned long long *prevslot = tslots + (gw + NWT - 1) % NWT;
    long long tclk = clock64();
    unsigned int mysteps = 0;
    for (int mi = team; mi < nmats; mi += nteams)
    {
        const sa_ts_mat M = mats[mi];
        const int n = M.n;
        double *Bd = M.band;
        double *T = M.T;
        const int nsweeps = n - 2; // sweeps 0 .. n-3
        for (int s = gw; s < nsweeps; s += NWT)
        {
            const int nsteps = 1 + (n - 2 - s) / TS_B;
            double vp = 0., taup = 0.;
            int st = 0, Lp = 0;
            if (a == 0)
                *myslot = s2_enc(mi, s, 0);
            for (int t = 0; t < nsteps; ++t)
            {
                // ---- wait until sweep s-1 is two steps ahead (or done)
                if (s > 0)
                {
                    if (a == 0)
                    {
                        while (true)
                        {
                            const unsigned long long pv = *prevslot;
                            const int pm = (int)(pv >> 40), ps = (int)((pv >> 16) & 0xffffffull),
                                      pt = (int)(pv & 0xffffull);
                            if (pm > mi + 1 || (pm == mi + 1 && (ps > s - 1 || (ps == s - 1 && pt >= t + 2))))
                                break;
                            __nanosleep(100);
                        }
                    }
                    __syncwarp();
                    __threadfence();
                }
                if (t == 0)
                {
                    const int i0 = s + 1;
                    const int L = min(TS_B, n - i0);
                    const double x = (a < L) ? __ldcg(Bd + (1 + a) + (size_t)TS_LDB * s) : 0.;
                    const double xn2 = warp_sum(a >= 1 ? x * x : 0.);
                    const double alpha = __shfl_sync(0xffffffffu, x, 0);
                    double beta, tau, scal;
                    ts_larfg(alpha, xn2, beta, tau, scal);
                    const double v = (a == 0) ? 1. : (a < L ? x * scal : 0.);
                    if (a < L)
                    {
                        Bd[(1 + a) + (size_t)TS_LDB * s] = (a == 0) ? beta : 0.;
                        T[s + (size_t)n * (i0 + a)] = (a == 0) ? tau : v;
                    }
                    s2_two_sided(Bd, i0, L, v, tau, Ds, vs, qs);
                    vp = v;
                    taup = tau;
                    st = i0;
                    Lp = L;
                }
                else
                {
                    const int i0 = st + TS_B;
                    const int L = min(TS_B, n - i0);
                    // O = A[i0 : i0+L, st : st+Lp]; lane a holds row a
                    double O[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        O[c] = (a < L && c < Lp) ? __ldcg(Bd + (TS_B + a - c) + (size_t)TS_LDB * (st + c)) : 0.;
                    vs[a] = vp;
                    __syncwarp();
                    double u = 0.;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        u += O[c] * vs[c];
                    u *= taup;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        O[c] -= u * vs[c];
                    __syncwarp();
                    // new reflector from column 0
                    const double x = O[0];
                    const double xn2 = warp_sum(a >= 1 ? x * x : 0.);
                    const double alpha = __shfl_sync(0xffffffffu, x, 0);
                    double beta, tau, scal;
                    ts_larfg(alpha, xn2, beta, tau, scal);
                    const double v = (a == 0) ? 1. : (a < L ? x * scal : 0.);
                    O[0] = (a == 0) ? beta : 0.;
                    // left application to columns 1..: w_c = tau sum_a v_a O[a][c], transposed
                    // through a half tile (16 columns at a time; lane l sums 16 rows of column l & 15)
                    vs[a] = v;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                    {
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            Os[a * 17 + c] = O[16 * h + c];
                        __syncwarp();
                        const int cc = a & 15, r0 = (a >> 4) * 16;
                        double wc = 0.;
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2)
                            wc += vs[r0 + r2] * Os[(r0 + r2) * 17 + cc];
                        wc += __shfl_xor_sync(0xffffffffu, wc, 16);
                        if (a < 16)
                            ws[16 * h + a] = (h == 0 && a == 0) ? 0. : tau * wc;
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = 1; c < 32; ++c)
                        O[c] -= v * ws[c];
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (a < L && c < Lp)
                            Bd[(TS_B + a - c) + (size_t)TS_LDB * (st + c)] = O[c];
                    if (a < L)
                        T[s + (size_t)n * (i0 + a)] = (a == 0) ? tau : v;
                    __syncwarp();
                    s2_two_sided(Bd, i0, L, v, tau, Ds, vs, qs);
                    vp = v;
                    taup = tau;
                    st = i0;
                    Lp = L;
                }
                ++mysteps;
                // ---- publish: the step's stores are visible before the slot changes
                __threadfence();
                __syncwarp();
                if (a == 0)
                    *myslot = s2_enc(mi, s, t + 1);
            }
        }
        // no sweeps left for this warp on this matrix: successors must not wait for it
        if (a == 0)
            *myslot = s2_enc(mi, 0xffffff, 0);
    }
    if (a == 0)
    {
        atomicAdd(&g_ts_clk[10], (unsigned long long)mysteps);
        atomicAdd(&g_ts_clk[9], (unsigned long long)(clock64() - tclk));
    }
}

/* d, e of the tridiagonal matrices (after every team is done) */
__global__ void k_sb2st_extract(const sa_ts_mat *__restrict__ mats, int nmats)
{
    const int mi = blockIdx.x;
    if (mi >= nmats)
        return;
    const sa_ts_mat M = mats[mi];
    for (int j = threadIdx.x; j < M.n; j += blockDim.x)
    {
        M.d[j] = M.band[(size_t)TS_LDB * j];
        M.e[j] = (j + 1 < M.n) ? M.band[1 + (size_t)TS_LDB * j] : 0.;
    }
}

// ------------------------------------------------------------------- back-transformation

/* z <- Q1 Q2 z for eigenvector t of the list (ev_slot / ev_idx as in the inverse iteration);
   slots that did not go through the two-stage path (ts_of_slot < 0) are skipped. */
__global__ void __launch_bounds__(256)
k_ts_back(const sa_ts_mat *__restrict__ mats, const int *__restrict__ ts_of_slot,
          const int *__restrict__ ev_slot, const int *__restrict__ ev_idx, int nev_total,
          const int64_t *__restrict__ evect_off_slot, double *__restrict__ evects)
{
    extern __shared__ double z[];
    __shared__ double s_red[8];
    const int tev = blockIdx.x;
    if (tev >= nev_total)
        return;
    const int slot = ev_slot[tev];
    const int mi = ts_of_slot[slot];
    if (mi < 0)
        return;
    const sa_ts_mat M = mats[mi];
    const int n = M.n;
    const double *T = M.T;
    double *Z = evects + evect_off_slot[slot] + (size_t)n * ev_idx[tev];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NW = blockDim.x >> 5;
    for (int i = tid; i < n; i += blockDim.x)
        z[i] = Z[i];
    __syncthreads();
    // stage-2 reflectors: sweeps in reverse order; the steps of a sweep act on disjoint rows
    for (int s = n - 3; s >= 0; --s)
    {
        const int nsteps = 1 + (n - 2 - s) / TS_B;
        for (int t = wid; t < nsteps; t += NW)
        {
            const int i0 = s + 1 + t * TS_B;
            const int L = min(TS_B, n - i0);
            const double val = (lane < L) ? T[s + (size_t)n * (i0 + lane)] : 0.;
            const double tau = __shfl_sync(0xffffffffu, val, 0);
            const double v = (lane == 0) ? 1. : val;
            const double zi = (lane < L) ? z[i0 + lane] : 0.;
            const double dot = warp_sum(v * zi);
            if (lane < L)
                z[i0 + lane] = zi - tau * dot * v;
        }
        __syncthreads();
    }
    // stage-1 reflectors: column k has its unit entry in row k + TS_B
    for (int k = n - TS_B - 2; k >= 0; --k)
    {
        const double tau = M.tau1[k];
        if (tau == 0.)
            continue; // (uniform)
        const int u = k + TS_B;
        const double *vk = T + (size_t)n * k;
        double part = 0.;
        for (int i = u + tid; i < n; i += blockDim.x)
            part += ((i == u) ? 1. : vk[i]) * z[i];
        part = warp_sum(part);
        if (lane == 0)
            s_red[wid] = part;
        __syncthreads();
        double dot = 0.;
        for (int w = 0; w < NW; ++w)
            dot += s_red[w];
        dot *= tau;
        for (int i = u + tid; i < n; i += blockDim.x)
            z[i] -= dot * ((i == u) ? 1. : vk[i]);
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x)
        Z[i] = z[i];
}

} // namespace

// ------------------------------------------------------------------------- host interface

size_t sa_ts_s1_smem_bytes(int nmax, int *w_in_out)
{
    // sub-panel width: the wider of 8 / 4 whose columns fit next to the fixed part
    const size_t cap = 200 * 1024;
    int w_in = 8;
    auto need = [&](int w) {
        const size_t u = std::max<size_t>((size_t)w * ((size_t)nmax | 1), (size_t)TS_S1_UMIN);
        return (TS_S1_FIXED + u) * sizeof(double);
    };
    while (w_in > 4 && need(w_in) > cap)
        w_in >>= 1;
    if (w_in_out)
        *w_in_out = w_in;
    return need(w_in);
}

void sa_ts_reduce(sa_gpu_ctx *ctx, const sa_ts_mat *d_mats, int nmats, int nmax, cudaStream_t st)
{
    if (nmats <= 0)
        return;
    int w_in = 8;
    const size_t smem = sa_ts_s1_smem_bytes(nmax, &w_in);
    if (smem > ctx->smem_optin)
        SA_FAIL("two-stage eigensolver: AE with %d dofs exceeds the supported size", nmax);
    SpectralWs &WS = ctx->sws;
    // clusters of CS blocks per matrix when there are far fewer matrices than SMs
    const int cs_env = getenv("SA_GPU_TS_CLUSTER") ? atoi(getenv("SA_GPU_TS_CLUSTER")) : -1; // (tests)
    int CS = 1;
    if (cs_env >= 1)
        CS = cs_env;
    else if (nmats * 2 <= ctx->num_sms && nmax >= 1024)
        CS = (nmats * 8 <= ctx->num_sms) ? 8 : (nmats * 4 <= ctx->num_sms ? 4 : 2);
    const int nunits = (CS > 1) ? nmats : std::min(nmats, ctx->num_sms);
    const int ldws = (nmax + 3) & ~3;
    WS.ts_wsV.ensure((size_t)nunits * ldws * TS_B);
    WS.ts_wsW.ensure((size_t)nunits * ldws * TS_B);
    WS.counters.ensure(4);
    SA_CUDA(cudaMemsetAsync(WS.counters.p, 0, 4 * sizeof(unsigned int), st));
    {
        ProfScope ps(ctx, "eig.ts_stage1");
        if (CS == 1)
        {
            SA_CUDA(cudaFuncSetAttribute(k_sy2sb<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_sy2sb<false><<<nunits, TS_NT, smem, st>>>(d_mats, nmats, WS.counters.p, WS.ts_wsV.p, WS.ts_wsW.p,
                                                        ldws, w_in);
        }
        else
        {
            SA_CUDA(cudaFuncSetAttribute(k_sy2sb<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg;
            std::memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(nunits * CS);
            cfg.blockDim = dim3(TS_NT);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CS;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            unsigned int *queue = WS.counters.p;
            double *wv = WS.ts_wsV.p, *ww = WS.ts_wsW.p;
            int ldws_ = ldws, w_in_ = w_in, nm = nmats;
            SA_CUDA(cudaLaunchKernelEx(&cfg, k_sy2sb<true>, d_mats, nm, queue, wv, ww, ldws_, w_in_));
        }
        ctx->launches++;
        SA_CUDA(cudaGetLastError());
    }
    {
        ProfScope ps(ctx, "eig.ts_stage2");
        // teams of G blocks per matrix: one block when there are many matrices, up to 8 when a
        // handful of large ones would otherwise leave most SMs idle
        // (measured at 630 matrices of n ~ 2000: teams of 5 blocks, which would keep the bands
        // of the matrices in flight in L2, are 40 % slower than one block per matrix -- passing
        // the band between SMs costs more than streaming it from HBM)
        // blocks in flight: every chase step re-reads what the previous sweep wrote two steps earlier,
        // so the reuse distance is (warps in flight) x 24 KB x 2 -- SA_GPU_TS_S2_BPS (blocks per SM)
        const int bps = getenv("SA_GPU_TS_S2_BPS") ? std::max(1, std::min(2, atoi(getenv("SA_GPU_TS_S2_BPS")))) : 2;
        const int cap = bps * ctx->num_sms;
        int G = std::max(1, std::min(8, cap / nmats));
        G = std::min(G, std::max(1, (nmax / (2 * TS_B) + TS_S2_NW - 1) / TS_S2_NW)); // useful concurrency
        const int nteams = std::min(nmats, cap / G);
        const size_t nslots = (size_t)nteams * G * TS_S2_NW;
        WS.ts_slots.ensure(nslots);
        SA_CUDA(cudaMemsetAsync(WS.ts_slots.p, 0, nslots * sizeof(unsigned long long), st));
        const size_t smem2 = (size_t)TS_S2_NW * TS_S2_PER_WARP * sizeof(double);
        SA_CUDA(cudaFuncSetAttribute(k_sb2st, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        // (all blocks are co-resident: grid <= 2 blocks per SM)
        k_sb2st<<<nteams * G, TS_S2_NW * 32, smem2, st>>>(d_mats, nmats, G,
                                                         (unsigned long long *)WS.ts_slots.p);
        ctx->launches++;
        SA_CUDA(cudaGetLastError());
        k_sb2st_extract<<<nmats, 256, 0, st>>>(d_mats, nmats);
        ctx->launches++;
        SA_CUDA(cudaGetLastError());
    }
}

void sa_ts_back(sa_gpu_ctx *ctx, const sa_ts_mat *d_mats, const int *d_ts_of_slot, const int *d_ev_slot,
                const int *d_ev_idx, int nev_total, const int64_t *d_evect_off_slot, double *d_evects,
                int nmax, cudaStream_t st)
{
    if (nev_total <= 0)
        return;
    const size_t smem = (size_t)nmax * sizeof(double);
    SA_CUDA(cudaFuncSetAttribute(k_ts_back, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)std::max<size_t>(smem, 1024)));
    k_ts_back<<<nev_total, 256, smem, st>>>(d_mats, d_ts_of_slot, d_ev_slot, d_ev_idx, nev_total,
                                            d_evect_off_slot, d_evects);
    ctx->launches++;
    SA_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------ diagnostics (tests)

/* Runs the two-stage reduction on one dense symmetric matrix given by the caller (host,
   n x n column-major).  Outputs (any may be NULL): T_out n x n (band + reflectors), tau1_out n,
   d_out n, e_out n.  The device state is kept for sa_gpu_debug_twostage_back. */
extern "C" int sa_gpu_debug_twostage(sa_gpu_ctx *ctx, int n, const double *A, double *T_out,
                                     double *tau1_out, double *d_out, double *e_out)
{
    SA_API_BEGIN
    cudaStream_t st = ctx->stream;
    SpectralWs &WS = ctx->sws;
    WS.Twork.upload(A, (size_t)n * n, st);
    WS.ts_band.ensure((size_t)n * TS_LDB);
    WS.ts_tau1.ensure(n);
    WS.d.ensure(n);
    WS.e.ensure(n);
    sa_ts_mat m;
    m.n = n;
    m.T = WS.Twork.p;
    m.band = WS.ts_band.p;
    m.d = WS.d.p;
    m.e = WS.e.p;
    m.tau1 = WS.ts_tau1.p;
    m.tauz = nullptr;
    WS.ts_mats.ensure(sizeof(sa_ts_mat) / sizeof(int64_t));
    SA_CUDA(cudaMemcpyAsync(WS.ts_mats.p, &m, sizeof m, cudaMemcpyHostToDevice, st));
    SA_CUDA(cudaStreamSynchronize(st));
    sa_ts_reduce(ctx, (const sa_ts_mat *)WS.ts_mats.p, 1, n, st);
    if (T_out)
        WS.Twork.download(T_out, (size_t)n * n, st);
    if (tau1_out)
        WS.ts_tau1.download(tau1_out, n, st);
    if (d_out)
        WS.d.download(d_out, n, st);
    if (e_out)
        WS.e.download(e_out, n, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

/* Y (n x nvec, column-major, in/out) <- Q1 Q2 Y with the reflectors of the last
   sa_gpu_debug_twostage call. */
extern "C" int sa_gpu_debug_twostage_back(sa_gpu_ctx *ctx, int n, int nvec, double *Y)
{
    SA_API_BEGIN
    cudaStream_t st = ctx->stream;
    SpectralWs &WS = ctx->sws;
    DevBuf<double> dY;
    DevBuf<int> slot, idx, tsof;
    DevBuf<int64_t> off;
    std::vector<int> h_slot(nvec, 0), h_idx(nvec);
    for (int j = 0; j < nvec; ++j)
        h_idx[j] = j;
    const int zero = 0;
    const int64_t zero64 = 0;
    dY.upload(Y, (size_t)n * nvec, st);
    slot.upload(h_slot.data(), nvec, st);
    idx.upload(h_idx.data(), nvec, st);
    tsof.upload(&zero, 1, st);
    off.upload(&zero64, 1, st);
    sa_ts_back(ctx, (const sa_ts_mat *)WS.ts_mats.p, tsof.p, slot.p, idx.p, nvec, off.p, dY.p, n, st);
    dY.download(Y, (size_t)n * nvec, st);
    SA_CUDA(cudaStreamSynchronize(st));
    SA_API_END
}

extern "C" int sa_gpu_debug_ts_clocks(double *out16)
{
    unsigned long long h[16], z[16];
    std::memset(z, 0, sizeof z);
    if (cudaMemcpyFromSymbol(h, g_ts_clk, sizeof h) != cudaSuccess)
        return 1;
    cudaMemcpyToSymbol(g_ts_clk, z, sizeof z);
    for (int i = 0; i < 16; ++i)
        out16[i] = (double)h[i];
    return 0;
}
