// Included inside eigen.cu's anonymous namespace.
//
// k_tridiag_coop: Householder tridiagonalisation of LARGE AE matrices (coarse levels,
// n ~ 10^3) that live in global memory, with a GROUP of thread blocks per matrix
// (cooperative launch: all blocks are co-resident, blocks of a group synchronise through a
// monotonic counter in global memory).
//
// One group barrier and ONE pass over the trailing block per Householder step:
//   - every block redundantly reconstructs column k from the stale stored column and the
//     pending rank-2 update (v, w of step k-1), forms the new reflector (norm, beta, tau,
//     v_new) -- no cross-block reduction needed;
//   - fused pass over the rows the block owns: apply the pending update
//     T_ij -= v_i w_j + w_i v_j and accumulate p_i = sum_j T_ij v_new_j in the same sweep
//     (2 r^2 words of traffic per step instead of 3 r^2);
//   - p (own rows) and the partial p.v go to global memory; barrier; every block reads p and
//     forms w_new.
// Column k of T is never written again after step k-1 started (it is "virtual"), which is
// what removes the second barrier; reflectors go to the packed array V.
// Rows are owned in blocks of 32 (row block rb belongs to block rb % G) so that work stays
// balanced as the trailing block shrinks; inside a block a warp owns (row block, column
// segment) and lanes are rows: all global accesses are 256-byte coalesced.

struct CoopMatrix
{
    int slot;       // chunk slot of the AE
    double *T;      // n x n scaled matrix (full, symmetric), overwritten
    double *pbuf;   // 2 n (double buffered by step parity)
    double *pvacc;  // 3 rotating accumulators
    unsigned int *counter;
};

__device__ __forceinline__ void group_barrier(unsigned int *counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned int *)counter) < target)
            __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512, 1)
k_tridiag_coop(LevelTables L, ChunkDev C, const CoopMatrix *mats, int G)
{
    extern __shared__ double sm[];
    const int gidx = blockIdx.x / G; // matrix of this block
    const int g = blockIdx.x % G;    // rank inside the group
    const CoopMatrix M = mats[gidx];
    const int slot = M.slot;
    const int part = C.ae_of_slot[slot];
    const int n = L.AE2d_I[part + 1] - L.AE2d_I[part];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = blockDim.x >> 5;
    double *T = M.T;
    double *v = sm;          // pending reflector (global row index)
    double *w = v + n;       // pending w
    double *vn = w + n;      // new reflector
    double *red = vn + n;    // 64 doubles of reduction scratch
    double *psum = red + 64; // NW * 32 partial sums
    double *dd = C.d + C.doff[slot], *ee = C.e + C.doff[slot], *tt = C.tau + C.doff[slot];
    double *Vp = C.V + C.voff[slot]; // packed lower triangle by columns
    unsigned int epoch = 0;
    bool pending = false; // is there a rank-2 update (v, w) not yet applied?

    for (int i = tid; i < n; i += blockDim.x)
    {
        v[i] = 0.;
        w[i] = 0.;
    }
    __syncthreads();

    for (int k = 0; k < n; ++k)
    {
        // ---- A. column k with the pending update applied (rows >= k), redundantly
        const double vk = v[k], wk = w[k];
        double nrm = 0.;
        for (int i = k + tid; i < n; i += blockDim.x)
        {
            double x = __ldcg(T + i + (int64_t)n * k);
            if (pending)
                x -= v[i] * wk + w[i] * vk;
            vn[i] = x;
            if (i >= k + 2)
                nrm += x * x;
        }
        for (int o = 16; o > 0; o >>= 1)
            nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0)
            red[wid] = nrm;
        __syncthreads();
        double xnorm2 = 0.;
        for (int s = 0; s < NW; ++s)
            xnorm2 += red[s];
        const double dk = vn[k];
        if (k == n - 1)
        {
            if (g == 0 && tid == 0)
            {
                dd[k] = dk;
                ee[k] = 0.;
                tt[k] = 0.;
            }
            break;
        }
        const double alpha = vn[k + 1];
        double tau = 0., beta = alpha, scal = 0.;
        if (xnorm2 > 0.)
        {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tau = (beta - alpha) / beta;
            scal = 1. / (alpha - beta);
        }
        __syncthreads(); // everyone has read vn[k], vn[k+1]
        const int64_t cjm = (int64_t)k * n - ((int64_t)k * (k - 1)) / 2 - k;
        for (int i = k + 1 + tid; i < n; i += blockDim.x)
        {
            const double vi = (i == k + 1) ? 1. : vn[i] * scal;
            vn[i] = vi;
            if (g == 0 && i >= k + 2)
                Vp[cjm + i] = vi; // reflector storage (packed column k)
        }
        if (g == 0 && tid == 0)
        {
            dd[k] = dk;
            ee[k] = beta;
            tt[k] = tau;
        }
        __syncthreads();

        // ---- B. fused pass over the owned rows of the trailing block (rows, cols >= k+1)
        const int rb0 = (k + 1) >> 5, rb1 = (n - 1) >> 5;
        // first owned row block >= rb0
        int first = rb0 + ((g - (rb0 % G)) % G + G) % G;
        const int nown = (first > rb1) ? 0 : (rb1 - first) / G + 1;
        double pvpart = 0.;
        if (nown > 0)
        {
            // S column segments per row block so that all warps have work
            int S = 1;
            while (S * 2 * nown <= NW)
                S *= 2;
            const int ncols = n - (k + 1);
            const int segw = (ncols + S - 1) / S;
            for (int ob = wid / S; ob < nown; ob += max(1, NW / S))
            {
                const int seg = wid % S;
                const int rb = first + ob * G;
                const int i = (rb << 5) + lane;
                const bool rowok = (i >= k + 1) && (i < n);
                const int j0 = k + 1 + seg * segw;
                const int j1 = min(n, j0 + segw);
                double acc = 0.;
                if (rowok)
                {
                    const double vi = v[i], wi = w[i];
                    double *Tp = T + i + (int64_t)n * j0;
                    int j = j0;
                    const int64_t ln = n;
                    if (pending)
                    {
                        for (; j + 8 <= j1; j += 8, Tp += 8 * ln)
                        {
                            double t[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                t[u] = Tp[u * ln];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                            {
                                t[u] -= vi * w[j + u] + wi * v[j + u];
                                Tp[u * ln] = t[u];
                                acc += t[u] * vn[j + u];
                            }
                        }
                        for (; j < j1; ++j, Tp += ln)
                        {
                            const double t0 = Tp[0] - (vi * w[j] + wi * v[j]);
                            Tp[0] = t0;
                            acc += t0 * vn[j];
                        }
                    }
                    else
                    {
                        for (; j + 8 <= j1; j += 8, Tp += 8 * ln)
                        {
                            double t[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                t[u] = Tp[u * ln];
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                acc += t[u] * vn[j + u];
                        }
                        for (; j < j1; ++j, Tp += ln)
                            acc += Tp[0] * vn[j];
                    }
                }
                if (S == 1)
                {
                    if (rowok)
                    {
                        const double p = tau * acc;
                        M.pbuf[(int64_t)(k & 1) * n + i] = p;
                        pvpart += p * vn[i];
                    }
                }
                else
                    psum[wid * 32 + lane] = rowok ? acc : 0.;
            }
            if (S > 1)
            {
                __syncthreads();
                // warp (ob * S) sums the segments of its row block
                if ((wid % S) == 0 && (wid / S) < nown)
                {
                    const int rb = first + (wid / S) * G;
                    const int i = (rb << 5) + lane;
                    if (i >= k + 1 && i < n)
                    {
                        double acc = 0.;
                        for (int s = 0; s < S; ++s)
                            acc += psum[(wid + s) * 32 + lane];
                        const double p = tau * acc;
                        M.pbuf[(int64_t)(k & 1) * n + i] = p;
                        pvpart += p * vn[i];
                    }
                }
            }
        }
        // partial p.v of this block -> rotating global accumulator
        for (int o = 16; o > 0; o >>= 1)
            pvpart += __shfl_xor_sync(0xffffffffu, pvpart, o);
        __syncthreads();
        if (lane == 0)
            red[32 + wid] = pvpart;
        __syncthreads();
        if (tid == 0)
        {
            double s = 0.;
            for (int q = 0; q < NW; ++q)
                s += red[32 + q];
            atomicAdd(M.pvacc + (k % 3), s);
            if (g == 0)
                M.pvacc[(k + 1) % 3] = 0.; // its last readers (step k-2) are past barrier k-1
        }
        ++epoch;
        group_barrier(M.counter, epoch * (unsigned int)G);

        // ---- D. w_new = p + alpha2 v_new (redundantly), becomes the pending update
        const double pv = __ldcg(M.pvacc + (k % 3));
        const double alpha2 = -0.5 * tau * pv;
        for (int i = k + 1 + tid; i < n; i += blockDim.x)
        {
            const double vi = vn[i];
            v[i] = vi;
            w[i] = (tau != 0.) ? __ldcg(M.pbuf + (int64_t)(k & 1) * n + i) + alpha2 * vi : 0.;
        }
        for (int i = tid; i <= k; i += blockDim.x)
        {
            v[i] = 0.;
            w[i] = 0.;
        }
        pending = true;
        __syncthreads();
    }
}

// k_tridiag_coop_sym: same algorithm and launch shape as k_tridiag_coop, but only the LOWER
// triangle of the trailing block is read and written: half the HBM traffic, which is what
// bounds this kernel.  The matrix is walked in 32 x 32 tiles (row block rb >= column block
// cb).  An entry T_ij of a tile serves both p_i += T_ij v_j (lane = row: a private running
// sum) and p_j += T_ij v_i (j < i): for the second one the updated tile half is parked in a
// per-warp shared-memory buffer and summed down its columns by lanes = columns, so there are
// no shuffles per entry and no atomics.
// Ownership / determinism:
//   - row blocks are dealt to the G blocks of the group in snake order (balanced triangular
//     work); inside a block, the column blocks are dealt to the warps (snake order): the column
//     sums of cb are then a single warp's register, written once per step to this block's row
//     of pcol[G][n];
//   - row sums go to a per-warp shared array and are added in warp order into prow[n];
//   - after the group barrier every block forms p = tau (prow + sum_g pcol[g]) in the same
//     order, the block-wide p.v reduction and w: all blocks hold bit-identical v, w.
// pbuf layout per matrix: prow[2][n], pcol[2][G][n] (double buffered by step parity).
__global__ void __launch_bounds__(512, 1)
k_tridiag_coop_sym(LevelTables L, ChunkDev C, const CoopMatrix *mats, int G, int QMAX)
{
    extern __shared__ double sm[];
    const int gidx = blockIdx.x / G;
    const int g = blockIdx.x % G;
    const CoopMatrix M = mats[gidx];
    const int slot = M.slot;
    const int part = C.ae_of_slot[slot];
    const int n = L.AE2d_I[part + 1] - L.AE2d_I[part];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int NW = blockDim.x >> 5;
    const int NRB = (n + 31) >> 5, npad = NRB << 5;
    double *T = M.T;
    double *v = sm;            // pending reflector (global row index), padded to npad
    double *w = v + npad;      // pending w
    double *vn = w + npad;     // new reflector
    double *red = vn + npad;   // 64 doubles of reduction scratch
    double *sbuf = red + 64 + (size_t)wid * (16 * 33); // per-warp tile half, [column][row] stride 33
    double *prw_all = red + 64 + (size_t)NW * (16 * 33);
    double *prw = prw_all + (size_t)wid * QMAX * 32;  // per-warp row sums of the owned row blocks
    double *dd = C.d + C.doff[slot], *ee = C.e + C.doff[slot], *tt = C.tau + C.doff[slot];
    double *Vp = C.V + C.voff[slot]; // packed lower triangle by columns
    double *prow_g = M.pbuf;
    double *pcol_g = M.pbuf + 2 * (int64_t)n;
    unsigned int epoch = 0;
    bool pending = false;

    for (int i = tid; i < npad; i += blockDim.x)
    {
        v[i] = 0.;
        w[i] = 0.;
        vn[i] = 0.;
    }
    __syncthreads();

    for (int k = 0; k < n; ++k)
    {
        // ---- A. column k with the pending update applied (rows >= k), redundantly
        const double vk = v[k], wk = w[k];
        double nrm = 0.;
        for (int i = k + tid; i < n; i += blockDim.x)
        {
            double x = __ldcg(T + i + (int64_t)n * k);
            if (pending)
                x -= v[i] * wk + w[i] * vk;
            vn[i] = x;
            if (i >= k + 2)
                nrm += x * x;
        }
        for (int o = 16; o > 0; o >>= 1)
            nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0)
            red[wid] = nrm;
        __syncthreads();
        double xnorm2 = 0.;
        for (int s = 0; s < NW; ++s)
            xnorm2 += red[s];
        const double dk = vn[k];
        if (k == n - 1)
        {
            if (g == 0 && tid == 0)
            {
                dd[k] = dk;
                ee[k] = 0.;
                tt[k] = 0.;
            }
            break;
        }
        const double alpha = vn[k + 1];
        double tau = 0., beta = alpha, scal = 0.;
        if (xnorm2 > 0.)
        {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tau = (beta - alpha) / beta;
            scal = 1. / (alpha - beta);
        }
        __syncthreads(); // everyone has read vn[k], vn[k+1]
        const int64_t cjm = (int64_t)k * n - ((int64_t)k * (k - 1)) / 2 - k;
        for (int i = k + 1 + tid; i < n; i += blockDim.x)
        {
            const double vi = (i == k + 1) ? 1. : vn[i] * scal;
            vn[i] = vi;
            if (g == 0 && i >= k + 2)
                Vp[cjm + i] = vi; // reflector storage (packed column k)
        }
        if (g == 0 && tid == 0)
        {
            dd[k] = dk;
            ee[k] = beta;
            tt[k] = tau;
        }
        // this warp's row-sum slots
        for (int q = lane; q < QMAX * 32; q += 32)
            prw[q] = 0.;
        __syncthreads();

        // ---- B. fused pass over the owned tiles of the lower triangle (rows, cols >= k+1)
        const int k1 = k + 1;
        const int cb0 = k1 >> 5;
        const int par = k & 1;
        // column blocks are dealt to the warps in snake order as well: block cb has about
        // (NRB - cb) / G tiles here, so a plain cyclic deal would load the low warps 1.6x more
        for (int cblk = cb0 / NW; cblk * NW < NRB; ++cblk)
        {
            const int cb = cblk * NW + ((cblk & 1) ? (NW - 1 - wid) : wid);
            if (cb < cb0 || cb >= NRB)
                continue;
            double colacc = 0.; // lane = column (cb << 5) + lane
            for (int q = 0; q < QMAX; ++q)
            {
                const int rb = q * G + ((q & 1) ? (G - 1 - g) : g);
                if (rb >= NRB)
                    break;
                if (rb < cb || (rb << 5) + 31 < k1)
                    continue;
                const int i = (rb << 5) + lane;
                const bool rowok = i >= k1 && i < n;
                const double vi = v[i], wi = w[i];
                double acc = 0.;
                for (int h = 0; h < 2; ++h)
                {
                    const int jb = (cb << 5) + h * 16;
                    // all 16 loads of the half tile are issued before anything depends on them
                    // (predicated, no branches: edge and diagonal tiles cost the same as
                    // interior ones).  Measured: 8-column double-buffered batches are slower
                    // (fewer loads in flight, register spills).
                    double *Tp = T + i + (int64_t)n * jb;
                    double t[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                    {
                        const int j = jb + c;
                        const bool ok = rowok && j >= k1 && j <= i;
                        t[c] = ok ? Tp[(int64_t)n * c] : 0.;
                    }
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                    {
                        const int j = jb + c;
                        const bool ok = rowok && j >= k1 && j <= i;
                        if (ok)
                        {
                            if (pending)
                            {
                                t[c] -= vi * w[j] + wi * v[j];
                                Tp[(int64_t)n * c] = t[c];
                            }
                            acc += t[c] * vn[j];
                            if (j == i)
                                t[c] = 0.; // the diagonal is not part of the transposed product
                        }
                        sbuf[c * 33 + lane] = t[c];
                    }
                    __syncwarp();
                    {
                        // column sums of this half: lane -> (column c, half of the rows)
                        const int c = lane & 15, hr = lane >> 4;
                        const double *sb = sbuf + c * 33 + hr * 16;
                        const double *vr = vn + (rb << 5) + hr * 16;
                        double s = 0.;
#pragma unroll
                        for (int r = 0; r < 16; ++r)
                            s += sb[r] * vr[r];
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (hr == h)
                            colacc += s;
                    }
                    __syncwarp();
                }
                if (rowok)
                    prw[q * 32 + lane] += acc;
            }
            const int j = (cb << 5) + lane;
            if (j >= k1 && j < n)
                pcol_g[((int64_t)par * G + g) * n + j] = colacc;
        }
        __syncthreads();
        // row sums of the owned row blocks: warp partials added in warp order
        for (int t = tid; t < QMAX * 32; t += blockDim.x)
        {
            const int q = t >> 5;
            const int rb = q * G + ((q & 1) ? (G - 1 - g) : g);
            const int i = (rb << 5) + (t & 31);
            if (rb < NRB && i >= k1 && i < n)
            {
                double s = 0.;
                for (int ww = 0; ww < NW; ++ww)
                    s += prw_all[(size_t)ww * QMAX * 32 + t];
                prow_g[(int64_t)par * n + i] = s;
            }
        }
        ++epoch;
        group_barrier(M.counter, epoch * (unsigned int)G);

        // ---- D. p = tau (row part + column parts), p.v, w_new = p + alpha2 v_new; redundantly
        double pvp = 0.;
        for (int i = k1 + tid; i < n; i += blockDim.x)
        {
            double p = __ldcg(prow_g + (int64_t)par * n + i);
            for (int gg = 0; gg < G; ++gg)
                p += __ldcg(pcol_g + ((int64_t)par * G + gg) * n + i);
            p *= tau;
            w[i] = p;
            pvp += p * vn[i];
        }
        for (int o = 16; o > 0; o >>= 1)
            pvp += __shfl_xor_sync(0xffffffffu, pvp, o);
        if (lane == 0)
            red[32 + wid] = pvp;
        __syncthreads();
        double pv = 0.;
        for (int s = 0; s < NW; ++s)
            pv += red[32 + s];
        const double alpha2 = -0.5 * tau * pv;
        for (int i = k1 + tid; i < n; i += blockDim.x)
        {
            const double vi = vn[i];
            v[i] = vi;
            w[i] = (tau != 0.) ? w[i] + alpha2 * vi : 0.;
        }
        for (int i = tid; i <= k; i += blockDim.x)
        {
            v[i] = 0.;
            w[i] = 0.;
        }
        pending = true;
        __syncthreads();
    }
}
