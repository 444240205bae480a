// Included inside eigen.cu's anonymous namespace (after ChunkDev / g_phase_clk).
//
// k_at_packed<max threads, min blocks per SM>: assemble + weighted-l1 scaling + Householder tridiagonalisation of one AE
// per thread block with the matrix held as a PACKED LOWER TRIANGLE in shared memory
// (entry (i,j), i >= j, at T[cjm[j] + i]).  Half the footprint of the square tile means
// 2-3 resident blocks per SM for n ~ 100-150 -- the kernel is bound by the latency of the
// ~n dependent Householder steps, so co-resident blocks convert directly into throughput --
// and half the rank-2 update work.  Both access patterns are conflict free:
//   row pattern    (fixed column j, lanes = consecutive rows)      -> consecutive words
//   column pattern (lane a walks down its own column a)            -> stride n - a
// Reflectors are written out in the same packed layout (halves the HBM traffic of the
// back-transformation).
// Shared layout (doubles): [slots 32][v n][w n][dg n][psum NT][cjm n ints][tile n(n+1)/2]

template <bool ABS>
__device__ __forceinline__ double packed_row_dot(const double *T, const int *cjm, int n, int k0,
                                                 int ia, int c, int ncls, const double *x)
{
    // sum over stored row part  j in [k0, ia], j == k0 + c (mod ncls):  A(ia, j) x_j
    // and column part           i in (ia, n),  i == ia + 1 + c (mod ncls): A(i, ia) x_i
    double s0 = 0., s1 = 0.;
    int j = k0 + c;
    // column offsets cjm[j] = j (2n - 1 - j) / 2 are computed, not loaded: the kernel is bound
    // by shared-memory wavefronts (ncu: 73 % of peak), integer issue slots are free
    const int tn1 = 2 * n - 1;
#define SA_CJM(jj) ((((jj) * (tn1 - (jj))) >> 1))
    for (; j + ncls <= ia; j += 2 * ncls)
    {
        const double a0 = T[SA_CJM(j) + ia], a1 = T[SA_CJM(j + ncls) + ia];
        s0 += (ABS ? fabs(a0) : a0) * x[j];
        s1 += (ABS ? fabs(a1) : a1) * x[j + ncls];
    }
    if (j <= ia)
    {
        const double a0 = T[SA_CJM(j) + ia];
        s0 += (ABS ? fabs(a0) : a0) * x[j];
    }
    const double *col = T + SA_CJM(ia);
    int i = ia + 1 + c;
    for (; i + ncls < n; i += 2 * ncls)
    {
        const double a0 = col[i], a1 = col[i + ncls];
        s0 += (ABS ? fabs(a0) : a0) * x[i];
        s1 += (ABS ? fabs(a1) : a1) * x[i + ncls];
    }
    if (i < n)
    {
        const double a0 = col[i];
        s0 += (ABS ? fabs(a0) : a0) * x[i];
    }
    return s0 + s1;
}

template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_at_packed(LevelTables L, ChunkDev C, const int *slot_list, double *ae_D, double *G, int GS)
{
    extern __shared__ double sm[];
    const int slot = slot_list[blockIdx.x];
    const int part = C.ae_of_slot[slot];
    const int rb = L.AE2d_I[part];
    const int n = L.AE2d_I[part + 1] - rb;
    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *slots_n = sm;
    double *slots_p = sm + 16;
    double *v = sm + 32;
    double *w = v + n;
    double *dg = w + n;
    double *psum = dg + n;
    int *cjm = (int *)(psum + NT);
    double *T = psum + NT + ((n + 1) >> 1);
    const int np = n * (n + 1) / 2;
    double *dd = C.d + C.doff[slot], *ee = C.e + C.doff[slot], *tt = C.tau + C.doff[slot],
           *sinv = C.sinv + C.doff[slot];
    double *Vout = C.V + C.voff[slot];

    long long tc0 = clock64();
    for (int j = tid; j < n; j += NT)
        cjm[j] = j * n - (j * (j - 1)) / 2 - j;
    __syncthreads();
    {
        PackedTile pt;
        pt.T = T;
        pt.cjm = cjm;
        // the vector / reduction area (v, w, dg, psum) is free until the scaling step
        if (!sa_dev_assemble_AE_staged(L, part, pt, (void *)v, (3 * n + NT) * (int)sizeof(double)))
            sa_dev_assemble_AE_tile(L, part, pt);
    }
    long long tc1 = clock64();
    if (tid == 0)
        atomicAdd(&g_phase_clk[0], (unsigned long long)(tc1 - tc0));

    // weighted-l1 diagonal: D_ii = sqrt(a_ii) * sum_j |a_ij| / sqrt(a_jj)  (amg/src/mbox.cpp:913-949)
    int bad = 0;
    for (int i = tid; i < n; i += NT)
    {
        const double a = T[cjm[i] + i];
        if (!(a > 0.))
            bad = 1;
        dg[i] = a;
        w[i] = rsqrt(a);
    }
    __syncthreads();
    {
        const int rpad = (n + 31) & ~31;
        const int ncls = max(1, NT / rpad);
        const int a = (ncls == 1) ? tid : tid % rpad;
        const int c = (ncls == 1) ? 0 : tid / rpad;
        for (int a0 = 0; a0 < n; a0 += (ncls == 1 ? NT : rpad))
        {
            const int i = a0 + a;
            double s = 0.;
            if (i < n && c < ncls)
                s = packed_row_dot<true>(T, cjm, n, 0, i, c, ncls, w);
            psum[tid] = s;
            __syncthreads();
            if (i < n && c == 0)
            {
                double t = s;
                for (int cc = 1; cc < ncls; ++cc)
                    t += psum[cc * rpad + a];
                const double sum = sqrt(dg[i]) * t;
                ae_D[rb + i] = sum;
                const double si = 1. / sqrt(sum);
                v[i] = si;
                sinv[i] = si;
                if (!(sum > 0.) || !isfinite(sum))
                    bad = 1;
            }
            __syncthreads();
        }
    }
    if (__syncthreads_or(bad))
    {
        if (tid == 0)
            C.status[slot] = 1;
        return;
    }
    // A^ = D^-1/2 A D^-1/2 on the stored triangle: one warp per column
    for (int j = wid; j < n; j += (NT >> 5))
    {
        const double sj = v[j];
        double *col = T + cjm[j];
        for (int i = j + lane; i < n; i += 32)
            col[i] *= v[i] * sj;
    }
    __syncthreads();
    long long tc2 = clock64();
    if (tid == 0)
        atomicAdd(&g_phase_clk[1], (unsigned long long)(tc2 - tc1));

    if (G)
    {
        // The reduction runs in k_tridiag_reg (eigen_reg.cuh): write the scaled matrix in its
        // layout -- full square, entry (16 a + r, 16 b + c) at (a GS + b) 256 + 16 r + c.
        double *Gt = G + (size_t)blockIdx.x * GS * GS * 256;
        const int tot = GS * GS * 256;
        for (int q = tid; q < tot; q += NT)
        {
            const int e = q >> 8, t = q & 255;
            const int i = 16 * (e / GS) + (t >> 4), j = 16 * (e % GS) + (t & 15);
            double x = 0.;
            if (i < n && j < n)
                x = i >= j ? T[cjm[j] + i] : T[cjm[i] + j];
            __stcs(Gt + q, x);
        }
        return;
    }
    // ---- Householder tridiagonalisation (dsytd2 recurrences, lower)
    {
        double part2 = 0.;
        for (int i = 2 + tid; i < n; i += NT)
        {
            const double x = T[i]; // column 0
            part2 += x * x;
        }
        for (int o = 16; o > 0; o >>= 1)
            part2 += __shfl_xor_sync(0xffffffffu, part2, o);
        if (lane == 0)
            slots_n[wid] = part2;
    }
    __syncthreads();
    int nslots_n = NT >> 5;
    int cur_rpad = -1, ncls = 1, my_a = tid, my_c = 0;
    int cur_rhpad = -1, ncls2 = 1, my_a2 = tid, my_c2 = 0; // row-pair mapping of the update
    for (int k = 0; k < n - 1; ++k)
    {
        const int r = n - k - 1;
        const int rpad = (r + 31) & ~31;
        if (rpad != cur_rpad)
        {
            cur_rpad = rpad;
            ncls = max(1, NT / rpad);
            my_a = (ncls == 1) ? tid : tid % rpad;
            my_c = (ncls == 1) ? 0 : tid / rpad;
        }
        const int rhpad = (((r + 1) >> 1) + 31) & ~31;
        if (rhpad != cur_rhpad)
        {
            cur_rhpad = rhpad;
            ncls2 = max(1, NT / rhpad);
            my_a2 = (ncls2 == 1) ? tid : tid % rhpad;
            my_c2 = (ncls2 == 1) ? 0 : tid / rhpad;
        }
        double xnorm2 = 0.;
        for (int s = 0; s < nslots_n; ++s)
            xnorm2 += slots_n[s];
        double *colk = T + cjm[k];
        const double alpha = colk[k + 1];
        double tau = 0., beta = alpha, scal = 0.;
        if (xnorm2 > 0.)
        {
            const double nx2 = alpha * alpha + xnorm2;
            const double rinv = rsqrt(nx2);
            const double nx = nx2 * rinv;
            beta = -copysign(nx, alpha);
            tau = 1. + alpha * copysign(rinv, alpha);
            scal = 1. / (alpha - beta);
        }
        for (int a = tid; a < r; a += NT)
        {
            const int i = k + 1 + a;
            const double vi = (a == 0) ? 1. : colk[i] * scal;
            v[i] = vi;
            if (a > 0)
                colk[i] = vi; // keep the reflector in place
        }
        if (tid == 0)
        {
            dd[k] = colk[k];
            ee[k] = beta;
            tt[k] = tau;
        }
        __syncthreads(); // (1)
        const int nrw = (min(r, NT) + 31) >> 5;
        double nrm = 0.;
        if (tau != 0.)
        {
            // p = tau * A22 v
            double pv = 0.;
            for (int a = my_a; a < r; a += (ncls == 1 ? NT : r))
            {
                // (ncls > 1: single pass with a == my_a; ncls == 1: rows strided over threads)
                double s = 0.;
                if (my_c < ncls)
                    s = packed_row_dot<false>(T, cjm, n, k + 1, k + 1 + a, my_c, ncls, v);
                if (ncls == 1)
                {
                    s *= tau;
                    w[k + 1 + a] = s;
                    pv += s * v[k + 1 + a];
                }
                else
                    psum[tid] = s;
                if (ncls > 1)
                    break;
            }
            if (ncls > 1)
            {
                if (my_a >= r)
                    psum[tid] = 0.;
                __syncthreads(); // (2)
                if (tid < r)
                {
                    double t = psum[tid];
#pragma unroll 4
                    for (int cc = 1; cc < ncls; ++cc)
                        t += psum[cc * rpad + tid];
                    t *= tau;
                    w[k + 1 + tid] = t;
                    pv = t * v[k + 1 + tid];
                }
            }
            for (int o = 16; o > 0; o >>= 1)
                pv += __shfl_xor_sync(0xffffffffu, pv, o);
            if (lane == 0)
                slots_p[wid] = pv;
            __syncthreads(); // (3)
            double pvs = 0.;
            for (int s = 0; s < nrw; ++s)
                pvs += slots_p[s];
            const double alpha2 = -0.5 * tau * pvs;
            // A22 -= v w^T + w v^T (stored triangle only), w = p + alpha2 v on the fly.
            // Row a of the stored triangle has a + 1 entries: a thread takes the PAIR of rows
            // (a2, r - 1 - a2), r + 1 entries together, split over ncls2 column classes, so
            // that every thread (and warp) has the same amount of work before the barrier.
            const int tn1 = 2 * n - 1;
            auto update_row = [&](int a, int c, int nc) {
                const int ia = k + 1 + a;
                const double vi = v[ia], wi = w[ia] + alpha2 * vi;
                int j = k + 1 + c;
                if (c == 0)
                {
                    // column k+1 also feeds the norm of the next Householder vector
                    const double vj = v[j], wj = w[j] + alpha2 * vj;
                    double *e0 = T + SA_CJM(j) + ia;
                    const double t = *e0 - (vi * wj + wi * vj);
                    *e0 = t;
                    if (a >= 2)
                        nrm += t * t;
                    j += nc;
                }
                for (; j + nc <= ia; j += 2 * nc)
                {
                    const double v0 = v[j], v1 = v[j + nc];
                    const double w0 = w[j] + alpha2 * v0, w1 = w[j + nc] + alpha2 * v1;
                    double *e0 = T + SA_CJM(j) + ia, *e1 = T + SA_CJM(j + nc) + ia;
                    const double t0 = *e0, t1 = *e1;
                    *e0 = t0 - (vi * w0 + wi * v0);
                    *e1 = t1 - (vi * w1 + wi * v1);
                }
                if (j <= ia)
                {
                    const double v0 = v[j], w0 = w[j] + alpha2 * v0;
                    double *e0 = T + SA_CJM(j) + ia;
                    *e0 = *e0 - (vi * w0 + wi * v0);
                }
            };
            const int rh = (r + 1) >> 1;
            if (ncls2 == 1)
            {
                for (int a2 = tid; a2 < rh; a2 += NT)
                {
                    update_row(a2, 0, 1);
                    if (r - 1 - a2 != a2)
                        update_row(r - 1 - a2, 0, 1);
                }
            }
            else if (my_c2 < ncls2 && my_a2 < rh)
            {
                update_row(my_a2, my_c2, ncls2);
                if (r - 1 - my_a2 != my_a2)
                    update_row(r - 1 - my_a2, my_c2, ncls2);
            }
        }
        else
        {
            for (int a = 2 + tid; a < r; a += NT)
            {
                const double x = T[cjm[k + 1] + (k + 1 + a)];
                nrm += x * x;
            }
        }
        for (int o = 16; o > 0; o >>= 1)
            nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0)
            slots_n[wid] = nrm;
        nslots_n = NT >> 5;
        __syncthreads(); // (4)
    }
    if (tid == 0)
    {
        dd[n - 1] = T[cjm[n - 1] + (n - 1)];
        ee[n - 1] = 0.;
        tt[n - 1] = 0.;
    }
    long long tc3 = clock64();
    if (tid == 0)
        atomicAdd(&g_phase_clk[2], (unsigned long long)(tc3 - tc2));
    for (int q = tid; q < np; q += NT)
        Vout[q] = T[q];
}
