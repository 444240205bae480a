// Device-side assembly of one AE's dense local matrix, executed cooperatively by
// one thread block.  One warp owns one row of the tile at a time and its lanes own
// distinct entries of that row, so every entry is summed by a single thread in
// ascending element order: deterministic, no atomics, and the same summation order
// as the CPU path.
//   mode with_global = 1: agg_build_AE_stiffm_with_global + agg_assemble_value
//                         (amg/src/aggregates.cpp:855-945, 68-184)
//   mode with_global = 0: agg_build_AE_stiffm (amg/src/aggregates.cpp:959-1086)
#ifndef SA_ASSEMBLE_CUH
#define SA_ASSEMBLE_CUH

#include <algorithm>

#include "sa_gpu_internal.cuh"

#define SA_AGG_BETWEEN_AES_FLAG 0x01
#define SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG 0x02

/// agg_map_id_glob_to_AE (amg/inc/aggregates.hpp:653-673)
__device__ __forceinline__ int sa_dev_map_id_glob_to_AE(const LevelTables &L, int glob_id,
                                                        int part)
{
    const int b = L.d2AE_I[glob_id], e = L.d2AE_I[glob_id + 1];
    for (int p = b; p < e; ++p)
        if (L.d2AE_J[p] == part)
            return L.dof_id_inAE[p];
    return -1;
}

/// Tile addressing: full column-major square, or lower triangle packed by columns
/// (entry (i,j), i >= j, at T[cjm[j] + i] with cjm[j] = j*n - j*(j-1)/2 - j).
struct FullTile
{
    double *T;
    int ld;
    __device__ __forceinline__ bool has(int, int) const { return true; }
    __device__ __forceinline__ double &at(int i, int j) const { return T[i + (int64_t)ld * j]; }
    __device__ __forceinline__ int64_t size(int n) const { return (int64_t)n * ld; }
};
struct PackedTile
{
    double *T;
    const int *cjm;
    __device__ __forceinline__ bool has(int i, int j) const { return i >= j; }
    __device__ __forceinline__ double &at(int i, int j) const { return T[cjm[j] + i]; }
    __device__ __forceinline__ int64_t size(int n) const { return (int64_t)n * (n + 1) / 2; }
};

template <class Tile>
static __device__ void sa_dev_assemble_AE_tile(const LevelTables &L, int part, Tile tile);

/// Fills the n x n column-major tile T (leading dimension ld) with the matrix of AE
/// `part`.  All threads of the block must call; ends with __syncthreads().
/// T may live in shared or global memory.
template <class TP>
static __device__ void sa_dev_assemble_AE(const LevelTables &L, int part, TP T, int ld)
{
    FullTile ft;
    ft.T = T;
    ft.ld = ld;
    sa_dev_assemble_AE_tile(L, part, ft);
}

template <class Tile>
static __device__ void sa_dev_assemble_AE_tile(const LevelTables &L, int part, Tile tile)
{
    const int rb = L.AE2d_I[part];
    const int n = L.AE2d_I[part + 1] - rb;
    const int *dofs = L.AE2d_J + rb;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    {
        const int64_t tot = tile.size(n);
        for (int64_t q = threadIdx.x; q < tot; q += blockDim.x)
            tile.T[q] = 0.;
    }
    __syncthreads();
    if (L.with_global)
    {
        for (int i = wid; i < n; i += nw)
        {
            const int glob_dof = dofs[i];
            const char fi = L.agg_flags[glob_dof];
            // (1) entries taken from the global operator: lanes own the row's nonzeros.
            //     Entries that the reference re-assembles from the AE's elements (both dofs on
            //     an AE interface, and not an off-diagonal touching an essential dof --
            //     bdr_cond_imposed = assemble_ess_diag = true, amg/src/elmat.cpp:51-52) stay
            //     zero here and are summed in (2).
            const int ab = L.A_I[glob_dof], ae = L.A_I[glob_dof + 1];
            for (int p = ab + lane; p < ae; p += 32)
            {
                const int glob_neigh = L.A_J[p];
                const int local_neigh = sa_dev_map_id_glob_to_AE(L, glob_neigh, part);
                if (local_neigh < 0 || !tile.has(i, local_neigh))
                    continue;
                const char fj = L.agg_flags[glob_neigh];
                const bool both_iface =
                    (fi & SA_AGG_BETWEEN_AES_FLAG) && (fj & SA_AGG_BETWEEN_AES_FLAG);
                const bool ess = (fi & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG) ||
                                 (fj & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG);
                if (!(both_iface && !(ess && !(glob_neigh == glob_dof))))
                    tile.at(i, local_neigh) = L.A_data[p];
            }
            if (!(fi & SA_AGG_BETWEEN_AES_FLAG))
                continue;
            __syncwarp();
            // (2) agg_assemble_value for the whole row at once: the AE's elements containing
            //     the row dof, ascending (the order of the reference's sum); the lanes take
            //     the element's columns, which are distinct tile entries.  The reference
            //     evaluates the (smaller local id, larger local id) entry and mirrors it.
            const int bi = L.d2e_I[glob_dof], ei = L.d2e_I[glob_dof + 1];
            for (int p = bi; p < ei; ++p)
            {
                const int elem = L.d2e_J[p];
                if (L.partitioning[elem] != part)
                    continue;
                const int eb = L.e2d_I[elem];
                const int sz = L.e2d_I[elem + 1] - eb;
                int k = -1;
                for (int q = 0; q < sz; ++q)
                    if (L.e2d_J[eb + q] == glob_dof)
                    {
                        k = q;
                        break;
                    }
                const double *Ke = L.elmat + L.elmat_off[elem];
                for (int j = lane; j < sz; j += 32)
                {
                    const int gj = L.e2d_J[eb + j];
                    const char fj = L.agg_flags[gj];
                    if (!(fj & SA_AGG_BETWEEN_AES_FLAG))
                        continue;
                    const bool ess = (fi & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG) ||
                                     (fj & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG);
                    if (ess && gj != glob_dof)
                        continue;
                    const int local_j = sa_dev_map_id_glob_to_AE(L, gj, part);
                    if (!tile.has(i, local_j))
                        continue;
                    tile.at(i, local_j) += (i <= local_j) ? Ke[(int64_t)j * sz + k]
                                                          : Ke[(int64_t)k * sz + j];
                }
                __syncwarp();
            }
        }
    }
    else
    {
        // elements of a row are visited one after the other (ascending); the lanes take
        // the element's columns, which map to distinct tile entries
        for (int i = wid; i < n; i += nw)
        {
            const int g = dofs[i];
            const int bi = L.d2e_I[g], ei = L.d2e_I[g + 1];
            for (int p = bi; p < ei; ++p)
            {
                const int elem = L.d2e_J[p];
                if (L.partitioning[elem] != part)
                    continue;
                const int eb = L.e2d_I[elem];
                const int sz = L.e2d_I[elem + 1] - eb;
                int k = -1;
                for (int q = 0; q < sz; ++q)
                    if (L.e2d_J[eb + q] == g)
                    {
                        k = q;
                        break;
                    }
                const double *Ke = L.elmat + L.elmat_off[elem];
                for (int j = lane; j < sz; j += 32)
                {
                    const double el = Ke[(int64_t)j * sz + k];
                    if (0. != el)
                    {
                        const int local_j = sa_dev_map_id_glob_to_AE(L, L.e2d_J[eb + j], part);
                        if (tile.has(i, local_j))
                            tile.at(i, local_j) += el;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
}

/// Same result as sa_dev_assemble_AE_tile, for AEs whose index structure fits a scratch area
/// in shared memory (the fine-level case: a few dozen small elements).  The generic routine
/// chases d2AE / d2e / e2d through global memory for every entry, one dependent load after
/// the other; here the AE's structure is staged once:
///   - open-addressing hash  global dof -> local id           (replaces agg_map_id_glob_to_AE)
///   - per element of the AE: its dofs as local ids, its block offset
///   - the AE dofs' flags
/// and the assembly itself only reads the operator row / the element blocks from global
/// memory.  Summation order per entry is unchanged (elements ascending).
/// Returns false (uniformly, nothing written) when the structure does not fit; the caller
/// then uses the generic routine.  Ends with __syncthreads().
template <class Tile>
static __device__ bool sa_dev_assemble_AE_staged(const LevelTables &L, int part, Tile tile,
                                                 void *scratch, int scratch_bytes)
{
    const int rb = L.AE2d_I[part];
    const int n = L.AE2d_I[part + 1] - rb;
    const int *dofs = L.AE2d_J + rb;
    const int eb0 = L.AE2e_I[part];
    const int ne = L.AE2e_I[part + 1] - eb0;
    const int *elems = L.AE2e_J + eb0;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
    int H = 32, Hlog = 5;
    while (H < 2 * n)
    {
        H <<= 1;
        ++Hlog;
    }
    if (n > 255 || ne > 1024 ||
        scratch_bytes < 8 * ne + 8 + 4 * H + 8 * ne + 4 + 8 * n + H + n)
        return false;
    long long *eoff = (long long *)scratch;       // ne: element block offsets
    int *hdr = (int *)(eoff + ne);                // [0] total element dofs, [1] largest element
    int *hkey = hdr + 2;                          // H
    int *ebase = hkey + H;                        // ne: e2d_I[elem]
    int *epre = ebase + ne;                       // ne + 1: prefix of the element sizes
    int *arow = epre + ne + 1;                    // 2 n: operator row begin / end of each dof
    unsigned char *hval = (unsigned char *)(arow + 2 * n); // H
    unsigned char *flg = hval + H;                // n
    unsigned char *eloc = flg + n;                // total element dofs
    if (tid < 2)
        hdr[tid] = 0;
    for (int q = tid; q < H; q += NT)
        hkey[q] = -1;
    __syncthreads();
    for (int s = tid; s < ne; s += NT)
    {
        const int e = elems[s];
        const int b = L.e2d_I[e];
        const int sz = L.e2d_I[e + 1] - b;
        ebase[s] = b;
        epre[s + 1] = sz;
        eoff[s] = L.elmat_off[e];
        atomicAdd(&hdr[0], sz);
        atomicMax(&hdr[1], sz);
    }
    for (int i = tid; i < n; i += NT)
    {
        const int key = dofs[i];
        flg[i] = (unsigned char)L.agg_flags[key];
        if (L.with_global)
        {
            arow[2 * i] = L.A_I[key];
            arow[2 * i + 1] = L.A_I[key + 1];
        }
        unsigned int slot = ((unsigned int)key * 0x9E3779B1u) >> (32 - Hlog);
        while (true)
        {
            const int old = atomicCAS(&hkey[slot], -1, key);
            if (old == -1)
            {
                hval[slot] = (unsigned char)i;
                break;
            }
            slot = (slot + 1) & (H - 1);
        }
    }
    __syncthreads();
    const int tot = hdr[0], maxsz = hdr[1];
    const int fixed = (int)((unsigned char *)eloc - (unsigned char *)scratch);
    if (maxsz > 32 || fixed + tot > scratch_bytes)
    {
        __syncthreads();
        return false;
    }
    if (tid == 0)
    {
        epre[0] = 0;
        for (int s = 0; s < ne; ++s)
            epre[s + 1] += epre[s];
    }
    {
        const int64_t tsz = tile.size(n);
        for (int64_t q = tid; q < tsz; q += NT)
            tile.T[q] = 0.;
    }
    __syncthreads();
    auto lookup = [&](int key) -> int {
        unsigned int slot = ((unsigned int)key * 0x9E3779B1u) >> (32 - Hlog);
        while (true)
        {
            const int k = hkey[slot];
            if (k == key)
                return (int)hval[slot];
            if (k == -1)
                return -1;
            slot = (slot + 1) & (H - 1);
        }
    };
    for (int s = wid; s < ne; s += nw)
    {
        const int sz = epre[s + 1] - epre[s];
        if (lane < sz)
            eloc[epre[s] + lane] = (unsigned char)lookup(L.e2d_J[ebase[s] + lane]);
    }
    __syncthreads();
    int gsz = 1;
    while (gsz < maxsz)
        gsz <<= 1;
    const int G = 32 / gsz; // elements scanned per warp step
    const bool wg = L.with_global != 0;
    for (int i = wid; i < n; i += nw)
    {
        const unsigned char fi = flg[i];
        if (wg)
        {
            // entries copied from the global operator (see sa_dev_assemble_AE_tile)
            const int ab = arow[2 * i], ae = arow[2 * i + 1];
            for (int p = ab + lane; p < ae; p += 32)
            {
                const int gj = L.A_J[p];
                const double aval = L.A_data[p]; // issued together with the index load
                const int lj = lookup(gj);
                if (lj < 0 || !tile.has(i, lj))
                    continue;
                const unsigned char fj = flg[lj];
                const bool both_iface =
                    (fi & SA_AGG_BETWEEN_AES_FLAG) && (fj & SA_AGG_BETWEEN_AES_FLAG);
                const bool ess = ((fi | fj) & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG) != 0;
                if (!(both_iface && !(ess && lj != i)))
                    tile.at(i, lj) = aval;
            }
            if (!(fi & SA_AGG_BETWEEN_AES_FLAG))
                continue;
            __syncwarp();
        }
        // the AE's elements containing dof i, ascending: collect up to NB hits (element slot,
        // position of dof i in it), fetch their block entries together, then add in order
        const int NB = 8;
        int nh = 0;
        int my_hit = 0; // lane h keeps hit h as (slot << 8 | k): no dynamically indexed arrays
        auto flush = [&]() {
            double val[NB];
            int ljs[NB];
#pragma unroll
            for (int h = 0; h < NB; ++h)
            {
                val[h] = 0.;
                ljs[h] = -1;
                const int packed = __shfl_sync(0xffffffffu, my_hit, h);
                if (h < nh)
                {
                    const int hs = packed >> 8, k = packed & 255;
                    const int sz = epre[hs + 1] - epre[hs];
                    if (lane < sz)
                    {
                        const int j = lane;
                        const int lj = eloc[epre[hs] + j];
                        if (tile.has(i, lj))
                        {
                            const double *Ke = L.elmat + eoff[hs];
                            if (wg)
                            {
                                const unsigned char fj = flg[lj];
                                const bool ess =
                                    ((fi | fj) & SA_AGG_ON_ESS_DOMAIN_BORDER_FLAG) != 0;
                                if ((fj & SA_AGG_BETWEEN_AES_FLAG) && !(ess && lj != i))
                                {
                                    val[h] = (i <= lj) ? Ke[(int64_t)j * sz + k]
                                                       : Ke[(int64_t)k * sz + j];
                                    ljs[h] = lj;
                                }
                            }
                            else
                            {
                                val[h] = Ke[(int64_t)j * sz + k];
                                if (0. != val[h])
                                    ljs[h] = lj;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < NB; ++h)
            {
                if (h < nh)
                {
                    if (ljs[h] >= 0)
                        tile.at(i, ljs[h]) += val[h];
                    __syncwarp();
                }
            }
            nh = 0;
        };
        const bool uniform = tot == ne * gsz; // all elements have gsz dofs
        for (int s0 = 0; s0 < ne; s0 += G)
        {
            bool hit = false;
            if (uniform)
            {
                const int idx = s0 * gsz + lane;
                hit = idx < tot && eloc[idx] == i;
            }
            else
            {
                const int s = s0 + lane / gsz, q = lane % gsz;
                if (s < ne && q < epre[s + 1] - epre[s])
                    hit = eloc[epre[s] + q] == i;
            }
            unsigned int ballot = __ballot_sync(0xffffffffu, hit);
            while (ballot)
            {
                const int b = __ffs(ballot) - 1;
                ballot &= ballot - 1;
                if (lane == nh)
                    my_hit = ((s0 + b / gsz) << 8) | (b % gsz);
                if (++nh == NB)
                    flush();
            }
        }
        if (nh)
            flush();
    }
    __syncthreads();
    return true;
}

/// Column-wise assembly of LARGE AE matrices (coarse levels: a few thousand dofs, elements
/// with a hundred dofs or more) straight into a column-major n x n matrix in global memory;
/// agg_build_AE_stiffm (amg/src/aggregates.cpp:959-1086) only (no operator rows).
/// grid = (column splits, matrices), one warp per column at a time:
///   - the block stages an open-addressing hash global dof -> local id in shared memory;
///   - a column is summed in a shared-memory buffer of n doubles (scatter by local row id),
///     over the AE's elements containing the column dof in ascending order (deterministic,
///     the reference's order); the element block column is read contiguously;
///   - the finished column is written out coalesced.
/// Shared memory: 2 * H ints (H = power of two >= 2 n) + warps * n doubles.
/// parts[m]: AE of matrix m; matrix m at Tbase + m * tstride (or Tbase + toffs[m]).
static __global__ void k_assemble_large(LevelTables L, const int *parts, const int *slot_list,
                                        const int *ae_of_slot, double *Tbase, int64_t tstride,
                                        const int64_t *toffs, int Hlog)
{
    extern __shared__ double smem_al[];
    const int m = blockIdx.y;
    const int part = parts ? parts[m] : ae_of_slot[slot_list[m]];
    const int rb = L.AE2d_I[part];
    const int n = L.AE2d_I[part + 1] - rb;
    const int *dofs = L.AE2d_J + rb;
    const int H = 1 << Hlog;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    int *hkey = (int *)smem_al;
    int *hval = hkey + H;
    double *buf = (double *)(hval + H) + (size_t)wid * n;
    double *T = Tbase + (toffs ? toffs[m] : (int64_t)m * tstride);
    for (int q = tid; q < H; q += NT)
        hkey[q] = -1;
    __syncthreads();
    for (int i = tid; i < n; i += NT)
    {
        const int key = dofs[i];
        unsigned int slot = ((unsigned int)key * 0x9E3779B1u) >> (32 - Hlog);
        while (true)
        {
            const int old = atomicCAS(&hkey[slot], -1, key);
            if (old == -1)
            {
                hval[slot] = i;
                break;
            }
            slot = (slot + 1) & (H - 1);
        }
    }
    __syncthreads();
    for (int col = blockIdx.x * NW + wid; col < n; col += gridDim.x * NW)
    {
        for (int i = lane; i < n; i += 32)
            buf[i] = 0.;
        __syncwarp();
        const int g = dofs[col];
        const int bi = L.d2e_I[g], ei = L.d2e_I[g + 1];
        for (int p = bi; p < ei; ++p)
        {
            const int elem = L.d2e_J[p];
            if (L.partitioning[elem] != part)
                continue;
            const int eb = L.e2d_I[elem];
            const int sz = L.e2d_I[elem + 1] - eb;
            int kc = -1; // position of the column dof in the element
            for (int q0 = 0; q0 < sz && kc < 0; q0 += 32)
            {
                const bool hit = (q0 + lane < sz) && L.e2d_J[eb + q0 + lane] == g;
                const unsigned int b = __ballot_sync(0xffffffffu, hit);
                if (b)
                    kc = q0 + __ffs(b) - 1;
            }
            const double *Kc = L.elmat + L.elmat_off[elem] + (int64_t)kc * sz; // column kc
            for (int jr = lane; jr < sz; jr += 32)
            {
                const double el = Kc[jr];
                if (0. != el)
                {
                    const int key = L.e2d_J[eb + jr];
                    unsigned int slot = ((unsigned int)key * 0x9E3779B1u) >> (32 - Hlog);
                    while (hkey[slot] != key)
                        slot = (slot + 1) & (H - 1);
                    buf[hval[slot]] += el; // distinct dofs of one element: distinct rows
                }
            }
            __syncwarp();
        }
        double *Tc = T + (int64_t)n * col;
        for (int i = lane; i < n; i += 32)
            Tc[i] = buf[i];
        __syncwarp();
    }
}

/// Launches k_assemble_large for nmat matrices (largest: nmax dofs) on \a stream; returns false
/// (nothing launched) when the shared-memory footprint does not fit or the level assembles
/// with the global operator.
static inline bool sa_launch_assemble_large(sa_gpu_ctx *ctx, const LevelTables &L, const int *d_parts,
                                            const int *d_slot_list, const int *d_ae_of_slot,
                                            int nmat, int nmax, double *Tbase, int64_t tstride,
                                            const int64_t *d_toffs, cudaStream_t stream)
{
    if (L.with_global || nmat <= 0)
        return false;
    int Hlog = 5;
    while ((1 << Hlog) < 2 * nmax)
        ++Hlog;
    int NW = 4;
    auto smem_of = [&](int nw) {
        return (size_t)2 * ((size_t)1 << Hlog) * sizeof(int) + (size_t)nw * nmax * sizeof(double);
    };
    while (NW > 1 && smem_of(NW) > ctx->smem_optin)
        NW >>= 1;
    const size_t smem = smem_of(NW);
    if (smem > ctx->smem_optin)
        return false;
    const int nsplit = std::max(1, std::min((nmax + NW - 1) / NW, (2 * ctx->num_sms + nmat - 1) / nmat));
    SA_CUDA(cudaFuncSetAttribute(k_assemble_large, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    k_assemble_large<<<dim3(nsplit, nmat), NW * 32, smem, stream>>>(L, d_parts, d_slot_list,
                                                                     d_ae_of_slot, Tbase, tstride, d_toffs, Hlog);
    ctx->launches++;
    SA_CUDA(cudaGetLastError());
    return true;
}

#endif
