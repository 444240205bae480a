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

/// agg_assemble_value (amg/src/aggregates.cpp:68-184): sum over the elements of
/// AE `part` containing both dofs of the element-matrix entry (di, dj).
__device__ __forceinline__ double sa_dev_assemble_value(const LevelTables &L, int di, int dj,
                                                        int part)
{
    double value = 0.;
    const int bi = L.d2e_I[di], ei = L.d2e_I[di + 1];
    for (int p = bi; p < ei; ++p)
    {
        const int elno = L.d2e_J[p];
        if (L.partitioning[elno] != part)
            continue;
        const int eb = L.e2d_I[elno];
        const int ndofs = L.e2d_I[elno + 1] - eb;
        int dii = -1, djj = -1;
        for (int k = 0; k < ndofs; ++k)
        {
            const int g = L.e2d_J[eb + k];
            if (g == di)
                dii = k;
            if (g == dj)
                djj = k;
        }
        if (djj < 0)
            continue; // element does not contain dj
        value += L.elmat[L.elmat_off[elno] + (int64_t)djj * ndofs + dii];
    }
    return value;
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
                // bdr_cond_imposed = assemble_ess_diag = true (amg/src/elmat.cpp:51-52)
                if (both_iface && !(ess && !(glob_neigh == glob_dof)))
                {
                    // the reference assembles (i, j) for i <= j and mirrors the value
                    const double value =
                        (i <= local_neigh) ? sa_dev_assemble_value(L, glob_dof, glob_neigh, part)
                                           : sa_dev_assemble_value(L, glob_neigh, glob_dof, part);
                    tile.at(i, local_neigh) = value;
                }
                else
                    tile.at(i, local_neigh) = L.A_data[p];
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

#endif
