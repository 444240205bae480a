// Table / sparse helpers with MFEM-compatible ordering semantics (see sa_types.hpp).
#include "sa_types.hpp"

#include <omp.h>

#include <algorithm>
#include <thread>

namespace saamge
{

void Transpose(const Table &A, Table &At, int ncols_A)
{
    int nc = ncols_A >= 0 ? ncols_A : A.ncols;
    const int nnz = A.Size_of_connections();
    if (ncols_A < 0 && nc <= 0)
    {
        for (int p = 0; p < nnz; ++p)
            nc = std::max(nc, A.J[p] + 1);
    }
    At.nrows = nc;
    At.ncols = A.nrows;
    At.I.assign((size_t)nc + 1, 0);
    At.J.resize(nnz);
    for (int p = 0; p < nnz; ++p)
        At.I[A.J[p] + 1]++;
    for (int i = 0; i < nc; ++i)
        At.I[i + 1] += At.I[i];
    std::vector<int> pos(At.I.begin(), At.I.end() - 1);
    for (int i = 0; i < A.nrows; ++i)
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            At.J[pos[A.J[p]]++] = i;
}

void TableFromArray(const int *arr, int n, int ncols, Table &T)
{
    T.nrows = n;
    T.ncols = ncols;
    T.I.resize((size_t)n + 1);
    T.J.assign(arr, arr + n);
    for (int i = 0; i <= n; ++i)
        T.I[i] = i;
}

static thread_local bool t_host_serial = false;
void sa_host_threads_serial_here(bool on) { t_host_serial = on; }

int sa_host_threads()
{
    if (t_host_serial)
        return 1;
    // threads for the host-side table products: SA_HOST_THREADS, else the hardware threads divided
    // by the ranks of this node (torchrun exports LOCAL_WORLD_SIZE and sets OMP_NUM_THREADS=1)
    static const int n = [] {
        if (const char *e = getenv("SA_HOST_THREADS"))
            return std::max(1, atoi(e));
        int hw = (int)std::thread::hardware_concurrency();
        if (hw <= 0)
            hw = 1;
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE"))
            ranks = std::max(1, atoi(e));
        return std::max(1, std::min(16, hw / ranks));
    }();
    return n;
}

void Mult(const Table &A, const Table &B, Table &C)
{
    const int nc = B.ncols;
    const int nrows = A.nrows;
    C.nrows = nrows;
    C.ncols = nc;
    C.I.assign((size_t)nrows + 1, 0);
    C.J.clear();
    // rows are independent and keep their first-encounter order: contiguous row blocks per thread
    // (own marker array), concatenated afterwards -- the result is that of the sequential loop
    const int nt = (A.Size_of_connections() < (1 << 16)) ? 1 : std::min(sa_host_threads(), std::max(1, nrows));
    std::vector<std::vector<int>> Jt((size_t)nt);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)(((int64_t)nrows * t) / nt), r1 = (int)(((int64_t)nrows * (t + 1)) / nt);
        std::vector<int> marker((size_t)std::max(nc, 1), -1);
        std::vector<int> &J = Jt[t];
        for (int i = r0; i < r1; ++i)
        {
            const size_t before = J.size();
            for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            {
                const int k = A.J[p];
                for (int q = B.I[k]; q < B.I[k + 1]; ++q)
                {
                    const int j = B.J[q];
                    if (marker[j] != i)
                    {
                        marker[j] = i;
                        J.push_back(j);
                    }
                }
            }
            C.I[i + 1] = (int)(J.size() - before); // row length; prefix sum below
        }
    }
    for (int i = 0; i < nrows; ++i)
        C.I[i + 1] += C.I[i];
    C.J.resize((size_t)C.I[nrows]);
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; ++t)
    {
        const int r0 = (int)(((int64_t)nrows * t) / nt);
        std::copy(Jt[t].begin(), Jt[t].end(), C.J.begin() + C.I[r0]);
    }
}

void SpMult(const SparseMatrix &A, const double *x, double *y)
{
    for (int i = 0; i < A.h; ++i)
    {
        double s = 0.;
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
            s += A.A[p] * x[A.J[p]];
        y[i] = s;
    }
}

void SpTranspose(const SparseMatrix &A, SparseMatrix &At)
{
    const int nnz = A.NumNonZeroElems();
    At.h = A.w;
    At.w = A.h;
    At.I.assign((size_t)A.w + 1, 0);
    At.J.resize(nnz);
    At.A.resize(nnz);
    for (int p = 0; p < nnz; ++p)
        At.I[A.J[p] + 1]++;
    for (int i = 0; i < A.w; ++i)
        At.I[i + 1] += At.I[i];
    std::vector<int> pos(At.I.begin(), At.I.end() - 1);
    for (int i = 0; i < A.h; ++i)
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
        {
            const int q = pos[A.J[p]]++;
            At.J[q] = i;
            At.A[q] = A.A[p];
        }
}

void SpMultMat(const SparseMatrix &A, const SparseMatrix &B, SparseMatrix &C)
{
    SA_ASSERT(A.w == B.h);
    C.h = A.h;
    C.w = B.w;
    C.I.assign((size_t)A.h + 1, 0);
    C.J.clear();
    C.A.clear();
    std::vector<int> marker((size_t)std::max(B.w, 1), -1);
    std::vector<double> acc((size_t)std::max(B.w, 1), 0.);
    std::vector<int> cols;
    for (int i = 0; i < A.h; ++i)
    {
        cols.clear();
        for (int p = A.I[i]; p < A.I[i + 1]; ++p)
        {
            const int k = A.J[p];
            const double a = A.A[p];
            for (int q = B.I[k]; q < B.I[k + 1]; ++q)
            {
                const int j = B.J[q];
                if (marker[j] != i)
                {
                    marker[j] = i;
                    acc[j] = 0.;
                    cols.push_back(j);
                }
                acc[j] += a * B.A[q];
            }
        }
        std::sort(cols.begin(), cols.end());
        for (size_t c = 0; c < cols.size(); ++c)
        {
            C.J.push_back(cols[c]);
            C.A.push_back(acc[cols[c]]);
        }
        C.I[i + 1] = (int)C.J.size();
    }
}

} // namespace saamge
