// Host-side plain data types standing in for the MFEM containers the reference
// passes across its hot-path interfaces (mfem::Table, mfem::SparseMatrix,
// mfem::DenseMatrix, mfem::Vector).  MFEM is not part of this tree; only the
// layout and the ordering semantics the reference relies on are kept:
//   Table        CSR int32 I/J                        (used by amg/inc/aggregates.hpp:120-179)
//   SparseMatrix CSR int32 I/J + double data          (AE matrices, P, A)
//   DenseMatrix  column-major double                  (element blocks, cut_evects, MIS blocks)
// Ordering rules chosen for the two Table products (SURVEY.md hard part 1):
//   Transpose -> rows hold ascending column ids
//   Mult      -> rows hold column ids in first-encounter order
#ifndef SAAMGE_B200_SA_TYPES_HPP
#define SAAMGE_B200_SA_TYPES_HPP

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace saamge
{

/* Error convention of the reference: SA_ASSERT prints "ASSERT: file, line: expr"
   and aborts (amg/inc/common.hpp:635-647).  Kept (without MPI_Abort). */
#define SA_ASSERT(expr)                                                        \
    do {                                                                       \
        if (!(expr)) {                                                         \
            std::fprintf(stderr, "ASSERT: %s, %d: %s\n", __FILE__, __LINE__,   \
                         #expr);                                               \
            std::abort();                                                      \
        }                                                                      \
    } while (0)

struct Table
{
    int nrows = 0;
    int ncols = 0;
    std::vector<int> I; // nrows+1
    std::vector<int> J;

    int Size() const { return nrows; }
    int Width() const { return ncols; }
    int RowSize(int i) const { return I[i + 1] - I[i]; }
    const int *GetRow(int i) const { return J.data() + I[i]; }
    int *GetRow(int i) { return J.data() + I[i]; }
    int Size_of_connections() const { return nrows ? I[nrows] : 0; }
    const int *GetI() const { return I.data(); }
    const int *GetJ() const { return J.data(); }
};

/// Common base so providers can hand back either kind of matrix
/// (mfem::Matrix in amg/inc/elmat.hpp:62).
struct Matrix
{
    virtual ~Matrix() {}
};

struct SparseMatrix : public Matrix
{
    int h = 0;
    int w = 0;
    std::vector<int> I;
    std::vector<int> J;
    std::vector<double> A;

    int Size() const { return h; }
    int Height() const { return h; }
    int Width() const { return w; }
    int RowSize(int i) const { return I[i + 1] - I[i]; }
    int NumNonZeroElems() const { return h ? I[h] : 0; }
    const int *GetI() const { return I.data(); }
    const int *GetJ() const { return J.data(); }
    const double *GetData() const { return A.data(); }
    /// Entry lookup; 0 when not in the pattern (mfem::SparseMatrix::operator()).
    double operator()(int i, int j) const
    {
        for (int p = I[i]; p < I[i + 1]; ++p)
            if (J[p] == j)
                return A[p];
        return 0.;
    }
};

struct DenseMatrix : public Matrix
{
    int h = 0;
    int w = 0;
    std::vector<double> d; // column-major, ld = h

    DenseMatrix() {}
    DenseMatrix(int h_, int w_) : h(h_), w(w_), d((size_t)h_ * w_, 0.) {}
    void SetSize(int h_, int w_)
    {
        h = h_;
        w = w_;
        d.assign((size_t)h_ * w_, 0.);
    }
    int Height() const { return h; }
    int Width() const { return w; }
    double *Data() { return d.data(); }
    const double *Data() const { return d.data(); }
    double &operator()(int i, int j) { return d[(size_t)j * h + i]; }
    double operator()(int i, int j) const { return d[(size_t)j * h + i]; }
};

typedef std::vector<double> Vector;

/// At = A^T.  Rows of At list the rows of A in ascending order.
void Transpose(const Table &A, Table &At, int ncols_A = -1);
/// Table from an array "row i -> single column arr[i]" (mfem::Table(int, int*)).
void TableFromArray(const int *arr, int n, int ncols, Table &T);
/// C = A*B (boolean product); row i of C lists columns in first-encounter order
/// while scanning row i of A and, for each entry, the row of B.
void Mult(const Table &A, const Table &B, Table &C);
/// Threads the host-side table products may use (SA_HOST_THREADS; default: hardware threads / ranks
/// of this node, at most 16).
int sa_host_threads();
/// Background helper threads (the topology prefetch, which runs beside the GPU stages and their host
/// work) keep their table products on the calling thread.
void sa_host_threads_serial_here(bool on);

/// y = A x
void SpMult(const SparseMatrix &A, const double *x, double *y);
/// At = A^T, rows ascending.
void SpTranspose(const SparseMatrix &A, SparseMatrix &At);
/// C = A*B, rows of C with ascending columns; entries accumulated in the order
/// "scan row of A, then row of B" (Gustavson).
void SpMultMat(const SparseMatrix &A, const SparseMatrix &B, SparseMatrix &C);

} // namespace saamge

#endif
