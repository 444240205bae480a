// Element-matrix plug interface -- the operator/plugin API through which a
// user feeds element matrices to the hierarchy builder.  Same shape as
// class ElementMatrixProvider (amg/inc/elmat.hpp:53-77): GetMatrix returns the
// matrix of element elno and tells the caller whether to free it;
// BuildAEStiff returns the assembled matrix of AE elno (caller owns).
//
// Product-side concrete providers only *describe* where the element blocks
// live; the batched device kernels do the arithmetic (sa_gpu_local_spectral).
// The two single-item virtuals are kept for API compatibility and run the
// same device path on a batch of one.
#ifndef SAAMGE_B200_ELMAT_HPP
#define SAAMGE_B200_ELMAT_HPP

#include "aggregates.hpp"
#include "sa_types.hpp"

namespace saamge
{

class ElementMatrixProvider
{
public:
    ElementMatrixProvider(const agg_partitioning_relations_t &agg_part_rels)
        : agg_part_rels(agg_part_rels), is_geometric(false)
    {
    }
    virtual ~ElementMatrixProvider() {}
    virtual Matrix *GetMatrix(int elno, bool &free_matr) const = 0;
    virtual SparseMatrix *BuildAEStiff(int elno) const = 0;
    bool IsGeometric() const { return is_geometric; }

    /* batched view used by the device path */
    /// contiguous dense element blocks (column-major) and their offsets, or NULL
    virtual const double *DenseBlocks() const { return NULL; }
    virtual const int64_t *DenseBlockOffsets() const { return NULL; }
    /// BC-eliminated assembled matrix (fine level "with global" assembly), or NULL
    virtual const SparseMatrix *AssembledMatrix() const { return NULL; }

protected:
    const agg_partitioning_relations_t &agg_part_rels;
    bool is_geometric;
};

} // namespace saamge

#endif
