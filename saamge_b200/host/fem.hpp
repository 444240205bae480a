// Host input producer: structured quad/hex H1 (Q1/Q2) diffusion problems.
// Stands in for the MFEM side of the reference (mesh, ParBilinearForm,
// DiffusionIntegrator, BC elimination; amg/inc/fem.hpp:427-451,
// amg/src/fem.cpp:87-140, 686-717), which stays on the host as *input* to the
// hot path (BASELINE.json north_star).  Not a GPU target.
#ifndef SAAMGE_B200_FEM_HPP
#define SAAMGE_B200_FEM_HPP

#include "sa_types.hpp"

namespace saamge
{

typedef char agg_dof_status_t; // amg/inc/aggregates.hpp:116

enum fem_coef_kind
{
    FEM_COEF_CONSTANT = 0,  // k == 1 (Poisson; BASELINE config C1)
    FEM_COEF_LOGNORMAL = 1, // exp(s*g), g smoothed Gaussian, max/min = contrast
    FEM_COEF_CHECKER = 2,   // checkerboard 1 / contrast on 4^d blocks
    FEM_COEF_MLTEST = 3     // checkboard_coef of the mltest driver evaluated at element
                            // centres (amg/test/mltest/mltest.cpp:156-175): 1e6 / 1
};

struct fem_problem_t
{
    int dim = 0, order = 1;
    int nx = 0, ny = 0, nz = 0;
    int NE = 0, ND = 0;
    int ne = 0;                   // dofs per element
    Table elem_to_dof;            // NE x ND
    Table elem_to_elem;           // face neighbours, no self loops
    std::vector<double> coef;     // per-element coefficient
    std::vector<double> elmat;    // NE blocks ne*ne, column-major, NOT BC-eliminated
    SparseMatrix A;               // assembled, essential BC eliminated keeping the diagonal
    Vector b;                     // load vector for f == 1, eliminated
    std::vector<agg_dof_status_t> bdr_dofs; // AGG_ON_ESS_DOMAIN_BORDER_FLAG on essential dofs
};

/// Unit square/cube, nx x ny (x nz) cells, homogeneous Dirichlet on the whole
/// boundary (reference default ess_bdr = 1, amg/test/mltest/mltest.cpp:491-494).
fem_problem_t *fem_generate_structured(int dim, int nx, int ny, int nz,
                                       int order, int coef_kind,
                                       double contrast, uint64_t seed);
/// Same with a choice of essential sides: bit 0: x=0, 1: x=1, 2: y=0, 3: y=1, 4: z=0,
/// 5: z=1 (the mltest fixture marks only x=0, amg/test/mltest/mltest.cpp:476-479).
fem_problem_t *fem_generate_structured_ex(int dim, int nx, int ny, int nz, int order,
                                          int coef_kind, double contrast, uint64_t seed,
                                          int ess_mask);

} // namespace saamge

#endif
