// Flat host-side record of what one level of the hierarchy setup produced.
// Both the product (downloaded from the device on request) and the CPU oracle
// fill this same record, so parity tests compare like with like.  The arrays
// are the flat forms of interp_data_t's per-AE / per-MIS matrices
// (amg/inc/interp.hpp:54-100) and of tg_data_t's operators (amg/inc/tg_data.hpp:47-83).
#ifndef SAAMGE_B200_LEVEL_RESULTS_HPP
#define SAAMGE_B200_LEVEL_RESULTS_HPP

#include "sa_types.hpp"

namespace saamge
{

struct sa_level_results_t
{
    int nparts = 0, num_mises = 0, ND = 0, NDc = 0;
    // per AE: cut_evects_arr[i] (n_i x m_i column-major), its eigenvalues, and
    // rhs_matrices_arr[i] (the weighted-l1 diagonal D, n_i entries)
    std::vector<int> ae_m;
    std::vector<int64_t> ae_eval_off;  // nparts+1
    std::vector<double> evals;
    std::vector<int64_t> ae_evect_off; // nparts+1
    std::vector<double> evects;
    std::vector<double> ae_D;          // offsets = AE_to_dof.I
    // per MIS: mis_tent_interps[mis] (s x k column-major)
    std::vector<int> mis_numcoarsedof;
    std::vector<int64_t> mis_off;      // num_mises+1
    std::vector<double> mis_tent;
    // operators
    SparseMatrix tent_interp, interp, Ac;
    Vector Dinv_neg;
    // coarse element matrices P_e^T A_AE P_e of every AE (only when a coarser
    // level was built from this one), nc_e x nc_e column-major
    std::vector<int64_t> celmat_off;   // nparts+1
    std::vector<double> celmat;
};

struct sa_pcg_results_t
{
    int iterations = 0;           // negative on failure (amg/src/mfem_addons.cpp:201,232)
    std::vector<double> brr;      // (B r, r) history, brr[0] = initial
    Vector x;                     // solution
    double final_res_norm = 0.;   // ||b - A x||_2
};

} // namespace saamge

#endif
