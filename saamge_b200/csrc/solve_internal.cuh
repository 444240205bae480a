// Solver state shared by solve.cu (single GPU) and dist.cu (row-partitioned, NCCL).
#pragma once
#include <vector>

#include "sa_gpu_internal.cuh"

struct SolverLevel
{
    sa_gpu_level *lev = nullptr;
    DevBuf<double> b, xa, xb, r; // rhs, ping/pong iterate, residual (size ND of the level)
    // user smoothers (smpr_ft plug, amg/inc/smpr.hpp:59-60): host callbacks; NULL = the fused
    // device SAS polynomial smoother
    sa_gpu_smoother_ft pre = nullptr, post = nullptr;
    void *smoother_data = nullptr;
    std::vector<double> hb, hx;
};

struct sa_gpu_solver
{
    sa_gpu_ctx *ctx = nullptr;
    std::vector<SolverLevel *> L;
    int degree = 0;
    std::vector<double> roots;
    // coarsest
    int nc = 0;
    DevBuf<double> Ainv; // dense nc x nc
    DevBuf<double> bc, xc;
    // PCG work vectors on level 0
    DevBuf<double> pb, px, pr, pd, pz;
    DevBuf<double> dots; // device scalars
    DevBuf<double> dot_partials; // per-block partial sums of k_dot (fixed summation order)
    ~sa_gpu_solver()
    {
        for (size_t i = 0; i < L.size(); ++i)
            delete L[i];
    }
};

