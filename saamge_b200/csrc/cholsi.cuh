// Cholesky + shift-invert subspace iteration for the lower eigenpairs of large AE matrices
// (cholsi.cu).
#pragma once
#include "sa_gpu_internal.cuh"

#define SA_CS_K 8 /* vectors of the subspace iteration: at most SA_CS_K - 1 eigenvalues <= theta */

struct sa_cs_mat
{
    int n;
    int slot;    // seed of the start vectors (the AE id: results independent of the batching)
    double *T;   // n x n column-major: scaled matrix on entry, L in the lower triangle on exit
    double *X;   // n x SA_CS_K: Ritz vectors (unit 2-norm, ascending Ritz value) on exit
    double *X2;  // n x SA_CS_K: work (ping-pong partner of X)
    double *Z;   // n x SA_CS_K: work
    double *small; // 160 doubles: Gram matrices and residuals of the Rayleigh-Ritz step
    double *lam; // SA_CS_K Ritz values
    int *info;   // [0]: number of eigenvalues <= theta (>= 0), or -1 non-positive pivot, -2 all
                 //      SA_CS_K Ritz values <= theta, -3 no convergence, -4 rank-deficient block;
                 // [1]: iterations
};

double sa_cs_sigma(double theta);
void sa_cs_factor_iterate(sa_gpu_ctx *ctx, const sa_cs_mat *d_mats, int nmats, int nmax, double theta,
                          cudaStream_t st);
void sa_cs_gather(sa_gpu_ctx *ctx, const sa_cs_mat *d_mats, int nmats, const int *d_nev,
                  const int64_t *d_eval_off, const int64_t *d_evect_off, const double *d_sinv,
                  const int *d_doff, double *evals, double *evects, double theta, int *borderline,
                  cudaStream_t st);
