// Bench driver: times the local spectral stage (a2-a7) of the finest level through the
// C ABI, either with inputs already resident on the device ("value") or end to end
// from host buffers with the H2D / D2H copies inside the timed region ("e2e").
// Used by bench.py only; see include/saamge_b200_driver.h.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "hierarchy.hpp"
#include "saamge.hpp"

extern "C" {
double sa_gpu_level_uploaded_bytes(sa_gpu_level *level);
int sa_gpu_host_register(const void *p, size_t bytes);
int sa_gpu_host_unregister(const void *p);
int sa_gpu_debug_phase_clocks(double *out8);
}

using namespace saamge;

struct sa_bench_t
{
    sa_problem_t *prob = NULL;
    sa_drv_params_t p;
    sa_gpu_ctx *ctx = NULL;
    sa_gpu_level *lev = NULL; // resident level
    sa_gpu_level_desc desc;
    std::vector<int64_t> offsets;
    std::vector<int> ae_m;
    std::vector<double> evals, evects;
    double h2d_bytes = 0., d2h_bytes = 0., h2d_bytes_last = 0., table_bytes = 0.;
    bool pinned = false;
    std::vector<const void *> registered;
    double e2e_phase_ms[3] = {0., 0., 0.}; // upload, compute, read back (last mode-1 step)
};

static bool pin(sa_bench_t *B, const void *p, size_t bytes)
{
    if (!p || bytes == 0)
        return true;
    if (sa_gpu_host_register(p, bytes) != 0)
        return false;
    B->registered.push_back(p);
    return true;
}

static double desc_bytes(const sa_gpu_level_desc &d)
{
    double b = 0.;
    b += 4. * (d.NE + 1 + d.elem_to_dof_I[d.NE]);
    b += 4. * (d.ND + 1 + d.dof_to_elem_I[d.ND]);
    b += 4. * (d.nparts + 1 + d.AE_to_elem_I[d.nparts]);
    b += 4. * (d.nparts + 1 + d.AE_to_dof_I[d.nparts]);
    b += 4. * (d.ND + 1 + 2. * d.dof_to_AE_I[d.ND]);
    b += 4. * d.NE + 1. * d.ND;
    b += 4. * (d.num_mises + 1 + d.mis_to_dof_I[d.num_mises]);
    b += 4. * (d.num_mises + 1 + d.mis_to_AE_I[d.num_mises]);
    b += 4. * (d.nparts + 1 + d.AE_to_mis_I[d.nparts]);
    b += 4. * d.ND;
    if (d.A_I)
        b += 4. * (d.ND + 1) + 12. * d.A_I[d.ND];
    if (d.elmat)
        b += 8. * (d.NE + 1) + 8. * d.elmat_off[d.NE];
    return b;
}

extern "C" void *sa_drv_bench_create(void *prob_, const sa_drv_params_t *p, int device)
{
    sa_problem_t *prob = (sa_problem_t *)prob_;
    SA_ASSERT(prob && prob->rels);
    sa_bench_t *B = new sa_bench_t;
    B->prob = prob;
    B->p = *p;
    B->ctx = proc_gpu_init(device);
    const fem_problem_t &f = *prob->fem;
    const agg_partitioning_relations_t &r = *prob->rels;
    B->offsets.resize((size_t)f.NE + 1);
    for (int e = 0; e <= f.NE; ++e)
        B->offsets[e] = (int64_t)e * f.ne * f.ne;
    sa_gpu_level_desc &d = B->desc;
    std::memset(&d, 0, sizeof d);
    d.ND = r.ND;
    d.NE = r.elem_to_dof->Size();
    d.nparts = r.nparts;
    d.num_mises = r.num_mises;
    d.elem_to_dof_I = r.elem_to_dof->GetI();
    d.elem_to_dof_J = r.elem_to_dof->GetJ();
    d.dof_to_elem_I = r.dof_to_elem->GetI();
    d.dof_to_elem_J = r.dof_to_elem->GetJ();
    d.AE_to_elem_I = r.AE_to_elem->GetI();
    d.AE_to_elem_J = r.AE_to_elem->GetJ();
    d.AE_to_dof_I = r.AE_to_dof->GetI();
    d.AE_to_dof_J = r.AE_to_dof->GetJ();
    d.dof_to_AE_I = r.dof_to_AE->GetI();
    d.dof_to_AE_J = r.dof_to_AE->GetJ();
    d.dof_id_inAE = r.dof_id_inAE;
    d.partitioning = r.partitioning;
    d.agg_flags = r.agg_flags;
    d.mis_to_dof_I = r.mis_to_dof->GetI();
    d.mis_to_dof_J = r.mis_to_dof->GetJ();
    d.mis_to_AE_I = r.mis_to_AE->GetI();
    d.mis_to_AE_J = r.mis_to_AE->GetJ();
    d.AE_to_mis_I = r.AE_to_mis->GetI();
    d.AE_to_mis_J = r.AE_to_mis->GetJ();
    d.mises = r.mises;
    d.A_I = f.A.GetI();
    d.A_J = f.A.GetJ();
    d.A_data = f.A.GetData();
    d.elmat = f.elmat.data();
    d.elmat_off = B->offsets.data();
    d.assemble_with_global = 1;
    B->h2d_bytes = desc_bytes(d);
    B->table_bytes = B->h2d_bytes - (d.A_I ? 12. * d.A_I[d.ND] : 0.) - (d.elmat ? 8. * d.elmat_off[d.NE] : 0.);
    // pin the large host arrays so the e2e copies run at full PCIe speed
    // (every array the level upload reads: element blocks, operator, relation tables)
    bool ok = pin(B, f.elmat.data(), f.elmat.size() * sizeof(double));
    ok = pin(B, f.A.GetData(), f.A.A.size() * sizeof(double)) && ok;
    ok = pin(B, f.A.GetJ(), f.A.J.size() * sizeof(int)) && ok;
    ok = pin(B, f.A.GetI(), ((size_t)d.ND + 1) * sizeof(int)) && ok;
    ok = pin(B, B->offsets.data(), B->offsets.size() * sizeof(int64_t)) && ok;
    const struct { const int *I, *J; int n; } tabs[] = {
        {d.elem_to_dof_I, d.elem_to_dof_J, d.NE},   {d.dof_to_elem_I, d.dof_to_elem_J, d.ND},
        {d.AE_to_elem_I, d.AE_to_elem_J, d.nparts}, {d.AE_to_dof_I, d.AE_to_dof_J, d.nparts},
        {d.dof_to_AE_I, d.dof_to_AE_J, d.ND},       {d.mis_to_dof_I, d.mis_to_dof_J, d.num_mises},
        {d.mis_to_AE_I, d.mis_to_AE_J, d.num_mises}, {d.AE_to_mis_I, d.AE_to_mis_J, d.nparts}};
    for (const auto &t : tabs)
    {
        ok = pin(B, t.I, ((size_t)t.n + 1) * sizeof(int)) && ok;
        ok = pin(B, t.J, (size_t)t.I[t.n] * sizeof(int)) && ok;
    }
    ok = pin(B, d.dof_id_inAE, (size_t)d.dof_to_AE_I[d.ND] * sizeof(int)) && ok;
    ok = pin(B, d.partitioning, (size_t)d.NE * sizeof(int)) && ok;
    ok = pin(B, d.agg_flags, (size_t)d.ND) && ok;
    ok = pin(B, d.mises, (size_t)d.ND * sizeof(int)) && ok;
    B->pinned = ok;
    sa_gpu_check(sa_gpu_level_create(B->ctx, &d, NULL, &B->lev), "sa_gpu_level_create");
    B->ae_m.resize(r.nparts);
    return B;
}

extern "C" void *sa_drv_bench_level(void *b_) { return ((sa_bench_t *)b_)->lev; }

extern "C" void sa_drv_bench_destroy(void *b_)
{
    sa_bench_t *B = (sa_bench_t *)b_;
    if (!B)
        return;
    for (const void *p : B->registered)
        sa_gpu_host_unregister(p);
    sa_gpu_level_destroy(B->lev);
    delete B;
}

/* mode 0: device-resident inputs, AEs [ae_begin, ae_end); returns device ms
   mode 1: end to end from host buffers: upload, compute, read back m / lambda / vectors
   mode 2: as 1 for one rank of a sharded stage: only the inputs AEs [ae_begin, ae_end) read are
           uploaded (desc.async_upload = 2) and only their results are read back */
extern "C" double sa_drv_bench_step(void *b_, int mode, int ae_begin, int ae_end)
{
    sa_bench_t *B = (sa_bench_t *)b_;
    const double theta = B->p.first_theta;
    if (mode == 0)
    {
        sa_gpu_ctx_timer(B->ctx, 1);
        sa_gpu_check(sa_gpu_local_spectral(B->lev, theta, ae_begin, ae_end, 0),
                     "sa_gpu_local_spectral");
        return sa_gpu_ctx_timer(B->ctx, 0);
    }
    // SA_BENCH_BREAKDOWN=1 synchronises between the three phases to time them separately
    // (diagnostic only: it removes the overlap the normal path has)
    static const bool breakdown = getenv("SA_BENCH_BREAKDOWN") != NULL;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    sa_gpu_ctx_timer(B->ctx, 1);
    auto t0 = now();
    sa_gpu_level *lev = NULL;
    // pipelined upload: the eigen stage starts on the first AEs while the rest of the
    // operator / element blocks is still in flight (SA_BENCH_SYNC_UPLOAD=1 turns it off)
    static const bool sync_upload = getenv("SA_BENCH_SYNC_UPLOAD") != NULL;
    sa_gpu_level_desc desc = B->desc;
    desc.async_upload = (breakdown || sync_upload) ? 0 : (mode == 2 ? 2 : 1);
    sa_gpu_check(sa_gpu_level_create(B->ctx, &desc, NULL, &lev), "sa_gpu_level_create");
    if (breakdown)
    {
        sa_gpu_ctx_sync(B->ctx);
        B->e2e_phase_ms[0] = ms_since(t0);
        t0 = now();
    }
    sa_gpu_check(sa_gpu_local_spectral(lev, theta, ae_begin, ae_end, 0), "sa_gpu_local_spectral");
    if (breakdown)
    {
        sa_gpu_ctx_sync(B->ctx);
        B->e2e_phase_ms[1] = ms_since(t0);
        t0 = now();
    }
    sa_gpu_check(sa_gpu_get_spectral_counts(lev, B->ae_m.data()), "sa_gpu_get_spectral_counts");
    const agg_partitioning_relations_t &r = *B->prob->rels;
    size_t ne = 0, nv = 0;
    for (int i = 0; i < r.nparts; ++i)
    {
        if (mode == 2 && (i < ae_begin || i >= ae_end))
            B->ae_m[i] = 0; // (counts outside the range are not this rank's)
        ne += B->ae_m[i];
        nv += (size_t)B->ae_m[i] * r.AE_to_dof->RowSize(i);
    }
    // the caller's result buffers are page-locked like its inputs (pageable memory made this read-back
    // 5 ms of the 63 ms step); they grow by 1.5x, so the warm-up steps do the registering
    auto grow_pinned = [&](std::vector<double> &v, size_t n) {
        if (n > v.capacity())
        {
            if (B->pinned && v.capacity())
            {
                sa_gpu_host_unregister(v.data());
                B->registered.erase(std::remove(B->registered.begin(), B->registered.end(), (const void *)v.data()),
                                    B->registered.end());
            }
            std::vector<double>().swap(v);
            v.reserve(n + n / 2 + 1024);
            if (B->pinned)
                pin(B, v.data(), v.capacity() * sizeof(double));
        }
        v.resize(n);
    };
    grow_pinned(B->evals, ne);
    grow_pinned(B->evects, nv);
    sa_gpu_check(sa_gpu_get_spectral(lev, B->evals.data(), B->evects.data(), NULL),
                 "sa_gpu_get_spectral");
    if (breakdown)
        B->e2e_phase_ms[2] = ms_since(t0);
    const double ms = sa_gpu_ctx_timer(B->ctx, 0);
    B->d2h_bytes = 8. * (ne + nv) + 4. * r.nparts;
    if (mode == 2)
        B->h2d_bytes_last = sa_gpu_level_uploaded_bytes(lev) + B->table_bytes;
    sa_gpu_level_destroy(lev);
    return ms;
}

extern "C" double sa_drv_bench_scalar(void *b_, const char *name_)
{
    sa_bench_t *B = (sa_bench_t *)b_;
    const std::string name(name_);
    const agg_partitioning_relations_t &r = *B->prob->rels;
    if (name == "h2d_bytes") return B->h2d_bytes;
    if (name == "h2d_bytes_last") return B->h2d_bytes_last;
    if (name == "d2h_bytes") return B->d2h_bytes;
    if (name == "pinned") return B->pinned ? 1. : 0.;
    if (name == "e2e.upload_ms") return B->e2e_phase_ms[0];
    if (name == "e2e.compute_ms") return B->e2e_phase_ms[1];
    if (name == "e2e.readback_ms") return B->e2e_phase_ms[2];
    if (name == "launches") return (double)sa_gpu_ctx_launch_count(B->ctx);
    if (name == "flops" || name == "bytes" || name == "sum_m")
    {
        // algorithmic work of the eigen stage:
        //   F = 4/3 n^3 + 2 n^2 m + 3 n^2 ; bytes = 8 n^2 + 8 n m + 8 m
        sa_gpu_check(sa_gpu_get_spectral_counts(B->lev, B->ae_m.data()),
                     "sa_gpu_get_spectral_counts");
        double F = 0., Bt = 0., M = 0.;
        for (int i = 0; i < r.nparts; ++i)
        {
            const double n = r.AE_to_dof->RowSize(i), m = B->ae_m[i];
            F += 4. / 3. * n * n * n + 2. * n * n * m + 3. * n * n;
            Bt += 8. * n * n + 8. * n * m + 8. * m;
            M += m;
        }
        return name == "flops" ? F : (name == "bytes" ? Bt : M);
    }
    if (name.compare(0, 6, "phase.") == 0)
    {
        static double clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int idx = name[6] - '0';
        if (idx == 0)
            sa_gpu_debug_phase_clocks(clk);
        return (idx >= 0 && idx < 8) ? clk[idx] : NAN;
    }
    return NAN;
}

extern "C" int sa_drv_gpu_profile(int enable, char *buf, int buflen)
{
    return sa_gpu_ctx_profile(proc_gpu_ctx(), enable, buf, buflen);
}

extern "C" void *sa_drv_ctx(void) { return proc_gpu_ctx(); }
/* the process's context on `device` (created when there is none yet) */
extern "C" void *sa_drv_ctx_on(int device) { return proc_gpu_init(device); }

/* kind 0: SpMV with the finest operator, kind 1: fused polynomial-smoother step;
   returns milliseconds per call (device-resident vectors, CUDA events) */
extern "C" double sa_drv_ml_spmv_bench(void *hier, int kind, int reps)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    struct impl_t
    {
        ml_data_t *ml;
    };
    ml_data_t *ml = ((impl_t *)H->impl)->ml;
    sa_gpu_level *lev = ml->levels_list.finest->tg_data->gpu;
    return kind == 0 ? sa_gpu_bench_spmv(lev, SA_GPU_MAT_A, reps) : sa_gpu_bench_smoother(lev, reps);
}

/* device handles of a built hierarchy (for the row-partitioned multi-GPU solve driven from
   Python: saamge_b200/dist_solve.py) */
extern "C" void *sa_drv_ml_gpu_level(void *hier, int level)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    struct impl_t
    {
        ml_data_t *ml;
    };
    ml_data_t *ml = ((impl_t *)H->impl)->ml;
    levels_level_t *l = levels_list_get_level(ml->levels_list, level);
    return l ? (void *)l->tg_data->gpu : NULL;
}

extern "C" void *sa_drv_ml_gpu_solver(void *hier)
{
    sa_hierarchy_t *H = (sa_hierarchy_t *)hier;
    struct impl_t
    {
        ml_data_t *ml;
    };
    return (void *)((impl_t *)H->impl)->ml->gpu_solver;
}
