// Structured-grid H1 diffusion input generator (see fem.hpp).
#include "fem.hpp"

#include <algorithm>
#include <cmath>
#include <random>

namespace saamge
{

namespace
{

// 1D Lagrange basis on equispaced nodes of [0,1] (order p), value and derivative.
void lagrange1d(int p, double x, double *val, double *der)
{
    const int n = p + 1;
    for (int i = 0; i < n; ++i)
    {
        const double xi = (double)i / p;
        double v = 1., d = 0.;
        for (int j = 0; j < n; ++j)
        {
            if (j == i)
                continue;
            const double xj = (double)j / p;
            v *= (x - xj) / (xi - xj);
        }
        for (int k = 0; k < n; ++k)
        {
            if (k == i)
                continue;
            const double xk = (double)k / p;
            double t = 1. / (xi - xk);
            for (int j = 0; j < n; ++j)
            {
                if (j == i || j == k)
                    continue;
                const double xj = (double)j / p;
                t *= (x - xj) / (xi - xj);
            }
            d += t;
        }
        val[i] = v;
        der[i] = d;
    }
}

// 1D mass M, stiffness S and load L on [0,1]; 4-point Gauss (exact to degree 7).
void matrices1d(int p, double *M, double *S, double *L)
{
    static const double gx[4] = {0.06943184420297371, 0.33000947820757187,
                                 0.66999052179242813, 0.93056815579702629};
    static const double gw[4] = {0.17392742256872693, 0.32607257743127307,
                                 0.32607257743127307, 0.17392742256872693};
    const int n = p + 1;
    std::fill(M, M + n * n, 0.);
    std::fill(S, S + n * n, 0.);
    std::fill(L, L + n, 0.);
    double v[4], d[4];
    for (int q = 0; q < 4; ++q)
    {
        lagrange1d(p, gx[q], v, d);
        for (int i = 0; i < n; ++i)
        {
            L[i] += gw[q] * v[i];
            for (int j = 0; j < n; ++j)
            {
                M[i * n + j] += gw[q] * v[i] * v[j];
                S[i * n + j] += gw[q] * d[i] * d[j];
            }
        }
    }
}

void box_filter_axis(std::vector<double> &g, int nx, int ny, int nz, int axis,
                     int radius)
{
    std::vector<double> out(g.size());
    const int n[3] = {nx, ny, nz};
    const int stride[3] = {1, nx, nx * ny};
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i)
            {
                const int idx[3] = {i, j, k};
                double s = 0.;
                int cnt = 0;
                for (int r = -radius; r <= radius; ++r)
                {
                    const int c = idx[axis] + r;
                    if (c < 0 || c >= n[axis])
                        continue;
                    s += g[(size_t)i + (size_t)nx * (j + (size_t)ny * k) +
                           (int64_t)r * stride[axis]];
                    ++cnt;
                }
                out[(size_t)i + (size_t)nx * (j + (size_t)ny * k)] = s / cnt;
            }
    g.swap(out);
}

} // namespace

fem_problem_t *fem_generate_structured(int dim, int nx, int ny, int nz,
                                       int order, int coef_kind,
                                       double contrast, uint64_t seed)
{
    return fem_generate_structured_ex(dim, nx, ny, nz, order, coef_kind, contrast, seed, 63);
}

fem_problem_t *fem_generate_structured_ex(int dim, int nx, int ny, int nz, int order,
                                          int coef_kind, double contrast, uint64_t seed,
                                          int ess_mask)
{
    SA_ASSERT(dim == 2 || dim == 3);
    SA_ASSERT(order == 1 || order == 2);
    if (dim == 2)
        nz = 1;
    fem_problem_t *fp = new fem_problem_t;
    fp->dim = dim;
    fp->order = order;
    fp->nx = nx;
    fp->ny = ny;
    fp->nz = nz;
    const int p = order;
    const int n1 = p + 1;
    const int ne = (dim == 2) ? n1 * n1 : n1 * n1 * n1;
    fp->ne = ne;
    const int NE = nx * ny * nz;
    fp->NE = NE;
    const int gx = p * nx + 1, gy = p * ny + 1, gz = (dim == 3) ? p * nz + 1 : 1;
    const int ND = gx * gy * gz;
    fp->ND = ND;

    // elem_to_dof: lexicographic nodes, x fastest, local order lexicographic.
    Table &e2d = fp->elem_to_dof;
    e2d.nrows = NE;
    e2d.ncols = ND;
    e2d.I.resize((size_t)NE + 1);
    e2d.J.resize((size_t)NE * ne);
    for (int e = 0; e <= NE; ++e)
        e2d.I[e] = e * ne;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < NE; ++e)
    {
        const int ex = e % nx, ey = (e / nx) % ny, ez = e / (nx * ny);
        int *row = &e2d.J[(size_t)e * ne];
        int l = 0;
        for (int c = 0; c < (dim == 3 ? n1 : 1); ++c)
            for (int b = 0; b < n1; ++b)
                for (int a = 0; a < n1; ++a)
                    row[l++] = (p * ex + a) +
                               gx * ((p * ey + b) + gy * (dim == 3 ? p * ez + c : 0));
    }

    // elem_to_elem: face neighbours, ascending.
    Table &e2e = fp->elem_to_elem;
    e2e.nrows = NE;
    e2e.ncols = NE;
    e2e.I.assign((size_t)NE + 1, 0);
    for (int e = 0; e < NE; ++e)
    {
        const int ex = e % nx, ey = (e / nx) % ny, ez = e / (nx * ny);
        int c = 0;
        c += (ex > 0) + (ex < nx - 1) + (ey > 0) + (ey < ny - 1);
        if (dim == 3)
            c += (ez > 0) + (ez < nz - 1);
        e2e.I[e + 1] = e2e.I[e] + c;
    }
    e2e.J.resize(e2e.I[NE]);
    for (int e = 0; e < NE; ++e)
    {
        const int ex = e % nx, ey = (e / nx) % ny, ez = e / (nx * ny);
        int q = e2e.I[e];
        if (dim == 3 && ez > 0)
            e2e.J[q++] = e - nx * ny;
        if (ey > 0)
            e2e.J[q++] = e - nx;
        if (ex > 0)
            e2e.J[q++] = e - 1;
        if (ex < nx - 1)
            e2e.J[q++] = e + 1;
        if (ey < ny - 1)
            e2e.J[q++] = e + nx;
        if (dim == 3 && ez < nz - 1)
            e2e.J[q++] = e + nx * ny;
    }

    // coefficient
    fp->coef.assign(NE, 1.);
    if (coef_kind == FEM_COEF_LOGNORMAL)
    {
        std::mt19937_64 rng(seed);
        std::normal_distribution<double> nd(0., 1.);
        std::vector<double> g(NE);
        for (int e = 0; e < NE; ++e)
            g[e] = nd(rng);
        box_filter_axis(g, nx, ny, nz, 0, 2);
        box_filter_axis(g, nx, ny, nz, 1, 2);
        if (dim == 3)
            box_filter_axis(g, nx, ny, nz, 2, 2);
        double mn = g[0], mx = g[0];
        for (int e = 0; e < NE; ++e)
        {
            mn = std::min(mn, g[e]);
            mx = std::max(mx, g[e]);
        }
        const double half = 0.5 * std::log(contrast);
        for (int e = 0; e < NE; ++e)
        {
            const double t = (mx > mn) ? (2. * (g[e] - mn) / (mx - mn) - 1.) : 0.;
            fp->coef[e] = std::exp(half * t);
        }
    }
    else if (coef_kind == FEM_COEF_CHECKER)
    {
        for (int e = 0; e < NE; ++e)
        {
            const int ex = e % nx, ey = (e / nx) % ny, ez = e / (nx * ny);
            const int par = (ex / 4 + ey / 4 + ez / 4) & 1;
            fp->coef[e] = par ? contrast : 1.;
        }
    }

    else if (coef_kind == FEM_COEF_MLTEST)
    {
        const double d = 10.;
        for (int e = 0; e < NE; ++e)
        {
            const int ex = e % nx, ey = (e / nx) % ny, ez = e / (nx * ny);
            const double x0 = (ex + 0.5) / nx, x1 = (ey + 0.5) / ny, x2 = (ez + 0.5) / nz;
            const int c0 = (int)std::ceil(x0 * d) & 1, c1 = (int)std::ceil(x1 * d) & 1,
                      c2 = (int)std::ceil(x2 * d) & 1;
            bool hi;
            if (dim == 2)
                hi = (c0 == c1);
            else
                hi = (c2 && c0 == c1) || (!c2 && c0 != c1);
            fp->coef[e] = hi ? 1e6 : 1e0;
        }
    }

    // reference element matrix for a cell hx x hy x hz
    const double hx = 1. / nx, hy = 1. / ny, hz = (dim == 3) ? 1. / nz : 1.;
    double M1[9], S1[9], L1[3];
    matrices1d(p, M1, S1, L1);
    std::vector<double> Kref((size_t)ne * ne), Lref(ne);
    for (int r = 0; r < ne; ++r)
    {
        const int ra = r % n1, rb = (r / n1) % n1, rc = r / (n1 * n1);
        Lref[r] = L1[ra] * hx * L1[rb] * hy * (dim == 3 ? L1[rc] * hz : 1.);
        for (int c = 0; c < ne; ++c)
        {
            const int ca = c % n1, cb = (c / n1) % n1, cc = c / (n1 * n1);
            const double Mx = M1[ra * n1 + ca] * hx, Sx = S1[ra * n1 + ca] / hx;
            const double My = M1[rb * n1 + cb] * hy, Sy = S1[rb * n1 + cb] / hy;
            double v;
            if (dim == 3)
            {
                const double Mz = M1[rc * n1 + cc] * hz, Sz = S1[rc * n1 + cc] / hz;
                v = Sx * My * Mz + Mx * Sy * Mz + Mx * My * Sz;
            }
            else
                v = Sx * My + Mx * Sy;
            Kref[(size_t)c * ne + r] = v;
        }
    }
    // symmetrize exactly (roundoff in the quadrature sums)
    for (int r = 0; r < ne; ++r)
        for (int c = r + 1; c < ne; ++c)
        {
            const double v = 0.5 * (Kref[(size_t)c * ne + r] + Kref[(size_t)r * ne + c]);
            Kref[(size_t)c * ne + r] = Kref[(size_t)r * ne + c] = v;
        }
    fp->elmat.resize((size_t)NE * ne * ne);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < NE; ++e)
    {
        double *dst = &fp->elmat[(size_t)e * ne * ne];
        const double k = fp->coef[e];
        for (int i = 0; i < ne * ne; ++i)
            dst[i] = k * Kref[i];
    }

    // essential dofs: whole boundary
    fp->bdr_dofs.assign(ND, 0);
    for (int d = 0; d < ND; ++d)
    {
        const int ix = d % gx, iy = (d / gx) % gy, iz = d / (gx * gy);
        bool on = (ix == 0 && (ess_mask & 1)) || (ix == gx - 1 && (ess_mask & 2)) ||
                  (iy == 0 && (ess_mask & 4)) || (iy == gy - 1 && (ess_mask & 8));
        if (dim == 3)
            on = on || (iz == 0 && (ess_mask & 16)) || (iz == gz - 1 && (ess_mask & 32));
        if (on)
            fp->bdr_dofs[d] = 0x02; // AGG_ON_ESS_DOMAIN_BORDER_FLAG
    }

    // global assembly (rows independent): element contributions in ascending
    // element order, columns sorted ascending
    Table d2e;
    Transpose(e2d, d2e, ND);
    SparseMatrix &A = fp->A;
    A.h = A.w = ND;
    A.I.assign((size_t)ND + 1, 0);
    const int maxrow = (dim == 2) ? (2 * p + 1) * (2 * p + 1)
                                  : (2 * p + 1) * (2 * p + 1) * (2 * p + 1);
    std::vector<int> rowcnt(ND);
#pragma omp parallel
    {
        std::vector<int> cols;
        cols.reserve(maxrow);
#pragma omp for schedule(static)
        for (int d = 0; d < ND; ++d)
        {
            cols.clear();
            for (int q = d2e.I[d]; q < d2e.I[d + 1]; ++q)
            {
                const int *ed = e2d.GetRow(d2e.J[q]);
                for (int l = 0; l < ne; ++l)
                    cols.push_back(ed[l]);
            }
            std::sort(cols.begin(), cols.end());
            rowcnt[d] = (int)(std::unique(cols.begin(), cols.end()) - cols.begin());
        }
    }
    for (int d = 0; d < ND; ++d)
        A.I[d + 1] = A.I[d] + rowcnt[d];
    A.J.resize(A.I[ND]);
    A.A.assign(A.I[ND], 0.);
    fp->b.assign(ND, 0.);
#pragma omp parallel
    {
        std::vector<int> cols;
        cols.reserve(maxrow);
#pragma omp for schedule(static)
        for (int d = 0; d < ND; ++d)
        {
            cols.clear();
            for (int q = d2e.I[d]; q < d2e.I[d + 1]; ++q)
            {
                const int *ed = e2d.GetRow(d2e.J[q]);
                for (int l = 0; l < ne; ++l)
                    cols.push_back(ed[l]);
            }
            std::sort(cols.begin(), cols.end());
            const int cnt = (int)(std::unique(cols.begin(), cols.end()) - cols.begin());
            int *Jrow = &A.J[A.I[d]];
            double *Arow = &A.A[A.I[d]];
            for (int c = 0; c < cnt; ++c)
                Jrow[c] = cols[c];
            double load = 0.;
            for (int q = d2e.I[d]; q < d2e.I[d + 1]; ++q)
            {
                const int e = d2e.J[q];
                const int *ed = e2d.GetRow(e);
                const double *Ke = &fp->elmat[(size_t)e * ne * ne];
                int lr = -1;
                for (int l = 0; l < ne; ++l)
                    if (ed[l] == d)
                        lr = l;
                load += Lref[lr];
                for (int l = 0; l < ne; ++l)
                {
                    const int pos = (int)(std::lower_bound(Jrow, Jrow + cnt, ed[l]) - Jrow);
                    Arow[pos] += Ke[(size_t)l * ne + lr];
                }
            }
            fp->b[d] = load;
        }
    }
    // eliminate essential BC keeping the diagonal (x_ess = 0)
#pragma omp parallel for schedule(static)
    for (int d = 0; d < ND; ++d)
    {
        const bool ess_row = fp->bdr_dofs[d] & 0x02;
        for (int q = A.I[d]; q < A.I[d + 1]; ++q)
        {
            const int c = A.J[q];
            if (c == d)
                continue;
            if (ess_row || (fp->bdr_dofs[c] & 0x02))
                A.A[q] = 0.;
        }
        if (ess_row)
            fp->b[d] = 0.;
    }
    return fp;
}

} // namespace saamge
