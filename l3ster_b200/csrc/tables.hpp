// Host-side reference-element tables for the B200 path: Gauss–Lobatto nodes, Gauss–Legendre rules, 1-D Lagrange
// interpolation / derivative matrices, and dense tensor tables at arbitrary reference points.
//
// Replaces (for the device path) the reference's basisfun/, quad/ and math/ layers:
//   math/LobattoRuleAbsc.hpp:10-35, math/ComputeGaussRule.hpp:25-61, basisfun/ReferenceBasisFunction.hpp:28-153,
//   quad/GenerateQuadrature.hpp:11-77, algsys/SumFactorization.hpp:25-65,
//   basisfun/ReferenceElementBasisAtQuadrature.hpp:10-96, mapping/ReferenceBoundaryToSideMapping.hpp:14-48.
// The reference builds its Lagrange polynomials as monomial coefficients and evaluates them with Horner; here the nodes
// and weights are computed by Newton iteration in long double and the Lagrange polynomials are evaluated in product form
// in long double, then rounded — agreement with the reference tables is at the 1e-15 level (checked against the oracle,
// which keeps the reference's algorithm).
#ifndef L3B_TABLES_HPP
#define L3B_TABLES_HPP

#include <cmath>
#include <stdexcept>
#include <vector>

namespace l3b::tables
{
using ld = long double;

inline void legendre(int n, ld x, ld& p, ld& dp, ld& ddp)
{
    if (n == 0)
    {
        p = 1;
        dp = ddp = 0;
        return;
    }
    ld p0 = 1, p1 = x;
    for (int k = 2; k <= n; ++k)
    {
        const ld pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0          = p1;
        p1          = pk;
    }
    p   = p1;
    dp  = n * (p0 - x * p1) / (1 - x * x);
    ddp = (2 * x * dp - static_cast< ld >(n) * (n + 1) * p1) / (1 - x * x);
}

// Gauss–Lobatto–Legendre nodes: {-1, roots of P'_{n-1}, 1}
inline std::vector< ld > gllNodes(int n_points)
{
    if (n_points < 2)
        throw std::invalid_argument{"GLL rule needs >= 2 points"};
    std::vector< ld > x(n_points);
    x.front()   = -1;
    x.back()    = 1;
    const int m = n_points - 1;
    for (int i = 1; i < n_points - 1; ++i)
    {
        ld xi = -std::cos(3.14159265358979323846264338327950288L * i / m);
        for (int it = 0; it < 100; ++it)
        {
            ld p, dp, ddp;
            legendre(m, xi, p, dp, ddp);
            const ld dx = dp / ddp;
            xi -= dx;
            if (std::fabs(dx) < 1e-19L)
                break;
        }
        x[i] = xi;
    }
    if (n_points % 2)
        x[n_points / 2] = 0;
    // enforce exact antisymmetry
    for (int i = 0; i < n_points / 2; ++i)
        x[n_points - 1 - i] = -x[i];
    return x;
}

// Gauss–Legendre rule on [-1, 1], ascending points
inline void glRule(int n, std::vector< ld >& x, std::vector< ld >& w)
{
    x.assign(n, 0);
    w.assign(n, 0);
    for (int i = 0; i < n; ++i)
    {
        ld xi = -std::cos(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
        ld p = 0, dp = 1, ddp = 0;
        if (n == 1)
            xi = 0;
        else
            for (int it = 0; it < 100; ++it)
            {
                legendre(n, xi, p, dp, ddp);
                const ld dx = p / dp;
                xi -= dx;
                if (std::fabs(dx) < 1e-19L)
                    break;
            }
        if (n > 1)
            legendre(n, xi, p, dp, ddp);
        x[i] = xi;
        w[i] = 2 / ((1 - xi * xi) * dp * dp);
    }
    if (n % 2)
        x[n / 2] = 0;
    for (int i = 0; i < n / 2; ++i)
    {
        x[n - 1 - i] = -x[i];
        w[n - 1 - i] = w[i];
    }
}

// value and derivative of the b-th Lagrange polynomial through `nodes`, at x (product form)
inline void lagrange(const std::vector< ld >& nodes, int b, ld x, ld& val, ld& der)
{
    const int n = static_cast< int >(nodes.size());
    ld        denom = 1;
    for (int j = 0; j < n; ++j)
        if (j != b)
            denom *= nodes[b] - nodes[j];
    val = 1;
    for (int j = 0; j < n; ++j)
        if (j != b)
            val *= x - nodes[j];
    val /= denom;
    der = 0;
    for (int k = 0; k < n; ++k)
    {
        if (k == b)
            continue;
        ld prod = 1;
        for (int j = 0; j < n; ++j)
            if (j != b and j != k)
                prod *= x - nodes[j];
        der += prod;
    }
    der /= denom;
}

// 1-D tables used by the sum-factorised kernels
struct Tables1D
{
    int                   nb = 0, nq = 0;
    std::vector< double > interp; // [b][q] = l_b(xi_q)            (SumFactorization.hpp:25-36)
    std::vector< double > der;    // [b][q] = l_b'(xi_q)           (SumFactorization.hpp:38-49)
    std::vector< double > colloc; // [m][q] = lq_m'(xi_q), lq = Lagrange basis through the quadrature points
    std::vector< double > w;      // GL weights
    std::vector< double > pts;    // GL points
    std::vector< double > interp1, der1; // order-1 (geometry) basis at the quadrature points, [v][q] (:431-436)
};

inline Tables1D makeTables1D(int order, int nq)
{
    Tables1D t;
    t.nb = order + 1;
    t.nq = nq;
    const auto        nodes = gllNodes(order + 1);
    std::vector< ld > qx, qw;
    glRule(nq, qx, qw);
    t.interp.resize(t.nb * nq);
    t.der.resize(t.nb * nq);
    t.colloc.resize(nq * nq);
    t.interp1.resize(2 * nq);
    t.der1.resize(2 * nq);
    for (int q = 0; q < nq; ++q)
    {
        t.w.push_back(static_cast< double >(qw[q]));
        t.pts.push_back(static_cast< double >(qx[q]));
        for (int b = 0; b < t.nb; ++b)
        {
            ld v, d;
            lagrange(nodes, b, qx[q], v, d);
            t.interp[b * nq + q] = static_cast< double >(v);
            t.der[b * nq + q]    = static_cast< double >(d);
        }
        for (int m = 0; m < nq; ++m)
        {
            ld v, d;
            lagrange(qx, m, qx[q], v, d);
            t.colloc[m * nq + q] = static_cast< double >(d);
        }
        t.interp1[0 * nq + q] = static_cast< double >((1 - qx[q]) / 2);
        t.interp1[1 * nq + q] = static_cast< double >((1 + qx[q]) / 2);
        t.der1[0 * nq + q]    = -0.5;
        t.der1[1 * nq + q]    = 0.5;
    }
    return t;
}

// Dense tables at a set of reference points: values [q][a], derivatives [q][d][a] (a = ix + nb*iy + nb^2*iz), the
// device-side layout of basis::ReferenceBasisAtQuadrature (basisfun/ReferenceBasisAtPoints.hpp:8-21)
struct DenseTables
{
    int                   dim = 0, n_bases = 0, n_qp = 0;
    std::vector< double > points;  // [q][dim]
    std::vector< double > weights; // [q]
    std::vector< double > values, derivatives;
};

inline DenseTables makeDenseTables(int dim, int order, const std::vector< ld >& pts, const std::vector< ld >& wts)
{
    DenseTables t;
    t.dim     = dim;
    const int nb = order + 1;
    t.n_bases = 1;
    for (int d = 0; d < dim; ++d)
        t.n_bases *= nb;
    t.n_qp = static_cast< int >(wts.size());
    t.points.resize(static_cast< std::size_t >(t.n_qp) * dim);
    t.weights.resize(t.n_qp);
    t.values.resize(static_cast< std::size_t >(t.n_qp) * t.n_bases);
    t.derivatives.resize(static_cast< std::size_t >(t.n_qp) * dim * t.n_bases);
    const auto        nodes = gllNodes(nb);
    std::vector< ld > v1(static_cast< std::size_t >(dim) * nb), d1(v1.size());
    for (int q = 0; q < t.n_qp; ++q)
    {
        t.weights[q] = static_cast< double >(wts[q]);
        for (int d = 0; d < dim; ++d)
        {
            t.points[static_cast< std::size_t >(q) * dim + d] = static_cast< double >(pts[static_cast< std::size_t >(q) * dim + d]);
            for (int b = 0; b < nb; ++b)
                lagrange(nodes, b, pts[static_cast< std::size_t >(q) * dim + d], v1[d * nb + b], d1[d * nb + b]);
        }
        for (int a = 0; a < t.n_bases; ++a)
        {
            int idx[3] = {a % nb, (a / nb) % nb, a / (nb * nb)};
            ld  val    = 1;
            for (int d = 0; d < dim; ++d)
                val *= v1[d * nb + idx[d]];
            t.values[static_cast< std::size_t >(q) * t.n_bases + a] = static_cast< double >(val);
            for (int dd = 0; dd < dim; ++dd)
            {
                ld der = 1;
                for (int d = 0; d < dim; ++d)
                    der *= d == dd ? d1[d * nb + idx[d]] : v1[d * nb + idx[d]];
                t.derivatives[(static_cast< std::size_t >(q) * dim + dd) * t.n_bases + a] = static_cast< double >(der);
            }
        }
    }
    return t;
}

// tensor Gauss–Legendre rule on the reference element
inline void domainQuadrature(int dim, int nq, std::vector< ld >& pts, std::vector< ld >& wts)
{
    std::vector< ld > x, w;
    glRule(nq, x, w);
    int Q = 1;
    for (int d = 0; d < dim; ++d)
        Q *= nq;
    pts.assign(static_cast< std::size_t >(Q) * dim, 0);
    wts.assign(Q, 1);
    for (int q = 0; q < Q; ++q)
    {
        int rem = q;
        for (int d = 0; d < dim; ++d) // x fastest — the order is immaterial to any result (sums over q)
        {
            const int i                              = rem % nq;
            rem /= nq;
            pts[static_cast< std::size_t >(q) * dim + d] = x[i];
            wts[q] *= w[i];
        }
    }
}

// (dim-1)-dimensional tensor rule mapped onto side `side` of the reference quad/hex: the exact images of
// mapping/ReferenceBoundaryToSideMapping.hpp:14-48 (rotations by multiples of pi/2 written out exactly)
inline void sideQuadrature(int dim, int nq, int side, std::vector< ld >& pts, std::vector< ld >& wts)
{
    std::vector< ld > x, w;
    glRule(nq, x, w);
    if (dim == 2)
    {
        pts.assign(static_cast< std::size_t >(nq) * 2, 0);
        wts.assign(nq, 0);
        for (int i = 0; i < nq; ++i)
        {
            const ld s = x[i];
            ld       px = 0, py = 0;
            switch (side)
            {
            case 0: // R(pi) (s, 0) + (0, -1)
                px = -s;
                py = -1;
                break;
            case 1:
                px = s;
                py = 1;
                break;
            case 2: // rot2D(pi/2) = [[0, 1], [-1, 0]] → (0, -s) + (-1, 0)
                px = -1;
                py = -s;
                break;
            case 3: // rot2D(-pi/2) = [[0, -1], [1, 0]] → (0, s) + (1, 0)
                px = 1;
                py = s;
                break;
            default:
                throw std::out_of_range{"quad side"};
            }
            pts[i * 2]     = px;
            pts[i * 2 + 1] = py;
            wts[i]         = w[i];
        }
        return;
    }
    pts.assign(static_cast< std::size_t >(nq) * nq * 3, 0);
    wts.assign(static_cast< std::size_t >(nq) * nq, 0);
    int idx = 0;
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < nq; ++j, ++idx)
        {
            const ld s = x[i], t = x[j]; // reference quad point (s, t, 0)
            ld       p[3] = {0, 0, 0};
            switch (side)
            {
            case 0: // RotX(pi): (s, -t, 0) + (0, 0, -1)
                p[0] = s;
                p[1] = -t;
                p[2] = -1;
                break;
            case 1:
                p[0] = s;
                p[1] = t;
                p[2] = 1;
                break;
            case 2: // RotX(-pi/2) = [[1,0,0],[0,0,1],[0,-1,0]]: (s, 0, -t) + (0, -1, 0)
                p[0] = s;
                p[1] = -1;
                p[2] = -t;
                break;
            case 3: // RotX(pi/2) = [[1,0,0],[0,0,-1],[0,1,0]]: (s, 0, t) + (0, 1, 0)
                p[0] = s;
                p[1] = 1;
                p[2] = t;
                break;
            case 4: // RotY(pi/2) = [[0,0,1],[0,1,0],[-1,0,0]]: (0, t, -s) + (-1, 0, 0)
                p[0] = -1;
                p[1] = t;
                p[2] = -s;
                break;
            case 5: // RotY(-pi/2) = [[0,0,-1],[0,1,0],[1,0,0]]: (0, t, s) + (1, 0, 0)
                p[0] = 1;
                p[1] = t;
                p[2] = s;
                break;
            default:
                throw std::out_of_range{"hex side"};
            }
            for (int d = 0; d < 3; ++d)
                pts[static_cast< std::size_t >(idx) * 3 + d] = p[d];
            wts[idx] = w[i] * w[j];
        }
}
} // namespace l3b::tables
#endif
