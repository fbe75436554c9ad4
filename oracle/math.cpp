// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp). math/, quad/, basisfun/ restatement.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>

namespace orc
{
// math/Legendre.hpp:8-49 — same three-term recurrence, in double, coefficients highest power first
std::vector< val_t > legendreCoefs(int N)
{
    std::vector< val_t > coefs(N + 1, 0.);
    if (N == 0)
        coefs[0] = 1.;
    else if (N == 1)
    {
        coefs[0] = 1.;
        coefs[1] = 0.;
    }
    else
    {
        std::vector< val_t > P_n1(N + 1, 0.), P_n2(N + 1, 0.);
        P_n1[N - 1]  = 1.;   // *(P_n1.rbegin() + 1) = 1   → P_1 = x
        coefs[N]     = -.5;  // P_2 = 1.5 x^2 - 0.5
        coefs[N - 2] = 1.5;
        const auto a = [](int x) { return static_cast< val_t >(2 * x - 1) / static_cast< val_t >(x); };
        const auto c = [](int x) { return static_cast< val_t >(x - 1) / static_cast< val_t >(x); };
        for (int i = 3; i <= N; ++i)
        {
            const int index = N - i;
            std::copy(P_n1.begin() + index, P_n1.end(), P_n2.begin() + index);
            std::copy(coefs.begin() + index, coefs.end(), P_n1.begin() + index);
            // coefs[index + k] = a(i) * P_n1[index + 1 + k] - c(i) * P_n2[index + k]
            for (int k = 0; index + 1 + k <= N; ++k)
                coefs[index + k] = a(i) * P_n1[index + 1 + k] - c(i) * P_n2[index + k];
            coefs[N] = -c(i) * P_n2[N];
        }
    }
    return coefs;
}

// math/Polynomial.hpp:80-95
std::vector< val_t > polyDerivative(const std::vector< val_t >& coefs)
{
    const int order = static_cast< int >(coefs.size()) - 1;
    if (order == 0)
        return {0.};
    std::vector< val_t > ret(order);
    for (int i = 0; i < order; ++i)
        ret[i] = coefs[i] * static_cast< val_t >(order - i);
    return ret;
}

// math/Polynomial.hpp:57-66 (Horner)
val_t polyEval(const std::vector< val_t >& coefs, val_t x)
{
    val_t ret = 0;
    for (val_t c : coefs)
    {
        ret *= x;
        ret += c;
    }
    return ret;
}

namespace
{
// Legendre P_n and P'_n, P''_n at x in extended precision via the Bonnet recurrence
struct LegVals
{
    long double p, dp, ddp;
};
LegVals legendreEval(int n, long double x)
{
    long double p0 = 1.L, p1 = x;
    if (n == 0)
        return {1.L, 0.L, 0.L};
    for (int k = 2; k <= n; ++k)
    {
        const long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0                   = p1;
        p1                   = pk;
    }
    // (1-x^2) P'_n = n (P_{n-1} - x P_n) ;  (1-x^2) P''_n = 2x P'_n - n(n+1) P_n
    const long double dp  = n * (p0 - x * p1) / (1.L - x * x);
    const long double ddp = (2.L * x * dp - static_cast< long double >(n) * (n + 1) * p1) / (1.L - x * x);
    return {p1, dp, ddp};
}
} // namespace

// math/LobattoRuleAbsc.hpp:10-35: {-1, roots of P'_{n-1}, 1}. The reference finds the roots as eigenvalues of the
// companion matrix of the monomial-form polynomial (math/Polynomial.hpp:98-122, Eigen::eigenvalues); Eigen is absent
// here, so the roots are found by Newton iteration on P'_{n-1} in long double and rounded to double. Both agree with the
// closed forms of tests/MathTests.cpp:165-214 to 1e-14 (the reference's own tolerance).
const std::vector< val_t >& lobattoAbsc(int n_points)
{
    static std::map< int, std::vector< val_t > > cache;
    static std::mutex                            mtx;
    std::lock_guard                              lock{mtx};
    if (auto it = cache.find(n_points); it != cache.end())
        return it->second;
    if (n_points < 2)
        throw std::logic_error{"Lobatto rule needs >= 2 points"};
    std::vector< val_t > r(n_points);
    r.front() = -1.;
    r.back()  = 1.;
    const int m = n_points - 1;
    for (int i = 1; i < n_points - 1; ++i)
    {
        long double x = -std::cos(M_PIl * i / m); // Chebyshev–Gauss–Lobatto initial guess
        for (int it = 0; it < 100; ++it)
        {
            const auto        v  = legendreEval(m, x);
            const long double dx = v.dp / v.ddp;
            x -= dx;
            if (std::fabs(dx) < 1e-19L)
                break;
        }
        r[i] = static_cast< val_t >(x);
    }
    if (n_points % 2 == 1)
        r[n_points / 2] = 0.; // the middle root is exactly 0 (math/LobattoRuleAbsc.hpp:20 special-cases n=3 the same way)
    return cache.emplace(n_points, std::move(r)).first->second;
}

// math/LagrangeInterpolation.hpp:12-40 — verbatim algorithm (monomial coefficients, "accurate until N ≈ 16")
std::vector< val_t > lagrangeInterp(const std::vector< val_t >& x, const std::vector< val_t >& y)
{
    const std::size_t    N = x.size();
    std::vector< val_t > lag_coefs(N, 0.);
    for (std::size_t i = 0; i != N; ++i)
    {
        std::vector< val_t > roots;
        roots.reserve(N - 1);
        for (std::size_t j = 0; j != N; ++j)
            if (j != i)
                roots.push_back(x[j]);
        std::vector< val_t > l_i(N, 0.);
        l_i.front() = 1.;
        for (std::size_t j = 0; j != N - 1; ++j)
            for (std::size_t k = j + 1; k > 0; --k)
                l_i[k] -= l_i[k - 1] * roots[j];
        const val_t s = y[i] / polyEval(l_i, x[i]);
        for (std::size_t j = 0; j != N; ++j)
            lag_coefs[j] += s * l_i[j];
    }
    return lag_coefs;
}

// math/ComputeGaussRule.hpp:25-61 with the Legendre recurrence of quad/ReferenceQuadrature.hpp:32-41. The reference
// diagonalises the Golub–Welsch Jacobi matrix in long double (Eigen::SelfAdjointEigenSolver) and rounds to double; here
// the same nodes are obtained by Newton iteration on P_n in long double, weights 2/((1-x^2) P'_n(x)^2), then rounded.
// Eigenvalues come out ascending, as here. Pinned by tests/QuadratureTests.cpp:10-61 (1e-10).
const GaussRule& gaussLegendre(int n)
{
    static std::map< int, GaussRule > cache;
    static std::mutex                 mtx;
    std::lock_guard                   lock{mtx};
    if (auto it = cache.find(n); it != cache.end())
        return it->second;
    GaussRule rule;
    rule.points.resize(n);
    rule.weights.resize(n);
    for (int i = 0; i < n; ++i)
    {
        long double x = -std::cos(M_PIl * (i + 0.75L) / (n + 0.5L));
        for (int it = 0; it < 100; ++it)
        {
            if (n == 1)
            {
                x = 0.L;
                break;
            }
            const auto        v  = legendreEval(n, x);
            const long double dx = v.p / v.dp;
            x -= dx;
            if (std::fabs(dx) < 1e-19L)
                break;
        }
        long double dp;
        if (n == 1)
            dp = 1.L;
        else
            dp = legendreEval(n, x).dp;
        rule.points[i]  = static_cast< val_t >(x);
        rule.weights[i] = static_cast< val_t >(2.L / ((1.L - x * x) * dp * dp));
    }
    if (n % 2 == 1)
        rule.points[n / 2] = 0.;
    return cache.emplace(n, std::move(rule)).first->second;
}

// quad/GenerateQuadrature.hpp:11-77
Quadrature makeQuadrature(ElementType et, int quad_order)
{
    const auto& ref = gaussLegendre(refQuadSize(quad_order));
    const int   n   = static_cast< int >(ref.points.size());
    Quadrature  q;
    q.dim  = nativeDim(et);
    q.size = ipow(n, q.dim);
    q.points.resize(static_cast< std::size_t >(q.size) * q.dim);
    q.weights.resize(q.size);
    if (et == Line)
    {
        q.points  = ref.points;
        q.weights = ref.weights;
    }
    else if (et == Quad)
    {
        int index = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
            {
                q.points[index * 2 + 0] = ref.points[i];
                q.points[index * 2 + 1] = ref.points[j];
                q.weights[index]        = ref.weights[i] * ref.weights[j];
                ++index;
            }
    }
    else
    {
        int index = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k)
                {
                    q.points[index * 3 + 0] = ref.points[i];
                    q.points[index * 3 + 1] = ref.points[j];
                    q.points[index * 3 + 2] = ref.points[k];
                    q.weights[index]        = ref.weights[i] * ref.weights[j] * ref.weights[k];
                    ++index;
                }
    }
    return q;
}

// basisfun/ReferenceBasisFunction.hpp:28-57
const LineBasis& lineBasis(int order)
{
    static std::map< int, LineBasis > cache;
    static std::mutex                 mtx;
    std::lock_guard                   lock{mtx};
    if (auto it = cache.find(order); it != cache.end())
        return it->second;
    LineBasis   b;
    b.order     = order;
    const auto& x = lobattoAbsc(order + 1);
    for (int I = 0; I <= order; ++I)
    {
        std::vector< val_t > vals(order + 1, 0.);
        vals[I] = 1.;
        b.polys.push_back(lagrangeInterp(x, vals));
        b.ders.push_back(polyDerivative(b.polys.back()));
    }
    return cache.emplace(order, std::move(b)).first->second;
}

namespace
{
val_t lineVal(const LineBasis& lb, int I, val_t x)
{
    return polyEval(lb.polys[I], x);
}
val_t lineDer(const LineBasis& lb, int I, val_t x)
{
    return polyEval(lb.ders[I], x);
}
} // namespace

// basisfun/ReferenceBasisFunction.hpp:74-153 — tensor-product expansion, I = ix + n*iy + n^2*iz, product order kept
val_t refBasisValue(ElementType et, int order, int I, const val_t* pt)
{
    const auto& lb = lineBasis(order);
    const int   n  = order + 1;
    switch (et)
    {
    case Line:
        return lineVal(lb, I, pt[0]);
    case Quad:
        return lineVal(lb, I % n, pt[0]) * lineVal(lb, I / n, pt[1]);
    case Hex: {
        const int xe = I % (n * n), ze = I / (n * n);
        return (lineVal(lb, xe % n, pt[0]) * lineVal(lb, xe / n, pt[1])) * lineVal(lb, ze, pt[2]);
    }
    }
    return 0.;
}

val_t refBasisDer(ElementType et, int order, int I, int d, const val_t* pt)
{
    const auto& lb = lineBasis(order);
    const int   n  = order + 1;
    const auto  f  = [&](int dir, int ind, val_t x) { return dir == d ? lineDer(lb, ind, x) : lineVal(lb, ind, x); };
    switch (et)
    {
    case Line:
        return f(0, I, pt[0]);
    case Quad:
        return f(0, I % n, pt[0]) * f(1, I / n, pt[1]);
    case Hex: {
        const int xe = I % (n * n), ze = I / (n * n);
        return (f(0, xe % n, pt[0]) * f(1, xe / n, pt[1])) * f(2, ze, pt[2]);
    }
    }
    return 0.;
}

namespace
{
RefBasisAtQuad evalAtQuadrature(ElementType et, int order, Quadrature quad)
{
    RefBasisAtQuad r;
    r.et      = et;
    r.order   = order;
    r.dim     = nativeDim(et);
    r.n_bases = numNodes(et, order);
    r.quad    = std::move(quad);
    const int Q = r.quad.size;
    r.values.resize(static_cast< std::size_t >(Q) * r.n_bases);
    r.derivatives.resize(static_cast< std::size_t >(Q) * r.dim * r.n_bases);
    for (int q = 0; q < Q; ++q)
    {
        const val_t* pt = &r.quad.points[static_cast< std::size_t >(q) * r.dim];
        for (int a = 0; a < r.n_bases; ++a)
        {
            r.values[static_cast< std::size_t >(q) * r.n_bases + a] = refBasisValue(et, order, a, pt);
            for (int d = 0; d < r.dim; ++d)
                r.derivatives[(static_cast< std::size_t >(q) * r.dim + d) * r.n_bases + a] = refBasisDer(et, order, a, d, pt);
        }
    }
    return r;
}
} // namespace

// basis::getBasisAtNodes + mesh::getNodeLocations (basisfun/ReferenceBasisAtNodes.hpp, mesh/NodeReferenceLocation.hpp:31-44): the basis
// and its reference derivatives at the element's own node locations; "point" a is local node a, weights are not used
RefBasisAtQuad makeRefBasisAtNodes(ElementType et, int order)
{
    const int   dim = nativeDim(et), nb = order + 1, nn = numNodes(et, order);
    const auto& gll = lobattoAbsc(nb);
    Quadrature  pts;
    pts.dim  = dim;
    pts.size = nn;
    pts.points.assign(static_cast< std::size_t >(nn) * dim, 0.);
    pts.weights.assign(nn, 1.);
    for (int a = 0; a < nn; ++a)
    {
        int rem = a;
        for (int d = 0; d < dim; ++d, rem /= nb)
            pts.points[static_cast< std::size_t >(a) * dim + d] = gll[rem % nb];
    }
    return evalAtQuadrature(et, order, std::move(pts));
}

// basisfun/ReferenceElementBasisAtQuadrature.hpp:10-19
RefBasisAtQuad makeRefBasisAtDomainQuad(ElementType et, int order, int quad_order)
{
    return evalAtQuadrature(et, order, makeQuadrature(et, quad_order));
}

// basisfun/ReferenceElementBasisAtQuadrature.hpp:21-96: (D-1)-dimensional rule, zero-padded, rotated and translated
// onto the side (mapping/ReferenceBoundaryToSideMapping.hpp)
RefBasisAtQuad makeRefBasisAtBoundaryQuad(ElementType et, int order, int quad_order, int side)
{
    const int  dim = nativeDim(et);
    Quadrature sideq;
    sideq.dim = dim;
    if (dim > 1)
    {
        const auto ref = makeQuadrature(et == Hex ? Quad : Line, quad_order);
        sideq.size     = ref.size;
        sideq.weights  = ref.weights;
        sideq.points.assign(static_cast< std::size_t >(ref.size) * dim, 0.);
        std::vector< val_t > rot(dim * dim), trans(dim);
        refBoundaryToSide(et, side, rot.data(), trans.data());
        for (int c = 0; c < ref.size; ++c)
        {
            std::array< val_t, 3 > in{};
            for (int r = 0; r < dim - 1; ++r)
                in[r] = ref.points[static_cast< std::size_t >(c) * (dim - 1) + r];
            for (int r = 0; r < dim; ++r)
            {
                val_t acc = 0.;
                for (int k = 0; k < dim; ++k)
                    acc += rot[r * dim + k] * in[k];
                sideq.points[static_cast< std::size_t >(c) * dim + r] = acc + trans[r];
            }
        }
    }
    else
    {
        std::vector< val_t > rot(1), trans(1);
        refBoundaryToSide(et, side, rot.data(), trans.data());
        sideq.size    = 1;
        sideq.weights = {1.};
        sideq.points  = {rot[0] * 0. + trans[0]};
    }
    return evalAtQuadrature(et, order, std::move(sideq));
}
} // namespace orc
