// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp). algsys/SumFactorization.hpp restatement.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <array>
#include <map>
#include <tuple>
#include <utility>

namespace orc
{
// algsys/SumFactorization.hpp:25-65
SumFactTables makeSumFactTables(int basis_order, int quad_order, bool odd_even)
{
    SumFactTables t;
    t.nb              = basis_order + 1;
    t.nq              = refQuadSize(quad_order);
    t.odd_even        = odd_even;
    const auto& rule  = gaussLegendre(t.nq);
    t.weights         = rule.weights;
    t.interp.resize(static_cast< std::size_t >(t.nb) * t.nq);
    t.der.resize(t.interp.size());
    t.interp_t.resize(t.interp.size());
    t.der_t.resize(t.interp.size());
    for (int q = 0; q < t.nq; ++q)
        for (int b = 0; b < t.nb; ++b)
        {
            const val_t pt             = rule.points[q];
            const val_t v              = refBasisValue(Line, basis_order, b, &pt);
            const val_t d              = refBasisDer(Line, basis_order, b, 0, &pt);
            t.interp[b * t.nq + q]     = v;
            t.der[b * t.nq + q]        = d;
            t.interp_t[q * t.nb + b]   = v;
            t.der_t[q * t.nb + b]      = d;
        }
    return t;
}

namespace
{
// out = in^T * M  (:67-86).  in col-major n_in x cols, out col-major cols x n_out, M row-major n_in x n_out
void sweepStandard(const val_t* in, val_t* out, int n_in, int n_out, int cols, const val_t* M, bool accumulate)
{
    for (int c = 0; c < cols; ++c)
        for (int o = 0; o < n_out; ++o)
        {
            val_t acc = 0.;
            for (int i = 0; i < n_in; ++i)
                acc += in[i + n_in * c] * M[i * n_out + o];
            val_t& dst = out[c + static_cast< std::size_t >(cols) * o];
            dst        = accumulate ? dst + acc : acc;
        }
}

// The same sweep with compile-time extents, as the reference instantiates it (Eigen fixed-size maps, :67-86): identical summation order,
// so the results do not change; it exists so that the CPU baseline timed beside the GPU numbers is not slowed down by run-time loop bounds.
template < int NI, int NO >
void sweepStandardFixed(const val_t* __restrict__ in, val_t* __restrict__ out, int cols, const val_t* __restrict__ M, bool accumulate)
{
    val_t m[NI * NO];
    for (int k = 0; k < NI * NO; ++k)
        m[k] = M[k];
    for (int c = 0; c < cols; ++c)
    {
        val_t v[NI];
        for (int i = 0; i < NI; ++i)
            v[i] = in[i + NI * c];
        for (int o = 0; o < NO; ++o)
        {
            val_t acc = 0.;
            for (int i = 0; i < NI; ++i)
                acc += v[i] * m[i * NO + o];
            val_t& dst = out[c + static_cast< std::size_t >(cols) * o];
            dst        = accumulate ? dst + acc : acc;
        }
    }
}
using sweep_fn_t = void (*)(const val_t*, val_t*, int, const val_t*, bool);
constexpr int max_fixed_extent = 10;
template < int NI, int... NOs >
constexpr std::array< sweep_fn_t, sizeof...(NOs) > sweepRow(std::integer_sequence< int, NOs... >)
{
    return {&sweepStandardFixed< NI, NOs + 1 >...};
}
template < int... NIs >
constexpr std::array< std::array< sweep_fn_t, max_fixed_extent >, sizeof...(NIs) > sweepTable(std::integer_sequence< int, NIs... >)
{
    return {sweepRow< NIs + 1 >(std::make_integer_sequence< int, max_fixed_extent >{})...};
}
constexpr auto fixed_sweeps = sweepTable(std::make_integer_sequence< int, max_fixed_extent >{});

// odd-even decomposition (:88-258). psi tables are rebuilt per call — this path exists for parity, not for speed.
void sweepOddEven(const val_t* in, val_t* out, int rows, int colsM, int cols, const val_t* M, bool is_der, bool accumulate)
{
    const int psi_p_rows = (rows + 1) / 2, psi_m_rows = rows / 2;
    const int psi_p_cols = is_der ? colsM / 2 : (colsM + 1) / 2;
    const int psi_m_cols = is_der ? (colsM + 1) / 2 : colsM / 2;
    std::vector< val_t > psi_p(static_cast< std::size_t >(psi_p_rows) * psi_p_cols), psi_m(static_cast< std::size_t >(psi_m_rows) * psi_m_cols);
    // makePsiPlusImpl (:88-98)
    for (int r = 0; r < rows / 2; ++r)
        for (int c = 0; c < psi_p_cols; ++c)
            psi_p[r * psi_p_cols + c] = M[r * colsM + c] + M[(rows - r - 1) * colsM + c];
    if (rows % 2)
        for (int c = 0; c < psi_p_cols; ++c)
            psi_p[(psi_p_rows - 1) * psi_p_cols + c] = M[(rows / 2) * colsM + c];
    // makePsiMinusImpl (:100-108)
    for (int r = 0; r < psi_m_rows; ++r)
        for (int c = 0; c < psi_m_cols; ++c)
            psi_m[r * psi_m_cols + c] = M[r * colsM + c] - M[(rows - r - 1) * colsM + c];

    std::vector< val_t > e(psi_p_rows), o(std::max(psi_m_rows, 1)), ep(std::max(psi_p_cols, 1)), op(std::max(psi_m_cols, 1));
    for (int c = 0; c < cols; ++c)
    {
        const val_t* v = in + static_cast< std::size_t >(rows) * c;
        // makeEO (:159-181)
        for (int r = 0; r < psi_m_rows; ++r)
        {
            const val_t v1 = v[r], v2 = v[rows - r - 1];
            e[r] = .5 * (v1 + v2);
            o[r] = .5 * (v1 - v2);
        }
        if (psi_m_rows < psi_p_rows)
            e[psi_m_rows] = v[psi_m_rows];
        for (int j = 0; j < psi_p_cols; ++j)
        {
            val_t acc = 0.;
            for (int r = 0; r < psi_p_rows; ++r)
                acc += e[r] * psi_p[r * psi_p_cols + j];
            ep[j] = acc;
        }
        for (int j = 0; j < psi_m_cols; ++j)
        {
            val_t acc = 0.;
            for (int r = 0; r < psi_m_rows; ++r)
                acc += o[r] * psi_m[r * psi_m_cols + j];
            op[j] = acc;
        }
        // reconstructOddEven (:183-203); derivative sweeps pass (o', e') (:233)
        const val_t* first    = is_der ? op.data() : ep.data();
        const val_t* second   = is_der ? ep.data() : op.data();
        const int    n_first  = is_der ? psi_m_cols : psi_p_cols;
        const int    n_second = is_der ? psi_p_cols : psi_m_cols;
        const int    full     = n_first + n_second;
        const auto   put      = [&](int col, val_t val) {
            val_t& dst = out[c + static_cast< std::size_t >(cols) * col];
            dst        = accumulate ? dst + val : val;
        };
        for (int j = 0; j < n_second; ++j)
        {
            put(j, first[j] + second[j]);
            put(full - j - 1, first[j] - second[j]);
        }
        if (n_first > n_second)
            put(n_second, first[n_first - 1]);
    }
}
// sweepOddEven with compile-time extents: the same operations in the same order (the results are bit-identical), tables and temporaries
// on the stack. This is the sweep the reference picks for 2 <= p <= 6 (AssembleLocalSystem.hpp:43-48), i.e. the one the CPU baseline runs
// at p = 4: with run-time extents and four heap allocations per call it cost more than the standard sweep it is meant to beat.
template < int ROWS, int COLSM, bool DER >
void sweepOddEvenFixed(const val_t* __restrict__ in, val_t* __restrict__ out, int cols, const val_t* __restrict__ M, bool accumulate)
{
    constexpr int PPR = (ROWS + 1) / 2, PMR = ROWS / 2;
    constexpr int PPC = DER ? COLSM / 2 : (COLSM + 1) / 2, PMC = DER ? (COLSM + 1) / 2 : COLSM / 2;
    constexpr int N_FIRST = DER ? PMC : PPC, N_SECOND = DER ? PPC : PMC, FULL = N_FIRST + N_SECOND;
    val_t psi_p[(PPR * PPC > 0 ? PPR * PPC : 1)], psi_m[(PMR * PMC > 0 ? PMR * PMC : 1)];
    for (int r = 0; r < ROWS / 2; ++r)
        for (int c = 0; c < PPC; ++c)
            psi_p[r * PPC + c] = M[r * COLSM + c] + M[(ROWS - r - 1) * COLSM + c];
    if constexpr (ROWS % 2 != 0)
        for (int c = 0; c < PPC; ++c)
            psi_p[(PPR - 1) * PPC + c] = M[(ROWS / 2) * COLSM + c];
    for (int r = 0; r < PMR; ++r)
        for (int c = 0; c < PMC; ++c)
            psi_m[r * PMC + c] = M[r * COLSM + c] - M[(ROWS - r - 1) * COLSM + c];
    for (int c = 0; c < cols; ++c)
    {
        const val_t* v = in + static_cast< std::size_t >(ROWS) * c;
        val_t        e[PPR > 0 ? PPR : 1], o[PMR > 0 ? PMR : 1], ep[PPC > 0 ? PPC : 1], op[PMC > 0 ? PMC : 1];
        for (int r = 0; r < PMR; ++r)
        {
            const val_t v1 = v[r], v2 = v[ROWS - r - 1];
            e[r] = .5 * (v1 + v2);
            o[r] = .5 * (v1 - v2);
        }
        if constexpr (PMR < PPR)
            e[PMR] = v[PMR];
        for (int j = 0; j < PPC; ++j)
        {
            val_t acc = 0.;
            for (int r = 0; r < PPR; ++r)
                acc += e[r] * psi_p[r * PPC + j];
            ep[j] = acc;
        }
        for (int j = 0; j < PMC; ++j)
        {
            val_t acc = 0.;
            for (int r = 0; r < PMR; ++r)
                acc += o[r] * psi_m[r * PMC + j];
            op[j] = acc;
        }
        const val_t* first  = DER ? op : ep;
        const val_t* second = DER ? ep : op;
        const auto   put    = [&](int col, val_t val) {
            val_t& dst = out[c + static_cast< std::size_t >(cols) * col];
            dst        = accumulate ? dst + val : val;
        };
        for (int j = 0; j < N_SECOND; ++j)
        {
            put(j, first[j] + second[j]);
            put(FULL - j - 1, first[j] - second[j]);
        }
        if constexpr (N_FIRST > N_SECOND)
            put(N_SECOND, first[N_FIRST - 1]);
    }
}
template < bool DER, int NI, int... NOs >
constexpr std::array< sweep_fn_t, sizeof...(NOs) > oddEvenRow(std::integer_sequence< int, NOs... >)
{
    return {&sweepOddEvenFixed< NI, NOs + 1, DER >...};
}
template < bool DER, int... NIs >
constexpr std::array< std::array< sweep_fn_t, max_fixed_extent >, sizeof...(NIs) > oddEvenTable(std::integer_sequence< int, NIs... >)
{
    return {oddEvenRow< DER, NIs + 1 >(std::make_integer_sequence< int, max_fixed_extent >{})...};
}
constexpr auto fixed_odd_even_interp = oddEvenTable< false >(std::make_integer_sequence< int, max_fixed_extent >{});
constexpr auto fixed_odd_even_der    = oddEvenTable< true >(std::make_integer_sequence< int, max_fixed_extent >{});
} // namespace

void sumFactSweep(const val_t* in, val_t* out, int n_in, int n_out, int cols, const val_t* M, bool is_der, bool accumulate, bool odd_even)
{
    const bool fixed = n_in >= 1 and n_in <= max_fixed_extent and n_out >= 1 and n_out <= max_fixed_extent;
    if (odd_even and fixed)
        (is_der ? fixed_odd_even_der : fixed_odd_even_interp)[n_in - 1][n_out - 1](in, out, cols, M, accumulate);
    else if (odd_even)
        sweepOddEven(in, out, n_in, n_out, cols, M, is_der, accumulate);
    else if (fixed)
        fixed_sweeps[n_in - 1][n_out - 1](in, out, cols, M, accumulate);
    else
        sweepStandard(in, out, n_in, n_out, cols, M, accumulate);
}

namespace
{
using buf_t = std::vector< val_t >;

// sumFactBackQuad (:438-467) / sumFactBackHex (:469-504). `fill` goes to r[dim]; returns {vals, d/dxi, d/deta[, d/dzeta]},
// each row-major Q x F with qi = (qz*nq + qy)*nq + qx.
std::array< buf_t, 4 > sumFactBack(const SumFactTables& t, int dim, int F, const val_t* fill, std::size_t fill_size)
{
    const int             nb = t.nb, nq = t.nq;
    const std::size_t     bufsz = static_cast< std::size_t >(std::max(ipow(nb, dim), ipow(nq, dim))) * F;
    std::array< buf_t, 4 > r;
    for (auto& b : r)
        b.assign(bufsz, 0.);
    buf_t      temp(bufsz, 0.);
    const bool oe = t.odd_even;
    const auto I  = [&](const buf_t& in, buf_t& out, int cols) {
        sumFactSweep(in.data(), out.data(), nb, nq, cols, t.interp.data(), false, false, oe);
    };
    const auto D = [&](const buf_t& in, buf_t& out, int cols) {
        sumFactSweep(in.data(), out.data(), nb, nq, cols, t.der.data(), true, false, oe);
    };
    if (dim == 2)
    {
        std::copy_n(fill, fill_size, r[1].begin());
        const int c0 = F * nb, c1 = F * nq;
        I(r[1], temp, c0);
        D(temp, r[2], c1);
        I(temp, r[0], c1);
        D(r[1], temp, c0);
        I(temp, r[1], c1);
    }
    else
    {
        std::copy_n(fill, fill_size, r[3].begin());
        const int c0 = F * nb * nb, c1 = F * nq * nb, c2 = F * nq * nq;
        I(r[3], r[0], c0);
        D(r[3], r[1], c0);
        I(r[1], r[3], c1);
        I(r[3], r[1], c2);
        D(r[0], r[3], c1);
        I(r[3], r[2], c2);
        I(r[0], temp, c1);
        I(temp, r[0], c2);
        D(temp, r[3], c2);
    }
    return r;
}

// sumFactForwardQuad (:758-782) / sumFactForwardHex (:784-814); result in ts[0], row-major n_nodes x F
void sumFactForward(const SumFactTables& t, int dim, int F, std::array< buf_t, 4 >& ts, buf_t& temp)
{
    const int  nb = t.nb, nq = t.nq;
    const bool oe = t.odd_even;
    const auto IA = [&](const buf_t& in, buf_t& out, int cols) {
        sumFactSweep(in.data(), out.data(), nq, nb, cols, t.interp_t.data(), false, false, oe);
    };
    const auto DA = [&](const buf_t& in, buf_t& out, int cols) {
        sumFactSweep(in.data(), out.data(), nq, nb, cols, t.der_t.data(), true, true, oe);
    };
    if (dim == 2)
    {
        const int c0 = F * nq, c1 = F * nb;
        IA(ts[0], temp, c0);
        DA(ts[1], temp, c0);
        IA(temp, ts[0], c1);
        IA(ts[2], temp, c0);
        DA(temp, ts[0], c1);
    }
    else
    {
        const int c0 = F * nq * nq, c1 = F * nb * nq, c2 = F * nb * nb;
        IA(ts[0], temp, c0);
        DA(ts[1], temp, c0);
        IA(temp, ts[1], c1);
        IA(ts[2], temp, c0);
        DA(temp, ts[1], c1);
        IA(ts[1], ts[0], c2);
        IA(ts[3], ts[1], c0);
        IA(ts[1], ts[3], c1);
        DA(ts[3], ts[0], c2);
    }
}
} // namespace

// evalLocalOperatorSumFact (:882-917) → sumFactImpl (:816-868) → evalAtQuadQPs / evalAtHexQPs (:614-756)
// DIM_T, E_T, U_T: compile-time sizes (0 = read them from the kernel). The reference's sizes are all template parameters; the
// instantiations listed in evalLocalOperatorSumFact give the compiler the same knowledge for the configurations that are timed.
namespace
{
template < int DIM_T, int E_T, int U_T >
void evalSumFactSized(const Kernel&          kernel,
                      ElementType            et,
                      int                    order,
                      const val_t*           verts,
                      const AssemblyOptions& opts,
                      val_t                  time,
                      int                    n_rhs_actual,
                      const val_t*           X,
                      val_t*                 Y)
{
    const int  dim = DIM_T > 0 ? DIM_T : nativeDim(et);
    const int  E = E_T > 0 ? E_T : kernel.params.n_equations, U = U_T > 0 ? U_T : kernel.params.n_unknowns, NF = kernel.params.n_fields;
    const int  n_ops   = U * n_rhs_actual;
    const int  F_total = n_ops + NF;
    const int  quad_order = 2 * opts.order(order); // make_basis_params (:425-430)
    const bool oe         = opts.useOddEven(order);
    // the reference's tables are compile-time constants (:25-65); rebuilding them per element (the 1-D basis is evaluated through the
    // generic polynomial code) cost more than the element itself, which made the CPU baseline unfairly slow: one copy per thread
    const auto cachedTables = [](int basis_order, int qo, bool odd_even) -> const SumFactTables& {
        thread_local std::map< std::tuple< int, int, bool >, SumFactTables > cache;
        const auto key = std::make_tuple(basis_order, qo, odd_even);
        auto       it  = cache.find(key);
        if (it == cache.end())
            it = cache.emplace(key, makeSumFactTables(basis_order, qo, odd_even)).first;
        return it->second;
    };
    const auto& tab  = cachedTables(order, quad_order, oe);
    const auto& gtab = cachedTables(1, quad_order, oe); // make_geom_basis_params (:431-436)
    const int  nq = tab.nq, n_nodes = numNodes(et, order), Q = ipow(nq, dim), nv = 1 << dim;

    auto back = sumFactBack(tab, dim, F_total, X, static_cast< std::size_t >(n_nodes) * F_total);

    // computeGeomDataLin (:506-537): coordinate "fields", vertex-major
    buf_t gfill(static_cast< std::size_t >(nv) * dim);
    for (int v = 0; v < nv; ++v)
        for (int s = 0; s < dim; ++s)
            gfill[v + static_cast< std::size_t >(s) * nv] = verts[v * 3 + s];
    const auto geom = sumFactBack(gtab, dim, dim, gfill.data(), gfill.size());

    std::array< buf_t, 4 > fwd;
    for (auto& b : fwd)
        b.assign(static_cast< std::size_t >(std::max(Q, n_nodes)) * n_ops, 0.);

    std::vector< val_t > A(static_cast< std::size_t >(dim + 1) * E * U), Fk(static_cast< std::size_t >(E) * kernel.params.n_rhs),
        Dm(static_cast< std::size_t >(dim) * E * U), tvec(static_cast< std::size_t >(E) * n_rhs_actual), fvals(std::max(NF, 1)),
        fders(static_cast< std::size_t >(3) * std::max(NF, 1));
    for (int qi = 0; qi < Q; ++qi)
    {
        // jacobian_mat(s, d) = d x_s / d xi_d   (:649, :716-724)
        val_t Jm[9], Ji[9];
        for (int s = 0; s < dim; ++s)
            for (int d = 0; d < dim; ++d)
                Jm[s * dim + d] = geom[1 + d][static_cast< std::size_t >(qi) * dim + s];
        inverse(dim, Jm, Ji);
        // evalFieldVals / evalFieldDers (:572-612): the rightmost n_fields columns
        for (int f = 0; f < NF; ++f)
        {
            fvals[f] = back[0][static_cast< std::size_t >(qi) * F_total + n_ops + f];
            for (int s = 0; s < dim; ++s)
            {
                val_t acc = 0.;
                for (int d = 0; d < dim; ++d)
                    acc += Ji[d * dim + s] * back[1 + d][static_cast< std::size_t >(qi) * F_total + n_ops + f];
                fders[static_cast< std::size_t >(s) * NF + f] = acc;
            }
        }
        KernelInput in{};
        in.field_vals = fvals.data();
        for (int d = 0; d < 3; ++d)
            in.field_ders[d] = fders.data() + static_cast< std::size_t >(d) * NF;
        // NOTE the reference passes z = 0 also for hexes (:656, :732) — replicated on purpose
        in.space[0] = geom[0][static_cast< std::size_t >(qi) * dim + 0];
        in.space[1] = geom[0][static_cast< std::size_t >(qi) * dim + 1];
        in.space[2] = 0.;
        in.time     = time;
        std::fill(A.begin(), A.end(), 0.);
        std::fill(Fk.begin(), Fk.end(), 0.);
        KernelOutput out{};
        for (int i = 0; i <= dim; ++i)
            out.operators[i] = Op{A.data() + static_cast< std::size_t >(i) * E * U, E};
        out.rhs = RhsView{Fk.data(), E};
        kernel.fn(in, out);
        // D_d = sum_s A_{s+1} * Ji(d, s)   (:660-661, :736-738)
        for (int d = 0; d < dim; ++d)
            for (int k = 0; k < E * U; ++k)
            {
                val_t acc = 0.;
                for (int s = 0; s < dim; ++s)
                    acc += A[static_cast< std::size_t >(s + 1) * E * U + k] * Ji[d * dim + s];
                Dm[static_cast< std::size_t >(d) * E * U + k] = acc;
            }
        // weights: wx*wy[*wz]*detJ, with qi = (qz*nq + qy)*nq + qx
        const int   qx = qi % nq, qy = (qi / nq) % nq, qz = qi / (nq * nq);
        const val_t jac = det(dim, Jm);
        val_t       wgt = dim == 2 ? tab.weights[qx] * tab.weights[qy] * jac : tab.weights[qx] * tab.weights[qy] * tab.weights[qz] * jac;
        // t = wgt * (A0 t0 + sum_d D_d t_d), operands are U x n_rhs col-major maps onto the Q-row of each back buffer
        for (int r = 0; r < n_rhs_actual; ++r)
            for (int e = 0; e < E; ++e)
            {
                val_t acc = 0.;
                for (int u = 0; u < U; ++u)
                {
                    acc += A[e + static_cast< std::size_t >(u) * E] * back[0][static_cast< std::size_t >(qi) * F_total + r * U + u];
                    for (int d = 0; d < dim; ++d)
                        acc += Dm[static_cast< std::size_t >(d) * E * U + e + static_cast< std::size_t >(u) * E] *
                               back[1 + d][static_cast< std::size_t >(qi) * F_total + r * U + u];
                }
                tvec[e + static_cast< std::size_t >(r) * E] = wgt * acc;
            }
        // r0 = A0^T t, r_d = D_d^T t, written col-major Q x n_ops   (:669-671, :747-750)
        for (int r = 0; r < n_rhs_actual; ++r)
            for (int u = 0; u < U; ++u)
            {
                val_t acc0 = 0.;
                for (int e = 0; e < E; ++e)
                    acc0 += A[e + static_cast< std::size_t >(u) * E] * tvec[e + static_cast< std::size_t >(r) * E];
                fwd[0][qi + static_cast< std::size_t >(Q) * (r * U + u)] = acc0;
                for (int d = 0; d < dim; ++d)
                {
                    val_t acc = 0.;
                    for (int e = 0; e < E; ++e)
                        acc += Dm[static_cast< std::size_t >(d) * E * U + e + static_cast< std::size_t >(u) * E] *
                               tvec[e + static_cast< std::size_t >(r) * E];
                    fwd[1 + d][qi + static_cast< std::size_t >(Q) * (r * U + u)] = acc;
                }
            }
    }
    buf_t temp(static_cast< std::size_t >(std::max(Q, n_nodes)) * n_ops, 0.);
    sumFactForward(tab, dim, n_ops, fwd, temp);
    std::copy_n(fwd[0].begin(), static_cast< std::size_t >(n_nodes) * n_ops, Y);
}
} // namespace

void evalLocalOperatorSumFact(const Kernel&          kernel,
                              ElementType            et,
                              int                    order,
                              const val_t*           verts,
                              const AssemblyOptions& opts,
                              val_t                  time,
                              int                    n_rhs_actual,
                              const val_t*           X,
                              val_t*                 Y)
{
    const int dim = nativeDim(et);
    if (kernel.params.dimension != dim or dim < 2)
        throw std::invalid_argument{"sum factorisation needs a quad/hex element matching the kernel dimension"};
    const int E = kernel.params.n_equations, U = kernel.params.n_unknowns;
    if (dim == 3 and E == 7 and U == 4) // benchmarks/Diffusion3D.hpp:50-79, the configuration bench.py times
        evalSumFactSized< 3, 7, 4 >(kernel, et, order, verts, opts, time, n_rhs_actual, X, Y);
    else if (dim == 2 and E == 4 and U == 3) // tests/Kernels.hpp diffusion 2D, examples/02
        evalSumFactSized< 2, 4, 3 >(kernel, et, order, verts, opts, time, n_rhs_actual, X, Y);
    else if (dim == 3 and E == 8 and U == 7) // benchmarks/Kernels.hpp ns3d
        evalSumFactSized< 3, 8, 7 >(kernel, et, order, verts, opts, time, n_rhs_actual, X, Y);
    else
        evalSumFactSized< 0, 0, 0 >(kernel, et, order, verts, opts, time, n_rhs_actual, X, Y);
}
} // namespace orc
