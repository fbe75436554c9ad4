#include <tuple>
// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp). mapping/ and element-local algsys/ restatement.
#include "l3ster_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace orc
{
// mapping/JacobiMat.hpp:17-45 — J(d, s) = sum_v X_v[s] * dN1_v/dxi_d, order-1 Lagrange shape functions, vertices only
void jacobiMat(ElementType et, const val_t* verts, const val_t* point, val_t* J)
{
    const int dim = nativeDim(et), nv = 1 << dim;
    std::fill_n(J, dim * dim, 0.);
    for (int v = 0; v < nv; ++v)
        for (int d = 0; d < dim; ++d)
        {
            const val_t sf = refBasisDer(et, 1, v, d, point);
            for (int s = 0; s < dim; ++s)
                J[d * dim + s] += verts[v * 3 + s] * sf;
        }
}

val_t det(int dim, const val_t* M)
{
    if (dim == 1)
        return M[0];
    if (dim == 2)
        return M[0] * M[3] - M[1] * M[2];
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// cofactor inverse, as Eigen's fixed-size inverse() does for 2x2 / 3x3 (mapping/ComputePhysBasisDer.hpp:14)
void inverse(int dim, const val_t* M, val_t* R)
{
    if (dim == 1)
    {
        R[0] = 1. / M[0];
        return;
    }
    if (dim == 2)
    {
        const val_t invdet = 1. / det(2, M);
        R[0]               = M[3] * invdet;
        R[1]               = -M[1] * invdet;
        R[2]               = -M[2] * invdet;
        R[3]               = M[0] * invdet;
        return;
    }
    const val_t c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const val_t d      = M[0] * c00 + M[1] * c01 + M[2] * c02;
    const val_t invdet = 1. / d;
    R[0]               = c00 * invdet;
    R[1]               = (M[2] * M[7] - M[1] * M[8]) * invdet;
    R[2]               = (M[1] * M[5] - M[2] * M[4]) * invdet;
    R[3]               = c01 * invdet;
    R[4]               = (M[0] * M[8] - M[2] * M[6]) * invdet;
    R[5]               = (M[2] * M[3] - M[0] * M[5]) * invdet;
    R[6]               = c02 * invdet;
    R[7]               = (M[1] * M[6] - M[0] * M[7]) * invdet;
    R[8]               = (M[0] * M[4] - M[1] * M[3]) * invdet;
}

// mapping/MapReferenceToPhysical.hpp:14-25 + basisfun/ValueAt.hpp:8-19
void mapToPhysicalSpace(ElementType et, const val_t* verts, const val_t* point, val_t* out3)
{
    const int nv = 1 << nativeDim(et);
    for (int dim = 0; dim < 3; ++dim)
    {
        val_t acc = 0.;
        for (int v = 0; v < nv; ++v)
        {
            const val_t term = verts[v * 3 + dim] * refBasisValue(et, 1, v, point);
            acc              = v == 0 ? term : acc + term;
        }
        out3[dim] = acc;
    }
}

// mapping/ReferenceBoundaryToSideMapping.hpp:14-48 with math/RotationMatrix.hpp:11-76
void refBoundaryToSide(ElementType et, int side, val_t* rot, val_t* trans)
{
    const val_t pi = M_PI;
    if (et == Hex)
    {
        const auto rotX = [&](val_t a) {
            const val_t s = std::sin(a), c = std::cos(a);
            const val_t m[9] = {1., 0., 0., 0., c, -s, 0., s, c};
            std::copy_n(m, 9, rot);
        };
        const auto rotY = [&](val_t a) {
            const val_t s = std::sin(a), c = std::cos(a);
            const val_t m[9] = {c, 0., s, 0., 1., 0., -s, 0., c};
            std::copy_n(m, 9, rot);
        };
        const auto setT = [&](val_t x, val_t y, val_t z) {
            trans[0] = x;
            trans[1] = y;
            trans[2] = z;
        };
        switch (side)
        {
        case 0:
            rotX(pi);
            setT(0., 0., -1.);
            break;
        case 1: {
            const val_t m[9] = {1., 0., 0., 0., 1., 0., 0., 0., 1.};
            std::copy_n(m, 9, rot);
            setT(0., 0., 1.);
            break;
        }
        case 2:
            rotX(-pi / 2.);
            setT(0., -1., 0.);
            break;
        case 3:
            rotX(pi / 2.);
            setT(0., 1., 0.);
            break;
        case 4:
            rotY(pi / 2.);
            setT(-1., 0., 0.);
            break;
        case 5:
            rotY(-pi / 2.);
            setT(1., 0., 0.);
            break;
        default:
            throw std::out_of_range{"hex side"};
        }
    }
    else if (et == Quad)
    {
        const auto rot2 = [&](val_t a) {
            const val_t s = std::sin(a), c = std::cos(a);
            rot[0] = c;
            rot[1] = s;
            rot[2] = -s;
            rot[3] = c;
        };
        switch (side)
        {
        case 0:
            rot2(pi);
            trans[0] = 0.;
            trans[1] = -1.;
            break;
        case 1:
            rot[0] = 1.;
            rot[1] = 0.;
            rot[2] = 0.;
            rot[3] = 1.;
            trans[0] = 0.;
            trans[1] = 1.;
            break;
        case 2:
            rot2(pi / 2);
            trans[0] = -1.;
            trans[1] = 0.;
            break;
        case 3:
            rot2(-pi / 2);
            trans[0] = 1.;
            trans[1] = 0.;
            break;
        default:
            throw std::out_of_range{"quad side"};
        }
    }
    else
    {
        rot[0]   = side == 0 ? -1. : 1.;
        trans[0] = side == 0 ? -1. : 1.;
    }
}

// mapping/BoundaryIntegralJacobian.hpp:9-28
val_t boundaryIntegralJacobian(ElementType et, int side, const val_t* J)
{
    const int dim = nativeDim(et);
    if (dim == 1)
        return 0.;
    val_t rot[9], trans[3];
    refBoundaryToSide(et, side, rot, trans);
    if (dim == 2)
    {
        // (J^T * rot.col(0)).norm()
        const val_t v0 = J[0] * rot[0] + J[2] * rot[2];
        const val_t v1 = J[1] * rot[0] + J[3] * rot[2];
        return std::sqrt(v0 * v0 + v1 * v1);
    }
    val_t a[3], b[3];
    for (int s = 0; s < 3; ++s)
    {
        a[s] = J[0 * 3 + s] * rot[0 * 3 + 0] + J[1 * 3 + s] * rot[1 * 3 + 0] + J[2 * 3 + s] * rot[2 * 3 + 0];
        b[s] = J[0 * 3 + s] * rot[0 * 3 + 1] + J[1 * 3 + s] * rot[1 * 3 + 1] + J[2 * 3 + s] * rot[2 * 3 + 1];
    }
    const val_t c0 = a[1] * b[2] - a[2] * b[1], c1 = a[2] * b[0] - a[0] * b[2], c2 = a[0] * b[1] - a[1] * b[0];
    return std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
}

// mapping/BoundaryNormal.hpp:8-64
void boundaryNormal(ElementType et, int side, const val_t* J, val_t* n)
{
    if (et == Line)
    {
        n[0] = side == 0 ? -1. : 1.;
        return;
    }
    if (et == Quad)
    {
        switch (side)
        {
        case 0:
            n[0] = J[1];
            n[1] = -J[0];
            break;
        case 1:
            n[0] = -J[1];
            n[1] = J[0];
            break;
        case 2:
            n[0] = -J[3];
            n[1] = J[2];
            break;
        case 3:
            n[0] = J[3];
            n[1] = -J[2];
            break;
        }
        const val_t nrm = std::sqrt(n[0] * n[0] + n[1] * n[1]);
        n[0] /= nrm;
        n[1] /= nrm;
        return;
    }
    const auto cross = [&](int r0, int r1, val_t sign) {
        const val_t* a = J + 3 * r0;
        const val_t* b = J + 3 * r1;
        n[0]           = sign * (a[1] * b[2] - a[2] * b[1]);
        n[1]           = sign * (a[2] * b[0] - a[0] * b[2]);
        n[2]           = sign * (a[0] * b[1] - a[1] * b[0]);
    };
    switch (side)
    {
    case 0:
        cross(0, 1, -1.);
        break;
    case 1:
        cross(0, 1, 1.);
        break;
    case 2:
        cross(0, 2, 1.);
        break;
    case 3:
        cross(0, 2, -1.);
        break;
    case 4:
        cross(1, 2, -1.);
        break;
    case 5:
        cross(1, 2, 1.);
        break;
    }
    const val_t nrm = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int i = 0; i < 3; ++i)
        n[i] /= nrm;
}

namespace
{
// Everything the reference computes per quadrature point before the kernel result is consumed:
// map::mapDomain / mapBoundary (mapping/MapReferenceToPhysical.hpp:28-89) + evalKernel (AssembleLocalSystem.hpp:218-232)
struct QpEval
{
    int                  dim, n_bases, E, U, NF, NRHS;
    std::vector< val_t > phys_ders; // dim x n_bases (row d contiguous)
    std::vector< val_t > A;         // (dim+1) operators, each E x U col-major
    std::vector< val_t > F;         // E x NRHS col-major
    std::vector< val_t > fvals, fders;
    val_t                jacobian = 0.; // det J (domain) or surface measure (boundary)
    val_t                detJ     = 0.;

    QpEval(const Kernel& k, int n_bases_)
        : dim{k.params.dimension},
          n_bases{n_bases_},
          E{k.params.n_equations},
          U{k.params.n_unknowns},
          NF{k.params.n_fields},
          NRHS{k.params.n_rhs},
          phys_ders(static_cast< std::size_t >(dim) * n_bases_),
          A(static_cast< std::size_t >(dim + 1) * E * U),
          F(static_cast< std::size_t >(E) * NRHS),
          fvals(std::max(NF, 1)),
          fders(static_cast< std::size_t >(std::max(NF, 1)) * 3)
    {}

    void eval(const Kernel&         kernel,
              ElementType           et,
              const val_t*          verts,
              const val_t*          node_vals,
              const RefBasisAtQuad& rbq,
              int                   q,
              val_t                 time,
              int                   side)
    {
        const val_t* point    = &rbq.quad.points[static_cast< std::size_t >(q) * dim];
        const val_t* bvals    = &rbq.values[static_cast< std::size_t >(q) * n_bases];
        const val_t* ref_ders = &rbq.derivatives[static_cast< std::size_t >(q) * dim * n_bases];
        val_t        J[9], Ji[9];
        jacobiMat(et, verts, point, J);
        inverse(dim, J, Ji);
        // phys = J^-1 * ref_ders   (mapping/ComputePhysBasisDer.hpp:9-15)
        for (int r = 0; r < dim; ++r)
            for (int a = 0; a < n_bases; ++a)
            {
                val_t acc = 0.;
                for (int k = 0; k < dim; ++k)
                    acc += Ji[r * dim + k] * ref_ders[k * n_bases + a];
                phys_ders[static_cast< std::size_t >(r) * n_bases + a] = acc;
            }
        detJ = det(dim, J);
        KernelInput in{};
        if (side < 0)
            jacobian = detJ;
        else
        {
            jacobian = boundaryIntegralJacobian(et, side, J);
            boundaryNormal(et, side, J, in.normal);
        }
        // computeFieldVals / computeFieldDers (AssembleLocalSystem.hpp:54-75)
        for (int f = 0; f < NF; ++f)
        {
            val_t v = 0.;
            for (int a = 0; a < n_bases; ++a)
                v += node_vals[static_cast< std::size_t >(a) * NF + f] * bvals[a];
            fvals[f] = v;
            for (int d = 0; d < dim; ++d)
            {
                val_t g = 0.;
                for (int a = 0; a < n_bases; ++a)
                    g += phys_ders[static_cast< std::size_t >(d) * n_bases + a] * node_vals[static_cast< std::size_t >(a) * NF + f];
                fders[static_cast< std::size_t >(d) * NF + f] = g;
            }
        }
        in.field_vals = fvals.data();
        for (int d = 0; d < 3; ++d)
            in.field_ders[d] = fders.data() + static_cast< std::size_t >(d) * NF;
        mapToPhysicalSpace(et, verts, point, in.space);
        in.time = time;
        // initKernelResult (common/KernelInterface.hpp:61-69)
        std::fill(A.begin(), A.end(), 0.);
        std::fill(F.begin(), F.end(), 0.);
        KernelOutput out{};
        for (int i = 0; i <= dim; ++i)
            out.operators[i] = Op{A.data() + static_cast< std::size_t >(i) * E * U, E};
        out.rhs = RhsView{F.data(), E};
        kernel.fn(in, out);
    }

    // block_a(u, e) = N_a * A0(e,u) + sum_d dN_a/dx_d * A_{d+1}(e,u)    (AssembleLocalSystem.hpp:131-142)
    val_t block(const val_t* bvals, int a, int u, int e) const
    {
        val_t v = bvals[a] * A[static_cast< std::size_t >(e) + static_cast< std::size_t >(u) * E];
        for (int d = 0; d < dim; ++d)
            v += phys_ders[static_cast< std::size_t >(d) * n_bases + a] *
                 A[static_cast< std::size_t >(d + 1) * E * U + e + static_cast< std::size_t >(u) * E];
        return v;
    }
};

constexpr int max_unknowns_fill = 16; // unknowns per kernel the batch fill of assembleLocalSystem keeps on the stack

void checkDims(const Kernel& kernel, ElementType et)
{
    if (kernel.params.dimension != nativeDim(et))
        throw std::invalid_argument{"kernel dimension does not match the element's native dimension"};
}

// K_lower += sign * B * B^T, B col-major L x ncols (Eigen selfadjointView<Lower>::rankUpdate, AssembleLocalSystem.hpp:191-208).
// Register-tiled so that the multi-threaded CPU baseline is a fair stand-in for Eigen's SYRK: a 4 x 16 accumulator tile (measured best of
// 3 x 16 ... 14 x 16, 4 x 32, 8 x 24 with 256- and 512-bit vectors: 38.7 GFLOP/s per core with -mprefer-vector-width=512, which the
// -march=native build of oracle/__init__.py passes; 25.0 with GCC's default 256-bit preference on the same AVX-512 core), the transposed
// panel in a per-thread buffer, no edge conditionals in the k-loop (rows are padded with zeros).
constexpr int syrk_rt = 4, syrk_ct = 16;
void syrkLower(val_t* K, int L, const val_t* B, int ncols, val_t sign)
{
    constexpr int RT = syrk_rt, CT = syrk_ct;
    const int     Lp = (L + CT - 1) / CT * CT + RT; // padded: a row tile starting below L may read up to RT - 1 rows past it
    thread_local std::vector< val_t > Bt;           // Bt[k][r]
    Bt.assign(static_cast< std::size_t >(ncols) * Lp, 0.);
    for (int k = 0; k < ncols; ++k)
        for (int r = 0; r < L; ++r)
            Bt[static_cast< std::size_t >(k) * Lp + r] = B[static_cast< std::size_t >(r) + static_cast< std::size_t >(k) * L];
    for (int r0 = 0; r0 < L; r0 += RT)
    {
        const int rn = std::min(RT, L - r0);
        for (int c0 = 0; c0 <= r0 + rn - 1; c0 += CT)
        {
            const int cn = std::min(CT, L - c0);
            val_t     acc[RT][CT] = {};
            const val_t* __restrict__ bk = Bt.data();
            for (int k = 0; k < ncols; ++k, bk += Lp)
            {
                const val_t* __restrict__ bc = bk + c0;
                const val_t* __restrict__ br = bk + r0;
#pragma GCC unroll 8
                for (int i = 0; i < RT; ++i)
                {
                    const val_t bri = br[i];
#pragma omp simd
                    for (int j = 0; j < CT; ++j)
                        acc[i][j] += bri * bc[j];
                }
            }
            for (int i = 0; i < rn; ++i)
                for (int j = 0; j < cn; ++j)
                    if (c0 + j <= r0 + i)
                        K[static_cast< std::size_t >(r0 + i) * L + c0 + j] += sign * acc[i][j];
        }
    }
}
} // namespace

// algsys/AssembleLocalSystem.hpp:77-216, 234-280
void assembleLocalSystem(const Kernel&         kernel,
                         ElementType           et,
                         int                   order,
                         const val_t*          verts,
                         const val_t*          node_vals,
                         const RefBasisAtQuad& rbq,
                         val_t                 time,
                         int                   side,
                         val_t*                K,
                         val_t*                Fout)
{
    checkDims(kernel, et);
    const int n_bases = numNodes(et, order);
    const int E = kernel.params.n_equations, U = kernel.params.n_unknowns, NRHS = kernel.params.n_rhs;
    const int L = n_bases * U;
    if (U > max_unknowns_fill)
        throw std::invalid_argument{"more unknowns per kernel than the oracle's batch fill keeps on the stack"};
    // LocalSystemManager batching (:80-82): target_update_size = 16 * simd_width / sizeof(val_t), simd_width = 32
    constexpr int target_update_size = 16 * 32 / 8;
    const int     updates_per_batch  = (target_update_size + E - 1) / E;
    const int     batch_cols         = E * updates_per_batch;
    std::vector< val_t > posw(static_cast< std::size_t >(L) * batch_cols), negw(posw.size());
    int                  pos_n = 0, neg_n = 0;
    std::fill_n(K, static_cast< std::size_t >(L) * L, 0.);
    std::fill_n(Fout, static_cast< std::size_t >(L) * NRHS, 0.);
    QpEval     qp{kernel, n_bases};
    const auto flush = [&](std::vector< val_t >& buf, int& n, val_t sign) {
        if (n > 0)
            syrkLower(K, L, buf.data(), n * E, sign);
        n = 0;
    };
    for (int q = 0; q < rbq.quad.size; ++q)
    {
        qp.eval(kernel, et, verts, node_vals, rbq, q, time, side);
        if (side < 0 and not(qp.jacobian > 0.))
            throw std::runtime_error{"Encountered degenerate element ( |J| <= 0 )"};
        const val_t  weight   = qp.jacobian * rbq.quad.weights[q];
        const bool   positive = weight >= 0.;
        auto&        buf      = positive ? posw : negw;
        int&         bn       = positive ? pos_n : neg_n;
        const val_t  wsqrt    = std::sqrt(std::fabs(weight));
        const val_t* bvals    = &rbq.values[static_cast< std::size_t >(q) * n_bases];
        // equation-major: column (bn E + e) of the batch is written top to bottom (rows a U + u), the order Eigen's column-major
        // block assignment takes (:159-166); per row the contributions to F_e still arrive in ascending e
        for (int e = 0; e < E; ++e)
        {
            val_t* __restrict__ col = &buf[static_cast< std::size_t >(bn * E + e) * L];
            // row e of the operators, per unknown: the same products and the same order as QpEval::block
            val_t       ce[4][max_unknowns_fill];
            const int   dim = qp.dim;
            for (int u = 0; u < U; ++u)
                for (int i = 0; i <= dim; ++i)
                    ce[i][u] = qp.A[static_cast< std::size_t >(i) * E * U + e + static_cast< std::size_t >(u) * E];
            const val_t* __restrict__ pd = qp.phys_ders.data();
            for (int a = 0; a < n_bases; ++a)
            {
                const val_t n = bvals[a];
                val_t       g[3] = {0., 0., 0.};
                for (int d = 0; d < dim; ++d)
                    g[d] = pd[static_cast< std::size_t >(d) * n_bases + a];
                for (int u = 0; u < U; ++u)
                {
                    const int row = a * U + u;
                    val_t     b   = n * ce[0][u];
                    for (int d = 0; d < dim; ++d)
                        b += g[d] * ce[d + 1][u];
                    for (int r = 0; r < NRHS; ++r)
                        Fout[static_cast< std::size_t >(row) + static_cast< std::size_t >(r) * L] +=
                            b * qp.F[static_cast< std::size_t >(e) + static_cast< std::size_t >(r) * E] * weight;
                    col[row] = b * wsqrt;
                }
            }
        }
        if (++bn == updates_per_batch)
            flush(buf, bn, positive ? 1. : -1.);
    }
    flush(posw, pos_n, 1.);
    flush(negw, neg_n, -1.);
    // m_system->first = selfadjointView<Lower>()   (:180)
    for (int r = 0; r < L; ++r)
        for (int c = r + 1; c < L; ++c)
            K[static_cast< std::size_t >(r) * L + c] = K[static_cast< std::size_t >(c) * L + r];
}

// post/Integral.hpp:11-53: sum_q weight_q * jacobian_q * kernel(input_q); the residual kernel writes its Rhs (E x n_rhs)
void evalElementIntegral(const Kernel&         kernel,
                         ElementType           et,
                         int                   order,
                         const val_t*          verts,
                         const val_t*          node_vals,
                         const RefBasisAtQuad& rbq,
                         val_t                 time,
                         int                   side,
                         bool                  squared,
                         val_t*                out)
{
    checkDims(kernel, et);
    const int nv = kernel.params.n_equations * kernel.params.n_rhs;
    std::fill_n(out, nv, 0.);
    QpEval qp{kernel, numNodes(et, order)};
    for (int q = 0; q < rbq.quad.size; ++q)
    {
        qp.eval(kernel, et, verts, node_vals, rbq, q, time, side);
        for (int i = 0; i < nv; ++i)
        {
            const val_t r = squared ? qp.F[i] * qp.F[i] : qp.F[i]; // post/NormL2.hpp:21-28
            out[i] += qp.jacobian * r * rbq.quad.weights[q];
        }
    }
}

// algsys/EvaluateLocalOperator.hpp:94-146, 211-263
void evaluateLocalOperator(const Kernel&         kernel,
                           ElementType           et,
                           int                   order,
                           const val_t*          verts,
                           const val_t*          node_vals,
                           const RefBasisAtQuad& rbq,
                           val_t                 time,
                           int                   side,
                           int                   n_cols,
                           const val_t*          x,
                           val_t*                y)
{
    checkDims(kernel, et);
    const int n_bases = numNodes(et, order);
    const int E = kernel.params.n_equations, U = kernel.params.n_unknowns;
    const int L = n_bases * U;
    std::fill_n(y, static_cast< std::size_t >(L) * n_cols, 0.);
    QpEval               qp{kernel, n_bases};
    std::vector< val_t > H(static_cast< std::size_t >(L) * E), t(E);
    for (int q = 0; q < rbq.quad.size; ++q)
    {
        qp.eval(kernel, et, verts, node_vals, rbq, q, time, side);
        if (side < 0 and not(qp.jacobian > 0.))
            throw std::runtime_error{"Encountered degenerate element ( |J| <= 0 )"};
        const val_t  weight = qp.jacobian * rbq.quad.weights[q];
        const val_t* bvals  = &rbq.values[static_cast< std::size_t >(q) * n_bases];
        for (int e = 0; e < E; ++e)
            for (int a = 0; a < n_bases; ++a)
                for (int u = 0; u < U; ++u)
                    H[static_cast< std::size_t >(a * U + u) + static_cast< std::size_t >(e) * L] = qp.block(bvals, a, u, e);
        for (int c = 0; c < n_cols; ++c)
        {
            const val_t* xc = x + static_cast< std::size_t >(c) * L;
            val_t*       yc = y + static_cast< std::size_t >(c) * L;
            for (int e = 0; e < E; ++e)
            {
                val_t acc = 0.;
                for (int r = 0; r < L; ++r)
                    acc += H[static_cast< std::size_t >(r) + static_cast< std::size_t >(e) * L] * xc[r];
                t[e] = acc * weight;
            }
            for (int e = 0; e < E; ++e)
                for (int r = 0; r < L; ++r)
                    yc[r] += H[static_cast< std::size_t >(r) + static_cast< std::size_t >(e) * L] * t[e];
        }
    }
}

// algsys/EvaluateLocalOperator.hpp:172-208, 276-328
void precomputeOperatorDiagonalAndRhs(const Kernel&         kernel,
                                      ElementType           et,
                                      int                   order,
                                      const val_t*          verts,
                                      const val_t*          node_vals,
                                      const RefBasisAtQuad& rbq,
                                      val_t                 time,
                                      int                   side,
                                      int                   n_dirichlet,
                                      const int*            dirichlet_inds,
                                      const val_t*          dirichlet_vals,
                                      val_t*                diag,
                                      val_t*                rhs)
{
    checkDims(kernel, et);
    const int n_bases = numNodes(et, order);
    const int E = kernel.params.n_equations, U = kernel.params.n_unknowns, NRHS = kernel.params.n_rhs;
    const int L = n_bases * U;
    std::fill_n(diag, L, 0.);
    std::fill_n(rhs, static_cast< std::size_t >(L) * NRHS, 0.);
    QpEval               qp{kernel, n_bases};
    std::vector< val_t > H(static_cast< std::size_t >(L) * E), inter(static_cast< std::size_t >(E) * NRHS);
    for (int q = 0; q < rbq.quad.size; ++q)
    {
        qp.eval(kernel, et, verts, node_vals, rbq, q, time, side);
        if (side < 0 and not(qp.jacobian > 0.))
            throw std::runtime_error{"Encountered degenerate element ( |J| <= 0 )"};
        const val_t  weight = qp.jacobian * rbq.quad.weights[q];
        const val_t* bvals  = &rbq.values[static_cast< std::size_t >(q) * n_bases];
        for (int a = 0; a < n_bases; ++a)
            for (int u = 0; u < U; ++u)
            {
                const int row = a * U + u;
                val_t     sq  = 0.;
                for (int e = 0; e < E; ++e)
                {
                    const val_t b                                                                = qp.block(bvals, a, u, e);
                    H[static_cast< std::size_t >(row) + static_cast< std::size_t >(e) * L] = b;
                    sq += b * b;
                }
                diag[row] += sq * weight;
                for (int r = 0; r < NRHS; ++r)
                {
                    val_t acc = 0.;
                    for (int e = 0; e < E; ++e)
                        acc += H[static_cast< std::size_t >(row) + static_cast< std::size_t >(e) * L] *
                               qp.F[static_cast< std::size_t >(e) + static_cast< std::size_t >(r) * E];
                    rhs[static_cast< std::size_t >(row) + static_cast< std::size_t >(r) * L] += acc * weight;
                }
            }
        if (n_dirichlet > 0)
        {
            // intermediate = H(dirichlet_inds, :)^T * dirichlet_vals * weight ; rhs -= H * intermediate    (:204-206)
            for (int e = 0; e < E; ++e)
                for (int r = 0; r < NRHS; ++r)
                {
                    val_t acc = 0.;
                    for (int i = 0; i < n_dirichlet; ++i)
                        acc += H[static_cast< std::size_t >(dirichlet_inds[i]) + static_cast< std::size_t >(e) * L] *
                               dirichlet_vals[static_cast< std::size_t >(i) + static_cast< std::size_t >(r) * n_dirichlet];
                    inter[static_cast< std::size_t >(e) + static_cast< std::size_t >(r) * E] = acc * weight;
                }
            for (int r = 0; r < NRHS; ++r)
                for (int row = 0; row < L; ++row)
                {
                    val_t acc = 0.;
                    for (int e = 0; e < E; ++e)
                        acc += H[static_cast< std::size_t >(row) + static_cast< std::size_t >(e) * L] *
                               inter[static_cast< std::size_t >(e) + static_cast< std::size_t >(r) * E];
                    rhs[static_cast< std::size_t >(row) + static_cast< std::size_t >(r) * L] -= acc;
                }
        }
    }
}
// algsys/ComputeValuesAtNodes.hpp:316-369, 450-506 with detail::zeroOut (:92-110) and the averaging of :112-154, single rank
void computeValuesAtNodes(const Mesh& mesh, const Kernel& kernel, val_t time, const val_t* fields, const std::vector< int >& field_inds,
                          const std::vector< int >& boundary_ids, int dpn, const std::vector< int >& dof_inds, val_t* values)
{
    const int         nn = mesh.nodesPerElem(), NF = kernel.params.n_fields, E = kernel.params.n_equations, R = kernel.params.n_rhs;
    const std::size_t ld = mesh.n_nodes * static_cast< std::size_t >(dpn);
    const auto        rb = makeRefBasisAtNodes(mesh.et, mesh.order);
    std::vector< val_t >  node_vals(static_cast< std::size_t >(nn) * std::max(NF, 1));
    std::vector< double > n_contribs(ld, 0.);
    // (element, side, local nodes to visit)
    std::vector< std::tuple< std::size_t, int, std::vector< int > > > work;
    if (not kernel.is_boundary)
    {
        std::vector< int > all(nn);
        for (int a = 0; a < nn; ++a)
            all[a] = a;
        for (std::size_t e = 0; e < mesh.n_elems; ++e)
            work.emplace_back(e, -1, all);
    }
    else
        for (const auto& b : mesh.boundary)
            if (std::find(boundary_ids.begin(), boundary_ids.end(), b.domain_id) != boundary_ids.end())
                work.emplace_back(b.parent, b.side, sideNodeInds(mesh.et, mesh.order, b.side));
    for (const auto& [e, side, locals] : work) // zeroOut
        for (int a : locals)
            for (int eq = 0; eq < E; ++eq)
                for (int r = 0; r < R; ++r)
                    values[mesh.elem_nodes[e * nn + a] * dpn + dof_inds[eq] + r * ld] = 0.;
    QpEval qp{kernel, nn};
    for (const auto& [e, side, locals] : work)
    {
        const n_id_t* el_nodes = &mesh.elem_nodes[e * nn];
        for (int a = 0; a < nn; ++a)
            for (int f = 0; f < NF; ++f)
                node_vals[static_cast< std::size_t >(a) * NF + f] =
                    fields[el_nodes[a] + static_cast< std::size_t >(field_inds.empty() ? f : field_inds[f]) * mesh.n_nodes];
        for (int a : locals)
        {
            qp.eval(kernel, mesh.et, &mesh.elem_verts[e * (1u << nativeDim(mesh.et)) * 3], node_vals.data(), rb, a, time, side);
            for (int eq = 0; eq < E; ++eq)
            {
                const std::size_t dof = el_nodes[a] * dpn + dof_inds[eq];
                n_contribs[dof] += 1.;
                for (int r = 0; r < R; ++r)
                    values[dof + r * ld] += qp.F[eq + r * E];
            }
        }
    }
    for (std::size_t dof = 0; dof < ld; ++dof)
        if (n_contribs[dof] > 0.)
            for (int r = 0; r < R; ++r)
                values[dof + r * ld] /= n_contribs[dof];
}

} // namespace orc
