// Matrix-free system initialisation for domain kernels: diag[dofs(e)] += diag(K_e), rhs[dofs(e)] += F_e
// (precomputeOperatorDiagonalAndRhs, algsys/EvaluateLocalOperator.hpp:172-208, 276-328; scatter MatrixFreeSystem.hpp:377-390).
// The Dirichlet lifting rhs -= K_e g is applied by the caller as one unmasked operator apply on the vector of prescribed values
// (capi.cu: l3b_mf_end_assembly_begin), so this kernel never forms K_e g.
//
// With the coefficient form of the assembly kernel (assemble_dmma.cuh),
//     sqrt(w) B[e,(a,u)] = c0 N_a + sum_d c_d dN_a/dxi_d,   c0 = sqrt(w) A0(e,u),  c_d = sqrt(w) sum_s A_s(e,u) Jinv(s,d),
// the diagonal is a sum of squares, diag[(a,u)] = sum_q sum_e (sqrt(w) B[e,(a,u)])^2, and F_e[(a,u)] = sum_q rc(q,u) . (N_a, dN_a) with
// rc = sum_e sqrt(w) f_e c(e,u): per (point, node) four table reads and a few fp64 operations per structurally non-zero
// (equation, unknown) entry, no reductions across threads, no barriers inside the point loop. One CTA per element, one thread per
// node; the per-point stage (mapping, user kernel, coefficients) runs for 64 points at a time, one thread each.
// Replaces localElementKernel<MODE_INIT> (one CTA-wide reduction per point: 0.48 s at 64^3 hex p=4) for domain kernels.
#ifndef L3B_MF_INIT_CUH
#define L3B_MF_INIT_CUH

#include "assemble_dmma.cuh"

namespace l3b
{
template < typename KernelT, int DIM, int P >
struct MfInitCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields, NRHS = params.n_rhs;
    static constexpr int  NN      = cpow(P + 1, DIM);
    static constexpr int  threads = NN >= 256 ? 256 : ((NN + 31) / 32) * 32;
    static constexpr int  SP      = 64; // points per pass of the per-point stage
    // (unknown, equation) entries with a structurally non-zero B: their index in the coefficient table
    static constexpr int n_ent = [] {
        int n = 0;
        for (int u = 0; u < U; ++u)
            for (int eq = 0; eq < E; ++eq)
                n += unknownInEquation< KernelT >(u, eq);
        return n;
    }();
    static constexpr int entIndex(int u, int eq)
    {
        int n = 0;
        for (int uu = 0; uu < U; ++uu)
            for (int e = 0; e < E; ++e)
            {
                if (uu == u and e == eq)
                    return n;
                n += unknownInEquation< KernelT >(uu, e);
            }
        return n;
    }
    // smem (doubles): coefficients [SP][n_ent][4] | rhs coefficients [SP][U][NRHS][4] | node field values | vertices
    static constexpr int off_rc = SP * n_ent * 4, off_nv = off_rc + SP * U * NRHS * 4, off_verts = off_nv + NN * NF,
                         total = off_verts + 8 * 3;
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
};

template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(MfInitCfg< KernelT, DIM, P >::threads)
    mfInitDiagRhsKernel(const KernelT kernel, const __grid_constant__ ElemArgs args)
{
    using Cfg = MfInitCfg< KernelT, DIM, P >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NRHS = Cfg::NRHS, NN = Cfg::NN, T = Cfg::threads, SP = Cfg::SP, n_ent = Cfg::n_ent;
    constexpr int nv = 1 << DIM;
    static_assert(not KernelT::is_boundary);
    extern __shared__ double smem[];
    double* const s_c     = smem;
    double* const s_rc    = smem + Cfg::off_rc;
    double* const s_nv    = smem + Cfg::off_nv;
    double* const s_verts = smem + Cfg::off_verts;

    const int       tid = threadIdx.x;
    const long long wi  = blockIdx.x;
    const long long e   = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const uint32_t* el_nodes = args.nodes + e * NN;
    for (int i = tid; i < nv * 3; i += T)
        s_verts[i] = args.verts[e * nv * 3 + i];
    if constexpr (NF > 0)
        for (int i = tid; i < NN * NF; i += T)
            s_nv[i] = args.fields[el_nodes[i / NF] + args.field_inds[i % NF] * args.field_stride];
    __syncthreads();

    constexpr int NPT = (NN + T - 1) / T; // nodes per thread
    double        dacc[NPT][U], racc[NPT][U][NRHS];
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            dacc[i][u] = 0.;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                racc[i][u][r] = 0.;
        }

    for (int q0 = 0; q0 < args.n_qp; q0 += SP)
    {
        const int n_here = min(SP, args.n_qp - q0);
        // ---- per-point stage: one thread per point
        for (int qc = tid; qc < n_here; qc += T)
        {
            const int q = q0 + qc;
            double    xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM];
            for (int d = 0; d < DIM; ++d)
                xi[d] = args.tab_pts[q * DIM + d];
            geometryAt< DIM >(s_verts, xi, xs, Jt);
            const double detJ = invert< DIM >(Jt, Jti);
            if (not(detJ > 0.))
                atomicOr(args.status, status_degenerate_element); // EvaluateLocalOperator.hpp:295
            typename KernelT::Input in;
            if constexpr (NF > 0)
            {
                double        sv[NF], sd[DIM][NF];
                const double* bv = args.tab_vals + static_cast< long long >(q) * NN;
                const double* bd = args.tab_ders + static_cast< long long >(q) * DIM * NN;
                for (int f = 0; f < NF; ++f)
                {
                    sv[f] = 0.;
                    for (int d = 0; d < DIM; ++d)
                        sd[d][f] = 0.;
                }
                for (int a = 0; a < NN; ++a)
                {
                    const double n = bv[a];
                    double       dr[DIM];
                    for (int d = 0; d < DIM; ++d)
                        dr[d] = bd[d * NN + a];
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                    {
                        const double val = s_nv[a * NF + f];
                        sv[f]            = fma(n, val, sv[f]);
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            sd[d][f] = fma(dr[d], val, sd[d][f]);
                    }
                }
#pragma unroll
                for (int f = 0; f < NF; ++f)
                {
                    in.field_vals[f] = sv[f];
#pragma unroll
                    for (int s_ = 0; s_ < DIM; ++s_)
                    {
                        double acc = 0.;
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            acc = fma(Jti[s_][d], sd[d][f], acc);
                        in.field_ders[s_][f] = acc;
                    }
                }
            }
            for (int s_ = 0; s_ < 3; ++s_)
                in.point.space.coords[s_] = xs[s_]; // the non-SF paths pass the true 3-D point (AssembleLocalSystem.hpp:229)
            in.point.time  = args.time;
            const auto res = kernel(in);
            bool       violated = false;
            staticFor< DIM + 1 >([&](auto op) {
                staticFor< E >([&](auto eq) {
                    staticFor< U >([&](auto uu) {
                        if constexpr (not Sp::nz(op, eq, uu))
                            violated |= res.operators[op](eq, uu) != 0.;
                    });
                });
            });
            if (violated)
                atomicOr(args.status, status_sparsity_violation);
            const double sw = sqrt(fabs(detJ * args.tab_wts[q]));
            staticFor< U >([&](auto uu) {
                double rsum[NRHS][4];
                for (int r = 0; r < NRHS; ++r)
                    for (int j = 0; j < 4; ++j)
                        rsum[r][j] = 0.;
                staticFor< E >([&](auto eq) {
                    if constexpr (unknownInEquation< KernelT >(uu, eq))
                    {
                        double cf[4] = {0., 0., 0., 0.};
                        if constexpr (Sp::nz(0, eq, uu))
                            cf[0] = sw * res.operators[0](eq, uu);
                        staticFor< DIM >([&](auto s_) {
                            if constexpr (Sp::nz(s_ + 1, eq, uu))
                            {
#pragma unroll
                                for (int d = 0; d < DIM; ++d)
                                    cf[1 + d] = fma(sw * res.operators[s_ + 1](eq, uu), Jti[s_][d], cf[1 + d]);
                            }
                        });
                        constexpr int ent = Cfg::entIndex(uu, eq);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            s_c[(qc * n_ent + ent) * 4 + j] = cf[j];
#pragma unroll
                        for (int r = 0; r < NRHS; ++r)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                rsum[r][j] = fma(sw * res.rhs(eq, r), cf[j], rsum[r][j]);
                    }
                });
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        s_rc[((qc * U + uu) * NRHS + r) * 4 + j] = rsum[r][j];
            });
        }
        __syncthreads();
        // ---- accumulate: one thread per node, every point of the pass
#pragma unroll
        for (int i = 0; i < NPT; ++i)
        {
            const int a = tid + i * T;
            if (a >= NN)
                continue;
#pragma unroll 2
            for (int qc = 0; qc < n_here; ++qc)
            {
                const int q      = q0 + qc;
                double    bas[4] = {__ldg(args.tab_vals + q * NN + a), 0., 0., 0.};
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    bas[1 + d] = __ldg(args.tab_ders + (q * DIM + d) * NN + a);
                staticFor< U >([&](auto uu) {
                    staticFor< E >([&](auto eq) {
                        if constexpr (unknownInEquation< KernelT >(uu, eq))
                        {
                            constexpr int ent = Cfg::entIndex(uu, eq);
                            const double2 c01 = *reinterpret_cast< const double2* >(s_c + (qc * n_ent + ent) * 4);
                            const double2 c23 = *reinterpret_cast< const double2* >(s_c + (qc * n_ent + ent) * 4 + 2);
                            double        b   = c01.x * bas[0];
                            b                 = fma(c01.y, bas[1], b);
                            if constexpr (DIM >= 2)
                                b = fma(c23.x, bas[2], b);
                            if constexpr (DIM >= 3)
                                b = fma(c23.y, bas[3], b);
                            dacc[i][uu] = fma(b, b, dacc[i][uu]);
                        }
                    });
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                    {
                        const double* c4  = s_rc + ((qc * U + uu) * NRHS + r) * 4;
                        double        acc = c4[0] * bas[0];
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            acc = fma(c4[1 + d], bas[1 + d], acc);
                        racc[i][uu][r] += acc;
                    }
                });
            }
        }
        __syncthreads();
    }
    // ---- scatter (MatrixFreeSystem.hpp:377-390)
#pragma unroll
    for (int i = 0; i < NPT; ++i)
    {
        const int a = tid + i * T;
        if (a >= NN)
            continue;
        const long long node = el_nodes[a];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long dof = node * args.dofs_per_node + args.dof_inds[u];
            atomicAdd(args.diag + dof, dacc[i][u]);
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                atomicAdd(args.rhs + dof + r * args.ld, racc[i][u][r]);
        }
    }
}
} // namespace l3b
#endif
