// Sum-factorised matrix-free apply for hexahedra, "planes + columns" formulation (kernel v5):
//     y[dofs(e)] += alpha * K_e * x[dofs(e)],
// same mathematics and parity contract as mfSumFactApplyKernel (mf_sumfact.cuh), which stays the path for quads and
// for quadratures too large for this kernel's register tiles. Replaces evalLocalOperatorSumFact +
// gatherSumFact/scatterSumFact of the reference (algsys/SumFactorization.hpp:438-917,
// algsys/MatrixFreeSystem.hpp:421-537).
//
// Why another formulation: the line-per-thread kernel moves every tensor through shared memory once per 1-D sweep
// (~136 KB and 12 barriers per p=4 element; ncu r1_v4: shared-memory pipe 46 %, fp64 pipe 25 %, 7 150 warp instructions
// per element). Here every thread owns a whole z-column or a whole xy-plane of one field in REGISTERS, so a sweep costs
// no shared-memory traffic at all; shared memory is only the transpose buffer between the two ownerships:
//
//   A  column (i, j)     gather the nodal z-column of every unknown from x (HBM, vector loads), interpolate along z
//                        in registers, store                                                   →  V[f][qz][j][i]
//   B  plane  (f, qz)    load the nb x nb plane, interpolate along x and y in registers, then differentiate the
//                        interpolated plane along x and y (collocation derivative at the Gauss points)
//                                                                                              →  V, DX, DY
//   C  column (qx, qy)   load values and x/y-derivatives of every field along z; z-derivative in registers;
//                        quadrature-point stage (geometry, user kernel, least-squares operator, fluxes r0, r_xi, r_eta,
//                        r_zeta) at the nq points of the column; transposed z-derivative accumulated in registers
//                                                                                              →  V (= r0 + Dz^T r_zeta), DX, DY
//   D  plane  (u, qz)    load the three planes, transposed x/y-derivatives, project to the nodes along y and x →  V
//   E  column (i, j)     load along z, project to the nodes along z, scatter with fp64 atomics (RED) into y (HBM)
//
// 4 barriers and ~64 KB of shared-memory traffic per p=4 element; 46 k DFMA per element (15 k per transform direction
// pair, 16 k in the point stage) is what remains, i.e. the kernel is built to be bound by the fp64 pipe.
//
// Shared-memory layout (doubles): three arrays V, DX, DY of [plane = f * NQ + qz][element slot][PSZ], PSZ = NQ^2 rounded
// up to an odd number. A column thread (slot, c) addresses plane * EPB * PSZ + slot * PSZ + c: consecutive threads →
// consecutive words, conflict-free. A plane thread w = plane * EPB + slot addresses PSZ * w + n: odd stride →
// conflict-free.
#ifndef L3B_MF_HEX_PLANES_CUH
#define L3B_MF_HEX_PLANES_CUH

#include "mf_sumfact.cuh"

namespace l3b
{
template < typename KernelT, int P, int NQ, int NRHS >
struct MfHexCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields;
    static constexpr int  NB = P + 1;
    static constexpr int  F0 = U * NRHS, F = F0 + NF;
    static constexpr int  NN  = NB * NB * NB;
    static constexpr int  CT  = NQ * NQ;                    // column threads per element
    static constexpr int  PSZ = CT % 2 == 0 ? CT + 1 : CT;  // plane stride
    static constexpr int  EPB = cmax(1, 128 / CT);          // elements per CTA
    static constexpr int  threads = ((EPB * CT + 31) / 32) * 32;
    static constexpr int  NPL     = F * NQ;                 // planes per element, back transform
    static constexpr int  NPLF    = F0 * NQ;                // planes per element, transposed transform
    static constexpr int  AS      = NPL * EPB * PSZ;        // doubles per array
    static constexpr int  geo_doubles = 36;                 // 8 x 3 monomial coefficients, Jti[3][3], detJ, affine flag (+1 pad)
    static constexpr size_t smem_bytes = (3 * static_cast< size_t >(AS) + EPB * geo_doubles) * sizeof(double);
    // registers: the point stage keeps a z-column of values and of accumulators for U unknowns + the NF field values
    static constexpr int col_doubles = (2 * U + NF) * NQ + NQ * NQ / 2;
    static constexpr int min_blocks  = col_doubles <= 56 and smem_bytes <= 72 * 1024 ? 3 : smem_bytes <= 110 * 1024 ? 2 : 1;
    static constexpr bool supported  = CT <= 49 and smem_bytes <= 200 * 1024;
    static_assert(params.dimension == 3);
    static_assert(NQ >= NB, "collocation differentiation at the Gauss points needs nq >= nb (value_order >= 1)");
};

// does unknown u's physical derivative along s enter any equation? (compile-time, from the probed operator sparsity)
template < typename KernelT >
constexpr bool gradNeeded(int s, int u)
{
    using Sp = KernelSparsity< KernelT >;
    for (size_t eq = 0; eq < KernelT::parameters.n_equations; ++eq)
        if (Sp::nz(s + 1, eq, u))
            return true;
    return false;
}

template < typename KernelT, int P, int NQ, int NRHS >
__global__ void __launch_bounds__(MfHexCfg< KernelT, P, NQ, NRHS >::threads, MfHexCfg< KernelT, P, NQ, NRHS >::min_blocks)
    mfHexPlanesKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ SumFactTables< P + 1, NQ > tab)
{
    using Cfg = MfHexCfg< KernelT, P, NQ, NRHS >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NB = Cfg::NB, F0 = Cfg::F0, F = Cfg::F, NN = Cfg::NN;
    constexpr int CT = Cfg::CT, PSZ = Cfg::PSZ, EPB = Cfg::EPB, AS = Cfg::AS, NPL = Cfg::NPL, NPLF = Cfg::NPLF;
    constexpr int T = Cfg::threads;
    extern __shared__ double smem[];
    double* const s_V   = smem;
    double* const s_DX  = smem + AS;
    double* const s_DY  = smem + 2 * AS;
    double* const s_geo = smem + 3 * AS;

    const int       tid      = threadIdx.x;
    const long long wi0      = static_cast< long long >(blockIdx.x) * EPB;
    const int       n_active = static_cast< int >(args.n_work - wi0 < EPB ? args.n_work - wi0 : EPB); // elements of this CTA
    // column role
    const int  slot    = tid / CT;
    const int  cc      = tid % CT;
    const bool col_on  = tid < EPB * CT and slot < n_active;
    const int  ci      = cc % NQ, cj = cc / NQ;
    const long long ce = col_on ? (args.work_elems ? args.work_elems[wi0 + slot] : args.first_elem + wi0 + slot) : 0;
    const uint32_t* el_nodes = args.nodes + ce * NN;
    double* const   geo      = s_geo + slot * Cfg::geo_doubles;
    const int       col_off  = slot * PSZ + cc; // + plane * EPB * PSZ

    // ---- A: gather (MatrixFreeSystem.hpp:421-467) + z-interpolation
    if (col_on)
    {
        buildGeometryCoefs< 3 >(args.verts + ce * 24, geo, cc, CT);
        if (ci < NB and cj < NB)
        {
            long long node[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k)
                node[k] = el_nodes[k * NB * NB + cj * NB + ci];
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                double v[U][NB];
                if (args.contiguous_dofs and U % 2 == 0)
                {
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        const long long dof = node[k] * U;
                        const double2*  src = reinterpret_cast< const double2* >(args.x + dof + r * args.ld);
#pragma unroll
                        for (int u2 = 0; u2 < U / 2; ++u2)
                        {
                            const double2 val = __ldg(src + u2);
                            v[2 * u2][k]      = isDirichlet(args.dir_mask, dof + 2 * u2) ? 0. : val.x;
                            v[2 * u2 + 1][k]  = isDirichlet(args.dir_mask, dof + 2 * u2 + 1) ? 0. : val.y;
                        }
                    }
                }
                else
                {
#pragma unroll
                    for (int k = 0; k < NB; ++k)
#pragma unroll
                        for (int u = 0; u < U; ++u)
                        {
                            const long long dof = node[k] * args.dofs_per_node + args.dof_inds[u];
                            v[u][k]             = isDirichlet(args.dir_mask, dof) ? 0. : __ldg(args.x + dof + r * args.ld);
                        }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        double acc = v[u][0] * tab.interp[q];
#pragma unroll
                        for (int k = 1; k < NB; ++k)
                            acc = fma(v[u][k], tab.interp[k * NQ + q], acc);
                        s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off] = acc;
                    }
            }
            if constexpr (NF > 0)
            {
#pragma unroll
                for (int f = 0; f < NF; ++f)
                {
                    double v[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        v[k] = __ldg(args.fields + node[k] + args.field_inds[f] * args.field_stride);
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        double acc = v[0] * tab.interp[q];
#pragma unroll
                        for (int k = 1; k < NB; ++k)
                            acc = fma(v[k], tab.interp[k * NQ + q], acc);
                        s_V[((F0 + f) * NQ + q) * (EPB * PSZ) + col_off] = acc;
                    }
                }
            }
        }
    }
    __syncthreads();

    // affine element (all mixed monomial coefficients vanish): one thread inverts the constant Jacobian for everybody
    if (col_on and cc == 0)
    {
        bool affine = true;
#pragma unroll
        for (int m = 0; m < 8; ++m)
            if (__popc(m) > 1)
                for (int s = 0; s < 3; ++s)
                    affine = affine and geo[m * 3 + s] == 0.;
        geo[34] = affine ? 1. : 0.;
        if (affine)
        {
            double Jt[3][3], Jti[3][3];
            for (int d = 0; d < 3; ++d)
                for (int s = 0; s < 3; ++s)
                    Jt[d][s] = geo[(1 << d) * 3 + s];
            geo[33] = invert< 3 >(Jt, Jti);
            for (int s = 0; s < 3; ++s)
                for (int d = 0; d < 3; ++d)
                    geo[24 + s * 3 + d] = Jti[s][d];
        }
    }

    // ---- B: xy-planes: interpolate along x and y, differentiate along x and y
    for (int w = tid; w < NPL * EPB; w += T)
    {
        if (w % EPB >= n_active)
            continue;
        double* const pv = s_V + w * PSZ;
        double        a[NQ][NQ];
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int i = 0; i < NB; ++i)
                a[j][i] = pv[j * NQ + i];
#pragma unroll
        for (int j = 0; j < NB; ++j) // x
        {
            double o[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
            {
                double acc = a[j][0] * tab.interp[q];
#pragma unroll
                for (int i = 1; i < NB; ++i)
                    acc = fma(a[j][i], tab.interp[i * NQ + q], acc);
                o[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                a[j][q] = o[q];
        }
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) // y
        {
            double o[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
            {
                double acc = a[0][qx] * tab.interp[q];
#pragma unroll
                for (int j = 1; j < NB; ++j)
                    acc = fma(a[j][qx], tab.interp[j * NQ + q], acc);
                o[q] = acc;
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                a[q][qx] = o[q];
        }
        double* const px = s_DX + w * PSZ;
        double* const py = s_DY + w * PSZ;
#pragma unroll
        for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
            for (int qx = 0; qx < NQ; ++qx)
            {
                pv[qy * NQ + qx] = a[qy][qx];
                double dx = a[qy][0] * tab.colloc[qx], dy = a[0][qx] * tab.colloc[qy];
#pragma unroll
                for (int m = 1; m < NQ; ++m)
                {
                    dx = fma(a[qy][m], tab.colloc[m * NQ + qx], dx);
                    dy = fma(a[m][qx], tab.colloc[m * NQ + qy], dy);
                }
                px[qy * NQ + qx] = dx;
                py[qy * NQ + qx] = dy;
            }
    }
    __syncthreads();

    // ---- C: quadrature-point stage along the z-column (SumFactorization.hpp:614-756)
    if (col_on)
    {
        const bool   affine = geo[34] != 0.;
        const double wxy    = tab.w[ci] * tab.w[cj];
        double       Jti[3][3], detJ = 0.;
        if (affine)
        {
            detJ = geo[33];
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    Jti[s][d] = geo[24 + s * 3 + d];
        }
        double fval[NF > 0 ? NF : 1][NQ];
        if constexpr (NF > 0)
        {
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    fval[f][q] = s_V[((F0 + f) * NQ + q) * (EPB * PSZ) + col_off];
        }
        bool violated = false;
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
        {
            double val[U][NQ], wacc[U][NQ];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                {
                    val[u][q]  = s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off];
                    wacc[u][q] = 0.;
                }
#pragma unroll
            for (int qz = 0; qz < NQ; ++qz)
            {
                // geometry + user kernel at (ci, cj, qz)
                typename KernelT::Input in;
                {
                    const double xi[3] = {tab.pts[ci], tab.pts[cj], tab.pts[qz]};
                    double       xs[3], Jt[3][3];
                    geometryFromCoefs< 3 >(geo, xi, xs, Jt); // dead code for affine elements whose kernel ignores the point
                    if (not affine)
                        detJ = invert< 3 >(Jt, Jti);
                    in.point.space.coords[0] = xs[0];
                    in.point.space.coords[1] = xs[1];
                    in.point.space.coords[2] = 0.; // SumFactorization.hpp:656, :732
                }
                in.point.time = args.time;
#pragma unroll
                for (int f = 0; f < NF; ++f)
                {
                    const int    off = ((F0 + f) * NQ + qz) * (EPB * PSZ) + col_off;
                    const double dxf = s_DX[off], dyf = s_DY[off];
                    double       dzf = fval[f][0] * tab.colloc[qz];
#pragma unroll
                    for (int m = 1; m < NQ; ++m)
                        dzf = fma(fval[f][m], tab.colloc[m * NQ + qz], dzf);
                    in.field_vals[f] = fval[f][qz];
#pragma unroll
                    for (int s = 0; s < 3; ++s)
                        in.field_ders[s][f] = fma(Jti[s][2], dzf, fma(Jti[s][1], dyf, Jti[s][0] * dxf));
                }
                const auto   res = kernel(in);
                const double wgt = wxy * tab.w[qz] * detJ;
                staticFor< 4 >([&](auto op) {
                    staticFor< E >([&](auto eq) {
                        staticFor< U >([&](auto u) {
                            if constexpr (not Sp::nz(op, eq, u))
                                violated |= res.operators[op](eq, u) != 0.;
                        });
                    });
                });
                // reference derivatives of the operand at this point
                double dref[3][U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const int off = ((r * U + u) * NQ + qz) * (EPB * PSZ) + col_off;
                    dref[0][u]    = s_DX[off];
                    dref[1][u]    = s_DY[off];
                    double dz     = val[u][0] * tab.colloc[qz];
#pragma unroll
                    for (int m = 1; m < NQ; ++m)
                        dz = fma(val[u][m], tab.colloc[m * NQ + qz], dz);
                    dref[2][u] = dz;
                }
                double g_phys[3][U];
                staticFor< U >([&](auto u) {
                    staticFor< 3 >([&](auto s) {
                        if constexpr (gradNeeded< KernelT >(s, u))
                            g_phys[s][u] = fma(Jti[s][2], dref[2][u], fma(Jti[s][1], dref[1][u], Jti[s][0] * dref[0][u]));
                    });
                });
                double tv[E];
                staticFor< E >([&](auto eq) {
                    double acc = 0.;
                    staticFor< U >([&](auto u) {
                        if constexpr (Sp::nz(0, eq, u))
                            acc = fma(res.operators[0](eq, u), val[u][qz], acc);
                        staticFor< 3 >([&](auto s) {
                            if constexpr (Sp::nz(s + 1, eq, u))
                                acc = fma(res.operators[s + 1](eq, u), g_phys[s][u], acc);
                        });
                    });
                    tv[eq] = acc * wgt;
                });
                staticFor< U >([&](auto u) {
                    double a0 = 0., ps[3] = {0., 0., 0.};
                    staticFor< E >([&](auto eq) {
                        if constexpr (Sp::nz(0, eq, u))
                            a0 = fma(res.operators[0](eq, u), tv[eq], a0);
                        staticFor< 3 >([&](auto s) {
                            if constexpr (Sp::nz(s + 1, eq, u))
                                ps[s] = fma(res.operators[s + 1](eq, u), tv[eq], ps[s]);
                        });
                    });
                    // r_d = sum_s Ji(d, s) p_s, only over the directions s this unknown is differentiated along
                    double rd[3] = {0., 0., 0.};
                    staticFor< 3 >([&](auto s) {
                        if constexpr (gradNeeded< KernelT >(s, u))
                        {
#pragma unroll
                            for (int d = 0; d < 3; ++d)
                                rd[d] = fma(Jti[s][d], ps[s], rd[d]);
                        }
                    });
                    const int off = ((r * U + u) * NQ + qz) * (EPB * PSZ) + col_off;
                    s_DX[off]     = rd[0];
                    s_DY[off]     = rd[1];
                    wacc[u][qz] += a0;
#pragma unroll
                    for (int m = 0; m < NQ; ++m)
                        wacc[u][m] = fma(tab.colloc[m * NQ + qz], rd[2], wacc[u][m]);
                });
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off] = wacc[u][q];
        }
        if (violated)
            atomicOr(args.status, status_sparsity_violation);
    }
    __syncthreads();

    // ---- D: xy-planes: transposed derivatives, projection to the nodes along y and x
    for (int w = tid; w < NPLF * EPB; w += T)
    {
        if (w % EPB >= n_active)
            continue;
        double* const       pv = s_V + w * PSZ;
        const double* const px = s_DX + w * PSZ;
        const double* const py = s_DY + w * PSZ;
        double              t[NQ][NQ];
#pragma unroll
        for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
            for (int qx = 0; qx < NQ; ++qx)
                t[qy][qx] = pv[qy * NQ + qx];
#pragma unroll
        for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
            for (int qx = 0; qx < NQ; ++qx)
            {
                const double rx = px[qy * NQ + qx], ry = py[qy * NQ + qx];
#pragma unroll
                for (int m = 0; m < NQ; ++m)
                {
                    t[qy][m] = fma(tab.colloc[m * NQ + qx], rx, t[qy][m]);
                    t[m][qx] = fma(tab.colloc[m * NQ + qy], ry, t[m][qx]);
                }
            }
#pragma unroll
        for (int qx = 0; qx < NQ; ++qx) // y: Gauss points → nodes
        {
            double o[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j)
            {
                double acc = t[0][qx] * tab.interp[j * NQ];
#pragma unroll
                for (int q = 1; q < NQ; ++q)
                    acc = fma(t[q][qx], tab.interp[j * NQ + q], acc);
                o[j] = acc;
            }
#pragma unroll
            for (int j = 0; j < NB; ++j)
                t[j][qx] = o[j];
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) // x
#pragma unroll
            for (int i = 0; i < NB; ++i)
            {
                double acc = t[j][0] * tab.interp[i * NQ];
#pragma unroll
                for (int q = 1; q < NQ; ++q)
                    acc = fma(t[j][q], tab.interp[i * NQ + q], acc);
                pv[j * NQ + i] = acc;
            }
    }
    __syncthreads();

    // ---- E: z-projection + scatter (MatrixFreeSystem.hpp:494-537): relaxed fp64 atomics, Dirichlet rows skipped
    if (col_on and ci < NB and cj < NB)
    {
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
        {
            double out[U][NB];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
                double v[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    v[q] = s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off];
#pragma unroll
                for (int k = 0; k < NB; ++k)
                {
                    double acc = v[0] * tab.interp[k * NQ];
#pragma unroll
                    for (int q = 1; q < NQ; ++q)
                        acc = fma(v[q], tab.interp[k * NQ + q], acc);
                    out[u][k] = acc * args.alpha;
                }
            }
#pragma unroll
            for (int k = 0; k < NB; ++k)
            {
                const long long node = el_nodes[k * NB * NB + cj * NB + ci];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const long long dof = node * args.dofs_per_node + args.dof_inds[u];
                    if (not isDirichlet(args.dir_mask, dof))
                        atomicAdd(args.y + dof + r * args.ld, out[u][k]);
                }
            }
        }
    }
}
} // namespace l3b
#endif
