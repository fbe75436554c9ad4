// Sum-factorised matrix-free local operator apply, fused with gather and scatter:
//     y[dofs(e)] += alpha * K_e * x[dofs(e)]      for every element e of the work list,
// K_e = sum_q [N_q (x) A0^T + sum_d d_xi_d N_q (x) D_d^T] w|J| [A0 u_q + sum_d D_d d_xi_d u_q]  (SURVEY Appendix A).
//
// Replaces evalLocalOperatorSumFact + gatherSumFact/scatterSumFact of the reference
// (algsys/SumFactorization.hpp:438-917, algsys/MatrixFreeSystem.hpp:421-537).
//
// Formulation (differs from the reference in rounding only, parity 1e-12):
//   * u is interpolated from the GLL nodes to the nq^D Gauss points with D sweeps; reference-space derivatives are then
//     taken *at the Gauss points* with the collocation derivative matrix of the Gauss-point Lagrange basis (exact for
//     nq >= nb, which AssemblyOptions guarantees for value_order >= 1) — 2D sweeps instead of the reference's 5 (quad) /
//     9 (hex);
//   * per quadrature point the physical gradient g_s = sum_d Ji(d,s) d_xi_d u is formed first, then
//     t = w|J| (A0 u + sum_s A_s g_s), r0 = A0^T t, p_s = A_s^T t, r_d = sum_s Ji(d,s) p_s — this never forms the
//     reference's D_d = sum_s A_s Ji(d,s) matrices (SumFactorization.hpp:736-738), saving (D*D*E*U - 2*D*D*U) FMAs per point;
//   * the hex path hands the user kernel point.space = (x, y, 0), replicating SumFactorization.hpp:732.
//
// Thread mapping: TPE = max(nb, nq)^D threads per element, EPB elements per CTA. All tensors of an element live in
// shared memory; every sweep assigns one 1-D line to a thread (coalesced, conflict-light strides).
#ifndef L3B_MF_SUMFACT_CUH
#define L3B_MF_SUMFACT_CUH

#include "device_common.cuh"

namespace l3b
{
template < int NB, int NQ >
struct SumFactTables
{
    double interp[NB * NQ]; // [b][q]
    double der[NB * NQ];    // [b][q]
    double colloc[NQ * NQ]; // [m][q]
    double w[NQ];
    double pts[NQ];
};

template < typename KernelT, int DIM, int P, int NQ, int NRHS >
struct MfSumFactCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields;
    static constexpr int  NB = P + 1;
    static constexpr int  F0 = U * NRHS, F = F0 + NF;
    static constexpr int  NN = cpow(NB, DIM), Q = cpow(NQ, DIM), M = cmax(NN, Q);
    static constexpr int  TPE = M;
    static constexpr int  EPB = cmax(1, 128 / TPE);
    static constexpr int  threads = TPE * EPB;
    // per element: two F x M ping-pong buffers and (DIM + 1) x F0 x Q result buffers
    static constexpr int    smem_doubles_per_elem = 2 * F * M + (DIM + 1) * F0 * Q;
    static constexpr size_t smem_bytes            = static_cast< size_t >(EPB) * smem_doubles_per_elem * sizeof(double);
    static_assert(params.dimension == DIM);
    static_assert(NQ >= NB, "collocation differentiation at the Gauss points needs nq >= nb (value_order >= 1)");
    static_assert(threads <= 1024, "element too large for one thread per tensor entry");
};

// out[line][o] = sum_i in[line][i] * Mtx[i][o]; line layout given by strides. One thread per line.
template < int N_IN, int N_OUT, bool TRANSPOSED_TABLE >
__device__ __forceinline__ void sweepLine(const double* __restrict__ in, int in_stride, double* __restrict__ out, int out_stride,
                                          const double* __restrict__ tab /* [N_IN][N_OUT] or transposed [N_OUT][N_IN] */)
{
    double v[N_IN];
#pragma unroll
    for (int i = 0; i < N_IN; ++i)
        v[i] = in[i * in_stride];
#pragma unroll
    for (int o = 0; o < N_OUT; ++o)
    {
        double acc = 0.;
#pragma unroll
        for (int i = 0; i < N_IN; ++i)
            acc = fma(v[i], TRANSPOSED_TABLE ? tab[o * N_IN + i] : tab[i * N_OUT + o], acc);
        out[o * out_stride] = acc;
    }
}

template < typename KernelT, int DIM, int P, int NQ, int NRHS >
__global__ void __launch_bounds__(MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::threads)
    mfSumFactApplyKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ SumFactTables< P + 1, NQ > tab)
{
    using Cfg = MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NB = Cfg::NB, F0 = Cfg::F0, F = Cfg::F, NN = Cfg::NN, Q = Cfg::Q, M = Cfg::M;
    extern __shared__ double smem[];
    const int       slot  = threadIdx.x / Cfg::TPE;
    const int       t     = threadIdx.x % Cfg::TPE;
    const long long wi    = static_cast< long long >(blockIdx.x) * Cfg::EPB + slot;
    const bool      active = wi < args.n_work;
    const long long e     = active ? (args.work_elems ? args.work_elems[wi] : args.first_elem + wi) : 0;
    double*         bufA  = smem + static_cast< size_t >(slot) * Cfg::smem_doubles_per_elem;
    double*         bufB  = bufA + F * M;
    double*         bufR  = bufB + F * M; // [(DIM+1)][F0][Q]
    const uint32_t* el_nodes = args.nodes + e * NN;

    // ---- gather (MatrixFreeSystem.hpp:421-467): bufA[f][a], f = rhs*U + u, then the NF external fields
    if (active)
        for (int a = t; a < NN; a += Cfg::TPE)
        {
            const long long node = el_nodes[a];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
                const long long dof = node * args.dofs_per_node + args.dof_inds[u];
                const bool      dir = isDirichlet(args.dir_mask, dof);
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                    bufA[(r * U + u) * M + a] = dir ? 0. : args.x[dof + r * args.ld];
            }
#pragma unroll
            for (int f = 0; f < NF; ++f)
                bufA[(F0 + f) * M + a] = args.fields[node + args.field_inds[f] * args.field_stride];
        }
    __syncthreads();

    // ---- interpolate to the Gauss points, one direction at a time (bufA → bufB → bufA [→ bufB])
    if constexpr (DIM == 2)
    {
        // x: lines (f, j): in bufA[f][j*NB + i] → bufB[f][j*NQ + qx]
        for (int l = t; l < F * NB; l += Cfg::TPE)
        {
            const int f = l / NB, j = l % NB;
            sweepLine< NB, NQ, false >(bufA + f * M + j * NB, 1, bufB + f * M + j * NQ, 1, tab.interp);
        }
        __syncthreads();
        // y: lines (f, qx): in bufB[f][j*NQ + qx] → bufA[f][qy*NQ + qx]
        for (int l = t; l < F * NQ; l += Cfg::TPE)
        {
            const int f = l / NQ, qx = l % NQ;
            sweepLine< NB, NQ, false >(bufB + f * M + qx, NQ, bufA + f * M + qx, NQ, tab.interp);
        }
        __syncthreads();
    }
    else
    {
        for (int l = t; l < F * NB * NB; l += Cfg::TPE)
        {
            const int f = l / (NB * NB), kj = l % (NB * NB);
            sweepLine< NB, NQ, false >(bufA + f * M + kj * NB, 1, bufB + f * M + kj * NQ, 1, tab.interp);
        }
        __syncthreads();
        // y: lines (f, k, qx): in bufB[f][(k*NB + j)*NQ + qx] → bufA[f][(k*NQ + qy)*NQ + qx]
        for (int l = t; l < F * NB * NQ; l += Cfg::TPE)
        {
            const int f = l / (NB * NQ), k = (l / NQ) % NB, qx = l % NQ;
            sweepLine< NB, NQ, false >(bufB + f * M + k * NB * NQ + qx, NQ, bufA + f * M + k * NQ * NQ + qx, NQ, tab.interp);
        }
        __syncthreads();
        // z: lines (f, qy, qx): in bufA[f][(k*NQ + qy)*NQ + qx] → bufB[f][(qz*NQ + qy)*NQ + qx]
        for (int l = t; l < F * NQ * NQ; l += Cfg::TPE)
        {
            const int f = l / (NQ * NQ), qyx = l % (NQ * NQ);
            sweepLine< NB, NQ, false >(bufA + f * M + qyx, NQ * NQ, bufB + f * M + qyx, NQ * NQ, tab.interp);
        }
        __syncthreads();
    }
    double* uq = DIM == 2 ? bufA : bufB; // values at the Gauss points, [f][q], q = (qz*NQ + qy)*NQ + qx
    double* vq = DIM == 2 ? bufB : bufA; // free buffer, receives the pre-projection result

    // ---- quadrature point stage (SumFactorization.hpp:614-756)
    if (active)
        for (int q = t; q < Q; q += Cfg::TPE)
        {
            int qi[3] = {q % NQ, (q / NQ) % NQ, DIM == 3 ? q / (NQ * NQ) : 0};
            // values and reference-space derivatives of all F fields
            double val[F], dref[DIM][F];
#pragma unroll
            for (int f = 0; f < F; ++f)
            {
                const double* line = uq + f * M;
                val[f]             = line[q];
                int stride         = 1;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    const int base = q - qi[d] * stride;
                    double    acc  = 0.;
#pragma unroll
                    for (int m = 0; m < NQ; ++m)
                        acc = fma(tab.colloc[m * NQ + qi[d]], line[base + m * stride], acc);
                    dref[d][f] = acc;
                    stride *= NQ;
                }
            }
            // geometry (computeGeomDataLin, :506-537)
            double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM];
            double wq = 1.;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                xi[d] = tab.pts[qi[d]];
                wq *= tab.w[qi[d]];
            }
            geometryAt< DIM >(args.verts + e * (1 << DIM) * 3, xi, xs, Jt);
            const double detJ = invert< DIM >(Jt, Jti);
            // Jt[d][s] = dx_s/dxi_d; its inverse satisfies Jti[s][d] = dxi_d/dx_s, i.e. the reference's jac_inv(d, s) = Jti[s][d]
            typename KernelT::Input in;
#pragma unroll
            for (int f = 0; f < NF; ++f)
            {
                in.field_vals[f] = val[F0 + f];
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                {
                    double acc = 0.;
#pragma unroll
                    for (int d = 0; d < DIM; ++d)
                        acc = fma(Jti[s][d], dref[d][F0 + f], acc);
                    in.field_ders[s][f] = acc;
                }
            }
            in.point.space.coords[0] = xs[0];
            in.point.space.coords[1] = xs[1];
            in.point.space.coords[2] = 0.; // SumFactorization.hpp:656, :732
            in.point.time            = args.time;
            const auto   res = kernel(in);
            const double wgt = wq * detJ;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                // physical gradients of the operand
                double g[DIM][U];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int s = 0; s < DIM; ++s)
                    {
                        double acc = 0.;
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            acc = fma(Jti[s][d], dref[d][r * U + u], acc);
                        g[s][u] = acc;
                    }
                double tv[E];
#pragma unroll
                for (int eq = 0; eq < E; ++eq)
                {
                    double acc = 0.;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        acc = fma(res.operators[0](eq, u), val[r * U + u], acc);
#pragma unroll
                        for (int s = 0; s < DIM; ++s)
                            acc = fma(res.operators[s + 1](eq, u), g[s][u], acc);
                    }
                    tv[eq] = acc * wgt;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    double r0 = 0., ps[DIM];
#pragma unroll
                    for (int s = 0; s < DIM; ++s)
                        ps[s] = 0.;
#pragma unroll
                    for (int eq = 0; eq < E; ++eq)
                    {
                        r0 = fma(res.operators[0](eq, u), tv[eq], r0);
#pragma unroll
                        for (int s = 0; s < DIM; ++s)
                            ps[s] = fma(res.operators[s + 1](eq, u), tv[eq], ps[s]);
                    }
                    bufR[(0 * F0 + r * U + u) * Q + q] = r0;
#pragma unroll
                    for (int d = 0; d < DIM; ++d)
                    {
                        double acc = 0.;
#pragma unroll
                        for (int s = 0; s < DIM; ++s)
                            acc = fma(Jti[s][d], ps[s], acc);
                        bufR[((d + 1) * F0 + r * U + u) * Q + q] = acc;
                    }
                }
            }
        }
    __syncthreads();

    // ---- transposed collocation derivative: v(q) = r0(q) + sum_d sum_n colloc[q_d][n] r_d(.., n, ..)
    if (active)
        for (int q = t; q < Q; q += Cfg::TPE)
        {
            int qi[3] = {q % NQ, (q / NQ) % NQ, DIM == 3 ? q / (NQ * NQ) : 0};
#pragma unroll
            for (int f = 0; f < F0; ++f)
            {
                double acc    = bufR[f * Q + q];
                int    stride = 1;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    const double* line = bufR + ((d + 1) * F0 + f) * Q + (q - qi[d] * stride);
#pragma unroll
                    for (int n = 0; n < NQ; ++n)
                        acc = fma(tab.colloc[qi[d] * NQ + n], line[n * stride], acc);
                    stride *= NQ;
                }
                vq[f * M + q] = acc;
            }
        }
    __syncthreads();

    // ---- project back to the nodes (transposed interpolation sweeps), last direction first
    double* res_nodes;
    if constexpr (DIM == 2)
    {
        // y: lines (f, qx): vq[f][qy*NQ + qx] → uq[f][j*NQ + qx]
        for (int l = t; l < F0 * NQ; l += Cfg::TPE)
        {
            const int f = l / NQ, qx = l % NQ;
            sweepLine< NQ, NB, true >(vq + f * M + qx, NQ, uq + f * M + qx, NQ, tab.interp);
        }
        __syncthreads();
        // x: lines (f, j): uq[f][j*NQ + qx] → vq[f][j*NB + i]
        for (int l = t; l < F0 * NB; l += Cfg::TPE)
        {
            const int f = l / NB, j = l % NB;
            sweepLine< NQ, NB, true >(uq + f * M + j * NQ, 1, vq + f * M + j * NB, 1, tab.interp);
        }
        __syncthreads();
        res_nodes = vq;
    }
    else
    {
        // z: lines (f, qy, qx): vq[f][(qz*NQ + qy)*NQ + qx] → uq[f][(k*NQ + qy)*NQ + qx]
        for (int l = t; l < F0 * NQ * NQ; l += Cfg::TPE)
        {
            const int f = l / (NQ * NQ), qyx = l % (NQ * NQ);
            sweepLine< NQ, NB, true >(vq + f * M + qyx, NQ * NQ, uq + f * M + qyx, NQ * NQ, tab.interp);
        }
        __syncthreads();
        // y: lines (f, k, qx): uq[f][(k*NQ + qy)*NQ + qx] → vq[f][(k*NB + j)*NQ + qx]
        for (int l = t; l < F0 * NB * NQ; l += Cfg::TPE)
        {
            const int f = l / (NB * NQ), k = (l / NQ) % NB, qx = l % NQ;
            sweepLine< NQ, NB, true >(uq + f * M + k * NQ * NQ + qx, NQ, vq + f * M + k * NB * NQ + qx, NQ, tab.interp);
        }
        __syncthreads();
        // x: lines (f, k, j): vq[f][(k*NB + j)*NQ + qx] → uq[f][(k*NB + j)*NB + i]
        for (int l = t; l < F0 * NB * NB; l += Cfg::TPE)
        {
            const int f = l / (NB * NB), kj = l % (NB * NB);
            sweepLine< NQ, NB, true >(vq + f * M + kj * NQ, 1, uq + f * M + kj * NB, 1, tab.interp);
        }
        __syncthreads();
        res_nodes = uq;
    }

    // ---- scatter (MatrixFreeSystem.hpp:494-537): relaxed fp64 atomics, Dirichlet rows skipped
    if (active)
        for (int a = t; a < NN; a += Cfg::TPE)
        {
            const long long node = el_nodes[a];
#pragma unroll
            for (int u = 0; u < U; ++u)
            {
                const long long dof = node * args.dofs_per_node + args.dof_inds[u];
                if (isDirichlet(args.dir_mask, dof))
                    continue;
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                    atomicAdd(args.y + dof + r * args.ld, args.alpha * res_nodes[(r * U + u) * M + a]);
            }
        }
}
} // namespace l3b
#endif
