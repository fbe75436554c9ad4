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
//     reference's D_d = sum_s A_s Ji(d,s) matrices (SumFactorization.hpp:736-738);
//   * multiply-adds against structurally zero operator entries are not issued (KernelSparsity, kernel_interface.cuh):
//     15 of 112 entries survive for 3-D diffusion. Every evaluation still checks the masked entries are zero;
//   * the multilinear geometry is evaluated from its monomial coefficients (8 x 3 numbers per hex, built once per element
//     in shared memory) instead of summing over the vertices at every point (computeGeomDataLin, :506-537);
//   * the hex path hands the user kernel point.space = (x, y, 0), replicating SumFactorization.hpp:732.
//
// Thread mapping: TPE = nq^D threads per element (one per Gauss point / node), EPB elements per CTA. All tensors of an
// element live in shared memory; every sweep assigns one 1-D line to a thread. Shared memory per element: two F x Q
// ping-pong buffers (+ extra when D*F0 > 2F), reused for the D x F0 x Q flux arrays r_d of the transposed stage.
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
    static constexpr int  NN = cpow(NB, DIM), Q = cpow(NQ, DIM), M = Q;
    static constexpr int  TPE = Q;
    static constexpr int  EPB = cmax(1, 128 / TPE);
    static constexpr int  threads = TPE * EPB;
    // resident CTAs per SM the register allocator should leave room for (~20 warps per SM)
    static constexpr int warps      = (threads + 31) / 32;
    static constexpr int min_blocks = cmax(1, 20 / warps);
    // the D*F0 flux arrays are laid over the two ping-pong buffers; anything beyond 2F arrays needs extra room
    static constexpr int extra_q     = cmax(0, DIM * F0 - 2 * F);
    static constexpr int geo_doubles = (1 << DIM) * 3;
    static constexpr int tab_doubles = NQ * NQ + 2 * NQ; // colloc, pts, w
    static constexpr int smem_doubles_per_elem = (2 * F + extra_q) * Q + geo_doubles;
    static constexpr size_t smem_bytes = (static_cast< size_t >(EPB) * smem_doubles_per_elem + tab_doubles) * sizeof(double);
    static_assert(params.dimension == DIM);
    static_assert(NQ >= NB, "collocation differentiation at the Gauss points needs nq >= nb (value_order >= 1)");
    static_assert(threads <= 1024, "element too large for one thread per tensor entry");
};

// out[o] = sum_i in[i] * tab; one thread per 1-D line
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
        double acc = v[0] * (TRANSPOSED_TABLE ? tab[o * N_IN] : tab[o]);
#pragma unroll
        for (int i = 1; i < N_IN; ++i)
            acc = fma(v[i], TRANSPOSED_TABLE ? tab[o * N_IN + i] : tab[i * N_OUT + o], acc);
        out[o * out_stride] = acc;
    }
}

// monomial coefficients of the multilinear map: x(xi) = sum_m c_m prod_{d in m} xi_d, m = bitmask over directions
template < int DIM >
__device__ __forceinline__ void buildGeometryCoefs(const double* __restrict__ verts, double* __restrict__ s_geo, int t, int n_threads)
{
    constexpr int nv = 1 << DIM;
    for (int i = t; i < nv * 3; i += n_threads)
    {
        const int m = i / 3, s = i % 3;
        double    acc = 0.;
#pragma unroll
        for (int v = 0; v < nv; ++v)
        {
            // sign of vertex v in monomial m: product over the directions in m of (+1 if the vertex sits at xi_d = +1 else -1)
            const int neg = __popc(m & ~v) & 1;
            acc += neg ? -verts[v * 3 + s] : verts[v * 3 + s];
        }
        s_geo[i] = acc * (1. / nv);
    }
}

// x_s (s < 3) and Jt[d][s] = dx_s/dxi_d at xi from the monomial coefficients
template < int DIM >
__device__ __forceinline__ void geometryFromCoefs(const double* __restrict__ c, const double (&xi)[DIM], double (&xs)[3], double (&Jt)[DIM][DIM])
{
    if constexpr (DIM == 2)
    {
#pragma unroll
        for (int s = 0; s < 3; ++s)
            xs[s] = fma(xi[0], fma(xi[1], c[3 * 3 + s], c[1 * 3 + s]), fma(xi[1], c[2 * 3 + s], c[s]));
#pragma unroll
        for (int s = 0; s < 2; ++s)
        {
            Jt[0][s] = fma(xi[1], c[3 * 3 + s], c[1 * 3 + s]);
            Jt[1][s] = fma(xi[0], c[3 * 3 + s], c[2 * 3 + s]);
        }
    }
    else
    {
        const double xy = xi[0] * xi[1], xz = xi[0] * xi[2], yz = xi[1] * xi[2];
        // monomial index bits: 1 = xi, 2 = eta, 4 = zeta
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            const double c0 = c[s], cx = c[3 + s], cy = c[6 + s], cxy = c[9 + s], cz = c[12 + s], cxz = c[15 + s], cyz = c[18 + s], cxyz = c[21 + s];
            Jt[0][s] = fma(yz, cxyz, fma(xi[2], cxz, fma(xi[1], cxy, cx)));
            Jt[1][s] = fma(xz, cxyz, fma(xi[2], cyz, fma(xi[0], cxy, cy)));
            Jt[2][s] = fma(xy, cxyz, fma(xi[1], cyz, fma(xi[0], cxz, cz)));
            xs[s]    = fma(xi[2], Jt[2][s], fma(xy, cxy, fma(xi[1], cy, fma(xi[0], cx, c0))));
        }
    }
}

template < typename KernelT, int DIM, int P, int NQ, int NRHS >
__global__ void __launch_bounds__(MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::threads, MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::min_blocks)
    mfSumFactApplyKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ SumFactTables< P + 1, NQ > tab)
{
    using Cfg = MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NB = Cfg::NB, F0 = Cfg::F0, F = Cfg::F, NN = Cfg::NN, Q = Cfg::Q;
    extern __shared__ double smem[];
    double*         s_colloc = smem; // [m][q]
    double*         s_pts    = s_colloc + NQ * NQ;
    double*         s_w      = s_pts + NQ;
    const int       slot     = threadIdx.x / Cfg::TPE;
    const int       t        = threadIdx.x % Cfg::TPE;
    const long long wi       = static_cast< long long >(blockIdx.x) * Cfg::EPB + slot;
    const bool      active   = wi < args.n_work;
    const long long e        = active ? (args.work_elems ? args.work_elems[wi] : args.first_elem + wi) : 0;
    double*         bufA     = s_w + NQ + static_cast< size_t >(slot) * Cfg::smem_doubles_per_elem;
    double*         bufB     = bufA + F * Q;
    double*         s_geo    = bufB + F * Q + Cfg::extra_q * Q;
    const uint32_t* el_nodes = args.nodes + e * NN;

    for (int i = threadIdx.x; i < NQ * NQ; i += Cfg::threads)
        s_colloc[i] = tab.colloc[i];
    for (int i = threadIdx.x; i < NQ; i += Cfg::threads)
    {
        s_pts[i] = tab.pts[i];
        s_w[i]   = tab.w[i];
    }
    if (active)
        buildGeometryCoefs< DIM >(args.verts + e * (1 << DIM) * 3, s_geo, t, Cfg::TPE);

    // ---- gather (MatrixFreeSystem.hpp:421-467): bufA[f][a], f = rhs*U + u, then the NF external fields
    if (active)
        for (int a = t; a < NN; a += Cfg::TPE)
        {
            const long long node = el_nodes[a];
            if (U % 2 == 0 and args.contiguous_dofs)
            {
                // dofs of a node are the U contiguous entries node*U .. node*U+U-1: vectorised 16-byte loads
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                {
                    const double2* src = reinterpret_cast< const double2* >(args.x + node * U + r * args.ld);
#pragma unroll
                    for (int u2 = 0; u2 < U / 2; ++u2)
                    {
                        double2 v = __ldg(src + u2);
                        if (args.dir_mask)
                        {
                            if (args.dir_mask[node * U + 2 * u2])
                                v.x = 0.;
                            if (args.dir_mask[node * U + 2 * u2 + 1])
                                v.y = 0.;
                        }
                        bufA[(r * U + 2 * u2) * Q + a]     = v.x;
                        bufA[(r * U + 2 * u2 + 1) * Q + a] = v.y;
                    }
                }
            }
            else
            {
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    const long long dof = node * args.dofs_per_node + args.dof_inds[u];
                    const bool      dir = isDirichlet(args.dir_mask, dof);
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                        bufA[(r * U + u) * Q + a] = dir ? 0. : args.x[dof + r * args.ld];
                }
            }
#pragma unroll
            for (int f = 0; f < NF; ++f)
                bufA[(F0 + f) * Q + a] = args.fields[node + args.field_inds[f] * args.field_stride];
        }
    __syncthreads();

    // ---- interpolate to the Gauss points, one direction at a time (bufA → bufB → bufA [→ bufB])
    if constexpr (DIM == 2)
    {
        for (int l = t; l < F * NB; l += Cfg::TPE)
        {
            const int f = l / NB, j = l % NB;
            sweepLine< NB, NQ, false >(bufA + f * Q + j * NB, 1, bufB + f * Q + j * NQ, 1, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F * NQ; l += Cfg::TPE)
        {
            const int f = l / NQ, qx = l % NQ;
            sweepLine< NB, NQ, false >(bufB + f * Q + qx, NQ, bufA + f * Q + qx, NQ, tab.interp);
        }
        __syncthreads();
    }
    else
    {
        for (int l = t; l < F * NB * NB; l += Cfg::TPE)
        {
            const int f = l / (NB * NB), kj = l % (NB * NB);
            sweepLine< NB, NQ, false >(bufA + f * Q + kj * NB, 1, bufB + f * Q + kj * NQ, 1, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F * NB * NQ; l += Cfg::TPE)
        {
            const int f = l / (NB * NQ), k = (l / NQ) % NB, qx = l % NQ;
            sweepLine< NB, NQ, false >(bufB + f * Q + k * NB * NQ + qx, NQ, bufA + f * Q + k * NQ * NQ + qx, NQ, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F * NQ * NQ; l += Cfg::TPE)
        {
            const int f = l / (NQ * NQ), qyx = l % (NQ * NQ);
            sweepLine< NB, NQ, false >(bufA + f * Q + qyx, NQ * NQ, bufB + f * Q + qyx, NQ * NQ, tab.interp);
        }
        __syncthreads();
    }
    double* uq = DIM == 2 ? bufA : bufB; // values at the Gauss points, [f][q], q = (qz*NQ + qy)*NQ + qx
    double* vq = DIM == 2 ? bufB : bufA; // the other ping-pong buffer
    // flux arrays r_d (d = 1..DIM), each F0 x Q: laid over [vq | uq | extra] *after* every thread has read uq
    double* const region2 = bufB + F * Q;
    const auto    r_ptr   = [&](int d, int f) -> double* {
        const int idx = d * F0 + f; // position in the virtual list of D*F0 arrays
        if (idx < F)
            return vq + idx * Q;
        if (idx < 2 * F)
            return uq + (idx - F) * Q;
        return region2 + (idx - 2 * F) * Q;
    };

    // ---- quadrature point stage (SumFactorization.hpp:614-756); one point per thread (TPE == Q)
    const int q     = t;
    const int qi[3] = {q % NQ, (q / NQ) % NQ, DIM == 3 ? q / (NQ * NQ) : 0};
    double    r0[F0], rd[DIM][F0];
    if (active)
    {
        double val[F], dref[DIM][F];
#pragma unroll
        for (int f = 0; f < F; ++f)
        {
            const double* line = uq + f * Q;
            val[f]             = line[q];
            int stride         = 1;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                const int base = q - qi[d] * stride;
                double    acc  = s_colloc[qi[d]] * line[base];
#pragma unroll
                for (int m = 1; m < NQ; ++m)
                    acc = fma(s_colloc[m * NQ + qi[d]], line[base + m * stride], acc);
                dref[d][f] = acc;
                stride *= NQ;
            }
        }
        double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM];
        double wq = 1.;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            xi[d] = s_pts[qi[d]];
            wq *= s_w[qi[d]];
        }
        geometryFromCoefs< DIM >(s_geo, xi, xs, Jt);
        const double detJ = invert< DIM >(Jt, Jti);
        // Jt[d][s] = dx_s/dxi_d, so Jti[s][d] = dxi_d/dx_s = the reference's jac_inv(d, s)
        typename KernelT::Input in;
#pragma unroll
        for (int f = 0; f < NF; ++f)
        {
            in.field_vals[f] = val[F0 + f];
#pragma unroll
            for (int s = 0; s < DIM; ++s)
            {
                double acc = Jti[s][0] * dref[0][F0 + f];
#pragma unroll
                for (int d = 1; d < DIM; ++d)
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
        // guard: entries the compile-time probe declared structurally zero must be zero (folds away when provable)
        bool violated = false;
        staticFor< DIM + 1 >([&](auto op) {
            staticFor< E >([&](auto eq) {
                staticFor< U >([&](auto u) {
                    if constexpr (not Sp::nz(op, eq, u))
                        violated |= res.operators[op](eq, u) != 0.;
                });
            });
        });
        if (violated)
            atomicOr(args.status, status_sparsity_violation);
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
        {
            double g[DIM][U]; // physical gradients of the operand
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                {
                    double acc = Jti[s][0] * dref[0][r * U + u];
#pragma unroll
                    for (int d = 1; d < DIM; ++d)
                        acc = fma(Jti[s][d], dref[d][r * U + u], acc);
                    g[s][u] = acc;
                }
            double tv[E];
            staticFor< E >([&](auto eq) {
                double acc = 0.;
                staticFor< U >([&](auto u) {
                    if constexpr (Sp::nz(0, eq, u))
                        acc = fma(res.operators[0](eq, u), val[r * U + u], acc);
                    staticFor< DIM >([&](auto s) {
                        if constexpr (Sp::nz(s + 1, eq, u))
                            acc = fma(res.operators[s + 1](eq, u), g[s][u], acc);
                    });
                });
                tv[eq] = acc * wgt;
            });
            staticFor< U >([&](auto u) {
                double a0 = 0., ps[DIM];
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                    ps[s] = 0.;
                staticFor< E >([&](auto eq) {
                    if constexpr (Sp::nz(0, eq, u))
                        a0 = fma(res.operators[0](eq, u), tv[eq], a0);
                    staticFor< DIM >([&](auto s) {
                        if constexpr (Sp::nz(s + 1, eq, u))
                            ps[s] = fma(res.operators[s + 1](eq, u), tv[eq], ps[s]);
                    });
                });
                r0[r * U + u] = a0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    double acc = Jti[0][d] * ps[0];
#pragma unroll
                    for (int s = 1; s < DIM; ++s)
                        acc = fma(Jti[s][d], ps[s], acc);
                    rd[d][r * U + u] = acc;
                }
            });
        }
    }
    __syncthreads(); // every thread is done reading uq: its storage may now be overwritten by the fluxes
    if (active)
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int f = 0; f < F0; ++f)
                r_ptr(d, f)[q] = rd[d][f];
    __syncthreads();

    // ---- transposed collocation derivative: v(q) = r0(q) + sum_d sum_n colloc[q_d][n] r_d(.., n, ..); kept in registers
    double vreg[F0];
    if (active)
    {
#pragma unroll
        for (int f = 0; f < F0; ++f)
        {
            double acc    = r0[f];
            int    stride = 1;
#pragma unroll
            for (int d = 0; d < DIM; ++d)
            {
                const double* line = r_ptr(d, f) + (q - qi[d] * stride);
#pragma unroll
                for (int n = 0; n < NQ; ++n)
                    acc = fma(s_colloc[qi[d] * NQ + n], line[n * stride], acc);
                stride *= NQ;
            }
            vreg[f] = acc;
        }
    }
    __syncthreads(); // all flux reads done: vq is free again
    if (active)
#pragma unroll
        for (int f = 0; f < F0; ++f)
            vq[f * Q + q] = vreg[f];
    __syncthreads();

    // ---- project back to the nodes (transposed interpolation sweeps), last direction first
    double* res_nodes;
    if constexpr (DIM == 2)
    {
        for (int l = t; l < F0 * NQ; l += Cfg::TPE)
        {
            const int f = l / NQ, qx = l % NQ;
            sweepLine< NQ, NB, true >(vq + f * Q + qx, NQ, uq + f * Q + qx, NQ, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F0 * NB; l += Cfg::TPE)
        {
            const int f = l / NB, j = l % NB;
            sweepLine< NQ, NB, true >(uq + f * Q + j * NQ, 1, vq + f * Q + j * NB, 1, tab.interp);
        }
        __syncthreads();
        res_nodes = vq;
    }
    else
    {
        for (int l = t; l < F0 * NQ * NQ; l += Cfg::TPE)
        {
            const int f = l / (NQ * NQ), qyx = l % (NQ * NQ);
            sweepLine< NQ, NB, true >(vq + f * Q + qyx, NQ * NQ, uq + f * Q + qyx, NQ * NQ, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F0 * NB * NQ; l += Cfg::TPE)
        {
            const int f = l / (NB * NQ), k = (l / NQ) % NB, qx = l % NQ;
            sweepLine< NQ, NB, true >(uq + f * Q + k * NQ * NQ + qx, NQ, vq + f * Q + k * NB * NQ + qx, NQ, tab.interp);
        }
        __syncthreads();
        for (int l = t; l < F0 * NB * NB; l += Cfg::TPE)
        {
            const int f = l / (NB * NB), kj = l % (NB * NB);
            sweepLine< NQ, NB, true >(vq + f * Q + kj * NQ, 1, uq + f * Q + kj * NB, 1, tab.interp);
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
                    atomicAdd(args.y + dof + r * args.ld, args.alpha * res_nodes[(r * U + u) * Q + a]);
            }
        }
}
} // namespace l3b
#endif
