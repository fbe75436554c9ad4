// Non-sum-factorised element kernels: one CTA per work item (element, or element + side for boundary kernels).
//
//   MODE_APPLY : y[dofs] += alpha * K_e x[dofs], K_e x = sum_q H_q w_q H_q^T x            (algsys/EvaluateLocalOperator.hpp:94-146, 211-263)
//   MODE_INIT  : diag[dofs] += diag(K_e), rhs[dofs] += F_e - K_e g   (g = Dirichlet values) (algsys/EvaluateLocalOperator.hpp:172-208, 276-328;
//                                                                                         scatter: MatrixFreeSystem.hpp:377-390)
// where H_q (L x E) holds, for node a and unknown u, block_a(u, e) = N_a A0(e,u) + sum_s dN_a/dx_s A_s(e,u)
// (AssembleLocalSystem.hpp:131-142). Used for every boundary kernel, for LocalEvalStrategy::LocalElement, and for the
// matrix-free system's diagonal/rhs initialisation (which the reference also runs non-sum-factorised).
#ifndef L3B_LOCAL_ELEMENT_CUH
#define L3B_LOCAL_ELEMENT_CUH

#include "device_common.cuh"

namespace l3b
{
constexpr int local_threads = 128;
enum LocalMode : int
{
    MODE_APPLY = 0,
    MODE_INIT  = 1
};

// sum `vals[0..N)` over the CTA; result broadcast through `red` (size >= N + N * n_warps)
template < int N >
__device__ __forceinline__ void blockReduce(double (&vals)[N], double* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
        double v = vals[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0)
            red[N + warp * N + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < N)
    {
        double s = 0.;
        for (int w = 0; w < n_warps; ++w)
            s += red[N + w * N + threadIdx.x];
        red[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i)
        vals[i] = red[i];
    __syncthreads();
}

template < typename KernelT, int DIM, int P, int NC /* operand columns */, int MODE >
struct LocalCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields, NRHS = params.n_rhs;
    static constexpr int  NN = cpow(P + 1, DIM), L = NN * U;
    static constexpr int  NCOL   = MODE == MODE_INIT ? NRHS : NC;
    static constexpr int  n_red  = cmax(E * NCOL, cmax(NF * (DIM + 1), 1));
    static constexpr int  n_warp = local_threads / 32;
    // smem: verts | x_e (L x NCOL) | y_e (L x NCOL) | node field values (NN x NF) | A ((DIM+1) E U) | F (E NRHS) | reduction
    static constexpr int    off_x = 8 * 3, off_y = off_x + L * NCOL, off_nv = off_y + L * NCOL, off_A = off_nv + NN * NF,
                         off_F = off_A + (DIM + 1) * E * U, off_red = off_F + E * NRHS, off_dg = off_red + n_red * (n_warp + 1),
                         total = off_dg + (MODE == MODE_INIT ? L + L * NRHS : 0);
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
};

template < typename KernelT, int DIM, int P, int NC, int MODE >
__global__ void __launch_bounds__(local_threads) localElementKernel(const KernelT kernel, const __grid_constant__ ElemArgs args)
{
    using Cfg = LocalCfg< KernelT, DIM, P, NC, MODE >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NRHS = Cfg::NRHS, NN = Cfg::NN, L = Cfg::L, NCOL = Cfg::NCOL;
    constexpr bool is_bnd = KernelT::is_boundary;
    extern __shared__ double smem[];
    double* s_verts = smem;
    double* s_x     = smem + Cfg::off_x;
    double* s_y     = smem + Cfg::off_y;
    double* s_nv    = smem + Cfg::off_nv;
    double* s_A     = smem + Cfg::off_A;
    double* s_F     = smem + Cfg::off_F;
    double* s_red   = smem + Cfg::off_red;
    double* s_diag  = smem + Cfg::off_dg;
    double* s_rhs   = s_diag + L;

    const long long wi   = blockIdx.x;
    const long long e    = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side = is_bnd ? args.work_sides[wi] : -1;
    const int       tid  = threadIdx.x;
    const uint32_t* el_nodes = args.nodes + e * NN;
    constexpr int   nv       = 1 << DIM;

    for (int i = tid; i < nv * 3; i += local_threads)
        s_verts[i] = args.verts[e * nv * 3 + i];
    for (int a = tid; a < NN; a += local_threads)
    {
        const long long node = el_nodes[a];
        for (int u = 0; u < U; ++u)
        {
            const long long dof = node * args.dofs_per_node + args.dof_inds[u];
            const bool      dir = isDirichlet(args.dir_mask, dof);
            for (int c = 0; c < NCOL; ++c)
            {
                double v;
                if constexpr (MODE == MODE_APPLY)
                    v = dir ? 0. : args.x[dof + c * args.ld]; // gather (MatrixFreeSystem.hpp:393-418)
                else
                    v = dir ? args.dir_vals[dof + c * args.ld] : 0.; // gatherDirichletVals (:364-375), zero-extended
                s_x[(a * U + u) + c * L] = v;
                s_y[(a * U + u) + c * L] = 0.;
            }
            if constexpr (MODE == MODE_INIT)
            {
                s_diag[a * U + u] = 0.;
                for (int c = 0; c < NRHS; ++c)
                    s_rhs[(a * U + u) + c * L] = 0.;
            }
        }
        for (int f = 0; f < NF; ++f)
            s_nv[a * NF + f] = args.fields[node + args.field_inds[f] * args.field_stride];
    }
    __syncthreads();

    const long long tab_off  = is_bnd ? static_cast< long long >(side) * args.n_qp : 0;
    const double*   tab_vals = args.tab_vals + tab_off * NN;
    const double*   tab_ders = args.tab_ders + tab_off * DIM * NN;
    const double*   tab_pts  = args.tab_pts + tab_off * DIM;
    const double*   tab_wts  = args.tab_wts + tab_off;

    double energy = 0.;
    for (int q = 0; q < args.n_qp; ++q)
    {
        // mapping (MapReferenceToPhysical.hpp:28-89) — every thread redundantly, it is a handful of FMAs
        double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], nrm[DIM];
        for (int d = 0; d < DIM; ++d)
            xi[d] = tab_pts[q * DIM + d];
        geometryAt< DIM >(s_verts, xi, xs, Jt);
        const double detJ = invert< DIM >(Jt, Jti);
        double       jac  = detJ;
        if constexpr (is_bnd)
            jac = boundaryMeasureAndNormal< DIM >(side, Jt, nrm);
        else if (not(detJ > 0.))
        {
            if (tid == 0)
                atomicOr(args.status, status_degenerate_element); // AssembleLocalSystem.hpp:249, EvaluateLocalOperator.hpp:229
        }
        const double  weight = jac * tab_wts[q];
        const double* bv     = tab_vals + static_cast< long long >(q) * NN;
        const double* bd     = tab_ders + static_cast< long long >(q) * DIM * NN;

        // field values / physical derivatives at the point (AssembleLocalSystem.hpp:54-75)
        constexpr int n_fred = NF * (DIM + 1);
        double        fred[n_fred > 0 ? n_fred : 1];
        if constexpr (NF > 0)
        {
            for (int i = 0; i < n_fred; ++i)
                fred[i] = 0.;
            for (int a = tid; a < NN; a += local_threads)
            {
                double pd[DIM];
                for (int s = 0; s < DIM; ++s)
                {
                    double acc = 0.;
                    for (int d = 0; d < DIM; ++d)
                        acc = fma(Jti[s][d], bd[d * NN + a], acc);
                    pd[s] = acc;
                }
                const double n = bv[a];
                for (int f = 0; f < NF; ++f)
                {
                    const double v = s_nv[a * NF + f];
                    fred[f]        = fma(n, v, fred[f]);
                    for (int s = 0; s < DIM; ++s)
                        fred[NF * (s + 1) + f] = fma(pd[s], v, fred[NF * (s + 1) + f]);
                }
            }
            blockReduce< n_fred >(fred, s_red);
        }
        if (tid == 0)
        {
            typename KernelT::Input in;
            for (int f = 0; f < NF; ++f)
            {
                in.field_vals[f] = fred[f];
                for (int s = 0; s < DIM; ++s)
                    in.field_ders[s][f] = fred[NF * (s + 1) + f];
            }
            for (int s = 0; s < 3; ++s)
                in.point.space.coords[s] = xs[s]; // the non-SF paths pass the true 3-D point (AssembleLocalSystem.hpp:229)
            in.point.time = args.time;
            if constexpr (is_bnd)
                for (int s = 0; s < DIM; ++s)
                    in.normal[s] = nrm[s];
            const auto res = kernel(in);
            for (int i = 0; i <= DIM; ++i)
                for (int k = 0; k < E * U; ++k)
                    s_A[i * E * U + k] = res.operators[i].v[k];
            for (int k = 0; k < E * NRHS; ++k)
                s_F[k] = res.rhs.v[k];
        }
        __syncthreads();

        // t = w * H^T x  (E x NCOL), reduced over the CTA
        double tv[E * NCOL];
        for (int i = 0; i < E * NCOL; ++i)
            tv[i] = 0.;
        for (int a = tid; a < NN; a += local_threads)
        {
            double pd[DIM];
            for (int s = 0; s < DIM; ++s)
            {
                double acc = 0.;
                for (int d = 0; d < DIM; ++d)
                    acc = fma(Jti[s][d], bd[d * NN + a], acc);
                pd[s] = acc;
            }
            const double n = bv[a];
            for (int u = 0; u < U; ++u)
                for (int eq = 0; eq < E; ++eq)
                {
                    double b = n * s_A[eq + u * E];
                    for (int s = 0; s < DIM; ++s)
                        b = fma(pd[s], s_A[(s + 1) * E * U + eq + u * E], b);
                    for (int c = 0; c < NCOL; ++c)
                        tv[eq + c * E] = fma(b, s_x[(a * U + u) + c * L], tv[eq + c * E]);
                    if constexpr (MODE == MODE_INIT)
                    {
                        s_diag[a * U + u] = fma(b * b, weight, s_diag[a * U + u]);
                        for (int c = 0; c < NRHS; ++c)
                            s_rhs[(a * U + u) + c * L] = fma(b * s_F[eq + c * E], weight, s_rhs[(a * U + u) + c * L]);
                    }
                }
        }
        blockReduce< E * NCOL >(tv, s_red);
        if constexpr (MODE == MODE_APPLY)
            if (tid == 0 and args.energy != nullptr) // x^T A x of column 0: sum_q w |H_q^T x_e|^2
                for (int eq = 0; eq < E; ++eq)
                    energy = fma(tv[eq] * weight, tv[eq], energy);
        // y += H (w t)
        for (int a = tid; a < NN; a += local_threads)
        {
            double pd[DIM];
            for (int s = 0; s < DIM; ++s)
            {
                double acc = 0.;
                for (int d = 0; d < DIM; ++d)
                    acc = fma(Jti[s][d], bd[d * NN + a], acc);
                pd[s] = acc;
            }
            const double n = bv[a];
            for (int u = 0; u < U; ++u)
                for (int eq = 0; eq < E; ++eq)
                {
                    double b = n * s_A[eq + u * E];
                    for (int s = 0; s < DIM; ++s)
                        b = fma(pd[s], s_A[(s + 1) * E * U + eq + u * E], b);
                    for (int c = 0; c < NCOL; ++c)
                        s_y[(a * U + u) + c * L] = fma(b, tv[eq + c * E] * weight, s_y[(a * U + u) + c * L]);
                }
        }
        __syncthreads();
    }

    if constexpr (MODE == MODE_APPLY)
        if (tid == 0 and args.energy != nullptr)
            atomicAdd(args.energy, energy);
    // scatter
    for (int a = tid; a < NN; a += local_threads)
    {
        const long long node = el_nodes[a];
        for (int u = 0; u < U; ++u)
        {
            const long long dof = node * args.dofs_per_node + args.dof_inds[u];
            if constexpr (MODE == MODE_APPLY)
            {
                if (isDirichlet(args.dir_mask, dof)) // scatter (MatrixFreeSystem.hpp:469-492)
                    continue;
                for (int c = 0; c < NCOL; ++c)
                    atomicAdd(args.y + dof + c * args.ld, args.alpha * s_y[(a * U + u) + c * L]);
            }
            else
            {
                atomicAdd(args.diag + dof, s_diag[a * U + u]); // scatterInit (:377-390)
                for (int c = 0; c < NRHS; ++c)
                    atomicAdd(args.rhs + dof + c * args.ld, s_rhs[(a * U + u) + c * L] - s_y[(a * U + u) + c * L]);
            }
        }
    }
}
} // namespace l3b
#endif
