// Integrals of residual kernels over domains and boundaries: computeIntegral / computeNormL2 of the reference (post/Integral.hpp:11-121,
// post/NormL2.hpp:21-60). One CTA per work item (element, or element + side), one thread per quadrature point; the integrand is
// `jacobian * kernel(input)` (Integral.hpp:22-27, 45-50) summed with the quadrature weights, the norm variant squares every
// component first (NormL2.hpp:21-28; the caller doubles the quadrature order and takes the square root). Not a hot path: the
// examples call it once per solve to validate the solution.
#ifndef L3B_INTEGRATE_CUH
#define L3B_INTEGRATE_CUH

#include "local_element.cuh"

namespace l3b
{
template < typename KernelT, int DIM, int P >
struct IntegrateCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, NF = params.n_fields, NRHS = params.n_rhs, NN = cpow(P + 1, DIM), NV = E * NRHS;
    static constexpr int  n_warp = local_threads / 32;
    static constexpr int  off_nv = 8 * 3, off_red = off_nv + NN * NF, total = off_red + NV * (n_warp + 1);
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
};

// args.y = the NV sums (accumulated with atomics, zeroed by the caller); args.n_cols != 0 selects the squared integrand
template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(local_threads) integrateKernel(const KernelT kernel, const __grid_constant__ ElemArgs args)
{
    using Cfg              = IntegrateCfg< KernelT, DIM, P >;
    constexpr int  NF = Cfg::NF, NN = Cfg::NN, NV = Cfg::NV;
    constexpr bool is_bnd = KernelT::is_boundary;
    extern __shared__ double smem[];
    double* s_verts = smem;
    double* s_nv    = smem + Cfg::off_nv;
    double* s_red   = smem + Cfg::off_red;

    const long long wi   = blockIdx.x;
    const long long e    = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side = is_bnd ? args.work_sides[wi] : -1;
    const int       tid  = threadIdx.x;
    constexpr int   nv   = 1 << DIM;
    for (int i = tid; i < nv * 3; i += local_threads)
        s_verts[i] = args.verts[e * nv * 3 + i];
    if constexpr (NF > 0)
        for (int a = tid; a < NN; a += local_threads)
        {
            const long long node = args.nodes[e * NN + a];
            for (int f = 0; f < NF; ++f)
                s_nv[a * NF + f] = args.fields[node + args.field_inds[f] * args.field_stride]; // FieldAccess::getGloballyIndexed
        }
    __syncthreads();

    const long long tab_off  = is_bnd ? static_cast< long long >(side) * args.n_qp : 0;
    const double*   tab_vals = args.tab_vals + tab_off * NN;
    const double*   tab_ders = args.tab_ders + tab_off * DIM * NN;
    const double*   tab_pts  = args.tab_pts + tab_off * DIM;
    const double*   tab_wts  = args.tab_wts + tab_off;

    double acc[NV];
    for (int i = 0; i < NV; ++i)
        acc[i] = 0.;
    for (int q = tid; q < args.n_qp; q += local_threads)
    {
        double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], nrm[DIM];
        for (int d = 0; d < DIM; ++d)
            xi[d] = tab_pts[q * DIM + d];
        geometryAt< DIM >(s_verts, xi, xs, Jt);
        double jac = invert< DIM >(Jt, Jti);
        if constexpr (is_bnd)
            jac = boundaryMeasureAndNormal< DIM >(side, Jt, nrm);
        typename KernelT::Input in;
        if constexpr (NF > 0)
        {
            const double* bv = tab_vals + static_cast< long long >(q) * NN;
            const double* bd = tab_ders + static_cast< long long >(q) * DIM * NN;
            for (int f = 0; f < NF; ++f)
            {
                in.field_vals[f] = 0.;
                for (int s = 0; s < DIM; ++s)
                    in.field_ders[s][f] = 0.;
            }
            for (int a = 0; a < NN; ++a) // computeFieldVals / computeFieldDers (AssembleLocalSystem.hpp:54-75)
            {
                double pd[DIM];
                for (int s = 0; s < DIM; ++s)
                {
                    double t = 0.;
                    for (int d = 0; d < DIM; ++d)
                        t = fma(Jti[s][d], bd[d * NN + a], t);
                    pd[s] = t;
                }
                const double n = bv[a];
                for (int f = 0; f < NF; ++f)
                {
                    const double v   = s_nv[a * NF + f];
                    in.field_vals[f] = fma(n, v, in.field_vals[f]);
                    for (int s = 0; s < DIM; ++s)
                        in.field_ders[s][f] = fma(pd[s], v, in.field_ders[s][f]);
                }
            }
        }
        for (int s = 0; s < 3; ++s)
            in.point.space.coords[s] = xs[s];
        in.point.time = args.time;
        if constexpr (is_bnd)
            for (int s = 0; s < DIM; ++s)
                in.normal[s] = nrm[s];
        const auto   res = kernel(in);
        const double w   = jac * tab_wts[q];
        for (int i = 0; i < NV; ++i)
            acc[i] = fma(args.n_cols != 0 ? res.v[i] * res.v[i] : res.v[i], w, acc[i]);
    }
    blockReduce< NV >(acc, s_red);
    if (tid < NV)
        atomicAdd(args.y + tid, s_red[tid]);
}

// computeValuesAtNodes (algsys/ComputeValuesAtNodes.hpp:316-593): a residual kernel evaluated AT THE NODES of the visited elements (or of
// the visited element sides, with the outward normal): the reference location of the node, the fields and their physical gradients
// there, the physical point and time go in; equation `eq` of the result goes to dof (node, dof_inds[eq]). A node shared by several
// elements receives one contribution per element and the contributions are averaged (averageElementContributions, :112-154).
// Launch passes (args.n_cols): 0 = zero the visited dofs (zeroOut, :92-110), 1 = accumulate values (args.y, ld = args.ld) and
// contribution counts (args.diag). The caller then divides by the counts (and, over ranks, export-adds both first).
// Dense tables (args.tab_*) are taken at the node locations: point a of the table is local node a.
template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(local_threads) valuesAtNodesKernel(const KernelT kernel, const __grid_constant__ ElemArgs args)
{
    using Cfg              = IntegrateCfg< KernelT, DIM, P >;
    constexpr int  E = Cfg::E, NF = Cfg::NF, NN = Cfg::NN, NRHS = Cfg::NRHS, NB = P + 1;
    constexpr bool is_bnd = KernelT::is_boundary;
    static_assert(E <= max_unknowns, "more equations than dof slots");
    extern __shared__ double smem[];
    double* s_verts = smem;
    double* s_nv    = smem + Cfg::off_nv;

    const long long wi   = blockIdx.x;
    const long long e    = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side = is_bnd ? args.work_sides[wi] : -1;
    const int       tid  = threadIdx.x;
    constexpr int   nv   = 1 << DIM;
    for (int i = tid; i < nv * 3; i += local_threads)
        s_verts[i] = args.verts[e * nv * 3 + i];
    if constexpr (NF > 0)
        for (int a = tid; a < NN; a += local_threads)
        {
            const long long node = args.nodes[e * NN + a];
            for (int f = 0; f < NF; ++f)
                s_nv[a * NF + f] = args.fields[node + args.field_inds[f] * args.field_stride];
        }
    __syncthreads();
    for (int a = tid; a < NN; a += local_threads)
    {
        if constexpr (is_bnd)
        {
            // side -> fixed lattice coordinate (mesh/ElementTraits.hpp:72-98, 118-137)
            const int idx[3] = {a % NB, (a / NB) % NB, a / (NB * NB)};
            const int axis   = DIM == 2 ? (side < 2 ? 1 : 0) : (side < 2 ? 2 : side < 4 ? 1 : 0);
            if (idx[axis] != ((side & 1) ? NB - 1 : 0))
                continue;
        }
        const long long node = args.nodes[e * NN + a];
        if (args.n_cols == 0)
        {
            for (int eq = 0; eq < E; ++eq)
                for (int r = 0; r < NRHS; ++r)
                    args.y[node * args.dofs_per_node + args.dof_inds[eq] + r * args.ld] = 0.;
            continue;
        }
        double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], nrm[DIM];
        for (int d = 0; d < DIM; ++d)
            xi[d] = args.tab_pts[a * DIM + d];
        geometryAt< DIM >(s_verts, xi, xs, Jt);
        invert< DIM >(Jt, Jti);
        if constexpr (is_bnd)
            boundaryMeasureAndNormal< DIM >(side, Jt, nrm);
        typename KernelT::Input in;
        if constexpr (NF > 0)
        {
            const double* bv = args.tab_vals + static_cast< long long >(a) * NN;
            const double* bd = args.tab_ders + static_cast< long long >(a) * DIM * NN;
            for (int f = 0; f < NF; ++f)
            {
                in.field_vals[f] = 0.;
                for (int s = 0; s < DIM; ++s)
                    in.field_ders[s][f] = 0.;
            }
            for (int b = 0; b < NN; ++b)
            {
                double pd[DIM];
                for (int s = 0; s < DIM; ++s)
                {
                    double t = 0.;
                    for (int d = 0; d < DIM; ++d)
                        t = fma(Jti[s][d], bd[d * NN + b], t);
                    pd[s] = t;
                }
                const double n = bv[b];
                for (int f = 0; f < NF; ++f)
                {
                    const double v   = s_nv[b * NF + f];
                    in.field_vals[f] = fma(n, v, in.field_vals[f]);
                    for (int s = 0; s < DIM; ++s)
                        in.field_ders[s][f] = fma(pd[s], v, in.field_ders[s][f]);
                }
            }
        }
        for (int s = 0; s < 3; ++s)
            in.point.space.coords[s] = xs[s];
        in.point.time = args.time;
        if constexpr (is_bnd)
            for (int s = 0; s < DIM; ++s)
                in.normal[s] = nrm[s];
        const auto res = kernel(in);
        for (int eq = 0; eq < E; ++eq)
        {
            const long long dof = node * args.dofs_per_node + args.dof_inds[eq];
            atomicAdd(args.diag + dof, 1.);
            for (int r = 0; r < NRHS; ++r)
                atomicAdd(args.y + dof + r * args.ld, res.v[eq + r * E]);
        }
    }
}
} // namespace l3b
#endif
