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
} // namespace l3b
#endif
