// CPU ORACLE — test infrastructure only (see l3ster_oracle.hpp).
// Restatements of the equation kernels the reference's own tests, benchmarks and examples use on the hot path.
#include "l3ster_oracle.hpp"

#include <map>

namespace orc
{
namespace
{
using In  = const KernelInput&;
using Out = const KernelOutput&;

// tests/Kernels.hpp:5-25
void diffusion_kernel_2D(In, Out out)
{
    auto& A0 = out.operators[0];
    auto& Ax = out.operators[1];
    auto& Ay = out.operators[2];
    constexpr double lambda = 1.;
    Ax(0, 1) = -lambda;
    Ay(0, 2) = -lambda;
    A0(1, 1) = -1.;
    Ax(1, 0) = 1.;
    A0(2, 2) = -1.;
    Ay(2, 0) = 1.;
    Ax(3, 2) = 1.;
    Ay(3, 1) = -1.;
}

// tests/Kernels.hpp:27-54
void diffusion_kernel_2D_var(In in, Out out)
{
    const auto lambda = in.field_vals[0];
    const auto dl_dx  = in.field_ders[0][0];
    const auto dl_dy  = in.field_ders[1][0];
    auto&      A0     = out.operators[0];
    auto&      Ax     = out.operators[1];
    auto&      Ay     = out.operators[2];
    A0(0, 1) = -dl_dx;
    A0(0, 2) = -dl_dy;
    Ax(0, 1) = -lambda;
    Ay(0, 2) = -lambda;
    A0(1, 1) = -1.;
    Ax(1, 0) = 1.;
    A0(2, 2) = -1.;
    Ay(2, 0) = 1.;
    Ax(3, 2) = 1.;
    Ay(3, 1) = -1.;
}

// tests/Kernels.hpp:56-83; benchmarks/Kernels.hpp:91-118 adds rhs[0] = s, benchmarks/Diffusion3D.hpp:50-79 likewise
template < bool with_source >
void diffusion_kernel_3D(In, Out out)
{
    auto& A0 = out.operators[0];
    auto& Ax = out.operators[1];
    auto& Ay = out.operators[2];
    auto& Az = out.operators[3];
    constexpr double k = 1.;
    constexpr double s = 1.;
    Ax(0, 1) = -k;
    Ay(0, 2) = -k;
    Az(0, 3) = -k;
    if constexpr (with_source)
        out.rhs[0] = s;
    A0(1, 1) = -1.;
    Ax(1, 0) = 1.;
    A0(2, 2) = -1.;
    Ay(2, 0) = 1.;
    A0(3, 3) = -1.;
    Az(3, 0) = 1.;
    Ay(4, 3) = 1.;
    Az(4, 2) = -1.;
    Ax(5, 3) = -1.;
    Az(5, 1) = 1.;
    Ax(6, 2) = 1.;
    Ay(6, 1) = -1.;
}

// tests/Kernels.hpp:85-118
void diffusion_kernel_3D_var(In in, Out out)
{
    const auto lambda = in.field_vals[0];
    const auto dl_dx  = in.field_ders[0][0];
    const auto dl_dy  = in.field_ders[1][0];
    const auto dl_dz  = in.field_ders[2][0];
    auto&      A0     = out.operators[0];
    auto&      Ax     = out.operators[1];
    auto&      Ay     = out.operators[2];
    auto&      Az     = out.operators[3];
    A0(0, 1) = -dl_dx;
    A0(0, 2) = -dl_dy;
    A0(0, 3) = -dl_dz;
    Ax(0, 1) = -lambda;
    Ay(0, 2) = -lambda;
    Az(0, 3) = -lambda;
    A0(1, 1) = -1.;
    Ax(1, 0) = 1.;
    A0(2, 2) = -1.;
    Ay(2, 0) = 1.;
    A0(3, 3) = -1.;
    Az(3, 0) = 1.;
    Ay(4, 3) = 1.;
    Az(4, 2) = -1.;
    Ax(5, 3) = -1.;
    Az(5, 1) = 1.;
    Ax(6, 2) = 1.;
    Ay(6, 1) = -1.;
}

// tests/Kernels.hpp:120-128
void adiabatic_bc_2D(In in, Out out)
{
    auto& A0 = out.operators[0];
    A0(0, 1) = in.normal[0];
    A0(0, 2) = in.normal[1];
}

// 3-D boundary kernel on (T, qx, qy, qz): row 0 the 3-D form of tests/Kernels.hpp:120-128 (adiabatic wall), row 1 a Robin condition
// in derivative form, the 3-D form of examples/06-matrix-free/source.cpp:79-86 with a point-dependent right-hand side
void robin_bc_3D(In in, Out out)
{
    auto& A0 = out.operators[0];
    auto& A1 = out.operators[1];
    auto& A2 = out.operators[2];
    auto& A3 = out.operators[3];
    constexpr double h = 0.5;
    A0(0, 1)   = in.normal[0];
    A0(0, 2)   = in.normal[1];
    A0(0, 3)   = in.normal[2];
    A0(1, 0)   = h;
    A1(1, 0)   = in.normal[0];
    A2(1, 0)   = in.normal[1];
    A3(1, 0)   = in.normal[2];
    out.rhs[1] = h * (1. + in.space[0] + 2. * in.space[1] - in.space[2]);
}

// tests/MultiDomainTest.cpp:38-46 (the value to set travels as the time argument)
void multidomain_mass(In in, Out out)
{
    out.operators[0](0, 0) = 1.;
    out.rhs[0]             = in.time;
}

// examples/02-diffusion-2D/source.cpp:45-67
void example02_domain(In, Out out)
{
    auto& A0 = out.operators[0];
    auto& A1 = out.operators[1];
    auto& A2 = out.operators[2];
    A1(0, 1)   = -1.;
    A2(0, 2)   = -1.;
    out.rhs[0] = 1.;
    A0(1, 1)   = -1.;
    A1(1, 0)   = 1.;
    A0(2, 2)   = -1.;
    A2(2, 0)   = 1.;
    A1(3, 2)   = 1.;
    A2(3, 1)   = -1.;
}

// examples/02-diffusion-2D/source.cpp:68-81
void example02_bc(In in, Out out)
{
    auto& A0 = out.operators[0];
    A0(0, 0) = 1.;
    A0(0, 1) = in.normal[0];
    A0(0, 2) = in.normal[1];
}

// examples/07-karman-2D/source.cpp:21-78: unknowns (u, v, vorticity, p), nu = 1 / Re, Re = 100
constexpr double karman_nu = 1. / 100., karman_dt = .05;
void karmanFillSteady(Out out, double u, double v, double du_dx, double dv_dx, double du_dy, double dv_dy)
{
    constexpr int IU = 0, IV = 1, IO = 2, IP = 3;
    auto&         A0 = out.operators[0];
    auto&         A1 = out.operators[1];
    auto&         A2 = out.operators[2];
    A0(0, IU) = du_dx;
    A0(0, IV) = du_dy;
    A1(0, IU) = u;
    A1(0, IP) = 1.;
    A2(0, IU) = v;
    A2(0, IO) = karman_nu;
    out.rhs(0, 0) = u * du_dx + v * du_dy;
    A0(1, IU) = dv_dx;
    A0(1, IV) = dv_dy;
    A1(1, IV) = u;
    A1(1, IO) = -karman_nu;
    A2(1, IV) = v;
    A2(1, IP) = 1.;
    out.rhs(1, 0) = u * dv_dx + v * dv_dy;
    A1(2, IU) = 1.;
    A2(2, IV) = 1.;
    A0(3, IO) = 1.;
    A1(3, IV) = -1.;
    A2(3, IU) = 1.;
}
// examples/07-karman-2D/source.cpp:84-98
void karman_steady(In in, Out out)
{
    karmanFillSteady(out, in.field_vals[0], in.field_vals[1], in.field_ders[0][0], in.field_ders[0][1], in.field_ders[1][0], in.field_ders[1][1]);
}
// examples/07-karman-2D/source.cpp:101-136 (BDF2)
void karman_transient(In in, Out out)
{
    const double u1 = in.field_vals[0], v1 = in.field_vals[1], u2 = in.field_vals[2], v2 = in.field_vals[3];
    const double u = 2 * u1 - u2, v = 2 * v1 - v2;
    const double du_dx = 2 * in.field_ders[0][0] - in.field_ders[0][2], dv_dx = 2 * in.field_ders[0][1] - in.field_ders[0][3];
    const double du_dy = 2 * in.field_ders[1][0] - in.field_ders[1][2], dv_dy = 2 * in.field_ders[1][1] - in.field_ders[1][3];
    karmanFillSteady(out, u, v, du_dx, dv_dx, du_dy, dv_dy);
    out.operators[0](0, 0) += 1.5 / karman_dt;
    out.operators[0](1, 1) += 1.5 / karman_dt;
    out.rhs(0, 0) += (2 * u1 - .5 * u2) / karman_dt;
    out.rhs(1, 0) += (2 * v1 - .5 * v2) / karman_dt;
    for (int op = 0; op < 3; ++op)
        for (int unknown = 0; unknown < 4; ++unknown)
            for (int eq = 0; eq < 2; ++eq)
                out.operators[op](eq, unknown) *= karman_dt;
    for (int eq = 0; eq < 2; ++eq)
        out.rhs(eq, 0) *= karman_dt;
}
// examples/07-karman-2D/source.cpp:139-155: outlet condition on the dofs (u, v, p)
void karman_outlet(In in, Out out)
{
    const double nx = in.normal[0], ny = in.normal[1];
    auto&        A0 = out.operators[0];
    auto&        A1 = out.operators[1];
    auto&        A2 = out.operators[2];
    A0(0, 2) = -nx;
    A1(0, 0) = karman_nu * nx;
    A2(0, 0) = karman_nu * ny;
    A0(1, 2) = -ny;
    A1(1, 1) = karman_nu * nx;
    A2(1, 1) = karman_nu * ny;
}

// benchmarks/Kernels.hpp:3-65
void ns3d_kernel(In in, Out out)
{
    const auto* vals = in.field_vals;
    const auto  u = vals[0], v = vals[1], w = vals[2];
    const auto* xd = in.field_ders[0];
    const auto* yd = in.field_ders[1];
    const auto* zd = in.field_ders[2];
    const auto  ux = xd[0], vx = xd[1], wx = xd[2];
    const auto  uy = yd[0], vy = yd[1], wy = yd[2];
    const auto  uz = zd[0], vz = zd[1], wz = zd[2];
    auto&       A0 = out.operators[0];
    auto&       A1 = out.operators[1];
    auto&       A2 = out.operators[2];
    auto&       A3 = out.operators[3];
    constexpr double Re_inv = 1e-3;
    A0(0, 0) = ux;
    A0(0, 1) = uy;
    A0(0, 2) = uz;
    A0(1, 0) = vx;
    A0(1, 1) = vy;
    A0(1, 2) = vz;
    A0(2, 0) = wx;
    A0(2, 1) = wy;
    A0(2, 2) = wz;
    A0(3, 4) = 1.;
    A0(4, 5) = 1.;
    A0(5, 6) = 1.;
    A1(0, 0) = u;
    A1(0, 3) = 1.;
    A1(1, 1) = u;
    A1(1, 6) = -Re_inv;
    A1(2, 2) = u;
    A1(2, 5) = Re_inv;
    A1(4, 2) = -1.;
    A1(5, 1) = 1.;
    A1(6, 0) = 1.;
    A1(7, 4) = 1.;
    A2(0, 0) = v;
    A2(0, 3) = 1.;
    A2(0, 6) = Re_inv;
    A2(1, 1) = v;
    A2(2, 2) = v;
    A2(2, 4) = -Re_inv;
    A2(3, 2) = 1.;
    A2(5, 0) = -1.;
    A2(6, 1) = 1.;
    A2(7, 5) = 1.;
    A3(0, 0) = w;
    A3(0, 3) = 1.;
    A3(0, 5) = -Re_inv;
    A3(1, 1) = w;
    A3(1, 4) = Re_inv;
    A3(2, 2) = w;
    A3(3, 1) = -1.;
    A3(4, 0) = 1.;
    A3(6, 2) = 1.;
    A3(7, 6) = 1.;
    out.rhs[0] = u * ux + v * uy + w * uz;
    out.rhs[1] = u * vx + v * vy + w * vz;
    out.rhs[2] = u * wx + v * wy + w * wz;
}

// A deliberately space- and time-dependent dense kernel (not in the reference): exercises point.space / time / every
// operator entry, so that layout or indexing mistakes cannot hide behind structural zeros. 3D: E=5, U=3, NF=2.
void dense_probe_3D(In in, Out out)
{
    const int E = 5, U = 3;
    for (int i = 0; i <= 3; ++i)
        for (int e = 0; e < E; ++e)
            for (int u = 0; u < U; ++u)
                out.operators[i](e, u) = 0.1 * (i + 1) + 0.01 * (e + 1) * (u + 2) + 0.3 * in.space[0] - 0.2 * in.space[1] * (i == 2) +
                                         0.05 * in.field_vals[0] * (e == u) + 0.07 * in.field_ders[(e + u) % 3][1] + 0.01 * in.time;
    for (int e = 0; e < E; ++e)
        out.rhs[e] = 1. + 0.5 * e + in.space[0] * in.field_vals[1];
}
void dense_probe_2D(In in, Out out)
{
    const int E = 4, U = 2;
    for (int i = 0; i <= 2; ++i)
        for (int e = 0; e < E; ++e)
            for (int u = 0; u < U; ++u)
                out.operators[i](e, u) = 0.1 * (i + 1) + 0.01 * (e + 1) * (u + 2) + 0.3 * in.space[0] - 0.2 * in.space[1] * (i == 2) +
                                         0.05 * in.field_vals[0] * (e == u) + 0.07 * in.field_ders[(e + u) % 2][1] + 0.01 * in.time;
    for (int e = 0; e < E; ++e)
        out.rhs[e] = 1. + 0.5 * e + in.space[1] * in.field_vals[1];
}

// ---- residual kernels (integrands of computeIntegral / computeNormL2): they fill `rhs` only
// tests/Diffusion2D.hpp:82-91 (node_dist.back() == 1), used as domain and as boundary residual kernel
void diffusion2d_error(In in, Out out)
{
    const auto T = in.field_vals[0], dT_dx = in.field_vals[1], dT_dy = in.field_vals[2];
    out.rhs[0] = T - in.space[0] / 1.;
    out.rhs[1] = dT_dx - 1. / 1.;
    out.rhs[2] = dT_dy;
}
// examples/07-karman-2D/source.cpp:158-166
void karman_flowrate(In in, Out out)
{
    out.rhs[0] = in.field_vals[0] * in.normal[0] + in.field_vals[1] * in.normal[1];
}
// examples/07-karman-2D/source.cpp:167-171
void karman_inlet(In in, Out out)
{
    out.rhs[0] = 1.5 * (1. - in.space[1] * in.space[1]);
    out.rhs[1] = 0.;
}
// not in the reference: polynomial / field probes with closed-form integrals on boxes
void integrand_probe_2D(In in, Out out)
{
    out.rhs[0] = 1.;
    out.rhs[1] = in.space[0] * in.space[0] * in.space[1] + in.time;
    out.rhs[2] = in.field_vals[0] + 0.5 * in.field_ders[0][1] - in.field_ders[1][0] * in.space[1];
}
void integrand_probe_3D(In in, Out out)
{
    out.rhs[0] = 1.;
    out.rhs[1] = in.space[0] * in.space[1] * in.space[1] * in.space[2] + in.time;
    out.rhs[2] = in.field_vals[0] + 0.5 * in.field_ders[0][1] - in.field_ders[2][0] * in.space[1];
}
void boundary_probe_2D(In in, Out out)
{
    out.rhs[0] = 1.;
    out.rhs[1] = in.space[0] * in.normal[0] + in.space[1] * in.normal[1]; // closed boundary: 2 x area
    out.rhs[2] = in.field_vals[0] * in.normal[0] + in.field_ders[1][1];
}
void boundary_probe_3D(In in, Out out)
{
    out.rhs[0] = 1.;
    out.rhs[1] = in.space[0] * in.normal[0] + in.space[1] * in.normal[1] + in.space[2] * in.normal[2]; // closed boundary: 3 x volume
    out.rhs[2] = in.field_vals[0] * in.normal[2] + in.field_ders[1][1];
}

std::map< std::string, Kernel > makeRegistry()
{
    std::map< std::string, Kernel > r;
    r["diffusion2d_error_dom"]   = Kernel{{2, 3, 0, 3, 1}, false, diffusion2d_error};
    r["diffusion2d_error_bnd"]   = Kernel{{2, 3, 0, 3, 1}, true, diffusion2d_error};
    r["karman_flowrate"]         = Kernel{{2, 1, 0, 2, 1}, true, karman_flowrate};
    r["karman_inlet"]            = Kernel{{2, 2, 0, 0, 1}, true, karman_inlet};
    r["integrand_probe_2D"]      = Kernel{{2, 3, 0, 2, 1}, false, integrand_probe_2D};
    r["integrand_probe_3D"]      = Kernel{{3, 3, 0, 2, 1}, false, integrand_probe_3D};
    r["boundary_probe_2D"]       = Kernel{{2, 3, 0, 2, 1}, true, boundary_probe_2D};
    r["boundary_probe_3D"]       = Kernel{{3, 3, 0, 2, 1}, true, boundary_probe_3D};
    r["diffusion_kernel_2D"]     = Kernel{{2, 4, 3, 0, 1}, false, diffusion_kernel_2D};
    r["diffusion_kernel_2D_var"] = Kernel{{2, 4, 3, 1, 1}, false, diffusion_kernel_2D_var};
    r["diffusion_kernel_3D"]     = Kernel{{3, 7, 4, 0, 1}, false, diffusion_kernel_3D< false >};
    r["diffusion_kernel_3D_var"] = Kernel{{3, 7, 4, 1, 1}, false, diffusion_kernel_3D_var};
    r["bench_diffusion3d"]       = Kernel{{3, 7, 4, 0, 1}, false, diffusion_kernel_3D< true >};
    r["adiabatic_bc_2D"]         = Kernel{{2, 1, 3, 0, 1}, true, adiabatic_bc_2D};
    r["robin_bc_3D"]             = Kernel{{3, 2, 4, 0, 1}, true, robin_bc_3D};
    r["multidomain_mass"]        = Kernel{{2, 1, 1, 0, 1}, false, multidomain_mass};
    r["example02_domain"]        = Kernel{{2, 4, 3, 0, 1}, false, example02_domain};
    r["example02_bc"]            = Kernel{{2, 1, 3, 0, 1}, true, example02_bc};
    r["ns3d_kernel"]             = Kernel{{3, 8, 7, 7, 1}, false, ns3d_kernel};
    r["karman_steady"]           = Kernel{{2, 4, 4, 2, 1}, false, karman_steady};
    r["karman_transient"]        = Kernel{{2, 4, 4, 4, 1}, false, karman_transient};
    r["karman_outlet"]           = Kernel{{2, 2, 3, 0, 1}, true, karman_outlet};
    r["dense_probe_3D"]          = Kernel{{3, 5, 3, 2, 1}, false, dense_probe_3D};
    r["dense_probe_2D"]          = Kernel{{2, 4, 2, 2, 1}, false, dense_probe_2D};
    return r;
}
} // namespace

const Kernel& getKernel(const std::string& name)
{
    static const auto registry = makeRegistry();
    const auto        it       = registry.find(name);
    if (it == registry.end())
        throw std::invalid_argument{"unknown oracle kernel: " + name};
    return it->second;
}
} // namespace orc
