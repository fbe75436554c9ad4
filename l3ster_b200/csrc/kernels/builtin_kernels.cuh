// Device-callable restatements of the equation kernels used by the reference's tests, benchmarks and examples, written in
// the reference's own idiom (structured bindings on `out`, `A(i, j) = ...`). Each is registered for the element orders
// the parity tests and the benchmarks need (see the .cu files next to this header).
#ifndef L3B_BUILTIN_KERNELS_CUH
#define L3B_BUILTIN_KERNELS_CUH

#include "../kernel_interface.cuh"

namespace l3b::kernels
{
// tests/Kernels.hpp:5-25
struct Diffusion2D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In&, Out& out) const
    {
        auto& [operators, rhs] = out;
        auto& [A0, Ax, Ay]     = operators;
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
};

// tests/Kernels.hpp:27-54
struct Diffusion2DVar
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [field_vals, field_ders, _] = in;
        const auto lambda                       = field_vals[0];
        const auto& [dx, dy]                    = field_ders;
        const auto dl_dx                        = dx[0];
        const auto dl_dy                        = dy[0];
        auto& [operators, rhs] = out;
        auto& [A0, Ax, Ay]     = operators;
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
};

// tests/Kernels.hpp:56-83; with `source`: benchmarks/Diffusion3D.hpp:50-79 and benchmarks/Kernels.hpp:91-118
template < bool with_source >
struct Diffusion3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In&, Out& out) const
    {
        auto& [operators, rhs] = out;
        auto& [A0, Ax, Ay, Az] = operators;
        constexpr double k = 1.; // diffusivity
        constexpr double s = 1.; // source
        // -k * div q = s
        Ax(0, 1) = -k;
        Ay(0, 2) = -k;
        Az(0, 3) = -k;
        if constexpr (with_source)
            rhs[0] = s;
        // grad T = q
        A0(1, 1) = -1.;
        Ax(1, 0) = 1.;
        A0(2, 2) = -1.;
        Ay(2, 0) = 1.;
        A0(3, 3) = -1.;
        Az(3, 0) = 1.;
        // rot q = 0
        Ay(4, 3) = 1.;
        Az(4, 2) = -1.;
        Ax(5, 3) = -1.;
        Az(5, 1) = 1.;
        Ax(6, 2) = 1.;
        Ay(6, 1) = -1.;
    }
};

// tests/Kernels.hpp:85-118
struct Diffusion3DVar
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [field_vals, field_ders, _] = in;
        const auto lambda                       = field_vals[0];
        const auto& [dx, dy, dz]                = field_ders;
        const auto dl_dx                        = dx[0];
        const auto dl_dy                        = dy[0];
        const auto dl_dz                        = dz[0];
        auto& [operators, rhs] = out;
        auto& [A0, Ax, Ay, Az] = operators;
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
};

// tests/Kernels.hpp:120-128
struct AdiabaticBC2D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [vals, ders, point, normal] = in;
        auto& [operators, rhs]                  = out;
        auto& [A0, A1, A2]                      = operators;
        A0(0, 1) = normal[0];
        A0(0, 2) = normal[1];
    }
};

// 3-D boundary equation kernel on the dofs (T, qx, qy, qz) of the diffusion system. Equation 0: adiabatic wall q.n = 0 — the 3-D form
// of tests/Kernels.hpp:120-128. Equation 1: a Robin condition written with DERIVATIVE operators and the normal, dT/dn + h T = h T_inf(x)
// — the 3-D form of the Neumann kernel of examples/06-matrix-free/source.cpp:79-86 plus a point-dependent right-hand side, so that
// A0, A1..A3, the normal and the physical point of the boundary quadrature are all exercised.
struct RobinBC3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [vals, ders, point, normal] = in;
        auto& [operators, rhs]                  = out;
        auto& [A0, A1, A2, A3]                  = operators;
        constexpr double h = 0.5;
        A0(0, 1) = normal[0];
        A0(0, 2) = normal[1];
        A0(0, 3) = normal[2];
        A0(1, 0) = h;
        A1(1, 0) = normal[0];
        A2(1, 0) = normal[1];
        A3(1, 0) = normal[2];
        rhs(1, 0) = h * (1. + point.space.x() + 2. * point.space.y() - point.space.z());
    }
};

// benchmarks/Kernels.hpp:3-65 — the incompressible Navier-Stokes kernel of LocalAssemblyBenchmarks.cpp:41-87 (velocity-vorticity-pressure
// first-order system: U = 7 unknowns (u, v, w, p, ox, oy, oz), E = 8 equations, linearised about the n_fields = 7 nodal fields)
struct NS3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [vals, ders, point]             = in;
        const auto& [u, v, w, p, ox, oy, oz]        = vals;
        const auto& [x_ders, y_ders, z_ders]        = ders;
        const auto& [ux, vx, wx, px, oxx, oyx, ozx] = x_ders;
        const auto& [uy, vy, wy, py, oxy, oyy, ozy] = y_ders;
        const auto& [uz, vz, wz, pz, oxz, oyz, ozz] = z_ders;
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2, A3] = operators;
        constexpr double Re_inv = 1e-3;
        // momentum (rows 0-2): convection linearised about the fields, pressure gradient, curl of the vorticity
        A0(0, 0) = ux, A0(0, 1) = uy, A0(0, 2) = uz;
        A0(1, 0) = vx, A0(1, 1) = vy, A0(1, 2) = vz;
        A0(2, 0) = wx, A0(2, 1) = wy, A0(2, 2) = wz;
        A1(0, 0) = u, A1(1, 1) = u, A1(2, 2) = u;
        A2(0, 0) = v, A2(1, 1) = v, A2(2, 2) = v;
        A3(0, 0) = w, A3(1, 1) = w, A3(2, 2) = w;
        A1(0, 3) = 1., A2(0, 3) = 1., A3(0, 3) = 1.;
        A2(0, 6) = Re_inv, A3(0, 5) = -Re_inv;
        A1(1, 6) = -Re_inv, A3(1, 4) = Re_inv;
        A1(2, 5) = Re_inv, A2(2, 4) = -Re_inv;
        // vorticity definition (rows 3-5)
        A0(3, 4) = 1., A2(3, 2) = 1., A3(3, 1) = -1.;
        A0(4, 5) = 1., A1(4, 2) = -1., A3(4, 0) = 1.;
        A0(5, 6) = 1., A1(5, 1) = 1., A2(5, 0) = -1.;
        // continuity (row 6), solenoidal vorticity (row 7)
        A1(6, 0) = 1., A2(6, 1) = 1., A3(6, 2) = 1.;
        A1(7, 4) = 1., A2(7, 5) = 1., A3(7, 6) = 1.;
        rhs(0, 0) = u * ux + v * uy + w * uz;
        rhs(1, 0) = u * vx + v * vy + w * vz;
        rhs(2, 0) = u * wx + v * wy + w * wz;
    }
};

// tests/MultiDomainTest.cpp:38-46: one unknown, A0 = 1, rhs = the value to set (the test captures it in the lambda; a registered functor
// is stateless, so the value travels as the `time` argument of assembleProblem)
struct MultiDomainMass
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        A0(0, 0)               = 1.;
        rhs(0, 0)              = in.point.time;
    }
};

// examples/02-diffusion-2D/source.cpp:45-67
struct Example02Domain
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In&, Out& out) const
    {
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        A1(0, 1) = -1.;
        A2(0, 2) = -1.;
        rhs[0]   = 1.;
        A0(1, 1) = -1.;
        A1(1, 0) = 1.;
        A0(2, 2) = -1.;
        A2(2, 0) = 1.;
        A1(3, 2) = 1.;
        A2(3, 1) = -1.;
    }
};

// examples/02-diffusion-2D/source.cpp:68-81
struct Example02BC
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& normal = in.normal;
        const auto  nx     = normal[0];
        const auto  ny     = normal[1];
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        A0(0, 0) = 1.;
        A0(0, 1) = nx;
        A0(0, 2) = ny;
    }
};

// examples/07-karman-2D/source.cpp:21-155 (BASELINE configs[3]): unknowns (u, v, vorticity, p), nu = 1 / Re with Re = 100, dt = 0.05
namespace karman
{
inline constexpr double nu = 1. / 100., dt = .05;
inline constexpr int    IU = 0, IV = 1, IO = 2, IP = 3;
template < typename Op, typename Rhs >
L3B_HD constexpr void fillSteady(Op& A0, Op& A1, Op& A2, Rhs& rhs, double u, double v, double du_dx, double dv_dx, double du_dy, double dv_dy)
{
    // momentum equations
    A0(0, IU) = du_dx;
    A0(0, IV) = du_dy;
    A1(0, IU) = u;
    A1(0, IP) = 1.;
    A2(0, IU) = v;
    A2(0, IO) = nu;
    rhs(0, 0) = u * du_dx + v * du_dy;
    A0(1, IU) = dv_dx;
    A0(1, IV) = dv_dy;
    A1(1, IV) = u;
    A1(1, IO) = -nu;
    A2(1, IV) = v;
    A2(1, IP) = 1.;
    rhs(1, 0) = u * dv_dx + v * dv_dy;
    // incompressibility
    A1(2, IU) = 1.;
    A2(2, IV) = 1.;
    // vorticity definition
    A0(3, IO) = 1.;
    A1(3, IV) = -1.;
    A2(3, IU) = 1.;
}
} // namespace karman
struct KarmanSteady
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [field_vals, field_ders, point] = in;
        const auto& [u, v]                          = field_vals; // velocity of the previous iteration
        const auto& [x_ders, y_ders]                = field_ders;
        const auto& [du_dx, dv_dx]                  = x_ders;
        const auto& [du_dy, dv_dy]                  = y_ders;
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        karman::fillSteady(A0, A1, A2, rhs, u, v, du_dx, dv_dx, du_dy, dv_dy);
    }
};
struct KarmanTransient
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [field_vals, field_ders, point]  = in;
        const auto& [u1, v1, u2, v2]                 = field_vals; // velocities of the 2 previous time steps
        const auto& [x_ders, y_ders]                 = field_ders;
        const auto& [du1_dx, dv1_dx, du2_dx, dv2_dx] = x_ders;
        const auto& [du1_dy, dv1_dy, du2_dy, dv2_dy] = y_ders;
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        const double u = 2 * u1 - u2, v = 2 * v1 - v2; // extrapolated in time
        const double du_dx = 2 * du1_dx - du2_dx, dv_dx = 2 * dv1_dx - dv2_dx;
        const double du_dy = 2 * du1_dy - du2_dy, dv_dy = 2 * dv1_dy - dv2_dy;
        karman::fillSteady(A0, A1, A2, rhs, u, v, du_dx, dv_dx, du_dy, dv_dy);
        A0(0, karman::IU) += 1.5 / karman::dt;
        A0(1, karman::IV) += 1.5 / karman::dt;
        rhs(0, 0) += (2 * u1 - .5 * u2) / karman::dt;
        rhs(1, 0) += (2 * v1 - .5 * v2) / karman::dt;
        for (size_t op = 0; op < 3; ++op) // scale the momentum equations by dt
            for (size_t unknown = 0; unknown != 4; ++unknown)
                for (size_t eq = 0; eq != 2; ++eq)
                    operators[op](eq, unknown) *= karman::dt;
        for (size_t eq = 0; eq != 2; ++eq)
            rhs(eq, 0) *= karman::dt;
    }
};
struct KarmanOutlet // on the dofs (u, v, p)
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const double nx        = in.normal[0];
        const double ny        = in.normal[1];
        auto& [operators, rhs] = out;
        auto& [A0, A1, A2]     = operators;
        A0(0, 2) = -nx;
        A1(0, 0) = karman::nu * nx;
        A2(0, 0) = karman::nu * ny;
        A0(1, 2) = -ny;
        A1(1, 1) = karman::nu * nx;
        A2(1, 1) = karman::nu * ny;
    }
};

// Not in the reference: dense, space-/time-/field-dependent probes (same formulas as oracle/kernels.cpp) that leave no
// structural zero for an indexing or layout mistake to hide behind.
struct DenseProbe3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        constexpr int E = 5, U = 3;
        for (int i = 0; i <= 3; ++i)
            for (int e = 0; e < E; ++e)
                for (int u = 0; u < U; ++u)
                    out.operators[i](e, u) = 0.1 * (i + 1) + 0.01 * (e + 1) * (u + 2) + 0.3 * in.point.space[0] -
                                             0.2 * in.point.space[1] * (i == 2) + 0.05 * in.field_vals[0] * (e == u) +
                                             0.07 * in.field_ders[(e + u) % 3][1] + 0.01 * in.point.time;
        for (int e = 0; e < E; ++e)
            out.rhs[e] = 1. + 0.5 * e + in.point.space[0] * in.field_vals[1];
    }
};
struct DenseProbe2D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        constexpr int E = 4, U = 2;
        for (int i = 0; i <= 2; ++i)
            for (int e = 0; e < E; ++e)
                for (int u = 0; u < U; ++u)
                    out.operators[i](e, u) = 0.1 * (i + 1) + 0.01 * (e + 1) * (u + 2) + 0.3 * in.point.space[0] -
                                             0.2 * in.point.space[1] * (i == 2) + 0.05 * in.field_vals[0] * (e == u) +
                                             0.07 * in.field_ders[(e + u) % 2][1] + 0.01 * in.point.time;
        for (int e = 0; e < E; ++e)
            out.rhs[e] = 1. + 0.5 * e + in.point.space[1] * in.field_vals[1];
    }
};

// ---- residual kernels: integrands of computeIntegral / computeNormL2 (post/Integral.hpp, post/NormL2.hpp)
// tests/Diffusion2D.hpp:82-91 (node_dist.back() == 1), used as domain and as boundary residual kernel
struct Diffusion2DError
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& error) const
    {
        const auto& point             = in.point;
        const auto& vals              = in.field_vals;
        const auto& [T, dT_dx, dT_dy] = vals;
        error[0]                      = T - point.space.x() / 1.;
        error[1]                      = dT_dx - 1. / 1.;
        error[2]                      = dT_dy;
    }
};
// examples/07-karman-2D/source.cpp:158-166
struct KarmanFlowRate
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const auto& [field_vals, field_ders, point, normal] = in;
        const auto& [u, v]                                  = field_vals;
        const double nx = normal[0], ny = normal[1];
        out[0] = u * nx + v * ny;
    }
};
// examples/07-karman-2D/source.cpp:167-171: parabolic inlet velocity profile, the kernel of setDirichletBCValues(kernel_inlet, {inlet}, {IU, IV})
struct KarmanInlet
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        const double y = in.point.space.y();
        out[0]         = 1.5 * (1. - y * y);
        out[1]         = 0.;
    }
};
// not in the reference: polynomial / field probes with closed-form integrals on boxes (same bodies as oracle/kernels.cpp)
struct IntegrandProbe2D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        out[0] = 1.;
        out[1] = in.point.space[0] * in.point.space[0] * in.point.space[1] + in.point.time;
        out[2] = in.field_vals[0] + 0.5 * in.field_ders[0][1] - in.field_ders[1][0] * in.point.space[1];
    }
};
struct IntegrandProbe3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        out[0] = 1.;
        out[1] = in.point.space[0] * in.point.space[1] * in.point.space[1] * in.point.space[2] + in.point.time;
        out[2] = in.field_vals[0] + 0.5 * in.field_ders[0][1] - in.field_ders[2][0] * in.point.space[1];
    }
};
struct BoundaryProbe2D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        out[0] = 1.;
        out[1] = in.point.space[0] * in.normal[0] + in.point.space[1] * in.normal[1];
        out[2] = in.field_vals[0] * in.normal[0] + in.field_ders[1][1];
    }
};
struct BoundaryProbe3D
{
    template < typename In, typename Out >
    L3B_HD constexpr void operator()(const In& in, Out& out) const
    {
        out[0] = 1.;
        out[1] = in.point.space[0] * in.normal[0] + in.point.space[1] * in.normal[1] + in.point.space[2] * in.normal[2];
        out[2] = in.field_vals[0] * in.normal[2] + in.field_ders[1][1];
    }
};
} // namespace l3b::kernels
#endif
