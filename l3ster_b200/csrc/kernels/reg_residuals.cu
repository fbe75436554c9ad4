#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
// residual kernels (integrands): tests/Diffusion2D.hpp:80-97, examples/07-karman-2D/source.cpp:158-166, and probes; listed per element order
L3B_REGISTER_DOMAIN_RESIDUAL_KERNEL(diffusion2d_error_dom, kernels::Diffusion2DError, (KernelParams{.dimension = 2, .n_equations = 3, .n_fields = 3}), 2, 4);
L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(diffusion2d_error_bnd, kernels::Diffusion2DError, (KernelParams{.dimension = 2, .n_equations = 3, .n_fields = 3}), 2, 4);
L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(karman_flowrate, kernels::KarmanFlowRate, (KernelParams{.dimension = 2, .n_equations = 1, .n_fields = 2}), 2, 4);
L3B_REGISTER_DOMAIN_RESIDUAL_KERNEL(integrand_probe_2D, kernels::IntegrandProbe2D, (KernelParams{.dimension = 2, .n_equations = 3, .n_fields = 2}), 1, 2, 4);
L3B_REGISTER_DOMAIN_RESIDUAL_KERNEL(integrand_probe_3D, kernels::IntegrandProbe3D, (KernelParams{.dimension = 3, .n_equations = 3, .n_fields = 2}), 1, 2, 3, 4);
L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(boundary_probe_2D, kernels::BoundaryProbe2D, (KernelParams{.dimension = 2, .n_equations = 3, .n_fields = 2}), 1, 2, 4);
L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(boundary_probe_3D, kernels::BoundaryProbe3D, (KernelParams{.dimension = 3, .n_equations = 3, .n_fields = 2}), 1, 2, 3, 4);
L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(karman_inlet, kernels::KarmanInlet, (KernelParams{.dimension = 2, .n_equations = 2}), 2, 4);
