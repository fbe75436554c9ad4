#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
// tests/LocalOperatorCommon.hpp fixtures: quad p=4, asm_opts{.value_order = 2} → nq = 9; tests/Diffusion2D.hpp: p=2, nq=3
L3B_REGISTER_DOMAIN_KERNEL(diffusion_kernel_2D, kernels::Diffusion2D,
                           (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 3, .n_rhs = 2}), L3B_PQ(4, 9), L3B_PQ(4, 5));
L3B_REGISTER_DOMAIN_KERNEL(diffusion_kernel_2D_var, kernels::Diffusion2DVar,
                           (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 3, .n_fields = 1, .n_rhs = 2}), L3B_PQ(4, 9),
                           L3B_PQ(3, 4));
L3B_REGISTER_DOMAIN_KERNEL(diffusion_kernel_2D_r1, kernels::Diffusion2D, (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 3}),
                           L3B_PQ(2, 3), L3B_PQ(4, 5));
L3B_REGISTER_BOUNDARY_KERNEL(adiabatic_bc_2D, kernels::AdiabaticBC2D, (KernelParams{.dimension = 2, .n_equations = 1, .n_unknowns = 3}),
                             L3B_PQ(2, 3), L3B_PQ(4, 5));
// examples/02-diffusion-2D (BASELINE config 1): quad p=4, default options → nq = 5
L3B_REGISTER_DOMAIN_KERNEL(example02_domain, kernels::Example02Domain, (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 3}),
                           L3B_PQ(4, 5), L3B_PQ(2, 3));
L3B_REGISTER_BOUNDARY_KERNEL(example02_bc, kernels::Example02BC, (KernelParams{.dimension = 2, .n_equations = 1, .n_unknowns = 3}),
                             L3B_PQ(4, 5), L3B_PQ(2, 3));
// BASELINE configs[4], 2-D half of the order sweep (benchmarks/LocalAssemblyBenchmarks.cpp:41-87 runs diff2d on quads): orders 1..8
L3B_REGISTER_DOMAIN_KERNEL(bench_diffusion2d, kernels::Diffusion2D, (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 3}),
                           L3B_PQ(1, 2), L3B_PQ(2, 3), L3B_PQ(3, 4), L3B_PQ(4, 5), L3B_PQ(5, 6), L3B_PQ(6, 7), L3B_PQ(7, 8), L3B_PQ(8, 9));
// tests/MultiDomainTest.cpp: one unknown per domain
L3B_REGISTER_DOMAIN_KERNEL(multidomain_mass, kernels::MultiDomainMass, (KernelParams{.dimension = 2, .n_equations = 1, .n_unknowns = 1}),
                           L3B_PQ(2, 3), L3B_PQ(4, 5));
