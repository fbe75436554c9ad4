#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"

using namespace l3b;
// BASELINE configs[4]: benchmarks/LocalAssemblyBenchmarks.cpp:41-87 — hex, U = 7, E = 8, n_fields = 7, quadrature order 4p - 1
// (nq = 2p: AssemblyOptions{.value_order = 2}: 2 (2p) / 2 + 1 would be 2p + 1, the benchmark passes QO = 4p - 1 directly, nq = (4p - 1) / 2 + 1 = 2p)
L3B_REGISTER_DOMAIN_KERNEL(ns3d_kernel, kernels::NS3D, (KernelParams{.dimension = 3, .n_equations = 8, .n_unknowns = 7, .n_fields = 7}),
                           L3B_PQ(2, 4), L3B_PQ(2, 3), L3B_PQ(4, 8), L3B_PQ(4, 5), L3B_PQ(3, 4));
// 3-D boundary equation kernel (adiabatic + Robin in derivative form) on the diffusion system's dofs
L3B_REGISTER_BOUNDARY_KERNEL(robin_bc_3D, kernels::RobinBC3D, (KernelParams{.dimension = 3, .n_equations = 2, .n_unknowns = 4}), L3B_PQ(2, 3),
                             L3B_PQ(3, 4), L3B_PQ(4, 5));
