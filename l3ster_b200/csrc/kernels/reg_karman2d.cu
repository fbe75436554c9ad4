#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
// examples/07-karman-2D (BASELINE configs[3]): quad p=4, AssemblyOptions{1, 1} -> nq = 8 (source.cpp:81-82); p=2 -> nq = 4 for small tests
L3B_REGISTER_DOMAIN_KERNEL(karman_steady, kernels::KarmanSteady, (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 4, .n_fields = 2}),
                           L3B_PQ(4, 8), L3B_PQ(2, 4));
L3B_REGISTER_DOMAIN_KERNEL(karman_transient, kernels::KarmanTransient,
                           (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 4, .n_fields = 4}), L3B_PQ(4, 8), L3B_PQ(2, 4));
L3B_REGISTER_BOUNDARY_KERNEL(karman_outlet, kernels::KarmanOutlet, (KernelParams{.dimension = 2, .n_equations = 2, .n_unknowns = 3}),
                             L3B_PQ(4, 8), L3B_PQ(2, 4), L3B_PQ(4, 5));
