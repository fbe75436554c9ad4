#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
// tests/LocalOperatorCommon.hpp fixtures: hex p=3, asm_opts{.value_order = 2} → nq = 7; n_rhs = 3 / 2
L3B_REGISTER_DOMAIN_KERNEL(diffusion_kernel_3D, kernels::Diffusion3D< false >,
                           (KernelParams{.dimension = 3, .n_equations = 7, .n_unknowns = 4, .n_rhs = 3}), L3B_PQ(3, 7), L3B_PQ(2, 3));
L3B_REGISTER_DOMAIN_KERNEL(diffusion_kernel_3D_var, kernels::Diffusion3DVar,
                           (KernelParams{.dimension = 3, .n_equations = 7, .n_unknowns = 4, .n_fields = 1, .n_rhs = 2}), L3B_PQ(3, 7),
                           L3B_PQ(2, 3));
