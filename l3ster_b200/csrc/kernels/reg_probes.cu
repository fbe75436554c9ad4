#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
L3B_REGISTER_DOMAIN_KERNEL(dense_probe_3D, kernels::DenseProbe3D,
                           (KernelParams{.dimension = 3, .n_equations = 5, .n_unknowns = 3, .n_fields = 2}), L3B_PQ(2, 3), L3B_PQ(3, 6));
L3B_REGISTER_DOMAIN_KERNEL(dense_probe_2D, kernels::DenseProbe2D,
                           (KernelParams{.dimension = 2, .n_equations = 4, .n_unknowns = 2, .n_fields = 2}), L3B_PQ(3, 4), L3B_PQ(2, 4));
