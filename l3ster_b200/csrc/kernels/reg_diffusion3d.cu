#include "../register_kernel.cuh"
#include "builtin_kernels.cuh"
using namespace l3b;
// BASELINE configs 2/3/5: benchmarks/Diffusion3D.hpp:50-79 — hex, U=4, E=7
L3B_REGISTER_DOMAIN_KERNEL(bench_diffusion3d, kernels::Diffusion3D< true >, (KernelParams{.dimension = 3, .n_equations = 7, .n_unknowns = 4}),
                           L3B_PQ(1, 2), L3B_PQ(2, 3), L3B_PQ(3, 4), L3B_PQ(4, 5), L3B_PQ(5, 6), L3B_PQ(6, 7), L3B_PQ(7, 8), L3B_PQ(8, 9));
#ifdef L3B_ASM_TIMING
// experiment support: per-warp phase clocks of assembleDmmaKernel (this translation unit's copy of the table)
extern "C" int l3b_debug_asm_timing(long long* out, int n)
{
    return cudaMemcpyFromSymbol(out, l3b::g_asm_timing, sizeof(long long) * 64 * n) == cudaSuccess ? 0 : 1;
}
#endif
