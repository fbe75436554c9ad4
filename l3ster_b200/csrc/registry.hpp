// Kernel registry: every user kernel is compiled by nvcc per (element dimension, order, quadrature size) instantiation and
// registered here; the C ABI dispatches on (kernel id, order, nq). This is the device-side counterpart of the
// reference instantiating its templates in the user's translation unit (surface (i)).
#ifndef L3B_REGISTRY_HPP
#define L3B_REGISTRY_HPP

#include "device_common.cuh"
#include "tables.hpp"

#include <memory>
#include <string>
#include <vector>

namespace l3b
{
struct KernelInfo
{
    std::string name;
    int         dim = 0, n_equations = 0, n_unknowns = 0, n_fields = 0, n_rhs = 1;
    bool        is_boundary = false;
    bool        is_residual = false; // integrand of computeIntegral / computeNormL2, not an equation kernel
};

// launchers return the CUDA error of the launch; `kernel_obj` points at the registered functor instance
using MfLaunch    = cudaError_t (*)(const void* kernel_obj, const ElemArgs&, const tables::Tables1D&, cudaStream_t);
using ElemLaunch  = cudaError_t (*)(const void* kernel_obj, const ElemArgs&, cudaStream_t);

struct KernelInstance
{
    int        order = 0, nq = 0;
    MfLaunch   mf_sumfact_full = nullptr; // n_cols == n_rhs
    MfLaunch   mf_sumfact_one  = nullptr; // n_cols == 1
    ElemLaunch local_apply_full = nullptr, local_apply_one = nullptr;
    ElemLaunch init     = nullptr;
    ElemLaunch init_fast = nullptr; // domain kernels: diag + F_e without the Dirichlet lifting (mf_init.cuh)
    ElemLaunch assemble = nullptr;
    ElemLaunch integrate = nullptr; // residual kernels only; nq == 0: any quadrature size (dense tables at run time)
    ElemLaunch values_at_nodes = nullptr; // residual kernels only: computeValuesAtNodes
    // work per launch unit, for occupancy/grid decisions and reporting
    int mf_elems_per_block = 1, asm_blocks_per_elem = 1;
};

struct KernelEntry
{
    KernelInfo                    info;
    std::shared_ptr< const void > object;
    std::vector< KernelInstance > instances;
    const KernelInstance*         find(int order, int nq) const
    {
        for (const auto& i : instances)
            if (i.order == order and (i.nq == nq or i.nq == 0))
                return &i;
        return nullptr;
    }
};

std::vector< KernelEntry >& kernelRegistry();
int                         findKernel(const std::string& name);
} // namespace l3b
#endif
