// Instantiation + registration of a user kernel for a list of (order, nq) pairs. Include from a .cu file:
//
//     struct MyKernel { template <class In, class Out> __host__ __device__ void operator()(const In& in, Out& out) const {...} };
//     L3B_REGISTER_DOMAIN_KERNEL(my_kernel, MyKernel, (l3b::KernelParams{.dimension = 3, .n_equations = 7, .n_unknowns = 4}),
//                                L3B_PQ(4, 5), L3B_PQ(6, 7));
//
// mirrors `wrapDomainEquationKernel<params>(lambda)` / `wrapBoundaryEquationKernel<params>(lambda)` of the reference
// (common/KernelInterface.hpp:178-190), plus the explicit list of element orders the reference gets implicitly from the
// mesh type.
#ifndef L3B_REGISTER_KERNEL_CUH
#define L3B_REGISTER_KERNEL_CUH

#include "assemble.cuh"
#include "assemble_dmma.cuh"
#include "integrate.cuh"
#include "local_element.cuh"
#include "mf_hex_planes.cuh"
#include "mf_init.cuh"
#include "mf_sumfact.cuh"
#include "registry.hpp"

#include <algorithm>
#include <cstdlib>

namespace l3b
{
template < auto KernelFn >
cudaError_t raiseSmemLimit(size_t bytes)
{
    static size_t current = 48 * 1024;
    if (bytes > current)
    {
        const auto err = cudaFuncSetAttribute(KernelFn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast< int >(bytes));
        if (err != cudaSuccess)
            return err;
        current = bytes;
    }
    return cudaSuccess;
}

template < typename KernelT, int DIM, int P, int NQ, int NC >
cudaError_t launchMfSumFact(const void* obj, const ElemArgs& args, const tables::Tables1D& t, cudaStream_t stream)
{
    using Cfg = MfSumFactCfg< KernelT, DIM, P, NQ, NC >;
    if (args.n_work == 0)
        return cudaSuccess;
    SumFactTables< P + 1, NQ > tab;
    std::copy(t.interp.begin(), t.interp.end(), tab.interp);
    std::copy(t.der.begin(), t.der.end(), tab.der);
    std::copy(t.colloc.begin(), t.colloc.end(), tab.colloc);
    std::copy(t.w.begin(), t.w.end(), tab.w);
    std::copy(t.pts.begin(), t.pts.end(), tab.pts);
    constexpr auto fn = mfSumFactApplyKernel< KernelT, DIM, P, NQ, NC >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    static const cudaError_t carveout =
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, static_cast< int >(cudaSharedmemCarveoutMaxShared));
    if (carveout != cudaSuccess)
        return carveout;
    const auto grid = static_cast< unsigned >((args.n_work + Cfg::EPB - 1) / Cfg::EPB);
    fn<<< grid, Cfg::threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args, tab);
    return cudaGetLastError();
}

// hexahedra: planes + columns kernel (mf_hex_planes.cuh) when its register tiles fit, else the line-per-thread kernel.
// L3B_MF_LINES=1 in the environment forces the line-per-thread kernel (A/B measurements, parity of both paths).
inline bool forceLineKernel()
{
    static const bool force = [] {
        const char* e = std::getenv("L3B_MF_LINES");
        return e != nullptr and e[0] == '1';
    }();
    return force;
}
template < typename KernelT, int P, int NQ, int NC, bool ENERGY >
cudaError_t launchMfHexImpl(const void* obj, const ElemArgs& args, const SumFactTables< P + 1, NQ >& tab, cudaStream_t stream)
{
    using Cfg         = MfHexCfg< KernelT, P, NQ, NC >;
    constexpr auto fn = mfHexPlanesKernel< KernelT, P, NQ, NC, ENERGY >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::launch_smem); err != cudaSuccess)
        return err;
    static const cudaError_t carveout =
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, static_cast< int >(cudaSharedmemCarveoutMaxShared));
    if (carveout != cudaSuccess)
        return carveout;
    // persistent grid: one CTA per resident slot of the device, batches strided over the (virtual) CTAs
    static const int resident = [] {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, Cfg::launch_threads, Cfg::launch_smem);
        return std::max(1, sms * std::max(1, per_sm));
    }();
    const long long n_batches = (args.n_work + Cfg::EPB - 1) / Cfg::EPB;
    const auto      grid      = static_cast< unsigned >(std::min< long long >((n_batches + Cfg::WG - 1) / Cfg::WG, std::max(1, resident - args.reserve_ctas)));
    fn<<< grid, Cfg::launch_threads, Cfg::launch_smem, stream >>>(*static_cast< const KernelT* >(obj), args, tab);
    return cudaGetLastError();
}
template < typename KernelT, int P, int NQ, int NC >
cudaError_t launchMfHex(const void* obj, const ElemArgs& args, const tables::Tables1D& t, cudaStream_t stream)
{
    if (forceLineKernel())
        return launchMfSumFact< KernelT, 3, P, NQ, NC >(obj, args, t, stream);
    if (args.n_work == 0)
        return cudaSuccess;
    SumFactTables< P + 1, NQ > tab;
    std::copy(t.interp.begin(), t.interp.end(), tab.interp);
    std::copy(t.der.begin(), t.der.end(), tab.der);
    std::copy(t.colloc.begin(), t.colloc.end(), tab.colloc);
    std::copy(t.w.begin(), t.w.end(), tab.w);
    std::copy(t.pts.begin(), t.pts.end(), tab.pts);
    // the energy variant only for single-column applies (CG): the full-n_rhs instantiations stay single
    if constexpr (NC == 1)
        if (args.energy != nullptr)
            return launchMfHexImpl< KernelT, P, NQ, NC, true >(obj, args, tab, stream);
    if (args.energy != nullptr)
        return cudaErrorNotSupported;
    return launchMfHexImpl< KernelT, P, NQ, NC, false >(obj, args, tab, stream);
}

template < typename KernelT, int DIM, int P, int NC, int MODE >
cudaError_t launchLocal(const void* obj, const ElemArgs& args, cudaStream_t stream)
{
    using Cfg = LocalCfg< KernelT, DIM, P, NC, MODE >;
    if (args.n_work == 0)
        return cudaSuccess;
    constexpr auto fn = localElementKernel< KernelT, DIM, P, NC, MODE >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    fn<<< static_cast< unsigned >(args.n_work), local_threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args);
    return cudaGetLastError();
}

// Which assembly kernel: the DMMA kernel (assemble_dmma.cuh) from 32 nodes per element up, the register-tiled DFMA kernel
// (assemble.cuh) below — the measured crossover (profiles/r1_order_sweep.jsonl, hex U=4 E=7: p=1 DFMA 5x faster, p=2 equal, p>=3 DMMA
// 2.8-3.7x faster; small elements waste the 32 x 32 DMMA warp tiles). L3B_ASM_FMA=1 / L3B_ASM_FMA=0 force one or the other.
inline int forcedAssemblyKernel() // -1 auto, 0 DMMA, 1 DFMA
{
    static const int force = [] {
        const char* e = std::getenv("L3B_ASM_FMA");
        return e == nullptr ? -1 : e[0] == '1' ? 1 : 0;
    }();
    return force;
}
template < typename KernelT, int DIM, int P >
cudaError_t launchMfInitFast(const void* obj, const ElemArgs& args, cudaStream_t stream)
{
    using Cfg = MfInitCfg< KernelT, DIM, P >;
    if (args.n_work == 0)
        return cudaSuccess;
    constexpr auto fn = mfInitDiagRhsKernel< KernelT, DIM, P >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    fn<<< static_cast< unsigned >(args.n_work), Cfg::threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args);
    return cudaGetLastError();
}

template < typename KernelT, int DIM, int P >
cudaError_t launchAssemble(const void* obj, const ElemArgs& args, cudaStream_t stream)
{
    using Cfg = AsmCfg< KernelT, DIM, P >;
    if (args.n_work == 0)
        return cudaSuccess;
    if (forcedAssemblyKernel() == 0 or (forcedAssemblyKernel() < 0 and AsmDmmaCfg< KernelT, DIM, P >::NN >= 32))
    {
        using DCfg                  = AsmDmmaCfg< KernelT, DIM, P >;
        static const AsmPairs pairs = DCfg::makePairs();
        if (pairs.tiles_per_elem == 0)
            return cudaSuccess;
        constexpr auto dfn = assembleDmmaKernel< KernelT, DIM, P >;
        if (const auto err = raiseSmemLimit< dfn >(DCfg::smem_bytes); err != cudaSuccess)
            return err;
        dfn<<< static_cast< unsigned >(args.n_work * pairs.tiles_per_elem), DCfg::threads, DCfg::smem_bytes, stream >>>(
            *static_cast< const KernelT* >(obj), args, pairs);
        return cudaGetLastError();
    }
    constexpr auto fn = assembleKernel< KernelT, DIM, P >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    fn<<< static_cast< unsigned >(args.n_work * Cfg::n_pairs), asm_threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args);
    return cudaGetLastError();
}

template < typename KernelT, int DIM, int P >
cudaError_t launchIntegrate(const void* obj, const ElemArgs& args, cudaStream_t stream)
{
    using Cfg = IntegrateCfg< KernelT, DIM, P >;
    if (args.n_work == 0)
        return cudaSuccess;
    constexpr auto fn = integrateKernel< KernelT, DIM, P >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    fn<<< static_cast< unsigned >(args.n_work), local_threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args);
    return cudaGetLastError();
}

template < typename KernelT, int DIM, int P >
cudaError_t launchValuesAtNodes(const void* obj, const ElemArgs& args, cudaStream_t stream)
{
    using Cfg = IntegrateCfg< KernelT, DIM, P >;
    if (args.n_work == 0)
        return cudaSuccess;
    constexpr auto fn = valuesAtNodesKernel< KernelT, DIM, P >;
    if (const auto err = raiseSmemLimit< fn >(Cfg::smem_bytes); err != cudaSuccess)
        return err;
    fn<<< static_cast< unsigned >(args.n_work), local_threads, Cfg::smem_bytes, stream >>>(*static_cast< const KernelT* >(obj), args);
    return cudaGetLastError();
}

template < int P_, int NQ_ >
struct PQ
{
    static constexpr int P = P_, NQ = NQ_;
};

template < typename KernelT, int DIM, typename pq >
KernelInstance makeInstance()
{
    constexpr int  P = pq::P, NQ = pq::NQ;
    constexpr int  NRHS = KernelT::parameters.n_rhs;
    KernelInstance inst;
    inst.order = P;
    inst.nq    = NQ;
    if constexpr (not KernelT::is_boundary)
    {
        inst.mf_sumfact_full    = launchMfSumFact< KernelT, DIM, P, NQ, NRHS >;
        inst.mf_sumfact_one     = launchMfSumFact< KernelT, DIM, P, NQ, 1 >;
        inst.mf_elems_per_block = MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::EPB;
        if constexpr (DIM == 3)
        {
            if constexpr (MfHexCfg< KernelT, P, NQ, NRHS >::supported)
            {
                inst.mf_sumfact_full    = launchMfHex< KernelT, P, NQ, NRHS >;
                inst.mf_elems_per_block = MfHexCfg< KernelT, P, NQ, NRHS >::EPB;
            }
            if constexpr (MfHexCfg< KernelT, P, NQ, 1 >::supported)
                inst.mf_sumfact_one = launchMfHex< KernelT, P, NQ, 1 >;
        }
    }
    inst.local_apply_full    = launchLocal< KernelT, DIM, P, NRHS, MODE_APPLY >;
    inst.local_apply_one     = launchLocal< KernelT, DIM, P, 1, MODE_APPLY >;
    inst.init                = launchLocal< KernelT, DIM, P, NRHS, MODE_INIT >;
    if constexpr (not KernelT::is_boundary)
        inst.init_fast = launchMfInitFast< KernelT, DIM, P >; // diag + F_e only: the caller applies the Dirichlet lifting
    inst.assemble            = launchAssemble< KernelT, DIM, P >;
    inst.asm_blocks_per_elem = AsmCfg< KernelT, DIM, P >::n_pairs;
    return inst;
}

// residual kernels: one instance per element order, any quadrature size
template < typename KernelT, int... orders >
int registerResidualKernel(const char* name, const KernelT& kernel)
{
    constexpr auto params = KernelT::parameters;
    KernelEntry    entry;
    entry.info.name        = name;
    entry.info.dim         = params.dimension;
    entry.info.n_equations = static_cast< int >(params.n_equations);
    entry.info.n_unknowns  = 0;
    entry.info.n_fields    = static_cast< int >(params.n_fields);
    entry.info.n_rhs       = static_cast< int >(params.n_rhs);
    entry.info.is_boundary = KernelT::is_boundary;
    entry.info.is_residual = true;
    entry.object           = std::make_shared< KernelT >(kernel);
    const auto add         = [&]< int P >(std::integral_constant< int, P >) {
        KernelInstance inst;
        inst.order     = P;
        inst.nq        = 0;
        inst.integrate = launchIntegrate< KernelT, params.dimension, P >;
        if constexpr (KernelT::parameters.n_equations <= static_cast< size_t >(max_unknowns))
            inst.values_at_nodes = launchValuesAtNodes< KernelT, KernelT::parameters.dimension, P >;
        entry.instances.push_back(inst);
    };
    (add(std::integral_constant< int, orders >{}), ...);
    kernelRegistry().push_back(std::move(entry));
    return static_cast< int >(kernelRegistry().size()) - 1;
}

template < typename KernelT, typename... pqs >
int registerKernel(const char* name, const KernelT& kernel)
{
    constexpr auto params = KernelT::parameters;
    KernelEntry    entry;
    entry.info.name        = name;
    entry.info.dim         = params.dimension;
    entry.info.n_equations = static_cast< int >(params.n_equations);
    entry.info.n_unknowns  = static_cast< int >(params.n_unknowns);
    entry.info.n_fields    = static_cast< int >(params.n_fields);
    entry.info.n_rhs       = static_cast< int >(params.n_rhs);
    entry.info.is_boundary = KernelT::is_boundary;
    entry.object           = std::make_shared< KernelT >(kernel);
    (entry.instances.push_back(makeInstance< KernelT, params.dimension, pqs >()), ...);
    kernelRegistry().push_back(std::move(entry));
    return static_cast< int >(kernelRegistry().size()) - 1;
}
} // namespace l3b

#define L3B_PQ(P, NQ) ::l3b::PQ< P, NQ >
#define L3B_REGISTER_DOMAIN_KERNEL(NAME, FUNCTOR, PARAMS, ...)                                                                  \
    static const int l3b_registered_##NAME =                                                                                   \
        ::l3b::registerKernel< ::l3b::DomainEquationKernel< FUNCTOR, PARAMS >, __VA_ARGS__ >(#NAME, ::l3b::wrapDomainEquationKernel< PARAMS >(FUNCTOR{}))
#define L3B_REGISTER_BOUNDARY_KERNEL(NAME, FUNCTOR, PARAMS, ...)                                                                \
    static const int l3b_registered_##NAME =                                                                                   \
        ::l3b::registerKernel< ::l3b::BoundaryEquationKernel< FUNCTOR, PARAMS >, __VA_ARGS__ >(#NAME, ::l3b::wrapBoundaryEquationKernel< PARAMS >(FUNCTOR{}))
// residual kernels (wrapDomainResidualKernel / wrapBoundaryResidualKernel, common/KernelInterface.hpp:192-204): list the element orders
#define L3B_REGISTER_DOMAIN_RESIDUAL_KERNEL(NAME, FUNCTOR, PARAMS, ...)                                                         \
    static const int l3b_registered_##NAME =                                                                                   \
        ::l3b::registerResidualKernel< ::l3b::ResidualDomainKernel< FUNCTOR, PARAMS >, __VA_ARGS__ >(#NAME, ::l3b::wrapDomainResidualKernel< PARAMS >(FUNCTOR{}))
#define L3B_REGISTER_BOUNDARY_RESIDUAL_KERNEL(NAME, FUNCTOR, PARAMS, ...)                                                       \
    static const int l3b_registered_##NAME =                                                                                   \
        ::l3b::registerResidualKernel< ::l3b::ResidualBoundaryKernel< FUNCTOR, PARAMS >, __VA_ARGS__ >(#NAME, ::l3b::wrapBoundaryResidualKernel< PARAMS >(FUNCTOR{}))
#endif
