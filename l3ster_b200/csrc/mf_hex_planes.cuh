// Sum-factorised matrix-free apply for hexahedra, "planes + columns" formulation (kernel v5):
//     y[dofs(e)] += alpha * K_e * x[dofs(e)],
// same mathematics and parity contract as mfSumFactApplyKernel (mf_sumfact.cuh), which stays the path for quads and
// for quadratures too large for this kernel's register tiles. Replaces evalLocalOperatorSumFact +
// gatherSumFact/scatterSumFact of the reference (algsys/SumFactorization.hpp:438-917,
// algsys/MatrixFreeSystem.hpp:421-537).
//
// Why another formulation: the line-per-thread kernel moves every tensor through shared memory once per 1-D sweep
// (~136 KB and 12 barriers per p=4 element; ncu r1_v4: shared-memory pipe 46 %, fp64 pipe 25 %, 7 150 warp instructions
// per element). Here every thread owns a whole z-column or a whole xy-plane of one field in REGISTERS, so a sweep costs
// no shared-memory traffic at all; shared memory is only the transpose buffer between the two ownerships:
//
//   A  column (i, j)     gather the nodal z-column of every unknown from x (HBM, vector loads), interpolate along z
//                        in registers, store                                                   →  V[f][qz][j][i]
//   B  plane  (f, qz)    load the nb x nb plane, interpolate along x and y in registers, then differentiate the
//                        interpolated plane along x and y (collocation derivative at the Gauss points)
//                                                                                              →  V, DX, DY
//   C  column (qx, qy)   load values and x/y-derivatives of every field along z; z-derivative in registers;
//                        quadrature-point stage (geometry, user kernel, least-squares operator, fluxes r0, r_xi, r_eta,
//                        r_zeta) at the nq points of the column; transposed z-derivative accumulated in registers
//                                                                                              →  V (= r0 + Dz^T r_zeta), DX, DY
//   D  plane  (u, qz)    load the three planes, transposed x/y-derivatives, project to the nodes along y and x →  V
//   E  column (i, j)     load along z, project to the nodes along z, scatter with fp64 atomics (RED) into y (HBM)
//
// 4 barriers and ~64 KB of shared-memory traffic per p=4 element; 46 k DFMA per element (15 k per transform direction
// pair, 16 k in the point stage) is what remains, i.e. the kernel is built to be bound by the fp64 pipe.
//
// HBM side (v7): the kernel is persistent (one CTA per resident slot, batches of EPB elements strided over the grid) and
// software-pipelined so that no warp waits on HBM inside the loop (ncu r1_v5: 44 % of the stall samples sat on the
// Dirichlet-mask and x loads of the gather/scatter):
//   iteration i   runs A..E on batch i;
//                 after the first barrier it issues, with cp.async (LDGSTS), the geometry record of batch i+1 and the
//                 node ids + element flags of batch i+2 (coalesced 4-byte copies into a 3-deep ring);
//                 at the start of D it issues the indexed gather of batch i+1 — one 256-bit load per node through the
//                 ids that arrived an iteration earlier — into registers that are only consumed by A of iteration i+1
//                 (D and E are the phases with register head-room; v6 staged x through cp.async instead, which cost one
//                 shared-memory wavefront per 16-byte piece: 45 % of all shared-memory traffic, ncu r1_v6).
// Dirichlet mask bytes are only looked at for elements flagged at system creation (elemDirichletFlagKernel).
//
// Shared-memory layout (doubles): three arrays V, DX, DY of [plane = f * NQ + qz][element slot][PSZ], PSZ = NQ^2 rounded
// up to an odd number. A column thread (slot, c) addresses plane * EPB * PSZ + slot * PSZ + c: consecutive threads →
// consecutive words, conflict-free. A plane thread w = plane * EPB + slot addresses PSZ * w + n: odd stride →
// conflict-free.
#ifndef L3B_MF_HEX_PLANES_CUH
#define L3B_MF_HEX_PLANES_CUH

#include "mf_sumfact.cuh"

namespace l3b
{
// per-element geometry record built once at mesh upload (hexGeometryKernel): monomial coefficients of the trilinear map
// [8][3], then — for affine elements — the constant inverse Jacobian Jti[s][d] = dxi_d/dx_s, detJ, and the affine flag
constexpr int hex_geo_doubles = 36;
constexpr int hex_geo_jti = 24, hex_geo_det = 33, hex_geo_affine = 34;

template < typename KernelT, int P, int NQ, int NRHS >
struct MfHexCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields;
    static constexpr int  NB = P + 1;
    static constexpr int  F0 = U * NRHS, F = F0 + NF;
    static constexpr int  NN  = NB * NB * NB;
    static constexpr int  CT  = NQ * NQ;                    // column threads per element
    static constexpr int  PSZ = CT % 2 == 0 ? CT + 1 : CT;  // plane stride
    // elements per batch: fill ~128 threads, but keep the batch's tensors within ~100 KB of shared memory
    static constexpr int  EPB = [] {
#ifdef L3B_HEX_EPB
        if (CT == 25)
            return L3B_HEX_EPB; // experiment override for the nq = 5 instances
#endif
        // nq >= 6: one 256-thread CTA per SM wastes fewer lanes than two of 128 (49 columns: 5 x 49 = 245 of 256 against 98 of 128);
        // measured per 48^3 / 40^3 apply: p=5 1.73 -> 1.60 ms, p=6 2.39 -> 1.95 ms
        if (CT >= 36)
        {
            int wide = 256 / CT;
            while (wide > 1 and 3 * F * NQ * wide * PSZ * 8 > 200 * 1024)
                --wide;
            return wide;
        }
        int epb = cmax(1, 128 / CT);
        while (epb > 1 and 3 * F * NQ * epb * PSZ * 8 > 100 * 1024)
            --epb;
        return epb;
    }();
    static constexpr int  threads = ((EPB * CT + 31) / 32) * 32;
    static constexpr int  NPL     = F * NQ;                 // planes per element, back transform
    static constexpr int  NPLF    = F0 * NQ;                // planes per element, transposed transform
    static constexpr int  AS      = NPL * EPB * PSZ;        // doubles per array
    static constexpr int  RING    = 3;                      // node-id / element-flag prefetch depth
    // shared memory map, in doubles: V, DX, DY | geometry x 2 | then u32: node ids x RING | mask words | flags x RING
    static constexpr int off_geo   = 3 * AS + (3 * AS) % 2; // 16-byte aligned: target of 16-byte cp.async
    static constexpr int off_u32   = off_geo + 2 * EPB * hex_geo_doubles; // doubles
    static constexpr int ids_words = RING * EPB * NN, mask_words = threads * NB, flag_words = RING * EPB + 1;
    static constexpr size_t smem_bytes = static_cast< size_t >(off_u32) * sizeof(double) + (ids_words + mask_words + flag_words) * sizeof(uint32_t);
#ifdef L3B_HEX_MIN_BLOCKS
    static constexpr int  min_blocks  = L3B_HEX_MIN_BLOCKS; // experiment override
#else
    // 2 CTAs/SM: at 3 (168 registers) the point stage spills, and the shared-memory carve-out leaves the spills no L1
    static constexpr int  min_blocks  = smem_bytes <= 112 * 1024 ? 2 : 1;
#endif
    static constexpr bool supported   = CT <= 49 and smem_bytes <= 220 * 1024;
    // Warpgroup slicing (L3B_HEX_WG, nq = 5 instances): ONE CTA per SM made of WG independent 128-thread groups, each running the whole
    // pipeline on its own batches with its own shared-memory slice and named barrier — a "virtual CTA". What it buys over WG resident
    // CTAs: registers move between the groups (setmaxnreg): a group holds reg_base registers per thread in the phases A, B, D, E and
    // reg_points only while it is in the quadrature-point stage C, so three groups fit where two 240-register CTAs did.
#if defined(L3B_HEX_WG)
    static constexpr int  WG = (CT == 25 and threads == 128 and 3 * smem_bytes <= 225 * 1024) ? L3B_HEX_WG : 1;
#else
    static constexpr int  WG = 1;
#endif
#ifndef L3B_HEX_REG_BASE
#define L3B_HEX_REG_BASE 128
#endif
#ifndef L3B_HEX_REG_POINTS
#define L3B_HEX_REG_POINTS 240
#endif
    static constexpr int    reg_base = L3B_HEX_REG_BASE, reg_points = L3B_HEX_REG_POINTS;
    static constexpr size_t wg_stride_doubles = ((smem_bytes + 15) / 16) * 2; // slice of one group, 16-byte aligned
    static constexpr size_t launch_smem    = WG > 1 ? WG * wg_stride_doubles * sizeof(double) : smem_bytes;
    static constexpr int    launch_threads = WG * threads;
    static constexpr int    launch_min_blocks = WG > 1 ? 1 : min_blocks;
    // second copy of the point stage for axis-aligned elements: measured per order (ms per apply with / without, profiles/r1_final_summary.md):
    // p=4 1.50 / 1.63, but p=3 1.52 / 1.44, p=2 1.21 / 1.19, p=5 2.09 / 1.75 (at the 255-register limit the extra code costs more than
    // the saved multiply-adds) — so only the nq = 5 instances carry it
#ifdef L3B_HEX_DIAG_FAST_MAX_NQ
    static constexpr bool diag_fast_path = NQ <= L3B_HEX_DIAG_FAST_MAX_NQ; // experiment override
#else
    static constexpr bool diag_fast_path = NQ == 5;
#endif
    static constexpr bool mask_bits_fit = NB * U <= 64;
    static_assert(params.dimension == 3);
    static_assert(NQ >= NB, "collocation differentiation at the Gauss points needs nq >= nb (value_order >= 1)");
};

// does unknown u's physical derivative along s enter any equation? (compile-time, from the probed operator sparsity)
template < typename KernelT >
constexpr bool gradNeeded(int s, int u)
{
    using Sp = KernelSparsity< KernelT >;
    for (size_t eq = 0; eq < KernelT::parameters.n_equations; ++eq)
        if (Sp::nz(s + 1, eq, u))
            return true;
    return false;
}

__device__ __forceinline__ void cpAsync4(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast< unsigned >(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cpAsync16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast< unsigned >(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
// gather loads of the software pipeline: volatile so that they stay where they are issued (far ahead of their use)
__device__ __forceinline__ void ldNc256(const double* p, double& a, double& b, double& c, double& d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ldNc128(const double* p, double& a, double& b)
{
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ double ldNc64(const double* p)
{
    double a;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(p));
    return a;
}
// predicated fire-and-forget fp64 atomic add: no divergent branch around the RED
__device__ __forceinline__ void redAddIf(double* p, double v, int go)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"(go) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll()
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ENERGY: also accumulate x^T A x (ElemArgs::energy). A separate instantiation: the plain apply keeps its registers (240 at p = 4;
// one more live double costs 6 registers and 3 % of the apply).
template < typename KernelT, int P, int NQ, int NRHS, bool ENERGY = false >
__global__ void __launch_bounds__(MfHexCfg< KernelT, P, NQ, NRHS >::launch_threads, MfHexCfg< KernelT, P, NQ, NRHS >::launch_min_blocks)
    mfHexPlanesKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ SumFactTables< P + 1, NQ > tab)
{
    using Cfg = MfHexCfg< KernelT, P, NQ, NRHS >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NB = Cfg::NB, F0 = Cfg::F0, NN = Cfg::NN;
    constexpr int CT = Cfg::CT, PSZ = Cfg::PSZ, EPB = Cfg::EPB, AS = Cfg::AS, NPL = Cfg::NPL, NPLF = Cfg::NPLF;
    constexpr int T = Cfg::threads, RING = Cfg::RING;
    extern __shared__ double smem_all[];
    constexpr int   WG      = Cfg::WG;
    const int       wg      = WG > 1 ? threadIdx.x / T : 0;                  // warpgroup = virtual CTA
    double* const   smem    = smem_all + wg * (WG > 1 ? Cfg::wg_stride_doubles : 0);
    double* const   s_V     = smem;
    double* const   s_DX    = smem + AS;
    double* const   s_DY    = smem + 2 * AS;
    double* const   s_geo   = smem + Cfg::off_geo;
    uint32_t* const s_ids   = reinterpret_cast< uint32_t* >(smem + Cfg::off_u32);
    uint32_t* const s_mask  = s_ids + Cfg::ids_words;
    uint32_t* const s_flag  = s_mask + Cfg::mask_words;
    // barrier over the group's T threads (named barrier wg + 1 when the CTA holds several groups)
    const auto groupSync = [&] {
        if constexpr (WG > 1)
            asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "n"(T) : "memory");
        else
            __syncthreads();
    };
    if constexpr (WG > 1)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::reg_base));

    const int       tid       = WG > 1 ? threadIdx.x % T : threadIdx.x;
    const long long vblock    = static_cast< long long >(blockIdx.x) * WG + wg, vgrid = static_cast< long long >(gridDim.x) * WG;
    const long long n_batches = (args.n_work + EPB - 1) / EPB;
    const int       n_it      = static_cast< int >((n_batches - vblock + vgrid - 1) / vgrid);
    // column role
    const int  slot    = tid / CT;
    const int  cc      = tid % CT;
    const int  ci      = cc % NQ, cj = cc / NQ;
    const bool col_thr = tid < EPB * CT;
    const bool node_thr = col_thr and ci < NB and cj < NB; // owns the nodal z-column (ci, cj)
    const int  col_off = slot * PSZ + cc;                   // + plane * EPB * PSZ
    // vector gather when the dofs of a node are contiguous and aligned; then (U == 4) one 4-byte mask word per node
    const bool vec_vals   = args.contiguous_dofs != 0 and U % 2 == 0;
    const bool stage_mask = vec_vals and U == 4 and args.dir_mask != nullptr;

    const auto batchOf = [&](int i) { return vblock + static_cast< long long >(i) * vgrid; };
    // element handled by this thread's slot in iteration i, or -1
    const auto elemOf = [&](int i) -> long long {
        const long long wi = batchOf(i) * EPB + slot;
        if (not col_thr or wi >= args.n_work)
            return -1;
        return args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    };
    const auto activeIn = [&](int i) { // number of elements in batch i
        const long long left = args.n_work - batchOf(i) * EPB;
        return static_cast< int >(left < EPB ? left : EPB);
    };
    // node ids (coalesced over the batch) and element flags of iteration i → ring slot i % RING
    const auto prefetchIds = [&](int i) {
        const int       ring  = i % RING;
        const int       n_act = activeIn(i);
        const long long wi0   = batchOf(i) * EPB;
        if (args.work_elems == nullptr)
        {
            const uint32_t* src = args.nodes + (args.first_elem + wi0) * NN;
#pragma unroll 1
            for (int w = tid; w < n_act * NN; w += T)
                cpAsync4(s_ids + ring * EPB * NN + w, src + w);
        }
        else
#pragma unroll 1
            for (int w = tid; w < n_act * NN; w += T)
                cpAsync4(s_ids + ring * EPB * NN + w, args.nodes + static_cast< long long >(args.work_elems[wi0 + w / NN]) * NN + w % NN);
        if (args.elem_dir != nullptr and tid < n_act)
            cpAsync4(s_flag + ring * EPB + tid, args.elem_dir + (args.work_elems ? args.work_elems[wi0 + tid] : args.first_elem + wi0 + tid));
    };
    const auto flagOf = [&](int i) -> bool { // does the element of iteration i touch a Dirichlet dof?
        if (args.dir_mask == nullptr)
            return false;
        return args.elem_dir == nullptr or s_flag[(i % RING) * EPB + slot] != 0;
    };
    const auto nodeOf = [&](int i, int k) -> uint32_t { return s_ids[((i % RING) * EPB + slot) * NN + k * NB * NB + cj * NB + ci]; };
    // geometry record of iteration i
    const auto prefetchGeo = [&](int i) {
        const long long e = elemOf(i);
        if (e < 0)
            return;
        for (int c = cc; c < hex_geo_doubles / 2; c += CT)
            cpAsync16(s_geo + ((i & 1) * EPB + slot) * hex_geo_doubles + c * 2, args.hex_geo + e * hex_geo_doubles + c * 2);
    };
    // the nodal z-column of this thread in iteration i: x → registers, mask words of flagged elements → shared memory
    double     xn[NRHS][NB][U];
    const auto loadX = [&](int i) {
        if (not(node_thr and slot < activeIn(i)))
            return;
        const bool flagged = stage_mask and flagOf(i);
#pragma unroll
        for (int k = 0; k < NB; ++k)
        {
            const long long node = nodeOf(i, k);
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                if (vec_vals)
                {
                    const double* src = args.x + node * U + r * args.ld;
                    if constexpr (U % 4 == 0)
                    {
                        if (args.ld % 4 == 0 and reinterpret_cast< uintptr_t >(args.x) % 32 == 0)
                        {
#pragma unroll
                            for (int u4 = 0; u4 < U / 4; ++u4)
                                ldNc256(src + 4 * u4, xn[r][k][4 * u4], xn[r][k][4 * u4 + 1], xn[r][k][4 * u4 + 2], xn[r][k][4 * u4 + 3]);
                            continue;
                        }
                    }
                    if constexpr (U % 2 == 0)
                    {
#pragma unroll
                        for (int u2 = 0; u2 < U / 2; ++u2)
                            ldNc128(src + 2 * u2, xn[r][k][2 * u2], xn[r][k][2 * u2 + 1]);
                    }
                }
                else
                {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        xn[r][k][u] = ldNc64(args.x + node * args.dofs_per_node + args.dof_inds[u] + r * args.ld);
                }
            }
            if (flagged)
                cpAsync4(s_mask + k * T + tid, args.dir_mask + node * U);
        }
    };

    // ---- prologue (the only exposed HBM latency of the CTA)
    if (n_it <= 0)
        return;
    prefetchIds(0);
    prefetchGeo(0);
    cpAsyncCommit();
    cpAsyncWaitAll();
    groupSync();
#ifndef L3B_HEX_NO_X_PREFETCH
    loadX(0);
#endif
    if (n_it > 1)
        prefetchIds(1);
    cpAsyncCommit();

    constexpr bool want_energy = ENERGY;
    double         energy      = 0.; // this thread's share of x^T A x = sum_q w |B_q x_e|^2 (operand column 0)
    for (int it = 0; it < n_it; ++it)
    {
        const int       n_active = activeIn(it);
        const bool      col_on   = col_thr and slot < n_active;
        const double*   geo      = s_geo + ((it & 1) * EPB + slot) * hex_geo_doubles;
        unsigned long long dirbits = 0; // bit k * U + u: dof (node k of this column, unknown u) is a Dirichlet dof

#ifdef L3B_HEX_NO_X_PREFETCH
        loadX(it);
        cpAsyncCommit();
#endif
        cpAsyncWaitAll();
        // ---- A: gather (MatrixFreeSystem.hpp:421-467) + z-interpolation
        if (col_on and node_thr)
        {
            const bool flagged = flagOf(it);
            if (flagged)
            {
                if (stage_mask)
                {
                    if constexpr (U == 4)
                    {
#pragma unroll
                        for (int k = 0; k < NB; ++k)
                        {
                            const uint32_t m = s_mask[k * T + tid]; // the node's 4 mask bytes, staged by loadX
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if ((m >> (8 * b)) & 0xffu)
                                    dirbits |= 1ull << (k * U + b);
                        }
                    }
                }
                else if constexpr (Cfg::mask_bits_fit)
                {
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        const long long node = nodeOf(it, k);
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (args.dir_mask[node * args.dofs_per_node + args.dof_inds[u]] != 0)
                                dirbits |= 1ull << (k * U + u);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                double v[U][NB];
#pragma unroll
                for (int k = 0; k < NB; ++k)
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        bool dir;
                        if constexpr (Cfg::mask_bits_fit)
                            dir = ((dirbits >> (k * U + u)) & 1ull) != 0;
                        else
                            dir = flagged and args.dir_mask[static_cast< long long >(nodeOf(it, k)) * args.dofs_per_node + args.dof_inds[u]] != 0;
                        v[u][k] = dir ? 0. : xn[r][k][u];
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        double acc = v[u][0] * tab.interp[q];
#pragma unroll
                        for (int k = 1; k < NB; ++k)
                            acc = fma(v[u][k], tab.interp[k * NQ + q], acc);
                        s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off] = acc;
                    }
            }
            if constexpr (NF > 0)
            {
#pragma unroll
                for (int f = 0; f < NF; ++f)
                {
                    double v[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        v[k] = __ldg(args.fields + nodeOf(it, k) + args.field_inds[f] * args.field_stride);
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        double acc = v[0] * tab.interp[q];
#pragma unroll
                        for (int k = 1; k < NB; ++k)
                            acc = fma(v[k], tab.interp[k * NQ + q], acc);
                        s_V[((F0 + f) * NQ + q) * (EPB * PSZ) + col_off] = acc;
                    }
                }
            }
        }
        groupSync();

        // ---- software pipeline: geometry of the next batch, node ids of the one after
        if (it + 1 < n_it)
            prefetchGeo(it + 1);
        if (it + 2 < n_it)
            prefetchIds(it + 2);
        cpAsyncCommit();

        // ---- B: xy-planes: interpolate along x and y, differentiate along x and y
        for (int w = tid; w < NPL * EPB; w += T)
        {
            if (w % EPB >= n_active)
                continue;
            double* const pv = s_V + w * PSZ;
            double        a[NQ][NQ];
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    a[j][i] = pv[j * NQ + i];
#pragma unroll
            for (int j = 0; j < NB; ++j) // x
            {
                double o[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                {
                    double acc = a[j][0] * tab.interp[q];
#pragma unroll
                    for (int i = 1; i < NB; ++i)
                        acc = fma(a[j][i], tab.interp[i * NQ + q], acc);
                    o[q] = acc;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    a[j][q] = o[q];
            }
#pragma unroll
            for (int qx = 0; qx < NQ; ++qx) // y
            {
                double o[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                {
                    double acc = a[0][qx] * tab.interp[q];
#pragma unroll
                    for (int j = 1; j < NB; ++j)
                        acc = fma(a[j][qx], tab.interp[j * NQ + q], acc);
                    o[q] = acc;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    a[q][qx] = o[q];
            }
            double* const px = s_DX + w * PSZ;
            double* const py = s_DY + w * PSZ;
#pragma unroll
            for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
                for (int qx = 0; qx < NQ; ++qx)
                {
                    pv[qy * NQ + qx] = a[qy][qx];
                    double dx = a[qy][0] * tab.colloc[qx], dy = a[0][qx] * tab.colloc[qy];
#pragma unroll
                    for (int m = 1; m < NQ; ++m)
                    {
                        dx = fma(a[qy][m], tab.colloc[m * NQ + qx], dx);
                        dy = fma(a[m][qx], tab.colloc[m * NQ + qy], dy);
                    }
                    px[qy * NQ + qx] = dx;
                    py[qy * NQ + qx] = dy;
                }
        }
        groupSync();

        // ---- C: quadrature-point stage along the z-column (SumFactorization.hpp:614-756)
        if constexpr (WG > 1)
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::reg_points)); // waits until the CTA's pool has them
        if (col_on)
        {
            const bool   affine = geo[hex_geo_affine] != 0.;
            const double wxy    = tab.w[ci] * tab.w[cj];
            double       Jti[3][3], detJ = 0.;
            if (affine)
            {
                detJ = geo[hex_geo_det];
#pragma unroll
                for (int s = 0; s < 3; ++s)
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        Jti[s][d] = geo[hex_geo_jti + s * 3 + d];
            }
            double fval[NF > 0 ? NF : 1][NQ];
            if constexpr (NF > 0)
            {
#pragma unroll
                for (int f = 0; f < NF; ++f)
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        fval[f][q] = s_V[((F0 + f) * NQ + q) * (EPB * PSZ) + col_off];
            }
            bool violated = false;
            // DIAGJ: the element is an axis-aligned box (J^-1 diagonal, flagged in the geometry record): physical gradients and the
            // pull-back of the fluxes are three multiplies per unknown instead of 3 x 3 multiply-adds
            const auto pointStage = [&](auto diag_tag) {
                constexpr bool DIAGJ = decltype(diag_tag)::value;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                double val[U][NQ], wacc[U][NQ];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                    {
                        val[u][q]  = s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off];
                        wacc[u][q] = 0.;
                    }
#pragma unroll
                for (int qz = 0; qz < NQ; ++qz)
                {
                    // geometry + user kernel at (ci, cj, qz)
                    typename KernelT::Input in;
                    {
                        const double xi[3] = {tab.pts[ci], tab.pts[cj], tab.pts[qz]};
                        double       xs[3], Jt[3][3];
                        geometryFromCoefs< 3 >(geo, xi, xs, Jt); // dead code for affine elements whose kernel ignores the point
                        if (not affine)
                            detJ = invert< 3 >(Jt, Jti);
                        in.point.space.coords[0] = xs[0];
                        in.point.space.coords[1] = xs[1];
                        in.point.space.coords[2] = 0.; // SumFactorization.hpp:656, :732
                    }
                    in.point.time = args.time;
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                    {
                        const int    off = ((F0 + f) * NQ + qz) * (EPB * PSZ) + col_off;
                        const double dxf = s_DX[off], dyf = s_DY[off];
                        double       dzf = fval[f][0] * tab.colloc[qz];
#pragma unroll
                        for (int m = 1; m < NQ; ++m)
                            dzf = fma(fval[f][m], tab.colloc[m * NQ + qz], dzf);
                        in.field_vals[f] = fval[f][qz];
#pragma unroll
                        for (int s = 0; s < 3; ++s)
                            in.field_ders[s][f] = DIAGJ ? Jti[s][s] * (s == 0 ? dxf : s == 1 ? dyf : dzf) : fma(Jti[s][2], dzf, fma(Jti[s][1], dyf, Jti[s][0] * dxf));
                    }
                    const auto   res = kernel(in);
                    const double wgt = wxy * tab.w[qz] * detJ;
                    staticFor< 4 >([&](auto op) {
                        staticFor< E >([&](auto eq) {
                            staticFor< U >([&](auto u) {
                                if constexpr (not Sp::nz(op, eq, u))
                                    violated |= res.operators[op](eq, u) != 0.;
                            });
                        });
                    });
                    // reference derivatives of the operand at this point
                    double dref[3][U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        const int off = ((r * U + u) * NQ + qz) * (EPB * PSZ) + col_off;
                        dref[0][u]    = s_DX[off];
                        dref[1][u]    = s_DY[off];
                        double dz     = val[u][0] * tab.colloc[qz];
#pragma unroll
                        for (int m = 1; m < NQ; ++m)
                            dz = fma(val[u][m], tab.colloc[m * NQ + qz], dz);
                        dref[2][u] = dz;
                    }
                    double g_phys[3][U];
                    staticFor< U >([&](auto u) {
                        staticFor< 3 >([&](auto s) {
                            if constexpr (gradNeeded< KernelT >(s, u))
                                g_phys[s][u] = DIAGJ ? Jti[s][s] * dref[s][u] : fma(Jti[s][2], dref[2][u], fma(Jti[s][1], dref[1][u], Jti[s][0] * dref[0][u]));
                        });
                    });
                    double tv[E];
                    staticFor< E >([&](auto eq) {
                        double acc = 0.;
                        staticFor< U >([&](auto u) {
                            if constexpr (Sp::nz(0, eq, u))
                                acc = fma(res.operators[0](eq, u), val[u][qz], acc);
                            staticFor< 3 >([&](auto s) {
                                if constexpr (Sp::nz(s + 1, eq, u))
                                    acc = fma(res.operators[s + 1](eq, u), g_phys[s][u], acc);
                            });
                        });
                        tv[eq] = acc * wgt;
                        if constexpr (want_energy)
                            if (r == 0)
                                energy = fma(acc, tv[eq], energy);
                    });
                    staticFor< U >([&](auto u) {
                        double a0 = 0., ps[3] = {0., 0., 0.};
                        staticFor< E >([&](auto eq) {
                            if constexpr (Sp::nz(0, eq, u))
                                a0 = fma(res.operators[0](eq, u), tv[eq], a0);
                            staticFor< 3 >([&](auto s) {
                                if constexpr (Sp::nz(s + 1, eq, u))
                                    ps[s] = fma(res.operators[s + 1](eq, u), tv[eq], ps[s]);
                            });
                        });
                        // r_d = sum_s Ji(d, s) p_s, only over the directions s this unknown is differentiated along
                        double rd[3] = {0., 0., 0.};
                        staticFor< 3 >([&](auto s) {
                            if constexpr (gradNeeded< KernelT >(s, u))
                            {
                                if constexpr (DIAGJ)
                                    rd[s] = Jti[s][s] * ps[s];
                                else
                                {
#pragma unroll
                                    for (int d = 0; d < 3; ++d)
                                        rd[d] = fma(Jti[s][d], ps[s], rd[d]);
                                }
                            }
                        });
                        const int off = ((r * U + u) * NQ + qz) * (EPB * PSZ) + col_off;
                        s_DX[off]     = rd[0];
                        s_DY[off]     = rd[1];
                        wacc[u][qz] += a0;
#pragma unroll
                        for (int m = 0; m < NQ; ++m)
                            wacc[u][m] = fma(tab.colloc[m * NQ + qz], rd[2], wacc[u][m]);
                    });
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off] = wacc[u][q];
            }
            };
            if constexpr (Cfg::diag_fast_path)
            {
                if (geo[hex_geo_affine] == 2.)
                    pointStage(std::true_type{});
                else
                    pointStage(std::false_type{});
            }
            else
                pointStage(std::false_type{});
            if (violated)
                atomicOr(args.status, status_sparsity_violation);
        }
        if constexpr (WG > 1)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::reg_base));
        groupSync();

        // ---- software pipeline: gather of the next batch into registers (consumed by A of the next iteration)
#ifndef L3B_HEX_NO_X_PREFETCH
        if (it + 1 < n_it)
            loadX(it + 1);
        cpAsyncCommit();
#endif

        // ---- D: xy-planes: transposed derivatives, projection to the nodes along y and x
        for (int w = tid; w < NPLF * EPB; w += T)
        {
            if (w % EPB >= n_active)
                continue;
            double* const       pv = s_V + w * PSZ;
            const double* const px = s_DX + w * PSZ;
            const double* const py = s_DY + w * PSZ;
            double              t[NQ][NQ];
#pragma unroll
            for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
                for (int qx = 0; qx < NQ; ++qx)
                    t[qy][qx] = pv[qy * NQ + qx];
#pragma unroll
            for (int qy = 0; qy < NQ; ++qy)
#pragma unroll
                for (int qx = 0; qx < NQ; ++qx)
                {
                    const double rx = px[qy * NQ + qx], ry = py[qy * NQ + qx];
#pragma unroll
                    for (int m = 0; m < NQ; ++m)
                    {
                        t[qy][m] = fma(tab.colloc[m * NQ + qx], rx, t[qy][m]);
                        t[m][qx] = fma(tab.colloc[m * NQ + qy], ry, t[m][qx]);
                    }
                }
#pragma unroll
            for (int qx = 0; qx < NQ; ++qx) // y: Gauss points → nodes
            {
                double o[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j)
                {
                    double acc = t[0][qx] * tab.interp[j * NQ];
#pragma unroll
                    for (int q = 1; q < NQ; ++q)
                        acc = fma(t[q][qx], tab.interp[j * NQ + q], acc);
                    o[j] = acc;
                }
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    t[j][qx] = o[j];
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) // x
#pragma unroll
                for (int i = 0; i < NB; ++i)
                {
                    double acc = t[j][0] * tab.interp[i * NQ];
#pragma unroll
                    for (int q = 1; q < NQ; ++q)
                        acc = fma(t[j][q], tab.interp[i * NQ + q], acc);
                    pv[j * NQ + i] = acc;
                }
        }
        groupSync();

#ifdef L3B_HEX_SCATTER_TASKS
        // ---- E: z-projection + scatter (MatrixFreeSystem.hpp:494-537): relaxed fp64 atomics (RED), Dirichlet rows skipped.
        // One task per (nodal column (i, j), unknown u), unknown fastest: the U lanes of a node hit one 32-byte sector, so a
        // warp's RED touches 32 / U sectors instead of 32 (ncu r1_v6: the scatter held 19 % of the stall samples).
        {
            const int n_tasks = n_active * NB * NB * U;
#pragma unroll 1
            for (int w = tid; w < n_tasks; w += T)
            {
                const int  u = w % U, ij = (w / U) % (NB * NB), sl = w / (U * NB * NB);
                const int  i = ij % NB, j = ij / NB;
                const bool flagged = args.dir_mask != nullptr and (args.elem_dir == nullptr or s_flag[(it % RING) * EPB + sl] != 0);
                const uint32_t* ids = s_ids + ((it % RING) * EPB + sl) * NN + j * NB + i;
                const long long dof_stride = vec_vals ? U : args.dofs_per_node;
                const int       dof_off    = vec_vals ? u : args.dof_inds[u];
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                {
                    double v[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        v[q] = s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + sl * PSZ + j * NQ + i];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        double acc = v[0] * tab.interp[k * NQ];
#pragma unroll
                        for (int q = 1; q < NQ; ++q)
                            acc = fma(v[q], tab.interp[k * NQ + q], acc);
                        const long long dof = ids[k * NB * NB] * dof_stride + dof_off;
                        const int       go  = not(flagged and args.dir_mask[dof] != 0);
                        redAddIf(args.y + dof + r * args.ld, acc * args.alpha, go);
                    }
                }
            }
        }
#else
        // ---- E: z-projection + scatter (MatrixFreeSystem.hpp:494-537): relaxed fp64 atomics (RED), Dirichlet rows skipped
#ifndef L3B_HEX_NO_RED_TRANSPOSE
        // U == 4, contiguous dofs: the four dofs of a node are one 32-byte sector. A thread owns the 4 x NB results of its nodal column;
        // issued as they are, every lane of a RED hits its own sector (32 wavefronts in the L1 per instruction). A 4 x 4 transpose over
        // each group of four lanes (two shuffle stages) gives lane g the unknown g of the four columns of its group, so the four lanes
        // of a group add into ONE sector: 8 wavefronts per RED instruction instead of 32.
        if (U == 4 and vec_vals)
        {
            const int  g       = tid & 3;
            const bool me_ok   = col_on and node_thr;
            const bool odd1    = (g & 1) != 0, odd2 = (g & 2) != 0;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                double out[4][NB];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    double v[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        v[q] = me_ok ? s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off] : 0.;
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        double acc = v[0] * tab.interp[k * NQ];
#pragma unroll
                        for (int q = 1; q < NQ; ++q)
                            acc = fma(v[q], tab.interp[k * NQ + q], acc);
                        out[u][k] = acc * args.alpha;
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; ++k)
                {
                    // stage 1: lanes g ^ 1 swap the (0,1) and (2,3) index pairs; stage 2: lanes g ^ 2 swap (0,2) and (1,3)
                    {
                        const double s0 = odd1 ? out[0][k] : out[1][k], s1 = odd1 ? out[2][k] : out[3][k];
                        const double r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                        if (odd1)
                            out[0][k] = r0, out[2][k] = r1;
                        else
                            out[1][k] = r0, out[3][k] = r1;
                    }
                    {
                        const double s0 = odd2 ? out[0][k] : out[2][k], s1 = odd2 ? out[1][k] : out[3][k];
                        const double r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
                        if (odd2)
                            out[0][k] = r0, out[1][k] = r1;
                        else
                            out[2][k] = r0, out[3][k] = r1;
                    }
                }
                // out[m][k] now holds unknown g of the column owned by lane (group base + m)
#pragma unroll
                for (int m = 0; m < 4; ++m)
                {
                    const int  tm    = (tid & ~3) + m;
                    const int  slotm = tm / CT, ccm = tm % CT;
                    const int  cim = ccm % NQ, cjm = ccm / NQ;
                    const bool okm = tm < EPB * CT and slotm < n_active and cim < NB and cjm < NB;
                    const unsigned long long dbm = __shfl_sync(0xffffffffu, dirbits, m, 4);
                    const uint32_t* ids = s_ids + ((it % RING) * EPB + (okm ? slotm : 0)) * NN + (okm ? cjm * NB + cim : 0);
                    uint32_t        nd[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        nd[k] = ids[k * NB * NB];
                    double* const yb = args.y + g + r * args.ld;
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        const int go = okm and not((dbm >> (k * U + g)) & 1ull);
#if defined(L3B_HEX_PROBE_PLAIN_STORES) // timing probe only (wrong results): what the apply costs without the atomic units
                        if (go)
                            yb[static_cast< long long >(nd[k]) * U] = out[m][k];
#else
                        redAddIf(yb + static_cast< long long >(nd[k]) * U, out[m][k], go); // predicated RED: no branch around it
#endif
                    }
                }
            }
        }
        else
#endif
        if (col_on and node_thr)
        {
            const bool flagged = flagOf(it);
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
            {
                double out[U][NB];
#pragma unroll
                for (int u = 0; u < U; ++u)
                {
                    double v[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        v[q] = s_V[((r * U + u) * NQ + q) * (EPB * PSZ) + col_off];
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                    {
                        double acc = v[0] * tab.interp[k * NQ];
#pragma unroll
                        for (int q = 1; q < NQ; ++q)
                            acc = fma(v[q], tab.interp[k * NQ + q], acc);
                        out[u][k] = acc * args.alpha;
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; ++k)
                {
                    const long long node = nodeOf(it, k);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                    {
                        const long long dof = vec_vals ? node * U + u : node * args.dofs_per_node + args.dof_inds[u];
                        bool dir; // dirbits: set in A from the staged mask words, no HBM access here
                        if constexpr (Cfg::mask_bits_fit)
                            dir = ((dirbits >> (k * U + u)) & 1ull) != 0;
                        else
                            dir = flagged and args.dir_mask[dof] != 0;
                        if (not dir)
                            atomicAdd(args.y + dof + r * args.ld, out[u][k]);
                    }
                }
            }
        }
#endif
    }
    if constexpr (want_energy)
    {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            energy += __shfl_xor_sync(0xffffffffu, energy, off);
        if ((tid & 31) == 0)
            atomicAdd(args.energy, energy);
    }
}
} // namespace l3b
#endif
