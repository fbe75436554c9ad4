// Element-local least-squares assembly on the fp64 tensor cores (DMMA, mma.sync.m8n8k4.f64), fused with the CRS scatter:
//     K_e = sum_q w_q |J_q| B_q^T B_q,   F_e = sum_q w_q |J_q| B_q^T f_q,   B_q[e,(a,u)] = N_a A0(e,u) + sum_s dN_a/dx_s A_s(e,u)
// (algsys/AssembleLocalSystem.hpp:77-280), added straight into the rank-local CRS values / rhs through the precomputed
// node-block slot maps (algsys/ScatterLocalSystem.hpp:25-54 with dofs/DofsFromNodes.hpp:72-87). K_e (2 MB at p=4, U=4) is
// never materialised in HBM.
//
// Decomposition. The local matrix is split by *unknown pair*: the (u, v) block K_uv[a][b] = K_e[(a,u)][(b,v)] is
//     K_uv = sum_q sum_{e in E_u ∩ E_v} (sqrt(w) B[e,(a,u)]) (sqrt(w) B[e,(b,v)]),
// where E_u is the set of equations in which unknown u appears at all — known at compile time from the probed operator
// sparsity (KernelSparsity, kernel_interface.cuh; the device re-checks it on every evaluation). Equations outside
// E_u ∩ E_v contribute exact zeros to the reference's dense rank update (AssembleLocalSystem.hpp:197-206), so leaving
// them out changes nothing but the flop count: 3-D diffusion (E=7, U=4) needs sum |E_u ∩ E_v| = 33 of the dense
// 16 x 7 = 112 products, and only the pairs u <= v are computed (K_e is symmetric, the reference mirrors the lower
// triangle, :176-182). The sqrt(w) split is the reference's own (:189-195); weights are positive wherever |J| > 0.
//
// One CTA owns one (element, unknown pair, TB x TB node tile). Quadrature points are streamed in chunks: the two panels
// sqrt(w) B[.,(a,u)] and sqrt(w) B[.,(b,v)] of chunk c+1 are built in shared memory (per-point mapping + user kernel one
// chunk further ahead) while the warps contract chunk c with DMMA — 32 x 32 accumulator tile per warp, fragments read
// with conflict-free 8-byte loads (panel leading dimension ≡ 4 mod 16 doubles) — one barrier per chunk.
// The DMMA and DFMA instructions share the fp64 pipe on B200 (bench microkernels: 33.7 DFMA, 37.1 DMMA, 34.9 TFLOP/s
// interleaved), so the win over the register-tiled FMA kernel (assemble.cuh) is issue bandwidth and registers, not peak:
// one DMMA retires 256 multiply-adds for 2 fragment loads.
#ifndef L3B_ASSEMBLE_DMMA_CUH
#define L3B_ASSEMBLE_DMMA_CUH

#include "device_common.cuh"

namespace l3b
{
constexpr int asm_max_equations = 16;
constexpr int asm_max_pairs     = max_unknowns * (max_unknowns + 1) / 2;

// unknown pairs with a non-empty equation intersection, and where their tiles start in the per-element tile list
struct AsmPairs
{
    int n_pairs, tiles_per_elem;
    struct Pair
    {
        uint8_t  u, v, n_eq, pad;
        uint16_t tile0, n_tiles;
        uint8_t  eq[asm_max_equations];
    } p[asm_max_pairs];
};

template < typename KernelT >
constexpr bool unknownInEquation(int u, int eq)
{
    using Sp = KernelSparsity< KernelT >;
    for (int op = 0; op <= KernelT::parameters.dimension; ++op)
        if (Sp::nz(op, eq, u))
            return true;
    return false;
}

template < typename KernelT, int DIM, int P >
struct AsmDmmaCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields, NRHS = params.n_rhs;
    static constexpr int  NN = cpow(P + 1, DIM), L = NN * U;
    static constexpr int  TB      = NN > 64 ? 128 : NN > 32 ? 64 : 32; // node tile edge
    static constexpr int  WR      = TB / 32;                            // warps per tile edge
    static constexpr int  threads = WR * WR * 32;
    static constexpr int  n_nb    = (NN + TB - 1) / TB;                 // node blocks
    static constexpr int  LD      = TB + 4;                             // panel leading dimension, ≡ 4 (mod 16)
    static constexpr int  KCMAX   = cmax(32, 4 * E);                    // panel rows (quadrature points x equations) per chunk
    static constexpr int  QCMAX   = 16;                                 // quadrature points per chunk, at most
    static constexpr int  qp_doubles = DIM * DIM + 2 + (DIM + 1) * E * U + E * NRHS; // Jti, sqrt(w), w, A, f
    // smem (doubles): panel A x 2 | panel B x 2 | per-point data x 2 | node field values | vertices
    static constexpr int    off_b = 2 * KCMAX * LD, off_qp = 4 * KCMAX * LD, off_nv = off_qp + 2 * QCMAX * qp_doubles,
                         off_verts = off_nv + NN * NF, total = off_verts + 8 * 3;
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
    static_assert(E <= asm_max_equations and U <= max_unknowns);

    static constexpr AsmPairs makePairs()
    {
        AsmPairs out{};
        int      tile = 0;
        for (int u = 0; u < U; ++u)
            for (int v = u; v < U; ++v)
            {
                AsmPairs::Pair pr{};
                pr.u = static_cast< uint8_t >(u);
                pr.v = static_cast< uint8_t >(v);
                for (int eq = 0; eq < E; ++eq)
                    if (unknownInEquation< KernelT >(u, eq) and unknownInEquation< KernelT >(v, eq))
                        pr.eq[pr.n_eq++] = static_cast< uint8_t >(eq);
                if (pr.n_eq == 0)
                    continue;
                pr.tile0   = static_cast< uint16_t >(tile);
                pr.n_tiles = static_cast< uint16_t >(u == v ? n_nb * (n_nb + 1) / 2 : n_nb * n_nb);
                tile += pr.n_tiles;
                out.p[out.n_pairs++] = pr;
            }
        out.tiles_per_elem = tile;
        return out;
    }
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(AsmDmmaCfg< KernelT, DIM, P >::threads)
    assembleDmmaKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ AsmPairs pairs)
{
    using Cfg = AsmDmmaCfg< KernelT, DIM, P >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int  E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NRHS = Cfg::NRHS, NN = Cfg::NN, TB = Cfg::TB, LD = Cfg::LD;
    constexpr int  T = Cfg::threads, n_warps = T / 32, n_nb = Cfg::n_nb;
    constexpr bool is_bnd = KernelT::is_boundary;
    constexpr int  nv     = 1 << DIM;
    extern __shared__ double smem[];
    double* const s_pa    = smem;
    double* const s_pb    = smem + Cfg::off_b;
    double* const s_qp    = smem + Cfg::off_qp;
    double* const s_nv    = smem + Cfg::off_nv;
    double* const s_verts = smem + Cfg::off_verts;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- which (element, unknown pair, node tile) is this?
    const long long wi = blockIdx.x / pairs.tiles_per_elem;
    int             t  = static_cast< int >(blockIdx.x % pairs.tiles_per_elem);
    int             pi = 0;
    while (pi + 1 < pairs.n_pairs and t >= pairs.p[pi + 1].tile0)
        ++pi;
    const AsmPairs::Pair& pr = pairs.p[pi];
    t -= pr.tile0;
    const int u = pr.u, v = pr.v, n_eq = pr.n_eq;
    int       bi, bj;
    if (u != v)
    {
        bi = t / n_nb;
        bj = t % n_nb;
    }
    else
    {
        bi = 0;
        while (t >= bi + 1)
        {
            t -= bi + 1;
            ++bi;
        }
        bj = t;
    }
    const bool      same_panel = u == v and bi == bj; // the tile is a diagonal block: one panel serves as both operands
    const long long e          = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side       = is_bnd ? args.work_sides[wi] : -1;
    const uint32_t* el_nodes   = args.nodes + e * NN;

    for (int i = tid; i < nv * 3; i += T)
        s_verts[i] = args.verts[e * nv * 3 + i];
    if constexpr (NF > 0)
        for (int i = tid; i < NN * NF; i += T)
            s_nv[i] = args.fields[el_nodes[i / NF] + args.field_inds[i % NF] * args.field_stride];

    const long long tab_off  = is_bnd ? static_cast< long long >(side) * args.n_qp : 0;
    const double*   tab_vals = args.tab_vals + tab_off * NN;
    const double*   tab_ders = args.tab_ders + tab_off * DIM * NN;
    const double*   tab_pts  = args.tab_pts + tab_off * DIM;
    const double*   tab_wts  = args.tab_wts + tab_off;

    // quadrature points per chunk: as many as fit KCMAX panel rows, a multiple of 4 (DMMA k = 4), at most QCMAX
    int QC = (Cfg::KCMAX / n_eq) & ~3;
    QC     = QC > Cfg::QCMAX ? Cfg::QCMAX : QC;
    const int n_chunks = (args.n_qp + QC - 1) / QC;
    const int n_k4     = QC * n_eq / 4;

    // ---- per-point data of chunk c → s_qp[c & 1]: warp w handles point c * QC + w (mapping, fields, user kernel)
    const auto perPoint = [&](int c) {
        double* const base = s_qp + (c & 1) * Cfg::QCMAX * Cfg::qp_doubles;
        for (int qc = warp; qc < QC; qc += n_warps)
        {
            const int q  = c * QC + qc;
            double*   qd = base + qc * Cfg::qp_doubles;
            if (q >= args.n_qp)
            {
                if (lane == 0)
                    qd[DIM * DIM] = 0.; // padding point: sqrt(w) = 0 zeroes its panel rows
                continue;
            }
            double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], nrm[DIM];
            for (int d = 0; d < DIM; ++d)
                xi[d] = tab_pts[q * DIM + d];
            geometryAt< DIM >(s_verts, xi, xs, Jt);
            const double detJ = invert< DIM >(Jt, Jti);
            double       jac  = detJ;
            if constexpr (is_bnd)
                jac = boundaryMeasureAndNormal< DIM >(side, Jt, nrm);
            else if (not(detJ > 0.) and lane == 0)
                atomicOr(args.status, status_degenerate_element); // AssembleLocalSystem.hpp:249
            typename KernelT::Input in;
            if constexpr (NF > 0)
            {
                double        fred[NF * (DIM + 1)];
                const double* bv = tab_vals + static_cast< long long >(q) * NN;
                const double* bd = tab_ders + static_cast< long long >(q) * DIM * NN;
                for (int i = 0; i < NF * (DIM + 1); ++i)
                    fred[i] = 0.;
                for (int a = lane; a < NN; a += 32)
                {
                    double pd[DIM];
                    for (int s = 0; s < DIM; ++s)
                    {
                        double val = 0.;
                        for (int d = 0; d < DIM; ++d)
                            val = fma(Jti[s][d], bd[d * NN + a], val);
                        pd[s] = val;
                    }
                    for (int f = 0; f < NF; ++f)
                    {
                        const double val = s_nv[a * NF + f];
                        fred[f]          = fma(bv[a], val, fred[f]);
                        for (int s = 0; s < DIM; ++s)
                            fred[NF * (s + 1) + f] = fma(pd[s], val, fred[NF * (s + 1) + f]);
                    }
                }
                for (int i = 0; i < NF * (DIM + 1); ++i)
                    for (int off = 16; off > 0; off >>= 1)
                        fred[i] += __shfl_xor_sync(0xffffffffu, fred[i], off);
                for (int f = 0; f < NF; ++f)
                {
                    in.field_vals[f] = fred[f];
                    for (int s = 0; s < DIM; ++s)
                        in.field_ders[s][f] = fred[NF * (s + 1) + f];
                }
            }
            if (lane == 0)
            {
                for (int s = 0; s < 3; ++s)
                    in.point.space.coords[s] = xs[s];
                in.point.time = args.time;
                if constexpr (is_bnd)
                    for (int s = 0; s < DIM; ++s)
                        in.normal[s] = nrm[s];
                const auto res = kernel(in);
                // guard: entries the compile-time probe declared structurally zero must be zero
                bool violated = false;
                staticFor< DIM + 1 >([&](auto op) {
                    staticFor< E >([&](auto eq) {
                        staticFor< U >([&](auto uu) {
                            if constexpr (not Sp::nz(op, eq, uu))
                                violated |= res.operators[op](eq, uu) != 0.;
                        });
                    });
                });
                if (violated)
                    atomicOr(args.status, status_sparsity_violation);
                for (int s = 0; s < DIM; ++s)
                    for (int d = 0; d < DIM; ++d)
                        qd[s * DIM + d] = Jti[s][d];
                const double w    = jac * tab_wts[q];
                qd[DIM * DIM]     = sqrt(fabs(w)); // AssembleLocalSystem.hpp:189-195
                qd[DIM * DIM + 1] = w;
                double* qa        = qd + DIM * DIM + 2;
                for (int i = 0; i <= DIM; ++i)
                    for (int k = 0; k < E * U; ++k)
                        qa[i * E * U + k] = res.operators[i].v[k];
                for (int k = 0; k < E * NRHS; ++k)
                    qa[(DIM + 1) * E * U + k] = res.rhs.v[k];
            }
        }
    };

    // ---- panels of chunk c → s_pa/s_pb[c & 1]: rows k = qc * n_eq + ei, columns = the TB nodes of block bi (resp. bj)
    // work item ↔ (panel column = node of block bi or bj, point of the chunk); a thread always meets the same A column
    const int  n_cols      = same_panel ? TB : 2 * TB;
    const int  my_a_col    = tid % n_cols;          // the panel-A column of this thread, if < TB
    const bool rhs_duty    = u == v and bj == 0;    // this CTA also owns F_e[(a, u)] for the nodes a of block bi
    double     f_acc[NRHS];
    for (int r = 0; r < NRHS; ++r)
        f_acc[r] = 0.;
    const auto build = [&](int c) {
        const double* const qbase = s_qp + (c & 1) * Cfg::QCMAX * Cfg::qp_doubles;
        for (int w = tid; w < n_cols * QC; w += T)
        {
            const int     pcol = w % n_cols, qc = w / n_cols;
            const bool    is_a = pcol < TB;
            const int     prow = pcol % TB;
            const int     pa   = (is_a ? bi : bj) * TB + prow; // local node
            const int     pu   = is_a ? u : v;
            double* const dst  = (is_a ? s_pa : s_pb) + (c & 1) * Cfg::KCMAX * LD + prow;
            const int     q    = c * QC + qc;
            const double* qd   = qbase + qc * Cfg::qp_doubles;
            const double  sw   = qd[DIM * DIM];
            if (q < args.n_qp and pa < NN)
            {
                const double* qa = qd + DIM * DIM + 2;
                const double  wq = qd[DIM * DIM + 1];
                const double  n  = __ldg(tab_vals + static_cast< long long >(q) * NN + pa);
                double        der[DIM], pd[DIM];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    der[d] = __ldg(tab_ders + (static_cast< long long >(q) * DIM + d) * NN + pa);
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                {
                    double val = qd[s * DIM] * der[0];
#pragma unroll
                    for (int d = 1; d < DIM; ++d)
                        val = fma(qd[s * DIM + d], der[d], val);
                    pd[s] = val;
                }
                for (int ei = 0; ei < n_eq; ++ei)
                {
                    const int eq = pr.eq[ei];
                    double    b  = n * qa[eq + pu * E];
#pragma unroll
                    for (int s = 0; s < DIM; ++s)
                        b = fma(pd[s], qa[(s + 1) * E * U + eq + pu * E], b);
                    if (rhs_duty and is_a)
                        for (int r = 0; r < NRHS; ++r)
                            f_acc[r] = fma(b * wq, qa[(DIM + 1) * E * U + eq + r * E], f_acc[r]);
                    dst[(qc * n_eq + ei) * LD] = b * sw;
                }
            }
            else
                for (int ei = 0; ei < n_eq; ++ei)
                    dst[(qc * n_eq + ei) * LD] = 0.;
        }
    };

    // ---- accumulators: warp (wy, wx) owns rows wy*32 .. +32, columns wx*32 .. +32 of the tile as 4 x 4 DMMA tiles
    const int wy = warp / Cfg::WR, wx = warp % Cfg::WR;
    const int g = lane >> 2, tq = lane & 3;
    double    acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j][0] = acc[i][j][1] = 0.;

    __syncthreads(); // vertices, node field values
    perPoint(0);
    __syncthreads();
    build(0);
    if (n_chunks > 1)
        perPoint(1);
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c)
    {
        if (c + 1 < n_chunks)
            build(c + 1);
        if (c + 2 < n_chunks)
            perPoint(c + 2);
        const double* const pa_c = s_pa + (c & 1) * Cfg::KCMAX * LD + wy * 32 + g;
        const double* const pb_c = (same_panel ? s_pa : s_pb) + (c & 1) * Cfg::KCMAX * LD + wx * 32 + g;
#pragma unroll 2
        for (int k4 = 0; k4 < n_k4; ++k4)
        {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                a[i] = pa_c[(k4 * 4 + tq) * LD + i * 8];
                b[i] = pb_c[(k4 * 4 + tq) * LD + i * 8];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    // ---- scatter into the CRS values: slot(row (a,u), col (b,v)) = row_ptr[dof(a,u)] + pos[e][a][b] * dofs_per_node + dof_inds[v]
    // DMMA accumulator layout: acc[i][j][h] = tile(row i*8 + g, column j*8 + 2*tq + h)
    const uint16_t* pos = args.slot_pos + e * static_cast< long long >(NN) * NN;
    const int       du = args.dof_inds[u], dv = args.dof_inds[v];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        const int a = bi * TB + wy * 32 + i * 8 + g;
        if (a >= NN)
            continue;
        const long long node_a = el_nodes[a];
        const long long rbeg   = args.row_ptr[node_a * args.dofs_per_node + du];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const int b = bj * TB + wx * 32 + j * 8 + 2 * tq + h;
                if (b >= NN)
                    continue;
                const double val = acc[i][j][h];
                atomicAdd(args.crs_vals + rbeg + static_cast< long long >(pos[a * NN + b]) * args.dofs_per_node + dv, val);
                if (not same_panel) // mirrored entry K_e[(b,v)][(a,u)]
                {
                    const long long node_b = el_nodes[b];
                    atomicAdd(args.crs_vals + args.row_ptr[node_b * args.dofs_per_node + dv] +
                                  static_cast< long long >(pos[b * NN + a]) * args.dofs_per_node + du,
                              val);
                }
            }
    }
    // ---- rhs (ScatterLocalSystem.hpp:47-52)
    if (rhs_duty and my_a_col < TB and bi * TB + my_a_col < NN and (T >= n_cols or tid < TB))
    {
        const long long grow = static_cast< long long >(el_nodes[bi * TB + my_a_col]) * args.dofs_per_node + du;
        for (int r = 0; r < NRHS; ++r)
            atomicAdd(args.rhs + grow + r * args.ld, f_acc[r]);
    }
}
} // namespace l3b
#endif
