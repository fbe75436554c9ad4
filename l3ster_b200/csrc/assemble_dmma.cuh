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
// One CTA owns one (element, unknown pair, TM x TN node tile), 128 x 128 nodes and 16 warps for NN > 64. Quadrature
// points are streamed in chunks: the two panels
// sqrt(w) B[.,(a,u)] and sqrt(w) B[.,(b,v)] of chunk c+1 are built in shared memory while the warps contract chunk c with
// DMMA — 32 x 32 accumulator tile per warp, fragments read with conflict-free 8-byte loads (panel leading dimension
// ≡ 4 mod 16 doubles) — one barrier per chunk. The per-point stage (mapping, user kernel, two chunks ahead) folds
// J^-1 and sqrt(w) into four coefficients per (point, equation, panel),
//     sqrt(w) B[e,(a,u)] = c0 N_a + sum_d c_d dN_a/dxi_d,   c0 = sqrt(w) A0(e,u),  c_d = sqrt(w) sum_s A_s(e,u) Jinv(s,d),
// so a panel entry costs 4 fp64 operations on the (L2-resident) reference tables. For u == v the warp tiles above the
// diagonal are skipped and the lower triangle is mirrored by the scatter.
// The DMMA and DFMA instructions share the fp64 pipe on B200 (bench microkernels: 33.7 DFMA, 37.1 DMMA, 34.9 TFLOP/s
// interleaved), so the win over the register-tiled FMA kernel (assemble.cuh) is issue bandwidth and registers, not peak:
// one DMMA retires 256 multiply-adds for 2 fragment loads.
#ifndef L3B_ASSEMBLE_DMMA_CUH
#define L3B_ASSEMBLE_DMMA_CUH

#include "device_common.cuh"

namespace l3b
{
#ifndef L3B_ASM_KCMAX
#define L3B_ASM_KCMAX 32
#endif
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
        int8_t   ei_of_eq[asm_max_equations]; // position of an equation in `eq`, or -1
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
    static constexpr int  TN      = NN > 64 ? 128 : NN > 32 ? 64 : 32; // tile columns (nodes b)
#ifdef L3B_ASM_RECT
    static constexpr int  TM      = NN > 64 ? 64 : TN; // experiment: 64 x 128 tiles, two 256-thread CTAs per SM (129 k vs 158 k elements/s)
#else
    static constexpr int  TM      = TN;                // tile rows (nodes a); TN is a multiple of TM
#endif
    static constexpr int  WM = TM / 32, WN = TN / 32;                  // warps along rows / columns, 32 x 32 each
    static constexpr int  threads = WM * WN * 32;
    static constexpr int  n_rb = (NN + TM - 1) / TM, n_cb = (NN + TN - 1) / TN; // row / column blocks
    static constexpr int  LDA = TM + 4, LDB = TN + 4;                  // panel leading dimensions, ≡ 4 (mod 16)
    static constexpr int  KCMAX = cmax(L3B_ASM_KCMAX, 4 * E);                     // panel rows (quadrature points x equations) per chunk
    static constexpr int  QCMAX = 32;                                  // quadrature points per chunk, at most (fewer chunk barriers: 61 k -> 52 k cycles per n_eq = 1 tile)
    // smem (doubles): panel A x 2 | panel B x 2 | coefficients A | coefficients B | rhs coefficients | node field values | vertices
    // the per-point stage runs for a whole super-chunk of points at once, all threads, one point each: SROWS coefficient rows
    // (points x equations of the pair) per panel — every point of a p <= 4 hexahedron in one go
    static constexpr int  SROWS = cmax(512, 16 * E), SPTS = 128;
    static constexpr int NQ1MAX = 16; // 1-D quadrature points the shared-memory copy of the 1-D tables has room for
    static constexpr int off_b = 2 * KCMAX * LDA, off_ca = off_b + 2 * KCMAX * LDB, off_cb = off_ca + SROWS * 4,
                         off_cr = off_cb + SROWS * 4, off_nv = off_cr + SPTS * NRHS * 4, off_verts = off_nv + NN * NF,
                         off_t1d = off_verts + 8 * 3, off_qidx = off_t1d + 2 * (P + 1) * NQ1MAX, total = off_qidx + SPTS / 2;
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
    static constexpr int    min_blocks = smem_bytes <= 112 * 1024 and threads <= 256 ? 2 : 1;
    static_assert(E <= asm_max_equations and U <= max_unknowns);

    // tiles of the pair (u, v): all n_rb x n_cb for u != v; for u == v those that touch the lower triangle
    static constexpr bool tileNeeded(bool diag_pair, int rb, int cb) { return not diag_pair or cb * TN <= rb * TM + TM - 1; }
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
                for (int eq = 0; eq < asm_max_equations; ++eq)
                    pr.ei_of_eq[eq] = -1;
                for (int eq = 0; eq < E; ++eq)
                    if (unknownInEquation< KernelT >(u, eq) and unknownInEquation< KernelT >(v, eq))
                    {
                        pr.ei_of_eq[eq]  = static_cast< int8_t >(pr.n_eq);
                        pr.eq[pr.n_eq++] = static_cast< uint8_t >(eq);
                    }
                if (pr.n_eq == 0)
                    continue;
                int n = 0;
                for (int rb = 0; rb < n_rb; ++rb)
                    for (int cb = 0; cb < n_cb; ++cb)
                        n += tileNeeded(u == v, rb, cb);
                pr.tile0   = static_cast< uint16_t >(tile);
                pr.n_tiles = static_cast< uint16_t >(n);
                tile += n;
                out.p[out.n_pairs++] = pr;
            }
        out.tiles_per_elem = tile;
        return out;
    }
};

#ifdef L3B_ASM_TIMING
// experiment: per-warp phase clocks inside the chunk loop [cta][warp][build, mma, barrier, n_eq]
static __device__ long long g_asm_timing[4096][16][4];
#endif
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(AsmDmmaCfg< KernelT, DIM, P >::threads, AsmDmmaCfg< KernelT, DIM, P >::min_blocks)
    assembleDmmaKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ AsmPairs pairs)
{
    using Cfg = AsmDmmaCfg< KernelT, DIM, P >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int  E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NRHS = Cfg::NRHS, NN = Cfg::NN, TM = Cfg::TM, TN = Cfg::TN;
    constexpr int  LDA = Cfg::LDA, LDB = Cfg::LDB, KCMAX = Cfg::KCMAX, QCMAX = Cfg::QCMAX;
    constexpr int  T = Cfg::threads, n_warps = T / 32;
    constexpr bool is_bnd = KernelT::is_boundary;
    constexpr int  nv     = 1 << DIM;
    extern __shared__ double smem[];
    double* const s_pa    = smem;
    double* const s_pb    = smem + Cfg::off_b;
    double* const s_ca    = smem + Cfg::off_ca; // [buf][k][4]
    double* const s_cb    = smem + Cfg::off_cb;
    double* const s_cr    = smem + Cfg::off_cr; // [buf][qc][r][4]
    double* const s_nv    = smem + Cfg::off_nv;
    double* const s_verts = smem + Cfg::off_verts;
    double* const s_t1d   = smem + Cfg::off_t1d;                                      // [interp | der][b][q]
    uint32_t* const s_qidx = reinterpret_cast< uint32_t* >(smem + Cfg::off_qidx);      // per point of the super-chunk: qx | qy << 8 | qz << 16

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- which (element, unknown pair, node tile) is this?
    const long long wi = blockIdx.x / pairs.tiles_per_elem;
    int             t  = static_cast< int >(blockIdx.x % pairs.tiles_per_elem);
    int             pi = 0;
    while (pi + 1 < pairs.n_pairs and t >= pairs.p[pi + 1].tile0)
        ++pi;
    const AsmPairs::Pair& pr = pairs.p[pi];
    t -= pr.tile0;
    const int  u = pr.u, v = pr.v, n_eq = pr.n_eq;
    const bool diag_pair = u == v;
    int        rb = 0, cb = 0;
    for (int i = 0, n = 0; i < Cfg::n_rb * Cfg::n_cb; ++i) // t-th needed tile of the pair
        if (Cfg::tileNeeded(diag_pair, i / Cfg::n_cb, i % Cfg::n_cb) and n++ == t)
        {
            rb = i / Cfg::n_cb;
            cb = i % Cfg::n_cb;
            break;
        }
    const int row0 = rb * TM, col0 = cb * TN;
    // u == v and the row block lies inside the column block: panel A is a window of panel B
    const bool      a_in_b   = diag_pair and row0 >= col0 and row0 + TM <= col0 + TN;
    const long long e        = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side     = is_bnd ? args.work_sides[wi] : -1;
    const uint32_t* el_nodes = args.nodes + e * NN;

    for (int i = tid; i < nv * 3; i += T)
        s_verts[i] = args.verts[e * nv * 3 + i];
    // tensor-product evaluation of the basis in the panel build (no table traffic in the chunk loop) when the 1-D tables are given
    const int  nq1    = args.nq1d;
    const bool tensor = not is_bnd and args.tab1d != nullptr and nq1 <= Cfg::NQ1MAX;
    if (tensor)
        for (int i = tid; i < 2 * (P + 1) * nq1; i += T)
            s_t1d[i] = args.tab1d[i];
    if constexpr (NF > 0)
        for (int i = tid; i < NN * NF; i += T)
            s_nv[i] = args.fields[el_nodes[i / NF] + args.field_inds[i % NF] * args.field_stride];

    const long long tab_off  = is_bnd ? static_cast< long long >(side) * args.n_qp : 0;
    const double*   tab_vals = args.tab_vals + tab_off * NN;
    const double*   tab_ders = args.tab_ders + tab_off * DIM * NN;
    const double*   tab_pts  = args.tab_pts + tab_off * DIM;
    const double*   tab_wts  = args.tab_wts + tab_off;

    // quadrature points per chunk: as many as fit KCMAX panel rows, a multiple of 4 (DMMA k = 4), at most QCMAX
    int QC = (KCMAX / n_eq) & ~3;
    QC     = QC > QCMAX ? QCMAX : QC;
    const int  n_chunks = (args.n_qp + QC - 1) / QC;
    // chunks per super-chunk of the per-point stage
    int cps = Cfg::SROWS / (QC * n_eq);
    cps     = cps * QC > Cfg::SPTS ? Cfg::SPTS / QC : cps;
    const int  n_k4     = QC * n_eq / 4;
    const bool rhs_duty = diag_pair and cb == 0; // this CTA also owns F_e[(a, u)] for the nodes a of row block rb

    // ---- per-point stage of super-chunk sc → coefficient tables. All threads, one point each: mapping, user kernel, then —
    // through compile-time loops over the structurally non-zero operator entries only —
    //     c0 = sqrt(w) A0(e,u),  c_d = sqrt(w) sum_s A_s(e,u) Jinv(s,d)        for the equations e of the pair and u, v.
    const auto perPoint = [&](int sc) {
        double* const ca  = s_ca;
        double* const cbp = s_cb;
        double* const cr  = s_cr;
        for (int qc = tid; qc < cps * QC; qc += T) // qc: point within the super-chunk = row block qc * n_eq of the tables
        {
            const int q = sc * cps * QC + qc;
            if (tensor)
            {
                const int qq = min(q, args.n_qp - 1);
                s_qidx[qc]   = static_cast< uint32_t >(qq % nq1) | static_cast< uint32_t >((qq / nq1) % nq1) << 8 |
                             static_cast< uint32_t >(qq / (nq1 * nq1)) << 16;
            }
            if (q >= args.n_qp) // padding point: zero coefficients zero its panel rows
            {
                for (int i = 0; i < n_eq * 4; ++i)
                    ca[qc * n_eq * 4 + i] = cbp[qc * n_eq * 4 + i] = 0.;
                for (int i = 0; i < NRHS * 4; ++i)
                    cr[qc * NRHS * 4 + i] = 0.;
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
            else if (not(detJ > 0.))
                atomicOr(args.status, status_degenerate_element); // AssembleLocalSystem.hpp:249
            typename KernelT::Input in;
            if constexpr (NF > 0)
            {
                // reference-space sums over the nodes first, then one J^-1 per field (AssembleLocalSystem.hpp:54-75)
                double        sv[NF], sd[DIM][NF];
                const double* bv = tab_vals + static_cast< long long >(q) * NN;
                const double* bd = tab_ders + static_cast< long long >(q) * DIM * NN;
                for (int f = 0; f < NF; ++f)
                {
                    sv[f] = 0.;
                    for (int d = 0; d < DIM; ++d)
                        sd[d][f] = 0.;
                }
                for (int a = 0; a < NN; ++a)
                {
                    const double n = bv[a];
                    double       dr[DIM];
                    for (int d = 0; d < DIM; ++d)
                        dr[d] = bd[d * NN + a];
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                    {
                        const double val = s_nv[a * NF + f];
                        sv[f]            = fma(n, val, sv[f]);
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            sd[d][f] = fma(dr[d], val, sd[d][f]);
                    }
                }
#pragma unroll
                for (int f = 0; f < NF; ++f)
                {
                    in.field_vals[f] = sv[f];
#pragma unroll
                    for (int s_ = 0; s_ < DIM; ++s_)
                    {
                        double acc = 0.;
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            acc = fma(Jti[s_][d], sd[d][f], acc);
                        in.field_ders[s_][f] = acc;
                    }
                }
            }
            for (int s_ = 0; s_ < 3; ++s_)
                in.point.space.coords[s_] = xs[s_];
            in.point.time = args.time;
            if constexpr (is_bnd)
                for (int s_ = 0; s_ < DIM; ++s_)
                    in.normal[s_] = nrm[s_];
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
            const double w  = jac * tab_wts[q];
            const double sw = sqrt(fabs(w)); // AssembleLocalSystem.hpp:189-195
            double       racc[NRHS][4];
            for (int r = 0; r < NRHS; ++r)
                for (int j = 0; j < 4; ++j)
                    racc[r][j] = 0.;
            staticFor< U >([&](auto uu) {
                if (uu != u and uu != v)
                    return;
                staticFor< E >([&](auto eq) {
                    if constexpr (unknownInEquation< KernelT >(uu, eq))
                    {
                        const int ei = pr.ei_of_eq[eq];
                        if (ei < 0)
                            return;
                        double cf[4] = {0., 0., 0., 0.};
                        if constexpr (Sp::nz(0, eq, uu))
                            cf[0] = sw * res.operators[0](eq, uu);
                        staticFor< DIM >([&](auto s_) {
                            if constexpr (Sp::nz(s_ + 1, eq, uu))
                            {
#pragma unroll
                                for (int d = 0; d < DIM; ++d)
                                    cf[1 + d] = fma(sw * res.operators[s_ + 1](eq, uu), Jti[s_][d], cf[1 + d]);
                            }
                        });
                        if (uu == u)
                        {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                ca[(qc * n_eq + ei) * 4 + j] = cf[j];
                            // rhs: F_e[(a,u)] += sum_e w B[e,(a,u)] f_e = sum_e (sqrt(w) f_e) (sqrt(w) B[e,(a,u)])
#pragma unroll
                            for (int r = 0; r < NRHS; ++r)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    racc[r][j] = fma(sw * res.rhs(eq, r), cf[j], racc[r][j]);
                        }
                        if (uu == v)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                cbp[(qc * n_eq + ei) * 4 + j] = cf[j];
                    }
                });
            });
            if (rhs_duty)
                for (int r = 0; r < NRHS; ++r)
                    for (int j = 0; j < 4; ++j)
                        cr[(qc * NRHS + r) * 4 + j] = racc[r][j];
        }
    };

    // ---- panels of chunk c → s_pa/s_pb[c & 1]: rows k = qc * n_eq + ei, columns = the nodes of the row / column block.
    // A thread owns one panel column (several when T < n_pcols) and a share `part` of the chunk's points, so at most one
    // A-side node — the one whose rhs entry it accumulates. Everything that does not depend on the chunk is hoisted here; the
    // item loop is branch-free: columns of padding nodes are zeroed once and never written, padding points read the tables of
    // the last real point against zero coefficients.
    // u != v over the same node block: both panels are functions of the same table values — one item builds both (`dual`)
    const bool dual    = not a_in_b and row0 == col0 and TM == TN;
    const int  n_pcols = a_in_b or dual ? TN : TM + TN;
    const int  n_parts = T >= n_pcols ? T / n_pcols : 1;
    double     f_acc[NRHS];
    for (int r = 0; r < NRHS; ++r)
        f_acc[r] = 0.;
    for (int i = tid; i < 2 * KCMAX * LDA; i += T)
        s_pa[i] = 0.;
    for (int i = tid; i < 2 * KCMAX * LDB; i += T)
        s_pb[i] = 0.;
    const auto buildColumn = [&]< int NEQ >(std::integral_constant< int, NEQ >, int c, int pcol, int part) {
        const bool is_b = pcol < TN; // panel B columns first (they exist in both layouts)
        const int  prow = is_b ? pcol : pcol - TN;
        const int  node = (is_b ? col0 : row0) + prow;
        if (node >= NN)
            return;
        const int     ld    = is_b ? LDB : LDA;
        const int     neq   = NEQ > 0 ? NEQ : n_eq;
        const int     c_in  = c % cps; // chunk within its super-chunk
        double*       dst   = (is_b ? s_pb + (c & 1) * KCMAX * LDB : s_pa + (c & 1) * KCMAX * LDA) + prow + part * neq * ld;
        const double* cf    = (is_b ? s_cb : s_ca) + (c_in * QC + part) * neq * 4;
        double*       dst2  = s_pa + (c & 1) * KCMAX * LDA + prow + part * neq * LDA; // dual: the A panel of the same node
        const double* cf2   = s_ca + (c_in * QC + part) * neq * 4;
        const double* cr    = s_cr + (c_in * QC + part) * NRHS * 4;
        // rhs rows: the A-side nodes (row block); with a_in_b they are the window [row0, row0 + TM) of panel B
        const bool    rhs_row = rhs_duty and (a_in_b ? (node >= row0 and node < row0 + TM) : not is_b);
        const int     q_last  = args.n_qp - 1;
        // the basis of `node` at a point: products of the 1-D tables in shared memory, or the dense tables (L2)
        constexpr int NB1     = P + 1;
        const double* const tI = s_t1d, * const tD = s_t1d + NB1 * nq1;
        const int     ix = node % NB1, iy = (node / NB1) % NB1, iz = DIM == 3 ? node / (NB1 * NB1) : 0;
        const auto    basisAt = [&](int qc_in_chunk, double (&bas)[4]) {
            if (tensor)
            {
                const uint32_t qi = s_qidx[c_in * QC + qc_in_chunk];
                const int      qx = qi & 0xff, qy = (qi >> 8) & 0xff, qz = qi >> 16;
                const double   bx = tI[ix * nq1 + qx], dx = tD[ix * nq1 + qx], by = tI[iy * nq1 + qy], dy = tD[iy * nq1 + qy];
                if constexpr (DIM == 2)
                {
                    bas[0] = bx * by;
                    bas[1] = dx * by;
                    bas[2] = bx * dy;
                    bas[3] = 0.;
                }
                else
                {
                    const double bz = tI[iz * nq1 + qz], dz = tD[iz * nq1 + qz];
                    const double xy = bx * by, yz = by * bz;
                    bas[0] = xy * bz;
                    bas[1] = dx * yz;
                    bas[2] = bx * dy * bz;
                    bas[3] = xy * dz;
                }
            }
            else
            {
                const int q = min(c * QC + qc_in_chunk, q_last);
                bas[0]      = __ldg(tab_vals + q * NN + node);
                bas[3]      = 0.;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    bas[1 + d] = __ldg(tab_ders + (q * DIM + d) * NN + node);
            }
        };
        // one panel entry: c . (N, dN/dxi) as two independent half sums — the fp64 pipe's dependent-issue latency, not its
        // throughput, is what a chain of four DFMAs would pay here
        const auto entry = [](const double* cf4, const double (&bas)[4]) {
            const double2 c01 = *reinterpret_cast< const double2* >(cf4);
            const double2 c23 = *reinterpret_cast< const double2* >(cf4 + 2);
            double        t0  = c01.x * bas[0];
            t0                = fma(c01.y, bas[1], t0);
            if constexpr (DIM == 2)
                return fma(c23.x, bas[2], t0);
            else
            {
                double t1 = c23.x * bas[2];
                t1        = fma(c23.y, bas[3], t1);
                return t0 + t1;
            }
        };
        if constexpr (NEQ > 0)
        {
            // two points per trip, every entry of both an independent chain (explicit ILP: the compiler keeps the loop rolled)
            constexpr int G = 2;
            for (int qc = part; qc < QC; qc += G * n_parts)
            {
                double bas[G][4];
                bool   on[G];
#pragma unroll
                for (int gi = 0; gi < G; ++gi)
                {
                    on[gi] = qc + gi * n_parts < QC;
                    basisAt(on[gi] ? qc + gi * n_parts : qc, bas[gi]);
                }
#pragma unroll
                for (int ei = 0; ei < NEQ; ++ei)
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
                    {
                        // (a padding trip writes rows of the next chunk's share that this same thread rewrites later)
                        if (on[gi])
                            dst[(gi * n_parts * NEQ + ei) * ld] = entry(cf + (gi * n_parts * NEQ + ei) * 4, bas[gi]);
                        if (dual and on[gi])
                            dst2[(gi * n_parts * NEQ + ei) * LDA] = entry(cf2 + (gi * n_parts * NEQ + ei) * 4, bas[gi]);
                    }
                if (rhs_row)
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
                        if (on[gi])
#pragma unroll
                            for (int r = 0; r < NRHS; ++r)
                            {
                                const double* c4  = cr + (gi * n_parts * NRHS + r) * 4;
                                double        acc = c4[0] * bas[gi][0];
#pragma unroll
                                for (int d = 0; d < DIM; ++d)
                                    acc = fma(c4[1 + d], bas[gi][1 + d], acc);
                                f_acc[r] += acc;
                            }
                dst += G * n_parts * NEQ * ld;
                cf += G * n_parts * NEQ * 4;
                dst2 += G * n_parts * NEQ * LDA;
                cf2 += G * n_parts * NEQ * 4;
                cr += G * n_parts * NRHS * 4;
            }
        }
        else
        {
            for (int qc = part; qc < QC; qc += n_parts)
            {
                double bas[4];
                basisAt(qc, bas);
                for (int ei = 0; ei < n_eq; ++ei)
                {
                    dst[ei * ld] = entry(cf + ei * 4, bas);
                    if (dual)
                        dst2[ei * LDA] = entry(cf2 + ei * 4, bas);
                }
                if (rhs_row)
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                    {
                        double acc = cr[r * 4] * bas[0];
#pragma unroll
                        for (int d = 0; d < DIM; ++d)
                            acc = fma(cr[r * 4 + 1 + d], bas[1 + d], acc);
                        f_acc[r] += acc;
                    }
                dst += n_parts * neq * ld;
                cf += n_parts * neq * 4;
                dst2 += n_parts * neq * LDA;
                cf2 += n_parts * neq * 4;
                cr += n_parts * NRHS * 4;
            }
        }
    };
    const auto build = [&](int c) {
        for (int pc = tid; pc < n_pcols * n_parts; pc += T)
        {
            const int pcol = pc % n_pcols, part = pc / n_pcols;
            switch (n_eq)
            {
            case 1: buildColumn(std::integral_constant< int, 1 >{}, c, pcol, part); break;
            case 2: buildColumn(std::integral_constant< int, 2 >{}, c, pcol, part); break;
            case 3: buildColumn(std::integral_constant< int, 3 >{}, c, pcol, part); break;
            case 4: buildColumn(std::integral_constant< int, 4 >{}, c, pcol, part); break;
            default: buildColumn(std::integral_constant< int, 0 >{}, c, pcol, part); break;
            }
        }
    };

    // ---- the build's table reads are L2 round trips (ncu r1 v3: 30 % of its samples on the long scoreboard): pull the slice of
    // chunk c into L1 one iteration before it is built — same thread, same addresses, no registers held
    const auto prefetchTables = [&](int c) {
        if (tensor)
            return;
        for (int pc = tid; pc < n_pcols * n_parts; pc += T)
        {
            const int  pcol = pc % n_pcols, part = pc / n_pcols;
            const bool is_b = pcol < TN;
            const int  node = (is_b ? col0 : row0) + (is_b ? pcol : pcol - TN);
            if (node >= NN or (not is_b and row0 == col0)) // same nodes as panel B: those lines are already on their way
                continue;
            for (int qc = part; qc < QC; qc += n_parts)
            {
                const int q = c * QC + qc;
                if (q >= args.n_qp)
                    break;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(tab_vals + q * NN + node));
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(tab_ders + (q * DIM + d) * NN + node));
            }
        }
    };

    // ---- accumulators: warp (wy, wx) owns rows wy*32 .. +32, columns wx*32 .. +32 of the tile as 4 x 4 DMMA tiles
    // For u == v the warp tiles strictly above the diagonal are dead, and a warp tile ON the diagonal only needs its 8 x 8 DMMA
    // tiles on or below it (10 of 16). Warps sit on the scheduler `warp % 4`, so the live tiles are dealt to the warps by cost:
    // warp t takes the t-th tile of the order [diagonal, diagonal, full, full, diagonal, diagonal, full, ...] — for the 4 x 4 case
    // the schedulers get 2 diagonal + 1 full, 2 diagonal + 1 full, 2 full, 2 full tiles (2.25 / 2.25 / 2 / 2 tile units).
    int  wy = warp / Cfg::WN, wx = warp % Cfg::WN;
    bool warp_live = true, warp_on_diag = false;
    if (diag_pair)
    {
        warp_live = false;
        int n_d = 0, n_f = 0; // diagonal / full live tiles met so far
        // position of the k-th diagonal (full) tile in the dealing order: two diagonals, two fulls, two diagonals, two fulls, rest
        const auto slotOfDiag = [](int k) { return k < 2 ? k : k < 4 ? k + 2 : -1; };
        const auto slotOfFull = [](int k) { return k < 2 ? k + 2 : k + 4; };
        int n_diag_total = 0;
        for (int i = 0; i < Cfg::WM * Cfg::WN; ++i)
            n_diag_total += col0 + (i % Cfg::WN) * 32 == row0 + (i / Cfg::WN) * 32;
        const bool pattern = n_diag_total == 4 and Cfg::WM * Cfg::WN == 16; // otherwise: diagonals first, then the full tiles
        int n_full_before = 0;
        for (int i = 0; i < Cfg::WM * Cfg::WN; ++i)
        {
            const int ty = i / Cfg::WN, tx = i % Cfg::WN;
            if (col0 + tx * 32 > row0 + ty * 32 + 31) // strictly above the diagonal
                continue;
            const bool on_diag = col0 + tx * 32 == row0 + ty * 32;
            const int  slot    = pattern ? (on_diag ? slotOfDiag(n_d) : slotOfFull(n_f)) : (on_diag ? n_d : n_diag_total + n_f);
            if (slot == warp)
            {
                wy           = ty;
                wx           = tx;
                warp_live    = true;
                warp_on_diag = on_diag;
            }
            n_d += on_diag;
            n_f += not on_diag;
            (void)n_full_before;
        }
    }
    const int g = lane >> 2, tq = lane & 3;
    double     acc[4][4][2];
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
        prefetchTables(1);
    __syncthreads();
#ifdef L3B_ASM_TIMING
    long long tm_build = 0, tm_mma = 0, tm_bar = 0;
#endif
    for (int c = 0; c < n_chunks; ++c)
    {
        if (c + 1 < n_chunks and (c + 1) % cps == 0)
        {
            // super-chunk boundary: the coefficients of the next chunks (nobody reads the old ones any more: build(c) is done)
            perPoint((c + 1) / cps);
            __syncthreads();
        }
        // every warp builds its share of the next panel, then contracts the current one (building in some warps while others
        // contract was measured slower: the build's short fp64 chains starve behind the DMMA stream)
        const bool build_first = true;
#ifdef L3B_ASM_TIMING
        long long t0 = clock64();
#endif
        const auto buildNext   = [&] {
            if (c + 2 < n_chunks)
                prefetchTables(c + 2);
            if (c + 1 < n_chunks)
                build(c + 1);
        };
        if (build_first)
            buildNext();
#ifdef L3B_ASM_TIMING
        long long t1 = clock64();
        if (build_first) tm_build += t1 - t0;
#endif
        if (warp_live)
        {
            const double* const pb_c = s_pb + (c & 1) * KCMAX * LDB + wx * 32 + g;
            const double* const pa_c = a_in_b ? s_pb + (c & 1) * KCMAX * LDB + (row0 - col0) + wy * 32 + g : s_pa + (c & 1) * KCMAX * LDA + wy * 32 + g;
            const int           lda  = a_in_b ? LDB : LDA;
#pragma unroll 2
            for (int k4 = 0; k4 < n_k4; ++k4)
            {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    a[i] = pa_c[(k4 * 4 + tq) * lda + i * 8];
                    b[i] = pb_c[(k4 * 4 + tq) * LDB + i * 8];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j <= i or not warp_on_diag) // a diagonal warp tile: nothing above its own diagonal is scattered
                            dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
#ifdef L3B_ASM_TIMING
        long long t2 = clock64();
        tm_mma += t2 - t1;
#endif
        if (not build_first)
            buildNext();
#ifdef L3B_ASM_TIMING
        long long t3 = clock64();
        if (not build_first) tm_build += t3 - t2;
#endif
        __syncthreads();
#ifdef L3B_ASM_TIMING
        tm_bar += clock64() - t3;
#endif
    }
#ifdef L3B_ASM_TIMING
    if (lane == 0 and blockIdx.x < 4096 and warp < 16)
    {
        g_asm_timing[blockIdx.x][warp][0] = tm_build;
        g_asm_timing[blockIdx.x][warp][1] = tm_mma;
        g_asm_timing[blockIdx.x][warp][2] = tm_bar;
        g_asm_timing[blockIdx.x][warp][3] = n_eq * 16 + (diag_pair ? 1 : 0);
    }
#endif

    // ---- scatter into the CRS values: slot(row (a,u), col (b,v)) = row_ptr[dof(a,u)] + dof_inds[v] * deg(node a) + pos[e][a][b]
    // (column-dof-major rows, device_common.cuh: consecutive nodes b are consecutive doubles, so the REDs of a warp share sectors)
    // DMMA accumulator layout: acc[i][j][h] = tile(row i*8 + g, column j*8 + 2*tq + h). For u == v only b <= a is
    // scattered, with its mirror image when b < a; for u != v every entry and its mirror K_e[(b,v)][(a,u)].
    if (warp_live)
    {
        const uint16_t* pos = args.slot_pos + e * static_cast< long long >(NN) * NN;
        const int       dpn = args.dofs_per_node, du = args.dof_inds[u], dv = args.dof_inds[v];
        long long       cbeg[4][2]; // row starts of the mirrored entries: rows (b, v)
        int             bcol[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                const int b = col0 + wx * 32 + j * 8 + 2 * tq + h;
                bcol[j][h]  = b;
                if (b < NN)
                {
                    const long long nb = el_nodes[b], np_b = args.node_ptr[nb], deg_b = args.node_ptr[nb + 1] - np_b;
                    cbeg[j][h]         = dpn * (dpn * np_b + dv * deg_b) + du * deg_b; // row (b, v), column dof u
                }
                else
                    cbeg[j][h] = 0;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const int a = row0 + wy * 32 + i * 8 + g;
            if (a >= NN)
                continue;
            const long long na = el_nodes[a], np_a = args.node_ptr[na], deg_a = args.node_ptr[na + 1] - np_a;
            double* const   rowp = args.crs_vals + dpn * (dpn * np_a + du * deg_a) + dv * deg_a; // row (a, u), column dof v
            const uint16_t* pa_  = pos + a * NN;
            // the slot positions of the whole row of tiles first (one memory round trip), then the atomics
            int  p_ab[4][2], p_ba[4][2];
            bool on[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                {
                    const int b = bcol[j][h];
                    on[j][h]    = b < NN and not(diag_pair and b > a);
                    p_ab[j][h]  = on[j][h] ? pa_[b] : 0;
                    p_ba[j][h]  = on[j][h] and (not diag_pair or b < a) ? pos[b * NN + a] : -1;
                }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                {
                    if (not on[j][h])
                        continue;
                    const double val = acc[i][j][h];
                    atomicAdd(rowp + p_ab[j][h], val);
                    if (p_ba[j][h] >= 0)
                        atomicAdd(args.crs_vals + cbeg[j][h] + p_ba[j][h], val);
                }
        }
    }
    // ---- rhs (ScatterLocalSystem.hpp:47-52): the A-side node this thread met in `build`
    if (rhs_duty)
        for (int pc = tid; pc < n_pcols * n_parts; pc += T)
        {
            const int  pcol = pc % n_pcols;
            const bool is_b = pcol < TN;
            const int  node = (is_b ? col0 : row0) + (is_b ? pcol : pcol - TN);
            if ((a_in_b ? (node >= row0 and node < row0 + TM) : not is_b) and node < NN)
            {
                const long long grow = static_cast< long long >(el_nodes[node]) * args.dofs_per_node + args.dof_inds[u];
                for (int r = 0; r < NRHS; ++r)
                    atomicAdd(args.rhs + grow + r * args.ld, f_acc[r]);
            }
        }
}
} // namespace l3b
#endif
