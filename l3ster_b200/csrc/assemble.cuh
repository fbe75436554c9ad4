// Element-local least-squares assembly fused with the CRS scatter:
//     K_e = sum_q w_q |J_q| B_q^T B_q,   F_e = sum_q w_q |J_q| B_q^T f_q,   B_q[e,(a,u)] = N_a A0(e,u) + sum_s dN_a/dx_s A_s(e,u)
// (algsys/AssembleLocalSystem.hpp:77-280), added straight into the rank-local CRS values / rhs through precomputed
// node-block slot maps (algsys/ScatterLocalSystem.hpp:25-54 with dofs/DofsFromNodes.hpp:72-87) — K_e (2 MB at p=4, U=4)
// is never materialised in HBM.
//
// One CTA owns one (element, row-block I, column-block J <= I) pair of the L x L local matrix, BLK x BLK entries, and
// streams the quadrature points in chunks: per chunk the two H panels (rows of block I scaled by w|J|, rows of block J
// unscaled — so negative weights need no sign split, cf. the reference's separate +/- batches, :117-119) are built in
// shared memory from the dense basis tables and the per-point kernel result, then contracted with an 8 x 8 register
// tile per thread (fp64 FMA). Off-diagonal blocks are scattered twice (K_e is symmetric, the reference mirrors the
// lower triangle, :176-182).
#ifndef L3B_ASSEMBLE_CUH
#define L3B_ASSEMBLE_CUH

#include "device_common.cuh"
#include "local_element.cuh"

namespace l3b
{
constexpr int asm_blk     = 128; // block edge of the K_e tiling
constexpr int asm_threads = 256; // 16 x 16 threads, 8 x 8 accumulators each

template < typename KernelT, int DIM, int P >
struct AsmCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields, NRHS = params.n_rhs;
    static constexpr int  NN = cpow(P + 1, DIM), L = NN * U;
    static constexpr int  n_blk   = (L + asm_blk - 1) / asm_blk;
    static constexpr int  n_pairs = n_blk * (n_blk + 1) / 2;
    static constexpr int  QPC     = cmax(1, 32 / E); // quadrature points per chunk
    static constexpr int  KC      = QPC * E;
    static constexpr int  qp_doubles = DIM * DIM + 2 + (DIM + 1) * E * U + E * NRHS; // Jti, weight, detJ, A, f
    // smem: H_I [KC][BLK] | H_J [KC][BLK] | per-point data [QPC] | node field values [NN][NF] | verts
    static constexpr int    off_hj = KC * asm_blk, off_qp = 2 * KC * asm_blk, off_nv = off_qp + QPC * qp_doubles,
                         off_verts = off_nv + NN * NF, total = off_verts + 8 * 3;
    static constexpr size_t smem_bytes = static_cast< size_t >(total) * sizeof(double);
};

template < typename KernelT, int DIM, int P >
__global__ void __launch_bounds__(asm_threads) assembleKernel(const KernelT kernel, const __grid_constant__ ElemArgs args)
{
    using Cfg = AsmCfg< KernelT, DIM, P >;
    constexpr int  E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NRHS = Cfg::NRHS, NN = Cfg::NN, L = Cfg::L, QPC = Cfg::QPC, KC = Cfg::KC;
    constexpr bool is_bnd = KernelT::is_boundary;
    constexpr int  nv     = 1 << DIM;
    extern __shared__ double smem[];
    double* s_hi    = smem;
    double* s_hj    = smem + Cfg::off_hj;
    double* s_qp    = smem + Cfg::off_qp;
    double* s_nv    = smem + Cfg::off_nv;
    double* s_verts = smem + Cfg::off_verts;

    const int       tid  = threadIdx.x;
    const long long wi   = blockIdx.x / Cfg::n_pairs;
    int             pair = blockIdx.x % Cfg::n_pairs;
    int             bi   = 0;
    while (pair >= bi + 1) // pair index → (bi, bj) with bj <= bi
    {
        pair -= bi + 1;
        ++bi;
    }
    const int       bj   = pair;
    const long long e    = args.work_elems ? args.work_elems[wi] : args.first_elem + wi;
    const int       side = is_bnd ? args.work_sides[wi] : -1;
    const uint32_t* el_nodes = args.nodes + e * NN;

    for (int i = tid; i < nv * 3; i += asm_threads)
        s_verts[i] = args.verts[e * nv * 3 + i];
    if constexpr (NF > 0)
        for (int i = tid; i < NN * NF; i += asm_threads)
            s_nv[i] = args.fields[el_nodes[i / NF] + args.field_inds[i % NF] * args.field_stride];
    __syncthreads();

    const long long tab_off  = is_bnd ? static_cast< long long >(side) * args.n_qp : 0;
    const double*   tab_vals = args.tab_vals + tab_off * NN;
    const double*   tab_ders = args.tab_ders + tab_off * DIM * NN;
    const double*   tab_pts  = args.tab_pts + tab_off * DIM;
    const double*   tab_wts  = args.tab_wts + tab_off;

    // the panel row this thread builds: threads [0,128) rows of block I (scaled), [128,256) rows of block J
    const bool build_i = tid < asm_blk;
    const int  prow    = (build_i ? bi : bj) * asm_blk + (tid % asm_blk); // local row index (a*U + u), may be >= L (padding)
    const bool prow_ok = prow < L;
    const int  pa = prow_ok ? prow / U : 0, pu = prow_ok ? prow % U : 0;
    double     f_acc[NRHS];
    for (int r = 0; r < NRHS; ++r)
        f_acc[r] = 0.;

    // accumulators: rows bi*BLK + m*32 + ty*2 + {0,1}, cols bj*BLK + m*32 + tx*2 + {0,1}, m = 0..3
    const int tx = tid % 16, ty = tid / 16;
    double    acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            acc[i][j] = 0.;

    const int warp = tid >> 5, lane = tid & 31;
    for (int q0 = 0; q0 < args.n_qp; q0 += QPC)
    {
        const int nq_here = min(QPC, args.n_qp - q0);
        // ---- per-point data: warp w handles point q0 + w (mapping, fields, user kernel)
        for (int qc = warp; qc < nq_here; qc += asm_threads / 32)
        {
            const int q = q0 + qc;
            double    xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], nrm[DIM];
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
                        double v = 0.;
                        for (int d = 0; d < DIM; ++d)
                            v = fma(Jti[s][d], bd[d * NN + a], v);
                        pd[s] = v;
                    }
                    for (int f = 0; f < NF; ++f)
                    {
                        const double v = s_nv[a * NF + f];
                        fred[f]        = fma(bv[a], v, fred[f]);
                        for (int s = 0; s < DIM; ++s)
                            fred[NF * (s + 1) + f] = fma(pd[s], v, fred[NF * (s + 1) + f]);
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
                double*    qd  = s_qp + qc * Cfg::qp_doubles;
                for (int s = 0; s < DIM; ++s)
                    for (int d = 0; d < DIM; ++d)
                        qd[s * DIM + d] = Jti[s][d];
                qd[DIM * DIM]     = jac * tab_wts[q];
                qd[DIM * DIM + 1] = detJ;
                double* qa        = qd + DIM * DIM + 2;
                for (int i = 0; i <= DIM; ++i)
                    for (int k = 0; k < E * U; ++k)
                        qa[i * E * U + k] = res.operators[i].v[k];
                for (int k = 0; k < E * NRHS; ++k)
                    qa[(DIM + 1) * E * U + k] = res.rhs.v[k];
            }
        }
        __syncthreads();
        // ---- build the two H panels for this chunk: s_h*[k][row], k = qc*E + eq
        {
            double* dst = (build_i ? s_hi : s_hj) + (tid % asm_blk);
            for (int qc = 0; qc < QPC; ++qc)
            {
                if (qc < nq_here and prow_ok)
                {
                    const int     q  = q0 + qc;
                    const double* qd = s_qp + qc * Cfg::qp_doubles;
                    const double* qa = qd + DIM * DIM + 2;
                    const double  w  = qd[DIM * DIM];
                    const double  n  = tab_vals[static_cast< long long >(q) * NN + pa];
                    double        pd[DIM];
                    for (int s = 0; s < DIM; ++s)
                    {
                        double v = 0.;
                        for (int d = 0; d < DIM; ++d)
                            v = fma(qd[s * DIM + d], tab_ders[(static_cast< long long >(q) * DIM + d) * NN + pa], v);
                        pd[s] = v;
                    }
#pragma unroll
                    for (int eq = 0; eq < E; ++eq)
                    {
                        double b = n * qa[eq + pu * E];
                        for (int s = 0; s < DIM; ++s)
                            b = fma(pd[s], qa[(s + 1) * E * U + eq + pu * E], b);
                        if (build_i)
                        {
                            if (bj == 0)
                                for (int r = 0; r < NRHS; ++r)
                                    f_acc[r] = fma(b * w, qa[(DIM + 1) * E * U + eq + r * E], f_acc[r]);
                            b *= w;
                        }
                        dst[(qc * E + eq) * asm_blk] = b;
                    }
                }
                else
#pragma unroll
                    for (int eq = 0; eq < E; ++eq)
                        dst[(qc * E + eq) * asm_blk] = 0.;
            }
        }
        __syncthreads();
        // ---- contract: acc[i][j] += H_I[k][row_i] * H_J[k][col_j]
#pragma unroll 4
        for (int k = 0; k < KC; ++k)
        {
            double hi[8], hj[8];
#pragma unroll
            for (int m = 0; m < 4; ++m)
            {
                const double2 a = *reinterpret_cast< const double2* >(s_hi + k * asm_blk + m * 32 + ty * 2);
                const double2 b = *reinterpret_cast< const double2* >(s_hj + k * asm_blk + m * 32 + tx * 2);
                hi[2 * m]     = a.x;
                hi[2 * m + 1] = a.y;
                hj[2 * m]     = b.x;
                hj[2 * m + 1] = b.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    acc[i][j] = fma(hi[i], hj[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- scatter into the CRS values: slot(row (a,u), col (b,v)) = row_ptr[dof(a,u)] + dof_inds[v] * deg(node a) + pos[e][a][b]
    // (device layout of the values: device_common.cuh, ElemArgs)
    const uint16_t* pos = args.slot_pos + e * static_cast< long long >(NN) * NN;
    const int       dpn = args.dofs_per_node;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        const int r = bi * asm_blk + (i / 2) * 32 + ty * 2 + (i % 2);
        if (r >= L)
            continue;
        const int       ra = r / U, ru = r % U;
        const long long na = el_nodes[ra], np_a = args.node_ptr[na], deg_a = args.node_ptr[na + 1] - np_a;
        const long long rbeg = dpn * (dpn * np_a + args.dof_inds[ru] * deg_a);
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int c = bj * asm_blk + (j / 2) * 32 + tx * 2 + (j % 2);
            if (c >= L)
                continue;
            const int ca = c / U, cu = c % U;
            atomicAdd(args.crs_vals + rbeg + args.dof_inds[cu] * deg_a + pos[ra * NN + ca], acc[i][j]);
            if (bi != bj) // mirrored entry K_e[c][r]
            {
                const long long nb = el_nodes[ca], np_b = args.node_ptr[nb], deg_b = args.node_ptr[nb + 1] - np_b;
                atomicAdd(args.crs_vals + dpn * (dpn * np_b + args.dof_inds[cu] * deg_b) + args.dof_inds[ru] * deg_b + pos[ca * NN + ra],
                          acc[i][j]);
            }
        }
    }
    // ---- rhs (ScatterLocalSystem.hpp:47-52)
    if (build_i and bj == 0 and prow_ok)
    {
        const long long grow = static_cast< long long >(el_nodes[pa]) * args.dofs_per_node + args.dof_inds[pu];
        for (int r = 0; r < NRHS; ++r)
            atomicAdd(args.rhs + grow + r * args.ld, f_acc[r]);
    }
}

} // namespace l3b
#endif
