// Sum-factorised matrix-free local operator apply, fused with gather and scatter:
//     y[dofs(e)] += alpha * K_e * x[dofs(e)]      for every element e of the work list,
// K_e = sum_q [N_q (x) A0^T + sum_d d_xi_d N_q (x) D_d^T] w|J| [A0 u_q + sum_d D_d d_xi_d u_q]  (SURVEY Appendix A).
//
// Replaces evalLocalOperatorSumFact + gatherSumFact/scatterSumFact of the reference
// (algsys/SumFactorization.hpp:438-917, algsys/MatrixFreeSystem.hpp:421-537).
//
// Formulation (differs from the reference in rounding only, parity 1e-12):
//   * u is interpolated from the GLL nodes to the nq^D Gauss points with D sweeps; reference-space derivatives are then
//     taken *at the Gauss points* with the collocation derivative matrix of the Gauss-point Lagrange basis (exact for
//     nq >= nb, which AssemblyOptions guarantees for value_order >= 1) — 2D sweeps instead of the reference's 5 (quad) /
//     9 (hex);
//   * per quadrature point the physical gradient g_s = sum_d Ji(d,s) d_xi_d u is formed first, then
//     t = w|J| (A0 u + sum_s A_s g_s), r0 = A0^T t, p_s = A_s^T t, r_d = sum_s Ji(d,s) p_s — this never forms the
//     reference's D_d = sum_s A_s Ji(d,s) matrices (SumFactorization.hpp:736-738);
//   * multiply-adds against structurally zero operator entries are not issued (KernelSparsity, kernel_interface.cuh):
//     15 of 112 entries survive for 3-D diffusion. Every evaluation still checks the masked entries are zero;
//   * the multilinear geometry is evaluated from its monomial coefficients (8 x 3 numbers per hex, built once per element
//     in shared memory) instead of summing over the vertices at every point (computeGeomDataLin, :506-537);
//   * the hex path hands the user kernel point.space = (x, y, 0), replicating SumFactorization.hpp:732.
//
// Thread mapping: TPE = nq^D threads per element (one per Gauss point / node), EPB elements per CTA. All tensors of an
// element live in shared memory: (D+1) x F arrays (values + D reference derivatives) in a padded nq^D layout in which
// every sweep runs in place — one 1-D line per thread, loaded into registers, transformed with constant-bank
// coefficients, stored back. The last interpolation sweep also differentiates its own direction, the remaining
// derivatives are line sweeps over the interpolated values, the per-point stage overwrites its own entries with the
// fluxes r_0..r_D, and the transposed stage mirrors all of it. Shared-memory traffic — the binding resource of the v1/v2
// kernels, which read 15 neighbour values + 15 table entries per field and point — drops to ~128 KB per p=4 element.
#ifndef L3B_MF_SUMFACT_CUH
#define L3B_MF_SUMFACT_CUH

#include "device_common.cuh"

namespace l3b
{
template < int NB, int NQ >
struct SumFactTables
{
    double interp[NB * NQ]; // [b][q]
    double der[NB * NQ];    // [b][q]
    double colloc[NQ * NQ]; // [m][q]
    double w[NQ];
    double pts[NQ];
};

template < typename KernelT, int DIM, int P, int NQ, int NRHS >
struct MfSumFactCfg
{
    static constexpr auto params = KernelT::parameters;
    static constexpr int  E = params.n_equations, U = params.n_unknowns, NF = params.n_fields;
    static constexpr int  NB = P + 1;
    static constexpr int  F0 = U * NRHS, F = F0 + NF;
    static constexpr int  NN = cpow(NB, DIM), Q = cpow(NQ, DIM);
    // shared-memory tensor layout of one field: idx(qx, qy, qz) = qz * PS + qy * LS + qx. The line stride is padded to an
    // odd number of doubles so that the x-lines of one half-warp fall into distinct banks for every nq.
    static constexpr int LS = NQ % 2 == 0 ? NQ + 1 : NQ;
    static constexpr int PS = LS * NQ;
    static constexpr int AQ_raw = DIM == 3 ? PS * NQ : PS;
    // doubles per field array, padded to AQ ≡ NQ (mod 16): lanes ordered (qx fastest, then field) then hit consecutive
    // even banks in the sweeps whose lines start at (qx, k) — the y-type sweeps, which conflicted 2-way in v3
    static constexpr int AQ = AQ_raw + ((NQ - AQ_raw) % 16 + 16) % 16;
    static constexpr int TPE = Q;
    static constexpr int EPB = cmax(1, 128 / TPE);
    static constexpr int threads = TPE * EPB;
    // resident CTAs per SM the register allocator should leave room for (~20 warps per SM)
    static constexpr int warps      = (threads + 31) / 32;
    static constexpr int min_blocks = cmax(1, 20 / warps);
    // per element: values + DIM reference derivatives of all F fields (the DIM+1 flux arrays of the transposed stage
    // overwrite them in place), the geometry coefficients and, for affine elements, the shared inverse Jacobian
    static constexpr int geo_doubles = (1 << DIM) * 3 + DIM * DIM + 2;
    static constexpr int smem_doubles_per_elem = (DIM + 1) * F * AQ + geo_doubles;
    static constexpr size_t smem_bytes = static_cast< size_t >(EPB) * smem_doubles_per_elem * sizeof(double);
    static_assert(params.dimension == DIM);
    static_assert(NQ >= NB, "collocation differentiation at the Gauss points needs nq >= nb (value_order >= 1)");
    static_assert(threads <= 1024, "element too large for one thread per tensor entry");
};

// ---- 1-D line kernels. A line lives in shared memory at `p[i * stride]`; each is processed by one thread, entirely in
// registers, and written back in place (or to a second array). Table indices are compile-time after unrolling, so the
// coefficients become constant-bank operands of the DFMAs.
template < int NB, int NQ >
__device__ __forceinline__ void loadLine(const double* __restrict__ p, int stride, double (&v)[NQ], int n)
{
#pragma unroll
    for (int i = 0; i < NQ; ++i)
        if (i < n)
            v[i] = p[i * stride];
}
// nodes → Gauss points: out[q] = sum_i v[i] interp[i][q]
template < int NB, int NQ >
__device__ __forceinline__ void interpolate(const double (&v)[NQ], double (&out)[NQ], const double* __restrict__ interp)
{
#pragma unroll
    for (int q = 0; q < NQ; ++q)
    {
        double acc = v[0] * interp[q];
#pragma unroll
        for (int i = 1; i < NB; ++i)
            acc = fma(v[i], interp[i * NQ + q], acc);
        out[q] = acc;
    }
}
// Gauss points → nodes (transpose): out[i] = sum_q v[q] interp[i][q]
template < int NB, int NQ >
__device__ __forceinline__ void interpolateT(const double (&v)[NQ], double (&out)[NQ], const double* __restrict__ interp)
{
#pragma unroll
    for (int i = 0; i < NB; ++i)
    {
        double acc = v[0] * interp[i * NQ];
#pragma unroll
        for (int q = 1; q < NQ; ++q)
            acc = fma(v[q], interp[i * NQ + q], acc);
        out[i] = acc;
    }
}
// collocation derivative at the Gauss points: out[q] = sum_m colloc[m][q] v[m]
template < int NQ >
__device__ __forceinline__ void differentiate(const double (&v)[NQ], double (&out)[NQ], const double* __restrict__ colloc)
{
#pragma unroll
    for (int q = 0; q < NQ; ++q)
    {
        double acc = v[0] * colloc[q];
#pragma unroll
        for (int m = 1; m < NQ; ++m)
            acc = fma(v[m], colloc[m * NQ + q], acc);
        out[q] = acc;
    }
}
// its transpose, accumulated: acc[m] += sum_q colloc[m][q] r[q]
template < int NQ >
__device__ __forceinline__ void differentiateTAccumulate(const double (&r)[NQ], double (&acc)[NQ], const double* __restrict__ colloc)
{
#pragma unroll
    for (int m = 0; m < NQ; ++m)
    {
        double a = acc[m];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            a = fma(r[q], colloc[m * NQ + q], a);
        acc[m] = a;
    }
}
template < int NQ >
__device__ __forceinline__ void storeLine(double* __restrict__ p, int stride, const double (&v)[NQ], int n)
{
#pragma unroll
    for (int i = 0; i < NQ; ++i)
        if (i < n)
            p[i * stride] = v[i];
}

// monomial coefficients of the multilinear map: x(xi) = sum_m c_m prod_{d in m} xi_d, m = bitmask over directions
template < int DIM >
__device__ __forceinline__ void buildGeometryCoefs(const double* __restrict__ verts, double* __restrict__ s_geo, int t, int n_threads)
{
    constexpr int nv = 1 << DIM;
    for (int i = t; i < nv * 3; i += n_threads)
    {
        const int m = i / 3, s = i % 3;
        double    acc = 0.;
#pragma unroll
        for (int v = 0; v < nv; ++v)
        {
            // sign of vertex v in monomial m: product over the directions in m of (+1 if the vertex sits at xi_d = +1 else -1)
            const int neg = __popc(m & ~v) & 1;
            acc += neg ? -verts[v * 3 + s] : verts[v * 3 + s];
        }
        s_geo[i] = acc * (1. / nv);
    }
}

// x_s (s < 3) and Jt[d][s] = dx_s/dxi_d at xi from the monomial coefficients
template < int DIM >
__device__ __forceinline__ void geometryFromCoefs(const double* __restrict__ c, const double (&xi)[DIM], double (&xs)[3], double (&Jt)[DIM][DIM])
{
    if constexpr (DIM == 2)
    {
#pragma unroll
        for (int s = 0; s < 3; ++s)
            xs[s] = fma(xi[0], fma(xi[1], c[3 * 3 + s], c[1 * 3 + s]), fma(xi[1], c[2 * 3 + s], c[s]));
#pragma unroll
        for (int s = 0; s < 2; ++s)
        {
            Jt[0][s] = fma(xi[1], c[3 * 3 + s], c[1 * 3 + s]);
            Jt[1][s] = fma(xi[0], c[3 * 3 + s], c[2 * 3 + s]);
        }
    }
    else
    {
        const double xy = xi[0] * xi[1], xz = xi[0] * xi[2], yz = xi[1] * xi[2];
        // monomial index bits: 1 = xi, 2 = eta, 4 = zeta
#pragma unroll
        for (int s = 0; s < 3; ++s)
        {
            const double c0 = c[s], cx = c[3 + s], cy = c[6 + s], cxy = c[9 + s], cz = c[12 + s], cxz = c[15 + s], cyz = c[18 + s], cxyz = c[21 + s];
            Jt[0][s] = fma(yz, cxyz, fma(xi[2], cxz, fma(xi[1], cxy, cx)));
            Jt[1][s] = fma(xz, cxyz, fma(xi[2], cyz, fma(xi[0], cxy, cy)));
            Jt[2][s] = fma(xy, cxyz, fma(xi[1], cyz, fma(xi[0], cxz, cz)));
            xs[s]    = fma(xi[2], Jt[2][s], fma(xy, cxy, fma(xi[1], cy, fma(xi[0], cx, c0))));
        }
    }
}

template < typename KernelT, int DIM, int P, int NQ, int NRHS >
__global__ void __launch_bounds__(MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::threads, MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >::min_blocks)
    mfSumFactApplyKernel(const KernelT kernel, const __grid_constant__ ElemArgs args, const __grid_constant__ SumFactTables< P + 1, NQ > tab)
{
    using Cfg = MfSumFactCfg< KernelT, DIM, P, NQ, NRHS >;
    using Sp  = KernelSparsity< KernelT >;
    constexpr int E = Cfg::E, U = Cfg::U, NF = Cfg::NF, NB = Cfg::NB, F0 = Cfg::F0, F = Cfg::F, NN = Cfg::NN;
    constexpr int LS = Cfg::LS, PS = Cfg::PS, AQ = Cfg::AQ, TPE = Cfg::TPE;
    constexpr int nv = 1 << DIM;
    extern __shared__ double smem[];
    const int       slot   = threadIdx.x / TPE;
    const int       t      = threadIdx.x % TPE;
    const long long wi     = static_cast< long long >(blockIdx.x) * Cfg::EPB + slot;
    const bool      active = wi < args.n_work;
    const long long e      = active ? (args.work_elems ? args.work_elems[wi] : args.first_elem + wi) : 0;
    // field arrays: s_val[f], s_der[d][f] (f < F), each AQ doubles
    double* const   s_val  = smem + static_cast< size_t >(slot) * Cfg::smem_doubles_per_elem;
    double* const   s_der  = s_val + F * AQ;
    double* const   s_geo  = s_val + (DIM + 1) * F * AQ; // [2^D][3] monomial coefficients, then Jti[D][D], detJ, affine flag
    const uint32_t* el_nodes = args.nodes + e * NN;
    const auto      der_arr = [&](int d, int f) { return s_der + (d * F + f) * AQ; };

    if (active)
        buildGeometryCoefs< DIM >(args.verts + e * nv * 3, s_geo, t, TPE);
    // x^T A x of operand column 0 (ElemArgs::energy): per-thread share of sum_q w |B_q x_e|^2, summed through shared memory
    __shared__ double s_energy;
    const bool        want_energy = args.energy != nullptr;
    double            energy      = 0.;
    if (threadIdx.x == 0)
        s_energy = 0.;

    // ---- gather (MatrixFreeSystem.hpp:421-467) into s_val[f], f = rhs*U + u, then the NF external fields.
    // One lane per (node, unknown): the U dofs of a node are adjacent in x, so a warp touches ~32/U sectors per load
    // instead of 32 (thread-per-node would).
    if (active)
    {
        for (int i = t; i < NN * U; i += TPE)
        {
            const int       a = i / U, u = i % U;
            const long long node = el_nodes[a];
            const int       pos  = DIM == 3 ? (a / (NB * NB)) * PS + ((a / NB) % NB) * LS + a % NB : (a / NB) * LS + a % NB;
            const long long dof  = node * args.dofs_per_node + args.dof_inds[u];
            const bool      dir  = isDirichlet(args.dir_mask, dof);
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                s_val[(r * U + u) * AQ + pos] = dir ? 0. : __ldg(args.x + dof + r * args.ld);
        }
        if constexpr (NF > 0)
            for (int i = t; i < NN * NF; i += TPE)
            {
                const int       a = i % NN, f = i / NN;
                const int       pos = DIM == 3 ? (a / (NB * NB)) * PS + ((a / NB) % NB) * LS + a % NB : (a / NB) * LS + a % NB;
                s_val[(F0 + f) * AQ + pos] = __ldg(args.fields + el_nodes[a] + args.field_inds[f] * args.field_stride);
            }
    }
    __syncthreads();
    // affine element (all mixed monomial coefficients vanish): one thread inverts the constant Jacobian for everybody
    if (active and t == 0)
    {
        bool affine = true;
        for (int m = 0; m < nv; ++m)
            if (__popc(m) > 1)
                for (int s = 0; s < 3; ++s)
                    affine = affine and s_geo[m * 3 + s] == 0.;
        double* g = s_geo + nv * 3;
        g[DIM * DIM + 1] = affine ? 1. : 0.;
        if (affine)
        {
            double Jt[DIM][DIM], Jti[DIM][DIM];
            for (int d = 0; d < DIM; ++d)
                for (int s = 0; s < DIM; ++s)
                    Jt[d][s] = s_geo[(1 << d) * 3 + s];
            g[DIM * DIM] = invert< DIM >(Jt, Jti);
            for (int s = 0; s < DIM; ++s)
                for (int d = 0; d < DIM; ++d)
                    g[s * DIM + d] = Jti[s][d];
        }
    }

    // ---- nodes → Gauss points, in place, one direction at a time; the last sweep also differentiates along its direction
    {
        double v[NQ], o[NQ];
        // x: lines (f, k, j), j, k < NB
        constexpr int n_x = DIM == 3 ? F * NB * NB : F * NB;
        for (int l = t; l < n_x; l += TPE)
        {
            const int f = l / (n_x / F), r = l % (n_x / F);
            double*   p = s_val + f * AQ + (DIM == 3 ? (r / NB) * PS + (r % NB) * LS : r * LS);
            if (active)
            {
                loadLine< NB, NQ >(p, 1, v, NB);
                interpolate< NB, NQ >(v, o, tab.interp);
                if constexpr (DIM == 2)
                    ; // y is the last direction in 2-D
                storeLine< NQ >(p, 1, o, NQ);
            }
        }
        __syncthreads();
        if constexpr (DIM == 3)
        {
            // y: lines (qx fastest, then f, then k < NB)
            for (int l = t; l < F * NB * NQ; l += TPE)
            {
                const int qx = l % NQ, f = (l / NQ) % F, k = l / (NQ * F);
                double*   p = s_val + f * AQ + k * PS + qx;
                if (active)
                {
                    loadLine< NB, NQ >(p, LS, v, NB);
                    interpolate< NB, NQ >(v, o, tab.interp);
                    storeLine< NQ >(p, LS, o, NQ);
                }
            }
            __syncthreads();
        }
        // last direction (y in 2-D, z in 3-D): lines over all (qx[, qy]); interpolate and differentiate in one pass
        constexpr int n_l    = DIM == 3 ? NQ * NQ : NQ;
        constexpr int stride = DIM == 3 ? PS : LS;
        for (int l = t; l < F * n_l; l += TPE)
        {
            const int f = l / n_l, r = l % n_l;
            const int off = DIM == 3 ? (r / NQ) * LS + r % NQ : r;
            if (active)
            {
                loadLine< NB, NQ >(s_val + f * AQ + off, stride, v, NB);
                interpolate< NB, NQ >(v, o, tab.interp);
                storeLine< NQ >(s_val + f * AQ + off, stride, o, NQ);
                differentiate< NQ >(o, v, tab.colloc);
                storeLine< NQ >(der_arr(DIM - 1, f) + off, stride, v, NQ);
            }
        }
        __syncthreads();
        // remaining reference derivatives: x (and y in 3-D) lines of the interpolated values
        constexpr int n_dl = DIM == 3 ? NQ * NQ : NQ;
        for (int l = t; l < (DIM - 1) * F * n_dl; l += TPE)
        {
            const int d = l / (F * n_dl);
            int       f, off, st;
            if (d == 0) // x-lines (qy[, qz]); tasks ordered (line, f)
            {
                const int r = l % n_dl;
                f           = (l / n_dl) % F;
                off         = DIM == 3 ? (r / NQ) * PS + (r % NQ) * LS : r * LS;
                st          = 1;
            }
            else // y-lines, 3-D only; tasks ordered (qx fastest, then f, then qz)
            {
                const int ll = l - F * n_dl;
                f            = (ll / NQ) % F;
                off          = (ll / (NQ * F)) * PS + ll % NQ;
                st           = LS;
            }
            if (active)
            {
                loadLine< NQ, NQ >(s_val + f * AQ + off, st, v, NQ);
                differentiate< NQ >(v, o, tab.colloc);
                storeLine< NQ >(der_arr(d, f) + off, st, o, NQ);
            }
        }
        __syncthreads();
    }

    // ---- quadrature point stage (SumFactorization.hpp:614-756); one point per thread, in place on its own entries
    const int q     = t;
    const int qi[3] = {q % NQ, (q / NQ) % NQ, DIM == 3 ? q / (NQ * NQ) : 0};
    const int qpos  = qi[2] * PS + qi[1] * LS + qi[0];
    if (active)
    {
        double val[F], dref[DIM][F];
#pragma unroll
        for (int f = 0; f < F; ++f)
        {
            val[f] = s_val[f * AQ + qpos];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                dref[d][f] = der_arr(d, f)[qpos];
        }
        double xi[DIM], xs[3], Jt[DIM][DIM], Jti[DIM][DIM], detJ;
        double wq = 1.;
        // table entries by run-time index: tiny select chains over constant-bank operands
#pragma unroll
        for (int d = 0; d < DIM; ++d)
        {
            double pt = tab.pts[0], w = tab.w[0];
#pragma unroll
            for (int k = 1; k < NQ; ++k)
                if (qi[d] == k)
                {
                    pt = tab.pts[k];
                    w  = tab.w[k];
                }
            xi[d] = pt;
            wq *= w;
        }
        geometryFromCoefs< DIM >(s_geo, xi, xs, Jt);
        const double* g = s_geo + nv * 3;
        if (g[DIM * DIM + 1] != 0.)
        {
            detJ = g[DIM * DIM];
#pragma unroll
            for (int s = 0; s < DIM; ++s)
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    Jti[s][d] = g[s * DIM + d];
        }
        else
            detJ = invert< DIM >(Jt, Jti);
        // Jt[d][s] = dx_s/dxi_d, so Jti[s][d] = dxi_d/dx_s = the reference's jac_inv(d, s)
        typename KernelT::Input in;
#pragma unroll
        for (int f = 0; f < NF; ++f)
        {
            in.field_vals[f] = val[F0 + f];
#pragma unroll
            for (int s = 0; s < DIM; ++s)
            {
                double acc = Jti[s][0] * dref[0][F0 + f];
#pragma unroll
                for (int d = 1; d < DIM; ++d)
                    acc = fma(Jti[s][d], dref[d][F0 + f], acc);
                in.field_ders[s][f] = acc;
            }
        }
        in.point.space.coords[0] = xs[0];
        in.point.space.coords[1] = xs[1];
        in.point.space.coords[2] = 0.; // SumFactorization.hpp:656, :732
        in.point.time            = args.time;
        const auto   res = kernel(in);
        const double wgt = wq * detJ;
        // guard: entries the compile-time probe declared structurally zero must be zero (folds away when provable)
        bool violated = false;
        staticFor< DIM + 1 >([&](auto op) {
            staticFor< E >([&](auto eq) {
                staticFor< U >([&](auto u) {
                    if constexpr (not Sp::nz(op, eq, u))
                        violated |= res.operators[op](eq, u) != 0.;
                });
            });
        });
        if (violated)
            atomicOr(args.status, status_sparsity_violation);
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
        {
            double g_phys[DIM][U]; // physical gradients of the operand
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                {
                    double acc = Jti[s][0] * dref[0][r * U + u];
#pragma unroll
                    for (int d = 1; d < DIM; ++d)
                        acc = fma(Jti[s][d], dref[d][r * U + u], acc);
                    g_phys[s][u] = acc;
                }
            double tv[E];
            staticFor< E >([&](auto eq) {
                double acc = 0.;
                staticFor< U >([&](auto u) {
                    if constexpr (Sp::nz(0, eq, u))
                        acc = fma(res.operators[0](eq, u), val[r * U + u], acc);
                    staticFor< DIM >([&](auto s) {
                        if constexpr (Sp::nz(s + 1, eq, u))
                            acc = fma(res.operators[s + 1](eq, u), g_phys[s][u], acc);
                    });
                });
                tv[eq] = acc * wgt;
                if (want_energy and r == 0)
                    energy = fma(acc, tv[eq], energy);
            });
            staticFor< U >([&](auto u) {
                double a0 = 0., ps[DIM];
#pragma unroll
                for (int s = 0; s < DIM; ++s)
                    ps[s] = 0.;
                staticFor< E >([&](auto eq) {
                    if constexpr (Sp::nz(0, eq, u))
                        a0 = fma(res.operators[0](eq, u), tv[eq], a0);
                    staticFor< DIM >([&](auto s) {
                        if constexpr (Sp::nz(s + 1, eq, u))
                            ps[s] = fma(res.operators[s + 1](eq, u), tv[eq], ps[s]);
                    });
                });
                // fluxes overwrite this thread's own entries: r0 → values, r_d → derivative arrays
                s_val[(r * U + u) * AQ + qpos] = a0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                {
                    double acc = Jti[0][d] * ps[0];
#pragma unroll
                    for (int s = 1; s < DIM; ++s)
                        acc = fma(Jti[s][d], ps[s], acc);
                    der_arr(d, r * U + u)[qpos] = acc;
                }
            });
        }
        if (want_energy)
            atomicAdd(&s_energy, energy);
    }
    __syncthreads();
    if (want_energy and threadIdx.x == 0)
        atomicAdd(args.energy, s_energy);

    // ---- transposed stage: v = r0 + sum_d D_d^T r_d, then Gauss points → nodes; all in place on s_val[f < F0]
    {
        double v[NQ], r[NQ], o[NQ];
        constexpr int n_dl = DIM == 3 ? NQ * NQ : NQ;
        // x- (and, in 3-D, y-) derivative transposes, one direction per phase (both update s_val)
        for (int d = 0; d < DIM - 1; ++d)
        {
            for (int l = t; l < F0 * n_dl; l += TPE)
            {
                int f, off, st;
                if (d == 0)
                {
                    const int rr = l % n_dl;
                    f            = l / n_dl;
                    off          = DIM == 3 ? (rr / NQ) * PS + (rr % NQ) * LS : rr * LS;
                    st           = 1;
                }
                else
                {
                    f   = (l / NQ) % F0;
                    off = (l / (NQ * F0)) * PS + l % NQ;
                    st  = LS;
                }
                if (active)
                {
                    loadLine< NQ, NQ >(s_val + f * AQ + off, st, v, NQ);
                    loadLine< NQ, NQ >(der_arr(d, f) + off, st, r, NQ);
                    differentiateTAccumulate< NQ >(r, v, tab.colloc);
                    storeLine< NQ >(s_val + f * AQ + off, st, v, NQ);
                }
            }
            __syncthreads();
        }
        // last direction: derivative transpose fused with the projection to the nodes
        constexpr int n_l    = DIM == 3 ? NQ * NQ : NQ;
        constexpr int stride = DIM == 3 ? PS : LS;
        for (int l = t; l < F0 * n_l; l += TPE)
        {
            const int f = l / n_l, rr = l % n_l;
            const int off = DIM == 3 ? (rr / NQ) * LS + rr % NQ : rr;
            if (active)
            {
                loadLine< NQ, NQ >(s_val + f * AQ + off, stride, v, NQ);
                loadLine< NQ, NQ >(der_arr(DIM - 1, f) + off, stride, r, NQ);
                differentiateTAccumulate< NQ >(r, v, tab.colloc);
                interpolateT< NB, NQ >(v, o, tab.interp);
                storeLine< NQ >(s_val + f * AQ + off, stride, o, NB);
            }
        }
        __syncthreads();
        if constexpr (DIM == 3)
        {
            // y: lines (qx fastest, then f, then k < NB)
            for (int l = t; l < F0 * NB * NQ; l += TPE)
            {
                const int qx = l % NQ, f = (l / NQ) % F0, k = l / (NQ * F0);
                double*   p = s_val + f * AQ + k * PS + qx;
                if (active)
                {
                    loadLine< NQ, NQ >(p, LS, v, NQ);
                    interpolateT< NB, NQ >(v, o, tab.interp);
                    storeLine< NQ >(p, LS, o, NB);
                }
            }
            __syncthreads();
        }
        // x: lines (f, k, j), j, k < NB
        constexpr int n_x = DIM == 3 ? F0 * NB * NB : F0 * NB;
        for (int l = t; l < n_x; l += TPE)
        {
            const int f = l / (n_x / F0), rr = l % (n_x / F0);
            double*   p = s_val + f * AQ + (DIM == 3 ? (rr / NB) * PS + (rr % NB) * LS : rr * LS);
            if (active)
            {
                loadLine< NQ, NQ >(p, 1, v, NQ);
                interpolateT< NB, NQ >(v, o, tab.interp);
                storeLine< NQ >(p, 1, o, NB);
            }
        }
        __syncthreads();
    }

    // ---- scatter (MatrixFreeSystem.hpp:494-537): relaxed fp64 atomics, Dirichlet rows skipped; one lane per (node, unknown)
    if (active)
        for (int i = t; i < NN * U; i += TPE)
        {
            const int       a = i / U, u = i % U;
            const long long node = el_nodes[a];
            const int       pos  = DIM == 3 ? (a / (NB * NB)) * PS + ((a / NB) % NB) * LS + a % NB : (a / NB) * LS + a % NB;
            const long long dof  = node * args.dofs_per_node + args.dof_inds[u];
            if (isDirichlet(args.dir_mask, dof))
                continue;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                atomicAdd(args.y + dof + r * args.ld, args.alpha * s_val[(r * U + u) * AQ + pos]);
        }
}
} // namespace l3b
#endif
