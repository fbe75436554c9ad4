// Static condensation, CondensationPolicy::ElementBoundary (algsys/StaticCondensationManager.hpp:135-535), on the device.
//
// The reference keeps, per element, the interior x interior block K_ii, the boundary x interior block K_pi and f_i, scatters K_pp as it
// assembles, and at endAssembly adds -K_pi K_ii^-1 K_ip and -K_pi K_ii^-1 f_i to the global system of the primary (element-boundary)
// dofs (:330-353); after the solve the interior values follow from x_i = K_ii^-1 (f_i - K_ip x_p) (:420-535).
// Here the element matrices are assembled by the same fused kernels into a block-diagonal CRS — the mesh with every element given its
// own node ids, so a "row" of the CRS is a row of K_e and nothing is shared — and one CTA per element then forms the Schur complement
// S_e = K_pp - K_pi K_ii^-1 K_ip straight out of that storage and adds it, with the condensed rhs, to the CRS of the primary nodes.
// W = K_ii^-1 K_ip and g = K_ii^-1 f_i overwrite K_ip and f_i in place: the recovery x_i = g - W x_p needs nothing else.
//
// Storage of element e (device_common.cuh, column-dof-major rows of a graph where node (e, a) neighbours (e, 0..NN-1)):
//   K_e[(a, u)][(b, v)] = vals[e NN^2 U^2 + ((a U + u) U + v) NN + b],   F_e[(a, u)][r] = rhs[(e NN + a) U + u + r ld]
#ifndef L3B_CONDENSE_CUH
#define L3B_CONDENSE_CUH

#include "device_common.cuh"

namespace l3b
{
constexpr int cond_threads = 256;
constexpr int cond_wcols = 16;               // columns of K_ip per pass of W = K_ii^-1 K_ip
constexpr int cond_tm = 64, cond_tn = 64;    // tile of the Schur update S = K_pp - K_pi W (two passes of 2 x 4 outputs per thread)

struct CondArgs
{
    // element-local storage (block-diagonal system)
    double*    ke;     // values
    double*    fe;     // rhs, leading dimension ld_e
    long long  ld_e;
    int        NN, U, n_rhs;
    int        nB, nI;        // boundary / interior nodes per element
    const int* bnd_idx;       // [nB] local node index of boundary node ib
    const int* int_idx;       // [nI]
    // condensed system over the primary nodes
    const uint32_t*  elem_prim; // [e][nB] primary node id
    const uint16_t*  pos;       // [e][ib][ib'] position of ib' in the row of ib
    const long long* node_ptr;
    double*          vals;
    double*          rhs;
    long long        ld_c;
    double*          work; // per CTA: nId x ldM doubles when K_ii does not fit shared memory, else null
    int*             status;
};

__host__ __device__ inline int condLdM(int nId)
{
    return (nId + 3) & ~3; // rows of K_ii^-1 padded to 32 bytes: 4 entries per vector load
}
// shared memory in doubles: phases 1-2 [M (unless in `work`) | colv | rowv | g | K_ip tile], phase 3 [K_pi tile | W tile] over the same bytes
__host__ __device__ inline size_t condSmemDoubles(int nId, int n_rhs, bool m_in_smem)
{
    const size_t ldM = condLdM(nId);
    const size_t p12 = (m_in_smem ? nId * ldM : 0) + 2 * ldM + ldM * n_rhs + static_cast< size_t >(nId) * cond_wcols;
    const size_t p3  = static_cast< size_t >(cond_tm) * (nId + 1) + static_cast< size_t >(nId) * cond_tn;
    return ((p12 > p3 ? p12 : p3) + 1) & ~size_t{1};
}
// plus the four offset tables (ints) behind the doubles
inline size_t condSmemBytes(int nId, int nPd, int n_rhs, bool m_in_smem)
{
    return condSmemDoubles(nId, n_rhs, m_in_smem) * sizeof(double) + 2 * static_cast< size_t >(nId + nPd) * sizeof(int);
}

__device__ __forceinline__ long long keIndex(const CondArgs& c, long long e, int a, int u, int b, int v)
{
    return e * c.NN * c.NN * c.U * c.U + ((static_cast< long long >(a) * c.U + u) * c.U + v) * c.NN + b;
}

// One CTA per element. Interior dofs i = (interior node i / U, unknown i % U); primary dofs are enumerated dof-major, q = v nB + ib,
// so that consecutive q are consecutive doubles of a row of K_e. After the kernel the element-local storage holds W = K_ii^-1 K_ip in
// place of K_ip and g = K_ii^-1 f_i in place of f_i: x_i = g - W x_p is all the recovery needs.
// MAXC: 32-column groups of K_ii a lane covers in the inverse (the host picks the smallest of 1, 2, 4, 8 with 32 MAXC >= interior dofs)
template < int MAXC >
__global__ void __launch_bounds__(cond_threads, 2) condenseKernel(const __grid_constant__ CondArgs c)
{
    extern __shared__ __align__(16) double smem[];
    const long long e   = blockIdx.x;
    const int       tid = threadIdx.x, T = cond_threads, lane = tid & 31, warp = tid >> 5, n_warps = T / 32;
    const int       U = c.U, nId = c.nI * U, nPd = c.nB * U, ldM = condLdM(nId);
    // offsets inside the element's block of K_e[(a, u)][(b, v)] = ke[((a U + u) U + v) NN + b]: a dof as a row, a dof as a column
    int* const iRow = reinterpret_cast< int* >(smem + condSmemDoubles(nId, c.n_rhs, c.work == nullptr));
    int* const iCol = iRow + nId;
    int* const pRow = iCol + nId;
    int* const pCol = pRow + nPd;
    for (int i = tid; i < nId; i += T)
    {
        const int a = c.int_idx[i / U], u = i % U;
        iRow[i]     = (a * U + u) * U * c.NN;
        iCol[i]     = u * c.NN + a;
    }
    for (int q = tid; q < nPd; q += T)
    {
        const int b = c.bnd_idx[q % c.nB], v = q / c.nB;
        pRow[q]     = (b * U + v) * U * c.NN;
        pCol[q]     = v * c.NN + b;
    }
    double* const       ke   = c.ke + e * c.NN * c.NN * U * U;
    double* const       fe   = c.fe + e * c.NN * U; // + (a U + u) + r ld_e
    double*             M    = c.work ? c.work + static_cast< long long >(blockIdx.x) * nId * ldM : smem;
    double*             colv = c.work ? smem : smem + nId * ldM;
    double*             rowv = colv + ldM;
    double*             g    = rowv + ldM;
    double*             Ks   = g + ldM * c.n_rhs; // [nId][cond_wcols]
    const uint32_t*     prim_e = c.elem_prim + e * c.nB;
    const auto          feOff  = [&](int row_off) { return row_off / (U * c.NN); }; // (a U + u) from a row offset
    __syncthreads();

    // ---- 1: K_ii, then its inverse in place (Gauss-Jordan without pivoting: K_ii is symmetric positive definite). A warp per row.
    for (int i = warp; i < nId; i += n_warps)
        for (int j = lane; j < ldM; j += 32)
            M[i * ldM + j] = j < nId ? ke[iRow[i] + iCol[j]] : 0.;
    __syncthreads();
    for (int k = 0; k < nId; ++k)
    {
        for (int i = tid; i < nId; i += T)
        {
            colv[i] = M[i * ldM + k];
            rowv[i] = M[k * ldM + i];
        }
        __syncthreads();
        const double pkk = colv[k];
        if (not(pkk > 0.) and tid == 0)
            atomicOr(c.status, status_singular_interior);
        const double piv = 1. / pkk;
        // a warp per row, a lane per column: the pivot row stays in registers over the rows
        double        rv[MAXC];
#pragma unroll
        for (int t = 0; t < MAXC; ++t)
            rv[t] = lane + 32 * t < nId ? rowv[lane + 32 * t] : 0.;
        for (int i = warp; i < nId; i += n_warps)
        {
            double* const Mi = M + i * ldM;
            if (i == k)
            {
#pragma unroll
                for (int t = 0; t < MAXC; ++t)
                    if (lane + 32 * t < nId)
                        Mi[lane + 32 * t] = lane + 32 * t == k ? piv : rv[t] * piv;
            }
            else
            {
                const double f = -colv[i] * piv;
#pragma unroll
                for (int t = 0; t < MAXC; ++t)
                    if (lane + 32 * t < nId)
                        Mi[lane + 32 * t] = lane + 32 * t == k ? f : fma(f, rv[t], Mi[lane + 32 * t]);
            }
        }
        __syncthreads();
    }
    // g = K_ii^-1 f_i, kept in the element-local rhs in place of f_i
    for (int idx = tid; idx < nId * c.n_rhs; idx += T)
    {
        const int i = idx % nId, r = idx / nId;
        double    acc = 0.;
        for (int j = 0; j < nId; ++j)
            acc = fma(M[i * ldM + j], fe[feOff(iRow[j]) + r * c.ld_e], acc);
        g[i + r * ldM] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < nId * c.n_rhs; idx += T)
    {
        const int i = idx % nId, r = idx / nId;
        fe[feOff(iRow[i]) + r * c.ld_e] = g[i + r * ldM];
    }

    // ---- 2: W = K_ii^-1 K_ip, cond_wcols columns at a time, written over K_ip. A thread owns one column and four rows of W; the inverse
    // is read by rows of four through its symmetry (W[i][q] = sum_j M[j][i] K_ip[j][q])
    {
        const int ql = tid % cond_wcols, grp = tid / cond_wcols, n_grp = T / cond_wcols, n_blk = ldM / 4;
        for (int q0 = 0; q0 < nPd; q0 += cond_wcols)
        {
            const int q = q0 + ql, qc = q < nPd ? pCol[q] : 0;
            for (int j = grp; j < nId; j += n_grp)
                Ks[j * cond_wcols + ql] = q < nPd ? ke[iRow[j] + qc] : 0.;
            __syncthreads();
            for (int blk = grp; blk < n_blk; blk += n_grp)
            {
                double acc[4] = {0., 0., 0., 0.};
#pragma unroll 4
                for (int j = 0; j < nId; ++j)
                {
                    const double  kv = Ks[j * cond_wcols + ql];
                    const double2 m0 = *reinterpret_cast< const double2* >(M + j * ldM + 4 * blk);
                    const double2 m1 = *reinterpret_cast< const double2* >(M + j * ldM + 4 * blk + 2);
                    acc[0] = fma(m0.x, kv, acc[0]);
                    acc[1] = fma(m0.y, kv, acc[1]);
                    acc[2] = fma(m1.x, kv, acc[2]);
                    acc[3] = fma(m1.y, kv, acc[3]);
                }
                if (q < nPd)
                    for (int r = 0; r < 4; ++r)
                        if (4 * blk + r < nId)
                            ke[iRow[4 * blk + r] + qc] = acc[r];
            }
            __syncthreads();
        }
    }

    // ---- 3: S = K_pp - K_pi W in 64 x 64 tiles from shared memory (two passes of 2 x 4 outputs per thread: each W tile loaded from global
    // serves 64 rows, the accumulators stay at 8 registers' worth), straight into the condensed CRS; the
    // condensed rhs f_p - K_pi g rides on the K_pi tile
    double* As = smem;                          // [cond_tm][nId + 1]: K_pi rows of the tile
    double* Bs = smem + cond_tm * (nId + 1);    // [nId][cond_tn]: W columns of the tile
    const int       tx = tid % 16, ty = tid / 16, ldA = nId + 1;
    const uint16_t* pos_e = c.pos + e * c.nB * c.nB;
    for (int p0 = 0; p0 < nPd; p0 += cond_tm)
    {
        __syncthreads();
        for (int r = warp; r < cond_tm; r += n_warps)
        {
            const int p = p0 + r, ro = p < nPd ? pRow[p] : 0;
            for (int i = lane; i < nId; i += 32)
                As[r * ldA + i] = p < nPd ? ke[ro + iCol[i]] : 0.;
        }
        __syncthreads();
        for (int idx = tid; idx < cond_tm * c.n_rhs; idx += T)
        {
            const int r = idx % cond_tm, col = idx / cond_tm, p = p0 + r;
            if (p < nPd)
            {
                double acc = fe[feOff(pRow[p]) + col * c.ld_e];
                for (int i = 0; i < nId; ++i)
                    acc = fma(-As[r * ldA + i], fe[feOff(iRow[i]) + col * c.ld_e], acc);
                atomicAdd(c.rhs + static_cast< long long >(prim_e[p % c.nB]) * U + p / c.nB + col * c.ld_c, acc);
            }
        }
        // W tiles: the loads of the next tile are issued before the current one is contracted (registers, then shared memory), so
        // their DRAM latency hides behind the multiply-adds (the element's W is re-read once per row tile: 13 x 338 kB at hex p=4)
        constexpr int PF = 28; // values of a tile per thread: nId / (T / cond_tn) <= PF, else the tile is loaded without prefetch
        const bool    use_pf = (nId + T / cond_tn - 1) / (T / cond_tn) <= PF;
        const int     bql = tid % cond_tn, bi0 = tid / cond_tn;
        double        pf[PF];
        const auto    loadTile = [&](int q0) {
            const int q = q0 + bql, qc = q < nPd ? pCol[q] : 0;
#pragma unroll
            for (int k = 0; k < PF; ++k)
            {
                const int i = bi0 + k * (T / cond_tn);
                pf[k]       = i < nId and q < nPd ? ke[iRow[i] + qc] : 0.;
            }
        };
        // S is symmetric: only the column tiles that reach the diagonal of this row tile are formed, entries right of the diagonal are
        // scattered twice (rows and columns share the dof-major enumeration)
        const int q_first = (p0 / cond_tn) * cond_tn;
        if (use_pf)
            loadTile(q_first);
        for (int q0 = q_first; q0 < nPd; q0 += cond_tn)
        {
            __syncthreads();
            if (use_pf)
            {
#pragma unroll
                for (int k = 0; k < PF; ++k)
                {
                    const int i = bi0 + k * (T / cond_tn);
                    if (i < nId)
                        Bs[i * cond_tn + bql] = pf[k];
                }
            }
            else
            {
                const int q = q0 + bql, qc = q < nPd ? pCol[q] : 0;
#pragma unroll 4
                for (int i = bi0; i < nId; i += T / cond_tn)
                    Bs[i * cond_tn + bql] = q < nPd ? ke[iRow[i] + qc] : 0.;
            }
            __syncthreads();
            if (use_pf and q0 + cond_tn < nPd)
                loadTile(q0 + cond_tn);
            // thread tile: rows 2 ty, 2 ty + 1 of each 32-row half; columns 2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx (16-byte stride between threads)
            for (int half = 0; half < cond_tm / 32; ++half)
            {
                const int r0 = 32 * half + 2 * ty;
                if (p0 + 32 * half >= nPd or q0 + cond_tn <= p0 + 32 * half) // rows past the end, or a tile wholly left of this half's diagonal
                    continue;
                double acc[2][4] = {{0., 0., 0., 0.}, {0., 0., 0., 0.}};
#pragma unroll 2
                for (int i = 0; i < nId; ++i)
                {
                    const double  a0 = As[r0 * ldA + i], a1 = As[(r0 + 1) * ldA + i];
                    const double2 b0 = *reinterpret_cast< const double2* >(Bs + i * cond_tn + 2 * tx);
                    const double2 b1 = *reinterpret_cast< const double2* >(Bs + i * cond_tn + cond_tn / 2 + 2 * tx);
                    acc[0][0] = fma(a0, b0.x, acc[0][0]);
                    acc[0][1] = fma(a0, b0.y, acc[0][1]);
                    acc[0][2] = fma(a0, b1.x, acc[0][2]);
                    acc[0][3] = fma(a0, b1.y, acc[0][3]);
                    acc[1][0] = fma(a1, b0.x, acc[1][0]);
                    acc[1][1] = fma(a1, b0.y, acc[1][1]);
                    acc[1][2] = fma(a1, b1.x, acc[1][2]);
                    acc[1][3] = fma(a1, b1.y, acc[1][3]);
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                {
                    const int p = p0 + r0 + rr;
                    if (p >= nPd)
                        continue;
                    const int       ib = p % c.nB, u = p / c.nB, ro = pRow[p];
                    const long long A = prim_e[ib], np = c.node_ptr[A], deg = c.node_ptr[A + 1] - np;
                    double* const   rowp = c.vals + U * (U * np + u * deg);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                    {
                        const int q = q0 + (cc / 2) * (cond_tn / 2) + 2 * tx + cc % 2;
                        if (q >= nPd or q < p)
                            continue;
                        const int    ib2 = q % c.nB, v2 = q / c.nB;
                        const double sv  = ke[ro + pCol[q]] - acc[rr][cc];
                        atomicAdd(rowp + v2 * deg + pos_e[ib * c.nB + ib2], sv);
                        if (q > p)
                        {
                            const long long A2 = prim_e[ib2], np2 = c.node_ptr[A2], deg2 = c.node_ptr[A2 + 1] - np2;
                            atomicAdd(c.vals + U * (U * np2 + v2 * deg2) + u * deg2 + pos_e[ib2 * c.nB + ib], sv);
                        }
                    }
                }
            }
        }
    }
}


// ---- elements with MANY interior dofs (more than 256: hex p >= 6 at U = 4, the configuration benchmarks/Diffusion3DBenchmark.cpp ships) --------
// Same three steps, blocked so that nothing of size nId^2 or nId x tile has to fit shared memory: K_ii^-1 lives in the global work
// buffer (2 MB per resident element at nId = 500: L2-resident), the pivot row and column of each Gauss-Jordan step and the K_ip / K_pi /
// W tiles go through shared memory in chunks of cond_kc interior dofs.
constexpr int cond_kc = 64;                  // interior dofs per chunk of the tile contractions
constexpr int cond_lt = 32;                  // tile edge of the Schur update in the large kernel
inline size_t condLargeSmemBytes(int nId, int nPd, int n_rhs)
{
    const size_t ldM = condLdM(nId);
    const size_t p12 = 2 * ldM + ldM * n_rhs + static_cast< size_t >(nId) * cond_wcols;
    const size_t p3  = static_cast< size_t >(cond_lt) * (cond_kc + 1) + static_cast< size_t >(cond_kc) * cond_lt;
    const size_t dbl = ((p12 > p3 ? p12 : p3) + 1) & ~size_t{1};
    return dbl * sizeof(double) + 2 * static_cast< size_t >(nId + nPd) * sizeof(int);
}
__global__ void __launch_bounds__(cond_threads) condenseLargeKernel(const __grid_constant__ CondArgs c)
{
    extern __shared__ __align__(16) double smem[];
    const long long e   = blockIdx.x;
    const int       tid = threadIdx.x, T = cond_threads, lane = tid & 31, warp = tid >> 5, n_warps = T / 32;
    const int       U = c.U, nId = c.nI * U, nPd = c.nB * U, ldM = condLdM(nId);
    const size_t    p12 = 2 * static_cast< size_t >(ldM) + static_cast< size_t >(ldM) * c.n_rhs + static_cast< size_t >(nId) * cond_wcols;
    const size_t    p3  = static_cast< size_t >(cond_lt) * (cond_kc + 1) + static_cast< size_t >(cond_kc) * cond_lt;
    int* const      iRow = reinterpret_cast< int* >(smem + (((p12 > p3 ? p12 : p3) + 1) & ~size_t{1}));
    int* const      iCol = iRow + nId;
    int* const      pRow = iCol + nId;
    int* const      pCol = pRow + nPd;
    for (int i = tid; i < nId; i += T)
    {
        const int a = c.int_idx[i / U], u = i % U;
        iRow[i]     = (a * U + u) * U * c.NN;
        iCol[i]     = u * c.NN + a;
    }
    for (int q = tid; q < nPd; q += T)
    {
        const int b = c.bnd_idx[q % c.nB], v = q / c.nB;
        pRow[q]     = (b * U + v) * U * c.NN;
        pCol[q]     = v * c.NN + b;
    }
    double* const   ke     = c.ke + e * c.NN * c.NN * U * U;
    double* const   fe     = c.fe + e * c.NN * U;
    double* const   M      = c.work + static_cast< long long >(blockIdx.x) * nId * ldM;
    double* const   colv   = smem;
    double* const   rowv   = colv + ldM;
    double* const   g      = rowv + ldM;
    double* const   Ks     = g + ldM * c.n_rhs;
    const uint32_t* prim_e = c.elem_prim + e * c.nB;
    const auto      feOff  = [&](int row_off) { return row_off / (U * c.NN); };
    __syncthreads();
    // ---- 1: K_ii^-1 in place (Gauss-Jordan, no pivoting: symmetric positive definite)
    for (int i = warp; i < nId; i += n_warps)
        for (int j = lane; j < ldM; j += 32)
            M[i * ldM + j] = j < nId ? ke[iRow[i] + iCol[j]] : 0.;
    __syncthreads();
    for (int k = 0; k < nId; ++k)
    {
        for (int i = tid; i < nId; i += T)
        {
            colv[i] = M[i * ldM + k];
            rowv[i] = M[k * ldM + i];
        }
        __syncthreads();
        const double pkk = colv[k];
        if (not(pkk > 0.) and tid == 0)
            atomicOr(c.status, status_singular_interior);
        const double piv = 1. / pkk;
        for (int i = warp; i < nId; i += n_warps)
        {
            double* const Mi = M + i * ldM;
            const double  f  = i == k ? 0. : -colv[i] * piv;
            for (int j = lane; j < nId; j += 32)
                Mi[j] = i == k ? (j == k ? piv : rowv[j] * piv) : (j == k ? f : fma(f, rowv[j], Mi[j]));
        }
        __syncthreads();
    }
    // g = K_ii^-1 f_i over f_i
    for (int idx = tid; idx < nId * c.n_rhs; idx += T)
    {
        const int i = idx % nId, r = idx / nId;
        double    acc = 0.;
        for (int j = 0; j < nId; ++j)
            acc = fma(M[i * ldM + j], fe[feOff(iRow[j]) + r * c.ld_e], acc);
        g[i + r * ldM] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < nId * c.n_rhs; idx += T)
    {
        const int i = idx % nId, r = idx / nId;
        fe[feOff(iRow[i]) + r * c.ld_e] = g[i + r * ldM];
    }
    // ---- 2: W = K_ii^-1 K_ip over K_ip, cond_wcols columns at a time (a thread: one column, rows strided over the thread groups)
    {
        const int ql = tid % cond_wcols, grp = tid / cond_wcols, n_grp = T / cond_wcols;
        for (int q0 = 0; q0 < nPd; q0 += cond_wcols)
        {
            const int q = q0 + ql, qc = q < nPd ? pCol[q] : 0;
            __syncthreads();
            for (int j = grp; j < nId; j += n_grp)
                Ks[j * cond_wcols + ql] = q < nPd ? ke[iRow[j] + qc] : 0.;
            __syncthreads();
            for (int i = grp; i < nId; i += n_grp)
            {
                double acc = 0.;
                for (int j = 0; j < nId; ++j)
                    acc = fma(M[j * ldM + i], Ks[j * cond_wcols + ql], acc); // symmetric inverse: column i read as row-strided
                if (q < nPd)
                    ke[iRow[i] + qc] = acc;
            }
        }
    }
    __syncthreads();
    // condensed rhs: f_p - K_pi g
    for (int idx = tid; idx < nPd * c.n_rhs; idx += T)
    {
        const int p = idx % nPd, col = idx / nPd;
        double    acc = fe[feOff(pRow[p]) + col * c.ld_e];
        for (int i = 0; i < nId; ++i)
            acc = fma(-ke[pRow[p] + iCol[i]], fe[feOff(iRow[i]) + col * c.ld_e], acc);
        atomicAdd(c.rhs + static_cast< long long >(prim_e[p % c.nB]) * U + p / c.nB + col * c.ld_c, acc);
    }
    // ---- 3: S = K_pp - K_pi W, 32 x 32 tiles, interior dofs in chunks of cond_kc; upper triangle formed, mirrored on scatter
    double* const   As    = smem;                            // [cond_lt][cond_kc + 1]
    double* const   Bs    = smem + cond_lt * (cond_kc + 1);  // [cond_kc][cond_lt]
    const int       tx = tid % 16, ty = tid / 16;            // thread: rows 2 ty, 2 ty + 1; columns 2 tx, 2 tx + 1
    const uint16_t* pos_e = c.pos + e * c.nB * c.nB;
    for (int p0 = 0; p0 < nPd; p0 += cond_lt)
        for (int q0 = p0; q0 < nPd; q0 += cond_lt)
        {
            double acc[2][2] = {{0., 0.}, {0., 0.}};
            for (int k0 = 0; k0 < nId; k0 += cond_kc)
            {
                __syncthreads();
                for (int idx = tid; idx < cond_lt * cond_kc; idx += T)
                {
                    const int r = idx / cond_kc, k = idx % cond_kc, p = p0 + r, i = k0 + k;
                    As[r * (cond_kc + 1) + k] = p < nPd and i < nId ? ke[pRow[p] + iCol[i]] : 0.;
                }
                for (int idx = tid; idx < cond_kc * cond_lt; idx += T)
                {
                    const int k = idx / cond_lt, cq = idx % cond_lt, q = q0 + cq, i = k0 + k;
                    Bs[k * cond_lt + cq] = q < nPd and i < nId ? ke[iRow[i] + pCol[q]] : 0.;
                }
                __syncthreads();
#pragma unroll 4
                for (int k = 0; k < cond_kc; ++k)
                {
                    const double a0 = As[(2 * ty) * (cond_kc + 1) + k], a1 = As[(2 * ty + 1) * (cond_kc + 1) + k];
                    const double b0 = Bs[k * cond_lt + 2 * tx], b1 = Bs[k * cond_lt + 2 * tx + 1];
                    acc[0][0] = fma(a0, b0, acc[0][0]);
                    acc[0][1] = fma(a0, b1, acc[0][1]);
                    acc[1][0] = fma(a1, b0, acc[1][0]);
                    acc[1][1] = fma(a1, b1, acc[1][1]);
                }
            }
            for (int rr = 0; rr < 2; ++rr)
            {
                const int p = p0 + 2 * ty + rr;
                if (p >= nPd)
                    continue;
                const int       ib = p % c.nB, u = p / c.nB, ro = pRow[p];
                const long long A = prim_e[ib], np = c.node_ptr[A], deg = c.node_ptr[A + 1] - np;
                double* const   rowp = c.vals + U * (U * np + u * deg);
                for (int cc = 0; cc < 2; ++cc)
                {
                    const int q = q0 + 2 * tx + cc;
                    if (q >= nPd or q < p)
                        continue;
                    const int    ib2 = q % c.nB, v2 = q / c.nB;
                    const double sv  = ke[ro + pCol[q]] - acc[rr][cc];
                    atomicAdd(rowp + v2 * deg + pos_e[ib * c.nB + ib2], sv);
                    if (q > p)
                    {
                        const long long A2 = prim_e[ib2], np2 = c.node_ptr[A2], deg2 = c.node_ptr[A2 + 1] - np2;
                        atomicAdd(c.vals + U * (U * np2 + v2 * deg2) + u * deg2 + pos_e[ib2 * c.nB + ib], sv);
                    }
                }
            }
        }
}

// x_i = K_ii^-1 (f_i - K_ip x_p) = g - W x_p per element (StaticCondensationManager.hpp:420-535), with g and W as condenseKernel left
// them; out: nodal solution over the mesh's nodes, out[node * U + u + r * ld_out]; primary values are copied from the condensed
// solution. One CTA per element.
struct RecoverArgs
{
    const double*   ke;
    const double*   fe;
    long long       ld_e;
    int             NN, U, n_rhs, nB, nI;
    const int*      bnd_idx;
    const int*      int_idx;
    const uint32_t* elem_prim;
    const uint32_t* elem_nodes; // [e][NN] node ids of the mesh
    const double*   x_c;        // condensed solution, ld_c
    long long       ld_c;
    double*         out;
    long long       ld_out;
};
__global__ void __launch_bounds__(cond_threads) recoverKernel(const __grid_constant__ RecoverArgs c)
{
    extern __shared__ __align__(16) double smem[];
    const long long e   = blockIdx.x;
    const int       tid = threadIdx.x, T = cond_threads, U = c.U, nId = c.nI * U, nPd = c.nB * U;
    double*         xp = smem; // [nPd], dof-major: q = v nB + ib
    for (int r = 0; r < c.n_rhs; ++r)
    {
        __syncthreads();
        for (int q = tid; q < nPd; q += T)
        {
            const int    ib = q % c.nB, v = q / c.nB;
            const double x  = c.x_c[static_cast< long long >(c.elem_prim[e * c.nB + ib]) * U + v + r * c.ld_c];
            xp[q]           = x;
            c.out[static_cast< long long >(c.elem_nodes[e * c.NN + c.bnd_idx[ib]]) * U + v + r * c.ld_out] = x;
        }
        __syncthreads();
        for (int i = tid; i < nId; i += T)
        {
            const int       a = c.int_idx[i / U], u = i % U;
            const long long row = e * c.NN * c.NN * U * U + (static_cast< long long >(a) * U + u) * U * c.NN;
            double          acc = c.fe[(e * c.NN + a) * U + u + r * c.ld_e]; // g
            for (int q = 0; q < nPd; ++q)
                acc = fma(-c.ke[row + (q / c.nB) * c.NN + c.bnd_idx[q % c.nB]], xp[q], acc); // W
            c.out[static_cast< long long >(c.elem_nodes[e * c.NN + a]) * U + u + r * c.ld_out] = acc;
        }
    }
}
} // namespace l3b
#endif
