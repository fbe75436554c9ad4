// Static condensation, CondensationPolicy::ElementBoundary (algsys/StaticCondensationManager.hpp:135-535), on the device.
//
// The reference keeps, per element, the interior x interior block K_ii, the boundary x interior block K_pi and f_i, scatters K_pp as it
// assembles, and at endAssembly adds -K_pi K_ii^-1 K_ip and -K_pi K_ii^-1 f_i to the global system of the primary (element-boundary)
// dofs (:330-353); after the solve the interior values follow from x_i = K_ii^-1 (f_i - K_ip x_p) (:420-535).
// Here the element matrices are assembled by the same fused kernels into a block-diagonal CRS — the mesh with every element given its
// own node ids, so a "row" of the CRS is a row of K_e and nothing is shared — and one CTA per element then forms the Schur complement
// S_e = K_pp - K_pi K_ii^-1 K_ip straight out of that storage and adds it, with the condensed rhs, to the CRS of the primary nodes.
// K_ii^-1 overwrites K_ii in place: the recovery kernel needs nothing else.
//
// Storage of element e (device_common.cuh, column-dof-major rows of a graph where node (e, a) neighbours (e, 0..NN-1)):
//   K_e[(a, u)][(b, v)] = vals[e NN^2 U^2 + ((a U + u) U + v) NN + b],   F_e[(a, u)][r] = rhs[(e NN + a) U + u + r ld]
#ifndef L3B_CONDENSE_CUH
#define L3B_CONDENSE_CUH

#include "device_common.cuh"

namespace l3b
{
constexpr int cond_threads = 256, cond_rows = 16; // boundary rows per pass of the Schur update

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
    double*          work; // per CTA: nId^2 doubles when K_ii does not fit shared memory, else null
    int*             status;
};

__device__ __forceinline__ long long keIndex(const CondArgs& c, long long e, int a, int u, int b, int v)
{
    return e * c.NN * c.NN * c.U * c.U + ((static_cast< long long >(a) * c.U + u) * c.U + v) * c.NN + b;
}

// one CTA per element. Shared memory: M (nId x nId, or in `work`) | colv (nId) | rowv (nId) | t (cond_rows x nId) | g (nId x n_rhs)
__global__ void __launch_bounds__(cond_threads) condenseKernel(const __grid_constant__ CondArgs c)
{
    extern __shared__ double smem[];
    const long long e   = blockIdx.x;
    const int       tid = threadIdx.x, T = cond_threads;
    const int       U = c.U, nId = c.nI * U, nPd = c.nB * U;
    double*         M    = c.work ? c.work + static_cast< long long >(blockIdx.x) * nId * nId : smem;
    double*         colv = c.work ? smem : smem + nId * nId;
    double*         rowv = colv + nId;
    double*         t    = rowv + nId;
    double*         g    = t + cond_rows * nId;
    const auto      intNode = [&](int i) { return c.int_idx[i / U]; };
    const auto      bndNode = [&](int p) { return c.bnd_idx[p / U]; };

    // ---- K_ii, then its inverse in place (Gauss-Jordan without pivoting: K_ii is symmetric positive definite)
    for (int idx = tid; idx < nId * nId; idx += T)
    {
        const int i = idx / nId, j = idx % nId;
        M[idx]      = c.ke[keIndex(c, e, intNode(i), i % U, intNode(j), j % U)];
    }
    __syncthreads();
    for (int k = 0; k < nId; ++k)
    {
        for (int i = tid; i < nId; i += T)
        {
            colv[i] = M[i * nId + k];
            rowv[i] = M[k * nId + i];
        }
        __syncthreads();
        const double pkk = colv[k];
        if (not(pkk > 0.) and tid == 0)
            atomicOr(c.status, status_degenerate_element);
        const double piv = 1. / pkk;
        for (int idx = tid; idx < nId * nId; idx += T)
        {
            const int i = idx / nId, j = idx % nId;
            double    v;
            if (i == k)
                v = j == k ? piv : rowv[j] * piv;
            else if (j == k)
                v = -colv[i] * piv;
            else
                v = fma(-colv[i] * piv, rowv[j], M[idx]);
            M[idx] = v;
        }
        __syncthreads();
    }
    // g = K_ii^-1 f_i
    for (int idx = tid; idx < nId * c.n_rhs; idx += T)
    {
        const int i = idx % nId, r = idx / nId;
        double    acc = 0.;
        for (int j = 0; j < nId; ++j)
            acc = fma(M[i * nId + j], c.fe[(e * c.NN + intNode(j)) * U + j % U + r * c.ld_e], acc);
        g[idx] = acc;
    }
    __syncthreads();

    // ---- Schur complement, cond_rows boundary rows at a time: t = K_pi[rows] K_ii^-1, S[rows][p'] = K_pp - t K_ip
    const uint16_t* pos_e  = c.pos + e * c.nB * c.nB;
    const uint32_t* prim_e = c.elem_prim + e * c.nB;
    for (int p0 = 0; p0 < nPd; p0 += cond_rows)
    {
        const int nr = min(cond_rows, nPd - p0);
        for (int idx = tid; idx < nr * nId; idx += T)
        {
            const int       r = idx / nId, i2 = idx % nId, p = p0 + r;
            const long long row = keIndex(c, e, bndNode(p), p % U, 0, 0);
            double          acc = 0.;
            for (int i = 0; i < nId; ++i)
                acc = fma(c.ke[row + (i % U) * c.NN + intNode(i)], M[i * nId + i2], acc);
            t[r * nId + i2] = acc;
        }
        __syncthreads();
        // condensed rhs of these rows: f_p - K_pi K_ii^-1 f_i = f_p - K_pi g
        for (int idx = tid; idx < nr * c.n_rhs; idx += T)
        {
            const int       r = idx % nr, col = idx / nr, p = p0 + r;
            const long long row = keIndex(c, e, bndNode(p), p % U, 0, 0);
            double          acc = c.fe[(e * c.NN + bndNode(p)) * U + p % U + col * c.ld_e];
            for (int i = 0; i < nId; ++i)
                acc = fma(-c.ke[row + (i % U) * c.NN + intNode(i)], g[i + col * nId], acc);
            atomicAdd(c.rhs + static_cast< long long >(prim_e[p / U]) * U + p % U + col * c.ld_c, acc);
        }
        for (int q = tid; q < nPd; q += T) // columns enumerated dof-major: consecutive threads read consecutive doubles of K_ip
        {
            const int ib2 = q % c.nB, v2 = q / c.nB, b2 = c.bnd_idx[ib2];
            double    acc[cond_rows];
#pragma unroll
            for (int r = 0; r < cond_rows; ++r)
                acc[r] = 0.;
            for (int i = 0; i < nId; ++i)
            {
                const double kip = c.ke[keIndex(c, e, intNode(i), i % U, b2, v2)]; // K_ip[i][p2]
#pragma unroll
                for (int r = 0; r < cond_rows; ++r)
                    acc[r] = fma(t[r * nId + i], kip, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < cond_rows; ++r)
                if (r < nr)
                {
                    const int       p = p0 + r, ib = p / U, u = p % U;
                    const double    s = c.ke[keIndex(c, e, bndNode(p), u, b2, v2)] - acc[r];
                    const long long A = prim_e[ib], np = c.node_ptr[A], deg = c.node_ptr[A + 1] - np;
                    atomicAdd(c.vals + U * (U * np + u * deg) + v2 * deg + pos_e[ib * c.nB + ib2], s);
                }
        }
        __syncthreads();
    }
    // ---- keep K_ii^-1 where K_ii was: all the recovery needs
    for (int idx = tid; idx < nId * nId; idx += T)
    {
        const int i = idx / nId, j = idx % nId;
        c.ke[keIndex(c, e, intNode(i), i % U, intNode(j), j % U)] = M[idx];
    }
}

// x_i = K_ii^-1 (f_i - K_ip x_p) per element (StaticCondensationManager.hpp:420-535); out: nodal solution over the mesh's nodes,
// out[node * U + u + r * ld_out]; primary values are copied from the condensed solution. One CTA per element.
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
    extern __shared__ double smem[];
    const long long e   = blockIdx.x;
    const int       tid = threadIdx.x, T = cond_threads, U = c.U, nId = c.nI * U, nPd = c.nB * U;
    double*         xp = smem;       // [nPd]
    double*         sv = xp + nPd;   // [nId]
    const auto      ke = [&](int a, int u, int b, int v) {
        return c.ke[e * c.NN * c.NN * U * U + ((static_cast< long long >(a) * U + u) * U + v) * c.NN + b];
    };
    for (int r = 0; r < c.n_rhs; ++r)
    {
        for (int p = tid; p < nPd; p += T)
        {
            const double v = c.x_c[static_cast< long long >(c.elem_prim[e * c.nB + p / U]) * U + p % U + r * c.ld_c];
            xp[p]          = v;
            c.out[static_cast< long long >(c.elem_nodes[e * c.NN + c.bnd_idx[p / U]]) * U + p % U + r * c.ld_out] = v;
        }
        __syncthreads();
        for (int i = tid; i < nId; i += T)
        {
            const int a = c.int_idx[i / U], u = i % U;
            double    acc = c.fe[(e * c.NN + a) * U + u + r * c.ld_e];
            for (int p = 0; p < nPd; ++p)
                acc = fma(-ke(a, u, c.bnd_idx[p / U], p % U), xp[p], acc);
            sv[i] = acc;
        }
        __syncthreads();
        for (int i = tid; i < nId; i += T)
        {
            const int a = c.int_idx[i / U], u = i % U;
            double    acc = 0.;
            for (int j = 0; j < nId; ++j)
                acc = fma(ke(a, u, c.int_idx[j / U], j % U), sv[j], acc); // K_ii^-1, stored in place by condenseKernel
            c.out[static_cast< long long >(c.elem_nodes[e * c.NN + a]) * U + u + r * c.ld_out] = acc;
        }
        __syncthreads();
    }
}
} // namespace l3b
#endif
