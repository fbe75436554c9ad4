// C ABI implementation (include/l3ster_b200.h): contexts, device meshes, the assembled and the matrix-free system, and the
// Krylov layer. Host orchestration only — all element math lives in the kernels instantiated through register_kernel.cuh.
#include "../../include/l3ster_b200.h"

#include "comm.cuh"
#include "mesh_host.hpp"
#include "partition_host.hpp"
#include "apply_plan_host.hpp"
#include "dofmap_host.hpp"
#include "mf_hex_planes.cuh"
#include "condense.cuh"
#include "registry.hpp"
#include "tables.hpp"

#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace l3b
{
std::vector< KernelEntry >& kernelRegistry()
{
    static std::vector< KernelEntry > registry;
    return registry;
}
int findKernel(const std::string& name)
{
    const auto& r = kernelRegistry();
    for (size_t i = 0; i < r.size(); ++i)
        if (r[i].info.name == name)
            return static_cast< int >(i);
    return -1;
}
} // namespace l3b

using namespace l3b;

namespace
{
std::string g_global_error;

struct Error
{
    int         code;
    std::string msg;
};
[[noreturn]] void fail(int code, std::string msg)
{
    throw Error{code, std::move(msg)};
}
void cudaCheck(cudaError_t err, const char* what)
{
    if (err != cudaSuccess)
        fail(L3B_ERR_CUDA, std::string{what} + ": " + cudaGetErrorString(err));
}

// NVTX ranges named after the reference's Caliper regions (util/Caliper.hpp: L3STER_PROFILE_FUNCTION / _REGION_BEGIN), so that a
// timeline of this library reads like a Caliper report of the reference. NVTX v3 is header-only: no tool attached, no cost.
struct ProfileRegion
{
    explicit ProfileRegion(const char* name) { nvtxRangePushA(name); }
    ~ProfileRegion() { nvtxRangePop(); }
    ProfileRegion(const ProfileRegion&)            = delete;
    ProfileRegion& operator=(const ProfileRegion&) = delete;
};

template < typename T >
struct DevBuf
{
    T*     ptr = nullptr;
    size_t n   = 0;
    DevBuf()   = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : ptr{o.ptr}, n{o.n} { o.ptr = nullptr, o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        release();
        ptr   = o.ptr;
        n     = o.n;
        o.ptr = nullptr;
        o.n   = 0;
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_)
    {
        release();
        n = n_;
        if (n > 0)
            cudaCheck(cudaMalloc(&ptr, n * sizeof(T)), "cudaMalloc");
    }
    void release()
    {
        if (ptr)
            cudaFree(ptr);
        ptr = nullptr;
        n   = 0;
    }
    void upload(const T* src, size_t count, cudaStream_t s)
    {
        if (count > 0)
            cudaCheck(cudaMemcpyAsync(ptr, src, count * sizeof(T), cudaMemcpyHostToDevice, s), "H2D copy");
    }
    void download(T* dst, size_t count, cudaStream_t s) const
    {
        if (count > 0)
            cudaCheck(cudaMemcpyAsync(dst, ptr, count * sizeof(T), cudaMemcpyDeviceToHost, s), "D2H copy");
    }
    void zero(cudaStream_t s)
    {
        if (n > 0)
            cudaCheck(cudaMemsetAsync(ptr, 0, n * sizeof(T), s), "memset");
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// vector kernels
__global__ void scaleKernel(double* y, long long n, double beta)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        y[i] = beta == 0. ? 0. : beta * y[i];
}
// Dirichlet rows are identity rows: y[d] += alpha x[d] (MatrixFreeSystem.hpp:1087-1103)
__global__ void dirichletRowsKernel(const uint8_t* mask, const double* x, double* y, long long n, long long ld, int n_cols, double alpha)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        if (mask[i])
            for (int c = 0; c < n_cols; ++c)
                y[i + c * ld] += alpha * x[i + c * ld];
}
// per element: any Dirichlet dof on any of its nodes? (lets the apply kernels skip the mask for interior elements)
__global__ void elemDirichletFlagKernel(const uint32_t* nodes, long long n_elems, int nn, int dpn, const uint8_t* mask, uint32_t* flag)
{
    for (long long e = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; e < n_elems; e += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        uint32_t f = 0;
        for (int a = 0; a < nn; ++a)
            for (int d = 0; d < dpn; ++d)
                f |= mask[static_cast< long long >(nodes[e * nn + a]) * dpn + d];
        flag[e] = f != 0;
    }
}
// hex geometry record (mf_hex_planes.cuh): monomial coefficients of the trilinear map, and for affine elements the constant
// inverse Jacobian, its determinant and the affine flag
__global__ void hexGeometryKernel(const double* verts, long long n_elems, double* geo)
{
    for (long long e = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; e < n_elems; e += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        double* g = geo + e * hex_geo_doubles;
        for (int i = 0; i < hex_geo_doubles; ++i)
            g[i] = 0.;
        buildGeometryCoefs< 3 >(verts + e * 24, g, 0, 1);
        bool affine = true;
        for (int m = 0; m < 8; ++m)
            if (__popc(m) > 1)
                for (int s = 0; s < 3; ++s)
                    affine = affine and g[m * 3 + s] == 0.;
        g[hex_geo_affine] = affine ? 1. : 0.;
        if (affine)
        {
            double Jt[3][3], Jti[3][3];
            for (int d = 0; d < 3; ++d)
                for (int s = 0; s < 3; ++s)
                    Jt[d][s] = g[(1 << d) * 3 + s];
            g[hex_geo_det] = invert< 3 >(Jt, Jti);
            bool diagonal = true;
            for (int s = 0; s < 3; ++s)
                for (int d = 0; d < 3; ++d)
                {
                    g[hex_geo_jti + s * 3 + d] = Jti[s][d];
                    diagonal                   = diagonal and (s == d or Jti[s][d] == 0.);
                }
            if (diagonal)
                g[hex_geo_affine] = 2.; // axis-aligned box
        }
    }
}
// the same over a compact list of the Dirichlet dofs (O(surface) instead of a pass over the whole vector)
// energy (may be null): x^T A x picks up x_d^2 per owned Dirichlet dof d of column 0 (the identity rows)
__global__ void dirichletRowsListKernel(const int32_t* dofs, long long n_dir, const double* x, double* y, long long ld, int n_cols, double alpha,
                                        double* energy, long long n_owned_dofs)
{
    double e = 0.;
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n_dir; i += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const long long d = dofs[i];
        for (int c = 0; c < n_cols; ++c)
            y[d + c * ld] += alpha * x[d + c * ld];
        if (d < n_owned_dofs)
            e = fma(x[d], x[d], e);
    }
    if (energy != nullptr)
    {
        for (int off = 16; off > 0; off >>= 1)
            e += __shfl_xor_sync(0xffffffffu, e, off);
        if ((threadIdx.x & 31) == 0 and e != 0.)
            atomicAdd(energy, e);
    }
}
// halo packing (comm/ImportExport.hpp): gather / scatter-add through a device index list
__global__ void vecGatherKernel(const double* src, long long ld, const int32_t* idx, long long n, int n_cols, double* dst)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n * n_cols; i += static_cast< long long >(gridDim.x) * blockDim.x)
        dst[i] = src[idx[i % n] + (i / n) * ld];
}
__global__ void vecScatterAddKernel(double* dst, long long ld, const int32_t* idx, long long n, int n_cols, const double* src)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n * n_cols; i += static_cast< long long >(gridDim.x) * blockDim.x)
        dst[idx[i % n] + (i / n) * ld] += src[i];
}
// handle_dirichlet_dof (MatrixFreeSystem.hpp:911-915)
__global__ void dirichletInitKernel(const uint8_t* mask, const double* vals, double* diag, double* rhs, long long n, long long ld, int n_rhs)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        if (mask[i])
        {
            diag[i] = 1.;
            for (int c = 0; c < n_rhs; ++c)
                rhs[i + c * ld] = vals[i + c * ld];
        }
}
// native Jacobi: M^-1 = damping * sign(d) / max(|d|, threshold) (solve/NativePreconditioners.hpp:86-100)
__global__ void jacobiInvertKernel(const double* diag, double* minv, long long n, double damping, double threshold)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const double v = diag[i];
        minv[i]        = (v < 0 ? -damping : damping) / fmax(fabs(v), threshold);
    }
}
__device__ __forceinline__ double blockSum(double v)
{
    __shared__ double s[32];
    for (int off = 16; off > 0; off >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0)
        s[threadIdx.x >> 5] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.;
    if (threadIdx.x < 32)
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    return v;
}
// out[0] += a.b ; out[1] += c.d (either pair may alias); two doubles per access (cudaMalloc'ed arrays), odd tail by one thread
__global__ void dot2Kernel(const double* a, const double* b, const double* c, const double* d, long long n, double* out)
{
    double          s0 = 0., s1 = 0.;
    const long long n2 = n / 2;
    const auto* const a2 = reinterpret_cast< const double2* >(a);
    const auto* const b2 = reinterpret_cast< const double2* >(b);
    const auto* const c2 = reinterpret_cast< const double2* >(c);
    const auto* const d2 = reinterpret_cast< const double2* >(d);
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n2; i += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const double2 u = a2[i], v = b2[i];
        s0 = fma(u.x, v.x, fma(u.y, v.y, s0));
        if (c)
        {
            const double2 w = c2[i], t = d2[i];
            s1 = fma(w.x, t.x, fma(w.y, t.y, s1));
        }
    }
    if ((n & 1) and blockIdx.x == 0 and threadIdx.x == 0)
    {
        s0 = fma(a[n - 1], b[n - 1], s0);
        if (c)
            s1 = fma(c[n - 1], d[n - 1], s1);
    }
    s0 = blockSum(s0);
    s1 = blockSum(s1);
    if (threadIdx.x == 0)
    {
        atomicAdd(out, s0);
        if (c)
            atomicAdd(out + 1, s1);
    }
}
// r -= a Ap ; z = minv r ; out[0] += r.r ; out[1] += r.z     (a = rz / pAp read from device scalars). Two doubles per access: the
// arrays are cudaMalloc'ed (16-byte aligned), an odd last element is handled by one thread.
__global__ void cgUpdateKernel(double* r, double* z, const double* Ap, const double* minv, long long n, const double* scal /* [rz, pAp] */,
                               double* out)
{
    const double    a  = scal[0] / scal[1];
    double          s0 = 0., s1 = 0.;
    const long long n2 = n / 2;
    auto* const       r2  = reinterpret_cast< double2* >(r);
    auto* const       z2  = reinterpret_cast< double2* >(z);
    const auto* const Ap2 = reinterpret_cast< const double2* >(Ap);
    const auto* const m2  = reinterpret_cast< const double2* >(minv);
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n2; i += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const double2 q = Ap2[i], m = m2[i];
        double2       ri = r2[i];
        ri.x = fma(-a, q.x, ri.x);
        ri.y = fma(-a, q.y, ri.y);
        r2[i] = ri;
        const double2 zi{m.x * ri.x, m.y * ri.y};
        z2[i] = zi;
        s0    = fma(ri.x, ri.x, fma(ri.y, ri.y, s0));
        s1    = fma(ri.x, zi.x, fma(ri.y, zi.y, s1));
    }
    if ((n & 1) and blockIdx.x == 0 and threadIdx.x == 0)
    {
        const long long i  = n - 1;
        const double    ri = fma(-a, Ap[i], r[i]);
        r[i]               = ri;
        const double zi    = minv[i] * ri;
        z[i]               = zi;
        s0                 = fma(ri, ri, s0);
        s1                 = fma(ri, zi, s1);
    }
    s0 = blockSum(s0);
    s1 = blockSum(s1);
    if (threadIdx.x == 0)
    {
        atomicAdd(out, s0);
        atomicAdd(out + 1, s1);
    }
}
// x += a p ; p = z + (rz_new / rz_old) p     (the x update rides on the pass that reads p anyway; scal = [rz_old, pAp, rr, rz_new])
__global__ void cgDirectionKernel(double* x, double* p, const double* z, long long n, const double* scal)
{
    const double    a = scal[0] / scal[1], b = scal[3] / scal[0];
    const long long n2 = n / 2;
    auto* const       x2 = reinterpret_cast< double2* >(x);
    auto* const       p2 = reinterpret_cast< double2* >(p);
    const auto* const z2 = reinterpret_cast< const double2* >(z);
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n2; i += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const double2 pi = p2[i], zi = z2[i];
        double2       xi = x2[i];
        xi.x = fma(a, pi.x, xi.x);
        xi.y = fma(a, pi.y, xi.y);
        x2[i] = xi;
        p2[i] = double2{fma(b, pi.x, zi.x), fma(b, pi.y, zi.y)};
    }
    if ((n & 1) and blockIdx.x == 0 and threadIdx.x == 0)
    {
        const long long i  = n - 1;
        const double    pi = p[i];
        x[i]               = fma(a, pi, x[i]);
        p[i]               = fma(b, pi, z[i]);
    }
}
// x += a p     (last iteration: no next direction)
__global__ void cgFinalKernel(double* x, const double* p, long long n, const double* scal)
{
    const double a = scal[0] / scal[1];
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        x[i] = fma(a, p[i], x[i]);
}
__global__ void hadamardKernel(double* z, const double* minv, const double* r, long long n)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        z[i] = minv[i] * r[i];
}
// classical Gram-Schmidt pass of GMRES: out[i] += V_i . w for i < m (V: m vectors of leading dimension ld)
__global__ void gsDotsKernel(const double* V, long long ld, int m, const double* w, long long n, double* out)
{
    for (int i = 0; i < m; ++i)
    {
        const double* v = V + i * ld;
        double        s = 0.;
        for (long long k = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; k < n; k += static_cast< long long >(gridDim.x) * blockDim.x)
            s = fma(v[k], w[k], s);
        s = blockSum(s);
        if (threadIdx.x == 0)
            atomicAdd(out + i, s);
    }
}
// w += sign * sum_{i < m} h[i] V_i
__global__ void gsAxpyKernel(const double* V, long long ld, int m, const double* h, double sign, double* w, long long n)
{
    for (long long k = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; k < n; k += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        double acc = 0.;
        for (int i = 0; i < m; ++i)
            acc = fma(h[i], V[i * ld + k], acc);
        w[k] = fma(sign, acc, w[k]);
    }
}
// w = minv * (b - w)
__global__ void precResidualKernel(double* w, const double* b, const double* minv, long long n)
{
    for (long long k = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; k < n; k += static_cast< long long >(gridDim.x) * blockDim.x)
        w[k] = minv[k] * (b[k] - w[k]);
}
// v = a * w
__global__ void scaledCopyKernel(double* v, const double* w, double a, long long n)
{
    for (long long k = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; k < n; k += static_cast< long long >(gridDim.x) * blockDim.x)
        v[k] = a * w[k];
}
// dof-level CRS helpers on the node-block layout: row (n, d) = for each neighbour m (ascending): dofs_per_node entries
__global__ void rowPtrKernel(const long long* node_ptr, long long n_nodes, int dpn, long long* row_ptr)
{
    const long long row = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x;
    if (row > n_nodes * dpn)
        return;
    if (row == n_nodes * dpn)
    {
        row_ptr[row] = node_ptr[n_nodes] * dpn * dpn;
        return;
    }
    const long long n = row / dpn, d = row % dpn, deg = node_ptr[n + 1] - node_ptr[n];
    row_ptr[row] = dpn * (dpn * node_ptr[n] + d * deg);
}
// slot map from the node graph: pos[e][a][b] = index of node b in the sorted neighbour list of node a
__global__ void slotMapKernel(const uint32_t* nodes, long long n_elems, int nn, const long long* node_ptr, const uint32_t* node_nbr,
                              uint16_t* pos, int* status)
{
    const long long idx = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x;
    if (idx >= n_elems * nn * nn)
        return;
    const long long e    = idx / (static_cast< long long >(nn) * nn);
    const int       a    = static_cast< int >((idx / nn) % nn), b = static_cast< int >(idx % nn);
    const long long na   = nodes[e * nn + a];
    const uint32_t  want = nodes[e * nn + b];
    long long       lo = node_ptr[na], hi = node_ptr[na + 1];
    const long long beg = lo, end = hi;
    while (lo < hi)
    {
        const long long mid = (lo + hi) / 2;
        if (node_nbr[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo >= end or node_nbr[lo] != want)
    {
        atomicOr(status, status_graph_entry_missing);
        return;
    }
    if (lo - beg > 0xFFFF) // more neighbours than the 16-bit slot position can address
    {
        atomicOr(status, status_slot_overflow);
        return;
    }
    pos[idx] = static_cast< uint16_t >(lo - beg);
}
// y = A x, one warp per row, node-block layout
__global__ void spmvKernel(const long long* node_ptr, const uint32_t* node_nbr, const double* vals, long long n_nodes, int dpn, const double* x,
                           double* y)
{
    const long long row  = (blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x) / 32;
    const int       lane = threadIdx.x & 31;
    if (row >= n_nodes * dpn)
        return;
    const long long n = row / dpn, d = row % dpn, deg = node_ptr[n + 1] - node_ptr[n];
    const long long beg = dpn * (dpn * node_ptr[n] + d * deg);
    double          acc = 0.;
    for (long long k = lane; k < deg * dpn; k += 32) // entry k of the row = (neighbour k % deg, column dof k / deg)
        acc = fma(vals[beg + k], x[static_cast< long long >(node_nbr[node_ptr[n] + k % deg]) * dpn + k / deg], acc);
    for (int off = 16; off > 0; off >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0)
        y[row] = acc;
}
// algebraic Dirichlet BCs (bcs/DirichletBC.hpp:82-150), one warp per row
__global__ void dirichletAlgebraicKernel(const long long* node_ptr, const uint32_t* node_nbr, double* vals, long long n_nodes, int dpn,
                                         const uint8_t* is_bc, const double* bc_vals, double* rhs, long long ld, int n_rhs,
                                         long long n_owned_dofs)
{
    const long long row  = (blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x) / 32;
    const int       lane = threadIdx.x & 31;
    if (row >= n_nodes * dpn)
        return;
    const long long n = row / dpn, d = row % dpn, deg = node_ptr[n + 1] - node_ptr[n];
    const long long beg = dpn * (dpn * node_ptr[n] + d * deg);
    if (is_bc[row])
    {
        for (long long k = lane; k < deg * dpn; k += 32)
        {
            const long long col = static_cast< long long >(node_nbr[node_ptr[n] + k % deg]) * dpn + k / deg;
            vals[beg + k]       = col == row and row < n_owned_dofs ? 1. : 0.; // ghost copies of a Dirichlet row stay empty
        }
        if (lane == 0)
            for (int c = 0; c < n_rhs; ++c)
                rhs[row + c * ld] = row < n_owned_dofs ? bc_vals[row + c * ld] : 0.;
        return;
    }
    for (int c = 0; c < n_rhs; ++c)
    {
        double acc = 0.;
        for (long long k = lane; k < deg * dpn; k += 32)
        {
            const long long col = static_cast< long long >(node_nbr[node_ptr[n] + k % deg]) * dpn + k / deg;
            if (is_bc[col])
                acc = fma(vals[beg + k], bc_vals[col + c * ld], acc);
        }
        for (int off = 16; off > 0; off >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0)
            rhs[row + c * ld] -= acc;
    }
    for (long long k = lane; k < deg * dpn; k += 32)
    {
        const long long col = static_cast< long long >(node_nbr[node_ptr[n] + k % deg]) * dpn + k / deg;
        if (is_bc[col])
            vals[beg + k] = 0.;
    }
}
__global__ void extractDiagKernel(const long long* node_ptr, const uint32_t* node_nbr, const double* vals, long long n_nodes, int dpn,
                                  double* diag)
{
    const long long row = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x;
    if (row >= n_nodes * dpn)
        return;
    const long long n = row / dpn, d = row % dpn, deg = node_ptr[n + 1] - node_ptr[n];
    const long long beg = dpn * (dpn * node_ptr[n] + d * deg);
    long long       lo = 0, hi = deg;
    while (lo < hi)
    {
        const long long mid = (lo + hi) / 2;
        if (node_nbr[node_ptr[n] + mid] < n)
            lo = mid + 1;
        else
            hi = mid;
    }
    diag[row] = vals[beg + d * deg + lo];
}

// one thread per item (kernels without a grid-stride loop): exact block count, at least 1 so that empty ranks launch harmlessly
unsigned blocksFor(long long n, int block = 256)
{
    return static_cast< unsigned >(std::max< long long >(1, (n + block - 1) / block));
}
unsigned gridFor(long long n, int block = 256)
{
    const long long g = (n + block - 1) / block;
    return static_cast< unsigned >(std::max< long long >(1, std::min< long long >(g, 148ll * 32)));
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct l3b_context
{
    int                 device = 0;
    cudaStream_t        stream = nullptr;
    mutable std::string err;
    DevBuf< int >       status;
    DevBuf< double >    scalars; // Krylov scalars
    // Work vectors of the Krylov drivers, kept between solves: a time loop solves the same system again and again, a fresh set of
    // cudaMalloc'ed 0.5 GB vectors per solve costs tens of milliseconds of allocator time — and lands at different relative offsets
    // every time, which showed as a 2.4 - 3.3 ms spread of the CG iteration (the five streams of the vector passes collide in the
    // memory system when their bases line up). One slab, vectors carved out with staggered starts.
    DevBuf< double >    krylov_ws;
    double*             pinned_scalars = nullptr; // 8 doubles of page-locked host memory for the drivers' read-backs (cudaMallocHost and
                                                  // cudaFreeHost synchronise the device and cost up to a second with tens of GB mapped:
                                                  // once per context, not once per solve)
    cudaEvent_t         krylov_event   = nullptr;
    double*             pinnedScalars()
    {
        if (pinned_scalars == nullptr)
            cudaCheck(cudaMallocHost(&pinned_scalars, 8 * sizeof(double)), "cudaMallocHost");
        return pinned_scalars;
    }
    cudaEvent_t krylovEvent()
    {
        if (krylov_event == nullptr)
            cudaCheck(cudaEventCreateWithFlags(&krylov_event, cudaEventDisableTiming), "cudaEventCreate");
        return krylov_event;
    }
    double*             workspace(size_t n_doubles)
    {
        if (krylov_ws.n < n_doubles)
        {
            cudaCheck(cudaStreamSynchronize(stream), "sync");
            krylov_ws.alloc(n_doubles);
        }
        return krylov_ws.ptr;
    }
    int                 sm_count = 148;
    std::map< std::pair< int, int >, tables::Tables1D > tab1d;
    struct DenseDev
    {
        int              n_qp = 0;
        DevBuf< double > vals, ders, pts, wts; // domain: one table; boundary: n_sides tables back to back
        DevBuf< double > t1d;                  // domain only: 1-D interpolation and derivative tables [2][nb][nq]
        int              nq1d = 0;
    };
    std::map< std::tuple< int, int, int, bool >, std::unique_ptr< DenseDev > > dense;

    const tables::Tables1D& tables1d(int order, int nq)
    {
        auto it = tab1d.find({order, nq});
        if (it == tab1d.end())
            it = tab1d.emplace(std::make_pair(order, nq), tables::makeTables1D(order, nq)).first;
        return it->second;
    }
    const DenseDev& denseTables(int dim, int order, int nq, bool boundary)
    {
        const auto key = std::make_tuple(dim, order, nq, boundary);
        auto       it  = dense.find(key);
        if (it != dense.end())
            return *it->second;
        auto                    d = std::make_unique< DenseDev >();
        std::vector< double >   vals, ders, pts, wts;
        const int               n_tab = boundary ? 2 * dim : 1;
        for (int s = 0; s < n_tab; ++s)
        {
            std::vector< tables::ld > p, w;
            if (boundary)
                tables::sideQuadrature(dim, nq, s, p, w);
            else
                tables::domainQuadrature(dim, nq, p, w);
            const auto t = tables::makeDenseTables(dim, order, p, w);
            d->n_qp      = t.n_qp;
            vals.insert(vals.end(), t.values.begin(), t.values.end());
            ders.insert(ders.end(), t.derivatives.begin(), t.derivatives.end());
            pts.insert(pts.end(), t.points.begin(), t.points.end());
            wts.insert(wts.end(), t.weights.begin(), t.weights.end());
        }
        if (not boundary)
        {
            const auto&           t = tables1d(order, nq);
            std::vector< double > both(t.interp);
            both.insert(both.end(), t.der.begin(), t.der.end());
            d->t1d.alloc(both.size());
            d->t1d.upload(both.data(), both.size(), stream);
            d->nq1d = nq;
            cudaCheck(cudaStreamSynchronize(stream), "1-D table upload"); // `both` dies at the end of this block
        }
        d->vals.alloc(vals.size());
        d->ders.alloc(ders.size());
        d->pts.alloc(pts.size());
        d->wts.alloc(wts.size());
        d->vals.upload(vals.data(), vals.size(), stream);
        d->ders.upload(ders.data(), ders.size(), stream);
        d->pts.upload(pts.data(), pts.size(), stream);
        d->wts.upload(wts.data(), wts.size(), stream);
        cudaCheck(cudaStreamSynchronize(stream), "table upload");
        return *dense.emplace(key, std::move(d)).first->second;
    }
    // dense tables at the NODE locations (basis::getBasisAtNodes, mesh::getNodeLocations): point a = local node a
    std::map< std::pair< int, int >, std::unique_ptr< DenseDev > > node_tabs;
    const DenseDev& nodeTables(int dim, int order)
    {
        auto it = node_tabs.find({dim, order});
        if (it != node_tabs.end())
            return *it->second;
        const auto                gll = tables::gllNodes(order + 1);
        const int                 nb = order + 1, nn = cpow(nb, dim);
        std::vector< tables::ld > p(static_cast< size_t >(nn) * dim), w(nn, 1);
        for (int a = 0; a < nn; ++a)
        {
            int rem = a;
            for (int d = 0; d < dim; ++d, rem /= nb)
                p[static_cast< size_t >(a) * dim + d] = gll[rem % nb];
        }
        const auto t = tables::makeDenseTables(dim, order, p, w);
        auto       d = std::make_unique< DenseDev >();
        d->n_qp      = t.n_qp;
        d->vals.alloc(t.values.size());
        d->ders.alloc(t.derivatives.size());
        d->pts.alloc(t.points.size());
        d->vals.upload(t.values.data(), t.values.size(), stream);
        d->ders.upload(t.derivatives.data(), t.derivatives.size(), stream);
        d->pts.upload(t.points.data(), t.points.size(), stream);
        cudaCheck(cudaStreamSynchronize(stream), "table upload");
        return *node_tabs.emplace(std::make_pair(dim, order), std::move(d)).first->second;
    }
    // read-and-clear the device status word; translate to the reference's exceptions
    void checkStatus()
    {
        int st = 0;
        status.download(&st, 1, stream);
        cudaCheck(cudaStreamSynchronize(stream), "status read");
        if (st != 0)
        {
            status.zero(stream);
            if (st & status_degenerate_element)
                fail(L3B_ERR_DEGENERATE, "Encountered degenerate element ( |J| <= 0 )");
            if (st & status_singular_interior)
                fail(L3B_ERR_SINGULAR, "static condensation: the interior block K_ii of an element is singular (non-positive pivot)");
            if (st & status_slot_overflow)
                fail(L3B_ERR_GRAPH, "a node has more than 65535 neighbours: the 16-bit slot map cannot address its row");
            if (st & status_sparsity_violation)
                fail(L3B_ERR_SPARSITY, "an operator entry found structurally zero by the compile-time probe was non-zero at run time: "
                                       "register the kernel with a functor that is not constexpr-evaluable, or fix the probe");
            fail(L3B_ERR_GRAPH, "entry not present in the sparsity graph");
        }
    }
};

struct l3b_host_mesh
{
    host::Mesh mesh;
};

struct l3b_mesh
{
    l3b_context*        ctx = nullptr;
    int                 dim = 0, order = 0, nn = 0, n_sides = 0;
    long long           n_elems = 0, n_local_nodes = 0, n_owned_nodes = 0;
    DevBuf< double >    verts;
    DevBuf< uint32_t >  nodes;
    DevBuf< double >    hex_geo; // dim 3: per-element geometry record of the planes + columns apply kernel
    std::vector< uint16_t > side_bnd; // host copy, used to build boundary work lists
    std::vector< int32_t >  elem_domains; // host copy (l3b_mesh_set_element_domains), empty = one domain
};

struct l3b_fields
{
    l3b_context*     ctx = nullptr;
    long long        n_nodes = 0;
    int              n_fields = 0;
    DevBuf< double > data;
};

// communicator: the NCCL communicator of this rank plus the library's communication stream (comm.cuh)
struct l3b_comm
{
    l3b_context* ctx   = nullptr;
    ncclComm_t   comm  = nullptr;
    bool         owned = false; // created here (destroyed with the object) or attached (the caller's)
    int          rank = 0, world = 1;
    cudaStream_t stream = nullptr;                  // communication stream C
    cudaEvent_t  ev_ready = nullptr, ev_done = nullptr; // for the all-reduce hop S -> C -> S
    ~l3b_comm()
    {
        if (ev_ready)
            cudaEventDestroy(ev_ready);
        if (ev_done)
            cudaEventDestroy(ev_done);
        if (stream)
            cudaStreamDestroy(stream);
        if (owned and comm and l3b::comm::nccl().handle)
            l3b::comm::nccl().CommDestroy(comm);
    }
};
// comm::ImportExportContext + Import + Export of one dof layout (comm/ImportExport.hpp:29-72, 131-215)
struct l3b_halo
{
    l3b_comm*                comm = nullptr;
    long long                n_owned = 0, n_ghost = 0;
    std::vector< int >       owned_nbrs, shared_nbrs;
    std::vector< long long > owned_ptr, shared_off; // CSR offsets into owned_idx; offsets of the neighbours' ranges in the ghost block
    DevBuf< int32_t >        owned_idx;
    DevBuf< long long >      owned_ptr_dev;
    DevBuf< double >         buf; // pack buffer of the Import, receive buffer of the Export
    cudaEvent_t              ev_ready = nullptr, ev_done = nullptr;
    bool active() const { return comm != nullptr and (not owned_nbrs.empty() or not shared_nbrs.empty()); }
    ~l3b_halo()
    {
        if (ev_ready)
            cudaEventDestroy(ev_ready);
        if (ev_done)
            cudaEventDestroy(ev_done);
    }
};

namespace
{
void ncclCheck(ncclResult_t rc, const char* what)
{
    if (rc != ncclSuccess)
        fail(L3B_ERR_COMM, std::string{what} + ": " + l3b::comm::nccl().GetErrorString(rc));
}
const l3b::comm::NcclApi& ncclApi()
{
    const auto& api = l3b::comm::nccl();
    if (not api.handle)
        fail(L3B_ERR_COMM, api.error);
    return api;
}
void haloReserve(l3b_halo* h, int n_cols)
{
    const size_t need = static_cast< size_t >(h->owned_ptr.back()) * n_cols;
    if (h->buf.n < need)
    {
        cudaCheck(cudaStreamSynchronize(h->comm->stream), "sync"); // a transfer may still read the old buffer
        cudaCheck(cudaStreamSynchronize(h->comm->ctx->stream), "sync");
        h->buf.alloc(need);
    }
}
// one ncclGroup with every message of the exchange; `import`: owned -> ghost copies, else ghost -> owner contributions
void haloTransfer(l3b_halo* h, double* v, int n_cols, bool import)
{
    const auto&     api = ncclApi();
    auto* const     c   = h->comm;
    const long long ld  = h->n_owned + h->n_ghost;
    ncclCheck(api.GroupStart(), "ncclGroupStart");
    for (size_t k = 0; k < h->owned_nbrs.size(); ++k)
    {
        const long long n_k = h->owned_ptr[k + 1] - h->owned_ptr[k];
        for (int col = 0; col < n_cols and n_k > 0; ++col)
        {
            double* const chunk = h->buf.ptr + h->owned_ptr[k] * n_cols + col * n_k;
            if (import)
                ncclCheck(api.Send(chunk, n_k, ncclDouble, h->owned_nbrs[k], c->comm, c->stream), "ncclSend");
            else
                ncclCheck(api.Recv(chunk, n_k, ncclDouble, h->owned_nbrs[k], c->comm, c->stream), "ncclRecv");
        }
    }
    for (size_t j = 0; j < h->shared_nbrs.size(); ++j)
    {
        const long long n_j = h->shared_off[j + 1] - h->shared_off[j];
        for (int col = 0; col < n_cols and n_j > 0; ++col)
        {
            double* const range = v + h->n_owned + h->shared_off[j] + col * ld; // in place: the neighbour's ghosts are contiguous
            if (import)
                ncclCheck(api.Recv(range, n_j, ncclDouble, h->shared_nbrs[j], c->comm, c->stream), "ncclRecv");
            else
                ncclCheck(api.Send(range, n_j, ncclDouble, h->shared_nbrs[j], c->comm, c->stream), "ncclSend");
        }
    }
    ncclCheck(api.GroupEnd(), "ncclGroupEnd");
}
// Import (comm/ImportExport.hpp:296-384): begin = pack + post, end = the context stream waits for the ghost values
int haloImportBegin(l3b_halo* h, double* x, int n_cols)
{
    if (h == nullptr or not h->active())
        return 0;
    auto* const c = h->comm;
    const auto  S = c->ctx->stream;
    int         launches = 0;
    haloReserve(h, n_cols);
    if (h->owned_ptr.back() > 0)
    {
        l3b::comm::haloPackKernel<<< gridFor(h->owned_ptr.back() * n_cols, 256), 256, 0, S >>>(
            x, h->n_owned + h->n_ghost, h->owned_idx.ptr, h->owned_ptr_dev.ptr, static_cast< int >(h->owned_nbrs.size()), n_cols, h->buf.ptr);
        ++launches;
    }
    cudaCheck(cudaEventRecord(h->ev_ready, S), "event record");
    cudaCheck(cudaStreamWaitEvent(c->stream, h->ev_ready, 0), "event wait");
    haloTransfer(h, x, n_cols, true);
    cudaCheck(cudaEventRecord(h->ev_done, c->stream), "event record");
    return launches;
}
void haloImportEnd(l3b_halo* h)
{
    if (h == nullptr or not h->active())
        return;
    cudaCheck(cudaStreamWaitEvent(h->comm->ctx->stream, h->ev_done, 0), "event wait");
}
// Export (comm/ImportExport.hpp:403-470): begin = post (ordered after the work already on the context stream), end = wait + unpack-add
void haloExportBegin(l3b_halo* h, double* y, int n_cols)
{
    if (h == nullptr or not h->active())
        return;
    auto* const c = h->comm;
    haloReserve(h, n_cols);
    cudaCheck(cudaEventRecord(h->ev_ready, c->ctx->stream), "event record");
    cudaCheck(cudaStreamWaitEvent(c->stream, h->ev_ready, 0), "event wait");
    haloTransfer(h, y, n_cols, false);
    cudaCheck(cudaEventRecord(h->ev_done, c->stream), "event record");
}
int haloExportEnd(l3b_halo* h, double* y, int n_cols)
{
    if (h == nullptr or not h->active())
        return 0;
    const auto S = h->comm->ctx->stream;
    cudaCheck(cudaStreamWaitEvent(S, h->ev_done, 0), "event wait");
    if (h->owned_ptr.back() == 0)
        return 0;
    l3b::comm::haloUnpackAddKernel<<< gridFor(h->owned_ptr.back() * n_cols, 256), 256, 0, S >>>(
        y, h->n_owned + h->n_ghost, h->owned_idx.ptr, h->owned_ptr_dev.ptr, static_cast< int >(h->owned_nbrs.size()), n_cols, h->buf.ptr);
    return 1;
}
// in-place sum over the ranks of n device doubles, ordered after the work on the context stream and before what follows on it
void commAllReduce(l3b_comm* c, double* scalars, int n)
{
    if (c == nullptr or c->world == 1 or n == 0)
        return;
    const auto& api = ncclApi();
    const auto  S   = c->ctx->stream;
    cudaCheck(cudaEventRecord(c->ev_ready, S), "event record");
    cudaCheck(cudaStreamWaitEvent(c->stream, c->ev_ready, 0), "event wait");
    ncclCheck(api.AllReduce(scalars, scalars, n, ncclDouble, ncclSum, c->comm, c->stream), "ncclAllReduce");
    cudaCheck(cudaEventRecord(c->ev_done, c->stream), "event record");
    cudaCheck(cudaStreamWaitEvent(S, c->ev_done, 0), "event wait");
}

// values /= counts where a contribution arrived (averageElementContributions, ComputeValuesAtNodes.hpp:112-154)
__global__ void averageKernel(double* values, const double* counts, long long n, long long ld, int n_cols)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        if (counts[i] > 0.)
            for (int c = 0; c < n_cols; ++c)
                values[i + c * ld] /= counts[i];
}
// updateSolution (algsys/MatrixFreeSystem.hpp:1231-1273, AssembledSystem::updateSolution): dof columns of the solution -> nodal fields
__global__ void updateSolutionKernel(const double* x, long long n_nodes, int dpn, const int* dof_inds, const int* field_inds, int n, double* fields,
                                     long long field_stride)
{
    for (long long t = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; t < n_nodes * n; t += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const long long node = t % n_nodes;
        const int       i    = static_cast< int >(t / n_nodes);
        fields[node + field_inds[i] * field_stride] = x[node * dpn + dof_inds[i]];
    }
}
// shared-row export, receive side: the neighbour's values of my rows arrive in ITS layout (node block of dpn rows, each column-dof-major
// over its deg columns); entry (j-th received node, k-th column there) goes to position pos of my row of that node
__global__ void rowExportAddKernel(const double* buf, const long long* entry_ptr, const int32_t* row_node, const uint32_t* pos, long long n_nodes_recv,
                                   int dpn, const long long* node_ptr, double* vals)
{
    const long long n_entries = entry_ptr[n_nodes_recv];
    for (long long t = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; t < n_entries; t += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        long long lo = 0, hi = n_nodes_recv; // received node j of entry t: last j with entry_ptr[j] <= t
        while (hi - lo > 1)
        {
            const long long mid = (lo + hi) / 2;
            if (entry_ptr[mid] <= t)
                lo = mid;
            else
                hi = mid;
        }
        const long long deg_s = entry_ptr[lo + 1] - entry_ptr[lo], k = t - entry_ptr[lo];
        const long long n     = row_node[lo];
        const long long deg_r = node_ptr[n + 1] - node_ptr[n];
        const double*   src   = buf + static_cast< long long >(dpn) * dpn * entry_ptr[lo];
        double*         dst   = vals + static_cast< long long >(dpn) * dpn * node_ptr[n];
        for (int d = 0; d < dpn; ++d)
            for (int v = 0; v < dpn; ++v)
                atomicAdd(dst + (static_cast< long long >(d) * dpn + v) * deg_r + pos[t], src[(static_cast< long long >(d) * dpn + v) * deg_s + k]);
    }
}

struct WorkList
{
    long long          n = 0;
    DevBuf< int32_t >  elems;
    DevBuf< uint8_t >  sides;
};
struct KernelUse
{
    int                         kernel_id = -1;
    const KernelInstance*       inst      = nullptr;
    l3b_asm_opts                opts{};
    int                         nq = 0;
    double                      time = 0.;
    int                         dof_inds[max_unknowns]{};
    const l3b_fields*           fields = nullptr;
    int                         field_inds[max_fields]{};
    std::shared_ptr< WorkList > boundary_work; // boundary kernels only
    std::shared_ptr< WorkList > domain_work;   // domain kernels restricted to some element domains: ascending element ids
    std::vector< int32_t >      domain_elems;  // host copy of that list (to cut it to an element range)
};

// the (element, side) pairs lying on the given boundary ids (the reference's BoundaryView, mesh/BoundaryView.hpp)
std::shared_ptr< WorkList > makeBoundaryWork(l3b_mesh* mesh, const int* boundary_ids, int n_boundary_ids)
{
    if (mesh->side_bnd.empty())
        fail(L3B_ERR_INVALID_ARG, "boundary kernel on a mesh without side boundary ids");
    std::vector< int32_t > el;
    std::vector< uint8_t > sd;
    for (long long e = 0; e < mesh->n_elems; ++e)
        for (int s = 0; s < mesh->n_sides; ++s)
        {
            const auto id = mesh->side_bnd[e * mesh->n_sides + s];
            if (id == host::no_boundary)
                continue;
            for (int k = 0; k < n_boundary_ids; ++k)
                if (boundary_ids[k] == id)
                {
                    el.push_back(static_cast< int32_t >(e));
                    sd.push_back(static_cast< uint8_t >(s));
                    break;
                }
        }
    auto wl = std::make_shared< WorkList >();
    wl->n   = static_cast< long long >(el.size());
    wl->elems.alloc(el.size());
    wl->sides.alloc(sd.size());
    wl->elems.upload(el.data(), el.size(), mesh->ctx->stream);
    wl->sides.upload(sd.data(), sd.size(), mesh->ctx->stream);
    cudaCheck(cudaStreamSynchronize(mesh->ctx->stream), "work list upload");
    return wl;
}

KernelUse makeUse(l3b_mesh* mesh, int dofs_per_node, int n_rhs, int kernel_id, l3b_asm_opts opts, double time, const int* dof_inds,
                  const l3b_fields* fields, const int* field_inds, const int* boundary_ids, int n_boundary_ids)
{
    auto& reg = kernelRegistry();
    if (kernel_id < 0 or kernel_id >= static_cast< int >(reg.size()))
        fail(L3B_ERR_INVALID_ARG, "invalid kernel id");
    const auto& entry = reg[kernel_id];
    const auto& info  = entry.info;
    if (info.is_residual)
        fail(L3B_ERR_INVALID_ARG, "kernel '" + info.name + "' is a residual kernel (an integrand), not an equation kernel");
    if (info.dim != mesh->dim)
        fail(L3B_ERR_INVALID_ARG, "The dimensions of the kernel do not match the dimensions of the domain");
    if (info.n_rhs != n_rhs)
        fail(L3B_ERR_INVALID_ARG, "kernel n_rhs does not match the system's n_rhs");
    if (info.n_unknowns > dofs_per_node)
        fail(L3B_ERR_INVALID_ARG, "kernel has more unknowns than the system has dofs per node");
    if (opts.value_order < 1)
        fail(L3B_ERR_INVALID_ARG, "value_order must be >= 1");
    KernelUse use;
    use.kernel_id = kernel_id;
    use.opts      = opts;
    AssemblyOptions ao;
    ao.value_order      = opts.value_order;
    ao.derivative_order = opts.derivative_order;
    use.nq              = ao.nq1d(mesh->order);
    use.inst            = entry.find(mesh->order, use.nq);
    if (not use.inst)
        fail(L3B_ERR_NO_INSTANCE, "kernel '" + info.name + "' is not compiled for order " + std::to_string(mesh->order) + ", nq " +
                                      std::to_string(use.nq) + " — add L3B_PQ(" + std::to_string(mesh->order) + ", " +
                                      std::to_string(use.nq) + ") to its registration");
    use.time = time;
    for (int u = 0; u < info.n_unknowns; ++u)
    {
        use.dof_inds[u] = dof_inds ? dof_inds[u] : u;
        if (use.dof_inds[u] < 0 or use.dof_inds[u] >= dofs_per_node)
            fail(L3B_ERR_INVALID_ARG, "dof index out of range");
    }
    if (info.n_fields > 0)
    {
        if (not fields)
            fail(L3B_ERR_INVALID_ARG, "kernel needs external fields but none were passed");
        if (fields->n_nodes != mesh->n_local_nodes)
            fail(L3B_ERR_INVALID_ARG, "field storage does not match the mesh's local node count");
        for (int f = 0; f < info.n_fields; ++f)
        {
            use.field_inds[f] = field_inds ? field_inds[f] : f;
            if (use.field_inds[f] < 0 or use.field_inds[f] >= fields->n_fields)
                fail(L3B_ERR_INVALID_ARG, "field index out of range");
        }
        use.fields = fields;
    }
    if (info.is_boundary)
        use.boundary_work = makeBoundaryWork(mesh, boundary_ids, n_boundary_ids);
    else if (n_boundary_ids > 0 and not mesh->elem_domains.empty())
    {
        // assembleProblem(kernel, domain_ids, ...): the elements of the listed domains only (mesh.visit(..., domain_ids))
        for (long long e = 0; e < mesh->n_elems; ++e)
            for (int k = 0; k < n_boundary_ids; ++k)
                if (mesh->elem_domains[e] == boundary_ids[k])
                {
                    use.domain_elems.push_back(static_cast< int32_t >(e));
                    break;
                }
        auto wl = std::make_shared< WorkList >();
        wl->n   = static_cast< long long >(use.domain_elems.size());
        wl->elems.alloc(std::max< size_t >(1, use.domain_elems.size()));
        wl->elems.upload(use.domain_elems.data(), use.domain_elems.size(), mesh->ctx->stream);
        cudaCheck(cudaStreamSynchronize(mesh->ctx->stream), "work list upload");
        use.domain_work = std::move(wl);
    }
    return use;
}

ElemArgs baseArgs(l3b_mesh* mesh, const KernelUse& use, int dofs_per_node, long long ld)
{
    ElemArgs a{};
    a.verts         = mesh->verts.ptr;
    a.nodes         = mesh->nodes.ptr;
    a.hex_geo       = mesh->hex_geo.ptr;
    a.dofs_per_node = dofs_per_node;
    a.ld            = ld;
    std::memcpy(a.dof_inds, use.dof_inds, sizeof(a.dof_inds));
    std::memcpy(a.field_inds, use.field_inds, sizeof(a.field_inds));
    if (use.fields)
    {
        a.fields       = use.fields->data.ptr;
        a.field_stride = use.fields->n_nodes;
    }
    a.time   = use.time;
    a.status = mesh->ctx->status.ptr;
    if (use.boundary_work)
    {
        a.n_work     = use.boundary_work->n;
        a.work_elems = use.boundary_work->elems.ptr;
        a.work_sides = use.boundary_work->sides.ptr;
    }
    else if (use.domain_work)
    {
        a.n_work     = use.domain_work->n;
        a.work_elems = use.domain_work->elems.ptr;
    }
    else
    {
        a.n_work     = mesh->n_elems;
        a.first_elem = 0;
    }
    return a;
}
void setDense(ElemArgs& a, l3b_mesh* mesh, const KernelUse& use, bool boundary)
{
    const auto& d = mesh->ctx->denseTables(mesh->dim, mesh->order, use.nq, boundary);
    a.tab_vals    = d.vals.ptr;
    a.tab_ders    = d.ders.ptr;
    a.tab_pts     = d.pts.ptr;
    a.tab_wts     = d.wts.ptr;
    a.n_qp        = d.n_qp;
    a.tab1d       = boundary ? nullptr : d.t1d.ptr;
    a.nq1d        = d.nq1d;
}
} // namespace

struct l3b_asm
{
    l3b_context*        ctx  = nullptr;
    l3b_mesh*           mesh = nullptr;
    int                 dpn = 0, n_rhs = 1;
    long long           n_nodes = 0; // rows / dpn; == mesh->n_local_nodes when there is a mesh (l3b_crs_create: none)
    long long           n_dofs = 0, nnz = 0;
    DevBuf< long long > node_ptr, row_ptr;
    std::vector< long long > node_ptr_host; // for the layout conversion of l3b_asm_download
    DevBuf< uint32_t >  node_nbr;
    DevBuf< uint16_t >  slot_pos;
    DevBuf< double >    values, rhs;
    bool                open = false;
    cudaEvent_t         ev0 = nullptr, ev1 = nullptr;
    double              last_ms = 0.;
    // more than one rank (l3b_asm_set_halo): halo of the row layout [owned | ghost]
    l3b_halo*           halo = nullptr;
    long long           n_owned_dofs = -1; // -1: all rows owned
    bool                rows_exported = false;
    const host::DofMap* dofmap = nullptr; // l3b_asm_set_dofmap: inactive (node, dof) pairs are closed as identity rows // l3b_asm_export_shared_rows ran: owned rows are complete, ghost rows are spent
    long long           ownedDofs() const { return n_owned_dofs < 0 ? n_dofs : n_owned_dofs; }
    l3b_comm*           comm() const { return halo ? halo->comm : nullptr; }
    ~l3b_asm()
    {
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
    }
};

// state of the streamed host-buffer apply (apply_plan_host.hpp): the schedule, the two copy streams and one event pair per item
struct HostApplyPipe
{
    l3b::host::ApplyPlan       plan;
    cudaStream_t               up = nullptr, down = nullptr; // copy-in / copy-out streams
    cudaEvent_t                begin = nullptr;
    std::vector< cudaEvent_t > landed, done; // per item: its x blocks are on the device / its final y blocks may leave
    HostApplyPipe()                                = default;
    HostApplyPipe(const HostApplyPipe&)            = delete;
    HostApplyPipe& operator=(const HostApplyPipe&) = delete;
    ~HostApplyPipe()
    {
        for (auto ev : landed)
            cudaEventDestroy(ev);
        for (auto ev : done)
            cudaEventDestroy(ev);
        if (begin)
            cudaEventDestroy(begin);
        if (up)
            cudaStreamDestroy(up);
        if (down)
            cudaStreamDestroy(down);
    }
};

struct l3b_mf
{
    l3b_context*             ctx  = nullptr;
    l3b_mesh*                mesh = nullptr;
    int                      dpn = 0, n_rhs = 1;
    long long                n_dofs = 0;
    DevBuf< uint8_t >        dir_mask;
    DevBuf< uint32_t >       elem_dir;
    DevBuf< int32_t >        dir_list; // the Dirichlet dofs, ascending
    long long                n_dir = 0;
    DevBuf< double >         dir_g;            // prescribed values, zero off the Dirichlet dofs: operand of the lifting apply
    bool                     lift_needed = false; // some prescribed value is non-zero
    DevBuf< double >         dir_vals, diag, rhs;
    bool                     has_bc = false, closed = false;
    std::vector< KernelUse > uses;
    DevBuf< double >         work_x, work_y; // staging for the host-buffer apply
    std::vector< int32_t >   dir_list_host;  // host copy of dir_list (the streamed host-buffer apply cuts it by node range)
    // host-buffer apply: 0 = serial (copy in, apply, copy out), 1 = streamed when it pays (the default), 2 = streamed whenever legal
    int                              host_apply_mode = 1, host_apply_chunks = 0; // 0 = by size, see hostApplyPipe
    long long                        host_apply_block_nodes = 65536;
    std::unique_ptr< HostApplyPipe > pipe;
    bool                             last_host_apply_streamed = false;
    int                      last_launches = 0;
    // more than one rank (l3b_mf_set_halo): the halo of the dof layout [owned | ghost]; elements [0, n_border) touch ghost nodes
    l3b_halo*                halo     = nullptr;
    long long                n_border = 0;
    int                      reserve_ctas = 0; // set around the launches that must leave room for the halo's NCCL kernels
    long long                ownedDofs() const { return static_cast< long long >(mesh->n_owned_nodes) * dpn; }
    l3b_comm*                comm() const { return halo ? halo->comm : nullptr; }
};

namespace
{
// every context-bearing entry point runs with the context's device current (and restores the caller's): two contexts in one
// process, another host thread, or a framework that switches devices must not redirect allocations and launches
struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(const l3b_context* ctx)
    {
        if (ctx == nullptr)
            return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess and cur != ctx->device)
        {
            prev = cur;
            cudaSetDevice(ctx->device);
        }
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&)            = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
template < typename F >
int guardedCtx(const l3b_context* ctx, F&& f)
{
    try
    {
        const DeviceGuard guard{ctx};
        // a stale, non-sticky error left in this thread by another library (NCCL's peer probing leaves cudaErrorInvalidDevice on a
        // one-GPU box) must not be blamed on the next launch checked with cudaGetLastError
        (void)cudaGetLastError();
        f();
        return L3B_OK;
    }
    catch (const Error& e)
    {
        if (ctx)
            ctx->err = e.msg;
        g_global_error = e.msg;
        return e.code;
    }
    catch (const std::exception& e)
    {
        if (ctx)
            ctx->err = e.what();
        g_global_error = e.what();
        return L3B_ERR_INVALID_ARG;
    }
}

// y[dofs(e)] += alpha K_e x[dofs(e)] for one registered kernel over the domain elements [elem_begin, elem_end) (boundary kernels:
// their whole side list, with the range that contains element 0). `masked`: honour the Dirichlet mask (the operator apply) or not
// (the Dirichlet lifting of the initialisation). Returns the number of kernels launched.
int applyUse(l3b_mf* sys, const KernelUse& use, const double* x, double* y, int n_cols, double alpha, bool masked, double* energy, long long elem_begin,
             long long elem_end, bool boundary_work)
{
    auto*       ctx  = sys->ctx;
    const auto& info = kernelRegistry()[use.kernel_id].info;
    ElemArgs    a    = baseArgs(sys->mesh, use, sys->dpn, sys->n_dofs);
    if (info.is_boundary)
    {
        if (not boundary_work) // the side work list is not split: it runs once per apply, when the caller says so
            return 0;
    }
    else
    {
        if (elem_end <= elem_begin)
            return 0;
        if (use.domain_work)
        {
            // the part of the (ascending) domain list that falls into [elem_begin, elem_end)
            const auto& el = use.domain_elems;
            const auto  lo = std::lower_bound(el.begin(), el.end(), static_cast< int32_t >(elem_begin)) - el.begin();
            const auto  hi = std::lower_bound(el.begin(), el.end(), static_cast< int32_t >(std::min< long long >(elem_end, 0x7fffffff))) - el.begin();
            if (hi <= lo)
                return 0;
            a.work_elems = use.domain_work->elems.ptr + lo;
            a.n_work     = hi - lo;
        }
        else
        {
            a.first_elem = elem_begin;
            a.n_work     = elem_end - elem_begin;
        }
    }
    a.x              = x;
    a.y              = y;
    a.n_cols         = n_cols;
    a.alpha          = alpha;
    a.energy         = energy;
    a.reserve_ctas   = sys->reserve_ctas;
    a.dir_mask       = sys->has_bc and masked ? sys->dir_mask.ptr : nullptr;
    a.elem_dir       = sys->has_bc and masked ? sys->elem_dir.ptr : nullptr;
    bool contiguous  = info.n_unknowns == sys->dpn and reinterpret_cast< uintptr_t >(x) % 16 == 0 and sys->n_dofs % 2 == 0;
    for (int u = 0; u < info.n_unknowns; ++u)
        contiguous = contiguous and use.dof_inds[u] == u;
    a.contiguous_dofs = contiguous;
    const bool full  = n_cols == sys->n_rhs;
    const bool sf    = not info.is_boundary and use.opts.eval_strategy != 1;
    cudaError_t err;
    if (sf)
    {
        const auto& t = ctx->tables1d(sys->mesh->order, use.nq);
        err           = (full ? use.inst->mf_sumfact_full : use.inst->mf_sumfact_one)(kernelRegistry()[use.kernel_id].object.get(), a, t, ctx->stream);
    }
    else
    {
        setDense(a, sys->mesh, use, info.is_boundary);
        err = (full ? use.inst->local_apply_full : use.inst->local_apply_one)(kernelRegistry()[use.kernel_id].object.get(), a, ctx->stream);
    }
    cudaCheck(err, "operator apply launch");
    return a.n_work > 0 ? 1 : 0;
}

// y = alpha A x + beta y on device pointers, in the phases of l3b_mf_apply_phase_device
// energy (may be null): the ELEMENTS and FINISH phases add x^T A x of column 0, restricted to this rank's elements and owned Dirichlet
// dofs, to this device scalar
void mfApplyPhases(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta, int phases, long long elem_begin,
                   long long elem_end, double* energy = nullptr)
{
    auto* ctx = sys->ctx;
    if (not sys->closed)
        fail(L3B_ERR_STATE, "`apply` was called before `endAssembly()`");
    if (n_cols != sys->n_rhs and n_cols != 1)
        fail(L3B_ERR_INVALID_ARG, "n_cols must equal the system's n_rhs or 1");
    if (elem_begin < 0 or elem_end > sys->mesh->n_elems or elem_begin > elem_end)
        fail(L3B_ERR_INVALID_ARG, "element range out of bounds");
    if (energy != nullptr and n_cols != 1)
        fail(L3B_ERR_INVALID_ARG, "the energy x^T A x is collected for single-column applies only");
    int launches = 0;
    if (phases & L3B_APPLY_INIT)
    {
        if (beta == 0.)
            cudaCheck(cudaMemsetAsync(y, 0, sizeof(double) * sys->n_dofs * n_cols, ctx->stream), "memset");
        else
        {
            scaleKernel<<< gridFor(sys->n_dofs * n_cols), 256, 0, ctx->stream >>>(y, sys->n_dofs * n_cols, beta);
            ++launches;
        }
    }
    // boundary kernels run with L3B_APPLY_BOUNDARY only (the public entry point adds the bit by its documented rule)
    const bool boundary_work = (phases & L3B_APPLY_BOUNDARY) != 0;
    if (phases & (L3B_APPLY_ELEMENTS | L3B_APPLY_BOUNDARY))
        for (const auto& use : sys->uses)
        {
            const bool is_bnd = kernelRegistry()[use.kernel_id].info.is_boundary;
            if (is_bnd ? boundary_work : (phases & L3B_APPLY_ELEMENTS) != 0)
                launches += applyUse(sys, use, x, y, n_cols, alpha, true, energy, elem_begin, elem_end, boundary_work);
        }
    if ((phases & L3B_APPLY_FINISH) and sys->has_bc and sys->n_dir > 0)
    {
        const ProfileRegion region{"Impose Dirichlet BCs"}; // MatrixFreeSystem.hpp:1101-1104
        dirichletRowsListKernel<<< gridFor(sys->n_dir), 256, 0, ctx->stream >>>(sys->dir_list.ptr, sys->n_dir, x, y, sys->n_dofs, n_cols, alpha, energy,
                                                                                static_cast< long long >(sys->mesh->n_owned_nodes) * sys->dpn);
        ++launches;
    }
    cudaCheck(cudaGetLastError(), "operator apply");
    sys->last_launches = launches; // of this call
}
// The whole operator apply. One rank: one pass over all elements. With a halo (MatrixFreeSystem::applyImpl, :1019-1140):
//     pack + post the Import of x | y <- beta y | first half of the interior elements | wait | border elements + boundary kernels |
//     post the Export of y | second half of the interior elements | wait + unpack-add | Dirichlet rows
// The reference runs its interior elements while it polls for the Import and exposes the Export; here each transfer hides behind half
// of the interior work. The ghost block of x is overwritten by the Import.
void mfApplyDevice(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta, double* energy = nullptr)
{
    const ProfileRegion region{"Evaluate matrix-free operator"}; // MatrixFreeSystem.hpp:1032-1139
    auto* const h = sys->halo;
    if (h == nullptr or not h->active())
    {
        mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_INIT | L3B_APPLY_ELEMENTS | L3B_APPLY_BOUNDARY | L3B_APPLY_FINISH, 0, sys->mesh->n_elems, energy);
        return;
    }
    int        launches = haloImportBegin(h, const_cast< double* >(x), n_cols);
    const auto n_elems  = sys->mesh->n_elems;
    // Two schedules. Border-first: Import behind the pass that zeroes y, border elements, Export behind ALL interior elements — the
    // cheapest (one interior launch), but the Import has only ~80 us of slack. Split: half of the interior elements before the border
    // elements (hiding the Import), half after (hiding the Export) — either exchange may lag by half an apply before anybody waits.
    // A rank with a single neighbour (2 ranks; the ends of a chain) runs border-first, a rank with several — whose neighbours drift
    // against each other — runs split. Measured, 64^3 hex p=4 per GPU, ms per apply: 2 GPUs 1.48 border-first / 1.67 split;
    // 8 GPUs 1.70 - 2.26 border-first / 1.51 split (1.43 on one GPU). L3B_APPLY_SPLIT=0 / 1 forces one schedule on every rank.
    static const int forced = [] {
        const char* e = std::getenv("L3B_APPLY_SPLIT");
        return e == nullptr ? -1 : e[0] != '0';
    }();
    int n_nbrs = static_cast< int >(h->owned_nbrs.size());
    for (int r : h->shared_nbrs)
        n_nbrs += std::find(h->owned_nbrs.begin(), h->owned_nbrs.end(), r) == h->owned_nbrs.end();
    // (round 2, later) with more than two ranks EVERY rank runs split: an end-of-chain rank on border-first stalls whenever its — split —
    // neighbour posts its sends late, and the stall travels back through the Export: 4 GPUs measured 1.52 ms and 1.73 ms per apply in two
    // runs of the mixed policy, 8 GPUs 1.51 / 1.52 with split forced everywhere
    const bool split = forced >= 0 ? forced != 0 : (n_nbrs >= 2 or h->comm->world > 2);
    const long long half = split ? sys->n_border + (n_elems - sys->n_border) / 2 : sys->n_border;
    // The persistent element kernel fills every resident CTA slot; an NCCL send / recv kernel queued while it runs would wait for it to
    // finish (registers, not SMs, are what is full). Leaving a few slots free lets the transfers start at once.
    static const int reserve = [] {
        const char* e = std::getenv("L3B_HALO_RESERVE_CTAS");
        return e != nullptr ? std::atoi(e) : 0; // measured on 2 and 8 GPUs: no gain from 8 free slots (NCCL's kernel gets in anyway)
    }();
    struct ReserveScope
    {
        l3b_mf* s;
        ReserveScope(l3b_mf* s_, int r) : s{s_} { s->reserve_ctas = r; }
        ~ReserveScope() { s->reserve_ctas = 0; }
    } const reserve_scope{sys, reserve};
    mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_INIT, 0, 0);
    launches += sys->last_launches;
    if (half > sys->n_border)
    {
        mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_ELEMENTS, sys->n_border, half, energy);
        launches += sys->last_launches;
    }
    haloImportEnd(h);
    mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_ELEMENTS | L3B_APPLY_BOUNDARY, 0, sys->n_border, energy);
    launches += sys->last_launches;
    haloExportBegin(h, y, n_cols);
    if (half < n_elems)
    {
        mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_ELEMENTS, half, n_elems, energy);
        launches += sys->last_launches;
    }
    launches += haloExportEnd(h, y, n_cols);
    mfApplyPhases(sys, x, y, n_cols, alpha, beta, L3B_APPLY_FINISH, 0, 0, energy);
    sys->last_launches += launches;
}

// ---- streamed host-buffer apply (apply_plan_host.hpp): x over PCIe, the element chunks and y over PCIe as three concurrent streams
HostApplyPipe& hostApplyPipe(l3b_mf* sys)
{
    if (sys->pipe)
        return *sys->pipe;
    auto        p    = std::make_unique< HostApplyPipe >();
    auto* const mesh = sys->mesh;
    const auto  S    = sys->ctx->stream;
    const bool  halo = sys->halo != nullptr and sys->halo->active();
    // the connectivity lives on the device only: one read-back per system (131 MB at 64^3 hex p=4), not per apply
    std::vector< uint32_t > nodes(static_cast< size_t >(mesh->n_elems) * mesh->nn);
    mesh->nodes.download(nodes.data(), nodes.size(), S);
    std::vector< int32_t > halo_nodes;
    if (halo and sys->halo->owned_ptr.back() > 0)
    {
        halo_nodes.resize(static_cast< size_t >(sys->halo->owned_ptr.back()));
        sys->halo->owned_idx.download(halo_nodes.data(), halo_nodes.size(), S);
    }
    cudaCheck(cudaStreamSynchronize(S), "connectivity read-back");
    for (auto& d : halo_nodes)
        d /= sys->dpn; // dof index -> node
    const long long n_border   = halo ? sys->n_border : 0;
    const long long n_interior = mesh->n_elems - n_border;
    // How many chunks: with K equal chunks the call costs T (1 + 1 / K) + K c — T the slower of the two PCIe directions when both run, the
    // T / K is the last chunk's y leaving after the last x block has landed (the bytes in flight never drop below one chunk), c the
    // ~25 us of copy set-up, event hand-over and launch ramp per item. Measured at 543 MB per vector (profiles/r2_host_apply_sweep.jsonl):
    // K = 12 / 24 / 48 / 96 / 192 -> 12.9 / 12.7 / 13.3 / 14.5 / 16.2 ms against T = 11.4 ms. Hence about one chunk per 22 MB, at most 32.
    const long long bytes      = sys->n_dofs * static_cast< long long >(sizeof(double));
    const int       n_chunks   = sys->host_apply_chunks > 0 ? sys->host_apply_chunks
                                                            : static_cast< int >(std::clamp< long long >(bytes / (22ll << 20), 2, 32));
    const long long chunk      = std::max< long long >(1, (n_interior + n_chunks - 1) / n_chunks);
    try
    {
        p->plan = l3b::host::makeApplyPlan(mesh->n_local_nodes, halo ? mesh->n_owned_nodes : mesh->n_local_nodes, mesh->n_elems, mesh->nn, nodes.data(),
                                           n_border, halo_nodes.data(), static_cast< long long >(halo_nodes.size()), chunk,
                                           std::max< long long >(1, sys->host_apply_block_nodes));
    }
    catch (const std::exception& e)
    {
        fail(L3B_ERR_INVALID_ARG, std::string{"host apply plan: "} + e.what());
    }
    cudaCheck(cudaStreamCreateWithFlags(&p->up, cudaStreamNonBlocking), "stream create");
    cudaCheck(cudaStreamCreateWithFlags(&p->down, cudaStreamNonBlocking), "stream create");
    cudaCheck(cudaEventCreateWithFlags(&p->begin, cudaEventDisableTiming), "event create");
    p->landed.assign(p->plan.n_items, nullptr);
    p->done.assign(p->plan.n_items, nullptr);
    for (int k = 0; k < p->plan.n_items; ++k)
    {
        cudaCheck(cudaEventCreateWithFlags(&p->landed[k], cudaEventDisableTiming), "event create");
        cudaCheck(cudaEventCreateWithFlags(&p->done[k], cudaEventDisableTiming), "event create");
    }
    sys->pipe = std::move(p);
    return *sys->pipe;
}
// may this call take the streamed form? y <- alpha A x only (beta y would need y over PCIe in both directions), domain kernels only
// (the side list of a boundary kernel is not cut into chunks), distinct host vectors (y blocks land in host memory while x blocks are
// still being read), and — unless forced — vectors large enough for the copies to matter
bool hostApplyStreams(const l3b_mf* sys, const double* x, const double* y, int n_cols, double beta)
{
    if (sys->host_apply_mode == 0 or beta != 0. or not sys->closed)
        return false;
    for (const auto& use : sys->uses)
        if (kernelRegistry()[use.kernel_id].info.is_boundary)
            return false;
    const size_t n = static_cast< size_t >(sys->n_dofs) * n_cols;
    const auto xb = reinterpret_cast< uintptr_t >(x), yb = reinterpret_cast< uintptr_t >(y);
    if (n == 0 or (xb < yb + n * sizeof(double) and yb < xb + n * sizeof(double)))
        return false;
    return sys->host_apply_mode == 2 or n * sizeof(double) >= (size_t{8} << 20);
}
void mfApplyHostStreamed(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha)
{
    const ProfileRegion region{"Evaluate matrix-free operator"};
    auto&           pipe = hostApplyPipe(sys);
    const auto&     plan = pipe.plan;
    auto* const     ctx  = sys->ctx;
    const auto      S    = ctx->stream;
    const long long ld   = sys->n_dofs;
    const int       dpn  = sys->dpn;
    double* const   wx   = sys->work_x.ptr;
    double* const   wy   = sys->work_y.ptr;
    // both copy streams start behind whatever the context stream still does with the staging vectors
    cudaCheck(cudaEventRecord(pipe.begin, S), "event record");
    cudaCheck(cudaStreamWaitEvent(pipe.up, pipe.begin, 0), "event wait");
    cudaCheck(cudaStreamWaitEvent(pipe.down, pipe.begin, 0), "event wait");
    for (int k = 0; k < plan.n_items; ++k)
    {
        for (long long r = plan.up_ptr[k]; r < plan.up_ptr[k + 1]; ++r)
        {
            const long long d0 = plan.up_ranges[2 * r] * dpn, cnt = (plan.up_ranges[2 * r + 1] - plan.up_ranges[2 * r]) * dpn;
            for (int c = 0; c < n_cols; ++c)
                cudaCheck(cudaMemcpyAsync(wx + d0 + c * ld, x + d0 + c * ld, cnt * sizeof(double), cudaMemcpyHostToDevice, pipe.up), "H2D copy");
        }
        cudaCheck(cudaEventRecord(pipe.landed[k], pipe.up), "event record");
    }
    int launches = 0;
    mfApplyPhases(sys, wx, wy, n_cols, alpha, 0., L3B_APPLY_INIT, 0, 0);
    launches += sys->last_launches;
    for (int k = 0; k < plan.n_items; ++k)
    {
        cudaCheck(cudaStreamWaitEvent(S, pipe.landed[k], 0), "event wait");
        const long long e0 = plan.item_elems[2 * k], e1 = plan.item_elems[2 * k + 1];
        if (plan.halo_item and k == plan.n_items - 1)
        {
            // border elements + exchange (MatrixFreeSystem.hpp:1046-1122), after every interior chunk: its blocks leave last
            launches += haloImportBegin(sys->halo, wx, n_cols);
            haloImportEnd(sys->halo);
            if (e1 > e0)
            {
                mfApplyPhases(sys, wx, wy, n_cols, alpha, 0., L3B_APPLY_ELEMENTS, e0, e1);
                launches += sys->last_launches;
            }
            haloExportBegin(sys->halo, wy, n_cols);
            launches += haloExportEnd(sys->halo, wy, n_cols);
        }
        else if (e1 > e0)
        {
            mfApplyPhases(sys, wx, wy, n_cols, alpha, 0., L3B_APPLY_ELEMENTS, e0, e1);
            launches += sys->last_launches;
        }
        // Dirichlet identity rows (:1087-1103) of the blocks that are final now, then these blocks leave
        for (long long r = plan.down_ptr[k]; r < plan.down_ptr[k + 1]; ++r)
        {
            if (not sys->has_bc or sys->n_dir == 0)
                break;
            const long long d0 = plan.down_ranges[2 * r] * dpn, d1 = plan.down_ranges[2 * r + 1] * dpn;
            const auto&     dl = sys->dir_list_host;
            const auto      lo = std::lower_bound(dl.begin(), dl.end(), static_cast< int32_t >(std::min< long long >(d0, 0x7fffffff))) - dl.begin();
            const auto      hi = std::lower_bound(dl.begin(), dl.end(), static_cast< int32_t >(std::min< long long >(d1, 0x7fffffff))) - dl.begin();
            if (hi > lo)
            {
                dirichletRowsListKernel<<< gridFor(hi - lo), 256, 0, S >>>(sys->dir_list.ptr + lo, hi - lo, wx, wy, ld, n_cols, alpha, nullptr,
                                                                           static_cast< long long >(sys->mesh->n_owned_nodes) * dpn);
                ++launches;
            }
        }
        cudaCheck(cudaEventRecord(pipe.done[k], S), "event record");
        cudaCheck(cudaStreamWaitEvent(pipe.down, pipe.done[k], 0), "event wait");
        for (long long r = plan.down_ptr[k]; r < plan.down_ptr[k + 1]; ++r)
        {
            const long long d0 = plan.down_ranges[2 * r] * dpn, cnt = (plan.down_ranges[2 * r + 1] - plan.down_ranges[2 * r]) * dpn;
            for (int c = 0; c < n_cols; ++c)
                cudaCheck(cudaMemcpyAsync(y + d0 + c * ld, wy + d0 + c * ld, cnt * sizeof(double), cudaMemcpyDeviceToHost, pipe.down), "D2H copy");
        }
    }
    cudaCheck(cudaGetLastError(), "streamed operator apply");
    cudaCheck(cudaStreamSynchronize(pipe.down), "apply");
    cudaCheck(cudaStreamSynchronize(S), "apply");
    cudaCheck(cudaStreamSynchronize(pipe.up), "apply");
    ctx->checkStatus();
    sys->last_launches            = launches;
    sys->last_host_apply_streamed = true;
}

// Preconditioned CG, Belos "Block CG" semantics for block size 1 (solve/BelosSolvers.hpp:76-89): left preconditioner,
// absolute 2-norm of the (unpreconditioned) residual against `tol`, x0 = 0. Two fused vector kernels (10 vector passes) and one
// 2-scalar read-back per iteration. The read-back goes to pinned memory behind an event, and the host waits for it only after it has
// queued the next direction update and operator apply: the device never idles on the convergence test (the speculative work
// touches p and Ap only, so x is final when the test succeeds).
struct PinnedScalars
{
    double* h = nullptr;
    PinnedScalars() { cudaCheck(cudaMallocHost(&h, 8 * sizeof(double)), "cudaMallocHost"); }
    ~PinnedScalars() { cudaFreeHost(h); }
    PinnedScalars(const PinnedScalars&)            = delete;
    PinnedScalars& operator=(const PinnedScalars&) = delete;
};
struct Event
{
    cudaEvent_t ev = nullptr;
    Event() { cudaCheck(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate"); }
    ~Event() { cudaEventDestroy(ev); }
    Event(const Event&)            = delete;
    Event& operator=(const Event&) = delete;
};
// r = b - Ap
__global__ void residualKernel(double* r, const double* b, const double* Ap, long long n)
{
    for (long long i = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; i < n; i += static_cast< long long >(gridDim.x) * blockDim.x)
        r[i] = b[i] - Ap[i];
}
// x0_zero: the caller vouches that the initial guess is zero (the reference's first solve: m_solution starts at 0) — x is zeroed and the
// apply for r0 = b - A x0 is skipped; otherwise x is the initial guess, as Belos takes the system's persistent solution vector
// (AssembledSystem.hpp:113-135, solve/BelosSolvers.hpp: LinearProblem(A, x, b)). Only the owned part of x is valid on return.
template < typename Apply, typename Reduce >
void pcg(l3b_context* ctx, long long n_local, long long n, Apply&& apply, Reduce&& reduce, const double* diag, const double* b,
         double* x /* device */, double tol, int max_iters, double* achieved, int* iters, bool x0_zero = true)
{
    // n = owned dofs: dots and updates run over those; p and Ap carry the ghost tail the operator needs
    struct Vec
    {
        double* ptr;
        size_t  n;
        void    zero(cudaStream_t st) const
        {
            if (n > 0)
                cudaCheck(cudaMemsetAsync(ptr, 0, n * sizeof(double), st), "memset");
        }
    };
    // starts staggered by an odd number of 256-byte lines so that the streams of one vector pass do not share a phase
    const auto       padded = [](long long len, int k) { return static_cast< size_t >((len + 31) / 32 * 32 + 32 * (17 + 2 * k)); };
    const size_t     total  = padded(n, 0) + padded(n, 1) + padded(n_local, 2) + padded(n_local, 3) + padded(n, 4) + 64;
    double*          ws     = ctx->workspace(total);
    const Vec        r{ws, static_cast< size_t >(n)};
    const Vec        z{r.ptr + padded(n, 0), static_cast< size_t >(n)};
    const Vec        p{z.ptr + padded(n, 1), static_cast< size_t >(n_local)};
    const Vec        Ap{p.ptr + padded(n_local, 2), static_cast< size_t >(n_local)};
    const Vec        minv{Ap.ptr + padded(n_local, 3), static_cast< size_t >(n)};
    const Vec        sc_buf{minv.ptr + padded(n, 4), 8};
    const cudaEvent_t read_back_ev = ctx->krylovEvent();
    double*    sc = sc_buf.ptr; // [0] rz, [1] pAp, [2] rr, [3] rz_new
    double*    h  = ctx->pinnedScalars();
    const auto s  = ctx->stream;
    const auto g  = gridFor(n);
    jacobiInvertKernel<<< g, 256, 0, s >>>(diag, minv.ptr, n, 1., 0.);
    p.zero(s);
    if (x0_zero)
    {
        cudaCheck(cudaMemsetAsync(x, 0, n * sizeof(double), s), "memset");
        cudaCheck(cudaMemcpyAsync(r.ptr, b, n * sizeof(double), cudaMemcpyDeviceToDevice, s), "copy");
    }
    else
    {
        cudaCheck(cudaMemcpyAsync(p.ptr, x, n * sizeof(double), cudaMemcpyDeviceToDevice, s), "copy"); // p carries the ghost tail the operator fills
        apply(p.ptr, Ap.ptr, nullptr);
        residualKernel<<< g, 256, 0, s >>>(r.ptr, b, Ap.ptr, n);
    }
    hadamardKernel<<< g, 256, 0, s >>>(z.ptr, minv.ptr, r.ptr, n);
    cudaCheck(cudaMemcpyAsync(p.ptr, z.ptr, n * sizeof(double), cudaMemcpyDeviceToDevice, s), "copy");
    cudaCheck(cudaMemsetAsync(sc, 0, 8 * sizeof(double), s), "memset");
    dot2Kernel<<< g, 256, 0, s >>>(r.ptr, r.ptr, r.ptr, z.ptr, n, sc + 2); // rr → sc[2], rz → sc[3]
    reduce(sc + 2, 2);
    cudaCheck(cudaMemcpyAsync(h, sc + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, s), "copy");
    cudaCheck(cudaStreamSynchronize(s), "sync");
    double rnorm = std::sqrt(h[0]);
    int    it    = 0;
    static const bool     trace = [] { const char* e = std::getenv("L3B_PCG_TRACE"); return e != nullptr and e[0] == '1'; }();
    std::vector< double > stamps;
    const auto            t_start = std::chrono::steady_clock::now();
    if (rnorm > tol and max_iters > 0)
    {
        cudaCheck(cudaMemcpyAsync(sc, sc + 3, sizeof(double), cudaMemcpyDeviceToDevice, s), "copy"); // rz
        // the operator may add p.Ap to sc[1] itself (matrix-free apply: the energy falls out of its quadrature-point stage)
        const auto applyAndDot = [&] {
            cudaCheck(cudaMemsetAsync(sc + 1, 0, 3 * sizeof(double), s), "memset");
            if (not apply(p.ptr, Ap.ptr, sc + 1))
                dot2Kernel<<< g, 256, 0, s >>>(p.ptr, Ap.ptr, nullptr, nullptr, n, sc + 1);
        };
        applyAndDot();
        while (true)
        {
            reduce(sc + 1, 1);
            cgUpdateKernel<<< g, 256, 0, s >>>(r.ptr, z.ptr, Ap.ptr, minv.ptr, n, sc, sc + 2);
            reduce(sc + 2, 2);
            cudaCheck(cudaMemcpyAsync(h, sc + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, s), "copy");
            cudaCheck(cudaEventRecord(read_back_ev, s), "event");
            ++it;
            const bool last = it >= max_iters;
            if (last)
                cgFinalKernel<<< g, 256, 0, s >>>(x, p.ptr, n, sc);
            else // queued before the convergence test is known: x update + next direction, and its operator apply
            {
                cgDirectionKernel<<< g, 256, 0, s >>>(x, p.ptr, z.ptr, n, sc);
                cudaCheck(cudaMemcpyAsync(sc, sc + 3, sizeof(double), cudaMemcpyDeviceToDevice, s), "copy");
                applyAndDot();
            }
            cudaCheck(cudaEventSynchronize(read_back_ev), "sync");
            if (trace)
                stamps.push_back(std::chrono::duration< double, std::milli >(std::chrono::steady_clock::now() - t_start).count());
            rnorm = std::sqrt(h[0]);
            if (last or not(rnorm > tol))
                break;
        }
        if (trace and stamps.size() > 2)
        {
            // L3B_PCG_TRACE=1: distribution of the host-observed iteration periods (diagnostics for run-to-run spread)
            std::vector< double > d;
            for (size_t i = 1; i < stamps.size(); ++i)
                d.push_back(stamps[i] - stamps[i - 1]);
            std::vector< double > sorted = d;
            std::sort(sorted.begin(), sorted.end());
            const auto pct = [&](double q) { return sorted[static_cast< size_t >(q * (sorted.size() - 1))]; };
            std::fprintf(stderr, "[l3b pcg trace] %zu iterations: period min %.3f, p10 %.3f, median %.3f, p90 %.3f, p99 %.3f, max %.3f ms; first 8:", d.size(),
                         sorted.front(), pct(.1), pct(.5), pct(.9), pct(.99), sorted.back());
            for (size_t i = 0; i < std::min< size_t >(8, d.size()); ++i)
                std::fprintf(stderr, " %.3f", d[i]);
            std::fprintf(stderr, "; 100-iteration means:");
            for (size_t i = 0; i + 100 <= d.size(); i += 100)
            {
                double m = 0.;
                for (size_t k = i; k < i + 100; ++k)
                    m += d[k];
                std::fprintf(stderr, " %.3f", m / 100);
            }
            std::fprintf(stderr, "\n");
        }
        cudaCheck(cudaStreamSynchronize(s), "sync"); // the speculative tail reads buffers this function owns
    }
    cudaCheck(cudaGetLastError(), "pcg");
    *achieved = rnorm;
    *iters    = it;
}
// Restarted GMRES, left-preconditioned with native Jacobi — Belos "Pseudoblock GMRES" as the reference configures it
// (solve/BelosSolvers.hpp:42-131: "Num Blocks" = restart_length, "Maximum Restarts", left preconditioner, absolute residual by
// default), x0 = 0. Orthogonalisation: classical Gram-Schmidt applied twice (two fused dot/update kernel pairs per iteration — the
// all-reduce friendly form; Belos' default DGKS is the same idea). The convergence test is Belos' implicit one: the norm of the
// preconditioned residual from the Givens recurrence, against `tol`.
template < typename Apply, typename Reduce >
void gmres(l3b_context* ctx, long long n_local, long long n, Apply&& apply, Reduce&& reduce, const double* diag, const double* b,
           double* x /* device */, double tol, int restart, int max_restarts, int max_iters, double* achieved, int* iters, bool x0_zero = true)
{
    const int        m = std::max(1, restart);
    DevBuf< double > V(static_cast< size_t >(m + 1) * n), w(n_local), xin(n_local), minv(n), hd(m + 2);
    const auto       s = ctx->stream;
    const auto       g = gridFor(n);
    std::vector< double > H(static_cast< size_t >(m + 1) * m, 0.), cs(m), sn(m), gvec(m + 1), hcol(m + 2), y(m);
    jacobiInvertKernel<<< g, 256, 0, s >>>(diag, minv.ptr, n, 1., 0.);
    if (x0_zero) // otherwise x is the initial guess (Belos: LinearProblem(A, x, b) with the system's persistent solution)
        cudaCheck(cudaMemsetAsync(x, 0, n * sizeof(double), s), "memset");
    xin.zero(s);
    w.zero(s);
    int    it = 0, restarts = 0;
    double res = 0.;
    const auto dotsTo = [&](int cnt, const double* vec) { // hd[0..cnt) = V_i . vec, all-reduced, copied to hcol
        cudaCheck(cudaMemsetAsync(hd.ptr, 0, (m + 2) * sizeof(double), s), "memset");
        gsDotsKernel<<< g, 256, 0, s >>>(V.ptr, n, cnt, vec, n, hd.ptr);
        reduce(hd.ptr, cnt);
    };
    while (true)
    {
        // r = M^-1 (b - A x)
        cudaCheck(cudaMemcpyAsync(xin.ptr, x, n * sizeof(double), cudaMemcpyDeviceToDevice, s), "copy");
        apply(xin.ptr, w.ptr);
        precResidualKernel<<< g, 256, 0, s >>>(w.ptr, b, minv.ptr, n); // w = M^-1 (b - A x)
        cudaCheck(cudaMemsetAsync(hd.ptr, 0, (m + 2) * sizeof(double), s), "memset");
        dot2Kernel<<< g, 256, 0, s >>>(w.ptr, w.ptr, nullptr, nullptr, n, hd.ptr);
        reduce(hd.ptr, 1);
        double beta2 = 0.;
        cudaCheck(cudaMemcpyAsync(&beta2, hd.ptr, sizeof(double), cudaMemcpyDeviceToHost, s), "copy");
        cudaCheck(cudaStreamSynchronize(s), "sync");
        const double beta = std::sqrt(beta2);
        res               = beta;
        if (not(beta > tol) or it >= max_iters or restarts > max_restarts)
            break;
        scaledCopyKernel<<< g, 256, 0, s >>>(V.ptr, w.ptr, 1. / beta, n);
        std::fill(gvec.begin(), gvec.end(), 0.);
        gvec[0] = beta;
        int j   = 0;
        for (; j < m and it < max_iters; ++j)
        {
            // w = M^-1 A v_j
            cudaCheck(cudaMemcpyAsync(xin.ptr, V.ptr + static_cast< size_t >(j) * n, n * sizeof(double), cudaMemcpyDeviceToDevice, s), "copy");
            apply(xin.ptr, w.ptr);
            hadamardKernel<<< g, 256, 0, s >>>(w.ptr, minv.ptr, w.ptr, n);
            std::fill(hcol.begin(), hcol.end(), 0.);
            for (int pass = 0; pass < 2; ++pass) // classical Gram-Schmidt, twice
            {
                dotsTo(j + 1, w.ptr);
                gsAxpyKernel<<< g, 256, 0, s >>>(V.ptr, n, j + 1, hd.ptr, -1., w.ptr, n);
                std::vector< double > part(j + 1);
                cudaCheck(cudaMemcpyAsync(part.data(), hd.ptr, (j + 1) * sizeof(double), cudaMemcpyDeviceToHost, s), "copy");
                cudaCheck(cudaStreamSynchronize(s), "sync");
                for (int i = 0; i <= j; ++i)
                    hcol[i] += part[i];
            }
            cudaCheck(cudaMemsetAsync(hd.ptr, 0, (m + 2) * sizeof(double), s), "memset");
            dot2Kernel<<< g, 256, 0, s >>>(w.ptr, w.ptr, nullptr, nullptr, n, hd.ptr);
            reduce(hd.ptr, 1);
            double nw2 = 0.;
            cudaCheck(cudaMemcpyAsync(&nw2, hd.ptr, sizeof(double), cudaMemcpyDeviceToHost, s), "copy");
            cudaCheck(cudaStreamSynchronize(s), "sync");
            hcol[j + 1] = std::sqrt(nw2);
            if (hcol[j + 1] > 0.)
                scaledCopyKernel<<< g, 256, 0, s >>>(V.ptr + static_cast< size_t >(j + 1) * n, w.ptr, 1. / hcol[j + 1], n);
            // Givens rotations on the new column
            for (int i = 0; i < j; ++i)
            {
                const double t = cs[i] * hcol[i] + sn[i] * hcol[i + 1];
                hcol[i + 1]    = -sn[i] * hcol[i] + cs[i] * hcol[i + 1];
                hcol[i]        = t;
            }
            const double d = std::hypot(hcol[j], hcol[j + 1]);
            cs[j]          = d > 0. ? hcol[j] / d : 1.;
            sn[j]          = d > 0. ? hcol[j + 1] / d : 0.;
            hcol[j]        = d;
            hcol[j + 1]    = 0.;
            gvec[j + 1]    = -sn[j] * gvec[j];
            gvec[j]        = cs[j] * gvec[j];
            for (int i = 0; i <= j; ++i)
                H[static_cast< size_t >(i) * m + j] = hcol[i];
            ++it;
            res = std::fabs(gvec[j + 1]);
            if (not(res > tol))
            {
                ++j;
                break;
            }
        }
        // x += V_j y, R y = g
        for (int i = j - 1; i >= 0; --i)
        {
            double acc = gvec[i];
            for (int k = i + 1; k < j; ++k)
                acc -= H[static_cast< size_t >(i) * m + k] * y[k];
            y[i] = acc / H[static_cast< size_t >(i) * m + i];
        }
        hd.upload(y.data(), j, s);
        gsAxpyKernel<<< g, 256, 0, s >>>(V.ptr, n, j, hd.ptr, 1., x, n);
        cudaCheck(cudaStreamSynchronize(s), "sync");
        if (not(res > tol) or it >= max_iters)
            break;
        ++restarts;
        if (restarts > max_restarts)
            break;
    }
    cudaCheck(cudaGetLastError(), "gmres");
    *achieved = res;
    *iters    = it;
}

bool hostAllZero(const double* x, long long n)
{
    for (long long i = 0; i < n; ++i)
        if (x[i] != 0.)
            return false;
    return true;
}
// ghost copies of a solution vector follow their owners (so that the whole local vector is valid on return)
void refreshGhosts(l3b_halo* h, double* x)
{
    haloImportBegin(h, x, 1);
    haloImportEnd(h);
}
// y = A x on the assembled rows. With a halo and rows [owned | ghost] kept as the rank's elements assembled them: Import x, local product,
// Export-sum of the ghost rows; after l3b_asm_export_shared_rows the owned rows are complete and the Export is skipped.
void asmSpmvDevice(l3b_asm* sys, const double* x, double* y)
{
    auto* const     h       = sys->halo;
    const long long threads = sys->n_dofs * 32;
    haloImportBegin(h, const_cast< double* >(x), 1);
    haloImportEnd(h);
    spmvKernel<<< blocksFor(threads), 256, 0, sys->ctx->stream >>>(sys->node_ptr.ptr, sys->node_nbr.ptr, sys->values.ptr, sys->n_nodes, sys->dpn, x, y);
    if (not sys->rows_exported)
    {
        haloExportBegin(h, y, 1);
        haloExportEnd(h, y, 1);
    }
    cudaCheck(cudaGetLastError(), "spmv");
}
// diagonal of the global operator over the local rows (owned part valid): the Export-sum of the local diagonals
void asmDiagDevice(l3b_asm* sys, double* diag)
{
    extractDiagKernel<<< blocksFor(sys->n_dofs), 256, 0, sys->ctx->stream >>>(sys->node_ptr.ptr, sys->node_nbr.ptr, sys->values.ptr, sys->n_nodes,
                                                                            sys->dpn, diag);
    if (not sys->rows_exported)
    {
        haloExportBegin(sys->halo, diag, 1);
        haloExportEnd(sys->halo, diag, 1);
    }
    cudaCheck(cudaGetLastError(), "diagonal");
}
// right-hand side of the global system over the owned rows: the Export-sum of the local one (a copy: the system's rhs stays as assembled)
void asmGlobalRhs(l3b_asm* sys, double* rhs)
{
    cudaCheck(cudaMemcpyAsync(rhs, sys->rhs.ptr, sys->n_dofs * sizeof(double), cudaMemcpyDeviceToDevice, sys->ctx->stream), "copy");
    if (not sys->rows_exported)
    {
        haloExportBegin(sys->halo, rhs, 1);
        haloExportEnd(sys->halo, rhs, 1);
    }
}
} // namespace

extern "C"
{
// ---- context
int l3b_context_create(int device, l3b_context** out)
{
    return guardedCtx(nullptr, [&] {
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess or n_dev == 0)
            fail(L3B_ERR_NO_DEVICE, "no CUDA device available: l3ster_b200 has no CPU fallback");
        if (device < 0 or device >= n_dev)
            fail(L3B_ERR_INVALID_ARG, "invalid device ordinal");
        cudaCheck(cudaSetDevice(device), "cudaSetDevice");
        auto ctx    = std::make_unique< l3b_context >();
        ctx->device = device;
        cudaCheck(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "stream create");
        ctx->status.alloc(1);
        ctx->status.zero(ctx->stream);
        cudaDeviceProp prop{};
        cudaCheck(cudaGetDeviceProperties(&prop, device), "device properties");
        ctx->sm_count = prop.multiProcessorCount;
        *out          = ctx.release();
    });
}
void l3b_context_destroy(l3b_context* ctx)
{
    if (not ctx)
        return;
    cudaSetDevice(ctx->device);
    ctx->dense.clear();
    ctx->node_tabs.clear();
    ctx->status.release();
    ctx->scalars.release();
    ctx->krylov_ws.release();
    if (ctx->pinned_scalars)
        cudaFreeHost(ctx->pinned_scalars);
    if (ctx->krylov_event)
        cudaEventDestroy(ctx->krylov_event);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char* l3b_last_error(const l3b_context* ctx)
{
    return ctx ? ctx->err.c_str() : g_global_error.c_str();
}
const char* l3b_global_error(void)
{
    return g_global_error.c_str();
}
int l3b_context_synchronize(l3b_context* ctx)
{
    return guardedCtx(ctx, [&] { cudaCheck(cudaStreamSynchronize(ctx->stream), "synchronize"); });
}
void* l3b_context_stream(l3b_context* ctx)
{
    return ctx->stream;
}

// ---- communicator and halo (comm.cuh)
int l3b_comm_unique_id(char id[L3B_COMM_ID_BYTES])
{
    return guardedCtx(nullptr, [&] {
        static_assert(sizeof(ncclUniqueId) == L3B_COMM_ID_BYTES);
        ncclUniqueId uid;
        ncclCheck(ncclApi().GetUniqueId(&uid), "ncclGetUniqueId");
        std::memcpy(id, &uid, sizeof(uid));
    });
}
static void commFinishSetup(l3b_comm* c)
{
    cudaCheck(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream create");
    cudaCheck(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming), "event create");
    cudaCheck(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming), "event create");
}
int l3b_comm_create(l3b_context* ctx, int rank, int world, const char id[L3B_COMM_ID_BYTES], l3b_comm** out)
{
    return guardedCtx(ctx, [&] {
        if (world < 1 or rank < 0 or rank >= world)
            fail(L3B_ERR_INVALID_ARG, "l3b_comm_create: invalid rank / world size");
        auto c   = std::make_unique< l3b_comm >();
        c->ctx   = ctx;
        c->rank  = rank;
        c->world = world;
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        ncclCheck(ncclApi().CommInitRank(&c->comm, world, uid, rank), "ncclCommInitRank");
        c->owned = true;
        commFinishSetup(c.get());
        *out = c.release();
    });
}
int l3b_comm_attach(l3b_context* ctx, void* nccl_comm, l3b_comm** out)
{
    return guardedCtx(ctx, [&] {
        if (nccl_comm == nullptr)
            fail(L3B_ERR_INVALID_ARG, "l3b_comm_attach: null communicator");
        auto c  = std::make_unique< l3b_comm >();
        c->ctx  = ctx;
        c->comm = static_cast< ncclComm_t >(nccl_comm);
        ncclCheck(ncclApi().CommCount(c->comm, &c->world), "ncclCommCount");
        ncclCheck(ncclApi().CommUserRank(c->comm, &c->rank), "ncclCommUserRank");
        commFinishSetup(c.get());
        *out = c.release();
    });
}
void l3b_comm_destroy(l3b_comm* c)
{
    if (c == nullptr)
        return;
    const DeviceGuard guard{c->ctx};
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    delete c;
}
int l3b_comm_rank(const l3b_comm* c)
{
    return c ? c->rank : 0;
}
int l3b_comm_size(const l3b_comm* c)
{
    return c ? c->world : 1;
}
int l3b_comm_allreduce_sum(l3b_comm* c, double* device_scalars, int n)
{
    return guardedCtx(c ? c->ctx : nullptr, [&] { commAllReduce(c, device_scalars, n); });
}
int l3b_halo_create(l3b_comm* c, int64_t n_owned, int64_t n_ghost, int n_owned_nbrs, const int* owned_nbr_ranks, const int64_t* owned_ptr,
                    const int32_t* owned_inds, int n_shared_nbrs, const int* shared_nbr_ranks, const int64_t* shared_offsets, l3b_halo** out)
{
    return guardedCtx(c ? c->ctx : nullptr, [&] {
        if (c == nullptr)
            fail(L3B_ERR_COMM, "l3b_halo_create needs a communicator");
        if (n_owned < 0 or n_ghost < 0 or n_owned_nbrs < 0 or n_shared_nbrs < 0)
            fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: negative size");
        auto h     = std::make_unique< l3b_halo >();
        h->comm    = c;
        h->n_owned = n_owned;
        h->n_ghost = n_ghost;
        h->owned_nbrs.assign(owned_nbr_ranks, owned_nbr_ranks + n_owned_nbrs);
        h->shared_nbrs.assign(shared_nbr_ranks, shared_nbr_ranks + n_shared_nbrs);
        h->owned_ptr.assign(1, 0);
        h->shared_off.assign(1, 0);
        if (n_owned_nbrs > 0)
            h->owned_ptr.assign(owned_ptr, owned_ptr + n_owned_nbrs + 1);
        if (n_shared_nbrs > 0)
            h->shared_off.assign(shared_offsets, shared_offsets + n_shared_nbrs + 1);
        for (int r : h->owned_nbrs)
            if (r < 0 or r >= c->world)
                fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: neighbour rank out of range");
        for (int r : h->shared_nbrs)
            if (r < 0 or r >= c->world)
                fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: neighbour rank out of range");
        if (h->owned_ptr.front() != 0 or h->shared_off.front() != 0 or h->shared_off.back() > n_ghost)
            fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: offsets do not fit the dof layout");
        for (size_t k = 0; k + 1 < h->owned_ptr.size(); ++k)
            if (h->owned_ptr[k] > h->owned_ptr[k + 1])
                fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: offsets must ascend");
        for (size_t k = 0; k + 1 < h->shared_off.size(); ++k)
            if (h->shared_off[k] > h->shared_off[k + 1])
                fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: offsets must ascend");
        const long long n_idx = h->owned_ptr.back();
        for (long long i = 0; i < n_idx; ++i)
            if (owned_inds[i] < 0 or owned_inds[i] >= n_owned)
                fail(L3B_ERR_INVALID_ARG, "l3b_halo_create: a packed index is not an owned dof");
        const auto S = c->ctx->stream;
        h->owned_idx.alloc(std::max< long long >(n_idx, 1));
        h->owned_idx.upload(owned_inds, n_idx, S);
        h->owned_ptr_dev.alloc(h->owned_ptr.size());
        h->owned_ptr_dev.upload(h->owned_ptr.data(), h->owned_ptr.size(), S);
        cudaCheck(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming), "event create");
        cudaCheck(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming), "event create");
        cudaCheck(cudaStreamSynchronize(S), "halo upload");
        *out = h.release();
    });
}
void l3b_halo_destroy(l3b_halo* h)
{
    if (h == nullptr)
        return;
    const DeviceGuard guard{h->comm->ctx};
    cudaStreamSynchronize(h->comm->stream);
    cudaStreamSynchronize(h->comm->ctx->stream);
    delete h;
}
int l3b_halo_import_begin(l3b_halo* h, double* x, int n_cols)
{
    return guardedCtx(h->comm->ctx, [&] { haloImportBegin(h, x, n_cols); });
}
int l3b_halo_import_end(l3b_halo* h)
{
    return guardedCtx(h->comm->ctx, [&] { haloImportEnd(h); });
}
int l3b_halo_export_begin(l3b_halo* h, double* y, int n_cols)
{
    return guardedCtx(h->comm->ctx, [&] { haloExportBegin(h, y, n_cols); });
}
int l3b_halo_export_end(l3b_halo* h, double* y, int n_cols)
{
    return guardedCtx(h->comm->ctx, [&] { haloExportEnd(h, y, n_cols); });
}
int l3b_mf_set_halo(l3b_mf* sys, l3b_halo* halo, int64_t n_border_elems)
{
    return guardedCtx(sys->ctx, [&] {
        if (halo != nullptr and (halo->n_owned != sys->ownedDofs() or halo->n_owned + halo->n_ghost != sys->n_dofs))
            fail(L3B_ERR_INVALID_ARG, "l3b_mf_set_halo: the halo does not describe this system's dof layout [owned | ghost]");
        if (n_border_elems < 0 or n_border_elems > sys->mesh->n_elems)
            fail(L3B_ERR_INVALID_ARG, "l3b_mf_set_halo: border element count out of range");
        sys->halo     = halo;
        sys->n_border = n_border_elems;
        sys->pipe.reset(); // the schedule of the streamed host-buffer apply depends on both
    });
}
int l3b_asm_set_halo(l3b_asm* sys, l3b_halo* halo)
{
    return guardedCtx(sys->ctx, [&] {
        if (halo != nullptr and halo->n_owned + halo->n_ghost != sys->n_dofs)
            fail(L3B_ERR_INVALID_ARG, "l3b_asm_set_halo: the halo does not describe this system's row layout [owned | ghost]");
        sys->halo         = halo;
        sys->n_owned_dofs = halo ? halo->n_owned : -1;
    });
}

// ---- registry
int l3b_kernel_count(void)
{
    return static_cast< int >(kernelRegistry().size());
}
int l3b_kernel_find(const char* name)
{
    return findKernel(name);
}
int l3b_kernel_get_info(int id, l3b_kernel_info* out)
{
    const auto& r = kernelRegistry();
    if (id < 0 or id >= static_cast< int >(r.size()))
        return L3B_ERR_INVALID_ARG;
    const auto& i = r[id].info;
    std::memset(out, 0, sizeof(*out));
    std::strncpy(out->name, i.name.c_str(), sizeof(out->name) - 1);
    out->dimension   = i.dim;
    out->n_equations = i.n_equations;
    out->n_unknowns  = i.n_unknowns;
    out->n_fields    = i.n_fields;
    out->n_rhs       = i.n_rhs;
    out->is_boundary = i.is_boundary;
    out->is_residual = i.is_residual;
    out->n_instances = static_cast< int >(r[id].instances.size());
    return L3B_OK;
}
int l3b_kernel_get_instance(int id, int i, int* order, int* nq)
{
    const auto& r = kernelRegistry();
    if (id < 0 or id >= static_cast< int >(r.size()) or i < 0 or i >= static_cast< int >(r[id].instances.size()))
        return L3B_ERR_INVALID_ARG;
    *order = r[id].instances[i].order;
    *nq    = r[id].instances[i].nq;
    return L3B_OK;
}

// ---- tables
int l3b_tables_gll(int n, double* nodes)
{
    return guardedCtx(nullptr, [&] {
        const auto x = tables::gllNodes(n);
        for (int i = 0; i < n; ++i)
            nodes[i] = static_cast< double >(x[i]);
    });
}
int l3b_tables_gauss(int n, double* points, double* weights)
{
    return guardedCtx(nullptr, [&] {
        std::vector< tables::ld > x, w;
        tables::glRule(n, x, w);
        for (int i = 0; i < n; ++i)
        {
            points[i]  = static_cast< double >(x[i]);
            weights[i] = static_cast< double >(w[i]);
        }
    });
}
int l3b_tables_1d(int order, int nq, double* interp, double* der, double* colloc)
{
    return guardedCtx(nullptr, [&] {
        const auto t = tables::makeTables1D(order, nq);
        std::copy(t.interp.begin(), t.interp.end(), interp);
        std::copy(t.der.begin(), t.der.end(), der);
        if (colloc)
            std::copy(t.colloc.begin(), t.colloc.end(), colloc);
    });
}
int l3b_tables_dense(int dim, int order, int nq, int side, int* n_qp, double* points, double* weights, double* values, double* derivatives)
{
    return guardedCtx(nullptr, [&] {
        std::vector< tables::ld > p, w;
        if (side < 0)
            tables::domainQuadrature(dim, nq, p, w);
        else
            tables::sideQuadrature(dim, nq, side, p, w);
        const auto t = tables::makeDenseTables(dim, order, p, w);
        *n_qp        = t.n_qp;
        if (points)
            std::copy(t.points.begin(), t.points.end(), points);
        if (weights)
            std::copy(t.weights.begin(), t.weights.end(), weights);
        if (values)
            std::copy(t.values.begin(), t.values.end(), values);
        if (derivatives)
            std::copy(t.derivatives.begin(), t.derivatives.end(), derivatives);
    });
}

// ---- host mesh
int l3b_host_mesh_cube(int nx, const double* x, int ny, const double* y, int nz, const double* z, int order, l3b_host_mesh** out)
{
    return guardedCtx(nullptr, [&] {
        if (order < 1)
            fail(L3B_ERR_INVALID_ARG, "order must be >= 1");
        *out = new l3b_host_mesh{host::makeStructured(3, {x, x + nx}, {y, y + ny}, {z, z + nz}, order)};
    });
}
int l3b_host_mesh_square(int nx, const double* x, int ny, const double* y, int order, l3b_host_mesh** out)
{
    return guardedCtx(nullptr, [&] {
        if (order < 1)
            fail(L3B_ERR_INVALID_ARG, "order must be >= 1");
        *out = new l3b_host_mesh{host::makeStructured(2, {x, x + nx}, {y, y + ny}, {}, order)};
    });
}
void l3b_host_mesh_destroy(l3b_host_mesh* m)
{
    delete m;
}
int l3b_host_mesh_info(const l3b_host_mesh* m, int64_t info[6])
{
    info[0] = m->mesh.dim;
    info[1] = m->mesh.order;
    info[2] = m->mesh.n_nodes;
    info[3] = m->mesh.n_elems;
    info[4] = m->mesh.nodes_per_elem;
    info[5] = m->mesh.n_sides;
    return L3B_OK;
}
const uint32_t* l3b_host_mesh_nodes(const l3b_host_mesh* m)
{
    return m->mesh.nodes.data();
}
const double* l3b_host_mesh_verts(const l3b_host_mesh* m)
{
    return m->mesh.verts.data();
}
const uint16_t* l3b_host_mesh_side_boundaries(const l3b_host_mesh* m)
{
    return m->mesh.side_bnd.data();
}
int l3b_node_graph(int64_t n_nodes, int64_t n_elems, int nn, const uint32_t* nodes, int64_t** ptr, uint32_t** nbr)
{
    return guardedCtx(nullptr, [&] {
        const auto g = host::makeNodeGraph(n_nodes, n_elems, nn, nodes);
        *ptr         = static_cast< int64_t* >(std::malloc(g.ptr.size() * sizeof(int64_t)));
        *nbr         = static_cast< uint32_t* >(std::malloc(std::max< size_t >(1, g.nbr.size()) * sizeof(uint32_t)));
        if (not *ptr or not *nbr)
            fail(L3B_ERR_INVALID_ARG, "out of host memory");
        for (size_t i = 0; i < g.ptr.size(); ++i)
            (*ptr)[i] = g.ptr[i];
        std::copy(g.nbr.begin(), g.nbr.end(), *nbr);
    });
}
int l3b_graph_expand(int64_t n_nodes, const int64_t* ptr, const uint32_t* nbr, int dpn, int64_t* row_ptr, int32_t* col_ind)
{
    for (int64_t n = 0; n < n_nodes; ++n)
    {
        const int64_t deg = ptr[n + 1] - ptr[n];
        for (int d = 0; d < dpn; ++d)
        {
            const int64_t beg    = dpn * (dpn * ptr[n] + d * deg);
            row_ptr[n * dpn + d] = beg;
            if (col_ind)
                for (int64_t k = 0; k < deg; ++k)
                    for (int v = 0; v < dpn; ++v)
                        col_ind[beg + k * dpn + v] = static_cast< int32_t >(nbr[ptr[n] + k]) * dpn + v;
        }
    }
    row_ptr[n_nodes * dpn] = ptr[n_nodes] * dpn * dpn;
    return L3B_OK;
}
void l3b_free(void* p)
{
    std::free(p);
}

// ---- device mesh
int l3b_mesh_upload(l3b_context* ctx, int dim, int order, int64_t n_elems, const double* verts, const uint32_t* nodes,
                    const uint16_t* side_boundaries, int64_t n_local_nodes, int64_t n_owned_nodes, l3b_mesh** out)
{
    return guardedCtx(ctx, [&] {
        if (dim != 2 and dim != 3)
            fail(L3B_ERR_INVALID_ARG, "only quad (dim 2) and hex (dim 3) elements are supported");
        auto m           = std::make_unique< l3b_mesh >();
        m->ctx           = ctx;
        m->dim           = dim;
        m->order         = order;
        m->nn            = cpow(order + 1, dim);
        m->n_sides       = 2 * dim;
        m->n_elems       = n_elems;
        m->n_local_nodes = n_local_nodes;
        m->n_owned_nodes = n_owned_nodes;
        const size_t nv  = static_cast< size_t >(n_elems) * (1 << dim) * 3;
        m->verts.alloc(nv);
        m->verts.upload(verts, nv, ctx->stream);
        m->nodes.alloc(static_cast< size_t >(n_elems) * m->nn);
        m->nodes.upload(nodes, static_cast< size_t >(n_elems) * m->nn, ctx->stream);
        if (side_boundaries)
            m->side_bnd.assign(side_boundaries, side_boundaries + n_elems * m->n_sides);
        if (dim == 3 and n_elems > 0)
        {
            m->hex_geo.alloc(static_cast< size_t >(n_elems) * hex_geo_doubles);
            hexGeometryKernel<<< gridFor(n_elems), 256, 0, ctx->stream >>>(m->verts.ptr, n_elems, m->hex_geo.ptr);
            cudaCheck(cudaGetLastError(), "hex geometry");
        }
        cudaCheck(cudaStreamSynchronize(ctx->stream), "mesh upload");
        *out = m.release();
    });
}
int l3b_mesh_update_verts(l3b_mesh* mesh, const double* verts)
{
    return guardedCtx(mesh->ctx, [&] {
        auto* ctx = mesh->ctx;
        mesh->verts.upload(verts, mesh->verts.n, ctx->stream);
        if (mesh->dim == 3 and mesh->n_elems > 0)
        {
            hexGeometryKernel<<< gridFor(mesh->n_elems), 256, 0, ctx->stream >>>(mesh->verts.ptr, mesh->n_elems, mesh->hex_geo.ptr);
            cudaCheck(cudaGetLastError(), "hex geometry");
        }
        cudaCheck(cudaStreamSynchronize(ctx->stream), "vertex upload"); // the host buffer may be reused on return
    });
}
void l3b_mesh_destroy(l3b_mesh* mesh)
{
    delete mesh;
}

// ---- fields
int l3b_fields_upload(l3b_context* ctx, int64_t n_local_nodes, int n_fields, const double* data, l3b_fields** out)
{
    return guardedCtx(ctx, [&] {
        auto f      = std::make_unique< l3b_fields >();
        f->ctx      = ctx;
        f->n_nodes  = n_local_nodes;
        f->n_fields = n_fields;
        f->data.alloc(static_cast< size_t >(n_local_nodes) * n_fields);
        f->data.upload(data, f->data.n, ctx->stream);
        cudaCheck(cudaStreamSynchronize(ctx->stream), "fields upload");
        *out = f.release();
    });
}
int l3b_fields_update(l3b_fields* f, const double* data)
{
    return guardedCtx(f->ctx, [&] {
        f->data.upload(data, f->data.n, f->ctx->stream);
        cudaCheck(cudaStreamSynchronize(f->ctx->stream), "fields upload");
    });
}
void l3b_fields_destroy(l3b_fields* f)
{
    delete f;
}

// ---- assembled system
int l3b_asm_create(l3b_context* ctx, l3b_mesh* mesh, int dpn, int n_rhs, const int64_t* node_ptr, const uint32_t* node_nbr, l3b_asm** out)
{
    return guardedCtx(ctx, [&] {
        auto s       = std::make_unique< l3b_asm >();
        s->ctx       = ctx;
        s->mesh      = mesh;
        s->dpn       = dpn;
        s->n_rhs     = n_rhs;
        const auto N = mesh->n_local_nodes;
        s->n_nodes   = N;
        s->n_dofs    = N * dpn;
        s->nnz       = node_ptr[N] * dpn * dpn;
        s->node_ptr.alloc(N + 1);
        s->node_nbr.alloc(node_ptr[N]);
        static_assert(sizeof(long long) == sizeof(int64_t));
        s->node_ptr.upload(reinterpret_cast< const long long* >(node_ptr), N + 1, ctx->stream);
        s->node_ptr_host.assign(node_ptr, node_ptr + N + 1);
        s->node_nbr.upload(node_nbr, node_ptr[N], ctx->stream);
        s->row_ptr.alloc(s->n_dofs + 1);
        rowPtrKernel<<< blocksFor(s->n_dofs + 1), 256, 0, ctx->stream >>>(s->node_ptr.ptr, N, dpn, s->row_ptr.ptr);
        s->values.alloc(s->nnz);
        s->rhs.alloc(s->n_dofs * n_rhs);
        const long long n_pos = mesh->n_elems * mesh->nn * mesh->nn;
        s->slot_pos.alloc(n_pos);
        slotMapKernel<<< blocksFor(n_pos), 256, 0, ctx->stream >>>(mesh->nodes.ptr, mesh->n_elems, mesh->nn,
                                                                                           s->node_ptr.ptr, s->node_nbr.ptr,
                                                                                           s->slot_pos.ptr, ctx->status.ptr);
        cudaCheck(cudaGetLastError(), "slot map");
        ctx->checkStatus();
        cudaCheck(cudaEventCreate(&s->ev0), "event");
        cudaCheck(cudaEventCreate(&s->ev1), "event");
        *out = s.release();
    });
}
// the same storage without a mesh: rows of `n_nodes` nodes, filled by l3b_condense (the condensed system of static condensation)
int l3b_crs_create(l3b_context* ctx, int64_t n_nodes, int dpn, int n_rhs, const int64_t* node_ptr, const uint32_t* node_nbr, l3b_asm** out)
{
    return guardedCtx(ctx, [&] {
        auto s     = std::make_unique< l3b_asm >();
        s->ctx     = ctx;
        s->dpn     = dpn;
        s->n_rhs   = n_rhs;
        s->n_nodes = n_nodes;
        s->n_dofs  = n_nodes * dpn;
        s->nnz     = node_ptr[n_nodes] * dpn * dpn;
        s->node_ptr.alloc(n_nodes + 1);
        s->node_nbr.alloc(node_ptr[n_nodes]);
        s->node_ptr.upload(reinterpret_cast< const long long* >(node_ptr), n_nodes + 1, ctx->stream);
        s->node_ptr_host.assign(node_ptr, node_ptr + n_nodes + 1);
        s->node_nbr.upload(node_nbr, node_ptr[n_nodes], ctx->stream);
        s->row_ptr.alloc(s->n_dofs + 1);
        rowPtrKernel<<< blocksFor(s->n_dofs + 1), 256, 0, ctx->stream >>>(s->node_ptr.ptr, n_nodes, dpn, s->row_ptr.ptr);
        s->values.alloc(s->nnz);
        s->rhs.alloc(s->n_dofs * n_rhs);
        cudaCheck(cudaGetLastError(), "crs create");
        cudaCheck(cudaEventCreate(&s->ev0), "event");
        cudaCheck(cudaEventCreate(&s->ev1), "event");
        *out = s.release();
    });
}
void l3b_asm_destroy(l3b_asm* sys)
{
    delete sys;
}
int64_t l3b_asm_nnz(const l3b_asm* sys)
{
    return sys->nnz;
}
int l3b_asm_begin_assembly(l3b_asm* sys)
{
    return guardedCtx(sys->ctx, [&] {
        sys->values.zero(sys->ctx->stream);
        sys->rhs.zero(sys->ctx->stream);
        sys->open          = true;
        sys->rows_exported = false;
    });
}
int l3b_asm_assemble(l3b_asm* sys, int kernel_id, l3b_asm_opts opts, double time, const int* dof_inds, const l3b_fields* fields,
                     const int* field_inds, const int* boundary_ids, int n_boundary_ids)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"assembleGlobalSystem"}; // AssembleGlobalSystem.hpp:32, 74
        if (not sys->open)
            fail(L3B_ERR_STATE, "`assembleProblem()` was called before `beginAssembly()`");
        if (sys->mesh == nullptr)
            fail(L3B_ERR_STATE, "this system has no mesh (l3b_crs_create): it takes contributions from l3b_condense only");
        const auto  use  = makeUse(sys->mesh, sys->dpn, sys->n_rhs, kernel_id, opts, time, dof_inds, fields, field_inds, boundary_ids, n_boundary_ids);
        const auto& info = kernelRegistry()[kernel_id].info;
        ElemArgs    a    = baseArgs(sys->mesh, use, sys->dpn, sys->n_dofs);
        setDense(a, sys->mesh, use, info.is_boundary);
        a.row_ptr  = sys->row_ptr.ptr;
        a.node_ptr = sys->node_ptr.ptr;
        a.slot_pos = sys->slot_pos.ptr;
        a.crs_vals = sys->values.ptr;
        a.rhs      = sys->rhs.ptr;
        cudaCheck(cudaEventRecord(sys->ev0, sys->ctx->stream), "event record");
        cudaCheck(use.inst->assemble(kernelRegistry()[kernel_id].object.get(), a, sys->ctx->stream), "assembly launch");
        cudaCheck(cudaEventRecord(sys->ev1, sys->ctx->stream), "event record");
        sys->ctx->checkStatus();
        float ms = 0.f;
        cudaCheck(cudaEventElapsedTime(&ms, sys->ev0, sys->ev1), "event elapsed");
        sys->last_ms = ms;
    });
}
int l3b_asm_end_assembly(l3b_asm* sys, int64_t n_dir, const int32_t* dofs, const double* vals)
{
    return l3b_asm_end_assembly_ranked(sys, n_dir, dofs, vals, sys->n_dofs);
}
int l3b_asm_end_assembly_ranked(l3b_asm* sys, int64_t n_dir, const int32_t* dofs, const double* vals, int64_t n_owned_dofs)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"Dirichlet BCs"}; // AssembledSystem.hpp:349-353 (inside endAssembly, :376)
        if (not sys->open)
            fail(L3B_ERR_STATE, "`endAssembly()` was called more than once");
        const bool close_inactive = sys->dofmap != nullptr and sys->dofmap->n_dofs < sys->n_dofs;
        if (n_dir > 0 or close_inactive)
        {
            std::vector< uint8_t > mask(sys->n_dofs, 0);
            std::vector< double >  bc(static_cast< size_t >(sys->n_dofs) * sys->n_rhs, 0.);
            if (close_inactive) // homogeneous "Dirichlet" rows: identity, rhs 0; their columns are empty anyway
                for (long long i = 0; i < sys->n_dofs; ++i)
                    mask[i] = sys->dofmap->active[i] ? 0 : 1;
            for (int64_t i = 0; i < n_dir; ++i)
            {
                if (dofs[i] < 0 or dofs[i] >= sys->n_dofs)
                    fail(L3B_ERR_INVALID_ARG, "Dirichlet dof out of range");
                mask[dofs[i]] = 1;
                for (int c = 0; c < sys->n_rhs; ++c)
                    bc[dofs[i] + static_cast< size_t >(c) * sys->n_dofs] = vals[i + c * n_dir];
            }
            DevBuf< uint8_t > d_mask(mask.size());
            DevBuf< double >  d_bc(bc.size());
            d_mask.upload(mask.data(), mask.size(), sys->ctx->stream);
            d_bc.upload(bc.data(), bc.size(), sys->ctx->stream);
            const long long threads = sys->n_dofs * 32;
            dirichletAlgebraicKernel<<< blocksFor(threads), 256, 0, sys->ctx->stream >>>(
                sys->node_ptr.ptr, sys->node_nbr.ptr, sys->values.ptr, sys->n_nodes, sys->dpn, d_mask.ptr, d_bc.ptr, sys->rhs.ptr,
                sys->n_dofs, sys->n_rhs, n_owned_dofs);
            cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "Dirichlet BC application");
        }
        sys->open = false;
    });
}
int l3b_asm_download(l3b_asm* sys, double* values, double* rhs)
{
    return guardedCtx(sys->ctx, [&] {
        if (values)
            sys->values.download(values, sys->nnz, sys->ctx->stream);
        if (rhs)
            sys->rhs.download(rhs, sys->rhs.n, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "download");
        if (values)
        {
            // device rows are column-dof-major (v * deg + k); the reference's Tpetra rows are node-major (k * dpn + v)
            const int             dpn = sys->dpn;
            std::vector< double > row;
            long long             beg = 0;
            for (long long n = 0; n < sys->n_nodes; ++n)
            {
                const long long deg = sys->node_ptr_host[n + 1] - sys->node_ptr_host[n];
                row.resize(static_cast< size_t >(deg) * dpn);
                for (int d = 0; d < dpn; ++d, beg += deg * dpn)
                {
                    std::copy(values + beg, values + beg + deg * dpn, row.begin());
                    for (long long k = 0; k < deg; ++k)
                        for (int v = 0; v < dpn; ++v)
                            values[beg + k * dpn + v] = row[v * deg + k];
                }
            }
        }
    });
}
double* l3b_asm_device_values(l3b_asm* sys)
{
    return sys->values.ptr;
}
int l3b_asm_spmv(l3b_asm* sys, const double* x, double* y)
{
    return guardedCtx(sys->ctx, [&] {
        DevBuf< double > dx(sys->n_dofs), dy(sys->n_dofs);
        dx.upload(x, sys->n_dofs, sys->ctx->stream);
        const long long threads = sys->n_dofs * 32;
        spmvKernel<<< blocksFor(threads), 256, 0, sys->ctx->stream >>>(
            sys->node_ptr.ptr, sys->node_nbr.ptr, sys->values.ptr, sys->n_nodes, sys->dpn, dx.ptr, dy.ptr);
        dy.download(y, sys->n_dofs, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "spmv");
    });
}
// method 0: CG, 1: restarted GMRES — both with native Jacobi; x on the device over the local rows, in = initial guess, out = solution
int l3b_asm_solve_device(l3b_asm* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, int x0_is_zero,
                         double* achieved_tol, int* iters)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"solve"}; // AssembledSystem::solve (AssembledSystem.hpp:115)
        if (sys->open)
            fail(L3B_ERR_STATE, "`solve()` was called before `endAssembly()`");
        if (method != 0 and method != 1)
            fail(L3B_ERR_INVALID_ARG, "solver method: 0 (CG) or 1 (GMRES)");
        const auto       n = sys->n_dofs;
        DevBuf< double > diag(std::max< long long >(n, 1)), rhs(std::max< long long >(n, 1));
        asmDiagDevice(sys, diag.ptr);
        asmGlobalRhs(sys, rhs.ptr);
        const auto reduce = [&](double* sc, int k) { commAllReduce(sys->comm(), sc, k); };
        if (method == 0)
            pcg(
                sys->ctx, n, sys->ownedDofs(),
                [&](const double* in, double* out, double*) {
                    asmSpmvDevice(sys, in, out);
                    return false; // p.Ap by a dot-product pass
                },
                reduce, diag.ptr, rhs.ptr, x, tol, max_iters, achieved_tol, iters, x0_is_zero != 0);
        else
            gmres(
                sys->ctx, n, sys->ownedDofs(), [&](const double* in, double* out) { asmSpmvDevice(sys, in, out); }, reduce, diag.ptr, rhs.ptr, x, tol,
                restart_length, max_restarts, max_iters, achieved_tol, iters, x0_is_zero != 0);
        refreshGhosts(sys->halo, x);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "solve"); // diag and rhs die here
    });
}
static int asmSolveHost(l3b_asm* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, double* achieved_tol,
                        int* iters)
{
    DevBuf< double > dx;
    const int        rc = guardedCtx(sys->ctx, [&] {
        dx.alloc(std::max< long long >(sys->n_dofs, 1));
        dx.upload(x, sys->n_dofs, sys->ctx->stream);
    });
    if (rc != L3B_OK)
        return rc;
    const int rc2 = l3b_asm_solve_device(sys, method, tol, max_iters, restart_length, max_restarts, dx.ptr, hostAllZero(x, sys->ownedDofs()),
                                         achieved_tol, iters);
    if (rc2 != L3B_OK)
        return rc2;
    return guardedCtx(sys->ctx, [&] {
        dx.download(x, sys->n_dofs, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "solve");
    });
}
int l3b_asm_solve_gmres(l3b_asm* sys, double tol, int restart_length, int max_restarts, int max_iters, double* x, double* achieved_tol,
                        int* iters)
{
    return asmSolveHost(sys, 1, tol, max_iters, restart_length, max_restarts, x, achieved_tol, iters);
}
int l3b_asm_spmv_device(l3b_asm* sys, const double* x, double* y)
{
    return guardedCtx(sys->ctx, [&] { asmSpmvDevice(sys, x, y); });
}
int l3b_asm_diag_device(l3b_asm* sys, double* diag)
{
    return guardedCtx(sys->ctx, [&] { asmDiagDevice(sys, diag); });
}
double* l3b_asm_device_rhs(l3b_asm* sys)
{
    return sys->rhs.ptr;
}
int l3b_asm_solve_cg(l3b_asm* sys, double tol, int max_iters, double* x, double* achieved_tol, int* iters)
{
    return asmSolveHost(sys, 0, tol, max_iters, 1, 0, x, achieved_tol, iters);
}
double l3b_asm_last_kernel_ms(const l3b_asm* sys)
{
    return sys->last_ms;
}

// ---- matrix-free system
int l3b_mf_create(l3b_context* ctx, l3b_mesh* mesh, int dpn, int n_rhs, const uint8_t* mask, const double* vals, l3b_mf** out)
{
    return guardedCtx(ctx, [&] {
        auto s    = std::make_unique< l3b_mf >();
        s->ctx    = ctx;
        s->mesh   = mesh;
        s->dpn    = dpn;
        s->n_rhs  = n_rhs;
        s->n_dofs = mesh->n_local_nodes * dpn;
        s->diag.alloc(s->n_dofs);
        s->rhs.alloc(s->n_dofs * n_rhs);
        if (mask)
        {
            s->has_bc = true;
            s->dir_mask.alloc(s->n_dofs);
            s->dir_mask.upload(mask, s->n_dofs, ctx->stream);
            // identity rows only for owned dofs: ghost rows are summed into their owner by the export (MatrixFreeSystem.hpp:1087-1103)
            std::vector< int32_t > list;
            for (long long i = 0; i < mesh->n_owned_nodes * dpn; ++i)
                if (mask[i])
                    list.push_back(static_cast< int32_t >(i));
            s->n_dir = static_cast< long long >(list.size());
            s->dir_list.alloc(list.size());
            s->dir_list.upload(list.data(), list.size(), ctx->stream);
            s->dir_list_host = list;
            s->dir_vals.alloc(s->n_dofs * n_rhs);
            if (vals)
                s->dir_vals.upload(vals, s->dir_vals.n, ctx->stream);
            else
                s->dir_vals.zero(ctx->stream);
            std::vector< double > g(static_cast< size_t >(s->n_dofs) * n_rhs, 0.);
            if (vals)
                for (int c = 0; c < n_rhs; ++c)
                    for (long long i = 0; i < s->n_dofs; ++i)
                        if (mask[i] and vals[i + c * s->n_dofs] != 0.)
                        {
                            g[i + c * s->n_dofs] = vals[i + c * s->n_dofs];
                            s->lift_needed       = true;
                        }
            if (s->lift_needed)
            {
                s->dir_g.alloc(g.size());
                s->dir_g.upload(g.data(), g.size(), ctx->stream);
                cudaCheck(cudaStreamSynchronize(ctx->stream), "Dirichlet values upload");
            }
            if (mesh->n_elems > 0)
            {
                s->elem_dir.alloc(mesh->n_elems);
                elemDirichletFlagKernel<<< gridFor(mesh->n_elems), 256, 0, ctx->stream >>>(mesh->nodes.ptr, mesh->n_elems, mesh->nn, dpn,
                                                                                       s->dir_mask.ptr, s->elem_dir.ptr);
                cudaCheck(cudaGetLastError(), "element Dirichlet flags");
            }
            cudaCheck(cudaStreamSynchronize(ctx->stream), "Dirichlet upload");
        }
        *out = s.release();
    });
}
void l3b_mf_destroy(l3b_mf* sys)
{
    delete sys;
}
int l3b_mf_assemble(l3b_mf* sys, int kernel_id, l3b_asm_opts opts, double time, const int* dof_inds, const l3b_fields* fields,
                    const int* field_inds, const int* boundary_ids, int n_boundary_ids)
{
    return guardedCtx(sys->ctx, [&] {
        if (sys->closed)
            fail(L3B_ERR_STATE, "`assembleProblem()` was called after `endAssembly()`");
        sys->uses.push_back(makeUse(sys->mesh, sys->dpn, sys->n_rhs, kernel_id, opts, time, dof_inds, fields, field_inds, boundary_ids, n_boundary_ids));
    });
}
int l3b_mf_end_assembly_begin(l3b_mf* sys)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"computeDiagAndRhs"}; // MatrixFreeSystem.hpp:887-941
        if (sys->closed)
            fail(L3B_ERR_STATE, "`endAssembly()` was called more than once");
        auto* ctx = sys->ctx;
        sys->diag.zero(ctx->stream);
        sys->rhs.zero(ctx->stream);
        // domain kernels: diag + F_e by the coefficient-form kernel (mf_init.cuh), then the Dirichlet lifting rhs -= A_unmasked g as
        // one operator apply; boundary kernels (and everything under L3B_MF_INIT_DENSE=1): the dense H_q formulation
        static const bool force_dense = [] {
            const char* e = std::getenv("L3B_MF_INIT_DENSE");
            return e != nullptr and e[0] == '1';
        }();
        for (const auto& use : sys->uses)
        {
            const auto& info = kernelRegistry()[use.kernel_id].info;
            ElemArgs    a    = baseArgs(sys->mesh, use, sys->dpn, sys->n_dofs);
            setDense(a, sys->mesh, use, info.is_boundary);
            a.dir_mask = sys->has_bc ? sys->dir_mask.ptr : nullptr;
            a.dir_vals = sys->has_bc ? sys->dir_vals.ptr : nullptr;
            a.diag     = sys->diag.ptr;
            a.rhs      = sys->rhs.ptr;
            const bool fast = use.inst->init_fast != nullptr and not force_dense;
            cudaCheck((fast ? use.inst->init_fast : use.inst->init)(kernelRegistry()[use.kernel_id].object.get(), a, ctx->stream), "init launch");
            if (fast and sys->lift_needed)
                applyUse(sys, use, sys->dir_g.ptr, sys->rhs.ptr, sys->n_rhs, -1., false, nullptr, 0, sys->mesh->n_elems, true);
        }
        cudaCheck(cudaGetLastError(), "init");
    });
}
int l3b_mf_end_assembly_finish(l3b_mf* sys)
{
    return guardedCtx(sys->ctx, [&] {
        if (sys->closed)
            fail(L3B_ERR_STATE, "`endAssembly()` was called more than once");
        auto* ctx = sys->ctx;
        if (sys->has_bc)
            dirichletInitKernel<<< gridFor(sys->n_dofs), 256, 0, ctx->stream >>>(sys->dir_mask.ptr, sys->dir_vals.ptr, sys->diag.ptr, sys->rhs.ptr,
                                                                             sys->n_dofs, sys->n_dofs, sys->n_rhs);
        cudaCheck(cudaGetLastError(), "init");
        ctx->checkStatus();
        sys->closed = true;
    });
}
int l3b_mf_end_assembly(l3b_mf* sys)
{
    int rc = l3b_mf_end_assembly_begin(sys);
    if (rc != L3B_OK)
        return rc;
    // more than one rank: the ghost parts of diag and rhs go to their owners (MatrixFreeSystem.hpp:925-938)
    rc = guardedCtx(sys->ctx, [&] {
        haloExportBegin(sys->halo, sys->diag.ptr, 1);
        haloExportEnd(sys->halo, sys->diag.ptr, 1);
        haloExportBegin(sys->halo, sys->rhs.ptr, sys->n_rhs);
        haloExportEnd(sys->halo, sys->rhs.ptr, sys->n_rhs);
    });
    return rc != L3B_OK ? rc : l3b_mf_end_assembly_finish(sys);
}
double* l3b_mf_device_diag(l3b_mf* sys)
{
    return sys->diag.ptr;
}
double* l3b_mf_device_rhs(l3b_mf* sys)
{
    return sys->rhs.ptr;
}
int l3b_gmres_device(l3b_context* ctx, int64_t n_local, int64_t n_owned, l3b_apply_callback apply, l3b_allreduce_callback allreduce,
                     void* user, const double* diag, const double* b, double* x, double tol, int restart_length, int max_restarts,
                     int max_iters, int x0_is_zero, double* achieved_tol, int* iters)
{
    return guardedCtx(ctx, [&] {
        if (apply == nullptr or n_owned > n_local or n_owned < 0 or restart_length < 1)
            fail(L3B_ERR_INVALID_ARG, "l3b_gmres_device: invalid arguments");
        gmres(
            ctx, n_local, n_owned,
            [&](const double* in, double* out) {
                const int rc = apply(user, in, out, nullptr);
                if (rc != 0 and rc != 1)
                    fail(L3B_ERR_INVALID_ARG, "l3b_gmres_device: the apply callback failed");
            },
            [&](double* sc, int n) {
                if (allreduce != nullptr and allreduce(user, sc, n) != 0)
                    fail(L3B_ERR_INVALID_ARG, "l3b_gmres_device: the all-reduce callback failed");
            },
            diag, b, x, tol, restart_length, max_restarts, max_iters, achieved_tol, iters, x0_is_zero != 0);
    });
}
int l3b_mf_solve_device(l3b_mf* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, int x0_is_zero,
                        double* achieved_tol, int* iters)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"solve"}; // MatrixFreeSystem::solve
        if (not sys->closed)
            fail(L3B_ERR_STATE, "`solve()` was called before `endAssembly()`");
        if (sys->n_rhs != 1)
            fail(L3B_ERR_INVALID_ARG, "the Krylov drivers handle one right-hand side");
        if (method != 0 and method != 1)
            fail(L3B_ERR_INVALID_ARG, "solver method: 0 (CG) or 1 (GMRES)");
        const auto reduce = [&](double* sc, int k) { commAllReduce(sys->comm(), sc, k); };
        if (method == 0)
            pcg(
                sys->ctx, sys->n_dofs, sys->ownedDofs(),
                [&](const double* in, double* out, double* energy) {
                    mfApplyDevice(sys, in, out, 1, 1., 0., energy);
                    return energy != nullptr; // p.Ap comes out of the apply
                },
                reduce, sys->diag.ptr, sys->rhs.ptr, x, tol, max_iters, achieved_tol, iters, x0_is_zero != 0);
        else
            gmres(
                sys->ctx, sys->n_dofs, sys->ownedDofs(), [&](const double* in, double* out) { mfApplyDevice(sys, in, out, 1, 1., 0.); }, reduce,
                sys->diag.ptr, sys->rhs.ptr, x, tol, restart_length, max_restarts, max_iters, achieved_tol, iters, x0_is_zero != 0);
        refreshGhosts(sys->halo, x);
    });
}
static int mfSolveHost(l3b_mf* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, double* achieved_tol,
                       int* iters)
{
    DevBuf< double > dx;
    const int        rc = guardedCtx(sys->ctx, [&] {
        dx.alloc(std::max< long long >(sys->n_dofs, 1));
        dx.upload(x, sys->n_dofs, sys->ctx->stream);
    });
    if (rc != L3B_OK)
        return rc;
    const int rc2 = l3b_mf_solve_device(sys, method, tol, max_iters, restart_length, max_restarts, dx.ptr, hostAllZero(x, sys->ownedDofs()),
                                        achieved_tol, iters);
    if (rc2 != L3B_OK)
        return rc2;
    return guardedCtx(sys->ctx, [&] {
        dx.download(x, sys->n_dofs, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "solve");
    });
}
int l3b_mf_solve_gmres(l3b_mf* sys, double tol, int restart_length, int max_restarts, int max_iters, double* x, double* achieved_tol, int* iters)
{
    return mfSolveHost(sys, 1, tol, max_iters, restart_length, max_restarts, x, achieved_tol, iters);
}
int l3b_pcg_device(l3b_context* ctx, int64_t n_local, int64_t n_owned, l3b_apply_callback apply, l3b_allreduce_callback allreduce,
                   void* user, const double* diag, const double* b, double* x, double tol, int max_iters, int x0_is_zero,
                   double* achieved_tol, int* iters)
{
    return guardedCtx(ctx, [&] {
        if (apply == nullptr or n_owned > n_local or n_owned < 0)
            fail(L3B_ERR_INVALID_ARG, "l3b_pcg_device: invalid arguments");
        pcg(
            ctx, n_local, n_owned,
            [&](const double* in, double* out, double* energy) {
                const int rc = apply(user, in, out, energy);
                if (rc != 0 and rc != 1)
                    fail(L3B_ERR_INVALID_ARG, "l3b_pcg_device: the apply callback failed");
                return rc == 1;
            },
            [&](double* sc, int n) {
                if (allreduce != nullptr and allreduce(user, sc, n) != 0)
                    fail(L3B_ERR_INVALID_ARG, "l3b_pcg_device: the all-reduce callback failed");
            },
            diag, b, x, tol, max_iters, achieved_tol, iters, x0_is_zero != 0);
    });
}
int l3b_mf_download(l3b_mf* sys, double* diag, double* rhs)
{
    return guardedCtx(sys->ctx, [&] {
        if (diag)
            sys->diag.download(diag, sys->n_dofs, sys->ctx->stream);
        if (rhs)
            sys->rhs.download(rhs, sys->rhs.n, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "download");
    });
}
int l3b_mf_apply_device(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta)
{
    return guardedCtx(sys->ctx, [&] { mfApplyDevice(sys, x, y, n_cols, alpha, beta); });
}
int l3b_mf_apply_energy_device(l3b_mf* sys, const double* x, double* y, double alpha, double beta, double* energy)
{
    return guardedCtx(sys->ctx, [&] { mfApplyDevice(sys, x, y, 1, alpha, beta, energy); });
}
int l3b_mf_apply_phase_device(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta, int phases, int64_t elem_begin,
                              int64_t elem_end, double* energy)
{
    return guardedCtx(sys->ctx, [&] {
        // boundary kernels: with L3B_APPLY_BOUNDARY, or with a NON-EMPTY element range that starts at element 0 — an empty range never
        // triggers them, so a caller without border elements (range [0, 0) followed by [0, n)) does not run them twice
        int ph = phases;
        if ((ph & L3B_APPLY_ELEMENTS) != 0 and elem_begin == 0 and elem_end > 0)
            ph |= L3B_APPLY_BOUNDARY;
        mfApplyPhases(sys, x, y, n_cols, alpha, beta, ph, elem_begin, elem_end, energy);
    });
}
int l3b_vec_gather(l3b_context* ctx, const double* src, int64_t ld, const int32_t* idx, int64_t n, int n_cols, double* dst)
{
    return guardedCtx(ctx, [&] {
        if (n > 0)
            vecGatherKernel<<< gridFor(n * n_cols), 256, 0, ctx->stream >>>(src, ld, idx, n, n_cols, dst);
        cudaCheck(cudaGetLastError(), "gather");
    });
}
int l3b_vec_scatter_add(l3b_context* ctx, double* dst, int64_t ld, const int32_t* idx, int64_t n, int n_cols, const double* src)
{
    return guardedCtx(ctx, [&] {
        if (n > 0)
            vecScatterAddKernel<<< gridFor(n * n_cols), 256, 0, ctx->stream >>>(dst, ld, idx, n, n_cols, src);
        cudaCheck(cudaGetLastError(), "scatter-add");
    });
}
int l3b_mf_apply(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta)
{
    return guardedCtx(sys->ctx, [&] {
        const size_t n = static_cast< size_t >(sys->n_dofs) * n_cols;
        if (sys->work_x.n < n)
        {
            sys->work_x.alloc(n);
            sys->work_y.alloc(n);
        }
        sys->last_host_apply_streamed = false;
        if (hostApplyStreams(sys, x, y, n_cols, beta))
        {
            mfApplyHostStreamed(sys, x, y, n_cols, alpha);
            return;
        }
        sys->work_x.upload(x, n, sys->ctx->stream);
        if (beta != 0.)
            sys->work_y.upload(y, n, sys->ctx->stream);
        mfApplyDevice(sys, sys->work_x.ptr, sys->work_y.ptr, n_cols, alpha, beta);
        sys->work_y.download(y, n, sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "apply");
        sys->ctx->checkStatus();
    });
}
int l3b_mf_set_host_apply(l3b_mf* sys, int mode, int n_chunks, int64_t block_nodes)
{
    return guardedCtx(sys->ctx, [&] {
        if (mode < 0 or mode > 2 or n_chunks < 0 or block_nodes < 1)
            fail(L3B_ERR_INVALID_ARG, "l3b_mf_set_host_apply: mode in {0, 1, 2}, n_chunks >= 0 (0 = by size), block_nodes >= 1");
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "sync");
        sys->pipe.reset(); // the schedule is rebuilt by the next streamed call
        sys->host_apply_mode        = mode;
        sys->host_apply_chunks      = n_chunks;
        sys->host_apply_block_nodes = block_nodes;
    });
}
int l3b_mf_host_apply_info(const l3b_mf* sys, int64_t info[4])
{
    info[0] = sys->last_host_apply_streamed ? 1 : 0;
    info[1] = sys->pipe ? sys->pipe->plan.n_items : 0;
    info[2] = sys->pipe ? static_cast< int64_t >(sys->pipe->plan.up_ranges.size() / 2) : 0;
    info[3] = sys->pipe ? static_cast< int64_t >(sys->pipe->plan.down_ranges.size() / 2) : 0;
    return L3B_OK;
}
int l3b_host_apply_plan(int64_t n_nodes, int64_t n_owned_nodes, int64_t n_elems, int nn, const uint32_t* nodes, int64_t n_border_elems,
                        const int32_t* halo_nodes, int64_t n_halo_nodes, int64_t chunk_elems, int64_t block_nodes, int* n_items,
                        int64_t** item_elems, int64_t** up_ptr, int64_t** up_ranges, int64_t** down_ptr, int64_t** down_ranges)
{
    return guardedCtx(nullptr, [&] {
        l3b::host::ApplyPlan p;
        try
        {
            p = l3b::host::makeApplyPlan(n_nodes, n_owned_nodes, n_elems, nn, nodes, n_border_elems, halo_nodes, n_halo_nodes, chunk_elems, block_nodes);
        }
        catch (const std::exception& e)
        {
            fail(L3B_ERR_INVALID_ARG, std::string{"host apply plan: "} + e.what());
        }
        const auto give = [](const std::vector< long long >& v, int64_t** out) {
            *out = static_cast< int64_t* >(std::malloc(std::max< size_t >(1, v.size()) * sizeof(int64_t)));
            if (not *out)
                fail(L3B_ERR_INVALID_ARG, "out of host memory");
            for (size_t i = 0; i < v.size(); ++i)
                (*out)[i] = v[i];
        };
        *n_items = p.n_items;
        give(p.item_elems, item_elems);
        give(p.up_ptr, up_ptr);
        give(p.up_ranges, up_ranges);
        give(p.down_ptr, down_ptr);
        give(p.down_ranges, down_ranges);
    });
}
int l3b_mf_solve_cg(l3b_mf* sys, double tol, int max_iters, double* x, double* achieved_tol, int* iters)
{
    return mfSolveHost(sys, 0, tol, max_iters, 1, 0, x, achieved_tol, iters);
}
int64_t l3b_mf_num_dofs(const l3b_mf* sys)
{
    return sys->n_dofs;
}
int l3b_mf_kernel_launches(const l3b_mf* sys)
{
    return sys->last_launches;
}
} // extern "C"

namespace
{
// computeIntegral / computeNormL2 (post/Integral.hpp:55-121, post/NormL2.hpp:10-60), this rank's elements
void integrate(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, l3b_asm_opts opts, double time, const l3b_fields* fields,
               const int* field_inds, const int* boundary_ids, int n_boundary_ids, bool norm, double* out)
{
    auto& reg = kernelRegistry();
    if (kernel_id < 0 or kernel_id >= static_cast< int >(reg.size()))
        fail(L3B_ERR_INVALID_ARG, "invalid kernel id");
    const auto& entry = reg[kernel_id];
    const auto& info  = entry.info;
    if (not info.is_residual)
        fail(L3B_ERR_INVALID_ARG, "kernel '" + info.name + "' is an equation kernel, integrals take residual kernels");
    if (info.dim != mesh->dim)
        fail(L3B_ERR_INVALID_ARG, "The dimensions of the kernel do not match the dimensions of the domain");
    if (opts.value_order < 1)
        fail(L3B_ERR_INVALID_ARG, "value_order must be >= 1");
    KernelUse use;
    use.kernel_id = kernel_id;
    use.opts      = opts;
    AssemblyOptions ao;
    ao.value_order      = opts.value_order * (norm ? 2 : 1); // NormL2.hpp:10-19
    ao.derivative_order = opts.derivative_order * (norm ? 2 : 1);
    use.nq              = ao.order(mesh->order) / 2 + 1; // QO = options.order(EO) (Integral.hpp:66-69), size QO / 2 + 1
    use.inst            = entry.find(mesh->order, 0);
    if (not use.inst or not use.inst->integrate)
        fail(L3B_ERR_NO_INSTANCE, "residual kernel '" + info.name + "' is not compiled for order " + std::to_string(mesh->order));
    use.time = time;
    if (info.n_fields > 0)
    {
        if (not fields)
            fail(L3B_ERR_INVALID_ARG, "kernel needs external fields but none were passed");
        if (fields->n_nodes != mesh->n_local_nodes)
            fail(L3B_ERR_INVALID_ARG, "field storage does not match the mesh's local node count");
        for (int f = 0; f < info.n_fields; ++f)
        {
            use.field_inds[f] = field_inds ? field_inds[f] : f;
            if (use.field_inds[f] < 0 or use.field_inds[f] >= fields->n_fields)
                fail(L3B_ERR_INVALID_ARG, "field index out of range");
        }
        use.fields = fields;
    }
    if (info.is_boundary)
        use.boundary_work = makeBoundaryWork(mesh, boundary_ids, n_boundary_ids);
    const int        nv = info.n_equations * info.n_rhs;
    DevBuf< double > sums(nv);
    sums.zero(ctx->stream);
    ElemArgs a = baseArgs(mesh, use, 1, 0);
    setDense(a, mesh, use, info.is_boundary);
    a.y      = sums.ptr;
    a.n_cols = norm ? 1 : 0;
    cudaCheck(use.inst->integrate(entry.object.get(), a, ctx->stream), "integrate");
    sums.download(out, nv, ctx->stream);
    cudaCheck(cudaStreamSynchronize(ctx->stream), "integrate");
    if (norm)
        for (int i = 0; i < nv; ++i)
            out[i] = std::sqrt(out[i]);
}
} // namespace

extern "C" {
int l3b_compute_integral(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, l3b_asm_opts opts, double time, const l3b_fields* fields,
                         const int* field_inds, const int* boundary_ids, int n_boundary_ids, double* out)
{
    return guardedCtx(ctx, [&] { integrate(ctx, mesh, kernel_id, opts, time, fields, field_inds, boundary_ids, n_boundary_ids, false, out); });
}
int l3b_compute_norm_l2(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, l3b_asm_opts opts, double time, const l3b_fields* fields,
                        const int* field_inds, const int* boundary_ids, int n_boundary_ids, double* out)
{
    return guardedCtx(ctx, [&] { integrate(ctx, mesh, kernel_id, opts, time, fields, field_inds, boundary_ids, n_boundary_ids, true, out); });
}
}


// ---- partition import (partition_host.hpp)
struct l3b_partition
{
    host::Partition part;
};
extern "C" {
int l3b_partition_create(int dim, int order, int64_t n_nodes, int64_t n_elems, const uint32_t* nodes, int n_parts, const int32_t* epart,
                         const int32_t* npart, l3b_partition** out)
{
    return guardedCtx(nullptr, [&] {
        if ((dim != 2 and dim != 3) or order < 1 or n_parts < 1)
            fail(L3B_ERR_INVALID_ARG, "l3b_partition_create: invalid dimension, order or part count");
        *out = new l3b_partition{host::makePartition(dim, order, n_nodes, n_elems, nodes, n_parts, epart, npart)};
    });
}
void l3b_partition_destroy(l3b_partition* p)
{
    delete p;
}
int l3b_partition_node_map(const l3b_partition* p, int64_t* new_id, int32_t* npart, int64_t* dist)
{
    return guardedCtx(nullptr, [&] {
        if (new_id)
            std::copy(p->part.new_id.begin(), p->part.new_id.end(), new_id);
        if (npart)
            std::copy(p->part.npart.begin(), p->part.npart.end(), npart);
        if (dist)
            std::copy(p->part.dist.begin(), p->part.dist.end(), dist);
    });
}
static void checkRank(const l3b_partition* p, int rank)
{
    if (rank < 0 or rank >= p->part.n_parts)
        fail(L3B_ERR_INVALID_ARG, "partition: rank out of range");
}
int l3b_partition_rank_info(const l3b_partition* p, int rank, int extended, int64_t info[8])
{
    return guardedCtx(nullptr, [&] {
        checkRank(p, rank);
        const auto& P = p->part;
        const auto  h = host::makeHaloLists(P, rank, extended != 0);
        info[0]       = static_cast< int64_t >(P.ranks[rank].elems.size());
        info[1]       = P.ranks[rank].n_border;
        info[2]       = P.nOwned(rank);
        info[3]       = P.nLocal(rank, extended != 0);
        info[4]       = static_cast< int64_t >(h.owned_nbrs.size());
        info[5]       = static_cast< int64_t >(h.shared_nbrs.size());
        info[6]       = h.owned_ptr.back();
        info[7]       = P.dist[rank];
    });
}
int l3b_partition_rank_mesh(const l3b_partition* p, int rank, int extended, int64_t* elem_ids, uint32_t* nodes, int64_t* local_to_global)
{
    return guardedCtx(nullptr, [&] {
        checkRank(p, rank);
        const auto& P  = p->part;
        const auto& R  = P.ranks[rank];
        const bool  ex = extended != 0;
        for (size_t i = 0; i < R.elems.size(); ++i)
        {
            if (elem_ids)
                elem_ids[i] = R.elems[i];
            if (nodes)
                for (int a = 0; a < P.nn; ++a)
                    nodes[i * P.nn + a] = static_cast< uint32_t >(P.localId(rank, ex, P.new_id[P.nodes[R.elems[i] * P.nn + a]]));
        }
        if (local_to_global)
        {
            const long long no = P.nOwned(rank);
            for (long long l = 0; l < no; ++l)
                local_to_global[l] = P.dist[rank] + l;
            const auto& g = P.ghostsOf(rank, ex);
            std::copy(g.begin(), g.end(), local_to_global + no);
        }
    });
}
int l3b_partition_rank_halo(const l3b_partition* p, int rank, int extended, int* owned_nbr_ranks, int64_t* owned_ptr, int32_t* owned_nodes,
                            int* shared_nbr_ranks, int64_t* shared_offsets)
{
    return guardedCtx(nullptr, [&] {
        checkRank(p, rank);
        const auto h = host::makeHaloLists(p->part, rank, extended != 0);
        std::copy(h.owned_nbrs.begin(), h.owned_nbrs.end(), owned_nbr_ranks);
        std::copy(h.owned_ptr.begin(), h.owned_ptr.end(), owned_ptr);
        std::copy(h.owned_nodes.begin(), h.owned_nodes.end(), owned_nodes);
        std::copy(h.shared_nbrs.begin(), h.shared_nbrs.end(), shared_nbr_ranks);
        std::copy(h.shared_off.begin(), h.shared_off.end(), shared_offsets);
    });
}
int l3b_partition_rank_graph(const l3b_partition* p, int rank, int64_t** ptr, uint32_t** nbr, int64_t** export_entry_ptr, uint32_t** export_pos)
{
    return guardedCtx(nullptr, [&] {
        checkRank(p, rank);
        const auto g = host::makeRankGraph(p->part, rank);
        const auto dup = [](const auto& v, auto** out) {
            using T = std::remove_pointer_t< std::remove_pointer_t< decltype(out) > >;
            *out    = static_cast< T* >(std::malloc(std::max< size_t >(1, v.size()) * sizeof(T)));
            if (not *out)
                fail(L3B_ERR_INVALID_ARG, "out of host memory");
            for (size_t i = 0; i < v.size(); ++i)
                (*out)[i] = static_cast< T >(v[i]);
        };
        dup(g.ptr, ptr);
        dup(g.nbr, nbr);
        if (export_entry_ptr and export_pos)
        {
            const auto h    = host::makeHaloLists(p->part, rank, true);
            const auto plan = host::makeRowExportPlan(p->part, rank, h, g);
            dup(plan.entry_ptr, export_entry_ptr);
            dup(plan.pos, export_pos);
        }
    });
}
// Tpetra FECrsMatrix::endAssembly / FEMultiVector::endAssembly as AssembledSystem::endAssembly calls them (AssembledSystem.hpp:384-389):
// the values of the shared (ghost) rows and the ghost block of the rhs go to the owners and are ADDED there; afterwards the owned rows
// are complete (the reference's row-complete owner matrix, what l3b_asm_download returns) and the ghost rows are zero.
int l3b_asm_export_shared_rows(l3b_asm* sys, const int64_t* recv_entry_ptr, const uint32_t* recv_pos)
{
    return guardedCtx(sys->ctx, [&] {
        const ProfileRegion region{"Matrix"}; // AssembledSystem.hpp:384-389: regions "RHS" and "Matrix" of endAssembly
        auto* const h = sys->halo;
        if (not sys->open)
            fail(L3B_ERR_STATE, "l3b_asm_export_shared_rows: call between `assembleProblem()` and `endAssembly()`");
        if (sys->rows_exported)
            fail(L3B_ERR_STATE, "l3b_asm_export_shared_rows was called twice");
        if (h == nullptr)
            fail(L3B_ERR_COMM, "l3b_asm_export_shared_rows needs the halo of the row layout (l3b_asm_set_halo)");
        const int       dpn = sys->dpn;
        const long long bs  = static_cast< long long >(dpn) * dpn;
        if (h->n_owned % dpn != 0)
            fail(L3B_ERR_INVALID_ARG, "row export: the halo is not node-blocked");
        const long long n_owned_nodes = h->n_owned / dpn;
        const auto&     np            = sys->node_ptr_host;
        // rhs first: the ordinary Export-sum of the ghost block
        haloExportBegin(h, sys->rhs.ptr, sys->n_rhs);
        haloExportEnd(h, sys->rhs.ptr, sys->n_rhs);
        if (h->active())
        {
            const auto&     api = ncclApi();
            auto* const     c   = h->comm;
            const auto      S   = sys->ctx->stream;
            const long long n_recv_nodes = h->owned_ptr.back() / dpn;
            // receive buffer + device copies of the plan
            const long long n_entries = n_recv_nodes > 0 ? recv_entry_ptr[n_recv_nodes] : 0;
            DevBuf< double >    buf(std::max< long long >(1, n_entries * bs));
            DevBuf< long long > d_ptr(n_recv_nodes + 1);
            DevBuf< int32_t >   d_row(std::max< long long >(1, n_recv_nodes));
            DevBuf< uint32_t >  d_pos(std::max< long long >(1, n_entries));
            std::vector< int32_t > row_node(n_recv_nodes);
            {
                std::vector< int32_t > idx(h->owned_ptr.back());
                h->owned_idx.download(idx.data(), idx.size(), S);
                cudaCheck(cudaStreamSynchronize(S), "halo index read");
                for (long long j = 0; j < n_recv_nodes; ++j)
                {
                    row_node[j] = idx[j * dpn] / dpn;
                    for (int d = 0; d < dpn; ++d)
                        if (idx[j * dpn + d] != row_node[j] * dpn + d)
                            fail(L3B_ERR_INVALID_ARG, "row export: the halo is not node-blocked");
                }
            }
            static_assert(sizeof(long long) == sizeof(int64_t));
            if (n_recv_nodes > 0)
                d_ptr.upload(reinterpret_cast< const long long* >(recv_entry_ptr), n_recv_nodes + 1, S);
            else
                d_ptr.zero(S);
            d_row.upload(row_node.data(), row_node.size(), S);
            d_pos.upload(recv_pos, n_entries, S);
            cudaCheck(cudaEventRecord(h->ev_ready, S), "event record");
            cudaCheck(cudaStreamWaitEvent(c->stream, h->ev_ready, 0), "event wait");
            ncclCheck(api.GroupStart(), "ncclGroupStart");
            for (size_t k = 0; k < h->owned_nbrs.size(); ++k)
            {
                const long long j0 = h->owned_ptr[k] / dpn, j1 = h->owned_ptr[k + 1] / dpn;
                const long long cnt = (recv_entry_ptr[j1] - recv_entry_ptr[j0]) * bs;
                if (cnt > 0)
                    ncclCheck(api.Recv(buf.ptr + recv_entry_ptr[j0] * bs, cnt, ncclDouble, h->owned_nbrs[k], c->comm, c->stream), "ncclRecv");
            }
            for (size_t j = 0; j < h->shared_nbrs.size(); ++j)
            {
                const long long n0 = n_owned_nodes + h->shared_off[j] / dpn, n1 = n_owned_nodes + h->shared_off[j + 1] / dpn;
                const long long cnt = (np[n1] - np[n0]) * bs;
                if (cnt > 0)
                    ncclCheck(api.Send(sys->values.ptr + np[n0] * bs, cnt, ncclDouble, h->shared_nbrs[j], c->comm, c->stream), "ncclSend");
            }
            ncclCheck(api.GroupEnd(), "ncclGroupEnd");
            cudaCheck(cudaEventRecord(h->ev_done, c->stream), "event record");
            cudaCheck(cudaStreamWaitEvent(S, h->ev_done, 0), "event wait");
            if (n_entries > 0)
                rowExportAddKernel<<< gridFor(n_entries), 256, 0, S >>>(buf.ptr, d_ptr.ptr, d_row.ptr, d_pos.ptr, n_recv_nodes, dpn, sys->node_ptr.ptr,
                                                                       sys->values.ptr);
            cudaCheck(cudaGetLastError(), "row export");
            // the ghost rows are spent
            const long long tail = (np[sys->n_nodes] - np[n_owned_nodes]) * bs;
            if (tail > 0)
                cudaCheck(cudaMemsetAsync(sys->values.ptr + np[n_owned_nodes] * bs, 0, tail * sizeof(double), S), "memset");
            const long long n_ghost_dofs = sys->n_dofs - h->n_owned;
            for (int col = 0; col < sys->n_rhs and n_ghost_dofs > 0; ++col)
                cudaCheck(cudaMemsetAsync(sys->rhs.ptr + h->n_owned + col * sys->n_dofs, 0, n_ghost_dofs * sizeof(double), S), "memset");
            cudaCheck(cudaStreamSynchronize(S), "row export"); // the staging buffers die here
        }
        sys->rows_exported = true;
    });
}
}

// ---- dof maps with inactive dofs / several domains (dofmap_host.hpp)
struct l3b_dofmap
{
    host::DofMap map;
};
extern "C" {
int l3b_mesh_set_element_domains(l3b_mesh* mesh, const int32_t* domain_ids)
{
    return guardedCtx(mesh->ctx, [&] {
        if (domain_ids)
            mesh->elem_domains.assign(domain_ids, domain_ids + mesh->n_elems);
        else
            mesh->elem_domains.clear();
    });
}
int l3b_dofmap_create(int dim, int order, int64_t n_nodes, int64_t n_elems, const uint32_t* nodes, const int32_t* elem_domains,
                      const uint16_t* side_boundaries, int dofs_per_node, int n_defs, const int* def_ptr, const int* def_domain_ids,
                      const uint32_t* def_dof_masks, int64_t base_dof, l3b_dofmap** out)
{
    return guardedCtx(nullptr, [&] {
        if ((dim != 2 and dim != 3) or order < 1 or n_defs < 1)
            fail(L3B_ERR_INVALID_ARG, "l3b_dofmap_create: invalid dimension, order or definition count");
        host::ProblemDefinition def;
        def.ptr.assign(def_ptr, def_ptr + n_defs + 1);
        def.ids.assign(def_domain_ids, def_domain_ids + def_ptr[n_defs]);
        def.mask.assign(def_dof_masks, def_dof_masks + n_defs);
        *out = new l3b_dofmap{host::makeDofMap(dim, order, n_nodes, n_elems, nodes, elem_domains, side_boundaries, dofs_per_node, def, base_dof)};
    });
}
void l3b_dofmap_destroy(l3b_dofmap* m)
{
    delete m;
}
int l3b_dofmap_info(const l3b_dofmap* m, int64_t info[4])
{
    info[0] = m->map.n_dofs;
    info[1] = static_cast< int64_t >(m->map.col_ind.size());
    info[2] = m->map.n_nodes;
    info[3] = m->map.dpn;
    return L3B_OK;
}
int l3b_dofmap_get(const l3b_dofmap* m, uint8_t* active, int64_t* dof, int64_t* row_ptr, int32_t* col_ind)
{
    return guardedCtx(nullptr, [&] {
        if (active)
            std::copy(m->map.active.begin(), m->map.active.end(), active);
        if (dof)
            std::copy(m->map.dof.begin(), m->map.dof.end(), dof);
        if (row_ptr)
            std::copy(m->map.row_ptr.begin(), m->map.row_ptr.end(), row_ptr);
        if (col_ind)
            std::copy(m->map.col_ind.begin(), m->map.col_ind.end(), col_ind);
    });
}
// the system closes every inactive (node, dof) pair as an identity row with zero rhs at endAssembly
int l3b_asm_set_dofmap(l3b_asm* sys, const l3b_dofmap* m)
{
    return guardedCtx(sys->ctx, [&] {
        if (m != nullptr and (m->map.n_nodes != sys->n_nodes or m->map.dpn != sys->dpn))
            fail(L3B_ERR_INVALID_ARG, "l3b_asm_set_dofmap: the dof map does not belong to this system's nodes");
        sys->dofmap = m ? &m->map : nullptr;
    });
}
// values and rhs in the reference's COMPACT numbering: CRS over the dof map's graph (row_ptr / col_ind of l3b_dofmap_get), rhs
// n_dofs x n_rhs column-major. Entries of the padded storage outside the compact graph must be zero (they are: no kernel writes there;
// the identity rows of the inactive pairs are not part of the compact system) — checked, L3B_ERR_GRAPH otherwise.
int l3b_asm_download_compact(l3b_asm* sys, const l3b_dofmap* m, double* values, double* rhs)
{
    return guardedCtx(sys->ctx, [&] {
        if (m == nullptr or m->map.n_nodes != sys->n_nodes or m->map.dpn != sys->dpn)
            fail(L3B_ERR_INVALID_ARG, "l3b_asm_download_compact: the dof map does not belong to this system's nodes");
        const auto&           M   = m->map;
        const int             dpn = sys->dpn;
        std::vector< double > pv(sys->nnz), pr(sys->rhs.n);
        sys->values.download(pv.data(), pv.size(), sys->ctx->stream);
        sys->rhs.download(pr.data(), pr.size(), sys->ctx->stream);
        std::vector< uint32_t > nbr(sys->node_nbr.n);
        sys->node_nbr.download(nbr.data(), nbr.size(), sys->ctx->stream);
        cudaCheck(cudaStreamSynchronize(sys->ctx->stream), "download");
        const auto&              np = sys->node_ptr_host;
        std::vector< long long > node_of(M.n_dofs), d_of(M.n_dofs);
        for (long long n = 0; n < M.n_nodes; ++n)
            for (int d = 0; d < dpn; ++d)
                if (M.dof[n * dpn + d] >= 0)
                {
                    node_of[M.dof[n * dpn + d] - M.base_dof] = n;
                    d_of[M.dof[n * dpn + d] - M.base_dof]    = d;
                }
        double kept2 = 0., all2 = 0.;
        for (double v : pv)
            all2 += v * v;
        for (long long r = 0; r < M.n_dofs; ++r)
        {
            const long long n = node_of[r], d = d_of[r], deg = np[n + 1] - np[n];
            const long long base = static_cast< long long >(dpn) * (dpn * np[n] + d * deg);
            for (long long k = M.row_ptr[r]; k < M.row_ptr[r + 1]; ++k)
            {
                const long long c = M.col_ind[k], cn = node_of[c], cv = d_of[c];
                const auto      it = std::lower_bound(nbr.begin() + np[n], nbr.begin() + np[n + 1], static_cast< uint32_t >(cn));
                if (it == nbr.begin() + np[n + 1] or *it != cn)
                    fail(L3B_ERR_GRAPH, "compact download: a column of the compact graph is missing from the padded storage");
                const double v = pv[base + cv * deg + (it - (nbr.begin() + np[n]))];
                kept2 += v * v;
                if (values)
                    values[k] = v;
            }
            if (rhs)
                for (int col = 0; col < sys->n_rhs; ++col)
                    rhs[r + col * M.n_dofs] = pr[n * dpn + d + col * sys->n_dofs];
        }
        // identity rows of closed (inactive) pairs are the only padded entries allowed outside the compact graph
        double closed2 = 0.;
        if (not sys->open and sys->dofmap == &M)
            for (size_t i = 0; i < M.active.size(); ++i)
                closed2 += M.active[i] ? 0. : 1.;
        if (std::fabs(all2 - kept2 - closed2) > 1e-20 * std::max(all2, 1.))
            fail(L3B_ERR_GRAPH, "compact download: the padded storage holds non-zero entries outside the compact sparsity graph");
    });
}
}

// ---- nodal values of residual kernels, solution -> fields
extern "C" {
// computeValuesAtNodes (algsys/ComputeValuesAtNodes.hpp:316-593, 595-721): what AssembledSystem / MatrixFreeSystem::setDirichletBCValues
// and setValues call to turn a residual kernel into nodal values
int l3b_compute_values_at_nodes(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, double time, const l3b_fields* fields, const int* field_inds,
                                const int* ids, int n_ids, int dofs_per_node, const int* dof_inds, double* values, int64_t ld, l3b_halo* halo)
{
    return guardedCtx(ctx, [&] {
        auto& reg = kernelRegistry();
        if (kernel_id < 0 or kernel_id >= static_cast< int >(reg.size()))
            fail(L3B_ERR_INVALID_ARG, "invalid kernel id");
        const auto& entry = reg[kernel_id];
        const auto& info  = entry.info;
        if (not info.is_residual)
            fail(L3B_ERR_INVALID_ARG, "kernel '" + info.name + "' is an equation kernel, nodal values take residual kernels");
        if (info.dim != mesh->dim)
            fail(L3B_ERR_INVALID_ARG, "The dimensions of the kernel do not match the dimensions of the domain");
        KernelUse use;
        use.kernel_id = kernel_id;
        use.inst      = entry.find(mesh->order, 0);
        if (not use.inst or not use.inst->values_at_nodes)
            fail(L3B_ERR_NO_INSTANCE, "residual kernel '" + info.name + "' is not compiled for order " + std::to_string(mesh->order));
        use.time = time;
        for (int eq = 0; eq < info.n_equations; ++eq)
        {
            use.dof_inds[eq] = dof_inds ? dof_inds[eq] : eq;
            if (use.dof_inds[eq] < 0 or use.dof_inds[eq] >= dofs_per_node)
                fail(L3B_ERR_INVALID_ARG, "dof index out of range");
        }
        if (info.n_fields > 0)
        {
            if (not fields or fields->n_nodes != mesh->n_local_nodes)
                fail(L3B_ERR_INVALID_ARG, "kernel needs external fields over the mesh's local nodes");
            for (int f = 0; f < info.n_fields; ++f)
            {
                use.field_inds[f] = field_inds ? field_inds[f] : f;
                if (use.field_inds[f] < 0 or use.field_inds[f] >= fields->n_fields)
                    fail(L3B_ERR_INVALID_ARG, "field index out of range");
            }
            use.fields = fields;
        }
        if (info.is_boundary)
            use.boundary_work = makeBoundaryWork(mesh, ids, n_ids);
        else if (n_ids > 0 and not mesh->elem_domains.empty())
        {
            for (long long e = 0; e < mesh->n_elems; ++e)
                for (int k = 0; k < n_ids; ++k)
                    if (mesh->elem_domains[e] == ids[k])
                    {
                        use.domain_elems.push_back(static_cast< int32_t >(e));
                        break;
                    }
            auto wl = std::make_shared< WorkList >();
            wl->n   = static_cast< long long >(use.domain_elems.size());
            wl->elems.alloc(std::max< size_t >(1, use.domain_elems.size()));
            wl->elems.upload(use.domain_elems.data(), use.domain_elems.size(), ctx->stream);
            cudaCheck(cudaStreamSynchronize(ctx->stream), "work list upload");
            use.domain_work = std::move(wl);
        }
        const long long  n_dofs = mesh->n_local_nodes * dofs_per_node;
        DevBuf< double > counts(std::max< long long >(n_dofs, 1));
        counts.zero(ctx->stream);
        ElemArgs a  = baseArgs(mesh, use, dofs_per_node, ld);
        const auto& t = ctx->nodeTables(mesh->dim, mesh->order);
        a.tab_vals  = t.vals.ptr;
        a.tab_ders  = t.ders.ptr;
        a.tab_pts   = t.pts.ptr;
        a.n_qp      = t.n_qp;
        a.y         = values;
        a.diag      = counts.ptr;
        for (int pass = 0; pass < 2; ++pass)
        {
            a.n_cols = pass;
            cudaCheck(use.inst->values_at_nodes(entry.object.get(), a, ctx->stream), "values at nodes");
        }
        // over ranks: contributions and counts of ghost nodes go to the owners, the averages come back (exchangeElementContributions, :156-194)
        haloExportBegin(halo, values, info.n_rhs);
        haloExportEnd(halo, values, info.n_rhs);
        haloExportBegin(halo, counts.ptr, 1);
        haloExportEnd(halo, counts.ptr, 1);
        if (n_dofs > 0)
            averageKernel<<< gridFor(n_dofs), 256, 0, ctx->stream >>>(values, counts.ptr, n_dofs, ld, info.n_rhs);
        cudaCheck(cudaGetLastError(), "values at nodes");
        haloImportBegin(halo, values, info.n_rhs);
        haloImportEnd(halo);
        cudaCheck(cudaStreamSynchronize(ctx->stream), "values at nodes"); // `counts` dies here
    });
}
double* l3b_fields_device(l3b_fields* f)
{
    return f->data.ptr;
}
// AssembledSystem / MatrixFreeSystem::updateSolution on the device: field field_inds[i] <- dof dof_inds[i] of the (padded) solution
// vector x over the local dofs; nothing leaves the GPU between a solve and the next assembly that reads the fields
int l3b_update_solution(l3b_context* ctx, const double* x, int dofs_per_node, const int* dof_inds, int n, l3b_fields* fields, const int* field_inds)
{
    return guardedCtx(ctx, [&] {
        if (n <= 0 or n > max_unknowns)
            fail(L3B_ERR_INVALID_ARG, "l3b_update_solution: between 1 and 8 components per call");
        for (int i = 0; i < n; ++i)
            if (dof_inds[i] < 0 or dof_inds[i] >= dofs_per_node or field_inds[i] < 0 or field_inds[i] >= fields->n_fields)
                fail(L3B_ERR_INVALID_ARG, "l3b_update_solution: index out of range");
        DevBuf< int > idx(2 * n);
        std::vector< int > h(dof_inds, dof_inds + n);
        h.insert(h.end(), field_inds, field_inds + n);
        idx.upload(h.data(), h.size(), ctx->stream);
        if (fields->n_nodes > 0)
            updateSolutionKernel<<< gridFor(fields->n_nodes * n), 256, 0, ctx->stream >>>(x, fields->n_nodes, dofs_per_node, idx.ptr, idx.ptr + n, n,
                                                                                         fields->data.ptr, fields->n_nodes);
        cudaCheck(cudaGetLastError(), "update solution");
        cudaCheck(cudaStreamSynchronize(ctx->stream), "update solution"); // the index buffer dies here
    });
}
}

// ---- static condensation (algsys/StaticCondensationManager.hpp:135-535), see condense.cuh
struct l3b_cond
{
    l3b_context*       ctx = nullptr;
    l3b_asm *          elem_sys = nullptr, *cond_sys = nullptr;
    long long          n_elems = 0;
    int                NN = 0, nB = 0, nI = 0;
    DevBuf< int >      bnd_idx, int_idx;
    DevBuf< uint32_t > elem_prim, elem_nodes;
    DevBuf< uint16_t > pos;
    DevBuf< double >   work;
    size_t             smem_condense = 0;
    int                maxc = 8;
    bool               large = false; // more than 256 interior dofs per element: condenseLargeKernel
    bool               condensed = false;
};
extern "C" {
int l3b_cond_create(l3b_context* ctx, l3b_asm* elem_sys, l3b_asm* cond_sys, int64_t n_elems, int nodes_per_elem, int n_bnd, const int* bnd_idx,
                    int n_int, const int* int_idx, const uint32_t* elem_prim, const uint32_t* elem_nodes, l3b_cond** out)
{
    return guardedCtx(ctx, [&] {
        if (elem_sys->dpn != cond_sys->dpn or elem_sys->n_rhs != cond_sys->n_rhs or n_bnd + n_int != nodes_per_elem or
            elem_sys->n_nodes != n_elems * nodes_per_elem)
            fail(L3B_ERR_INVALID_ARG, "l3b_cond_create: the element-local and the condensed system do not fit together");
        auto c      = std::make_unique< l3b_cond >();
        c->ctx      = ctx;
        c->elem_sys = elem_sys;
        c->cond_sys = cond_sys;
        c->n_elems  = n_elems;
        c->NN       = nodes_per_elem;
        c->nB       = n_bnd;
        c->nI       = n_int;
        c->bnd_idx.alloc(std::max(n_bnd, 1));
        c->int_idx.alloc(std::max(n_int, 1));
        c->bnd_idx.upload(bnd_idx, n_bnd, ctx->stream);
        c->int_idx.upload(int_idx, n_int, ctx->stream);
        c->elem_prim.alloc(std::max< long long >(n_elems * n_bnd, 1));
        c->elem_nodes.alloc(std::max< long long >(n_elems * nodes_per_elem, 1));
        c->elem_prim.upload(elem_prim, n_elems * n_bnd, ctx->stream);
        c->elem_nodes.upload(elem_nodes, n_elems * nodes_per_elem, ctx->stream);
        const long long n_pos = n_elems * n_bnd * n_bnd;
        c->pos.alloc(std::max< long long >(n_pos, 1));
        if (n_pos > 0)
            slotMapKernel<<< blocksFor(n_pos), 256, 0, ctx->stream >>>(
                c->elem_prim.ptr, n_elems, n_bnd, cond_sys->node_ptr.ptr, cond_sys->node_nbr.ptr, c->pos.ptr, ctx->status.ptr);
        cudaCheck(cudaGetLastError(), "slot map");
        ctx->checkStatus();
        const int nId = n_int * elem_sys->dpn;
        const int nPd    = n_bnd * elem_sys->dpn;
        if (nId > 256 or condSmemBytes(nId, nPd, elem_sys->n_rhs, false) > 220 * 1024)
        {
            // more interior dofs than the register-resident inverse (256) or the 64 x 64 Schur tiles (shared memory) take: hex p >= 5 at U = 4 (benchmarks/Diffusion3DBenchmark.cpp ships p = 6: 500 interior dofs): the blocked kernel, K_ii^-1 in a
            // global work buffer
            c->large         = true;
            c->smem_condense = condLargeSmemBytes(nId, nPd, elem_sys->n_rhs);
            if (c->smem_condense > 220 * 1024)
                fail(L3B_ERR_INVALID_ARG, "static condensation: the element's interior block is too large for this kernel");
            c->work.alloc(n_elems * static_cast< long long >(nId) * condLdM(nId));
            if (c->smem_condense > 48 * 1024)
                cudaCheck(cudaFuncSetAttribute(condenseLargeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast< int >(c->smem_condense)),
                          "smem");
            cudaCheck(cudaStreamSynchronize(ctx->stream), "cond create");
            *out = c.release();
            return;
        }
        c->smem_condense = condSmemBytes(nId, nPd, elem_sys->n_rhs, true);
        if (c->smem_condense > 113 * 1024) // K_ii does not fit shared memory (two CTAs per SM): invert it in a global work buffer
        {
            c->work.alloc(n_elems * static_cast< long long >(nId) * condLdM(nId));
            c->smem_condense = condSmemBytes(nId, nPd, elem_sys->n_rhs, false);
            if (c->smem_condense > 220 * 1024)
                fail(L3B_ERR_INVALID_ARG, "static condensation: the element's interior block is too large for this kernel");
        }
        c->maxc = nId <= 32 ? 1 : nId <= 64 ? 2 : nId <= 128 ? 4 : 8;
        const auto raise = [&](auto kernel) {
            if (c->smem_condense > 48 * 1024)
                cudaCheck(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast< int >(c->smem_condense)), "smem");
        };
        switch (c->maxc)
        {
        case 1: raise(condenseKernel< 1 >); break;
        case 2: raise(condenseKernel< 2 >); break;
        case 4: raise(condenseKernel< 4 >); break;
        default: raise(condenseKernel< 8 >);
        }
        cudaCheck(cudaStreamSynchronize(ctx->stream), "cond create");
        *out = c.release();
    });
}
void l3b_cond_destroy(l3b_cond* c)
{
    delete c;
}
/* endAssembly of the condensation manager (:330-353): Schur complements and condensed right-hand sides of all elements are added to
 * the (zeroed by its beginAssembly) condensed system; K_ii^-1 replaces K_ii in the element-local storage */
int l3b_cond_condense(l3b_cond* c)
{
    return guardedCtx(c->ctx, [&] {
        const ProfileRegion region{"Static condensation"}; // AssembledSystem.hpp:379-383
        if (not c->elem_sys->open or not c->cond_sys->open)
            fail(L3B_ERR_STATE, "l3b_cond_condense: both systems must be open for assembly");
        if (c->n_elems > 0)
        {
            CondArgs a{};
            a.ke        = c->elem_sys->values.ptr;
            a.fe        = c->elem_sys->rhs.ptr;
            a.ld_e      = c->elem_sys->n_dofs;
            a.NN        = c->NN;
            a.U         = c->elem_sys->dpn;
            a.n_rhs     = c->elem_sys->n_rhs;
            a.nB        = c->nB;
            a.nI        = c->nI;
            a.bnd_idx   = c->bnd_idx.ptr;
            a.int_idx   = c->int_idx.ptr;
            a.elem_prim = c->elem_prim.ptr;
            a.pos       = c->pos.ptr;
            a.node_ptr  = c->cond_sys->node_ptr.ptr;
            a.vals      = c->cond_sys->values.ptr;
            a.rhs       = c->cond_sys->rhs.ptr;
            a.ld_c      = c->cond_sys->n_dofs;
            a.work      = c->work.ptr;
            a.status    = c->ctx->status.ptr;
            const auto launch = [&](auto kernel) { kernel<<< static_cast< unsigned >(c->n_elems), cond_threads, c->smem_condense, c->ctx->stream >>>(a); };
            if (c->large)
                launch(condenseLargeKernel);
            else
            switch (c->maxc)
            {
            case 1: launch(condenseKernel< 1 >); break;
            case 2: launch(condenseKernel< 2 >); break;
            case 4: launch(condenseKernel< 4 >); break;
            default: launch(condenseKernel< 8 >);
            }
            cudaCheck(cudaGetLastError(), "condense");
        }
        c->ctx->checkStatus();
        c->elem_sys->open = false; // K_ip and f_i are now W and g: no further assembly into this storage
        c->condensed      = true;
    });
}
/* recoverSolution (:420-535): x_c = solution of the condensed system (host, n_primary_dofs x n_rhs column-major); out = nodal values
 * over the mesh's nodes (host, n_nodes x dofs_per_node row-major per rhs: out[node * U + u + r * n_nodes * U]) */
int l3b_cond_recover(l3b_cond* c, const double* x_c, int64_t n_nodes, double* out)
{
    return guardedCtx(c->ctx, [&] {
        if (not c->condensed)
            fail(L3B_ERR_STATE, "l3b_cond_recover before l3b_cond_condense");
        const int        U = c->elem_sys->dpn, R = c->elem_sys->n_rhs;
        DevBuf< double > dx(c->cond_sys->n_dofs * R), dout(n_nodes * U * R);
        dx.upload(x_c, c->cond_sys->n_dofs * R, c->ctx->stream);
        dout.zero(c->ctx->stream);
        if (c->n_elems > 0)
        {
            RecoverArgs a{};
            a.ke         = c->elem_sys->values.ptr;
            a.fe         = c->elem_sys->rhs.ptr;
            a.ld_e       = c->elem_sys->n_dofs;
            a.NN         = c->NN;
            a.U          = U;
            a.n_rhs      = R;
            a.nB         = c->nB;
            a.nI         = c->nI;
            a.bnd_idx    = c->bnd_idx.ptr;
            a.int_idx    = c->int_idx.ptr;
            a.elem_prim  = c->elem_prim.ptr;
            a.elem_nodes = c->elem_nodes.ptr;
            a.x_c        = dx.ptr;
            a.ld_c       = c->cond_sys->n_dofs;
            a.out        = dout.ptr;
            a.ld_out     = n_nodes * U;
            const size_t smem = static_cast< size_t >(c->nB) * U * sizeof(double);
            recoverKernel<<< static_cast< unsigned >(c->n_elems), cond_threads, smem, c->ctx->stream >>>(a);
            cudaCheck(cudaGetLastError(), "recover");
        }
        dout.download(out, n_nodes * U * R, c->ctx->stream);
        cudaCheck(cudaStreamSynchronize(c->ctx->stream), "recover");
    });
}
}
