// Shared device helpers: small dense inverses, the vertex-based (sub-parametric) geometric mapping, argument blocks.
#ifndef L3B_DEVICE_COMMON_CUH
#define L3B_DEVICE_COMMON_CUH

#include "kernel_interface.cuh"

#include <cstdint>
#include <cuda_runtime.h>

namespace l3b
{
constexpr int cpow(int b, int e)
{
    int r = 1;
    for (int i = 0; i < e; ++i)
        r *= b;
    return r;
}
constexpr int cmax(int a, int b)
{
    return a > b ? a : b;
}

constexpr int max_unknowns = 8;
constexpr int max_fields   = 8;

// status word bits written by the device (mirrors the reference's exceptions, see l3ster_b200.h)
enum StatusBits : int
{
    status_degenerate_element = 1, // "Encountered degenerate element ( |J| <= 0 )" (AssembleLocalSystem.hpp:249)
    status_graph_entry_missing = 2,
    status_sparsity_violation  = 4, // an operator entry the compile-time probe found structurally zero was non-zero at run time
    status_slot_overflow       = 8, // a node with more than 65535 neighbours (slot positions are 16-bit)
    status_singular_interior   = 16 // static condensation: non-positive pivot in K_ii
};

// Everything an element kernel needs. Plain pointers and sizes only.
struct ElemArgs
{
    // mesh (mesh/LocalMeshView.hpp:13-57 condensed): per element 2^D vertices x 3 coords, n_nodes local node ids
    const double*   verts;
    const uint32_t* nodes;
    // work list: n_work items; elem = work_elems ? work_elems[i] : first_elem + i; side = work_sides ? work_sides[i] : -1
    long long       n_work;
    long long       first_elem;
    const int32_t*  work_elems;
    const uint8_t*  work_sides;
    // DOF layout: local dof of (node, kernel unknown u) = node * dofs_per_node + dof_inds[u]
    int dofs_per_node;
    int dof_inds[max_unknowns];
    int contiguous_dofs; // dofs_per_node == n_unknowns, dof_inds == identity, x 16-byte aligned: vectorised gathers allowed
    // operand / result (column-major, leading dimension ld)
    const double* x;
    double*       y;
    long long     ld;
    int           n_cols;
    double        alpha;
    // operator apply only, may be null: x^T A x of operand column 0 over the elements of this launch is added here (for CG's p.Ap:
    // sum_q w |B_q x_e|^2 falls out of the point stage, which saves the separate dot-product pass over both vectors)
    double* energy;
    // host-side launch hint (not read on the device): resident CTA slots a persistent launch leaves free, so that a communication
    // kernel queued on another stream (NCCL send / recv) can start while it runs
    int reserve_ctas;
    // Dirichlet mask per local dof (may be null) and prescribed values (ld-strided, may be null)
    const uint8_t* dir_mask;
    const double*  dir_vals;
    // per element: does any dof of any of its nodes carry a Dirichlet condition? (may be null = unknown, always look)
    const uint32_t* elem_dir;
    // hexahedra: per-element geometry record (mf_hex_planes.cuh: hex_geo_doubles), built at mesh upload
    const double* hex_geo;
    // init outputs
    double* diag;
    double* rhs;
    // external fields: field-major SoA, data[node + field_inds[f] * field_stride] (post/FieldAccess.hpp:40-47)
    const double* fields;
    long long     field_stride;
    int           field_inds[max_fields];
    double        time;
    // dense tables (non sum-factorised paths): values [q][a], derivatives [q][d][a], points [q][dim], weights [q];
    // for boundary work the tables of side s start at offsets s * (size of one side's table)
    const double* tab_vals;
    const double* tab_ders;
    const double* tab_pts;
    const double* tab_wts;
    int           n_qp;
    // domain quadrature of tensor elements: the 1-D tables behind the dense ones, [interp (nb x nq) | der (nb x nq)], b-major, and
    // nq; point q = qx + nq (qy + nq qz), node a = ix + nb (iy + nb iz). Null for side quadratures.
    const double* tab1d;
    int           nq1d;
    // CRS (assembly). Device layout of the values: row (n, d) starts at row_ptr = dpn * (dpn * node_ptr[n] + d * deg(n)) and holds its
    // entries column-dof-major: (neighbour k, column dof v) sits at v * deg(n) + k, so that consecutive neighbour nodes are consecutive
    // doubles (the scatter's atomics then share 32-byte sectors). l3b_asm_download returns the reference's node-major row layout.
    const long long* node_ptr;
    const long long* row_ptr;
    const uint16_t*  slot_pos; // [elem][a][b]: position (in node units) of node b's block in the rows of node a
    double*          crs_vals;
    int*             status;
};

template < int N >
struct Mat
{
    double v[N][N];
};

L3B_HD double det2(const double J[2][2])
{
    return J[0][0] * J[1][1] - J[0][1] * J[1][0];
}
L3B_HD double inv2(const double J[2][2], double Ji[2][2])
{
    const double d = det2(J), id = 1. / d;
    Ji[0][0] = J[1][1] * id;
    Ji[0][1] = -J[0][1] * id;
    Ji[1][0] = -J[1][0] * id;
    Ji[1][1] = J[0][0] * id;
    return d;
}
L3B_HD double inv3(const double M[3][3], double R[3][3])
{
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double d   = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
    const double id  = 1. / d;
    R[0][0] = c00 * id;
    R[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) * id;
    R[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) * id;
    R[1][0] = c01 * id;
    R[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) * id;
    R[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) * id;
    R[2][0] = c02 * id;
    R[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) * id;
    R[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) * id;
    return d;
}
template < int DIM >
L3B_HD double invert(const double (&M)[DIM][DIM], double (&R)[DIM][DIM])
{
    if constexpr (DIM == 2)
        return inv2(M, R);
    else
        return inv3(M, R);
}

// Geometry at reference point xi from the 2^D vertices (mapping/JacobiMat.hpp:17-45, MapReferenceToPhysical.hpp:14-25):
//   x_s = sum_v N1_v(xi) X_v[s];   Jt[d][s] = d x_s / d xi_d = sum_v dN1_v/dxi_d X_v[s]   (the reference's J(d, s))
template < int DIM >
L3B_HD void geometryAt(const double* verts, const double* xi, double (&xs)[3], double (&Jt)[DIM][DIM])
{
    double n1[DIM][2], d1[DIM][2];
    for (int d = 0; d < DIM; ++d)
    {
        n1[d][0] = 0.5 * (1. - xi[d]);
        n1[d][1] = 0.5 * (1. + xi[d]);
        d1[d][0] = -0.5;
        d1[d][1] = 0.5;
    }
    for (int s = 0; s < 3; ++s)
        xs[s] = 0.;
    for (int d = 0; d < DIM; ++d)
        for (int s = 0; s < DIM; ++s)
            Jt[d][s] = 0.;
    constexpr int nv = 1 << DIM;
    for (int v = 0; v < nv; ++v)
    {
        double nval = 1.;
        for (int d = 0; d < DIM; ++d)
            nval *= n1[d][(v >> d) & 1];
        for (int s = 0; s < 3; ++s)
            xs[s] += nval * verts[v * 3 + s];
        for (int d = 0; d < DIM; ++d)
        {
            double dval = 1.;
            for (int dd = 0; dd < DIM; ++dd)
                dval *= dd == d ? d1[dd][(v >> dd) & 1] : n1[dd][(v >> dd) & 1];
            for (int s = 0; s < DIM; ++s)
                Jt[d][s] += dval * verts[v * 3 + s];
        }
    }
}

// mapping/BoundaryIntegralJacobian.hpp:9-28 and mapping/BoundaryNormal.hpp:8-64, written out per side.
// Jt[d][s] is the reference's jacobi_mat(d, s). Returns the surface measure; fills the outward unit normal.
template < int DIM >
L3B_HD double boundaryMeasureAndNormal(int side, const double (&Jt)[DIM][DIM], double (&nrm)[DIM])
{
    if constexpr (DIM == 2)
    {
        // tangent = J^T * rot.col(0); rot.col(0) = (-1,0), (1,0), (0,-1), (0,1) for sides 0..3 → row 0 or row 1 of Jt
        const int    r       = side < 2 ? 0 : 1;
        const double measure = sqrt(Jt[r][0] * Jt[r][0] + Jt[r][1] * Jt[r][1]);
        switch (side)
        {
        case 0:
            nrm[0] = Jt[0][1];
            nrm[1] = -Jt[0][0];
            break;
        case 1:
            nrm[0] = -Jt[0][1];
            nrm[1] = Jt[0][0];
            break;
        case 2:
            nrm[0] = -Jt[1][1];
            nrm[1] = Jt[1][0];
            break;
        default:
            nrm[0] = Jt[1][1];
            nrm[1] = -Jt[1][0];
            break;
        }
        const double inv = 1. / sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1]);
        nrm[0] *= inv;
        nrm[1] *= inv;
        return measure;
    }
    else
    {
        // the two in-plane reference directions of the side: rows (0,1), (0,2) or (1,2) of Jt; |a x b| is the measure
        const int    r0 = side < 4 ? 0 : 1, r1 = side < 2 ? 1 : 2;
        const double cx = Jt[r0][1] * Jt[r1][2] - Jt[r0][2] * Jt[r1][1];
        const double cy = Jt[r0][2] * Jt[r1][0] - Jt[r0][0] * Jt[r1][2];
        const double cz = Jt[r0][0] * Jt[r1][1] - Jt[r0][1] * Jt[r1][0];
        const double measure = sqrt(cx * cx + cy * cy + cz * cz);
        // BoundaryNormal.hpp:38-58 signs: 0:-(r0 x r1) 1:+ 2:+(r0 x r2) 3:- 4:-(r1 x r2) 5:+
        const double sign = (side == 1 or side == 2 or side == 5) ? 1. : -1.;
        const double inv  = sign / measure;
        nrm[0]            = cx * inv;
        nrm[1]            = cy * inv;
        nrm[2]            = cz * inv;
        return measure;
    }
}

L3B_HD bool isDirichlet(const uint8_t* mask, long long dof)
{
    return mask != nullptr and mask[dof] != 0;
}
} // namespace l3b
#endif
