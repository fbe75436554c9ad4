// CPU ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Plain C++20 restatement (no Eigen/TBB/Trilinos/MPI) of the L3STER hot path named in BASELINE.json:
// reference-element tables, geometric mapping, element-local least-squares assembly, CRS scatter, the non-sum-factorised
// and sum-factorised local operator, the matrix-free system apply and CG+Jacobi. Every function cites the reference
// file:line it follows (paths relative to the reference's include/l3ster/ unless they start with tests/ or benchmarks/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library. The
// product (l3ster_b200/) never links, imports or calls anything in oracle/.
//
// Parity pin: the reference cannot be compiled in this image (needs gcc>=14, Eigen, oneTBB, MPI, Trilinos, Metis —
// none present), and it ships no stored numeric dumps. The oracle is pinned against every known-answer test the
// reference's own test-suite holds for this path (tests/test_oracle_known_answers.py lists them one by one).
#ifndef L3STER_ORACLE_HPP
#define L3STER_ORACLE_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc
{
using val_t        = double;         // common/Typedefs.h:23
using n_id_t       = std::uint64_t;  // common/Typedefs.h:13
using global_dof_t = long long int;  // common/Typedefs.h:24
using local_dof_t  = int;            // common/Typedefs.h:25

enum ElementType : int
{
    Line = 1,
    Quad = 2,
    Hex  = 3
};
inline int nativeDim(ElementType et)
{
    return static_cast< int >(et);
}
inline int ipow(int b, int e)
{
    int r = 1;
    while (e-- > 0)
        r *= b;
    return r;
}
inline int numNodes(ElementType et, int order)
{
    return ipow(order + 1, nativeDim(et));
}
inline int numSides(ElementType et)
{
    return 2 * nativeDim(et);
}

// ---------------------------------------------------------------------------------------------------------------------
// math/  (Polynomial.hpp, Legendre.hpp, Lobatto.hpp, LobattoRuleAbsc.hpp, LagrangeInterpolation.hpp, ComputeGaussRule.hpp)
// Polynomial coefficient arrays are stored highest power first, as in math/Polynomial.hpp:57-66 (Horner order).
std::vector< val_t > legendreCoefs(int n);                                   // math/Legendre.hpp:8-49
std::vector< val_t > polyDerivative(const std::vector< val_t >& coefs);      // math/Polynomial.hpp:80-95
val_t                polyEval(const std::vector< val_t >& coefs, val_t x);   // math/Polynomial.hpp:57-66
const std::vector< val_t >& lobattoAbsc(int n_points);                       // math/LobattoRuleAbsc.hpp:10-35
std::vector< val_t > lagrangeInterp(const std::vector< val_t >& x, const std::vector< val_t >& y); // LagrangeInterpolation.hpp:12-40
struct GaussRule
{
    std::vector< val_t > points, weights;
};
const GaussRule& gaussLegendre(int n_points); // math/ComputeGaussRule.hpp:25-61 + quad/ReferenceQuadrature.hpp:24-51
inline int       refQuadSize(int quad_order)  // quad/ReferenceQuadrature.hpp:13-22
{
    return quad_order / 2 + 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// quad/GenerateQuadrature.hpp:11-77 — tensor rules; point index = i*nq^2 + j*nq + k ↦ (xi_i, xi_j, xi_k)
struct Quadrature
{
    int                  dim = 0, size = 0;
    std::vector< val_t > points; // [size][dim]
    std::vector< val_t > weights;
};
Quadrature makeQuadrature(ElementType et, int quad_order);

// ---------------------------------------------------------------------------------------------------------------------
// basisfun/ — 1-D Lagrange basis on Gauss–Lobatto nodes as monomial-coefficient polynomials; tensor products
struct LineBasis
{
    int                                 order = 0;
    std::vector< std::vector< val_t > > polys, ders; // per basis function, highest power first
};
const LineBasis& lineBasis(int order);                                                   // ReferenceBasisFunction.hpp:28-57
val_t refBasisValue(ElementType et, int order, int I, const val_t* point);               // ReferenceBasisFunction.hpp:74-153
val_t refBasisDer(ElementType et, int order, int I, int der_dim, const val_t* point);    // ReferenceBasisFunction.hpp:74-153

// basisfun/ReferenceBasisAtPoints.hpp:8-35, ReferenceBasisAtQuadrature.hpp:11-24
struct RefBasisAtQuad
{
    ElementType          et{};
    int                  order = 0, dim = 0, n_bases = 0;
    Quadrature           quad;
    std::vector< val_t > values;      // [q][a]
    std::vector< val_t > derivatives; // [q][d][a]   (row-major dim x n_bases per q)
};
RefBasisAtQuad makeRefBasisAtDomainQuad(ElementType et, int order, int quad_order);             // ReferenceElementBasisAtQuadrature.hpp:10-19
RefBasisAtQuad makeRefBasisAtBoundaryQuad(ElementType et, int order, int quad_order, int side); // ReferenceElementBasisAtQuadrature.hpp:56-96
RefBasisAtQuad makeRefBasisAtNodes(ElementType et, int order); // basis::getBasisAtNodes, at mesh::getNodeLocations

// ---------------------------------------------------------------------------------------------------------------------
// mapping/
// Vertices: 2^D points of 3 coordinates each, lexicographic x-fastest (mesh/ElementData.hpp:14-30)
void  jacobiMat(ElementType et, const val_t* verts, const val_t* point, val_t* J /*[d][s] dim x dim row-major*/); // JacobiMat.hpp:17-45
val_t det(int dim, const val_t* M);
void  inverse(int dim, const val_t* M, val_t* Minv);
void  mapToPhysicalSpace(ElementType et, const val_t* verts, const val_t* point, val_t* out3);              // MapReferenceToPhysical.hpp:14-25
void  refBoundaryToSide(ElementType et, int side, val_t* rot /*dim x dim row-major*/, val_t* trans);        // ReferenceBoundaryToSideMapping.hpp:14-48
val_t boundaryIntegralJacobian(ElementType et, int side, const val_t* J);                                   // BoundaryIntegralJacobian.hpp:9-28
void  boundaryNormal(ElementType et, int side, const val_t* J, val_t* normal);                              // BoundaryNormal.hpp:8-64

// ---------------------------------------------------------------------------------------------------------------------
// common/KernelInterface.hpp:13-57 — runtime-sized mirror of KernelParams / KernelInterface<params>
struct KernelParams
{
    int dimension = 0, n_equations = 0, n_unknowns = 1, n_fields = 0, n_rhs = 1;
};
struct Op // view of one E x U operator, Eigen column-major (KernelInterface.hpp:32)
{
    val_t* p;
    int    rows;
    val_t& operator()(int i, int j) const { return p[i + j * rows]; }
};
struct RhsView // E x n_rhs column-major (KernelInterface.hpp:33)
{
    val_t* p;
    int    rows;
    val_t& operator[](int i) const { return p[i]; }
    val_t& operator()(int i, int j) const { return p[i + j * rows]; }
};
struct KernelInput // DomainInput / BoundaryInput (KernelInterface.hpp:44-56)
{
    const val_t* field_vals;              // [n_fields]
    const val_t* field_ders[3];           // [dim][n_fields]
    val_t        space[3];
    val_t        time;
    val_t        normal[3];
};
struct KernelOutput // Result (KernelInterface.hpp:34-38), zero-initialised before the call (:61-69)
{
    Op      operators[4];
    RhsView rhs;
};
struct Kernel
{
    KernelParams                                                  params;
    bool                                                          is_boundary = false;
    std::function< void(const KernelInput&, const KernelOutput&) > fn;
};
const Kernel& getKernel(const std::string& name); // kernels.cpp: restatements of tests/Kernels.hpp, benchmarks/*.hpp, examples/*

// algsys/AssembleLocalSystem.hpp:24-49
struct AssemblyOptions
{
    int value_order = 1, derivative_order = 0;
    int eval_strategy = 0; // 0 Auto, 1 LocalElement, 2 SumFactorization, 3 SumFactorizationOddEvenDecomposition
    int order(int elem_order) const { return value_order * elem_order + derivative_order * (elem_order - 1); }
    bool useOddEven(int EO) const
    {
        if (eval_strategy == 0)
            return EO >= 2 && EO <= 6;
        return eval_strategy != 2;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// algsys/AssembleLocalSystem.hpp:234-280  — K_e row-major L x L (full, mirrored), F_e column-major L x n_rhs
// node_vals: row-major n_nodes x n_fields. side < 0 → domain kernel, otherwise boundary kernel on that side.
// Throws std::runtime_error("Encountered degenerate element ( |J| <= 0 )") as the reference does (:249).
void assembleLocalSystem(const Kernel&         kernel,
                         ElementType           et,
                         int                   order,
                         const val_t*          verts,
                         const val_t*          node_vals,
                         const RefBasisAtQuad& rbq,
                         val_t                 time,
                         int                   side,
                         val_t*                K_out,
                         val_t*                F_out);

// post/Integral.hpp:11-53 evalElementIntegral / evalElementBoundaryIntegral: out[E * n_rhs] = sum_q w_q jac_q kernel(input_q), the kernel
// being a residual kernel (it fills `rhs` only); squared: every component is squared first (post/NormL2.hpp:21-28)
void evalElementIntegral(const Kernel&         kernel,
                         ElementType           et,
                         int                   order,
                         const val_t*          verts,
                         const val_t*          node_vals,
                         const RefBasisAtQuad& rbq,
                         val_t                 time,
                         int                   side,
                         bool                  squared,
                         val_t*                out);

// algsys/EvaluateLocalOperator.hpp:211-263 — y = K_e x without forming K_e; x, y column-major L x n_cols
void evaluateLocalOperator(const Kernel&         kernel,
                           ElementType           et,
                           int                   order,
                           const val_t*          verts,
                           const val_t*          node_vals,
                           const RefBasisAtQuad& rbq,
                           val_t                 time,
                           int                   side,
                           int                   n_cols,
                           const val_t*          x,
                           val_t*                y);

// algsys/EvaluateLocalOperator.hpp:172-208, 276-328 — diagonal (L), rhs (L x n_rhs col-major) incl. Dirichlet lifting
void precomputeOperatorDiagonalAndRhs(const Kernel&         kernel,
                                      ElementType           et,
                                      int                   order,
                                      const val_t*          verts,
                                      const val_t*          node_vals,
                                      const RefBasisAtQuad& rbq,
                                      val_t                 time,
                                      int                   side,
                                      int                   n_dirichlet,
                                      const int*            dirichlet_inds,
                                      const val_t*          dirichlet_vals, // n_dirichlet x n_rhs col-major
                                      val_t*                diag_out,
                                      val_t*                rhs_out);

// ---------------------------------------------------------------------------------------------------------------------
// algsys/SumFactorization.hpp
struct SumFactTables // :11-65, 88-157
{
    int                  nb = 0, nq = 0;
    bool                 odd_even = false;
    std::vector< val_t > interp, der;             // row-major nb x nq
    std::vector< val_t > interp_t, der_t;         // row-major nq x nb
    std::vector< val_t > weights;                 // 1-D GL weights
};
SumFactTables makeSumFactTables(int basis_order, int quad_order, bool odd_even);
// one sweep: `in` viewed col-major (n_in x cols), out col-major (cols x n_out); out = in^T * M  (:67-86) — or its
// odd-even variant (:159-383); M row-major n_in x n_out
void sumFactSweep(const val_t* in, val_t* out, int n_in, int n_out, int cols, const val_t* M, bool is_der, bool accumulate, bool odd_even);

// :882-917 — X: gather buffer in the reference layout dest[rhs*(U*n_nodes) + dof*n_nodes + node] followed by n_fields
// node-major columns; Y: result, row-major n_nodes x (U*n_rhs) (src index node*(U*n_rhs) + rhs*U + dof)
void evalLocalOperatorSumFact(const Kernel&          kernel,
                              ElementType            et,
                              int                    order,
                              const val_t*           verts,
                              const AssemblyOptions& opts,
                              val_t                  time,
                              int                    n_rhs_actual,
                              const val_t*           X, // n_nodes x (U*n_rhs_actual + n_fields) col-major
                              val_t*                 Y);

// ---------------------------------------------------------------------------------------------------------------------
// mesh/ — structured primitives + order conversion; see mesh.cpp
struct Mesh
{
    ElementType              et{};     // type of the volume elements
    int                      order = 1;
    std::size_t              n_nodes = 0;
    std::size_t              n_elems = 0;
    std::vector< n_id_t >    elem_nodes; // [n_elems][nodes_per_elem] lexicographic local order
    std::vector< val_t >     elem_verts; // [n_elems][2^D][3]
    std::vector< n_id_t >    elem_ids;
    // boundary elements (one dimension lower), in the reference's generation order
    struct BoundaryElem
    {
        int      domain_id;
        n_id_t   id;
        std::vector< n_id_t > nodes; // lexicographic local order of the lower-dimensional element
        std::vector< val_t >  verts; // [2^(D-1)][3]
        // filled by matchBoundaries(): parent volume element index + side
        std::size_t parent = 0;
        int         side   = -1;
    };
    std::vector< BoundaryElem > boundary;
    int nodesPerElem() const { return numNodes(et, order); }
};
Mesh makeCubeMesh(const std::vector< val_t >& dx, const std::vector< val_t >& dy, const std::vector< val_t >& dz); // mesh/primitives/CubeMesh.hpp:16-138
Mesh makeSquareMesh(const std::vector< val_t >& dx, const std::vector< val_t >& dy);                             // mesh/primitives/SquareMesh.hpp:14-76
Mesh convertMeshToOrder(const Mesh& mesh, int order);                                                            // mesh/ConvertMeshToOrder.hpp:52-104
void matchBoundaries(Mesh& mesh);                                                                                 // mesh/MeshPartition.hpp:505-596
const std::vector< int >& boundaryNodeInds(ElementType et, int order);   // mesh/ElementTraits.hpp:29-57
const std::vector< int >& internalNodeInds(ElementType et, int order);
std::vector< int >        sideNodeInds(ElementType et, int order, int side); // mesh/ElementTraits.hpp:72-98, 118-137

// ---------------------------------------------------------------------------------------------------------------------
// systems — see system.cpp
struct CrsGraph
{
    std::vector< std::int64_t > row_ptr;
    std::vector< local_dof_t >  col_ind;
};
// algsys/SparsityGraph.hpp:25-81 + :254-278 for a single rank with all `U` unknowns active on every node:
// global dof = node*U + u (dofs/NodeToDofMap.hpp:249-264), rows sorted by local column id
CrsGraph makeSparsityGraph(const Mesh& mesh, int n_unknowns);

struct AssembledSystem
{
    const Mesh*          mesh = nullptr;
    int                  U = 0, n_rhs = 1;
    CrsGraph             graph;
    std::vector< val_t > values;
    std::vector< val_t > rhs; // col-major n_dofs x n_rhs
};
// algsys/AssembleGlobalSystem.hpp:13-96 + ScatterLocalSystem.hpp:25-54 (CondensationPolicy::None)
void assembleGlobalSystem(AssembledSystem& sys, const Kernel& kernel, const AssemblyOptions& opts, val_t time,
                          const val_t* fields /*field-major [n_fields][n_nodes] or null*/, int n_threads,
                          const std::vector< int >& boundary_ids /*for boundary kernels*/,
                          const std::vector< int >& dof_inds = {}, const std::vector< int >& field_inds = {});
// post/Integral.hpp:55-121 computeIntegral and post/NormL2.hpp:31-60 computeNormL2 on one rank: quadrature order opts.order(EO) — not
// doubled as in the assembly (Integral.hpp:66-69); the norm doubles both option orders, squares the components and takes square roots
// algsys/ComputeValuesAtNodes.hpp:316-369 (domain kernel) / :450-506 (boundary kernel): zero the visited dofs, add the kernel's value at
// every node of every visited element (side), count, average. values: [n_nodes * dpn][n_rhs] column-major, in/out.
void computeValuesAtNodes(const Mesh& mesh, const Kernel& kernel, val_t time, const val_t* fields, const std::vector< int >& field_inds,
                          const std::vector< int >& boundary_ids, int dpn, const std::vector< int >& dof_inds, val_t* values);
std::vector< val_t > computeIntegral(const Mesh& mesh, const Kernel& kernel, const AssemblyOptions& opts, val_t time, const val_t* fields,
                                     const std::vector< int >& boundary_ids, const std::vector< int >& field_inds, bool norm_l2);
// bcs/DirichletBC.hpp:82-150 — algebraic Dirichlet application (row/col zeroing, rhs lifting)
void applyDirichletAlgebraic(AssembledSystem& sys, const std::vector< local_dof_t >& dofs, const std::vector< val_t >& vals);

struct MatrixFreeSystem
{
    const Mesh*                mesh = nullptr;
    int                        U = 0, n_rhs = 1;
    std::vector< std::uint8_t > is_dirichlet; // per local dof
    std::vector< val_t >       dirichlet_vals; // col-major n_dofs x n_rhs
    std::vector< val_t >       diag, rhs;
    struct Entry
    {
        const Kernel*      kernel;
        AssemblyOptions    opts;
        val_t              time;
        std::vector< int > boundary_ids; // boundary kernels only
        const val_t*       fields;       // field-major [n_fields][n_nodes]
    };
    std::vector< Entry > kernels;
};
void mfComputeDiagAndRhs(MatrixFreeSystem& sys, int n_threads);                                        // algsys/MatrixFreeSystem.hpp:887-941
void mfApply(const MatrixFreeSystem& sys, const val_t* x, val_t* y, int n_cols, val_t alpha, val_t beta, int n_threads); // :1019-1140
// Preconditioned CG with the native Jacobi preconditioner (solve/NativePreconditioners.hpp:86-100); Belos "Block CG"
// semantics for block size 1: left preconditioner, absolute residual 2-norm test (solve/BelosSolvers.hpp:42-84)
struct SolveResult
{
    val_t tol;
    int   iters;
};
SolveResult cgJacobi(const std::function< void(const val_t*, val_t*) >& apply, const val_t* diag, const val_t* b, val_t* x,
                     std::size_t n, val_t tol, int max_iters);
} // namespace orc
#endif
