/*
 * l3ster_b200 — C ABI of the B200-native implementation of L3STER's element-local least-squares assembly and
 * matrix-free operator hot path (BASELINE.json: north_star).
 *
 * The reference (kubagalecki/L3STER) has no FFI: its boundary is a header-only C++23 template API. This header is the
 * boundary a replacement of that path binds to — plain pointers and sizes, opaque handles, int status codes — and each
 * entry point cites the reference interface it replaces (paths relative to the reference's include/l3ster/).
 * INTEGRATION.md shows the C++ shim a reference maintainer would put on top.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero l3b_status otherwise; l3b_last_error(ctx) gives the message.
 *     The reference signals the same conditions with util::throwingAssert → std::runtime_error (util/Assertion.hpp:85-93).
 *   - element types: dim 2 = Quad, dim 3 = Hex; local node order lexicographic x-fastest
 *     (basisfun/ReferenceBasisFunction.hpp:107-117); vertices 2^dim x 3 doubles per element, same order
 *     (mesh/ElementData.hpp:14-30); sides as mesh/ElementTraits.hpp:88-93.
 *   - local dof of (local node n, system dof d) = n * dofs_per_node + d (dofs/NodeToDofMap.hpp:249-264 for a single
 *     domain with all dofs active); vectors are column-major n_local_dofs x n_cols (Tpetra LayoutLeft).
 *   - all floating point data is fp64 (common/Typedefs.h:23), local node ids u32, local dofs i32, row pointers i64.
 *   - one host thread drives one context; work is issued on the context's CUDA stream.
 */
#ifndef L3STER_B200_H
#define L3STER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

typedef enum l3b_status
{
    L3B_OK                 = 0,
    L3B_ERR_INVALID_ARG    = 1,
    L3B_ERR_CUDA           = 2,
    L3B_ERR_NO_INSTANCE    = 3, /* kernel not compiled for this (order, nq): add it to its L3B_REGISTER_* list */
    L3B_ERR_DEGENERATE     = 4, /* "Encountered degenerate element ( |J| <= 0 )" (algsys/AssembleLocalSystem.hpp:249) */
    L3B_ERR_STATE          = 5, /* assembly state machine violated (algsys/AssembledSystem.hpp:455-461) */
    L3B_ERR_GRAPH          = 6, /* entry not present in the sparsity graph */
    L3B_ERR_NOT_CONVERGED  = 7,
    L3B_ERR_NO_DEVICE      = 8,
    L3B_ERR_SPARSITY       = 9, /* run-time guard of the compile-time operator sparsity probe tripped (kernel_interface.cuh) */
    L3B_ERR_SINGULAR       = 10, /* static condensation: singular interior block K_ii (the reference's inverse() would return garbage) */
    L3B_ERR_COMM           = 11  /* NCCL failure, or a distributed entry point used without a communicator */
} l3b_status;

typedef struct l3b_context   l3b_context;
typedef struct l3b_host_mesh l3b_host_mesh;
typedef struct l3b_mesh      l3b_mesh;
typedef struct l3b_fields    l3b_fields;
typedef struct l3b_asm       l3b_asm;
typedef struct l3b_mf        l3b_mf;
typedef struct l3b_comm      l3b_comm;
typedef struct l3b_halo      l3b_halo;

/* ---- context ----------------------------------------------------------------------------------------------------- */
/* replaces util::L3sterScopeGuard's device-side duties (util/ScopeGuards.hpp:186-201). Fails with L3B_ERR_NO_DEVICE
 * when no CUDA device is present: there is no CPU fallback. */
int         l3b_context_create(int device, l3b_context** out);
void        l3b_context_destroy(l3b_context* ctx);
const char* l3b_last_error(const l3b_context* ctx);
const char* l3b_global_error(void); /* errors raised before a context exists */
int         l3b_context_synchronize(l3b_context* ctx);
void*       l3b_context_stream(l3b_context* ctx); /* cudaStream_t */

/* ---- communicator (replaces comm/MpiComm.hpp:404-600 on the device data path: NCCL over NVLink, one rank per GPU) ------------ */
/* NCCL is resolved at run time (libnccl.so.2, the copy already in the process if there is one); without it these calls return
 * L3B_ERR_COMM and everything single-rank still works. Rank 0 draws the id and hands the bytes to the other ranks out of band (the
 * reference's MPI_Bcast; torch.distributed in the tests); every rank then creates its communicator — collective over the ranks.
 * Transfers run on a communication stream owned by the communicator, ordered against the context's stream with events. */
#define L3B_COMM_ID_BYTES 128
int  l3b_comm_unique_id(char id[L3B_COMM_ID_BYTES]);
int  l3b_comm_create(l3b_context* ctx, int rank, int world, const char id[L3B_COMM_ID_BYTES], l3b_comm** out);
int  l3b_comm_attach(l3b_context* ctx, void* nccl_comm /* an existing ncclComm_t; stays the caller's */, l3b_comm** out);
void l3b_comm_destroy(l3b_comm* comm);
int  l3b_comm_rank(const l3b_comm* comm);
int  l3b_comm_size(const l3b_comm* comm);
/* in-place sum over the ranks of n device doubles (MpiComm::allReduce, :452-466), ordered after the work on the context stream */
int  l3b_comm_allreduce_sum(l3b_comm* comm, double* device_scalars, int n);

/* ---- halo engine: comm::ImportExportContext + Import + Export (comm/ImportExport.hpp:29-72, 131-215, 296-470) ------------------
 * Vectors cover the local dofs [owned | ghost] (n_owned + n_ghost rows, column-major, ld = n_owned + n_ghost). As in the reference's
 * context (segmented ownership: every rank owns one contiguous global range, ghosts sorted by global id) the description has two halves:
 *   owned neighbours : ranks holding ghost copies of my owned dofs; owned_inds[owned_ptr[k] .. owned_ptr[k+1]) = the local owned dofs
 *                      shared with neighbour k, in the order the neighbour stores them (ascending global id) — packed for the Import,
 *                      targets of the Export's combine (sum);
 *   shared neighbours: ranks owning my ghost dofs; neighbour j's ghosts are the contiguous range
 *                      [shared_offsets[j], shared_offsets[j+1]) of the ghost block — received in place by the Import, sent in place
 *                      by the Export.
 * A rank may be its own neighbour (NCCL copies locally). All right-hand sides travel in one ncclGroup per exchange. */
int  l3b_halo_create(l3b_comm* comm, int64_t n_owned, int64_t n_ghost, int n_owned_nbrs, const int* owned_nbr_ranks,
                     const int64_t* owned_ptr, const int32_t* owned_inds, int n_shared_nbrs, const int* shared_nbr_ranks,
                     const int64_t* shared_offsets, l3b_halo** out);
void l3b_halo_destroy(l3b_halo* halo);
/* Import::postComms / wait (:296-384): begin packs on the context stream and posts the transfers on the communication stream; work
 * queued on the context stream before _end overlaps them; after _end the ghost block of x holds the owners' values. */
int l3b_halo_import_begin(l3b_halo* halo, double* x, int n_cols);
int l3b_halo_import_end(l3b_halo* halo);
/* Export::postRecvs / postSends / wait with AtomicSumInto (:403-470): begin (ordered after the work already on the context stream)
 * posts the ghost block of y to the owners; _end waits and adds the received contributions into the owned dofs. */
int l3b_halo_export_begin(l3b_halo* halo, double* y, int n_cols);
int l3b_halo_export_end(l3b_halo* halo, double* y, int n_cols);

/* ---- kernel registry (common/KernelInterface.hpp:13-20, 178-190) ---------------------------------------------------- */
typedef struct l3b_kernel_info
{
    char name[64];
    int  dimension, n_equations, n_unknowns, n_fields, n_rhs, is_boundary;
    int  n_instances; /* compiled (order, nq) pairs; residual kernels: one per order with nq = 0 (any quadrature size) */
    int  is_residual; /* an integrand for l3b_compute_integral / l3b_compute_norm_l2, not an equation kernel */
} l3b_kernel_info;
int l3b_kernel_count(void);
int l3b_kernel_find(const char* name); /* id or -1 */
int l3b_kernel_get_info(int kernel_id, l3b_kernel_info* out);
int l3b_kernel_get_instance(int kernel_id, int i, int* order, int* nq);

/* AssemblyOptions (algsys/AssembleLocalSystem.hpp:24-49). eval_strategy: 0 Auto, 1 LocalElement, 2 SumFactorization,
 * 3 SumFactorizationOddEvenDecomposition (2 and 3 select the same device kernel: odd-even is a CPU-side optimisation) */
typedef struct l3b_asm_opts
{
    int value_order, derivative_order, eval_strategy;
} l3b_asm_opts;

/* ---- reference-element tables (basisfun/, quad/, math/), exposed for parity tests ------------------------------------ */
int l3b_tables_gll(int n_points, double* nodes);                                 /* math/LobattoRuleAbsc.hpp:10-35 */
int l3b_tables_gauss(int n_points, double* points, double* weights);             /* quad/ReferenceQuadrature.hpp:24-51 */
int l3b_tables_1d(int order, int nq, double* interp, double* der, double* colloc); /* algsys/SumFactorization.hpp:25-65 */
/* dense tables at the domain (side < 0) or side quadrature: values [q][a], derivatives [q][d][a];
 * basisfun/ReferenceElementBasisAtQuadrature.hpp:10-96. Returns the number of points through *n_qp. */
int l3b_tables_dense(int dim, int order, int nq, int side, int* n_qp, double* points, double* weights, double* values, double* derivatives);

/* ---- host mesh front end (mesh/primitives/CubeMesh.hpp:16-138, SquareMesh.hpp:14-76, ConvertMeshToOrder.hpp:52-104) -- */
int  l3b_host_mesh_cube(int nx, const double* x, int ny, const double* y, int nz, const double* z, int order, l3b_host_mesh** out);
int  l3b_host_mesh_square(int nx, const double* x, int ny, const double* y, int order, l3b_host_mesh** out);
void l3b_host_mesh_destroy(l3b_host_mesh* m);
/* info: [dim, order, n_nodes, n_elems, nodes_per_elem, n_sides] */
int l3b_host_mesh_info(const l3b_host_mesh* m, int64_t info[6]);
/* borrowed pointers, valid until destroy */
const uint32_t* l3b_host_mesh_nodes(const l3b_host_mesh* m);
const double*   l3b_host_mesh_verts(const l3b_host_mesh* m);
const uint16_t* l3b_host_mesh_side_boundaries(const l3b_host_mesh* m); /* 0xFFFF = not on a boundary */
/* node-level sparsity graph (algsys/SparsityGraph.hpp:25-81, 254-278); expands to the dof-level CRS with
 * l3b_graph_expand. Arrays are malloc'ed by the library; release with l3b_free. */
int  l3b_node_graph(int64_t n_nodes, int64_t n_elems, int nodes_per_elem, const uint32_t* nodes, int64_t** ptr, uint32_t** nbr);
int  l3b_graph_expand(int64_t n_nodes, const int64_t* ptr, const uint32_t* nbr, int dofs_per_node, int64_t* row_ptr, int32_t* col_ind);
void l3b_free(void* p);

/* ---- partition import (mesh/PartitionMesh.hpp:322-440, dofs/NodeToDofMap.hpp:144-163, comm/ImportExport.hpp:29-72,
 *      algsys/SparsityGraph.hpp:83-278) --------------------------------------------------------------------------------------------
 * The reference partitions with METIS_PartMeshNodal (third party, not rebuilt); this takes the partition vectors from outside —
 * epart[element] and, optionally, npart[node] (NULL: the lowest part among the elements holding the node) — and performs the
 * deterministic rest exactly as the reference does: node assignment incl. the repair of "disjoint" nodes, the global renumbering that
 * gives every rank one contiguous id range, local numbering [owned | ghost by global id], neighbour lists for the halo engine, the
 * row-complete sparsity graph of the owned rows with the column map the neighbours' contributions induce, and the receive plan of
 * the shared-row export. `nodes` is the global order-p connectivity (every rank builds the same object: no set-up communication).
 * Views: extended = 0: [owned | ghost] (matrix-free system); extended = 1: [owned | ghost + extra columns] (assembled system). */
typedef struct l3b_partition l3b_partition;
int  l3b_partition_create(int dim, int order, int64_t n_nodes, int64_t n_elems, const uint32_t* nodes, int n_parts, const int32_t* epart,
                          const int32_t* npart, l3b_partition** out);
void l3b_partition_destroy(l3b_partition* p);
/* new_id[n_nodes]: input node id -> global id (renumberNodes); npart[n_nodes] after the repair; dist[n_parts + 1]: rank r owns
 * [dist[r], dist[r+1]). Any pointer may be NULL. */
int l3b_partition_node_map(const l3b_partition* p, int64_t* new_id, int32_t* npart, int64_t* dist);
/* info: [n_elems, n_border_elems, n_owned_nodes, n_local_nodes, n_owned_nbrs, n_shared_nbrs, n_nodes_in_owned_lists, first owned gid] */
int l3b_partition_rank_info(const l3b_partition* p, int rank, int extended, int64_t info[8]);
/* the rank's elements (global element ids, border elements first), their node lists in local ids, local id -> global id */
int l3b_partition_rank_mesh(const l3b_partition* p, int rank, int extended, int64_t* elem_ids, uint32_t* nodes, int64_t* local_to_global);
/* node-level description for l3b_halo_create (multiply by dofs_per_node for the dof-level lists) */
int l3b_partition_rank_halo(const l3b_partition* p, int rank, int extended, int* owned_nbr_ranks, int64_t* owned_ptr, int32_t* owned_nodes,
                            int* shared_nbr_ranks, int64_t* shared_offsets);
/* extended view: node graph for l3b_asm_create (owned rows complete, ghost rows as the rank's elements fill them, extra rows empty)
 * and, optionally, the receive plan of l3b_asm_export_shared_rows. Arrays are malloc'ed by the library; release with l3b_free. */
int l3b_partition_rank_graph(const l3b_partition* p, int rank, int64_t** ptr, uint32_t** nbr, int64_t** export_entry_ptr, uint32_t** export_pos);

/* ---- device mesh (mesh/LocalMeshView.hpp:13-57: vertices + local node ids + side → boundary id) ---------------------- */
int  l3b_mesh_upload(l3b_context* ctx, int dim, int order, int64_t n_elems, const double* verts, const uint32_t* nodes,
                     const uint16_t* side_boundaries /* may be NULL */, int64_t n_local_nodes, int64_t n_owned_nodes, l3b_mesh** out);
/* new vertex coordinates for the same connectivity (moving / re-read geometry): H2D copy + geometry records, no reallocation */
int  l3b_mesh_update_verts(l3b_mesh* mesh, const double* verts);
void l3b_mesh_destroy(l3b_mesh* mesh);

/* element -> domain id (mesh/Domain.hpp): with it, the `ids` of l3b_asm_assemble / l3b_mf_assemble select the elements a DOMAIN kernel
 * visits (AssembledSystem::assembleProblem(kernel, domain_ids, ...), MeshPartition::visit(..., domain_ids)); without it, or with zero ids,
 * a domain kernel visits every element. NULL clears. */
int l3b_mesh_set_element_domains(l3b_mesh* mesh, const int32_t* domain_ids);

/* ---- dof maps for problem definitions with inactive (node, dof) pairs and several domains ------------------------------------------
 * ProblemDefinition -> NodeToGlobalDofMap -> sparsity graph (dofs/NodeToDofMap.hpp:144-163, 188-264, 335-357,
 * algsys/SparsityGraph.hpp:25-81; tests/MultiDomainTest.cpp, tests/SparsityGraphTest.cpp:99-146). Definition i activates the dofs of bit
 * mask def_dof_masks[i] on the nodes of the elements whose domain id — and of the element sides whose boundary id — is one of
 * def_domain_ids[def_ptr[i] .. def_ptr[i+1]). The map numbers the ACTIVE pairs node-major from base_dof (the reference's global dofs:
 * pass the number of active pairs of the lower ranks as base, nodes in global-id order) and holds the reference's compact CRS graph
 * (per definition the clique of element nodes x its dofs, rows sorted). Device storage stays padded (node * dofs_per_node + d): see
 * l3b_asm_set_dofmap / l3b_asm_download_compact for the translation; for the matrix-free system pass the inactive pairs in the
 * Dirichlet mask with zero values. */
typedef struct l3b_dofmap l3b_dofmap;
int  l3b_dofmap_create(int dim, int order, int64_t n_nodes, int64_t n_elems, const uint32_t* nodes, const int32_t* elem_domains /* NULL: all 0 */,
                       const uint16_t* side_boundaries /* may be NULL */, int dofs_per_node, int n_defs, const int* def_ptr,
                       const int* def_domain_ids, const uint32_t* def_dof_masks, int64_t base_dof, l3b_dofmap** out);
void l3b_dofmap_destroy(l3b_dofmap* map);
int  l3b_dofmap_info(const l3b_dofmap* map, int64_t info[4]); /* [n_dofs, graph nnz, n_nodes, dofs_per_node] */
/* active[n_nodes * dpn]; dof[n_nodes * dpn] (compact id or -1); row_ptr[n_dofs + 1]; col_ind[nnz] (compact local ids). NULL = skip */
int  l3b_dofmap_get(const l3b_dofmap* map, uint8_t* active, int64_t* dof, int64_t* row_ptr, int32_t* col_ind);

/* ---- nodal fields (post/SolutionManager.hpp:54-101, post/FieldAccess.hpp:10-53): field-major [n_fields][n_local_nodes] */
int  l3b_fields_upload(l3b_context* ctx, int64_t n_local_nodes, int n_fields, const double* data, l3b_fields** out);
int  l3b_fields_update(l3b_fields* f, const double* data);
void l3b_fields_destroy(l3b_fields* f);

/* ---- assembled system (algsys/AssembledSystem.hpp:22-83, AssembleGlobalSystem.hpp:13-96, ScatterLocalSystem.hpp:25-54) */
/* node_ptr/node_nbr: node-level graph of l3b_node_graph (host). The device CRS has the dof-level layout of the reference's
 * Tpetra graph: row (n, d) holds, for each neighbour node m in ascending order, the dofs_per_node columns of m. */
int l3b_asm_create(l3b_context* ctx, l3b_mesh* mesh, int dofs_per_node, int n_rhs, const int64_t* node_ptr, const uint32_t* node_nbr,
                   l3b_asm** out);
void    l3b_asm_destroy(l3b_asm* sys);
int64_t l3b_asm_nnz(const l3b_asm* sys);
int     l3b_asm_begin_assembly(l3b_asm* sys); /* AssembledSystem::beginAssembly: zero matrix and rhs */
/* AssembledSystem::assembleProblem (:406-436). dof_inds: kernel unknown u → system dof (NULL = identity);
 * boundary_ids: the ids of the domains to visit — boundary ids for a boundary kernel, element domain ids for a domain kernel on a mesh
 * with l3b_mesh_set_element_domains (none given: every element). */
int l3b_asm_assemble(l3b_asm* sys, int kernel_id, l3b_asm_opts opts, double time, const int* dof_inds, const l3b_fields* fields,
                     const int* field_inds, const int* boundary_ids, int n_boundary_ids);
/* AssembledSystem::endAssembly (:373-397) with algebraic Dirichlet BCs (bcs/DirichletBC.hpp:82-150): rows → identity,
 * columns eliminated into the rhs. vals: n_dirichlet x n_rhs column-major. */
int l3b_asm_end_assembly(l3b_asm* sys, int64_t n_dirichlet, const int32_t* dirichlet_dofs, const double* dirichlet_vals);
int l3b_asm_download(l3b_asm* sys, double* values /* nnz, may be NULL */, double* rhs /* n_dofs x n_rhs, may be NULL */);
/* device values in the library's own row layout (row (n, d): column-dof-major, entry (neighbour k, dof v) at v * deg(n) + k);
 * l3b_asm_download converts to the reference's node-major Tpetra layout */
double* l3b_asm_device_values(l3b_asm* sys);
/* y = A x on the device CRS (Tpetra::CrsMatrix::apply), host buffers */
int l3b_asm_spmv(l3b_asm* sys, const double* x, double* y);
/* the same on device vectors over the local dofs, asynchronous on the context stream; and the diagonal of the local matrix.
 * More than one rank: every rank keeps the rows of its local nodes [owned | ghost] as its elements assembled them (the reference
 * instead export-adds the shared rows to their owners at endAssembly, AssembledSystem.hpp:384-389); the global operator is then
 * Import x, local product, Export-sum of the ghost rows of y (l3ster_b200/slab.py: SlabAssembledOperator), the same halo as the
 * matrix-free apply, and the global diagonal the Export-sum of the local ones. */
int l3b_asm_spmv_device(l3b_asm* sys, const double* x, double* y);
int l3b_asm_diag_device(l3b_asm* sys, double* diag);
/* l3b_asm_set_dofmap: endAssembly closes the inactive (node, dof) pairs as identity rows with zero rhs (they receive no contributions);
 * l3b_asm_download_compact: values over the dof map's compact graph + rhs over its dofs — the reference's matrix, bit-exact graph. */
int l3b_asm_set_dofmap(l3b_asm* sys, const l3b_dofmap* map);
int l3b_asm_download_compact(l3b_asm* sys, const l3b_dofmap* map, double* values, double* rhs);
/* AssembledSystem::endAssembly's export of the shared rows (m_matrix->endAssembly(), m_rhs->endAssembly(): AssembledSystem.hpp:384-389).
 * Needs the halo (l3b_asm_set_halo) and a system created on l3b_partition_rank_graph's graph. Every rank sends the values of its ghost
 * rows (contiguous per owner) and the ghost block of its rhs to the owners, which add them into their rows through the receive plan
 * (recv_entry_ptr / recv_pos of l3b_partition_rank_graph). Afterwards the owned rows are the reference's row-complete owner matrix
 * (l3b_asm_download), ghost rows are zero, and spmv / solvers skip the Export of y. Call between assembleProblem and endAssembly. */
int l3b_asm_export_shared_rows(l3b_asm* sys, const int64_t* recv_entry_ptr, const uint32_t* recv_pos);
/* the halo of the row layout [owned | ghost]: l3b_asm_spmv_device, l3b_asm_diag_device and the solvers then act as the global operator
 * (Import x, local product, Export-sum of the ghost rows; dots over the owned rows, all-reduced) */
int l3b_asm_set_halo(l3b_asm* sys, l3b_halo* halo);
double* l3b_asm_device_rhs(l3b_asm* sys);
/* l3b_asm_end_assembly for a rank that also holds ghost rows: Dirichlet rows become identity rows for the first n_owned_dofs dofs
 * and zero rows for the ghost dofs (their owner holds the identity row; the Export-sum must add nothing to it) */
int l3b_asm_end_assembly_ranked(l3b_asm* sys, int64_t n_dirichlet, const int32_t* dirichlet_dofs, const double* dirichlet_vals,
                                int64_t n_owned_dofs);
/* CG + native Jacobi on the assembled matrix (solve/BelosSolvers.hpp:116-123, NativePreconditioners.hpp:36-100).
 * x (host, n_dofs): in = initial guess — the reference hands Belos its persistent solution vector, so a repeated solve (time step,
 * Newton iteration) starts from the previous solution (AssembledSystem.hpp:113-135); out = solution. A zero guess skips the apply for
 * r0. With a halo (l3b_asm_set_halo) the solve runs over all ranks: dots over the owned rows, all-reduced through the communicator. */
int l3b_asm_solve_cg(l3b_asm* sys, double tol, int max_iters, double* x /* host, n_dofs */, double* achieved_tol, int* iters);
/* GMRES + native Jacobi on the assembled matrix (solve/BelosSolvers.hpp:125-131) */
int l3b_asm_solve_gmres(l3b_asm* sys, double tol, int restart_length, int max_restarts, int max_iters, double* x, double* achieved_tol,
                        int* iters);
/* the same solvers on a device vector over the local rows (the solution stays on the GPU between solves — the reference's
 * AssembledSystem::solve + updateSolution round trip without the host copy): method 0 = CG, 1 = GMRES; x in = initial guess
 * (x0_is_zero != 0: the caller vouches it is zero), out = solution with the ghost copies refreshed */
int l3b_asm_solve_device(l3b_asm* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, int x0_is_zero,
                         double* achieved_tol, int* iters);
/* timing of the last l3b_asm_assemble kernel launches (ms, CUDA events on the context stream) */
double l3b_asm_last_kernel_ms(const l3b_asm* sys);

/* ---- matrix-free system (algsys/MatrixFreeSystem.hpp:47-120) -------------------------------------------------------- */
/* dirichlet_mask: n_local_dofs bytes (bcs/LocalDirichletBC.hpp), may be NULL; dirichlet_vals: n_local_dofs x n_rhs */
int  l3b_mf_create(l3b_context* ctx, l3b_mesh* mesh, int dofs_per_node, int n_rhs, const uint8_t* dirichlet_mask,
                   const double* dirichlet_vals, l3b_mf** out);
void l3b_mf_destroy(l3b_mf* sys);
/* MatrixFreeSystem::assembleProblem: registers the kernel for init and apply (:585-787) */
int l3b_mf_assemble(l3b_mf* sys, int kernel_id, l3b_asm_opts opts, double time, const int* dof_inds, const l3b_fields* fields,
                    const int* field_inds, const int* boundary_ids, int n_boundary_ids);
/* More than one rank: the system takes the halo of its dof layout (mesh: n_owned_nodes / n_local_nodes; elements [0, n_border_elems)
 * are the ones touching ghost nodes, mesh/SplitMesh.hpp) and from then on l3b_mf_end_assembly, l3b_mf_apply[_device] and the solvers work
 * over all ranks like MatrixFreeSystem does (:887-941, 1019-1140): one call = pack + Import of x behind the pass that zeroes y, border
 * elements, Export of y behind the interior elements, unpack-add, Dirichlet rows; dot products all-reduced. Call before endAssembly. */
int l3b_mf_set_halo(l3b_mf* sys, l3b_halo* halo, int64_t n_border_elems);
/* MatrixFreeSystem::endAssembly → computeDiagAndRhs (:877-941) */
int l3b_mf_end_assembly(l3b_mf* sys);
/* the same in two steps for more than one rank: _begin accumulates the element contributions to diag and rhs over the local
 * dofs [owned | ghost]; the caller export-adds the ghost parts to their owners (MatrixFreeSystem.hpp:925-938) through the
 * device pointers below; _finish sets the owned Dirichlet dofs (diag = 1, rhs = g, :911-915) and closes the system. */
int     l3b_mf_end_assembly_begin(l3b_mf* sys);
int     l3b_mf_end_assembly_finish(l3b_mf* sys);
double* l3b_mf_device_diag(l3b_mf* sys);
double* l3b_mf_device_rhs(l3b_mf* sys);
int l3b_mf_download(l3b_mf* sys, double* diag, double* rhs);
/* Tpetra::Operator::apply → MatrixFreeSystem::applyImpl (:34-41, 1019-1140): y = alpha A x + beta y.
 * _device: x, y are device pointers (column-major, ld = n_local_dofs), asynchronous on the context stream.
 * host version: copies x (and y if beta != 0) in, y out. */
int l3b_mf_apply_device(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta);
int l3b_mf_apply(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta);
/* How the host version moves its vectors. The reference's Krylov vectors live in host memory, so this is the call a drop-in makes per
 * iteration, and at the benchmark's size its two PCIe copies cost 14 x the apply. mode 0: serial (x in, apply, y out). mode 1 (default):
 * streamed when beta == 0, only domain kernels are registered, x and y do not overlap and the vectors are at least 8 MB — the interior
 * elements are cut into n_chunks chunks; x travels in blocks of block_nodes nodes on a copy-in stream in the order the chunks first
 * touch them, chunk k runs when its blocks have landed, and every y block leaves on a copy-out stream once the last chunk that adds to
 * it has run and its Dirichlet rows are set; with a halo the border elements and the exchange form the last item. Any element order is
 * legal (no locality = serial behaviour). mode 2: streamed whenever legal, whatever the size. n_chunks 0 (default) = about one chunk per
 * 22 MB of vector, 2 to 32 (more chunks shorten the tail — the last chunk's y leaves after the last x block landed — but each costs
 * ~25 us of copy set-up and event hand-over); block_nodes defaults to 65536. Measured, 64^3 hex p=4 (543 MB per vector): 12.7 ms
 * against 21.1 ms serial and 11.4 ms for the two bare copies run concurrently. */
int l3b_mf_set_host_apply(l3b_mf* sys, int mode, int n_chunks, int64_t block_nodes);
/* info = {1 if the last l3b_mf_apply ran streamed, items, upload ranges, download ranges of the current schedule} */
int l3b_mf_host_apply_info(const l3b_mf* sys, int64_t info[4]);
/* the schedule itself (host only, csrc/apply_plan_host.hpp): items = chunks of chunk_elems interior elements [n_border_elems, n_elems),
 * then — with border elements, halo nodes or ghost nodes — one item for the border elements + exchange. item_elems[2 k .. 2 k + 1]:
 * element range of item k; up_ranges / down_ranges [2 r .. 2 r + 1]: node ranges to copy in before / out after the item, CSR offsets
 * in up_ptr / down_ptr [n_items + 1]. Arrays are malloc'ed by the library; release with l3b_free. */
int l3b_host_apply_plan(int64_t n_nodes, int64_t n_owned_nodes, int64_t n_elems, int nodes_per_elem, const uint32_t* nodes,
                        int64_t n_border_elems, const int32_t* halo_nodes, int64_t n_halo_nodes, int64_t chunk_elems, int64_t block_nodes,
                        int* n_items, int64_t** item_elems, int64_t** up_ptr, int64_t** up_ranges, int64_t** down_ptr, int64_t** down_ranges);
/* single-column apply that also adds this rank's share of x^T A x (its elements, its owned Dirichlet dofs) to the device scalar
 * `energy`: CG takes p.Ap from the quadrature-point stage instead of a dot-product pass (see l3b_mf_apply_phase_device) */
int l3b_mf_apply_energy_device(l3b_mf* sys, const double* x, double* y, double alpha, double beta, double* energy);
/* CG + native Jacobi, rhs = system rhs column 0 (benchmarks/Diffusion3D.hpp:115-118); x (host, n_dofs): in = initial guess (zero in the
 * benchmark), out = solution over the local dofs, ghost copies refreshed. With a halo (l3b_mf_set_halo) the solve runs over all ranks. */
int l3b_mf_solve_cg(l3b_mf* sys, double tol, int max_iters, double* x /* host */, double* achieved_tol, int* iters);
int l3b_mf_solve_gmres(l3b_mf* sys, double tol, int restart_length, int max_restarts, int max_iters, double* x /* host */,
                       double* achieved_tol, int* iters);
int l3b_mf_solve_device(l3b_mf* sys, int method, double tol, int max_iters, int restart_length, int max_restarts, double* x, int x0_is_zero,
                        double* achieved_tol, int* iters); /* as l3b_asm_solve_device */
/* The same apply split into the phases MatrixFreeSystem::applyImpl overlaps with its halo exchange (:1046-1122):
 *   L3B_APPLY_INIT     y <- beta y                                   (:1038)
 *   L3B_APPLY_ELEMENTS y[dofs(e)] += alpha K_e x[dofs(e)] for the domain elements e in [elem_begin, elem_end) — the caller passes
 *                      its interior range while comm::Import (owner -> ghost copies of x) is in flight, then the border range.
 *                      Boundary kernels (side lists are not split) run once per apply: in the call that carries L3B_APPLY_BOUNDARY,
 *                      or — without that bit anywhere — with the NON-EMPTY range that starts at element 0 (an empty range [0, 0),
 *                      as a rank without border elements passes, never triggers them)
 *   L3B_APPLY_FINISH   Dirichlet identity rows y[d] += alpha x[d]    (:1087-1103)
 * All phases are asynchronous on the context stream; x, y are device pointers over the LOCAL dofs [owned | ghost]. */
enum
{
    L3B_APPLY_INIT     = 1,
    L3B_APPLY_ELEMENTS = 2,
    L3B_APPLY_FINISH   = 4,
    L3B_APPLY_BOUNDARY = 8 /* run the boundary kernels (their whole side list) in this call */
};
/* energy: NULL, or a device scalar to which the ELEMENTS and FINISH phases add x^T A x of column 0 — this rank's elements and owned
 * Dirichlet dofs, so that the sum over the phases and over the ranks is the global x^T A x (alpha and beta do not enter). CG uses it for
 * p.Ap: the quadrature-point stage has w |B x_e|^2 at hand, which saves the dot-product pass over both vectors. */
int l3b_mf_apply_phase_device(l3b_mf* sys, const double* x, double* y, int n_cols, double alpha, double beta, int phases,
                              int64_t elem_begin, int64_t elem_end, double* energy);
/* halo packing for comm::Import / comm::Export (comm/ImportExport.hpp:29-472) with device index lists:
 * gather: dst[i + c n] = src[idx[i] + c ld];  scatter_add: dst[idx[i] + c ld] += src[i + c n]. Asynchronous on the context stream. */
int l3b_vec_gather(l3b_context* ctx, const double* src, int64_t ld, const int32_t* idx, int64_t n, int n_cols, double* dst);
int l3b_vec_scatter_add(l3b_context* ctx, double* dst, int64_t ld, const int32_t* idx, int64_t n, int n_cols, const double* src);
/* Preconditioned CG with callbacks, for operators and reductions the library does not own (the multi-rank apply with its halo
 * exchange, MPI/NCCL all-reduce of the dot products): Belos Block-CG semantics as l3b_mf_solve_cg, native Jacobi from `diag`.
 * Vectors are device pointers over n_local dofs of which the first n_owned are owned (dots and updates run over those).
 * apply(user, x, y, energy): y = A x, enqueued on the context stream; energy is NULL or a zeroed device scalar: an operator that can
 * add this rank's share of x^T A x to it while applying (l3b_mf_apply_phase_device) returns 1, otherwise it leaves it alone and returns 0
 * (the solver then runs its own dot product); any other return value is a failure. allreduce(user, s, n): in-place sum over the ranks of
 * the n device scalars at s, ordered after the work already on the context stream (may be NULL on one rank); returns 0 on success. */
typedef int (*l3b_apply_callback)(void* user, const double* x, double* y, double* energy);
typedef int (*l3b_allreduce_callback)(void* user, double* scalars, int n);
/* Restarted GMRES with the same callbacks: Belos "Pseudoblock GMRES" as solve/BelosSolvers.hpp:42-131 configures it ("Num Blocks" =
 * restart_length 250, "Maximum Restarts" 39 by default, solve/SolverInterface.hpp:26-37), left Jacobi preconditioner,
 * convergence on the preconditioned residual norm of the Givens recurrence (Belos' implicit test, absolute). */
/* x: in = the initial guess (Belos takes the system's persistent solution vector, AssembledSystem.hpp:113-135), out = the solution;
 * only its first n_owned entries are read and valid on return. x0_is_zero != 0: the caller vouches that the guess is zero (the first
 * solve of the reference's benchmarks) — x is zeroed and the operator apply for r0 = b - A x0 is skipped. */
int l3b_gmres_device(l3b_context* ctx, int64_t n_local, int64_t n_owned, l3b_apply_callback apply, l3b_allreduce_callback allreduce,
                     void* user, const double* diag, const double* b, double* x, double tol, int restart_length, int max_restarts,
                     int max_iters, int x0_is_zero, double* achieved_tol, int* iters);
int l3b_pcg_device(l3b_context* ctx, int64_t n_local, int64_t n_owned, l3b_apply_callback apply, l3b_allreduce_callback allreduce,
                   void* user, const double* diag, const double* b, double* x, double tol, int max_iters, int x0_is_zero,
                   double* achieved_tol, int* iters);
/* ---- integrals and norms of residual kernels (post/Integral.hpp:102-121 computeIntegral, post/NormL2.hpp:31-60 computeNormL2) ----
 * The kernel is a registered residual kernel `(const Input&, Rhs&) -> void` (common/KernelInterface.hpp:140-176); a domain kernel is
 * integrated over all elements of the mesh, a boundary kernel over the sides carrying one of `boundary_ids`. The integrand is
 * jacobian * kernel(input) on the Gauss-Legendre rule of order opts.order(element order) (Integral.hpp:66-69 — not doubled as in the
 * assembly); the norm squares each component, doubles value_order and derivative_order and returns the square roots
 * (NormL2.hpp:10-28, 44). out: n_equations * n_rhs doubles on the host, this rank's share (the caller all-reduces the integral, resp.
 * the squares of the norm, over the ranks: Integral.hpp:116-119). */
int l3b_compute_integral(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, l3b_asm_opts opts, double time, const l3b_fields* fields,
                         const int* field_inds, const int* boundary_ids, int n_boundary_ids, double* out);
int l3b_compute_norm_l2(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, l3b_asm_opts opts, double time, const l3b_fields* fields,
                        const int* field_inds, const int* boundary_ids, int n_boundary_ids, double* out);
/* ---- nodal values of residual kernels (algsys/ComputeValuesAtNodes.hpp:316-721) and solution -> fields ------------------------------
 * computeValuesAtNodes: the residual kernel is evaluated at the nodes of the elements of the domains `ids` (domain kernel; no ids = all
 * elements) or of the element sides on the boundaries `ids` (boundary kernel, with the outward normal); equation eq of the result is
 * written to dof (node, dof_inds[eq]) of `values` (device, padded dofs over the local nodes, leading dimension ld, n_rhs columns);
 * nodes shared by several elements get the average of their contributions; other dofs of `values` are left alone. halo (may be NULL):
 * the contributions of ghost nodes are combined at the owners and the result copied back. This is what setDirichletBCValues(kernel,
 * boundaries, dof_inds) and setValues(...) of the reference's systems do. */
int l3b_compute_values_at_nodes(l3b_context* ctx, l3b_mesh* mesh, int kernel_id, double time, const l3b_fields* fields, const int* field_inds,
                                const int* ids, int n_ids, int dofs_per_node, const int* dof_inds, double* values, int64_t ld, l3b_halo* halo);
/* updateSolution (AssembledSystem.hpp:140-160, MatrixFreeSystem.hpp:1231-1273) without leaving the device: field field_inds[i] <- dof
 * dof_inds[i] of the solution vector x (device, padded local dofs, e.g. the x of l3b_asm_solve_device); l3b_fields_device exposes the
 * field storage [n_fields][n_local_nodes] */
int     l3b_update_solution(l3b_context* ctx, const double* x, int dofs_per_node, const int* dof_inds, int n, l3b_fields* fields, const int* field_inds);
double* l3b_fields_device(l3b_fields* f);

/* ---- static condensation, CondensationPolicy::ElementBoundary (algsys/StaticCondensationManager.hpp:135-535) ---------------------
 * The element matrices are assembled, with the ordinary l3b_asm_* calls, into an element-local system: an l3b_asm over the same
 * elements with node ids e * nodes_per_elem + a (every element its own nodes: the CRS is then the dense K_e, block by block). The
 * condensed system lives on the primary (element-boundary) nodes: l3b_crs_create makes its storage from the node graph of the
 * elements' boundary-node lists (no mesh: it only receives l3b_cond_condense's contributions, then takes l3b_asm_end_assembly — the
 * algebraic Dirichlet conditions in primary dof numbering — and the solvers like any l3b_asm).
 *   l3b_cond_create: bnd_idx / int_idx = local node indices of the boundary / interior nodes of an element; elem_prim[e][ib] = primary
 *     id of boundary node ib; elem_nodes[e][a] = node id in the original mesh (for the recovered solution).
 *   l3b_cond_condense (StaticCondensationManager::endAssemblyImpl :330-353, with condenseSystemImpl's K_pp scatter :355-418 folded in):
 *     adds K_pp - K_pi K_ii^-1 K_ip and f_p - K_pi K_ii^-1 f_i of every element to the condensed system; closes the element-local one.
 *   l3b_cond_recover (recoverSolutionImpl :420-535): x_i = K_ii^-1 (f_i - K_ip x_p); out[node * dofs_per_node + d + r * n_nodes * dpn]. */
typedef struct l3b_cond l3b_cond;
int  l3b_crs_create(l3b_context* ctx, int64_t n_nodes, int dofs_per_node, int n_rhs, const int64_t* node_ptr, const uint32_t* node_nbr,
                    l3b_asm** out);
int  l3b_cond_create(l3b_context* ctx, l3b_asm* elem_sys, l3b_asm* cond_sys, int64_t n_elems, int nodes_per_elem, int n_bnd,
                     const int* bnd_idx, int n_int, const int* int_idx, const uint32_t* elem_prim, const uint32_t* elem_nodes,
                     l3b_cond** out);
void l3b_cond_destroy(l3b_cond* cond);
int  l3b_cond_condense(l3b_cond* cond);
int  l3b_cond_recover(l3b_cond* cond, const double* x_condensed, int64_t n_nodes, double* out);
int64_t l3b_mf_num_dofs(const l3b_mf* sys);
int     l3b_mf_kernel_launches(const l3b_mf* sys); /* device kernels launched by the last apply */

/* ---- roofline denominators measured in place (MEASURED_PEAKS.json has no fp64 figure) ------------------------------ */
/* mode 0: fp64 FMA TFLOP/s (CUDA cores); 1: fp64 DMMA TFLOP/s (mma.sync.m8n8k4.f64); 2: HBM copy GB/s (read + write) */
int l3b_microbench(l3b_context* ctx, int mode, double* result);

#ifdef __cplusplus
}
#endif
#endif
