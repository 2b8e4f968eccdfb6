/*
 * florence_b200.h -- C ABI of libflorence_b200.so: the B200 (sm_100a) back end for the element-assembly hot
 * path of romeric/florence.  Plain pointers and sizes only; every array pointer is a DEVICE pointer (obtained by the
 * host layer from DLPack capsules / torch tensors) unless its name ends in `_host`.  All functions return 0 on
 * success or a negative FL_ERR_* code (the reference's natives return void and abort).
 * Launches are asynchronous on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference checkout).
 */
#ifndef FLORENCE_B200_H
#define FLORENCE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built -fvisibility=hidden; only this ABI is exported */
#endif

#define FL_OK 0
#define FL_ERR_INVALID (-1)     /* bad argument / unsupported shape                                  */
#define FL_ERR_UNSUPPORTED (-2) /* material or element without a device kernel -> NotImplementedError */
#define FL_ERR_CUDA (-3)        /* CUDA runtime error, see fl_last_error()                            */
#define FL_ERR_STATE (-4)       /* call order (e.g. CSR assembly before fl_pattern_build)             */

/* material numbers: Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.pyx:72-109 */
#define FL_MAT_EXPLICIT_MOONEY_RIVLIN 0
#define FL_MAT_NEOHOOKEAN 1
#define FL_MAT_MOONEY_RIVLIN 2
#define FL_MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN 3
#define FL_MAT_ELECTRO_101 4
#define FL_MAT_ELECTRO_105 5
#define FL_MAT_ELECTRO_108 8
#define FL_MAT_EXPLICIT_ELECTRO_108 9
#define FL_MAT_LINEAR_ELASTIC 10

/* scatter modes of fill_global_data (Florence/VariationalPrinciple/_Mass_/_MassIntegrand_.h:115-166) */
#define FL_MODE_COO 0 /* recompute_sparsity_pattern=True : (I,J,V) triplets, ndof^2 per element, no summation */
#define FL_MODE_CSR 1 /* recompute_sparsity_pattern=False: V aligned with the CSR pattern (indptr/indices)    */

typedef struct fl_handle fl_handle;

/* The mesh + function-space tables a reference assembler receives on every call
 * (_LowLevelAssemblyDF_.pyx:56-70): cached once on the device by fl_create. */
typedef struct {
    int32_t ndim;           /* 2 or 3                                                              */
    int32_t nodeperelem;    /* mesh.elements.shape[1]                                              */
    int32_t ngauss;         /* function_space.AllGauss.shape[0]                                    */
    int32_t reserved;
    int64_t nelem;          /* mesh.nelem                                                          */
    int64_t nnode;          /* mesh.points.shape[0]                                                */
    const double *points;   /* (nnode x ndim)  mesh.points, C order                                */
    const uint64_t *elements; /* (nelem x nodeperelem) mesh.elements, uint64 as Mesh.ChangeType leaves it */
    const double *bases;    /* (nodeperelem x ngauss) function_space.Bases                          */
    const double *Jm;       /* (ndim x nodeperelem x ngauss) function_space.Jm                      */
    const double *AllGauss; /* (ngauss) function_space.AllGauss                                     */
} fl_mesh_desc;

/* Material constants in the order of the reference C signature (_LowLevelAssemblyDF_.pyx:38-48). */
typedef struct {
    int32_t material_number; /* FL_MAT_*                                                            */
    int32_t reserved;
    double rho, mu, mu1, mu2, mu3, mue, lamb, eps_1, eps_2, eps_3, eps_e;
} fl_material;

const char *fl_last_error(void);
int fl_version(void);

/* Handle: device copies of connectivity (int32), tables, node->element adjacency for the deterministic
 * segmented reductions, scratch.  Replaces the per-call unpacking of _LowLevelAssemblyDF_.pyx:56-113. */
int fl_create(const fl_mesh_desc *mesh, fl_handle **out);
int fl_destroy(fl_handle *h);

/* Matrix-free internal force: _GlobalAssemblyExplicit_DF_DPF_
 * (Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.h:728-824; .pyx:44-152).
 * formulation_number 0 = mechanics (nvar=ndim), 1 = electro_mechanics (nvar=ndim+1).
 * T (nnode*nvar) is overwritten (the reference wrapper zero-fills then accumulates).  Eulerp may be NULL for mechanics. */
int fl_assemble_explicit(fl_handle *h, const double *Eulerx, const double *Eulerp, const fl_material *mat,
                         int formulation_number, double *T, void *stream);

/* Sparsity pattern: ComputeSparsityPattern (Florence/FiniteElements/Assembly/_Assembly_/ComputeSparsityPattern.pyx:44-112,
 * .h:22-62).  fl_pattern_build builds the node-level pattern on the device and returns nnz of the (nvar*nnode)^2 CSR
 * matrix; fl_pattern_export writes int32 indptr (nvar*nnode+1) and indices (nnz), bit-identical to the reference's. */
int fl_pattern_build(fl_handle *h, int nvar, int64_t *nnz_host);
int fl_pattern_export(fl_handle *h, int nvar, int32_t *indptr, int32_t *indices, void *stream);
/* data_local_indices / data_global_indices (ComputeSparsityPattern.h:67-122), each ndof^2*nelem int32 (optional export;
 * the device path uses its own compact node-rank map instead). */
int fl_pattern_export_data_indices(fl_handle *h, int nvar, int32_t *data_local_indices, int32_t *data_global_indices, void *stream);

/* Rigid-plane penalty contact of the explicit solver: ExplicitPenaltyContactFormulation.AssembleTractions
 * (Florence/VariationalPrinciple/ExplicitPenaltyContactFormulation.py:145-184), added to the internal forces after every
 * AssembleExplicit (Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:190-197).
 * fl_set_contact: surface_nodes (device int32, the unique nodes of mesh.faces / mesh.edges), plane_normal (HOST, ndim doubles),
 *   distance, kappa, contact_gap_tolerance as the formulation's members; n_surface = 0 switches contact off.
 *   While set, fl_explicit_steps adds the contact tractions to every internal force it evaluates (and to the T it returns).
 * fl_assemble_contact: accumulate = 0 -> T = T_contact (zeros elsewhere, what AssembleTractions returns);
 *                      accumulate = 1 -> T += T_contact. */
int fl_set_contact(fl_handle *h, const int32_t *surface_nodes, int64_t n_surface, const double *plane_normal, double distance,
                   double kappa, double contact_gap_tolerance, void *stream);
int fl_assemble_contact(fl_handle *h, const double *Eulerx, double *T, int accumulate, void *stream);

/* Dirichlet reduction on the device: BoundaryCondition.GetReducedMatrices and
 * BoundaryCondition.ApplyDirichletGetReducedMatrices (Florence/BoundaryCondition/BoundaryCondition.py:842-858, :861-891),
 * which the reference runs with scipy fancy indexing on the host in every Newton iteration (Florence/Solver/FEMSolver.py:951).
 * fl_dirichlet_build: columns_out (device, int32, strictly ascending) are the prescribed dofs (BoundaryCondition.py:375-396);
 *   builds the free/prescribed maps and the pattern of K[columns_in,:][:,columns_in]; returns n_in and its nnz.
 *   Requires fl_pattern_build(nvar).
 * fl_dirichlet_export: int32 indptr_b (n_in+1), indices_b (nnz_b), columns_in (n_in); any pointer may be NULL.
 * fl_dirichlet_apply: V = CSR values aligned with fl_pattern_export (stiffness or consistent mass).
 *   V_b (nnz_b, may be NULL)        <- values of V[columns_in,:][:,columns_in]                         (:850, :884)
 *   applied (n_out, may be NULL)    -> F[columns_in] -= (V[columns_in,:][:,columns_out[nz]] @ applied[nz]) * load_factor,
 *                                      nz = ~isclose(applied, 0), summed like scipy's csr_matvec          (:873-875)
 *   F (nvar*nnode, in place; may be NULL when neither applied nor F_b is given), F_b (n_in, may be NULL) <- F[columns_in]. */
int fl_dirichlet_build(fl_handle *h, int nvar, const int32_t *columns_out, int64_t n_out, int64_t *n_in_host, int64_t *nnz_b_host);
int fl_dirichlet_export(fl_handle *h, int32_t *indptr_b, int32_t *indices_b, int32_t *columns_in, void *stream);
int fl_dirichlet_apply(fl_handle *h, const double *V, double *V_b, const double *applied, double load_factor, double *F,
                       double *F_b, void *stream);

/* Implicit K and T: _GlobalAssemblyDF_<Material> / _GlobalAssemblyDPF_<Material>
 * (Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyDF_.h:8-176, _LowLevelAssemblyDPF_.h:45-198).
 * mode FL_MODE_COO: I, J (int32) and V have ndof^2*nelem entries, element-major, row-major within the element
 *                   (fill_triplet, _MassIntegrand_.h:69-110).
 * mode FL_MODE_CSR: I, J ignored; V has nnz entries aligned with fl_pattern_export (requires fl_pattern_build(nvar)).
 * V and T (nnode*nvar) are overwritten. */
int fl_assemble_implicit(fl_handle *h, const double *Eulerx, const double *Eulerp, const fl_material *mat,
                         int formulation_number, int requires_geometry_update, int mode, int32_t *I, int32_t *J, double *V,
                         double *T, void *stream);

/* Poisson stiffness: _GlobalAssemblyPerfectLaplacian_
 * (Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyPerfectLaplacian_.h:834-916; .pyx:81-82).
 * e_tensor_host: (ndim x ndim) HOST array, the tensor the wrapper passes (= -material.e). */
int fl_assemble_laplacian(fl_handle *h, const double *e_tensor_host, int is_hessian_symmetric, int mode, int32_t *I, int32_t *J,
                          double *V, void *stream);

/* Mass: __TotalConstantMassIntegrand__ (Florence/VariationalPrinciple/_Mass_/_MassIntegrand_.pyx:192-349, .h:249-395).
 * mass_type 0 = lumped: mass (nnode*nvar) overwritten; 1 = consistent: COO/CSR like fl_assemble_implicit. */
int fl_assemble_mass(fl_handle *h, double rho, int nvar, int mass_type, int mode, double *mass, int32_t *I, int32_t *J, double *V,
                     void *stream);

/* Device-resident explicit central-difference steps, lumped mass, mechanics:
 * ExplicitStructuralDynamicIntegrator.Solver time loop (Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:95-188).
 * State vectors are length nnode*ndim.  One call advances `nsteps` increments starting at `increment`:
 *   R = F_ext(inc) - T + (2/dt^2) M U0 - (1/dt^2) M U00 ; U = dt^2 invM R ; U[fixed] = 0 ; Eulerx = X + U + IncDirichlet ;
 *   U00 <- U0 <- U ; T = internal force(Eulerx)
 * F_ext(inc) = fext_scale0 + (inc) * fext_scale_step times `fext` (ramp loading, :121-128).  fixed_mask: uint8, 1 = Dirichlet dof;
 * inc_dirichlet (nnode*ndim, may be NULL): prescribed displacement of fixed dofs.  T holds the internal force of the current
 * Eulerx on entry and exit. */
typedef struct {
    double dt;
    double fext_scale0, fext_scale_step;   /* F_ext(inc)       = (fext_scale0 + inc * fext_scale_step) * fext           (:121-128) */
    double incd_scale0, incd_scale_step;   /* IncDirichlet(inc) = (incd_scale0 + inc * incd_scale_step) * inc_dirichlet  (:101-108) */
    int64_t increment, nsteps;
} fl_explicit_ctrl;
/* status_host (2 x int32, may be NULL): [0] bit 0 = a NaN was produced, bit 1 = abs(U.max() / (U0.max() + 1e-14)) exceeded the
 * reference's tolerance (1e200 before increment 5, 10 afterwards, :175-180); [1] = increment of the first detection. */
int fl_explicit_steps(fl_handle *h, const fl_material *mat, const fl_explicit_ctrl *ctrl, const double *M, const double *fext,
                      const uint8_t *fixed_mask, const double *inc_dirichlet, double *U0, double *U00, double *Eulerx, double *T,
                      int32_t *status_host, void *stream);

/* ---- Element-partitioned runs: one process per GPU replaces the reference's process pool / MPI launchers
 * (Florence/FiniteElements/Assembly/Assembly.py:1126-1358).  Each rank owns a contiguous block of elements (Mesh.Partition,
 * Florence/MeshGeneration/Mesh.py:7395-7447) and every node it touches; only the partial internal forces of INTERFACE nodes
 * cross NVLink (the reference broadcasts the whole Eulerx and reduces the whole T every step, Assembly.py:1153-1167). */

/* gather / scatter of a nodal vector at a list of local node ids: T[node_ids] -> buf, T[node_ids] += buf, T[node_ids] = buf */
int fl_pack_nodes(const double *T, const int32_t *node_ids, int64_t n, int nvar, double *buf, void *stream);
int fl_unpack_add_nodes(double *T, const int32_t *node_ids, int64_t n, int nvar, const double *buf, void *stream);
int fl_scatter_nodes(double *T, const int32_t *node_ids, int64_t n, int nvar, const double *buf, void *stream);
/* out[k] (nvar doubles) = sum over j in [ptr[k], ptr[k+1]) of all[idx[j] .. idx[j]+nvar), added in list order.  The host layer
 * lists the contributions of every rank sharing an interface node in ascending rank order, so all sharers compute bit-identical
 * sums (`T_all[pnodes] += T_p`, Assembly.py:1352-1354, made independent of the summation order of the ranks). */
int fl_sum_ordered(const double *all, const int64_t *ptr, const int64_t *idx, int64_t n, int nvar, double *out, void *stream);

/* Split form of fl_explicit_steps (mechanics): per step  update -> forces(interface elements) -> gather_pack -> [exchange, host layer]
 * -> forces(remaining elements, overlapping the exchange) -> sum_ordered -> next update.
 * fl_explicit_forces: per-element internal forces of elements [e0, e1) into the handle's scratch (the element kernel of
 *   fl_assemble_explicit on a sub-range; _LowLevelAssemblyExplicit_DF_DPF_.h:458-717).
 * fl_gather_pack_nodes: buf[k] = nodal sum of that scratch at node_ids[k] (ascending element order).
 * fl_gather_nodes: the full nodal reduction T (nnode*nvar) of the scratch (RHSAssemblyNative.pyx:30-39). */
int fl_explicit_forces(fl_handle *h, const double *Eulerx, const fl_material *mat, int64_t e0, int64_t e1, void *stream);
int fl_gather_pack_nodes(fl_handle *h, int nvar, const int32_t *node_ids, int64_t n, double *buf, void *stream);
int fl_gather_nodes(fl_handle *h, int nvar, double *T, void *stream);
/* fl_explicit_update: one central-difference update (ExplicitStructuralDynamicIntegrator.py:131-157).
 *   use_element_forces = 0: the internal force is the caller's T (already reduced and exchanged);
 *   use_element_forces = 1: it is reduced here from the scratch of fl_explicit_forces, except at nodes with iface_slot[n] >= 0,
 *     where T_iface[iface_slot[n]] (the rank-ordered interface sum) is used; write_T stores the reduced force in T.
 *   status_dev (2 x int32) / growth_keys_dev (2 x int64, may be NULL): see fl_explicit_check. */
typedef struct {
    double dt, fext_scale, incd_scale;
    const double *M, *fext;
    const uint8_t *fixed_mask;
    const double *inc_dirichlet;
    double *T;
    const int32_t *iface_slot;
    const double *T_iface;
    double *U0, *U00, *Eulerx;
    int32_t *status_dev;
    int64_t *growth_keys_dev;
    int32_t use_element_forces, write_T;
} fl_update_args;
int fl_explicit_update(fl_handle *h, const fl_update_args *args, void *stream);
/* Blow-up test of the explicit loop (:175-180) on the running maxima the update kernel left in growth_keys_dev (order-preserving
 * int64 images of U.max() and U0.max(); initialise both to INT64_MIN; the host layer takes the MAX over ranks first):
 * sets status_dev[0] bit 1 / status_dev[1] = increment and resets the maxima. */
int fl_explicit_check(fl_handle *h, int64_t *growth_keys_dev, int64_t increment, int32_t *status_dev, void *stream);

/* Implicit multi-GPU: "each rank emits the CSR row block it owns" (the reference sums per-partition triplets on
 * the parent, Assembly.py:1000-1041).  owned_nodes: ascending local node ids owned by this rank; node_map: global node id of every
 * local node.  fl_row_block_build writes the int64 row pointer of the block (n_owned*nvar+1 entries, starting at 0) and returns
 * its nnz; fl_row_block_emit copies the owned rows of V (aligned with fl_pattern_export) into vals and writes GLOBAL column dof
 * numbers into cols_global.  Either output may be NULL: the columns belong to the pattern and are emitted once, the values on
 * every assembly. */
int fl_row_block_build(fl_handle *h, int nvar, const int32_t *owned_nodes, int64_t n_owned, int64_t *indptr_block,
                       int64_t *nnz_block_host, void *stream);
int fl_row_block_emit(fl_handle *h, int nvar, const double *V, const int32_t *owned_nodes, int64_t n_owned, const int64_t *node_map,
                      const int64_t *indptr_block, int64_t *cols_global, double *vals, void *stream);

/* Space-filling-curve element order (BASELINE north_star; Mesh.Partition cuts np.array_split(arange(nelem)), Mesh.py:7403, so the
 * element order decides the interface size): perm (nelem int64) = elements sorted by the Morton key of their centroid, stable.
 * points / elements are DEVICE arrays as in fl_mesh_desc. */
int fl_sfc_order(const double *points, const uint64_t *elements, int64_t nelem, int nodeperelem, int ndim, int64_t nnode,
                 int64_t *perm, void *stream);

/* Tuning switches (for tests and A/B timing): option 0 = use the tensor-core (DMMA) explicit kernels for
 * hex8/hex27 (default 1); option 1 = use the DMMA implicit kernels for hex64 and electro-mechanical tet20 (default 1; 2 = also hex27 and mechanical tet20; 3 = hex64 with
 * the K_e scratch as dof-pair planes instead of per-row-node planes); option 2 = use the
 * warp-autonomous LinearElastic kernel: 1 = tet10 (default), 2 = tet10 and hex8, 0 = off; option 3 = CSR value reduction of the
 * element-order paths: 0 = shared-memory row-buffer kernels, 1 = register-resident slot-owner gather (nvar 2..4, low-order
 * elements whose K_e row blocks are multiples of 16 bytes), 2 = whichever was measured faster for the shape (default: the gather for
 * 2-D elements, hex8 mechanics and tet10 electro-mechanics); option 4 = tet10 mechanics (any material; hex8 LinearElastic when option 2 is 2) in CSR
 * mode: 1 = K_e stored along a space-filling curve and reduced in completion order (default), 2 = the same with the element kernel
 * and the reduction running concurrently on two streams (measured slower), 3 = the kernels of 2 one after the other, 0 = off;
 * option 5 = 1: hex64 / nvar 4 CSR reduction without the cross-node software pipeline (A/B timing; default 0). */
int fl_set_option(fl_handle *h, int option, int value);

/* Per-kernel device timing of the most recent fl_assemble_* call (CUDA events recorded on the call's stream):
 * ms[0] = element kernel, ms[1] = stiffness scatter / CSR reduction, ms[2] = nodal (T) reduction.  Used by bench.py for
 * the roofline of the dominant kernel.  fl_get_timing synchronises on the last event. */
int fl_set_timing(fl_handle *h, int enabled);
int fl_get_timing(fl_handle *h, float *ms3_host);

/* fp64 pipe peaks of the device the roofline fractions are quoted against (measured, not nominal):
 * runs a register-resident DFMA loop / DMMA loop for `iters` iterations and returns TFLOP/s. */
int fl_measure_fp64_peak(int use_dmma, int iters, double *tflops_host);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FLORENCE_B200_H */
