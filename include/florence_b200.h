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
 * Eulerx on entry and exit.  status_host[0] is set to 1 if a NaN was produced (blow-up test, :175-180). */
typedef struct {
    double dt;
    double fext_scale0, fext_scale_step;
    int64_t increment, nsteps;
} fl_explicit_ctrl;
int fl_explicit_steps(fl_handle *h, const fl_material *mat, const fl_explicit_ctrl *ctrl, const double *M, const double *fext,
                      const uint8_t *fixed_mask, const double *inc_dirichlet, double *U0, double *U00, double *Eulerx, double *T,
                      int32_t *status_host, void *stream);

/* Interface exchange support for element-partitioned meshes (replaces the whole-vector Bcast/Reduce of
 * Florence/FiniteElements/Assembly/Assembly.py:1135-1167): gather/scatter-add of T at a list of local node ids. */
int fl_pack_nodes(const double *T, const int32_t *node_ids, int64_t n, int nvar, double *buf, void *stream);
int fl_unpack_add_nodes(double *T, const int32_t *node_ids, int64_t n, int nvar, const double *buf, void *stream);

/* Split form of fl_explicit_steps for multi-GPU runs (force -> exchange -> update):
 * fl_explicit_update applies one central-difference update given the (already exchanged) internal force T. */
int fl_explicit_update(fl_handle *h, double dt, double fext_scale, const double *M, const double *fext, const uint8_t *fixed_mask,
                       const double *inc_dirichlet, const double *T, double *U0, double *U00, double *Eulerx, int32_t *nan_flag_dev,
                       void *stream);

/* Tuning switches (for tests and A/B timing): option 0 = use the tensor-core (DMMA) explicit kernels for
 * hex8/hex27 (default 1); option 1 = use the DMMA implicit kernels for hex27/hex64 (default 1); option 2 = use the
 * warp-autonomous LinearElastic kernel: 1 = tet10 (default), 2 = tet10 and hex8, 0 = off. */
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
