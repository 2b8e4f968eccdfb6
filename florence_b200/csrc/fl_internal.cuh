// fl_internal.cuh -- handle layout and launcher declarations shared by the translation units of libflorence_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/florence_b200.h"
#include "fl_math.cuh"

namespace fl {

void set_error(const char* fmt, ...);

#define FL_CUDA_CHECK(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            fl::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return FL_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

// Node-level sparsity pattern (a14): row-node n has neighbour nodes nbr_idx[nbr_ptr[n] .. nbr_ptr[n+1]) ascending.
// The CSR pattern of the (nvar*nnode)^2 matrix follows as indptr[nvar*n+i] = nvar*(nvar*nbr_ptr[n] + i*cnt_n),
// indices = nvar*m+l.  rank[(e*npe+a)*npe+b] = position of node conn[e][b] in the neighbour list of conn[e][a].
struct Pattern {
    int64_t nnzb = 0;          // number of node pairs
    int64_t* nbr_ptr = nullptr;  // nnode+1
    int32_t* nbr_idx = nullptr;  // nnzb
    uint16_t* rank = nullptr;    // nelem*npe*npe
    uint16_t* rank_adj = nullptr;  // the same rows in visit (adjacency) order: contiguous per node for the CSR reduction
    int max_cnt = 0;           // widest row in nodes
};

// Dirichlet reduction (fl_dirichlet.cu): free/prescribed dof maps and the row pointer of K[columns_in][:, columns_in]
struct Dirichlet {
    int nvar = 0;
    int64_t n_in = 0, n_out = 0, nnz_b = 0;
    int32_t* new_id = nullptr;    // nvar*nnode: reduced index of a free dof, -(k+1) for the k-th prescribed dof
    int32_t* cols_in = nullptr;   // n_in
    int64_t* rowptr_b = nullptr;  // n_in+1
    uint8_t* cmask = nullptr;     // per node pair of the pattern: bit l set <=> column dof l of the neighbour node is free
};

// Rigid-plane penalty contact (ExplicitPenaltyContactFormulation.py:145-184): nodes flagged in `surf` (the unique nodes of the
// boundary faces) with gap = x.n + L < tol receive the force kappa*gap*n.
struct Contact {
    uint8_t* surf = nullptr;   // nnode flags; nullptr = no contact
    double n[3] = {0, 0, 0};
    double L = 0, kappa = 0, tol = 0;
};

// Plan of the register-resident CSR value reduction (fl_gather.cuh): the neighbour list of every node cut into work items of up to 96 slots,
// and per (item, visit) one packed record: flat connectivity index + slot -> local-node fields.  All pointers are device arrays.
struct GatherItem {
    int64_t rec_off;    // first record of the item, in 32-bit words from GatherPlan::recs
    int64_t v_base;     // offset in V of row 0, first column of the item
    int32_t nvis;       // visits (elements) of the node
    int32_t w;          // row width nvar * neighbours
    int32_t nslots;     // node slots (neighbour columns) of the item (<= 96)
    int32_t ng;         // groups of 32 slots the item spans (1..3); its records have 2 + FWG*ng words
};
struct GatherPlan {
    int nvar = 0, bits = 0;
    int64_t nitems = 0, nwords = 0;
    GatherItem* items = nullptr;    // nitems, node-major (streamed assembly: in completion order)
    uint32_t* recs = nullptr;       // nwords: the records of all items, [item][visit][2 + FWG*ng words]
};

// Curve-ordered CSR assembly (fl_stream.cu): the element kernel walks the elements in space-filling-curve order; the CSR reduction
// follows in completion order, optionally beside it on a second stream (then the element kernel publishes its progress).
struct StreamPlan {
    int npe = 0, nvar = 0;
    int64_t ngroups = 0;
    int32_t* conn_p = nullptr;      // connectivity in storage (curve) order
    int32_t* adj_idx_p = nullptr;   // adjacency (node -> visits, ascending ORIGINAL element number) as storage flat indices
    GatherPlan gp;                  // records hold storage flat indices, items in completion order
    int32_t* flags = nullptr;       // ngroups: epoch of the call whose K_e rows of the group are complete
    int32_t* err = nullptr;         // set by a reduction warp that gave up waiting
    int32_t epoch = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

template <int D>
__device__ __forceinline__ bool contact_force(const Contact& c, const double* __restrict__ x, double* f) {
    double gap = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) gap = __dadd_rn(gap, __dmul_rn(x[i], c.n[i]));
    gap = __dadd_rn(gap, c.L);
    if (!(gap < c.tol)) return false;
#pragma unroll
    for (int i = 0; i < D; ++i) f[i] = __dmul_rn(c.kappa, __dmul_rn(gap, c.n[i]));
    return true;
}

}  // namespace fl

struct fl_handle {
    int ndim = 0, npe = 0, ng = 0;
    int64_t nelem = 0, nnode = 0;
    int ldg = 0;                 // padded (odd) leading dimension of the Jm table over gauss points
    // device copies
    int32_t* conn = nullptr;     // nelem x npe
    double* points = nullptr;    // nnode x ndim
    double* jm = nullptr;        // [k][a][ldg]: Jm[k][a][g]
    double* jmT = nullptr;       // [g][k][a]  : the same table, node index fastest
    double* bases = nullptr;     // [a][ng]
    double* gw = nullptr;        // [ng]
    // node -> flat connectivity index (e*npe+a), ascending in e: the order the reference's element loop sums in
    int64_t* adj_ptr = nullptr;  // nnode+1
    int32_t* adj_idx = nullptr;  // nelem*npe
    int max_adj = 0;
    fl::Pattern pat;
    fl::Dirichlet dir;
    fl::Contact contact;
    // scratch (grown on demand)
    double* te = nullptr;  size_t te_bytes = 0;   // per-element traction buffer nelem*ndof
    double* ke = nullptr;  size_t ke_bytes = 0;   // per-element stiffness buffer nelem*ndof^2 (CSR mode)
    fl::GatherPlan gplan;        // built on the first CSR reduction that can use it (fl_gather.cu)
    fl::StreamPlan splan;
    int use_stream = 1;          // LinearElastic tet10 (hex8 with option 2 = 2) in CSR mode (fl_stream.cu, fl_set_option 4): 1 = K_e along
                                 // a space-filling curve, reduction in completion order; 2 = the same with both kernels running
                                 // concurrently (measured slower); 3 = the kernels of 2 one after the other; 0 = off
    int use_reg_gather = 2;      // CSR value reduction of the element-order paths (fl_set_option 3): 1 = register-resident slot-owner
                                 // gather (fl_gather.cuh) wherever it is instantiated, 0 = shared-memory row-buffer kernels of
                                 // fl_pattern.cu, 2 = whichever was measured faster for the shape (reg_gather_preferred)
    double* ch = nullptr;  size_t ch_bytes = 0;   // per-element Chat_g blocks between the prologue and the DMMA kernel (p >= 2 hexahedra)
    int32_t* flag = nullptr;                       // device status: [0] bit 0 NaN, bit 1 growth blow-up; [1] increment of first detection
    int64_t* growth = nullptr;                     // running maxima (ordered keys) of U and U0 for the blow-up test
    int64_t el0 = 0, el1 = -1;                     // element range of the explicit force call in flight (el1 < 0: all)
    int sm_count = 148;
    int max_smem_optin = 0;
    int timing = 0;
    int use_mma = 1;             // explicit path: DMMA kernels for hex8 / hex27 (fl_set_option 0)
    int ke_plane_major = 0;      // layout of the CSR-mode K_e scratch for the call in flight: 0 = [(a,i)][(b,j)] (the COO layout),
                                 // 1 = dof-pair planes [(i,j)][a][b] (DMMA kernels: every fragment store fills whole sectors),
                                 // 2 = per-row-node planes [a][(i,j)][b] (same stores; one visit of the reduction is one contiguous run)
    int use_warp_iso = 1;        // implicit path: warp-autonomous LinearElastic kernel, 1 = tet10, 2 = tet10 + hex8 (fl_set_option 2)
    int wide_unpipelined = 0;    // 1: hex64 / nvar 4 CSR reduction without the cross-node software pipeline (fl_set_option 5; A/B timing)
    int use_mma_implicit = 1;    // implicit path: DMMA kernels for hex64 / electro tet20; 2 = also hex27 / mechanics tet20 (fl_set_option 1)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace fl {

int ensure_scratch(double** p, size_t* have, size_t need);

// fl_explicit.cu
int launch_explicit_elements(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation,
                             double* te, cudaStream_t st);
int launch_gather_nodes(fl_handle* h, int nvar, const double* te, double* T, cudaStream_t st);
int launch_explicit_update(fl_handle* h, int fused_gather, const double* te, double dt, double fext_scale, double incd_scale,
                           const double* M, const double* fext, const uint8_t* fixed, const double* inc_dir, double* T,
                           const int32_t* iface_slot, const double* T_iface, double* U0, double* U00, double* Eulerx, int32_t* nan_flag,
                           int64_t* growth, cudaStream_t st);
int launch_gather_pack(fl_handle* h, int nvar, const double* te, const int32_t* ids, int64_t n, double* buf, cudaStream_t st);
int launch_growth_check(int64_t* growth, int64_t increment, int32_t* status, cudaStream_t st);
int launch_contact(fl_handle* h, const double* Eulerx, double* T, int accumulate, cudaStream_t st);
// fl_implicit.cu
int launch_implicit_elements(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation,
                             int update, double* ke, double* te, cudaStream_t st);
int launch_laplacian_elements(fl_handle* h, const double* e_tensor, int symmetric, double* ke, cudaStream_t st);
int launch_mass_elements(fl_handle* h, double rho, int nvar, int lumped, double* out, cudaStream_t st);
// fl_pattern.cu
int build_adjacency(fl_handle* h);
int pattern_build(fl_handle* h);
int launch_pattern_export(fl_handle* h, int nvar, int32_t* indptr, int32_t* indices, cudaStream_t st);
int launch_data_indices(fl_handle* h, int nvar, int32_t* dl, int32_t* dg, cudaStream_t st);
int launch_coo_indices(fl_handle* h, int nvar, int32_t* I, int32_t* J, cudaStream_t st);
int launch_csr_gather(fl_handle* h, int nvar, const double* ke, double* V, cudaStream_t st);
int launch_row_block_build(fl_handle* h, int nvar, const int32_t* owned, int64_t n_owned, int64_t* indptr_block, int64_t* nnz_host,
                           cudaStream_t st);
int launch_row_block_emit(fl_handle* h, int nvar, const double* V, const int32_t* owned, int64_t n_owned, const int64_t* node_map,
                          const int64_t* indptr_block, int64_t* cols, double* vals, cudaStream_t st);
int launch_sfc_order(const double* points, const int64_t* elements, int64_t nelem, int npe, int ndim, int64_t nnode, int64_t* perm,
                     cudaStream_t st);
int launch_sfc_order_conn(const double* points, const int32_t* conn, int64_t nelem, int npe, int ndim, int64_t nnode, int64_t* perm,
                          cudaStream_t st);
// fl_gather.cu
void gather_plan_free(fl_handle* h);
void gather_plan_release(GatherPlan& g);
int gather_plan_build(fl_handle* h, int nvar, const int32_t* flat_store, bool by_completion, GatherPlan* out);
bool reg_gather_supported(const fl_handle* h, int nvar);
bool reg_gather_preferred(const fl_handle* h, int nvar);
int gather_plan_ensure(fl_handle* h, int nvar);
int launch_csr_gather_reg(fl_handle* h, int nvar, const double* ke, double* V, cudaStream_t st);
// fl_stream.cu
void stream_plan_free(fl_handle* h);
bool stream_csr_supported(const fl_handle* h);
int launch_stream_iso_csr(fl_handle* h, const double* Eulerx, const fl_material* mat, int update, double* V, double* T, cudaStream_t st);
bool curve_csr_preferred(const fl_handle* h, int nvar);
int launch_curve_csr(fl_handle* h, int nvar, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation, int update,
                     double* V, double* T, cudaStream_t st, cudaEvent_t after_elements, cudaEvent_t after_reduction);
// fl_dirichlet.cu
void dirichlet_free(fl_handle* h);
int dirichlet_build(fl_handle* h, int nvar, const int32_t* cols_out, int64_t n_out);
int launch_dirichlet_export(fl_handle* h, int32_t* indptr_b, int32_t* indices_b, int32_t* columns_in, cudaStream_t st);
int launch_dirichlet_apply(fl_handle* h, const double* V, double* V_b, const double* applied, double load_factor, double* F, double* F_b,
                           cudaStream_t st);

}  // namespace fl
