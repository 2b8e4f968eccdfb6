// fl_api.cu -- the C ABI of libflorence_b200.so (include/florence_b200.h): handle management, argument checks and the
// sequencing of kernels for each reference entry point.
#include <cstdarg>
#include <cstring>

#include "fl_internal.cuh"

namespace fl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ensure_scratch(double** p, size_t* have, size_t need) {
    if (*have >= need && *p) return FL_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    if (need == 0) need = 8;
    FL_CUDA_CHECK(cudaMalloc(p, need));
    *have = need;
    return FL_OK;
}

__global__ void narrow_conn_kernel(const uint64_t* __restrict__ in, int64_t n, int64_t nnode, int32_t* __restrict__ out, int* __restrict__ bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t v = in[i];
    if (v >= (uint64_t)nnode) *bad = 1;
    out[i] = (int32_t)v;
}

__global__ void transpose_jm_kernel(const double* __restrict__ in, int D, int npe, int ng, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * npe * ng) return;
    const int g = i / (D * npe), r = i - g * D * npe, k = r / npe, a = r - k * npe;
    out[i] = in[(k * npe + a) * ng + g];
}

__global__ void pad_jm_kernel(const double* __restrict__ in, int rows, int ng, int ldg, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ldg) return;
    const int r = i / ldg, g = i - r * ldg;
    out[i] = g < ng ? in[r * ng + g] : 0.0;
}

// ---- fp64 pipe peak micro-benchmarks -------------------------------------------------------------------
__global__ void dfma_peak_kernel(int iters, double* out) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}

__global__ void dmma_peak_kernel(int iters, double* out) {
    // mma.sync.aligned.m8n8k4.row.col.f64: 8x8x4 = 256 FMAs per warp instruction
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

}  // namespace fl

using namespace fl;

extern "C" {

const char* fl_last_error(void) { return g_err; }
int fl_version(void) { return 100; }

int fl_create(const fl_mesh_desc* m, fl_handle** out) {
    if (!m || !out) { set_error("null argument"); return FL_ERR_INVALID; }
    if (m->ndim != 2 && m->ndim != 3) { set_error("ndim must be 2 or 3, got %d", m->ndim); return FL_ERR_INVALID; }
    if (m->nodeperelem < 1 || m->ngauss < 1 || m->nelem < 0 || m->nnode < 1) {
        set_error("bad sizes: nodeperelem=%d ngauss=%d nelem=%lld nnode=%lld", m->nodeperelem, m->ngauss, (long long)m->nelem, (long long)m->nnode);
        return FL_ERR_INVALID;
    }
    if (m->nnode >= (int64_t)1 << 31) { set_error("nnode exceeds int32 indexing"); return FL_ERR_INVALID; }
    if (!m->points || !m->Jm || !m->AllGauss || (m->nelem > 0 && !m->elements)) { set_error("null mesh/table pointer"); return FL_ERR_INVALID; }
    fl_handle* h = new fl_handle();
    h->ndim = m->ndim; h->npe = m->nodeperelem; h->ng = m->ngauss; h->nelem = m->nelem; h->nnode = m->nnode;
    h->ldg = m->ngauss | 1;
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(ce)); delete h; return FL_ERR_CUDA; }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int64_t nk = h->nelem * h->npe;
    int rc = FL_OK;
    auto fail = [&](int code) { fl_destroy(h); return code; };
#define FL_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                            \
            return fail(FL_ERR_CUDA);                                                             \
        }                                                                                         \
    } while (0)
    FL_TRY(cudaMalloc(&h->conn, sizeof(int32_t) * (nk > 0 ? nk : 1)));
    FL_TRY(cudaMalloc(&h->points, sizeof(double) * h->nnode * h->ndim));
    FL_TRY(cudaMalloc(&h->jm, sizeof(double) * h->ndim * h->npe * h->ldg));
    FL_TRY(cudaMalloc(&h->jmT, sizeof(double) * h->ndim * h->npe * h->ng));
    FL_TRY(cudaMalloc(&h->bases, sizeof(double) * h->npe * h->ng));
    FL_TRY(cudaMalloc(&h->gw, sizeof(double) * h->ng));
    FL_TRY(cudaMalloc(&h->flag, 4 * sizeof(int32_t)));
    FL_TRY(cudaMemset(h->flag, 0, 4 * sizeof(int32_t)));
    FL_TRY(cudaMalloc(&h->growth, 2 * sizeof(int64_t)));
    {
        const int64_t lowest[2] = {INT64_MIN, INT64_MIN};
        FL_TRY(cudaMemcpy(h->growth, lowest, sizeof(lowest), cudaMemcpyHostToDevice));
    }
    FL_TRY(cudaMemcpy(h->points, m->points, sizeof(double) * h->nnode * h->ndim, cudaMemcpyDeviceToDevice));
    FL_TRY(cudaMemcpy(h->gw, m->AllGauss, sizeof(double) * h->ng, cudaMemcpyDeviceToDevice));
    if (m->bases) FL_TRY(cudaMemcpy(h->bases, m->bases, sizeof(double) * h->npe * h->ng, cudaMemcpyDeviceToDevice));
    else FL_TRY(cudaMemset(h->bases, 0, sizeof(double) * h->npe * h->ng));
    {
        const int tot = h->ndim * h->npe * h->ldg;
        pad_jm_kernel<<<(tot + 255) / 256, 256>>>(m->Jm, h->ndim * h->npe, h->ng, h->ldg, h->jm);
        const int tot2 = h->ndim * h->npe * h->ng;
        transpose_jm_kernel<<<(tot2 + 255) / 256, 256>>>(m->Jm, h->ndim, h->npe, h->ng, h->jmT);
    }
    if (nk > 0) {
        narrow_conn_kernel<<<(unsigned)((nk + 255) / 256), 256>>>(m->elements, nk, h->nnode, h->conn, h->flag);
        int bad = 0;
        FL_TRY(cudaMemcpy(&bad, h->flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad) { set_error("mesh.elements holds a node number >= nnode"); return fail(FL_ERR_INVALID); }
    }
    FL_TRY(cudaGetLastError());
    rc = build_adjacency(h);
    if (rc) return fail(rc);
    FL_TRY(cudaDeviceSynchronize());
#undef FL_TRY
    *out = h;
    return FL_OK;
}

static inline void mark(fl_handle* h, int k, cudaStream_t st) {
    if (h->timing && h->ev[k]) cudaEventRecord(h->ev[k], st);
}

int fl_set_option(fl_handle* h, int option, int value) {
    if (!h) { set_error("null handle"); return FL_ERR_INVALID; }
    if (option == 0) { h->use_mma = value ? 1 : 0; return FL_OK; }
    if (option == 1) { h->use_mma_implicit = value; return FL_OK; }
    if (option == 5) { h->wide_unpipelined = value ? 1 : 0; return FL_OK; }
    if (option == 2) { h->use_warp_iso = value; return FL_OK; }
    if (option == 3) { h->use_reg_gather = value; return FL_OK; }
    if (option == 4) { h->use_stream = value; return FL_OK; }
    set_error("unknown option %d", option);
    return FL_ERR_INVALID;
}

int fl_set_timing(fl_handle* h, int enabled) {
    if (!h) { set_error("null handle"); return FL_ERR_INVALID; }
    if (enabled && !h->ev[0])
        for (int k = 0; k < 4; ++k) FL_CUDA_CHECK(cudaEventCreate(&h->ev[k]));
    h->timing = enabled ? 1 : 0;
    return FL_OK;
}

int fl_get_timing(fl_handle* h, float* ms) {
    if (!h || !ms || !h->ev[0]) { set_error("timing is not enabled"); return FL_ERR_STATE; }
    FL_CUDA_CHECK(cudaEventSynchronize(h->ev[3]));
    for (int k = 0; k < 3; ++k) FL_CUDA_CHECK(cudaEventElapsedTime(&ms[k], h->ev[k], h->ev[k + 1]));
    return FL_OK;
}

int fl_destroy(fl_handle* h) {
    if (!h) return FL_OK;
    for (int k = 0; k < 4; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    cudaFree(h->conn); cudaFree(h->points); cudaFree(h->jm); cudaFree(h->jmT); cudaFree(h->bases); cudaFree(h->gw);
    cudaFree(h->adj_ptr); cudaFree(h->adj_idx); cudaFree(h->pat.nbr_ptr); cudaFree(h->pat.nbr_idx); cudaFree(h->pat.rank); cudaFree(h->pat.rank_adj);
    dirichlet_free(h);
    gather_plan_free(h);
    stream_plan_free(h);
    cudaFree(h->contact.surf);
    cudaFree(h->te); cudaFree(h->ke); cudaFree(h->ch); cudaFree(h->flag); cudaFree(h->growth);
    delete h;
    return FL_OK;
}

int fl_assemble_explicit(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation_number,
                         double* T, void* stream) {
    if (!h || !Eulerx || !mat || !T) { set_error("null argument"); return FL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nvar = h->ndim + (formulation_number == 1 ? 1 : 0);
    int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * h->nelem * h->npe * nvar);
    if (rc) return rc;
    mark(h, 0, st);
    rc = launch_explicit_elements(h, Eulerx, Eulerp, mat, formulation_number, h->te, st);
    if (rc) return rc;
    mark(h, 1, st);
    mark(h, 2, st);
    rc = launch_gather_nodes(h, nvar, h->te, T, st);
    mark(h, 3, st);
    return rc;
}

int fl_pattern_build(fl_handle* h, int nvar, int64_t* nnz_host) {
    if (!h || nvar < 1 || nvar > 4) { set_error("bad argument"); return FL_ERR_INVALID; }
    int rc = pattern_build(h);
    if (rc) return rc;
    // the plan of the register-resident CSR reduction depends on the pattern and nvar only: built here, outside the assembly calls
    if (reg_gather_preferred(h, nvar)) {
        rc = gather_plan_ensure(h, nvar);
        if (rc) return rc;
    }
    if (nnz_host) *nnz_host = h->pat.nnzb * nvar * nvar;
    return FL_OK;
}

int fl_pattern_export(fl_handle* h, int nvar, int32_t* indptr, int32_t* indices, void* stream) {
    if (!h || !indptr || !indices) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_pattern_export(h, nvar, indptr, indices, (cudaStream_t)stream);
}

int fl_pattern_export_data_indices(fl_handle* h, int nvar, int32_t* dl, int32_t* dg, void* stream) {
    if (!h || !dl || !dg) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_data_indices(h, nvar, dl, dg, (cudaStream_t)stream);
}

__global__ void flag_nodes_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t nnode, uint8_t* __restrict__ flags, int32_t* bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k = ids[i];
    if (k < 0 || k >= nnode) { *bad = 1; return; }
    flags[k] = 1;
}

int fl_set_contact(fl_handle* h, const int32_t* surface_nodes, int64_t n_surface, const double* plane_normal, double distance, double kappa,
                   double contact_gap_tolerance, void* stream) {
    if (!h) { set_error("null argument"); return FL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_surface <= 0) {
        FL_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(h->contact.surf);
        h->contact = Contact();
        return FL_OK;
    }
    if (!surface_nodes || !plane_normal) { set_error("null argument"); return FL_ERR_INVALID; }
    uint8_t* flags = nullptr;
    FL_CUDA_CHECK(cudaMalloc(&flags, h->nnode > 0 ? h->nnode : 1));
    FL_CUDA_CHECK(cudaMemsetAsync(flags, 0, h->nnode, st));
    FL_CUDA_CHECK(cudaMemsetAsync(h->flag, 0, sizeof(int32_t), st));
    flag_nodes_kernel<<<(unsigned)((n_surface + 255) / 256), 256, 0, st>>>(surface_nodes, n_surface, h->nnode, flags, h->flag);
    int32_t bad = 0;
    FL_CUDA_CHECK(cudaMemcpyAsync(&bad, h->flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FL_CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad) { cudaFree(flags); set_error("surface node id out of range"); return FL_ERR_INVALID; }
    cudaFree(h->contact.surf);
    h->contact = Contact();
    h->contact.surf = flags;
    for (int i = 0; i < h->ndim; ++i) h->contact.n[i] = plane_normal[i];
    h->contact.L = distance;
    h->contact.kappa = kappa;
    h->contact.tol = contact_gap_tolerance;
    return FL_OK;
}

int fl_assemble_contact(fl_handle* h, const double* Eulerx, double* T, int accumulate, void* stream) {
    if (!h || !Eulerx || !T) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_contact(h, Eulerx, T, accumulate ? 1 : 0, (cudaStream_t)stream);
}

int fl_dirichlet_build(fl_handle* h, int nvar, const int32_t* columns_out, int64_t n_out, int64_t* n_in_host, int64_t* nnz_b_host) {
    if (!h || nvar < 1 || nvar > 4 || (n_out > 0 && !columns_out)) { set_error("bad argument"); return FL_ERR_INVALID; }
    int rc = dirichlet_build(h, nvar, columns_out, n_out);
    if (rc) return rc;
    if (n_in_host) *n_in_host = h->dir.n_in;
    if (nnz_b_host) *nnz_b_host = h->dir.nnz_b;
    return FL_OK;
}

int fl_dirichlet_export(fl_handle* h, int32_t* indptr_b, int32_t* indices_b, int32_t* columns_in, void* stream) {
    if (!h) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_dirichlet_export(h, indptr_b, indices_b, columns_in, (cudaStream_t)stream);
}

int fl_dirichlet_apply(fl_handle* h, const double* V, double* V_b, const double* applied, double load_factor, double* F, double* F_b,
                       void* stream) {
    if (!h) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_dirichlet_apply(h, V, V_b, applied, load_factor, F, F_b, (cudaStream_t)stream);
}

static int scatter_stiffness(fl_handle* h, int nvar, int mode, double* ke, int32_t* I, int32_t* J, double* V, cudaStream_t st) {
    if (mode == FL_MODE_COO) {
        if (I && J) return launch_coo_indices(h, nvar, I, J, st);
        return FL_OK;
    }
    return launch_csr_gather(h, nvar, ke, V, st);
}

int fl_assemble_implicit(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation_number,
                         int requires_geometry_update, int mode, int32_t* I, int32_t* J, double* V, double* T, void* stream) {
    if (!h || !Eulerx || !mat || !V || !T) { set_error("null argument"); return FL_ERR_INVALID; }
    if (mode != FL_MODE_COO && mode != FL_MODE_CSR) { set_error("mode must be FL_MODE_COO or FL_MODE_CSR"); return FL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nvar = h->ndim + (formulation_number == 1 ? 1 : 0);
    const size_t ndof = (size_t)h->npe * nvar;
    if (mode == FL_MODE_CSR && !h->pat.nbr_ptr) { set_error("CSR assembly requires fl_pattern_build first"); return FL_ERR_STATE; }
    int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * h->nelem * ndof);
    if (rc) return rc;
    double* ke = V;  // COO mode: the element kernel writes the triplet values in place
    if (mode == FL_MODE_CSR) {
        rc = ensure_scratch(&h->ke, &h->ke_bytes, sizeof(double) * h->nelem * ndof * ndof);
        if (rc) return rc;
        ke = h->ke;
    }
    // CSR values of the isotropic constant-tangent material on tet10 / hex8: K_e stored along a space-filling curve, reduction in
    // completion order (fl_stream.cu)
    if (mode == FL_MODE_CSR && mat->material_number == MAT_LINEAR_ELASTIC && formulation_number == 0 && stream_csr_supported(h)) {
        mark(h, 0, st);
        return launch_stream_iso_csr(h, Eulerx, mat, requires_geometry_update ? 1 : 0, V, T, st);
    }
    // other materials on the shapes for which it pays: the ordinary element kernel on the curve-ordered connectivity + the same reduction
    if (mode == FL_MODE_CSR && curve_csr_preferred(h, nvar)) {
        mark(h, 0, st);
        rc = launch_curve_csr(h, nvar, Eulerx, Eulerp, mat, formulation_number, requires_geometry_update ? 1 : 0, V, T, st,
                              h->timing ? h->ev[1] : nullptr, h->timing ? h->ev[2] : nullptr);
        mark(h, 3, st);
        return rc;
    }
    // CSR mode with the DMMA kernel of p = 3 hexahedra: the K_e scratch is laid out as dof-pair planes (full-sector fragment stores);
    // the wide CSR reduction reads the same layout.  COO mode always keeps the reference's element-major triplet order.
    h->ke_plane_major = (mode == FL_MODE_CSR && h->ndim == 3 && h->use_mma_implicit && h->npe == 64 && h->ng == 64) ? (h->use_mma_implicit == 3 ? 1 : 2) : 0;
    mark(h, 0, st);
    rc = launch_implicit_elements(h, Eulerx, Eulerp, mat, formulation_number, requires_geometry_update ? 1 : 0, ke, h->te, st);
    if (rc) return rc;
    mark(h, 1, st);
    rc = scatter_stiffness(h, nvar, mode, ke, I, J, V, st);
    h->ke_plane_major = 0;
    if (rc) return rc;
    mark(h, 2, st);
    rc = launch_gather_nodes(h, nvar, h->te, T, st);
    mark(h, 3, st);
    return rc;
}

int fl_assemble_laplacian(fl_handle* h, const double* e_tensor_host, int is_hessian_symmetric, int mode, int32_t* I, int32_t* J, double* V,
                          void* stream) {
    if (!h || !e_tensor_host || !V) { set_error("null argument"); return FL_ERR_INVALID; }
    h->ke_plane_major = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == FL_MODE_CSR && !h->pat.nbr_ptr) { set_error("CSR assembly requires fl_pattern_build first"); return FL_ERR_STATE; }
    // the d x d tensor rides in the traction scratch buffer
    int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * 16);
    if (rc) return rc;
    FL_CUDA_CHECK(cudaMemcpyAsync(h->te, e_tensor_host, sizeof(double) * h->ndim * h->ndim, cudaMemcpyHostToDevice, st));
    double* ke = V;
    if (mode == FL_MODE_CSR) {
        rc = ensure_scratch(&h->ke, &h->ke_bytes, sizeof(double) * h->nelem * h->npe * h->npe);
        if (rc) return rc;
        ke = h->ke;
    }
    mark(h, 0, st);
    rc = launch_laplacian_elements(h, h->te, is_hessian_symmetric ? 1 : 0, ke, st);
    if (rc) return rc;
    mark(h, 1, st);
    rc = scatter_stiffness(h, 1, mode, ke, I, J, V, st);
    mark(h, 2, st);
    mark(h, 3, st);
    return rc;
}

int fl_assemble_mass(fl_handle* h, double rho, int nvar, int mass_type, int mode, double* mass, int32_t* I, int32_t* J, double* V,
                     void* stream) {
    if (!h || nvar < 1 || nvar > 4) { set_error("bad argument"); return FL_ERR_INVALID; }
    h->ke_plane_major = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ndof = (size_t)h->npe * nvar;
    if (mass_type == 0) {
        if (!mass) { set_error("null mass"); return FL_ERR_INVALID; }
        int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * h->nelem * ndof);
        if (rc) return rc;
        rc = launch_mass_elements(h, rho, nvar, 1, h->te, st);
        if (rc) return rc;
        return launch_gather_nodes(h, nvar, h->te, mass, st);
    }
    if (!V) { set_error("null V"); return FL_ERR_INVALID; }
    if (mode == FL_MODE_CSR && !h->pat.nbr_ptr) { set_error("CSR assembly requires fl_pattern_build first"); return FL_ERR_STATE; }
    double* ke = V;
    if (mode == FL_MODE_CSR) {
        int rc = ensure_scratch(&h->ke, &h->ke_bytes, sizeof(double) * h->nelem * ndof * ndof);
        if (rc) return rc;
        ke = h->ke;
    }
    int rc = launch_mass_elements(h, rho, nvar, 0, ke, st);
    if (rc) return rc;
    return scatter_stiffness(h, nvar, mode, ke, I, J, V, st);
}

int fl_explicit_update(fl_handle* h, const fl_update_args* a, void* stream) {
    if (!h || !a || !a->M || !a->T || !a->U0 || !a->U00 || !a->Eulerx) { set_error("null argument"); return FL_ERR_INVALID; }
    if (a->use_element_forces && !h->te) { set_error("fl_explicit_forces has not been called"); return FL_ERR_STATE; }
    if ((a->iface_slot == nullptr) != (a->T_iface == nullptr)) { set_error("iface_slot and T_iface go together"); return FL_ERR_INVALID; }
    return launch_explicit_update(h, a->use_element_forces ? (a->write_T ? 2 : 1) : 0, h->te, a->dt, a->fext_scale, a->incd_scale, a->M, a->fext,
                                  a->fixed_mask, a->inc_dirichlet, a->T, a->iface_slot, a->T_iface, a->U0, a->U00, a->Eulerx,
                                  a->status_dev ? a->status_dev : h->flag, a->growth_keys_dev, (cudaStream_t)stream);
}

int fl_explicit_check(fl_handle* h, int64_t* growth_keys_dev, int64_t increment, int32_t* status_dev, void* stream) {
    if (!h || !growth_keys_dev || !status_dev) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_growth_check(growth_keys_dev, increment, status_dev, (cudaStream_t)stream);
}

int fl_explicit_forces(fl_handle* h, const double* Eulerx, const fl_material* mat, int64_t e0, int64_t e1, void* stream) {
    if (!h || !Eulerx || !mat) { set_error("null argument"); return FL_ERR_INVALID; }
    if (e0 < 0 || e1 > h->nelem || e0 > e1) { set_error("bad element range [%lld, %lld)", (long long)e0, (long long)e1); return FL_ERR_INVALID; }
    int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * h->nelem * h->npe * h->ndim);
    if (rc) return rc;
    h->el0 = e0; h->el1 = e1;
    rc = launch_explicit_elements(h, Eulerx, nullptr, mat, 0, h->te, (cudaStream_t)stream);
    h->el0 = 0; h->el1 = -1;
    return rc;
}

int fl_gather_pack_nodes(fl_handle* h, int nvar, const int32_t* node_ids, int64_t n, double* buf, void* stream) {
    if (n == 0) return FL_OK;
    if (!h || !node_ids || !buf || !h->te) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_gather_pack(h, nvar, h->te, node_ids, n, buf, (cudaStream_t)stream);
}

int fl_gather_nodes(fl_handle* h, int nvar, double* T, void* stream) {
    if (!h || !T || !h->te) { set_error("null argument"); return FL_ERR_INVALID; }
    return launch_gather_nodes(h, nvar, h->te, T, (cudaStream_t)stream);
}

int fl_explicit_steps(fl_handle* h, const fl_material* mat, const fl_explicit_ctrl* c, const double* M, const double* fext,
                      const uint8_t* fixed_mask, const double* inc_dirichlet, double* U0, double* U00, double* Eulerx, double* T,
                      int32_t* status_host, void* stream) {
    if (!h || !mat || !c || !M || !U0 || !U00 || !Eulerx || !T) { set_error("null argument"); return FL_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nvar = h->ndim;
    int rc = ensure_scratch(&h->te, &h->te_bytes, sizeof(double) * h->nelem * h->npe * nvar);
    if (rc) return rc;
    FL_CUDA_CHECK(cudaMemsetAsync(h->flag, 0, 2 * sizeof(int32_t), st));
    for (int64_t s = 0; s < c->nsteps; ++s) {
        const double fs = c->fext_scale0 + (double)(c->increment + s) * c->fext_scale_step;
        const double ds = c->incd_scale0 + (double)(c->increment + s) * c->incd_scale_step;
        // first step consumes the caller's T; later steps reduce the per-element tractions and update in one kernel
        rc = launch_explicit_update(h, s == 0 ? 0 : 1, h->te, c->dt, fs, ds, M, fext, fixed_mask, inc_dirichlet, T, nullptr, nullptr, U0, U00,
                                    Eulerx, h->flag, h->growth, st);
        if (rc) return rc;
        rc = launch_growth_check(h->growth, c->increment + s, h->flag, st);
        if (rc) return rc;
        rc = launch_explicit_elements(h, Eulerx, nullptr, mat, 0, h->te, st);
        if (rc) return rc;
    }
    if (c->nsteps > 0) {
        rc = launch_gather_nodes(h, nvar, h->te, T, st);
        if (rc) return rc;
        if (h->contact.surf) {   // the returned T is the reference's TractionForces: internal + contact at the final geometry
            rc = launch_contact(h, Eulerx, T, 1, st);
            if (rc) return rc;
        }
    }
    if (status_host) {
        FL_CUDA_CHECK(cudaMemcpyAsync(status_host, h->flag, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FL_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return FL_OK;
}

__global__ void pack_nodes_kernel(const double* __restrict__ T, const int32_t* __restrict__ ids, int64_t n, int nvar, double* __restrict__ buf) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * nvar) return;
    const int64_t k = i / nvar;
    buf[i] = T[(int64_t)ids[k] * nvar + (i - k * nvar)];
}
__global__ void unpack_add_nodes_kernel(double* __restrict__ T, const int32_t* __restrict__ ids, int64_t n, int nvar,
                                        const double* __restrict__ buf) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * nvar) return;
    const int64_t k = i / nvar;
    T[(int64_t)ids[k] * nvar + (i - k * nvar)] += buf[i];
}

// out[k] = sum_j all[idx[j]..idx[j]+nvar) for j in [ptr[k], ptr[k+1]), added in list order (the lists are sorted by owner rank)
__global__ void sum_ordered_kernel(const double* __restrict__ all, const int64_t* __restrict__ ptr, const int64_t* __restrict__ idx, int64_t n,
                                   int nvar, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * nvar) return;
    const int64_t k = i / nvar;
    const int c = (int)(i - k * nvar);
    double acc = 0.0;
    for (int64_t j = ptr[k]; j < ptr[k + 1]; ++j) acc = __dadd_rn(acc, all[idx[j] + c]);
    out[i] = acc;
}
__global__ void scatter_nodes_kernel(double* __restrict__ T, const int32_t* __restrict__ ids, int64_t n, int nvar, const double* __restrict__ buf) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * nvar) return;
    const int64_t k = i / nvar;
    T[(int64_t)ids[k] * nvar + (i - k * nvar)] = buf[i];
}

int fl_sum_ordered(const double* all, const int64_t* ptr, const int64_t* idx, int64_t n, int nvar, double* out, void* stream) {
    if (n == 0) return FL_OK;
    if (!all || !ptr || !idx || !out) { set_error("null argument"); return FL_ERR_INVALID; }
    sum_ordered_kernel<<<(unsigned)((n * nvar + 255) / 256), 256, 0, (cudaStream_t)stream>>>(all, ptr, idx, n, nvar, out);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int fl_scatter_nodes(double* T, const int32_t* node_ids, int64_t n, int nvar, const double* buf, void* stream) {
    if (n == 0) return FL_OK;
    if (!T || !node_ids || !buf) { set_error("null argument"); return FL_ERR_INVALID; }
    scatter_nodes_kernel<<<(unsigned)((n * nvar + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, node_ids, n, nvar, buf);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int fl_pack_nodes(const double* T, const int32_t* node_ids, int64_t n, int nvar, double* buf, void* stream) {
    if (n == 0) return FL_OK;
    if (!T || !node_ids || !buf) { set_error("null argument"); return FL_ERR_INVALID; }
    pack_nodes_kernel<<<(unsigned)((n * nvar + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, node_ids, n, nvar, buf);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int fl_unpack_add_nodes(double* T, const int32_t* node_ids, int64_t n, int nvar, const double* buf, void* stream) {
    if (n == 0) return FL_OK;
    if (!T || !node_ids || !buf) { set_error("null argument"); return FL_ERR_INVALID; }
    unpack_add_nodes_kernel<<<(unsigned)((n * nvar + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, node_ids, n, nvar, buf);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int fl_row_block_build(fl_handle* h, int nvar, const int32_t* owned_nodes, int64_t n_owned, int64_t* indptr_block, int64_t* nnz_block_host,
                       void* stream) {
    if (!h || nvar < 1 || nvar > 4 || n_owned < 0 || (n_owned > 0 && !owned_nodes) || !indptr_block) { set_error("bad argument"); return FL_ERR_INVALID; }
    return launch_row_block_build(h, nvar, owned_nodes, n_owned, indptr_block, nnz_block_host, (cudaStream_t)stream);
}

int fl_row_block_emit(fl_handle* h, int nvar, const double* V, const int32_t* owned_nodes, int64_t n_owned, const int64_t* node_map,
                      const int64_t* indptr_block, int64_t* cols_global, double* vals, void* stream) {
    if (!h || nvar < 1 || nvar > 4 || !indptr_block || (n_owned > 0 && !owned_nodes) || (vals && !V) || (cols_global && !node_map) ||
        (!vals && !cols_global)) {
        set_error("bad argument");
        return FL_ERR_INVALID;
    }
    return launch_row_block_emit(h, nvar, V, owned_nodes, n_owned, node_map, indptr_block, cols_global, vals, (cudaStream_t)stream);
}

int fl_sfc_order(const double* points, const uint64_t* elements, int64_t nelem, int nodeperelem, int ndim, int64_t nnode, int64_t* perm,
                 void* stream) {
    if (!points || (nelem > 0 && (!elements || !perm)) || nodeperelem < 1 || (ndim != 2 && ndim != 3) || nnode < 1) {
        set_error("bad argument");
        return FL_ERR_INVALID;
    }
    return launch_sfc_order(points, reinterpret_cast<const int64_t*>(elements), nelem, nodeperelem, ndim, nnode, perm, (cudaStream_t)stream);
}

int fl_measure_fp64_peak(int use_dmma, int iters, double* tflops_host) {
    if (!tflops_host || iters < 1) { set_error("bad argument"); return FL_ERR_INVALID; }
    int dev = 0, sms = 0;
    FL_CUDA_CHECK(cudaGetDevice(&dev));
    FL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    FL_CUDA_CHECK(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    FL_CUDA_CHECK(cudaEventCreate(&e0));
    FL_CUDA_CHECK(cudaEventCreate(&e1));
    const int threads = 256, blocks = sms * 8;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FL_CUDA_CHECK(cudaEventRecord(e0));
        if (use_dmma) dmma_peak_kernel<<<blocks, threads>>>(iters, out);
        else dfma_peak_kernel<<<blocks, threads>>>(iters, out);
        FL_CUDA_CHECK(cudaEventRecord(e1));
        FL_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        FL_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = use_dmma ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0 : (double)blocks * threads * iters * 16.0 * 2.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops_host = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    return FL_OK;
}

}  // extern "C"
