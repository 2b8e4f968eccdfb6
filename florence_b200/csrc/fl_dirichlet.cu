// fl_dirichlet.cu -- Dirichlet reduction of the assembled system on the device.
//
// Device replacement of
//   BoundaryCondition.GetReducedMatrices               (Florence/BoundaryCondition/BoundaryCondition.py:842-858)
//   BoundaryCondition.ApplyDirichletGetReducedMatrices (Florence/BoundaryCondition/BoundaryCondition.py:861-891)
// which slice K[columns_in,:][:,columns_in] with scipy on the host in every Newton iteration (FEMSolver.py:951) and fold the
// prescribed dofs into the right hand side, F[in] -= K[in, out(nnz)] * AppliedDirichlet(nnz) * LoadFactor.
//
// The full matrix never leaves the node-level pattern of fl_pattern.cu: a dof row (n,i) has the columns {nbr m, l < nvar}
// ascending, so the reduced pattern is a stream compaction of every free row and the reduced values are a pure copy
// (bit-identical to scipy's fancy indexing).  The Dirichlet term reproduces scipy's csr_matvec summation: ascending
// columns, separate multiply and add, the sum scaled by LoadFactor afterwards; columns with np.isclose(applied, 0) are
// skipped exactly as the reference does (:873).
#include <cub/cub.cuh>

#include "fl_internal.cuh"

namespace fl {

__global__ void mark_out_kernel(const int32_t* __restrict__ cols_out, int64_t n_out, int64_t N, int32_t* __restrict__ is_in,
                                int32_t* __restrict__ bad) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    const int64_t c = cols_out[k];
    if (c < 0 || c >= N || (k > 0 && cols_out[k - 1] >= c)) { *bad = 1; return; }
    is_in[c] = 0;
}

__global__ void fill_i32_kernel(int32_t* v, int64_t n, int32_t x) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}

// new_id[c] = reduced index of a free dof; -(k+1) for the k-th prescribed dof
__global__ void new_id_kernel(const int32_t* __restrict__ is_in, const int32_t* __restrict__ scan, int64_t N,
                              const int32_t* __restrict__ cols_out, int64_t n_out, int32_t* __restrict__ new_id,
                              int32_t* __restrict__ cols_in) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N && is_in[i]) { new_id[i] = scan[i]; cols_in[scan[i]] = (int32_t)i; }
    if (i < n_out) new_id[cols_out[i]] = -(int32_t)(i + 1);
}

// cmask[q] bit l = 1 <=> dof l of neighbour node nbr_idx[q] is free: the row kernels then read one contiguous byte run per node
// instead of chasing nbr_idx -> new_id for every column
__global__ void col_mask_kernel(const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ new_id, int64_t nnzb, int nvar,
                                uint8_t* __restrict__ cmask) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nnzb) return;
    const int64_t m = nbr_idx[q];
    unsigned bits = 0;
    for (int l = 0; l < nvar; ++l) bits |= (new_id[m * nvar + l] >= 0 ? 1u : 0u) << l;
    cmask[q] = (uint8_t)bits;
}

// One warp per NODE of the full matrix: the nvar dof rows of a node share their column list, so the free/prescribed split of
// the columns (new_id lookups, ballots, positions) is computed once and applied to every free row of the node.
// MODE 0: count the free columns of each free row (cnt_b[rr]).  MODE 1: write the reduced column indices.
// MODE 2: copy values / fold the prescribed columns into F / gather F_b.
template <int MODE, int NV>
__global__ void __launch_bounds__(256)
dirichlet_rows_kernel(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ new_id,
                      const uint8_t* __restrict__ cmask, int64_t nnode, const int64_t* __restrict__ rowptr_b, int64_t* __restrict__ cnt_b,
                      int32_t* __restrict__ indices_b, const double* __restrict__ V, double* __restrict__ V_b,
                      const double* __restrict__ applied, double load_factor, double* __restrict__ F, double* __restrict__ F_b) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = warp0; n < nnode; n += nwarp) {
        int rr[NV];
        bool any = false;
#pragma unroll
        for (int i = 0; i < NV; ++i) { rr[i] = new_id[n * NV + i]; any |= rr[i] >= 0; }
        if (!any) continue;
        const int64_t p0 = nbr_ptr[n];
        const int cnt = (int)(nbr_ptr[n + 1] - p0);
        // MODE 2 with neither values to copy nor prescribed dofs to fold in only gathers F_b: no row walk
        const int w = (MODE != 2 || V_b || applied) ? cnt * NV : 0;
        int64_t ob[NV];
        double s[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) { ob[i] = (MODE && rr[i] >= 0) ? rowptr_b[rr[i]] : 0; s[i] = 0.0; }
        int running = 0;
        for (int t0 = 0; t0 < w; t0 += 32) {
            const int t = t0 + lane;
            const bool valid = t < w;
            // free / prescribed from the per-pair mask (contiguous per node); the index itself is only needed for the reduced
            // column numbers (MODE 1) and for prescribed columns with an applied value (MODE 2)
            bool in = false;
            int id = -1;
            if (valid) {
                const int r = t / NV, l = t - r * NV;
                in = (cmask[p0 + r] >> l) & 1;
                if (MODE == 1 ? in : (MODE == 2 && !in && applied)) id = new_id[(int64_t)nbr_idx[p0 + r] * NV + l];
            }
            const unsigned m = __ballot_sync(0xffffffffu, in);
            const int pos = running + __popc(m & ((1u << lane) - 1u));
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (in && rr[i] >= 0) indices_b[ob[i] + pos] = id;
            }
            if (MODE == 2) {
                double v[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    v[i] = (valid && rr[i] >= 0) ? V[(int64_t)NV * (NV * p0 + (int64_t)i * cnt) + t] : 0.0;
                if (V_b) {
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        if (in && rr[i] >= 0) V_b[ob[i] + pos] = v[i];
                }
                if (applied) {
                    bool use = false;
                    double a = 0.0;
                    if (valid && !in) {
                        a = applied[-id - 1];
                        use = !(fabs(a) <= 1e-8);          // ~np.isclose(a, 0.0)
                    }
                    unsigned mo = __ballot_sync(0xffffffffu, use);
                    if (mo) {
                        double prod[NV];
#pragma unroll
                        for (int i = 0; i < NV; ++i) prod[i] = __dmul_rn(v[i], a);
                        while (mo) {                         // ascending column order, one add per entry (scipy csr_matvec)
                            const int l = __ffs(mo) - 1;
#pragma unroll
                            for (int i = 0; i < NV; ++i) s[i] = __dadd_rn(s[i], __shfl_sync(0xffffffffu, prod[i], l));
                            mo &= mo - 1;
                        }
                    }
                }
            }
            running += __popc(m);
        }
        if (lane < NV) {
            int r_l = rr[0];
            double s_l = s[0];
#pragma unroll
            for (int i = 1; i < NV; ++i)
                if (lane == i) { r_l = rr[i]; s_l = s[i]; }
            if (r_l >= 0) {
                if (MODE == 0) cnt_b[r_l] = running;
                if (MODE == 2 && F) {
                    double f = F[n * NV + lane];
                    if (applied) {
                        f = __dadd_rn(f, -__dmul_rn(s_l, load_factor));
                        F[n * NV + lane] = f;
                    }
                    if (F_b) F_b[r_l] = f;
                }
            }
        }
    }
}

template <int MODE>
static void launch_rows(const fl_handle* h, const Dirichlet& d, int64_t* cnt_b, int32_t* indices_b, const double* V, double* V_b,
                        const double* applied, double load_factor, double* F, double* F_b, cudaStream_t st) {
    const int64_t need = (h->nnode + 7) / 8;
    const int64_t cap = (int64_t)h->sm_count * 32;
    const unsigned grid = (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
#define FL_ROWS(NV_)                                                                                                              \
    dirichlet_rows_kernel<MODE, NV_><<<grid, 256, 0, st>>>(h->pat.nbr_ptr, h->pat.nbr_idx, d.new_id, d.cmask, h->nnode, d.rowptr_b, cnt_b, \
                                                           indices_b, V, V_b, applied, load_factor, F, F_b)
    switch (d.nvar) {
        case 1: FL_ROWS(1); break;
        case 2: FL_ROWS(2); break;
        case 3: FL_ROWS(3); break;
        default: FL_ROWS(4); break;
    }
#undef FL_ROWS
}

__global__ void narrow_ptr_kernel(const int64_t* __restrict__ p, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)p[i];
}

void dirichlet_free(fl_handle* h) {
    Dirichlet& d = h->dir;
    cudaFree(d.new_id); cudaFree(d.cols_in); cudaFree(d.rowptr_b); cudaFree(d.cmask);
    d = Dirichlet();
}

namespace {
// temporary device allocation released on every exit path
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() { cudaFree(p); }
    template <typename T> T* as() { return static_cast<T*>(p); }
};
}  // namespace

int dirichlet_build(fl_handle* h, int nvar, const int32_t* cols_out, int64_t n_out) {
    if (!h->pat.nbr_ptr) { set_error("fl_dirichlet_build requires fl_pattern_build first"); return FL_ERR_STATE; }
    dirichlet_free(h);
    Dirichlet& d = h->dir;
    const int64_t N = h->nnode * nvar;
    if (N >= ((int64_t)1 << 31)) { set_error("dof count exceeds int32 indexing"); return FL_ERR_INVALID; }
    if (n_out < 0 || n_out > N) { set_error("bad number of prescribed dofs"); return FL_ERR_INVALID; }
    DevTmp is_in_b, scan_b, bad_b, cnt_b_b, tmp_b;
    size_t tmp_bytes = 0, tmp2 = 0;
    const unsigned gN = (unsigned)((N + 255) / 256 > 0 ? (N + 255) / 256 : 1);
    FL_CUDA_CHECK(cudaMalloc(&is_in_b.p, sizeof(int32_t) * (N + 1)));
    FL_CUDA_CHECK(cudaMalloc(&scan_b.p, sizeof(int32_t) * (N + 1)));
    FL_CUDA_CHECK(cudaMalloc(&bad_b.p, sizeof(int32_t)));
    int32_t *is_in = is_in_b.as<int32_t>(), *scan = scan_b.as<int32_t>(), *bad = bad_b.as<int32_t>();
    FL_CUDA_CHECK(cudaMemset(bad, 0, sizeof(int32_t)));
    fill_i32_kernel<<<gN, 256>>>(is_in, N, 1);
    if (n_out) mark_out_kernel<<<(unsigned)((n_out + 255) / 256), 256>>>(cols_out, n_out, N, is_in, bad);
    int32_t bad_h = 0;
    FL_CUDA_CHECK(cudaMemcpy(&bad_h, bad, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (bad_h) {
        set_error("columns_out must be strictly ascending and within [0, nvar*nnode)");
        return FL_ERR_INVALID;
    }
    d.n_in = N - n_out;
    d.n_out = n_out;
    d.nvar = nvar;
    FL_CUDA_CHECK(cudaMalloc(&d.new_id, sizeof(int32_t) * (N > 0 ? N : 1)));
    FL_CUDA_CHECK(cudaMalloc(&d.cols_in, sizeof(int32_t) * (d.n_in > 0 ? d.n_in : 1)));
    FL_CUDA_CHECK(cudaMalloc(&d.rowptr_b, sizeof(int64_t) * (d.n_in + 1)));
    FL_CUDA_CHECK(cudaMalloc(&cnt_b_b.p, sizeof(int64_t) * (d.n_in + 1)));
    int64_t* cnt_b = cnt_b_b.as<int64_t>();
    FL_CUDA_CHECK(cudaMemset(cnt_b, 0, sizeof(int64_t) * (d.n_in + 1)));
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, is_in, scan, (int)N);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, cnt_b, d.rowptr_b, (int)(d.n_in + 1));
    if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    FL_CUDA_CHECK(cudaMalloc(&tmp_b.p, tmp_bytes > 0 ? tmp_bytes : 1));
    void* tmp = tmp_b.p;
    if (N) {
        FL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_in, scan, (int)N));
        const int64_t m = N > n_out ? N : n_out;
        new_id_kernel<<<(unsigned)((m + 255) / 256), 256>>>(is_in, scan, N, cols_out, n_out, d.new_id, d.cols_in);
        FL_CUDA_CHECK(cudaMalloc(&d.cmask, h->pat.nnzb > 0 ? h->pat.nnzb : 1));
        col_mask_kernel<<<(unsigned)((h->pat.nnzb + 255) / 256), 256>>>(h->pat.nbr_idx, d.new_id, h->pat.nnzb, nvar, d.cmask);
        launch_rows<0>(h, d, cnt_b, nullptr, nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, 0);
    }
    FL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt_b, d.rowptr_b, (int)(d.n_in + 1)));
    FL_CUDA_CHECK(cudaMemcpy(&d.nnz_b, d.rowptr_b + d.n_in, sizeof(int64_t), cudaMemcpyDeviceToHost));
    FL_CUDA_CHECK(cudaGetLastError());
    if (d.nnz_b >= ((int64_t)1 << 31)) { set_error("reduced nnz exceeds int32 indexing"); return FL_ERR_INVALID; }
    return FL_OK;
}

int launch_dirichlet_export(fl_handle* h, int32_t* indptr_b, int32_t* indices_b, int32_t* columns_in, cudaStream_t st) {
    Dirichlet& d = h->dir;
    if (!d.new_id) { set_error("fl_dirichlet_build has not been called"); return FL_ERR_STATE; }
    const int64_t N = h->nnode * d.nvar;
    if (indptr_b) narrow_ptr_kernel<<<(unsigned)((d.n_in + 256) / 256), 256, 0, st>>>(d.rowptr_b, d.n_in + 1, indptr_b);
    if (indices_b && N)
        launch_rows<1>(h, d, nullptr, indices_b, nullptr, nullptr, nullptr, 0.0, nullptr, nullptr, st);
    if (columns_in && d.n_in)
        FL_CUDA_CHECK(cudaMemcpyAsync(columns_in, d.cols_in, sizeof(int32_t) * d.n_in, cudaMemcpyDeviceToDevice, st));
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int launch_dirichlet_apply(fl_handle* h, const double* V, double* V_b, const double* applied, double load_factor, double* F, double* F_b,
                           cudaStream_t st) {
    Dirichlet& d = h->dir;
    if (!d.new_id) { set_error("fl_dirichlet_build has not been called"); return FL_ERR_STATE; }
    if ((V_b || applied) && !V) { set_error("V is required to reduce values or apply prescribed dofs"); return FL_ERR_INVALID; }
    if ((applied || F_b) && !F) { set_error("F is required"); return FL_ERR_INVALID; }
    const int64_t N = h->nnode * d.nvar;
    if (N == 0) return FL_OK;
    launch_rows<2>(h, d, nullptr, nullptr, V, V_b, applied, load_factor, F, F_b, st);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
