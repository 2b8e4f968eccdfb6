// fl_pattern.cu -- node adjacency, CSR sparsity pattern, slot maps and the deterministic global reductions.
//
// Device replacement of
//   ComputeSparsityPattern / _ComputeSparsityPattern_ / _ComputeDataIndices_
//       (Florence/FiniteElements/Assembly/_Assembly_/ComputeSparsityPattern.pyx:44-112, .h:22-62, :67-122)
//   fill_triplet, SparseAssemblyNativeCSR_   (Florence/VariationalPrinciple/_Mass_/_MassIntegrand_.h:69-110,
//       Florence/FiniteElements/Assembly/_Assembly_/SparseAssemblyNative.h:32-45)
//   RHSAssemblyNative_                        (…/_Assembly_/RHSAssemblyNative.pyx:30-39)
//
// The pattern is kept at NODE level (row node -> sorted unique neighbour nodes): every dof row of a node has the same
// column nodes, so the (nvar*nnode)^2 CSR pattern and the element->CSR slot map follow from npe^2 uint16 ranks per
// element instead of the reference's 2*ndof^2 int32 -- 36x less map traffic for tet10/nvar=3.
// The CSR value reduction is node-centric: one warp owns the nvar rows of a node, walks the node's elements in ascending
// order (the reference's summation order) and adds the matching K_e rows into a shared-memory row buffer.  No atomics,
// bit-reproducible, and every global access is a coalesced row.
#include <cub/cub.cuh>

#include "fl_internal.cuh"

namespace fl {

// Set-up temporaries: released on every exit path (FL_CUDA_CHECK returns early on error)
struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { release(); return cudaMalloc(&p, bytes > 0 ? bytes : 8); }
    void release() { if (p) cudaFree(p); p = nullptr; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

__global__ void iota_kernel(int32_t* v, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// ptr[n] = first position in sorted `keys` whose value >= n  (n = 0..nrow)
template <typename K>
__global__ void lower_bound_kernel(const K* __restrict__ keys, int64_t nkeys, K scale, int64_t nrow, int64_t* __restrict__ ptr) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n > nrow) return;
    const K target = (K)n * scale;
    int64_t lo = 0, hi = nkeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    ptr[n] = lo;
}

// rank rows reordered by visit (adjacency position): the reduction of node n then reads one contiguous run of uint16
__global__ void reorder_rank_kernel(const uint16_t* __restrict__ rank, const int32_t* __restrict__ adj_idx, int64_t nvisit, int npe,
                                    uint16_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nvisit * npe) return;
    const int64_t k = i / npe;
    out[i] = rank[(int64_t)adj_idx[k] * npe + (i - k * npe)];
}

__global__ void max_diff_kernel(const int64_t* __restrict__ ptr, int64_t n, int* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicMax(out, (int)(ptr[i + 1] - ptr[i]));
}

int build_adjacency(fl_handle* h) {
    const int64_t nk = h->nelem * h->npe;
    if (nk >= (int64_t)1 << 31) {
        set_error("nelem*nodeperelem = %lld exceeds int32 indexing", (long long)nk);
        return FL_ERR_INVALID;
    }
    FL_CUDA_CHECK(cudaMalloc(&h->adj_ptr, sizeof(int64_t) * (h->nnode + 1)));
    FL_CUDA_CHECK(cudaMalloc(&h->adj_idx, sizeof(int32_t) * (nk > 0 ? nk : 1)));
    if (nk == 0) {
        FL_CUDA_CHECK(cudaMemset(h->adj_ptr, 0, sizeof(int64_t) * (h->nnode + 1)));
        return FL_OK;
    }
    DevBuf keys_out, vals_in, tmp, dmax;
    size_t tmp_bytes = 0;
    FL_CUDA_CHECK(keys_out.alloc(sizeof(int32_t) * nk));
    FL_CUDA_CHECK(vals_in.alloc(sizeof(int32_t) * nk));
    iota_kernel<<<(unsigned)((nk + 255) / 256), 256>>>(vals_in.as<int32_t>(), nk);
    int end_bit = 1;
    while (((int64_t)1 << end_bit) < h->nnode && end_bit < 32) ++end_bit;
    // stable LSD radix sort: equal node ids keep ascending flat index, i.e. ascending element number
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->conn, keys_out.as<int32_t>(), vals_in.as<int32_t>(), h->adj_idx, (int)nk, 0, end_bit);
    FL_CUDA_CHECK(tmp.alloc(tmp_bytes));
    FL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, h->conn, keys_out.as<int32_t>(), vals_in.as<int32_t>(), h->adj_idx, (int)nk, 0,
                                                  end_bit));
    lower_bound_kernel<int32_t><<<(unsigned)((h->nnode + 256) / 256), 256>>>(keys_out.as<int32_t>(), nk, 1, h->nnode, h->adj_ptr);
    FL_CUDA_CHECK(dmax.alloc(sizeof(int)));
    FL_CUDA_CHECK(cudaMemset(dmax.p, 0, sizeof(int)));
    max_diff_kernel<<<(unsigned)((h->nnode + 255) / 256), 256>>>(h->adj_ptr, h->nnode, dmax.as<int>());
    FL_CUDA_CHECK(cudaMemcpy(&h->max_adj, dmax.p, sizeof(int), cudaMemcpyDeviceToHost));
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// ------------------------------------------------------------------------------------------------ pattern
__global__ void pair_keys_kernel(const int32_t* __restrict__ conn, int64_t nelem, int npe, int64_t nnode, int64_t* __restrict__ keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = nelem * npe * npe;
    if (i >= total) return;
    const int64_t e = i / (npe * npe);
    const int r = (int)(i - e * npe * npe);
    const int a = r / npe, b = r - a * npe;
    keys[i] = (int64_t)conn[e * npe + a] * nnode + conn[e * npe + b];
}

__global__ void split_keys_kernel(const int64_t* __restrict__ keys, int64_t n, int64_t nnode, int32_t* __restrict__ col) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) col[i] = (int32_t)(keys[i] % nnode);
}

__global__ void rank_kernel(const int32_t* __restrict__ conn, int64_t nelem, int npe, const int64_t* __restrict__ nbr_ptr,
                            const int32_t* __restrict__ nbr_idx, uint16_t* __restrict__ rank) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = nelem * npe * npe;
    if (i >= total) return;
    const int64_t e = i / (npe * npe);
    const int r = (int)(i - e * npe * npe);
    const int a = r / npe, b = r - a * npe;
    const int32_t n = conn[e * npe + a], m = conn[e * npe + b];
    int64_t lo = nbr_ptr[n], hi = nbr_ptr[n + 1];
    const int64_t base = lo;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (nbr_idx[mid] < m) lo = mid + 1; else hi = mid;
    }
    rank[i] = (uint16_t)(lo - base);
}

int pattern_build(fl_handle* h) {
    Pattern& p = h->pat;
    if (p.nbr_ptr) return FL_OK;
    const int npe = h->npe;
    const int64_t total = h->nelem * npe * npe;
    if (total == 0) {
        set_error("empty mesh has no sparsity pattern");
        return FL_ERR_INVALID;
    }
    // Memory: two int64 key arrays of nelem*npe^2 entries + the radix sort's temporary.  Every mesh whose CSR values and K_e
    // scratch fit the device also fits this (the keys are 16 of the 8*nvar^2 + ... bytes per node pair the assembly itself needs).
    DevBuf keys, keys_sorted, d_num, tmp, dmax;
    size_t tb1 = 0, tb2 = 0;
    FL_CUDA_CHECK(keys.alloc(sizeof(int64_t) * total));
    FL_CUDA_CHECK(keys_sorted.alloc(sizeof(int64_t) * total));
    pair_keys_kernel<<<(unsigned)((total + 255) / 256), 256>>>(h->conn, h->nelem, npe, h->nnode, keys.as<int64_t>());
    int end_bit = 1;
    while (end_bit < 63 && ((int64_t)1 << end_bit) < h->nnode * h->nnode) ++end_bit;
    cub::DeviceRadixSort::SortKeys(nullptr, tb1, keys.as<int64_t>(), keys_sorted.as<int64_t>(), total, 0, end_bit);
    FL_CUDA_CHECK(tmp.alloc(tb1));
    FL_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(tmp.p, tb1, keys.as<int64_t>(), keys_sorted.as<int64_t>(), total, 0, end_bit));
    tmp.release();
    int64_t* uniq = keys.as<int64_t>();  // reuse
    FL_CUDA_CHECK(d_num.alloc(sizeof(int64_t)));
    cub::DeviceSelect::Unique(nullptr, tb2, keys_sorted.as<int64_t>(), uniq, d_num.as<int64_t>(), total);
    FL_CUDA_CHECK(tmp.alloc(tb2));
    FL_CUDA_CHECK(cub::DeviceSelect::Unique(tmp.p, tb2, keys_sorted.as<int64_t>(), uniq, d_num.as<int64_t>(), total));
    FL_CUDA_CHECK(cudaMemcpy(&p.nnzb, d_num.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
    tmp.release(); d_num.release(); keys_sorted.release();
    auto fail = [&](int code) {   // leave the handle without a half-built pattern
        cudaFree(p.nbr_ptr); cudaFree(p.nbr_idx); cudaFree(p.rank); cudaFree(p.rank_adj);
        p = Pattern();
        return code;
    };
#define FL_PTRY(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                            \
            return fail(FL_ERR_CUDA);                                                             \
        }                                                                                         \
    } while (0)
    FL_PTRY(cudaMalloc(&p.nbr_ptr, sizeof(int64_t) * (h->nnode + 1)));
    FL_PTRY(cudaMalloc(&p.nbr_idx, sizeof(int32_t) * p.nnzb));
    lower_bound_kernel<int64_t><<<(unsigned)((h->nnode + 256) / 256), 256>>>(uniq, p.nnzb, h->nnode, h->nnode, p.nbr_ptr);
    split_keys_kernel<<<(unsigned)((p.nnzb + 255) / 256), 256>>>(uniq, p.nnzb, h->nnode, p.nbr_idx);
    FL_PTRY(cudaDeviceSynchronize());
    keys.release();
    FL_PTRY(dmax.alloc(sizeof(int)));
    FL_PTRY(cudaMemset(dmax.p, 0, sizeof(int)));
    max_diff_kernel<<<(unsigned)((h->nnode + 255) / 256), 256>>>(p.nbr_ptr, h->nnode, dmax.as<int>());
    FL_PTRY(cudaMemcpy(&p.max_cnt, dmax.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (p.max_cnt >= 65536) {
        set_error("a node with %d neighbours exceeds the uint16 rank map", p.max_cnt);
        return fail(FL_ERR_UNSUPPORTED);
    }
    FL_PTRY(cudaMalloc(&p.rank, sizeof(uint16_t) * total));
    rank_kernel<<<(unsigned)((total + 255) / 256), 256>>>(h->conn, h->nelem, npe, p.nbr_ptr, p.nbr_idx, p.rank);
    FL_PTRY(cudaMalloc(&p.rank_adj, sizeof(uint16_t) * (total > 0 ? total : 1)));
    reorder_rank_kernel<<<(unsigned)((total + 255) / 256), 256>>>(p.rank, h->adj_idx, h->nelem * npe, npe, p.rank_adj);
    FL_PTRY(cudaGetLastError());
    FL_PTRY(cudaDeviceSynchronize());
#undef FL_PTRY
    return FL_OK;
}

// indptr (nvar*nnode+1) and indices: ComputeSparsityPattern.h:48-60 ordering
__global__ void export_indptr_kernel(const int64_t* __restrict__ nbr_ptr, int64_t nnode, int nvar, int32_t* __restrict__ indptr) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n > nnode) return;
    if (n == nnode) {
        indptr[nvar * nnode] = (int32_t)(nbr_ptr[nnode] * nvar * nvar);
        return;
    }
    const int64_t base = nbr_ptr[n] * nvar * nvar, w = (nbr_ptr[n + 1] - nbr_ptr[n]) * nvar;
    for (int i = 0; i < nvar; ++i) indptr[nvar * n + i] = (int32_t)(base + i * w);
}

__global__ void export_indices_kernel(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr_idx, int64_t nnode,
                                      int64_t nnzb, int nvar, int32_t* __restrict__ indices) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= nnzb) return;
    // row node of pair p: last n with nbr_ptr[n] <= p
    int64_t lo = 0, hi = nnode;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (nbr_ptr[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const int64_t n = lo, k = p - nbr_ptr[n], w = (nbr_ptr[n + 1] - nbr_ptr[n]) * nvar;
    const int64_t base = nbr_ptr[n] * nvar * nvar;
    const int32_t m = nbr_idx[p];
    for (int i = 0; i < nvar; ++i)
        for (int l = 0; l < nvar; ++l) indices[base + i * w + k * nvar + l] = nvar * m + l;
}

int launch_pattern_export(fl_handle* h, int nvar, int32_t* indptr, int32_t* indices, cudaStream_t st) {
    const Pattern& p = h->pat;
    if (!p.nbr_ptr) { set_error("fl_pattern_build has not been called"); return FL_ERR_STATE; }
    if (p.nnzb * nvar * nvar >= (int64_t)1 << 31) {
        set_error("nnz = %lld does not fit the reference's int32 indptr", (long long)(p.nnzb * nvar * nvar));
        return FL_ERR_INVALID;
    }
    export_indptr_kernel<<<(unsigned)((h->nnode + 256) / 256), 256, 0, st>>>(p.nbr_ptr, h->nnode, nvar, indptr);
    export_indices_kernel<<<(unsigned)((p.nnzb + 255) / 256), 256, 0, st>>>(p.nbr_ptr, p.nbr_idx, h->nnode, p.nnzb, nvar, indices);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// data_local_indices / data_global_indices in the reference's per-element node-sorted order (ComputeSparsityPattern.h:84-118)
__global__ void data_indices_kernel(const int32_t* __restrict__ conn, int64_t nelem, int npe, int nvar, const int64_t* __restrict__ nbr_ptr,
                                    const uint16_t* __restrict__ rank, int32_t* __restrict__ dl, int32_t* __restrict__ dg) {
    // one block per element; thread 0 argsorts the element's nodes in shared memory
    extern __shared__ int sh[];
    int* srt = sh;  // npe
    const int64_t e = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int a = 0; a < npe; ++a) srt[a] = a;
        for (int a = 1; a < npe; ++a) {
            const int v = srt[a];
            int q = a - 1;
            while (q >= 0 && conn[e * npe + srt[q]] > conn[e * npe + v]) { srt[q + 1] = srt[q]; --q; }
            srt[q + 1] = v;
        }
    }
    __syncthreads();
    const int ndof = nvar * npe;
    const int64_t cap = (int64_t)ndof * ndof;
    for (int64_t t = threadIdx.x; t < cap; t += blockDim.x) {
        const int i = (int)(t / ndof), j = (int)(t - (int64_t)i * ndof);
        const int ca = i / nvar, ii = i - ca * nvar, cb = j / nvar, jj = j - cb * nvar;
        const int a = srt[ca], b = srt[cb];
        const int64_t n = conn[e * npe + a];
        const int64_t w = (nbr_ptr[n + 1] - nbr_ptr[n]) * nvar;
        const int64_t row0 = nbr_ptr[n] * nvar * nvar + ii * w;
        dg[e * cap + t] = (int32_t)(row0 + (int64_t)rank[(e * npe + a) * npe + b] * nvar + jj);
        dl[e * cap + t] = (a * nvar + ii) * ndof + (b * nvar + jj);
    }
}

int launch_data_indices(fl_handle* h, int nvar, int32_t* dl, int32_t* dg, cudaStream_t st) {
    const Pattern& p = h->pat;
    if (!p.nbr_ptr) { set_error("fl_pattern_build has not been called"); return FL_ERR_STATE; }
    if (h->nelem == 0) return FL_OK;
    data_indices_kernel<<<(unsigned)h->nelem, 128, sizeof(int) * h->npe, st>>>(h->conn, h->nelem, h->npe, nvar, p.nbr_ptr, p.rank, dl, dg);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// COO row/column indices: fill_triplet (_MassIntegrand_.h:86-107)
__global__ void coo_indices_kernel(const int32_t* __restrict__ conn, int64_t nelem, int npe, int nvar, int32_t* __restrict__ I,
                                   int32_t* __restrict__ J) {
    const int ndof = npe * nvar;
    const int64_t cap = (int64_t)ndof * ndof;
    const int64_t total = nelem * cap;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = k / cap;
        const int rc = (int)(k - e * cap);
        const int r = rc / ndof, c = rc - r * ndof;
        I[k] = nvar * conn[e * npe + r / nvar] + r % nvar;
        J[k] = nvar * conn[e * npe + c / nvar] + c % nvar;
    }
}

int launch_coo_indices(fl_handle* h, int nvar, int32_t* I, int32_t* J, cudaStream_t st) {
    const int64_t total = h->nelem * (int64_t)(h->npe * nvar) * (h->npe * nvar);
    if (total == 0) return FL_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)h->sm_count * 64) blocks = (int64_t)h->sm_count * 64;
    coo_indices_kernel<<<(unsigned)blocks, 256, 0, st>>>(h->conn, h->nelem, h->npe, nvar, I, J);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// V rows of node n = sum over its elements (ascending) of the K_e rows of n, placed by node rank (SparseAssemblyNativeCSR_).
// The NV rows (a,0..NV-1) of K_e are adjacent, so one (element, node) visit is a single contiguous run of NV*ndof doubles.
// Two visits are loaded per trip (independent loads in flight) and added in element order.
template <int NV, int NPE>
__global__ void __launch_bounds__(256)
csr_gather_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx, const int64_t* __restrict__ nbr_ptr,
                  const uint16_t* __restrict__ rank_adj, const double* __restrict__ ke, int64_t nnode, int npe_rt, int wmax,
                  double* __restrict__ V) {
    extern __shared__ double rowbuf[];
    const int npe = NPE ? NPE : npe_rt;
    const int ndof = npe * NV;
    const int run = NV * ndof;           // doubles per (element, node) visit
    constexpr int MAXT = 4;              // register tile: up to 4*32 = 128 doubles per visit, else the generic loop
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    double* buf = rowbuf + (size_t)warp * NV * wmax;
    // per-warp stage of the rank rows of up to 32 visits (behind the row buffers of all warps)
    uint16_t* rks = reinterpret_cast<uint16_t*>(rowbuf + (size_t)wpb * NV * wmax) + (size_t)warp * 32 * npe;
    for (int64_t n = blockIdx.x * (int64_t)wpb + warp; n < nnode; n += (int64_t)gridDim.x * wpb) {
        const int w = (int)(nbr_ptr[n + 1] - nbr_ptr[n]) * NV;
        for (int t = lane; t < NV * w; t += 32) buf[t] = 0.0;
        const int64_t k0 = adj_ptr[n], k1 = adj_ptr[n + 1];
        for (int64_t kc = k0; kc < k1; kc += 32) {
            // One dependent global round trip per 32 visits instead of two per pair of visits: every lane fetches the flat
            // connectivity index of one visit, and the visits' rank rows (contiguous in adjacency order) are staged in shared
            // memory.  ncu had 60 % of this kernel's stall samples on the integer ops consuming the per-visit rank loads.
            const int nv = (int)min((int64_t)32, k1 - kc);
            const int32_t fl = lane < nv ? adj_idx[kc + lane] : 0;
            __syncwarp();
            for (int t = lane; t < nv * npe; t += 32) rks[t] = rank_adj[kc * npe + t];
            __syncwarp();
            if (run <= MAXT * 32) {
                // destination offsets depend only on (t, rank): i*w + rank[b]*NV + l
                for (int v = 0; v < nv; v += 2) {
                    const bool two = (v + 1 < nv);
                    const int64_t f0 = __shfl_sync(0xffffffffu, fl, v), f1 = __shfl_sync(0xffffffffu, fl, two ? v + 1 : v);
                    const int64_t ea = f0 / npe, eb = f1 / npe;
                    const int aa = (int)(f0 - ea * npe), ab = (int)(f1 - eb * npe);
                    const double* ra = ke + ea * (int64_t)ndof * ndof + (int64_t)(aa * NV) * ndof;
                    const double* rb = ke + eb * (int64_t)ndof * ndof + (int64_t)(ab * NV) * ndof;
                    const uint16_t* qa = rks + v * npe;
                    const uint16_t* qb = rks + (two ? v + 1 : v) * npe;
                    double va[MAXT], vb[MAXT];
                    int da[MAXT], db[MAXT];
#pragma unroll
                    for (int u = 0; u < MAXT; ++u) {
                        const int t = lane + 32 * u;
                        if (t < run) {
                            const int i = t / ndof, c = t - i * ndof;
                            const int b = c / NV, l = c - b * NV;
                            va[u] = ra[t];
                            da[u] = i * w + (int)qa[b] * NV + l;
                            if (two) {
                                vb[u] = rb[t];
                                db[u] = i * w + (int)qb[b] * NV + l;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < MAXT; ++u)
                        if (lane + 32 * u < run) buf[da[u]] += va[u];
                    __syncwarp();
                    if (two) {
#pragma unroll
                        for (int u = 0; u < MAXT; ++u)
                            if (lane + 32 * u < run) buf[db[u]] += vb[u];
                        __syncwarp();
                    }
                }
            } else {
                for (int v = 0; v < nv; ++v) {
                    const int64_t flat = __shfl_sync(0xffffffffu, fl, v);
                    const int64_t e = flat / npe;
                    const int a = (int)(flat - e * npe);
                    const uint16_t* rk = rks + v * npe;
                    const double* krow = ke + e * (int64_t)ndof * ndof + (int64_t)(a * NV) * ndof;
                    for (int t = lane; t < run; t += 32) {
                        const int i = t / ndof, c = t - i * ndof;
                        const int b = c / NV, l = c - b * NV;
                        buf[i * w + (int)rk[b] * NV + l] += krow[t];
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        const int64_t base = nbr_ptr[n] * NV * NV;
        for (int t = lane; t < NV * w; t += 32) V[base + t] = buf[t];
        __syncwarp();
    }
}

// Plane stride of the component-plane row buffer [(i,l)][rank]: the write-out reads rank r of plane l for consecutive CSR columns
// c = r*NV + l, so the NV planes must start 16/NV banks (of 8 bytes) apart -- with the plain stride (64, 112, 343 neighbours ...)
// the common hex64 node classes read with a 4-way bank conflict.
template <int NV>
__host__ __device__ __forceinline__ int plane_stride(int cn) {
    constexpr int TARGET = NV == 4 ? 4 : (NV == 3 ? 5 : (NV == 2 ? 8 : 0));
    return NV == 1 ? cn : cn + ((TARGET - cn) & 15);
}

// Wide rows (high-order elements: nvar*ndof > 128 doubles per visit): one block per node, the threads split each visit's run.
template <int NV, int NPE_T>
__global__ void __launch_bounds__(256)
csr_gather_wide_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx, const int64_t* __restrict__ nbr_ptr,
                       const uint16_t* __restrict__ rank_adj, const double* __restrict__ ke, int64_t nnode, int npe_rt, int plane_major,
                       int wmax, int max_adj, double* __restrict__ V) {
    extern __shared__ double rowbuf[];
    const int npe = NPE_T ? NPE_T : npe_rt;   // compile-time for hex27 / hex64: the index arithmetic below becomes shifts and constants
    const int ndof = npe * NV;
    const int run = NV * ndof;
    // register tile: the next visit's values are loaded before the current visit is added
    constexpr int MAXR = NPE_T ? (NV * NV * NPE_T + 255) / 256 : 4;
    // behind the row buffer: flat connectivity index and rank row of every visit of the node (one global round trip per node
    // instead of two dependent ones per visit)
    int32_t* flats = reinterpret_cast<int32_t*>(rowbuf + (size_t)NV * wmax);
    uint16_t* rks = reinterpret_cast<uint16_t*>(flats + ((max_adj + 1) & ~1));
    for (int64_t n = blockIdx.x; n < nnode; n += gridDim.x) {
        const int w = (int)(nbr_ptr[n + 1] - nbr_ptr[n]) * NV;
        int cnp = plane_stride<NV>(w / NV);           // plane-major modes: stride of a component plane of the row buffer
        if (cnp * NV > wmax) cnp = w / NV;            // (the widest nodes when the padded stride would cost a block per SM)
        const int64_t k0 = adj_ptr[n];
        const int nvis = (int)(adj_ptr[n + 1] - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < (plane_major ? NV * NV * cnp : NV * w); t += blockDim.x) rowbuf[t] = 0.0;
        for (int t = threadIdx.x; t < nvis; t += blockDim.x) flats[t] = adj_idx[k0 + t];
        for (int t = threadIdx.x; t < nvis * npe; t += blockDim.x) rks[t] = rank_adj[k0 * npe + t];
        __syncthreads();
        // source offset (relative to ke) and destination slot of item t of visit v
        auto src_of = [&](int v, int t) -> int64_t {
            const int64_t flat = flats[v];
            const int64_t e = flat / npe;
            const int a = (int)(flat - e * npe);
            if (plane_major == 2) {
                // per-row-node planes [a][(i,l)][b]: the whole visit is one contiguous run
                return e * (int64_t)ndof * ndof + (int64_t)a * (NV * NV) * npe + t;
            }
            if (plane_major) {
                // K_e scratch stored as dof-pair planes [(i,l)][a][b] (DMMA element kernels): npe-long contiguous runs
                const int il = t / npe, b = t - il * npe;
                return e * (int64_t)ndof * ndof + ((int64_t)il * npe + a) * npe + b;
            }
            return e * (int64_t)ndof * ndof + (int64_t)(a * NV) * ndof + t;
        };
        auto dst_of = [&](int v, int t) -> int {
            const uint16_t* rk = rks + v * npe;
            if (plane_major) {
                // consecutive lanes hold consecutive column nodes b of one (i,l) plane: the row buffer is kept as component
                // planes [(i,l)][rank] so that they land in distinct banks (rank*NV + l would be an NV*2-way conflict)
                const int il = t / npe, b = t - il * npe;
                return il * cnp + (int)rk[b];
            }
            const int i = t / ndof, c = t - i * ndof;
            const int b = c / NV, l = c - b * NV;
            return i * w + (int)rk[b] * NV + l;
        };
        if (NPE_T == 64 && NV == 4 && plane_major && blockDim.x == 256) {
            // hex64 / nvar 4, dof-pair planes: warp w owns planes 2w and 2w+1 outright (128 items per visit = 4 per lane), so the
            // destinations of different warps never meet and the visits are ordered by __syncwarp alone -- no block barrier per visit
            // (ncu: 3.0 barrier stalls per issue with the items dealt round-robin over the block)
            const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
            const int cn = cnp;
            int il[4], bb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { il[u] = 2 * wp + (u >> 1); bb[u] = ln + 32 * (u & 1); }
            auto src = [&](int v, int u) -> int64_t {
                const int64_t flat = flats[v];
                const int64_t e = flat >> 6;
                const int a = (int)(flat & 63);
                return plane_major == 2 ? e * (int64_t)(256 * 256) + ((int64_t)a * 16 + il[u]) * 64 + bb[u]
                                        : e * (int64_t)(256 * 256) + ((int64_t)il[u] * 64 + a) * 64 + bb[u];
            };
            double nxt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) nxt[u] = nvis > 0 ? ke[src(0, u)] : 0.0;
            for (int v = 0; v < nvis; ++v) {
                double cur[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
                if (v + 1 < nvis) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) nxt[u] = ke[src(v + 1, u)];
                }
                const uint16_t* rk = rks + v * 64;
#pragma unroll
                for (int u = 0; u < 4; ++u) rowbuf[il[u] * cn + (int)rk[bb[u]]] += cur[u];
                __syncwarp();
            }
            __syncthreads();
        } else if (run <= MAXR * (int)blockDim.x) {
            double nxt[MAXR];
#pragma unroll
            for (int u = 0; u < MAXR; ++u) {
                const int t = threadIdx.x + u * blockDim.x;
                nxt[u] = (nvis > 0 && t < run) ? ke[src_of(0, t)] : 0.0;
            }
            for (int v = 0; v < nvis; ++v) {
                double cur[MAXR];
#pragma unroll
                for (int u = 0; u < MAXR; ++u) cur[u] = nxt[u];
                if (v + 1 < nvis) {
#pragma unroll
                    for (int u = 0; u < MAXR; ++u) {
                        const int t = threadIdx.x + u * blockDim.x;
                        if (t < run) nxt[u] = ke[src_of(v + 1, t)];
                    }
                }
#pragma unroll
                for (int u = 0; u < MAXR; ++u) {
                    const int t = threadIdx.x + u * blockDim.x;
                    if (t < run) rowbuf[dst_of(v, t)] += cur[u];
                }
                __syncthreads();
            }
        } else {
            for (int v = 0; v < nvis; ++v) {
                for (int t = threadIdx.x; t < run; t += blockDim.x) rowbuf[dst_of(v, t)] += ke[src_of(v, t)];
                __syncthreads();
            }
        }
        const int64_t base = nbr_ptr[n] * NV * NV;
        if (plane_major) {
            // row i of the node, CSR column c = r*NV + l  <-  plane (i,l), rank r.  No division by a run-time number: this loop was
            // 43 % of the kernel's instructions when it decoded a flat index with `/ w` (ncu source page, prof_r2_hi)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double* Vi = V + base + (int64_t)i * w;
                const double* rb = rowbuf + (i * NV) * cnp;
                for (int c = threadIdx.x; c < w; c += blockDim.x) {
                    const int r = c / NV, l = c - r * NV;
                    Vi[c] = rb[l * cnp + r];
                }
            }
        } else {
            for (int t = threadIdx.x; t < NV * w; t += blockDim.x) V[base + t] = rowbuf[t];
        }
    }
}

// hex64 with nvar 4 and a plane-major K_e scratch (config 4): the block-per-node reduction above, software-pipelined ACROSS nodes.
// In the kernel above a node costs three dependent global round trips (row pointers -> connectivity index and rank rows of its
// visits -> K_e values) with five blocks per SM to hide them: 2.9 ms for 24^3 elements, and the time did not move when 43 % of the
// instructions were removed (division-free write-out), i.e. it is bound by that latency chain.  Here the row pointers of the node
// after next travel to shared memory by cp.async, the visit lists of the next node wait in three registers per thread while the
// current node is reduced, and the first visit of the next node is loaded before the current node's rows are written out.
// A warp owns its component planes of the row buffer outright (nvar 4: planes 2w and 2w+1, 4 items per lane and visit): zeroing and the
// visits of a node need __syncwarp only; two block barriers per node (row buffer complete / row buffer free).
// Same sums in the same order as csr_gather_wide_kernel: bit-identical V.  Needs max_adj <= 8 (two rank entries per thread).
// nvar 4 (config 4): 8 warps x 2 planes; nvar 3 (mechanics on hex64): 9 warps x 1 plane.
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int NV>
struct hex64_gather_cfg {
    static constexpr int PPW = NV == 4 ? 2 : 1;                 // component planes per warp
    static constexpr int NW = (NV * NV + PPW - 1) / PPW;        // nvar 4: 8 warps x 2 planes, nvar 3: 9 warps x 1 plane
    static constexpr int THREADS = NW * 32, IPL = 2 * PPW;      // items per lane and visit (a plane row is 64 doubles = 2 per lane)
    static constexpr int MAXV = 8;                              // visits whose rank rows fit two entries per thread
    static constexpr size_t META = 12 * sizeof(int64_t) + MAXV * sizeof(int32_t) + MAXV * 64 * sizeof(uint16_t);
};

template <int NV>
__global__ void __launch_bounds__(hex64_gather_cfg<NV>::THREADS)
csr_gather_hex64_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx, const int64_t* __restrict__ nbr_ptr,
                        const uint16_t* __restrict__ rank_adj, const double* __restrict__ ke, int64_t nnode, int plane_major, int wmax,
                        double* __restrict__ V) {
    using C = hex64_gather_cfg<NV>;
    constexpr int NPE = 64, MAXV = C::MAXV, PPW = C::PPW, IPL = C::IPL, NT = C::THREADS, NDOF = NV * NPE;
    extern __shared__ double rowbuf[];                                          // [(i,l)][plane stride]
    int64_t* ptrs = reinterpret_cast<int64_t*>(rowbuf + (size_t)NV * wmax);    // [3][adj_ptr n, n+1, nbr_ptr n, n+1]
    int32_t* flats = reinterpret_cast<int32_t*>(ptrs + 12);                     // [MAXV]      flat connectivity index per visit
    uint16_t* rks = reinterpret_cast<uint16_t*>(flats + MAXV);                 // [MAXV][64]  rank rows per visit
    uint32_t* rks2 = reinterpret_cast<uint32_t*>(rks);
    const int tid = threadIdx.x, wp = tid >> 5, ln = tid & 31;
    const int64_t stride = gridDim.x;
    auto issue_ptrs = [&](int64_t n, int slot) {
        if (tid < 4 && n < nnode) cp_async_8(&ptrs[slot * 4 + tid], tid < 2 ? adj_ptr + n + tid : nbr_ptr + n + (tid - 2));
    };
    int il[IPL], bb[IPL];
#pragma unroll
    for (int u = 0; u < IPL; ++u) { il[u] = PPW * wp + (u >> 1); bb[u] = ln + 32 * (u & 1); }
    auto src = [&](int32_t flat, int u) -> int64_t {
        const int64_t e = flat >> 6;
        const int a = flat & 63;
        return plane_major == 2 ? e * (int64_t)(NDOF * NDOF) + ((int64_t)a * (NV * NV) + il[u]) * 64 + bb[u]
                                : e * (int64_t)(NDOF * NDOF) + ((int64_t)il[u] * 64 + a) * 64 + bb[u];
    };
    int64_t n = blockIdx.x;
    issue_ptrs(n, 0);
    issue_ptrs(n + stride, 1);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    double nxt[IPL];
#pragma unroll
    for (int u = 0; u < IPL; ++u) nxt[u] = 0.0;
    if (n < nnode) {
        const int64_t k0 = ptrs[0];
        const int nvis = (int)(ptrs[1] - k0);
        if (tid < nvis) flats[tid] = adj_idx[k0 + tid];
        if (2 * tid < nvis * NPE) rks2[tid] = *reinterpret_cast<const uint32_t*>(rank_adj + k0 * NPE + 2 * tid);
        if (nvis > 0) {
            const int32_t f = adj_idx[k0];
#pragma unroll
            for (int u = 0; u < IPL; ++u) nxt[u] = ke[src(f, u)];
        }
    }
    __syncthreads();
    for (int j = 0; n < nnode; ++j, n += stride) {
        const int slot = j % 3, slot1 = (j + 1) % 3;
        issue_ptrs(n + 2 * stride, (j + 2) % 3);
        // the next node's visit lists: plain loads into registers, first used after this node's visits
        uint32_t rkn = 0;
        int32_t fln = 0, f0n = 0;
        int nvis1 = 0;
        if (n + stride < nnode) {
            const int64_t k1 = ptrs[slot1 * 4];
            nvis1 = (int)(ptrs[slot1 * 4 + 1] - k1);
            if (tid < nvis1) fln = adj_idx[k1 + tid];
            if (2 * tid < nvis1 * NPE) rkn = *reinterpret_cast<const uint32_t*>(rank_adj + k1 * NPE + 2 * tid);
            if (nvis1 > 0) f0n = adj_idx[k1];
        }
        const int64_t k0 = ptrs[slot * 4], nb0 = ptrs[slot * 4 + 2];
        const int nvis = (int)(ptrs[slot * 4 + 1] - k0);
        const int cn = (int)(ptrs[slot * 4 + 3] - nb0);
        const int w = cn * NV;
        int cnp = plane_stride<NV>(cn);
        if (cnp * NV > wmax) cnp = cn;                      // the widest nodes when the padded stride would cost a block per SM
        double* mine = rowbuf + PPW * wp * cnp;             // this warp's component planes
        for (int t = ln; t < PPW * cnp; t += 32) mine[t] = 0.0;
        __syncwarp();
        for (int v = 0; v < nvis; ++v) {
            double cur[IPL];
#pragma unroll
            for (int u = 0; u < IPL; ++u) cur[u] = nxt[u];
            if (v + 1 < nvis) {
                const int32_t f = flats[v + 1];
#pragma unroll
                for (int u = 0; u < IPL; ++u) nxt[u] = ke[src(f, u)];
            }
            const uint16_t* rk = rks + v * NPE;
#pragma unroll
            for (int u = 0; u < IPL; ++u) rowbuf[il[u] * cnp + (int)rk[bb[u]]] += cur[u];
            __syncwarp();
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                                    // row buffer complete, visit lists no longer read, row pointers landed
        if (tid < nvis1) flats[tid] = fln;
        if (2 * tid < nvis1 * NPE) rks2[tid] = rkn;
        if (nvis1 > 0) {
#pragma unroll
            for (int u = 0; u < IPL; ++u) nxt[u] = ke[src(f0n, u)];
        }
        const int64_t base = nb0 * NV * NV;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double* Vi = V + base + (int64_t)i * w;
            const double* rb = rowbuf + (i * NV) * cnp;
            for (int c = tid; c < w; c += NT) {
                const int r = c / NV, l = c - r * NV;
                Vi[c] = rb[l * cnp + r];
            }
        }
        __syncthreads();                                    // row buffer free, next visit lists visible
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

template <int NV, int NPE_T>
static int launch_csr_gather_wide(fl_handle* h, const double* ke, double* V, cudaStream_t st) {
    const Pattern& p = h->pat;
    // row buffer of the widest node: NV rows of max_cnt*NV entries, or NV*NV component planes of the padded stride
    int wmax = (h->ke_plane_major ? plane_stride<NV>(p.max_cnt) : p.max_cnt) * NV;
    size_t smem = sizeof(double) * NV * (size_t)wmax + sizeof(int32_t) * ((h->max_adj + 1) & ~1) +
                        ((sizeof(uint16_t) * (size_t)h->max_adj * h->npe + 7) & ~(size_t)7);
    if (smem > (size_t)h->max_smem_optin) {
        set_error("CSR rows of %d entries do not fit shared memory", p.max_cnt * NV);
        return FL_ERR_UNSUPPORTED;
    }
    if constexpr ((NV == 3 || NV == 4) && NPE_T == 64) {
        using C = hex64_gather_cfg<NV>;
        constexpr size_t meta = C::META;
        auto k64 = csr_gather_hex64_kernel<NV>;
        if (h->ke_plane_major && !h->wide_unpipelined && h->max_adj <= C::MAXV &&
            sizeof(double) * NV * (size_t)wmax + meta <= (size_t)h->max_smem_optin) {
            auto occ_of = [&](int wm, int* occ) {
                return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k64, C::THREADS, sizeof(double) * NV * (size_t)wm + meta);
            };
            // the conflict-free plane stride for every node if that costs no block per SM, else for all but the widest nodes
            int wm = wmax, occ_pad = 0, occ_raw = 0;
            FL_CUDA_CHECK(cudaFuncSetAttribute(k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * NV * (size_t)wmax + meta)));
            FL_CUDA_CHECK(occ_of(wmax, &occ_pad));
            FL_CUDA_CHECK(occ_of(p.max_cnt * NV, &occ_raw));
            if (occ_raw > occ_pad) wm = p.max_cnt * NV;
            const int occ2 = occ_raw > occ_pad ? occ_raw : occ_pad;
            if (occ2 >= 1) {
                // persistent blocks: each walks nodes b, b + grid, ... so that its prefetches always have a next node
                int64_t blocks = (int64_t)h->sm_count * occ2;
                if (blocks > h->nnode) blocks = h->nnode;
                k64<<<(unsigned)blocks, C::THREADS, sizeof(double) * NV * (size_t)wm + meta, st>>>(
                    h->adj_ptr, h->adj_idx, p.nbr_ptr, p.rank_adj, ke, h->nnode, h->ke_plane_major, wm, V);
                FL_CUDA_CHECK(cudaGetLastError());
                return FL_OK;
            }
        }
    }
    auto kern = csr_gather_wide_kernel<NV, NPE_T>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (h->ke_plane_major) {   // padded plane stride only where it does not cost a block per SM
        const size_t raw = smem - sizeof(double) * NV * (size_t)(wmax - p.max_cnt * NV);
        int o_pad = 0, o_raw = 0;
        FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o_pad, kern, 256, smem));
        FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o_raw, kern, 256, raw));
        if (o_raw > o_pad) { wmax = p.max_cnt * NV; smem = raw; }
    }
    // a visit is NV*NV*npe doubles: 256 threads for the hexahedral shapes (243 .. 1024 per visit); elements with shorter visits
    // (tet20: 180 / 320) run more, smaller blocks per SM
    // (measured on 82 944 tet20, profiles/tet20_bench.py: nvar 3 2.53 -> 1.69 ms with 64 threads, nvar 4 3.23 -> 2.72 ms with 128;
    // the register prefetch of the kernel covers 4 items per thread, so the block must keep 4 * threads >= visit length)
    int threads = 256;
    if (NPE_T == 0) {
        const int run = NV * NV * h->npe;
        threads = ((run + 3) / 4 + 63) / 64 * 64;
        if (threads > 256) threads = 256;
    }
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = h->nnode;
    const int64_t cap = (int64_t)h->sm_count * occ * 8;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, threads, smem, st>>>(h->adj_ptr, h->adj_idx, p.nbr_ptr, p.rank_adj, ke, h->nnode, h->npe, h->ke_plane_major, wmax,
                                              h->max_adj, V);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

template <int NV, int NPE>
static int launch_csr_gather_T(fl_handle* h, const double* ke, double* V, cudaStream_t st) {
    const Pattern& p = h->pat;
    const int wmax = p.max_cnt * NV;
    const size_t per_warp = sizeof(double) * NV * wmax + ((sizeof(uint16_t) * 32 * h->npe + 7) & ~(size_t)7);   // row buffer + rank stage
    int wpb = 8;
    while (wpb > 1 && per_warp * wpb > 48 * 1024) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > (size_t)h->max_smem_optin) {
        set_error("CSR row of %d entries does not fit shared memory", wmax);
        return FL_ERR_UNSUPPORTED;
    }
    auto kern = csr_gather_kernel<NV, NPE>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = (h->nnode + wpb - 1) / wpb;
    const int64_t cap = (int64_t)h->sm_count * occ * 4;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, wpb * 32, smem, st>>>(h->adj_ptr, h->adj_idx, p.nbr_ptr, p.rank_adj, ke, h->nnode, h->npe, wmax, V);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

template <int NV>
static int launch_csr_gather_NV(fl_handle* h, const double* ke, double* V, cudaStream_t st) {
    // visits longer than 128 doubles: one block per node (tet20 measured both ways, profiles/tet20_bench.py: the warp-per-node
    // row-buffer kernel below takes 3.12 ms (nvar 3) / 9.5 ms (nvar 4) where the block-per-node kernel takes 1.69 / 2.72 ms)
    if (NV * NV * h->npe > 128) {
        if (h->npe == 64) return launch_csr_gather_wide<NV, 64>(h, ke, V, st);
        if (h->npe == 27) return launch_csr_gather_wide<NV, 27>(h, ke, V, st);
        return launch_csr_gather_wide<NV, 0>(h, ke, V, st);
    }
    switch (h->npe) {
        case 4: return launch_csr_gather_T<NV, 4>(h, ke, V, st);
        case 8: return launch_csr_gather_T<NV, 8>(h, ke, V, st);
        case 10: return launch_csr_gather_T<NV, 10>(h, ke, V, st);
        case 27: return launch_csr_gather_T<NV, 27>(h, ke, V, st);
        // Poisson on p = 3 / p = 4 hexahedra (config 1): a visit is one row of 64 / 125 doubles; compile-time node counts turn the
        // index decoding (64-bit `flat / npe`, `t / ndof`) into multiplies
        case 64: if constexpr (NV == 1) return launch_csr_gather_T<NV, 64>(h, ke, V, st); else break;
        case 125: if constexpr (NV == 1) return launch_csr_gather_T<NV, 125>(h, ke, V, st); else break;
        default: break;
    }
    return launch_csr_gather_T<NV, 0>(h, ke, V, st);
}

int launch_csr_gather(fl_handle* h, int nvar, const double* ke, double* V, cudaStream_t st) {
    if (!h->pat.nbr_ptr) { set_error("fl_pattern_build has not been called"); return FL_ERR_STATE; }
    if (reg_gather_preferred(h, nvar)) return launch_csr_gather_reg(h, nvar, ke, V, st);
    switch (nvar) {
        case 1: return launch_csr_gather_NV<1>(h, ke, V, st);
        case 2: return launch_csr_gather_NV<2>(h, ke, V, st);
        case 3: return launch_csr_gather_NV<3>(h, ke, V, st);
        case 4: return launch_csr_gather_NV<4>(h, ke, V, st);
        default: set_error("nvar=%d unsupported", nvar); return FL_ERR_INVALID;
    }
}

// ------------------------------------------------------------------------------------------------ owned CSR row blocks
// "Each rank emits the CSR row block it owns": the rows of the owned nodes of a locally assembled matrix, compacted, with the
// column indices translated from local to GLOBAL dof numbers (Assembly.py:1000-1041 does this on the parent with the per-partition
// `partitioned_maps`).  The nvar rows of a node are contiguous in V, so the values are a concatenation of contiguous runs.
__global__ void row_block_counts_kernel(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ owned, int64_t n_owned, int nvar,
                                        int64_t* __restrict__ cnt) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n_owned) return;
    cnt[k] = k < n_owned ? (nbr_ptr[owned[k] + 1] - nbr_ptr[owned[k]]) * nvar * nvar : 0;
}

__global__ void row_block_indptr_kernel(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ owned, int64_t n_owned, int nvar,
                                        const int64_t* __restrict__ node_off, int64_t* __restrict__ indptr) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > n_owned) return;
    if (k == n_owned) { indptr[n_owned * nvar] = node_off[n_owned]; return; }
    const int64_t w = (nbr_ptr[owned[k] + 1] - nbr_ptr[owned[k]]) * nvar;
    for (int i = 0; i < nvar; ++i) indptr[k * nvar + i] = node_off[k] + i * w;
}

__global__ void row_block_emit_kernel(const int64_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ owned,
                                      int64_t n_owned, int nvar, const int64_t* __restrict__ node_map, const int64_t* __restrict__ indptr,
                                      const double* __restrict__ V, int64_t* __restrict__ cols, double* __restrict__ vals) {
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_owned) return;
    const int64_t node = owned[k], p0 = nbr_ptr[node];
    const int64_t w = (nbr_ptr[node + 1] - p0) * nvar;
    const int64_t src = p0 * nvar * nvar, dst = indptr[k * nvar];
    for (int64_t t = lane; t < nvar * w; t += 32) {
        const int64_t c = t % w;
        const int64_t q = c / nvar;
        if (vals) vals[dst + t] = V[src + t];
        if (cols) cols[dst + t] = node_map[nbr_idx[p0 + q]] * nvar + (c - q * nvar);
    }
}

int launch_row_block_build(fl_handle* h, int nvar, const int32_t* owned, int64_t n_owned, int64_t* indptr_block, int64_t* nnz_host,
                           cudaStream_t st) {
    const Pattern& p = h->pat;
    if (!p.nbr_ptr) { set_error("fl_pattern_build has not been called"); return FL_ERR_STATE; }
    DevBuf cnt, off, tmp;
    FL_CUDA_CHECK(cnt.alloc(sizeof(int64_t) * (n_owned + 1)));
    FL_CUDA_CHECK(off.alloc(sizeof(int64_t) * (n_owned + 1)));
    const unsigned blocks = (unsigned)((n_owned + 256) / 256);
    row_block_counts_kernel<<<blocks, 256, 0, st>>>(p.nbr_ptr, owned, n_owned, nvar, cnt.as<int64_t>());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<int64_t>(), off.as<int64_t>(), (int)(n_owned + 1), st);
    FL_CUDA_CHECK(tmp.alloc(tb));
    FL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.as<int64_t>(), off.as<int64_t>(), (int)(n_owned + 1), st));
    row_block_indptr_kernel<<<blocks, 256, 0, st>>>(p.nbr_ptr, owned, n_owned, nvar, off.as<int64_t>(), indptr_block);
    FL_CUDA_CHECK(cudaGetLastError());
    if (nnz_host) FL_CUDA_CHECK(cudaMemcpyAsync(nnz_host, off.as<int64_t>() + n_owned, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FL_CUDA_CHECK(cudaStreamSynchronize(st));
    return FL_OK;
}

int launch_row_block_emit(fl_handle* h, int nvar, const double* V, const int32_t* owned, int64_t n_owned, const int64_t* node_map,
                          const int64_t* indptr_block, int64_t* cols, double* vals, cudaStream_t st) {
    const Pattern& p = h->pat;
    if (!p.nbr_ptr) { set_error("fl_pattern_build has not been called"); return FL_ERR_STATE; }
    if (n_owned == 0) return FL_OK;
    const int64_t threads = n_owned * 32;
    row_block_emit_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p.nbr_ptr, p.nbr_idx, owned, n_owned, nvar, node_map, indptr_block, V,
                                                                             cols, vals);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// ------------------------------------------------------------------------------------------------ space-filling-curve element order
// Morton (Z-order) key of the element centroid, 21 bits per axis, relative to the bounding box of the nodes.  Sorting the elements
// by this key before cutting them into contiguous blocks (Mesh.Partition, Mesh.py:7403: np.array_split of the element range) gives
// compact blocks -- small interfaces -- whatever order the mesh file came in.  Every operation is an individually rounded IEEE
// operation so that the host twin (partition.sfc_order on numpy arrays) produces the same keys bit for bit.
__device__ __forceinline__ long long okey(double v) {
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFLL);
}
__device__ __forceinline__ double okey_inv(long long k) { return __longlong_as_double(k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL)); }

__global__ void bbox_kernel(const double* __restrict__ pts, int64_t nnode, int D, long long* __restrict__ box /* [2*D]: lo keys, hi keys */) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int k = 0; k < D; ++k) {
        double lo = n < nnode ? pts[n * D + k] : INFINITY, hi = n < nnode ? pts[n * D + k] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0 && lo <= hi) {
            atomicMin(box + k, okey(lo));
            atomicMax(box + D + k, okey(hi));
        }
    }
}

__device__ __forceinline__ uint64_t spread3(uint64_t x) {   // 21 bits -> every third bit
    x &= 0x1FFFFFULL;
    x = (x | x << 32) & 0x1F00000000FFFFULL;
    x = (x | x << 16) & 0x1F0000FF0000FFULL;
    x = (x | x << 8) & 0x100F00F00F00F00FULL;
    x = (x | x << 4) & 0x10C30C30C30C30C3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

template <typename IdxT>
__global__ void morton_keys_kernel(const double* __restrict__ pts, const IdxT* __restrict__ els, int64_t nelem, int npe, int D,
                                   const long long* __restrict__ box, uint64_t* __restrict__ keys, int64_t* __restrict__ iota) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nelem) return;
    uint64_t key = 0;
    for (int k = 0; k < D; ++k) {
        double c = 0.0;
        for (int a = 0; a < npe; ++a) c = __dadd_rn(c, pts[els[e * npe + a] * D + k]);
        c = __ddiv_rn(c, (double)npe);
        const double lo = okey_inv(box[k]), hi = okey_inv(box[D + k]);
        const double span = __dadd_rn(hi, -lo);
        double u = span > 0.0 ? __ddiv_rn(__dadd_rn(c, -lo), span) : 0.0;
        u = __dmul_rn(u, 2097152.0);
        int64_t q = (int64_t)u;
        q = q < 0 ? 0 : (q > 2097151 ? 2097151 : q);
        key |= spread3((uint64_t)q) << k;
    }
    keys[e] = key;
    iota[e] = e;
}

template <typename IdxT>
static int sfc_order_T(const double* points, const IdxT* elements, int64_t nelem, int npe, int ndim, int64_t nnode, int64_t* perm,
                       cudaStream_t st) {
    if (nelem == 0) return FL_OK;
    DevBuf box, keys, keys_out, iota, tmp;
    FL_CUDA_CHECK(box.alloc(sizeof(long long) * 6));
    long long init[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < ndim; ++k) { init[k] = INT64_MAX; init[ndim + k] = INT64_MIN; }
    FL_CUDA_CHECK(cudaMemcpyAsync(box.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    FL_CUDA_CHECK(cudaStreamSynchronize(st));   // `init` is a stack array
    bbox_kernel<<<(unsigned)((nnode + 255) / 256), 256, 0, st>>>(points, nnode, ndim, box.as<long long>());
    FL_CUDA_CHECK(keys.alloc(sizeof(uint64_t) * nelem));
    FL_CUDA_CHECK(keys_out.alloc(sizeof(uint64_t) * nelem));
    FL_CUDA_CHECK(iota.alloc(sizeof(int64_t) * nelem));
    morton_keys_kernel<IdxT><<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>(points, elements, nelem, npe, ndim, box.as<long long>(),
                                                                              keys.as<uint64_t>(), iota.as<int64_t>());
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<uint64_t>(), keys_out.as<uint64_t>(), iota.as<int64_t>(), perm, nelem, 0, 63, st);
    FL_CUDA_CHECK(tmp.alloc(tb));
    // stable: elements with equal keys keep their original relative order
    FL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<uint64_t>(), keys_out.as<uint64_t>(), iota.as<int64_t>(), perm, nelem, 0, 63, st));
    FL_CUDA_CHECK(cudaStreamSynchronize(st));
    return FL_OK;
}

int launch_sfc_order(const double* points, const int64_t* elements, int64_t nelem, int npe, int ndim, int64_t nnode, int64_t* perm,
                     cudaStream_t st) {
    return sfc_order_T<int64_t>(points, elements, nelem, npe, ndim, nnode, perm, st);
}

int launch_sfc_order_conn(const double* points, const int32_t* conn, int64_t nelem, int npe, int ndim, int64_t nnode, int64_t* perm,
                          cudaStream_t st) {
    return sfc_order_T<int32_t>(points, conn, nelem, npe, ndim, nnode, perm, st);
}

}  // namespace fl
