// fl_gather.cu -- plan and stand-alone launch of the register-resident CSR value reduction (fl_gather.cuh).
//
// Replaces the slot-map scatter of the reference (SparseAssemblyNativeCSR_, SparseAssemblyNative.h:32-45; fill_global_data,
// _MassIntegrand_.h:115-166) for nvar = 2..4 on the low-order elements whose K_e row blocks are multiples of 16 bytes.  The plan is
// a one-off per (mesh, nvar), built on the device from the node-level pattern: it only depends on the sparsity pattern, like the
// reference's data_global_indices.
#include <cub/cub.cuh>

#include <cstdlib>

#include "fl_gather.cuh"

namespace fl {

namespace {
struct Buf {
    void* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes > 0 ? bytes : 8); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace

namespace {
struct Layout {   // run-time twin of gather_cfg
    int nvar, bits, fpw, fwg, maxg;
    __host__ __device__ int groups(int64_t nbr) const { return (int)((nbr + GATHER_SLOTS - 1) / GATHER_SLOTS); }
    __host__ __device__ int items(int64_t nbr) const { return (groups(nbr) + maxg - 1) / maxg; }
    __host__ __device__ int rw(int ng) const { return 2 + fwg * ng; }
    // words of ONE visit's records over all items of a node with nbr neighbours
    __host__ __device__ int64_t words_per_visit(int64_t nbr) const { return 2 * (int64_t)items(nbr) + (int64_t)fwg * groups(nbr); }
    // words per visit of the items before item t: they all span maxg groups
    __host__ __device__ int64_t item_words_before(int t) const { return (int64_t)t * rw(maxg); }
};

}  // namespace

static Layout make_layout(int nvar, int bits) {
    Layout L;
    L.nvar = nvar; L.bits = bits; L.fpw = 32 / bits; L.fwg = GATHER_SLOTS / L.fpw; L.maxg = nvar <= 3 ? 3 : 1;
    return L;
}

__global__ void item_counts_kernel(const int64_t* __restrict__ adj_ptr, const int64_t* __restrict__ nbr_ptr, int64_t nnode, Layout L,
                                   int64_t* __restrict__ nitems, int64_t* __restrict__ nwords) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int64_t nbr = nbr_ptr[n + 1] - nbr_ptr[n];
    nitems[n] = L.items(nbr);
    nwords[n] = L.words_per_visit(nbr) * (adj_ptr[n + 1] - adj_ptr[n]);
}

// keys (optional): the largest stored flat index among the node's visits -- when the element kernel that walks the elements in
// storage order has passed it, every K_e row of the node exists.
__global__ void items_kernel(const int64_t* __restrict__ adj_ptr, const int64_t* __restrict__ nbr_ptr, const int64_t* __restrict__ item_ptr,
                             const int64_t* __restrict__ word_ptr, int64_t nnode, Layout L, GatherItem* __restrict__ items,
                             const int32_t* __restrict__ flat_store, int32_t* __restrict__ keys, int32_t* __restrict__ iota) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    const int64_t nbr = nbr_ptr[n + 1] - nbr_ptr[n], nvis = adj_ptr[n + 1] - adj_ptr[n];
    const int64_t i0 = item_ptr[n];
    const int ni = L.items(nbr), ng_all = L.groups(nbr);
    int32_t key = 0;
    if (keys)
        for (int64_t k = adj_ptr[n]; k < adj_ptr[n + 1]; ++k) key = max(key, flat_store[k]);
    for (int t = 0; t < ni; ++t) {
        GatherItem it;
        it.rec_off = word_ptr[n] + nvis * L.item_words_before(t);
        it.v_base = nbr_ptr[n] * L.nvar * L.nvar + (int64_t)t * L.maxg * GATHER_SLOTS * L.nvar;
        it.nvis = (int32_t)nvis;
        it.w = (int32_t)(nbr * L.nvar);
        it.nslots = (int32_t)min((int64_t)L.maxg * GATHER_SLOTS, nbr - (int64_t)t * L.maxg * GATHER_SLOTS);
        it.ng = min(L.maxg, ng_all - t * L.maxg);
        items[i0 + t] = it;
        if (keys) { keys[i0 + t] = key; iota[i0 + t] = (int32_t)(i0 + t); }
    }
}

__global__ void permute_items_kernel(const GatherItem* __restrict__ in, const int32_t* __restrict__ idx, int64_t n, GatherItem* __restrict__ out) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[idx[k]];
}

// The records start as all ones (every field NONE).  Pass 1, one thread per visit: word 0 = flat index of the visit's K_e row block
// (`flat_store`, optional: where the element kernel of the streamed assembly, which walks the elements in space-filling-curve order,
// stores it; default adj_idx), word 1 = 0 (spare) in the visit's record of every item of the node.
__global__ void record_heads_kernel(const int32_t* __restrict__ conn, const int64_t* __restrict__ adj_ptr, const int64_t* __restrict__ nbr_ptr,
                                    const int32_t* __restrict__ adj_idx, const int32_t* __restrict__ flat_store,
                                    const int64_t* __restrict__ word_ptr, int64_t nvisit, Layout L, uint32_t* __restrict__ recs) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nvisit) return;
    const int32_t flat = adj_idx[k];
    const int64_t n = conn[flat];
    const int64_t v = k - adj_ptr[n], nvis = adj_ptr[n + 1] - adj_ptr[n];
    const int64_t nbr = nbr_ptr[n + 1] - nbr_ptr[n];
    const int ni = L.items(nbr), ng_all = L.groups(nbr);
    const uint32_t stored = (uint32_t)(flat_store ? flat_store[k] : flat);
    for (int t = 0; t < ni; ++t) {
        const int ng = min(L.maxg, ng_all - t * L.maxg);
        uint32_t* rec = recs + word_ptr[n] + nvis * L.item_words_before(t) + v * L.rw(ng);
        rec[0] = stored;
        rec[1] = 0u;
    }
}

// Pass 2, one thread per (visit, local node b): the field of slot rank(a,b) -- clearing a field from NONE to b is a single atomic
// AND.
__global__ void record_fields_kernel(const int32_t* __restrict__ conn, const int64_t* __restrict__ adj_ptr, const int64_t* __restrict__ nbr_ptr,
                                     const int32_t* __restrict__ adj_idx, const uint16_t* __restrict__ rank_adj,
                                     const int64_t* __restrict__ word_ptr, int64_t nvisit, int npe, Layout L, uint32_t* __restrict__ recs) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nvisit * npe) return;
    const int64_t k = t / npe;
    const int b = (int)(t - k * npe);
    const int64_t n = conn[adj_idx[k]];
    const int64_t v = k - adj_ptr[n], nvis = adj_ptr[n + 1] - adj_ptr[n];
    const int64_t nbr = nbr_ptr[n + 1] - nbr_ptr[n];
    const int r = rank_adj[t];
    const int gg = r / GATHER_SLOTS, f = r - gg * GATHER_SLOTS;   // group of the node's neighbour list, slot within the group
    const int it = gg / L.maxg, g = gg - it * L.maxg;
    const int ng = min(L.maxg, L.groups(nbr) - it * L.maxg);
    uint32_t* rec = recs + word_ptr[n] + nvis * L.item_words_before(it) + v * L.rw(ng);
    const uint32_t none = (1u << L.bits) - 1u;
    const int sh = L.bits * (f % L.fpw);
    atomicAnd(&rec[2 + g * L.fwg + f / L.fpw], ~(none << sh) | ((uint32_t)b << sh));
}

constexpr int GW = 8;   // warps per block of the stand-alone kernel

template <int NV, int BITS, int NPE>
__global__ void __launch_bounds__(GW * 32)
csr_gather_reg_kernel(const GatherPlan gp, const double* __restrict__ ke, double* __restrict__ V) {
    using SM = gather_warp_smem<NV, BITS, gather_batch<NV, NPE>::B, NPE>;
    extern __shared__ __align__(16) unsigned char smem_g[];
    SM* sm = reinterpret_cast<SM*>(smem_g) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    gather_warp_loop<NV, BITS, gather_batch<NV, NPE>::B, NPE, false>(gp, wid, nw, ke, V, lane, *sm, nullptr, 0, nullptr);
}

static int bits_for(int nvar, int npe) {
    if (nvar == 2) return npe <= 15 ? 4 : 0;
    if (nvar == 3) return (npe <= 15 && npe % 2 == 0) ? 4 : 0;   // odd node counts: 72*npe bytes per row block is not a multiple of 16
    if (nvar == 4) return npe <= 255 ? 8 : 0;
    return 0;
}

static bool shape_instantiated(int nvar, int npe) {
    if (nvar == 2) return npe == 3 || npe == 4 || npe == 6 || npe == 9;
    if (nvar == 3) return npe == 4 || npe == 8 || npe == 10;
    if (nvar == 4) return npe == 4 || npe == 8 || npe == 10 || npe == 27;
    return false;
}

template <int NV, int BITS, int NPE>
static int launch_T(fl_handle* h, const double* ke, double* V, cudaStream_t st) {
    using SM = gather_warp_smem<NV, BITS, gather_batch<NV, NPE>::B, NPE>;
    auto kern = csr_gather_reg_kernel<NV, BITS, NPE>;
    const size_t smem = sizeof(SM) * GW;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GW * 32, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = (h->gplan.nitems + GW - 1) / GW;
    const int64_t cap = (int64_t)h->sm_count * occ;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return FL_OK;
    kern<<<(unsigned)blocks, GW * 32, smem, st>>>(h->gplan, ke, V);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

void gather_plan_release(GatherPlan& g) {
    cudaFree(g.items); cudaFree(g.recs);
    g = GatherPlan();
}

void gather_plan_free(fl_handle* h) { gather_plan_release(h->gplan); }

bool reg_gather_supported(const fl_handle* h, int nvar) {
    return h->pat.nbr_ptr != nullptr && !h->ke_plane_major && bits_for(nvar, h->npe) != 0 && shape_instantiated(nvar, h->npe) &&
           h->nelem > 0 && h->nelem * (int64_t)h->npe < ((int64_t)1 << 31);
}

// Where the register gather beats the row-buffer kernels in element order (profiles/gather_shapes_bench.py, 0.5-3 M elements, B200):
// every 2-D shape (tri3 3.0x, tri6 2.7x, quad4 1.3x, quad9 2.6x), hex8 with nvar = 3 (1.23x) and tet10 with nvar = 4 (2.9x); it is
// slower for tet4 (0.8x) and tet10 with nvar = 3 (0.9x) and equal for hex8 / hex27 with nvar = 4.
bool reg_gather_preferred(const fl_handle* h, int nvar) {
    if (!reg_gather_supported(h, nvar)) return false;
    if (h->use_reg_gather == 1) return true;
    if (h->use_reg_gather != 2) return false;
    return nvar == 2 || (nvar == 3 && h->npe == 8) || (nvar == 4 && h->npe == 10);
}

int gather_plan_ensure(fl_handle* h, int nvar) {
    if (h->gplan.nvar == nvar && h->gplan.recs) return FL_OK;
    return gather_plan_build(h, nvar, nullptr, false, &h->gplan);
}

// flat_store = nullptr: records hold adj_idx (K_e stored in element order).  by_completion: items sorted by the largest stored
// flat index of their node (stable: the items of a node stay adjacent), the order in which a streamed element kernel completes them.
int gather_plan_build(fl_handle* h, int nvar, const int32_t* flat_store, bool by_completion, GatherPlan* out) {
    GatherPlan& g = *out;
    gather_plan_release(g);
    const int bits = bits_for(nvar, h->npe);
    if (!bits) { set_error("register gather does not support nvar=%d with %d nodes per element", nvar, h->npe); return FL_ERR_UNSUPPORTED; }
    const Layout L = make_layout(nvar, bits);
    const int64_t nnode = h->nnode, nvisit = h->nelem * h->npe;
    Buf cnt_items, cnt_words, item_ptr, word_ptr, tmp;
    auto fail = [&](int code) { gather_plan_release(g); return code; };
#define FL_GTRY(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                            \
            return fail(FL_ERR_CUDA);                                                             \
        }                                                                                         \
    } while (0)
    FL_GTRY(cnt_items.alloc(sizeof(int64_t) * (nnode + 1)));
    FL_GTRY(cnt_words.alloc(sizeof(int64_t) * (nnode + 1)));
    FL_GTRY(item_ptr.alloc(sizeof(int64_t) * (nnode + 1)));
    FL_GTRY(word_ptr.alloc(sizeof(int64_t) * (nnode + 1)));
    FL_GTRY(cudaMemset(cnt_items.p, 0, sizeof(int64_t) * (nnode + 1)));
    FL_GTRY(cudaMemset(cnt_words.p, 0, sizeof(int64_t) * (nnode + 1)));
    item_counts_kernel<<<(unsigned)((nnode + 255) / 256), 256>>>(h->adj_ptr, h->pat.nbr_ptr, nnode, L, cnt_items.as<int64_t>(),
                                                                 cnt_words.as<int64_t>());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt_items.as<int64_t>(), item_ptr.as<int64_t>(), nnode + 1);
    FL_GTRY(tmp.alloc(tb));
    FL_GTRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt_items.as<int64_t>(), item_ptr.as<int64_t>(), nnode + 1));
    FL_GTRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt_words.as<int64_t>(), word_ptr.as<int64_t>(), nnode + 1));
    FL_GTRY(cudaMemcpy(&g.nitems, item_ptr.as<int64_t>() + nnode, sizeof(int64_t), cudaMemcpyDeviceToHost));
    FL_GTRY(cudaMemcpy(&g.nwords, word_ptr.as<int64_t>() + nnode, sizeof(int64_t), cudaMemcpyDeviceToHost));
    FL_GTRY(cudaMalloc(&g.items, sizeof(GatherItem) * (size_t)(g.nitems > 0 ? g.nitems : 1)));
    const size_t rec_bytes = sizeof(uint32_t) * (size_t)(g.nwords > 0 ? g.nwords : 2);
    FL_GTRY(cudaMalloc(&g.recs, rec_bytes));
    FL_GTRY(cudaMemset(g.recs, 0xFF, rec_bytes));
    Buf keys, keys_out, iota, order, tmp2, items_unsorted;
    const bool sorted = by_completion && g.nitems > 0;
    if (sorted) {
        FL_GTRY(keys.alloc(sizeof(int32_t) * g.nitems));
        FL_GTRY(keys_out.alloc(sizeof(int32_t) * g.nitems));
        FL_GTRY(iota.alloc(sizeof(int32_t) * g.nitems));
        FL_GTRY(order.alloc(sizeof(int32_t) * g.nitems));
        FL_GTRY(items_unsorted.alloc(sizeof(GatherItem) * g.nitems));
    }
    items_kernel<<<(unsigned)((nnode + 255) / 256), 256>>>(h->adj_ptr, h->pat.nbr_ptr, item_ptr.as<int64_t>(), word_ptr.as<int64_t>(), nnode, L,
                                                           sorted ? items_unsorted.as<GatherItem>() : g.items,
                                                           flat_store ? flat_store : h->adj_idx, sorted ? keys.as<int32_t>() : nullptr,
                                                           iota.as<int32_t>());
    if (sorted) {
        size_t tb2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb2, keys.as<int32_t>(), keys_out.as<int32_t>(), iota.as<int32_t>(), order.as<int32_t>(),
                                        (int)g.nitems);
        FL_GTRY(tmp2.alloc(tb2));
        FL_GTRY(cub::DeviceRadixSort::SortPairs(tmp2.p, tb2, keys.as<int32_t>(), keys_out.as<int32_t>(), iota.as<int32_t>(),
                                                order.as<int32_t>(), (int)g.nitems));
        permute_items_kernel<<<(unsigned)((g.nitems + 255) / 256), 256>>>(items_unsorted.as<GatherItem>(), order.as<int32_t>(), g.nitems,
                                                                         g.items);
    }
    if (nvisit > 0) {
        record_heads_kernel<<<(unsigned)((nvisit + 255) / 256), 256>>>(h->conn, h->adj_ptr, h->pat.nbr_ptr, h->adj_idx, flat_store,
                                                                      word_ptr.as<int64_t>(), nvisit, L, g.recs);
        record_fields_kernel<<<(unsigned)((nvisit * h->npe + 255) / 256), 256>>>(h->conn, h->adj_ptr, h->pat.nbr_ptr, h->adj_idx,
                                                                                h->pat.rank_adj, word_ptr.as<int64_t>(), nvisit, h->npe, L,
                                                                                g.recs);
    }
    FL_GTRY(cudaGetLastError());
    FL_GTRY(cudaDeviceSynchronize());
#undef FL_GTRY
    g.nvar = nvar;
    g.bits = bits;
    return FL_OK;
}

int launch_csr_gather_reg(fl_handle* h, int nvar, const double* ke, double* V, cudaStream_t st) {
    int rc = gather_plan_ensure(h, nvar);
    if (rc) return rc;
#define FL_GCASE(NV_, BITS_, NPE_) \
    if (nvar == NV_ && h->npe == NPE_) return launch_T<NV_, BITS_, NPE_>(h, ke, V, st)
    FL_GCASE(2, 4, 3); FL_GCASE(2, 4, 4); FL_GCASE(2, 4, 6); FL_GCASE(2, 4, 9);
    FL_GCASE(3, 4, 4); FL_GCASE(3, 4, 8); FL_GCASE(3, 4, 10);
    FL_GCASE(4, 8, 4); FL_GCASE(4, 8, 8); FL_GCASE(4, 8, 10); FL_GCASE(4, 8, 27);
#undef FL_GCASE
    set_error("register gather: unsupported nvar=%d, %d nodes per element", nvar, h->npe);
    return FL_ERR_UNSUPPORTED;
}

}  // namespace fl
