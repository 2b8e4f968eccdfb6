// fl_stream.cu -- CSR assembly with K_e stored along a space-filling curve: the warp-autonomous LinearElastic kernel on tet10 / hex8
// (launch_stream_iso_csr), any material through the ordinary element-kernel dispatch (launch_curve_csr), and the experiment of
// running the element kernel and the CSR reduction CONCURRENTLY.
//
// Same result, bit for bit, as implicit_iso_warp_kernel followed by csr_gather_kernel (reference: _GlobalAssemblyDF_,
// _LowLevelAssemblyDF_.h:8-176, with the slot-map scatter of SparseAssemblyNative.h:32-45).  The two-pass form writes 7.2 kB of K_e per
// tet10 element to HBM and reads it back: 14.4 of the 19 GB the step moves for 6.6 GB of algorithmic traffic (profiles/r1_summary.md).
//   * The elements are walked in Morton order of their centroids.  That is only a STORAGE order -- row block (position p, local node
//     a) of the K_e scratch; the summation order of every CSR entry stays ascending ORIGINAL element number, because the adjacency is
//     not reordered, only the addresses it points to.
//   * The reduction (fl_gather.cuh) walks the (node, <= 96 slots) items in the order in which the element walk completes them.
// use_stream = 1 (default): the two kernels run one after the other (the traction reduction on a side stream).  The reduction reads the row blocks of one element, which
//     straddle 128-byte lines, close together in time: 2.10 ms against 2.48 ms for csr_gather_kernel on config 2 (998 250 tet10).
// use_stream = 2: the reduction runs on a second stream BESIDE the element kernel (4 warps per SM fit next to its two blocks: the
//     register file holds 2 x 128 x 200 + 128 x 112), waits (acquire) for per-group release flags the element kernel publishes, and
//     pulls the rows from L2 with cp.async.cg.  93 % of the rows of config 2 are produced within 11 000 curve positions (80 MB) of the
//     moment their node completes, so they are still in L2.  Correct and deadlock-free (the element kernel never waits; a reduction
//     warp that waits ~0.5 s gives up and raises an error flag) -- but measured 3x SLOWER than mode 1 (profiles/r2_config2_study.md):
//     the reduction needs ~1250 warp instructions per element, and one warp per scheduler issues them at ~0.2 per cycle; its
//     throughput is proportional to the resident warps, and the element kernel leaves room for a quarter of what it needs.
// use_stream = 3: the kernels of mode 2 one after the other (flags already set when the reduction starts); for A/B timing and tests.
#include <cstdlib>

#include "fl_gather.cuh"
#include "fl_implicit_warp.cuh"

namespace fl {

namespace {
struct Buf {
    void* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes > 0 ? bytes : 8); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace

__global__ void positions_kernel(const int64_t* __restrict__ perm, int64_t nelem, int32_t* __restrict__ pos) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < nelem) pos[perm[p]] = (int32_t)p;
}

__global__ void permute_conn_kernel(const int32_t* __restrict__ conn, const int64_t* __restrict__ perm, int64_t nelem, int npe,
                                    int32_t* __restrict__ conn_p) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nelem * npe) return;
    const int64_t p = t / npe;
    conn_p[t] = conn[perm[p] * npe + (t - p * npe)];
}

__global__ void permute_adj_kernel(const int32_t* __restrict__ adj_idx, const int32_t* __restrict__ pos, int64_t nvisit, int npe,
                                   int32_t* __restrict__ adj_idx_p) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nvisit) return;
    const int32_t flat = adj_idx[k], e = flat / npe;
    adj_idx_p[k] = pos[e] * npe + (flat - e * npe);
}

constexpr int SG_WARPS = 4;   // reduction warps per SM beside the two element blocks (register file: 2 x 128 x 200 + 128 x 112)
constexpr int SG_B = 4;       // visits per step: 2 x 4 x 720 B of row blocks per warp fit beside the element blocks' shared memory

template <int NPE>
__global__ void __maxnreg__(112)
csr_gather_stream_kernel(const GatherPlan gp, const double* __restrict__ ke, double* __restrict__ V, const int32_t* __restrict__ flags,
                         int32_t epoch, int32_t* __restrict__ err) {
    using SM = gather_warp_smem<3, 4, SG_B, NPE>;
    extern __shared__ __align__(16) unsigned char smem_s[];
    SM* sm = reinterpret_cast<SM*>(smem_s) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    gather_warp_loop<3, 4, SG_B, NPE, true>(gp, wid, nw, ke, V, lane, *sm, flags, epoch, err);
}

// Two-pass mode (use_stream == 1, the default): same plan, no flags -- K_e in curve order, items in completion order, two blocks of
// eight warps per SM, launched behind the element kernel on the same stream.
template <int NPE>
__global__ void __launch_bounds__(256)
csr_gather_curve_kernel(const GatherPlan gp, const double* __restrict__ ke, double* __restrict__ V) {
    using SM = gather_warp_smem<3, 4, 6, NPE>;
    extern __shared__ __align__(16) unsigned char smem_s[];
    SM* sm = reinterpret_cast<SM*>(smem_s) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    gather_warp_loop<3, 4, 6, NPE, false>(gp, wid, nw, ke, V, lane, *sm, nullptr, 0, nullptr);
}

// Any material / element kernel (launch_curve_csr): the same reduction for every (nvar, nodes per element) the gather is instantiated for
template <int NV, int BITS, int NPE>
__global__ void __launch_bounds__(256)
csr_gather_curve_any_kernel(const GatherPlan gp, const double* __restrict__ ke, double* __restrict__ V) {
    using SM = gather_warp_smem<NV, BITS, gather_batch<NV, NPE>::B, NPE>;
    extern __shared__ __align__(16) unsigned char smem_s[];
    SM* sm = reinterpret_cast<SM*>(smem_s) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    gather_warp_loop<NV, BITS, gather_batch<NV, NPE>::B, NPE, false>(gp, wid, nw, ke, V, lane, *sm, nullptr, 0, nullptr);
}

// T[n] = sum over the visits of node n (ascending original element number) of the per-element tractions stored in curve order
template <int NV>
__global__ void gather_traction_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx_p, const double* __restrict__ te,
                                       int64_t nnode, double* __restrict__ T) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    const int64_t k1 = adj_ptr[n + 1];
    for (int64_t k = adj_ptr[n]; k < k1; ++k) {
        const int64_t idx = adj_idx_p[k];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += te[idx * NV + i];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) T[n * NV + i] = acc[i];
}

static int plan_build(fl_handle* h, int nvar) {
    StreamPlan& sp = h->splan;
    const int npe = h->npe;
    const int64_t nelem = h->nelem, nvisit = nelem * npe;
    const int epw = npe >= 2 ? 32 / (npe / 2) : 32;   // element-kernel group (only the iso path publishes flags)
    Buf perm, pos;
    auto fail = [&](int code) { stream_plan_free(h); return code; };
#define FL_STRY(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                            \
            return fail(FL_ERR_CUDA);                                                             \
        }                                                                                         \
    } while (0)
    FL_STRY(perm.alloc(sizeof(int64_t) * nelem));
    FL_STRY(pos.alloc(sizeof(int32_t) * nelem));
    int rc = launch_sfc_order_conn(h->points, h->conn, nelem, npe, h->ndim, h->nnode, perm.as<int64_t>(), nullptr);
    if (rc) return fail(rc);
    FL_STRY(cudaMalloc(&sp.conn_p, sizeof(int32_t) * nvisit));
    FL_STRY(cudaMalloc(&sp.adj_idx_p, sizeof(int32_t) * nvisit));
    positions_kernel<<<(unsigned)((nelem + 255) / 256), 256>>>(perm.as<int64_t>(), nelem, pos.as<int32_t>());
    permute_conn_kernel<<<(unsigned)((nvisit + 255) / 256), 256>>>(h->conn, perm.as<int64_t>(), nelem, npe, sp.conn_p);
    permute_adj_kernel<<<(unsigned)((nvisit + 255) / 256), 256>>>(h->adj_idx, pos.as<int32_t>(), nvisit, npe, sp.adj_idx_p);
    FL_STRY(cudaGetLastError());
    rc = gather_plan_build(h, nvar, sp.adj_idx_p, true, &sp.gp);
    if (rc) return fail(rc);
    sp.ngroups = (nelem + epw - 1) / epw;
    FL_STRY(cudaMalloc(&sp.flags, sizeof(int32_t) * (sp.ngroups + 1)));
    FL_STRY(cudaMemset(sp.flags, 0, sizeof(int32_t) * (sp.ngroups + 1)));
    FL_STRY(cudaMalloc(&sp.err, sizeof(int32_t)));
    FL_STRY(cudaMemset(sp.err, 0, sizeof(int32_t)));
    FL_STRY(cudaStreamCreateWithFlags(&sp.side, cudaStreamNonBlocking));
    FL_STRY(cudaEventCreateWithFlags(&sp.fork, cudaEventDisableTiming));
    FL_STRY(cudaEventCreateWithFlags(&sp.join, cudaEventDisableTiming));
    FL_STRY(cudaDeviceSynchronize());
#undef FL_STRY
    sp.npe = npe;
    sp.nvar = nvar;
    sp.epoch = 0;
    return FL_OK;
}

template <int NPE>
static int launch_T(fl_handle* h, const double* Eulerx, const MatParams& prm, int update, double* V, double* T, cudaStream_t st) {
    using S = iso_warp_shape<NPE, 8>;
    StreamPlan& sp = h->splan;
    sp.epoch = sp.epoch == INT32_MAX ? 1 : sp.epoch + 1;
    const bool concurrent = h->use_stream == 2;
    auto ekern = h->use_stream == 1 ? implicit_iso_warp_kernel<NPE, 8, false> : implicit_iso_warp_kernel<NPE, 8, true>;
    auto gkern = csr_gather_stream_kernel<NPE>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(ekern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ekern, IW_WARPS * 32, S::SMEM));
    if (occ < 1) occ = 1;
    const int64_t ngroups = sp.ngroups;
    const int64_t nblk = (ngroups + IW_WARPS - 1) / IW_WARPS;
    const int egrid = (int)(nblk < (int64_t)occ * h->sm_count ? nblk : (int64_t)occ * h->sm_count);
    int64_t gblocks = (sp.gp.nitems + SG_WARPS - 1) / SG_WARPS;
    if (gblocks > h->sm_count) gblocks = h->sm_count;
    cudaStream_t gs = concurrent ? sp.side : st;
    if (concurrent) {
        FL_CUDA_CHECK(cudaEventRecord(sp.fork, st));
        FL_CUDA_CHECK(cudaStreamWaitEvent(sp.side, sp.fork, 0));
    }
    ekern<<<egrid, IW_WARPS * 32, S::SMEM, st>>>(sp.conn_p, h->points, Eulerx, h->jm, h->gw, h->nelem, h->ldg, update, prm, h->ke, h->te,
                                                  sp.flags, sp.epoch);
    FL_CUDA_CHECK(cudaGetLastError());
    if (h->timing && h->ev[1]) cudaEventRecord(h->ev[1], st);
    if (h->use_stream == 1) {
        // the traction reduction only needs the element kernel: it runs on the side stream beside the CSR reduction
        FL_CUDA_CHECK(cudaEventRecord(sp.fork, st));
        FL_CUDA_CHECK(cudaStreamWaitEvent(sp.side, sp.fork, 0));
        gather_traction_kernel<3><<<(unsigned)((h->nnode + 255) / 256), 256, 0, sp.side>>>(h->adj_ptr, sp.adj_idx_p, h->te, h->nnode, T);
        FL_CUDA_CHECK(cudaGetLastError());
        FL_CUDA_CHECK(cudaEventRecord(sp.join, sp.side));
        auto ckern = csr_gather_curve_kernel<NPE>;
        const size_t csmem = sizeof(gather_warp_smem<3, 4, 6, NPE>) * 8;
        FL_CUDA_CHECK(cudaFuncSetAttribute(ckern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        ckern<<<2 * h->sm_count, 256, csmem, st>>>(sp.gp, h->ke, V);
        FL_CUDA_CHECK(cudaGetLastError());
        if (h->timing && h->ev[2]) cudaEventRecord(h->ev[2], st);
        FL_CUDA_CHECK(cudaStreamWaitEvent(st, sp.join, 0));
        if (h->timing && h->ev[3]) cudaEventRecord(h->ev[3], st);
        return FL_OK;
    } else if (gblocks > 0) {
        const size_t gsmem = sizeof(gather_warp_smem<3, 4, SG_B, NPE>) * SG_WARPS;
        FL_CUDA_CHECK(cudaFuncSetAttribute(gkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        gkern<<<(unsigned)gblocks, SG_WARPS * 32, gsmem, gs>>>(sp.gp, h->ke, V, sp.flags, sp.epoch, sp.err);
        FL_CUDA_CHECK(cudaGetLastError());
    }
    if (concurrent) {
        FL_CUDA_CHECK(cudaEventRecord(sp.join, sp.side));
        FL_CUDA_CHECK(cudaStreamWaitEvent(st, sp.join, 0));
    }
    if (h->timing && h->ev[2]) cudaEventRecord(h->ev[2], st);
    gather_traction_kernel<3><<<(unsigned)((h->nnode + 255) / 256), 256, 0, st>>>(h->adj_ptr, sp.adj_idx_p, h->te, h->nnode, T);
    FL_CUDA_CHECK(cudaGetLastError());
    if (h->timing && h->ev[3]) cudaEventRecord(h->ev[3], st);
    return FL_OK;
}

void stream_plan_free(fl_handle* h) {
    StreamPlan& sp = h->splan;
    if (sp.side) { cudaStreamSynchronize(sp.side); cudaStreamDestroy(sp.side); }
    if (sp.fork) cudaEventDestroy(sp.fork);
    if (sp.join) cudaEventDestroy(sp.join);
    cudaFree(sp.conn_p); cudaFree(sp.adj_idx_p); cudaFree(sp.flags); cudaFree(sp.err);
    gather_plan_release(sp.gp);
    sp = StreamPlan();
}

bool stream_csr_supported(const fl_handle* h) {
    return h->use_stream && h->use_warp_iso && h->ndim == 3 && h->ng == 8 && h->nelem > 0 && h->pat.nbr_ptr != nullptr &&
           (h->npe == 10 || (h->npe == 8 && h->use_warp_iso == 2)) && h->nelem * (int64_t)h->npe < ((int64_t)1 << 31);
}

// Error flag of the most recent streamed call(s): blocks until the stream has drained (debugging / tests only).
static int stream_check(fl_handle* h, cudaStream_t st) {
    if (!h->splan.err) return FL_OK;
    int32_t e = 0;
    FL_CUDA_CHECK(cudaMemcpyAsync(&e, h->splan.err, sizeof(e), cudaMemcpyDeviceToHost, st));
    FL_CUDA_CHECK(cudaStreamSynchronize(st));
    if (e) {
        cudaMemsetAsync(h->splan.err, 0, sizeof(int32_t), st);
        set_error("streamed CSR assembly: the reduction kernel gave up waiting for the element kernel");
        return FL_ERR_CUDA;
    }
    return FL_OK;
}

int launch_stream_iso_csr(fl_handle* h, const double* Eulerx, const fl_material* mat, int update, double* V, double* T, cudaStream_t st) {
    if (h->splan.npe != h->npe || h->splan.nvar != 3) {
        stream_plan_free(h);
        int rc = plan_build(h, 3);
        if (rc) return rc;
    }
    MatParams p;
    p.mu = mat->mu; p.mu1 = mat->mu1; p.mu2 = mat->mu2; p.mu3 = mat->mu3; p.mue = mat->mue; p.lamb = mat->lamb;
    p.eps_1 = mat->eps_1; p.eps_2 = mat->eps_2; p.eps_3 = mat->eps_3; p.eps_e = mat->eps_e;
    int rc = h->npe == 10 ? launch_T<10>(h, Eulerx, p, update, V, T, st) : launch_T<8>(h, Eulerx, p, update, V, T, st);
    if (rc) return rc;
    // FL_STREAM_CHECK (tests): wait for the call and turn a raised error flag into an error code
    return (h->use_stream >= 2 && getenv("FL_STREAM_CHECK") != nullptr) ? stream_check(h, st) : FL_OK;
}

// ---- any material: element kernel of the ordinary dispatch on the curve-ordered connectivity + the reduction in completion order
template <int NV, int BITS, int NPE>
static int launch_curve_gather(fl_handle* h, double* V, cudaStream_t st) {
    using SM = gather_warp_smem<NV, BITS, gather_batch<NV, NPE>::B, NPE>;
    auto kern = csr_gather_curve_any_kernel<NV, BITS, NPE>;
    const size_t smem = sizeof(SM) * 8;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
    if (occ < 1) occ = 1;
    kern<<<occ * h->sm_count, 256, smem, st>>>(h->splan.gp, h->ke, V);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// Shapes for which K_e along the curve + the register gather in completion order was measured faster than the best element-order
// reduction (profiles/gather_shapes_bench.py)
bool curve_csr_preferred(const fl_handle* h, int nvar) {
    if (h->use_stream != 1 || h->ke_plane_major || !reg_gather_supported(h, nvar)) return false;
    if (getenv("FL_CURVE_ALL") != nullptr) return true;      // tests: every shape the gather is instantiated for
    return nvar == 3 && h->npe == 10;                         // tet10 mechanics: reduction 1.34 -> 1.18 ms per 547 k elements; no gain elsewhere
}

int launch_curve_csr(fl_handle* h, int nvar, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation, int update,
                     double* V, double* T, cudaStream_t st, cudaEvent_t after_elements, cudaEvent_t after_reduction) {
    if (h->splan.npe != h->npe || h->splan.nvar != nvar) {
        stream_plan_free(h);
        int rc = plan_build(h, nvar);
        if (rc) return rc;
    }
    StreamPlan& sp = h->splan;
    int32_t* conn = h->conn;
    h->conn = sp.conn_p;                       // the element kernels walk the elements in storage (curve) order
    int rc = launch_implicit_elements(h, Eulerx, Eulerp, mat, formulation, update, h->ke, h->te, st);
    h->conn = conn;
    if (rc) return rc;
    if (after_elements) cudaEventRecord(after_elements, st);
    {   // the traction reduction only needs the element kernel: it runs on the side stream beside the CSR reduction
        FL_CUDA_CHECK(cudaEventRecord(sp.fork, st));
        FL_CUDA_CHECK(cudaStreamWaitEvent(sp.side, sp.fork, 0));
        const unsigned nb = (unsigned)((h->nnode + 255) / 256);
        switch (nvar) {
            case 2: gather_traction_kernel<2><<<nb, 256, 0, sp.side>>>(h->adj_ptr, sp.adj_idx_p, h->te, h->nnode, T); break;
            case 3: gather_traction_kernel<3><<<nb, 256, 0, sp.side>>>(h->adj_ptr, sp.adj_idx_p, h->te, h->nnode, T); break;
            default: gather_traction_kernel<4><<<nb, 256, 0, sp.side>>>(h->adj_ptr, sp.adj_idx_p, h->te, h->nnode, T); break;
        }
        FL_CUDA_CHECK(cudaGetLastError());
        FL_CUDA_CHECK(cudaEventRecord(sp.join, sp.side));
    }
    rc = FL_ERR_UNSUPPORTED;
#define FL_CCASE(NV_, BITS_, NPE_) \
    if (nvar == NV_ && h->npe == NPE_) rc = launch_curve_gather<NV_, BITS_, NPE_>(h, V, st)
    FL_CCASE(2, 4, 3); FL_CCASE(2, 4, 4); FL_CCASE(2, 4, 6); FL_CCASE(2, 4, 9);
    FL_CCASE(3, 4, 4); FL_CCASE(3, 4, 8); FL_CCASE(3, 4, 10);
    FL_CCASE(4, 8, 4); FL_CCASE(4, 8, 8); FL_CCASE(4, 8, 10); FL_CCASE(4, 8, 27);
#undef FL_CCASE
    if (rc) {
        cudaStreamWaitEvent(st, sp.join, 0);      // never leave the side stream un-joined
        if (rc == FL_ERR_UNSUPPORTED) set_error("curve-ordered CSR assembly: unsupported shape");
        return rc;
    }
    if (after_reduction) cudaEventRecord(after_reduction, st);
    FL_CUDA_CHECK(cudaStreamWaitEvent(st, sp.join, 0));
    return FL_OK;
}

}  // namespace fl
