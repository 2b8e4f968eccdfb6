// fl_gather.cuh -- register-resident CSR value reduction fed by asynchronous copies ("slot-owner gather").
//
// Same sum as csr_gather_kernel (fl_pattern.cu) and the reference's slot-map scatter (SparseAssemblyNative.h:32-45,
// _MassIntegrand_.h:115-166): the NV rows of node n in V are the sum, over the elements of n in ascending order, of the NV rows of
// K_e that belong to n, placed by the rank of the column node.  The shared-memory row buffer of csr_gather_kernel costs ~400
// shared-memory wavefronts per tet10 element, 44 % of them bank conflicts of the rank-scattered read-modify-writes
// (profiles/r1_summary.md).  Here the roles are turned round: a lane OWNS a node slot of the CSR rows and pulls what belongs to it.
//
//   work item  = (node n, up to 96 consecutive node slots of its neighbour list = up to 3 groups of 32 slots); one warp per item
//   lane       = slot 32 g + lane of the item; accumulators acc[g][i][j] for the NV x NV block of that slot in registers
//   step       = up to B visits (elements) of the item.  The K_e row blocks of the step's visits (NV rows x ndof columns, contiguous)
//                are copied into shared memory with cp.async.cg (16-byte pieces, all lanes busy), one step ahead of the step
//                being reduced: the data in flight lives in shared memory, not in registers, and no lane waits for a global
//                load.  Every sector of K_e is fetched once.  (cp.async.bulk, one copy per visit, was measured first: the
//                per-lane copies are issued through the uniform datapath one lane at a time, 33 instructions per visit.)
//   per visit  = one packed record: the flat index of the (element, local node a) visit and, per node slot of the item, the local
//                node b of the element that sits in that slot, or NONE.  A lane reads the one word that holds its field (a
//                broadcast) and, if the slot is present, adds the block K_e[(a,i)][(b,j)] from the staged row block.
//   write-out  = V[row i][NV*slot + j] from the owning lane (NV consecutive doubles per lane and row).
// Visits are added in ascending element order and every accumulator starts from +0.0: bit-identical to csr_gather_kernel.
#pragma once
#include "fl_internal.cuh"

namespace fl {

constexpr int GATHER_SLOTS = 32;   // node slots per group: one per lane

template <int NV, int BITS>
struct gather_cfg {
    static constexpr unsigned NONE = (1u << BITS) - 1u;       // field value of an empty slot
    static constexpr int FPW = 32 / BITS;                     // fields per 32-bit word (no field straddles a word)
    static constexpr int FWG = GATHER_SLOTS / FPW;            // field words per group (= BITS)
    static constexpr int MAXG = NV <= 3 ? 3 : 1;              // groups per item: NV*NV*MAXG accumulators per lane
    static constexpr int RWMAX = 2 + FWG * MAXG;              // record: flat index, spare word, fields of up to MAXG groups
    static_assert(FWG % 2 == 0, "record layout (records stay 8-byte aligned)");
};

struct GatherStep {   // what the three pipeline stages of a warp pass to each other (shared memory)
    int64_t v_base;
    int32_t nv, w, nslots, ng, flags, pad;  // flags: 1 = first step of the item, 2 = last step, 4 = no more work
};

// visits per step: ~4.5 kB of row blocks per buffer (two buffers per warp)
template <int NV, int NPE>
struct gather_batch {
    static constexpr int RAW = 4608 / (NV * NV * NPE * 8);
    static constexpr int B = RAW < 2 ? 2 : (RAW > 8 ? 8 : RAW);
};

// Depth of the per-warp software pipeline, in steps: at iteration s the records of step s + GP_DR are requested (cp.async), the
// row blocks of step s + GP_DW are requested (cp.async.cg, 16-byte pieces) and step s is reduced.  The item descriptors are
// requested GP_DI items ahead.  A warp that owns few registers and runs beside the element kernel has no other way of covering the latencies: nothing it
// waits for may have been requested less than a few steps ago.
constexpr int GP_DR = 4, GP_DW = 1, GP_NROW = GP_DW + 1, GP_NREC = 8, GP_DI = GP_DR - GP_DW + 1, GP_NDESC = 8;
static_assert(GP_NREC >= GP_DR + 2 && (GP_NREC & (GP_NREC - 1)) == 0 && GP_NDESC >= GP_DI + 1, "ring sizes");

// Shared memory of one warp
template <int NV, int BITS, int B, int NPE>
struct gather_warp_smem {
    static constexpr int ROWD = NV * NV * NPE;                // doubles of one (element, node) row block of K_e
    static_assert((ROWD * 8) % 16 == 0, "the row blocks are copied in 16-byte pieces");
    alignas(16) double rows[GP_NROW][B][ROWD];
    alignas(16) GatherItem desc[GP_NDESC];
    alignas(16) GatherStep steps[GP_NREC];
    alignas(8) uint32_t recs[GP_NREC][B * gather_cfg<NV, BITS>::RWMAX];
};

// The work loop of one warp: items first, first + stride, ...
//   STREAM = the element kernel is still running (fl_stream.cu): before the row blocks of a step are requested, wait until the
//            element kernel has published every group the step reads (flags[group] == epoch, acquire).
template <int NV, int BITS, int B, int NPE, bool STREAM>
__device__ __forceinline__ void gather_warp_loop(const GatherPlan& gp, int64_t first, int64_t stride, const double* __restrict__ ke,
                                                 double* __restrict__ V, int lane, gather_warp_smem<NV, BITS, B, NPE>& sm,
                                                 const int32_t* __restrict__ flags, int32_t epoch, int32_t* __restrict__ err) {
    using C = gather_cfg<NV, BITS>;
    constexpr int MG = C::MAXG, ROWD = NV * NV * NPE, NDOF = NV * NPE;
    constexpr int GD = NPE * (32 / (NPE / 2));                // flat connectivity indices per element-kernel group (STREAM)
    static_assert(B <= 32, "one lane per visit of a step");
    // per-lane constants: word (within a group's field words) and shift of the lane's slot field
    const int word = 2 + lane / C::FPW, shift = BITS * (lane % C::FPW);
    // descriptors of the warp's first GP_DI items: plain loads; later ones arrive by cp.async, GP_DI items ahead of the cursor
    if (lane < GP_DI) {
        GatherItem it;
        it.rec_off = 0; it.v_base = 0; it.nvis = 0; it.w = 0; it.nslots = 0; it.ng = 0;
        const int64_t k = first + lane * stride;
        if (k < gp.nitems) it = gp.items[k];
        sm.desc[lane] = it;
    }
    __syncwarp();
    // ---- stage R: the cursor walks (item, visit batch); the step's descriptor and records travel to shared memory
    int64_t ck = first;            // cursor: item, its index in the warp's sequence, first visit of the next step
    int ci = 0, cvb = 0;
    auto stage_R = [&](int s) {
        GatherStep st;
        st.v_base = 0; st.w = 0; st.nslots = 0; st.ng = 0; st.pad = 0; st.nv = 0; st.flags = 4;
        if (ck < gp.nitems) {
            const GatherItem cd = sm.desc[ci & (GP_NDESC - 1)];
            if (cvb == 0 && lane < 2) {     // entering an item: request the descriptor GP_DI items ahead (two 16-byte pieces)
                const int64_t kn = ck + GP_DI * stride;
                if (kn < gp.nitems) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.desc[(ci + GP_DI) & (GP_NDESC - 1)]) + 16 * lane;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const char*>(gp.items + kn) + 16 * lane) : "memory");
                }
            }
            const int rw = 2 + C::FWG * cd.ng;
            st.v_base = cd.v_base; st.w = cd.w; st.nslots = cd.nslots; st.ng = cd.ng;
            st.nv = min(B, cd.nvis - cvb);
            st.flags = (cvb == 0 ? 1 : 0) | (cvb + B >= cd.nvis ? 2 : 0);
            // the step's records are contiguous: nv * rw words, copied as 8-byte pieces
            const uint2* src = reinterpret_cast<const uint2*>(gp.recs + cd.rec_off + (int64_t)cvb * rw);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm.recs[s & (GP_NREC - 1)]);
            for (int c = lane; c < st.nv * rw / 2; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8 * c), "l"(src + c) : "memory");
            cvb += B;
            if (cvb >= cd.nvis) { ck += stride; ++ci; cvb = 0; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (lane == 0) sm.steps[s & (GP_NREC - 1)] = st;
    };
    // ---- stage W: records of step s have landed -> (wait for the producer) -> request the row blocks (L2 -> shared memory, never
    //      through this SM's L1: another SM may have written them during this kernel)
    auto stage_W = [&](int s, int buf) {
        asm volatile("cp.async.wait_group %0;" ::"n"(2 * (GP_DR - GP_DW)) : "memory");
        __syncwarp();
        const GatherStep st = sm.steps[s & (GP_NREC - 1)];
        const int rw = 2 + C::FWG * st.ng;
        const uint32_t* rec = sm.recs[s & (GP_NREC - 1)];
        if (STREAM && st.nv > 0) {
            const int32_t* fp = flags + (lane < st.nv ? rec[lane * rw] : 0u) / (unsigned)GD;
            int spins = 0;
            while (true) {
                int32_t f = epoch;
                if (lane < st.nv) asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(f) : "l"(fp) : "memory");
                if (__all_sync(0xffffffffu, f == epoch)) break;
                if (++spins > (1 << 21)) {        // ~0.5 s: the element kernel is not running; give up loudly instead of hanging
                    if (lane == 0) atomicExch(err, 1);
                    break;
                }
                __nanosleep(256);
            }
            __syncwarp();
        }
        // 16-byte pieces: lane o copies piece o (and o + 32, ...) of every visit's row block
        constexpr int PPV = ROWD / 2;
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(sm.rows[buf][0]) + 16 * lane;
        const double* src0 = ke + 2 * lane;
        for (int v = 0; v < st.nv; ++v) {
            const double* src = src0 + (int64_t)rec[v * rw] * ROWD;
            const unsigned dst = dst0 + v * (ROWD * 8);
#pragma unroll
            for (int o = 0; o < PPV; o += 32)
                if (o + 32 <= PPV || lane < PPV - o)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * o), "l"(src + 2 * o) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // prologue: steps 0 .. GP_DR-1 have their records requested, steps 0 .. GP_DW-1 their rows.  Every copy is a cp.async group:
    // in the steady state iteration t commits R(t + GP_DR) and then W(t + GP_DW), so "all but the 2*(GP_DR-GP_DW) most recent
    // groups" covers the records stage W needs and "all but the 2*GP_DW most recent" the rows the reduction needs.
#pragma unroll 1
    for (int s = 0; s < GP_DR; ++s) {
        stage_R(s);
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    int bufW = 0;
#pragma unroll 1
    for (int s = 0; s < GP_DW; ++s) {
        stage_W(s, bufW);
        bufW = bufW + 1 == GP_NROW ? 0 : bufW + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    double acc[MG][NV][NV];
    int bufC = 0;
    for (int s = 0;; ++s) {
        stage_R(s + GP_DR);
        stage_W(s + GP_DW, bufW);
        bufW = bufW + 1 == GP_NROW ? 0 : bufW + 1;
        const GatherStep st = sm.steps[s & (GP_NREC - 1)];
        if (st.flags & 4) break;
        if (st.flags & 1) {
#pragma unroll
            for (int g = 0; g < MG; ++g)
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < NV; ++j) acc[g][i][j] = 0.0;
        }
        asm volatile("cp.async.wait_group %0;" ::"n"(2 * GP_DW) : "memory");
        __syncwarp();
        const int rw = 2 + C::FWG * st.ng;
        const uint32_t* rec = sm.recs[s & (GP_NREC - 1)] + word;
        const double* row = sm.rows[bufC][0];
        // one visit: the lane's slot of group g receives the block of local node b (if the element has a node in that slot)
        auto add_group = [&](const uint32_t* r, const double* rowv, int g) {
            const unsigned b = (r[g * C::FWG] >> shift) & C::NONE;
            const bool on = b != C::NONE;
            const double* src = rowv + (on ? (int)(b * NV) : 0);   // absent slots load (and drop) the block of local node 0
            double val[NV][NV];
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < NV; ++j) val[i][j] = src[i * NDOF + j];
            if (on) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < NV; ++j) acc[g][i][j] = __dadd_rn(acc[g][i][j], val[i][j]);
            }
        };
        if (st.ng == 1) {              // nodes with at most 32 neighbours: the common case, no group tests in the loop
#pragma unroll 2
            for (int v = 0; v < st.nv; ++v, rec += rw, row += ROWD) add_group(rec, row, 0);
        } else {
            for (int v = 0; v < st.nv; ++v, rec += rw, row += ROWD) {
#pragma unroll
                for (int g = 0; g < MG; ++g)
                    if (g < st.ng) add_group(rec, row, g);
            }
        }
        if (st.flags & 2) {
#pragma unroll
            for (int g = 0; g < MG; ++g)
                if (g * GATHER_SLOTS + lane < st.nslots) {
                    double* out = V + st.v_base + (int64_t)(g * GATHER_SLOTS + lane) * NV;
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int j = 0; j < NV; ++j) __stcs(out + (int64_t)i * st.w + j, acc[g][i][j]);
                }
        }
        bufC = bufC + 1 == GP_NROW ? 0 : bufC + 1;
        __syncwarp();   // every lane is done with the step's buffers before the stages of the next iteration overwrite them
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace fl
