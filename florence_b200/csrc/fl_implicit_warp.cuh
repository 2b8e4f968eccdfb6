// fl_implicit_warp.cuh -- warp-autonomous implicit element kernel for the isotropic constant-tangent material
// (LinearElastic, _LinearElastic_.h:24-57) on small 3-D elements with 8 Gauss points and an even node count (tet10, hex8).
//
// Same arithmetic as the ISO path of implicit_elements_kernel (fl_implicit.cuh) -- K_ab = lamb S + mu S^T + mu tr(S) I with
// S_ab = sum_g detJ grad N_a (x) grad N_b -- but organised so that one WARP carries a group of elements through every phase
// with __syncwarp only (the block-wide kernel idles at its five barriers, profiles/r1_summary.md), and so that the shared
// memory pipe, which bounds both kernels, moves as few wavefronts per element as possible:
//   phase 1  lane = (element, Gauss point): kinematics, stress, spatial gradients -> shared memory
//   phase 3  lane = (element, PAIR of adjacent column nodes 2t, 2t+1).  The lane keeps the 8 x 6 weighted gradients of its
//            two column nodes in registers and, per (row node a, Gauss point), reads the 3 gradient components of a once
//            for 18 FMAs.  All lanes of a warp read at most EPW distinct addresses, one per element, placed 4 banks apart:
//            every load is a conflict-free broadcast.  S^T and tr(S) are register-local.
//            The three K_e rows of row node a (720 contiguous bytes per element) are written to a small per-warp tile
//            (conflict-free 16-byte stores, 48 contiguous bytes per lane and row) and leave through the TMA engine
//            (cp.async.bulk shared -> global, one copy per element and row node).  Stores issued from the lanes either
//            fill a third to a half of each 32-byte sector (L1 -> crossbar write path saturates) or, staged and
//            coalesced, cost the LSU data pipe -- the resource that bounds this kernel -- ~180 more wavefronts per element.
//   phase 4  same lanes: traction of the two nodes
// The next group's nodal coordinates are prefetched with cp.async while the current group is computed.
#pragma once
#include "fl_internal.cuh"

namespace fl {

#ifndef FL_IW_MINB
#define FL_IW_MINB 2
#endif
constexpr int IW_WARPS = 4;   // warps per block

template <int NPE, int NG>
struct iso_warp_shape {
    static constexpr int D = 3;
    static constexpr int LPE = NPE / 2;              // lanes per element in phase 3 (one per node pair)
    static constexpr int EPW = 32 / LPE;             // elements per warp group (tet10: 6, hex8: 8)
    static constexpr int EPR = 32 / NG;              // elements per phase-1 round
    static constexpr int ROUNDS = (EPW + EPR - 1) / EPR;
    static constexpr int NDOF = NPE * D;
    static constexpr int XS = NDOF | 1;              // per-element stride of the coordinate tiles
    static constexpr int SGS = NDOF | 1;             // per-Gauss-point stride of the gradient tile (odd: the phase-1 stores of
                                                     // one element spread over distinct banks)
    static constexpr int SGE = NG * SGS + ((25 - (NG * SGS) % 16) % 16);   // per-element stride == 9 (mod 16) doubles: the per-element
                                                     // broadcast addresses of phase 3 fall into distinct banks (0,9,2,11,4,13,...)
    static constexpr int SSS = 9;                    // sigma * detJ
    static constexpr int SSE = NG * SSS + 1;         // per-element stride of the stress tile (distinct banks per element)
    // per warp: 2 x (X, x) double-buffered, gradients, stresses, detJ
    static constexpr int X_OFF = 0;
    static constexpr int x_OFF = X_OFF + 2 * EPW * XS;
    static constexpr int SG_OFF = (x_OFF + 2 * EPW * XS + 1) & ~1;
    static constexpr int SS_OFF = SG_OFF + EPW * SGE;
    static constexpr int DJ_OFF = SS_OFF + EPW * SSE;
    static constexpr int KT_OFF = (DJ_OFF + EPW * NG + 1) & ~1;
    static constexpr int KTE = D * NDOF + ((30 - (D * NDOF) % 16) % 16);   // K row tile, per-element stride == 14 (mod 16) doubles:
                                                     // the 16-byte stores of every quarter warp hit 8 distinct bank quads
    static constexpr int WARP_DOUBLES = KT_OFF + EPW * KTE;
    static constexpr int LDG = NG | 1;
    static constexpr int JM_DOUBLES = D * NPE * LDG;
    static constexpr size_t SMEM = sizeof(double) * (JM_DOUBLES + NG + (size_t)IW_WARPS * WARP_DOUBLES);
    static_assert(NPE % 2 == 0 && LPE * EPW <= 32 && EPR >= 1, "a group must fit a warp");
    static_assert((JM_DOUBLES + NG) % 2 == 0 && (NDOF * NDOF) % 2 == 0 && NDOF % 2 == 0 && KTE % 2 == 0 && (D * NDOF) % 2 == 0,
                  "16-byte stores need 16-byte aligned rows");
};

// STREAM: the kernel publishes its progress for a concurrently running CSR reduction (fl_stream.cu).  flags[group] = epoch is
// stored with release semantics once every K_e row of the group has reached global memory (the bulk copies have completed, not
// merely been read out of shared memory); the stores of a group are published after the kinematics phase of the warp's NEXT group,
// when their completion costs no wait.
template <int NPE, int NG, bool STREAM = false>
__global__ void __launch_bounds__(IW_WARPS * 32, FL_IW_MINB)
implicit_iso_warp_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                         const double* __restrict__ jm_g, const double* __restrict__ gw_g, int64_t nelem, int ldg_g, int update,
                         MatParams prm, double* __restrict__ ke, double* __restrict__ te, int32_t* __restrict__ flags, int32_t epoch) {
    using S = iso_warp_shape<NPE, NG>;
    constexpr int D = 3, EPW = S::EPW, NDOF = S::NDOF, LPE = S::LPE;
    extern __shared__ __align__(16) double smem_w[];
    double* jm = smem_w;                     // [k][a][LDG]
    double* gw = jm + S::JM_DOUBLES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* wbase = gw + NG + warp * S::WARP_DOUBLES;
    double* Xs0 = wbase + S::X_OFF;
    double* xs0 = wbase + S::x_OFF;
    double* SG = wbase + S::SG_OFF;
    double* Ss = wbase + S::SS_OFF;
    double* dJ = wbase + S::DJ_OFF;
    double* KT = wbase + S::KT_OFF;
    for (int i = threadIdx.x; i < D * NPE * NG; i += blockDim.x) {
        const int g = i % NG, ka = i / NG;
        jm[ka * S::LDG + g] = jm_g[ka * ldg_g + g];
    }
    if (threadIdx.x < NG) gw[threadIdx.x] = gw_g[threadIdx.x];
    __syncthreads();

    const int64_t ngroups = (nelem + EPW - 1) / EPW;
    const int64_t wstride = (int64_t)gridDim.x * IW_WARPS;
    auto gather = [&](int64_t grp, int buf) {
        const int64_t e0 = grp * EPW;
        const int ne = (int)min((int64_t)EPW, nelem - e0);
        for (int it = lane; it < ne * NPE; it += 32) {
            const int el = it / NPE, a = it - el * NPE;
            const int64_t n = conn[e0 * NPE + it];
            const unsigned sX = (unsigned)__cvta_generic_to_shared(Xs0 + (buf * EPW + el) * S::XS + a * D);
            const unsigned sx = (unsigned)__cvta_generic_to_shared(xs0 + (buf * EPW + el) * S::XS + a * D);
#pragma unroll
            for (int l = 0; l < D; ++l) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sX + 8 * l), "l"(X + n * D + l));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sx + 8 * l), "l"(x + n * D + l));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    // all K_e rows of group `g` are in global memory: publish
    auto publish = [&](int64_t g) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(flags + g), "r"(epoch) : "memory");
        }
    };
    int64_t grp = (int64_t)blockIdx.x * IW_WARPS + warp;
    if (grp < ngroups) gather(grp, 0);
    int buf = 0;
    int64_t prev = -1;
    for (; grp < ngroups; grp += wstride, buf ^= 1) {
        const int64_t e0 = grp * EPW;
        const int ne = (int)min((int64_t)EPW, nelem - e0);
        asm volatile("cp.async.wait_all;");
        __syncwarp();
        if (grp + wstride < ngroups) gather(grp + wstride, buf ^ 1);
        // ---- phase 1: lane = (element, Gauss point), EPR elements per round
#pragma unroll 1
        for (int round = 0; round < S::ROUNDS; ++round) {
            const int el = round * S::EPR + lane / NG, g = lane % NG;
            if (el < ne) {
                const double* Xe = Xs0 + (buf * EPW + el) * S::XS;
                const double* xe = xs0 + (buf * EPW + el) * S::XS;
                double JX[9], Jx[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) JX[i] = Jx[i] = 0.0;
#pragma unroll
                for (int a = 0; a < NPE; ++a) {
                    double j[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) j[k] = jm[(k * NPE + a) * S::LDG + g];
#pragma unroll
                    for (int l = 0; l < D; ++l) {
                        const double Xa = Xe[a * D + l], xa = xe[a * D + l];
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            JX[k * D + l] = fma(j[k], Xa, JX[k * D + l]);
                            Jx[k * D + l] = fma(j[k], xa, Jx[k * D + l]);
                        }
                    }
                }
                double iJX[9], iJx[9];
                const double detX = invdet(JX, iJX);
                const double detx = invdet(Jx, iJx);
                const double detJ = gw[g] * fabs(update == 1 ? detx : detX);   // _KinematicMeasures_.h:94-99
                double F[9];
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int l = 0; l < D; ++l) {
                        double v = 0;
#pragma unroll
                        for (int k = 0; k < D; ++k) v = fma(iJX[l * D + k], Jx[k * D + i], v);
                        F[i * D + l] = v;
                    }
                double sig[9];
                double hess_unused[36];
                kinetic_measures<3, MAT_LINEAR_ELASTIC, false>(F, nullptr, prm, sig, nullptr, hess_unused);
                double* So = Ss + el * S::SSE + g * S::SSS;
#pragma unroll
                for (int i = 0; i < 9; ++i) So[i] = sig[i] * detJ;
                dJ[el * NG + g] = detJ;
                double* sgo = SG + el * S::SGE + g * S::SGS;
#pragma unroll
                for (int a = 0; a < NPE; ++a) {
                    double j[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) j[k] = jm[(k * NPE + a) * S::LDG + g];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        double v = 0;
#pragma unroll
                        for (int jj = 0; jj < D; ++jj) v = fma(iJx[k * D + jj], j[jj], v);
                        sgo[a * D + k] = v;
                    }
                }
            }
        }
        __syncwarp();
        if (STREAM) {
            if (prev >= 0) publish(prev);
            prev = grp;
        }
        // ---- phase 3 / 4: lane = (element, node pair)
        {
            const int el = lane / LPE, t = lane - el * LPE;
            const bool act = el < ne;                // also false for the spare lanes (el >= EPW)
            const int elc = act ? el : 0;            // idle lanes shadow element 0 (loads only)
            const double* sge = SG + elc * S::SGE;
            double bgv[NG][2 * D];                   // weighted gradients of column nodes 2t, 2t+1
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double d = dJ[elc * NG + g];
#pragma unroll
                for (int c = 0; c < 2 * D; ++c) bgv[g][c] = sge[g * S::SGS + t * 2 * D + c] * d;
            }
            double* ktl = KT + elc * S::KTE + t * 2 * D;
#pragma unroll 1
            for (int a = 0; a < NPE; ++a) {
                double Sm[2][D][D];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int jj = 0; jj < D; ++jj) Sm[q][i][jj] = 0.0;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double* ap = sge + g * S::SGS + a * D;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double ai = ap[i];
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int jj = 0; jj < D; ++jj) Sm[q][i][jj] = fma(ai, bgv[g][q * D + jj], Sm[q][i][jj]);
                    }
                }
                double tr[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    tr[q] = 0;
#pragma unroll
                    for (int i = 0; i < D; ++i) tr[q] += Sm[q][i][i];
                }
                // the previous row block must have been read out of the tile before it is overwritten
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                if (act) {
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double kv[2 * D];
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int jj = 0; jj < D; ++jj)
                                kv[q * D + jj] = __dadd_rn(fma(prm.lamb, Sm[q][i][jj], __dmul_rn(prm.mu, Sm[q][jj][i])),
                                                           i == jj ? __dmul_rn(prm.mu, tr[q]) : 0.0);
                        double2* row = reinterpret_cast<double2*>(ktl + i * NDOF);
                        row[0] = make_double2(kv[0], kv[1]);
                        row[1] = make_double2(kv[2], kv[3]);
                        row[2] = make_double2(kv[4], kv[5]);
                    }
                }
                // rows 3a .. 3a+2 of each element (720 contiguous bytes in the tile and in K_e) leave through the TMA engine:
                // one bulk shared -> global copy per element, issued by one lane; the LSU data pipe never sees them
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane < ne) {
                    const unsigned src = (unsigned)__cvta_generic_to_shared(KT + lane * S::KTE);
                    double* dst = ke + (size_t)(e0 + lane) * NDOF * NDOF + (size_t)a * D * NDOF;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(D * NDOF * 8) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            // traction t_b = sum_g grad N_b . (sigma detJ)   (only when the geometry is updated, _LowLevelAssemblyDF_.h:136-148)
            if (act) {
                double tv[2 * D];
#pragma unroll
                for (int c = 0; c < 2 * D; ++c) tv[c] = 0.0;
                if (update == 1) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const double* So = Ss + el * S::SSE + g * S::SSS;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const double* ag = sge + g * S::SGS + (2 * t + q) * D;
#pragma unroll
                            for (int i = 0; i < D; ++i)
#pragma unroll
                                for (int l = 0; l < D; ++l) tv[q * D + i] = fma(ag[l], l <= i ? So[l * D + i] : So[i * D + l], tv[q * D + i]);
                        }
                    }
                }
                double2* trow = reinterpret_cast<double2*>(te + ((e0 + el) * NPE + 2 * t) * D);
                trow[0] = make_double2(tv[0], tv[1]);
                trow[1] = make_double2(tv[2], tv[3]);
                trow[2] = make_double2(tv[4], tv[5]);
            }
        }
        __syncwarp();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (STREAM && prev >= 0) publish(prev);
}

template <int NPE, int NG>
int launch_impl_iso_warp(fl_handle* h, const double* Eulerx, const MatParams& prm, int update, double* ke, double* te, cudaStream_t st) {
    using S = iso_warp_shape<NPE, NG>;
    auto kern = implicit_iso_warp_kernel<NPE, NG>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IW_WARPS * 32, S::SMEM));
    if (occ < 1) occ = 1;
    const int64_t ngroups = (h->nelem + S::EPW - 1) / S::EPW;
    const int64_t nblk = (ngroups + IW_WARPS - 1) / IW_WARPS;
    const int grid = (int)(nblk < (int64_t)occ * h->sm_count ? nblk : (int64_t)occ * h->sm_count);
    if (grid == 0) return FL_OK;
    kern<<<grid, IW_WARPS * 32, S::SMEM, st>>>(h->conn, h->points, Eulerx, h->jm, h->gw, h->nelem, h->ldg, update, prm, ke, te, nullptr, 0);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
