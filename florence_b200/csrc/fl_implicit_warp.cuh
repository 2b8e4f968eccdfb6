// fl_implicit_warp.cuh -- warp-autonomous implicit element kernel for the isotropic constant-tangent material
// (LinearElastic, _LinearElastic_.h:24-57) on small 3-D elements with 8 Gauss points (tet10, hex8: ndof <= 32).
//
// Same arithmetic as the ISO path of implicit_elements_kernel (fl_implicit.cuh) -- K_ab = lamb S + mu S^T + mu tr(S) I with
// S_ab = sum_g detJ grad N_a (x) grad N_b -- but organised so that one WARP carries a group of 32/NG elements through every
// phase with __syncwarp only (the block-wide kernel idles at its five barriers, profiles/r1_summary.md):
//   phase 1  lane = (element, Gauss point): kinematics, stress, spatial gradients -> shared memory
//   phase 3  lane = (element, column node b): the lane owns the 3x3 blocks K_ab of its column node for every row node a.
//            Its 8 x 3 weighted column gradients stay in registers, the row gradients are (per element) broadcast loads
//            feeding 9 FMAs each -- a third of the shared-memory wavefronts of a lane-per-dof mapping -- and S^T, tr(S) are
//            register-local.  The six K_e rows of a row-node pair (1440 contiguous bytes per element) pass through a small
//            per-warp tile and leave as 16-byte, fully coalesced stores: 8-byte stores straight from the lanes fill a third
//            of every 32-byte sector they touch and saturate the L1 -> crossbar write path (profiles/r2 notes).
//   phase 4  lane = (element, node a): traction
// The next group's nodal coordinates are prefetched with cp.async while the current group is computed.
#pragma once
#include "fl_internal.cuh"

namespace fl {

#ifndef FL_IW_MINB
#define FL_IW_MINB 3
#endif
#ifndef FL_IW_UNROLL_A
#define FL_IW_UNROLL_A 1
#endif
constexpr int IW_WARPS = 4;   // warps per block

template <int NPE, int NG>
struct iso_warp_shape {
    static constexpr int D = 3;
    static constexpr int EPW = (32 / NG < 32 / NPE) ? 32 / NG : 32 / NPE;   // elements per warp group (tet10: 3, hex8: 4)
    static constexpr int NDOF = NPE * D;
    static constexpr int XS = NDOF | 1;              // per-element stride of the coordinate tiles
    static constexpr int SGS = NDOF;                 // per-Gauss-point stride of the gradient tile (even: 16-byte row-pair loads)
    static constexpr int SGE = NG * SGS + ((18 - (NG * SGS) % 16) % 16);   // per-element stride == 2 (mod 16): the elements of a
                                                                          // group sit 4 banks apart, so their broadcasts never collide
    static constexpr int SSS = 9;                    // sigma * detJ
    static constexpr int SSE = NG * SSS + 1;         // per-element stride of the stress tile (distinct banks per element)
    static constexpr int LDG = NG | 1;
    // per warp: 2 x (X, x) double-buffered, gradients, stresses, detJ
    static constexpr int X_OFF = 0;
    static constexpr int x_OFF = X_OFF + 2 * EPW * XS;
    static constexpr int SG_OFF = (x_OFF + 2 * EPW * XS + 1) & ~1;
    static constexpr int SS_OFF = SG_OFF + EPW * SGE;
    static constexpr int DJ_OFF = SS_OFF + EPW * SSE;
    static constexpr int KT_OFF = (DJ_OFF + EPW * NG + 1) & ~1;
    static constexpr int KTE = 2 * D * NDOF + 2;     // per-element stride of the K row-pair tile (6 rows of ndof doubles)
    static constexpr int WARP_DOUBLES = KT_OFF + EPW * KTE;
    static constexpr int JM_DOUBLES = D * NPE * LDG;
    static constexpr size_t SMEM = sizeof(double) * (JM_DOUBLES + NG + (size_t)IW_WARPS * WARP_DOUBLES);
    static_assert(NG * EPW <= 32 && NPE * EPW <= 32 && EPW >= 1, "a group must fit a warp in both mappings");
    static_assert(NPE % 2 == 0 && (JM_DOUBLES + NG) % 2 == 0 && SGE % 2 == 0 && SGS % 2 == 0 && (NDOF * NDOF) % 2 == 0 && KTE % 2 == 0,
                  "row-pair loads and the 16-byte write-out need 16-byte alignment");
};

template <int NPE, int NG>
__global__ void __launch_bounds__(IW_WARPS * 32, FL_IW_MINB)
implicit_iso_warp_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                         const double* __restrict__ jm_g, const double* __restrict__ gw_g, int64_t nelem, int ldg_g, int update,
                         MatParams prm, double* __restrict__ ke, double* __restrict__ te) {
    using S = iso_warp_shape<NPE, NG>;
    constexpr int D = 3, EPW = S::EPW, NDOF = S::NDOF;
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
    int64_t grp = (int64_t)blockIdx.x * IW_WARPS + warp;
    if (grp < ngroups) gather(grp, 0);
    int buf = 0;
    for (; grp < ngroups; grp += wstride, buf ^= 1) {
        const int64_t e0 = grp * EPW;
        const int ne = (int)min((int64_t)EPW, nelem - e0);
        asm volatile("cp.async.wait_all;");
        __syncwarp();
        if (grp + wstride < ngroups) gather(grp + wstride, buf ^ 1);
        // ---- phase 1: lane = (element, Gauss point)
        {
            const int el = lane / NG, g = lane - el * NG;
            const double* Xe = Xs0 + (buf * EPW + el) * S::XS;
            const double* xe = xs0 + (buf * EPW + el) * S::XS;
            double JX[9], Jx[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) JX[i] = Jx[i] = 0.0;
            if (el < ne) {   // el >= EPW for the spare lanes of a 3-element group
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
        // ---- phase 3 / 4: lane = (element, node)
        {
            const int el = lane / NPE, b = lane - el * NPE;
            const bool act = el < ne;
            const int elc = act ? el : 0;            // idle lanes shadow element 0 (loads only)
            const double* sge = SG + elc * S::SGE;
            double bgv[NG][D];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double d = dJ[elc * NG + g];
#pragma unroll
                for (int jj = 0; jj < D; ++jj) bgv[g][jj] = sge[g * S::SGS + b * D + jj] * d;
            }
            // two row nodes per trip: their 6 gradient components are three 16-byte loads
            constexpr int UA = FL_IW_UNROLL_A;
#pragma unroll(UA)
            for (int a2 = 0; a2 < NPE / 2; ++a2) {
                double Sm[2][D][D];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int jj = 0; jj < D; ++jj) Sm[q][i][jj] = 0.0;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const double2* ap2 = reinterpret_cast<const double2*>(sge + g * S::SGS + a2 * 2 * D);
                    const double2 p0 = ap2[0], p1 = ap2[1], p2 = ap2[2];
                    const double ap[2 * D] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y};
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int jj = 0; jj < D; ++jj) Sm[q][i][jj] = fma(ap[q * D + i], bgv[g][jj], Sm[q][i][jj]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    double tr = 0;
#pragma unroll
                    for (int i = 0; i < D; ++i) tr += Sm[q][i][i];
                    double* kt = KT + elc * S::KTE + (q * D) * NDOF + b * D;
                    if (act) {
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int jj = 0; jj < D; ++jj)
                                kt[i * NDOF + jj] =
                                    __dadd_rn(fma(prm.lamb, Sm[q][i][jj], __dmul_rn(prm.mu, Sm[q][jj][i])), i == jj ? __dmul_rn(prm.mu, tr) : 0.0);
                    }
                }
                __syncwarp();
                // rows 6 a2 .. 6 a2 + 5 of every element of the group: D * NDOF double2 per element, contiguous in K_e
                {
                    constexpr int PER = D * NDOF;          // double2 per element
                    double2* dst = reinterpret_cast<double2*>(ke + (size_t)e0 * NDOF * NDOF + (size_t)a2 * 2 * D * NDOF);
                    for (int f = lane; f < ne * PER; f += 32) {
                        const int el2 = f / PER, r = f - el2 * PER;
                        dst[(size_t)el2 * (NDOF * NDOF / 2) + r] = reinterpret_cast<const double2*>(KT + el2 * S::KTE)[r];
                    }
                }
                __syncwarp();
            }
            // traction t_b = sum_g grad N_b . (sigma detJ)   (only when the geometry is updated, _LowLevelAssemblyDF_.h:136-148)
            if (act) {
                double t[D];
#pragma unroll
                for (int i = 0; i < D; ++i) t[i] = 0.0;
                if (update == 1) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const double* ag = sge + g * S::SGS + b * D;
                        const double* So = Ss + el * S::SSE + g * S::SSS;
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int l = 0; l < D; ++l) t[i] = fma(ag[l], l <= i ? So[l * D + i] : So[i * D + l], t[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < D; ++i) te[((e0 + el) * NPE + b) * D + i] = t[i];
            }
        }
        __syncwarp();
    }
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
    kern<<<grid, IW_WARPS * 32, S::SMEM, st>>>(h->conn, h->points, Eulerx, h->jm, h->gw, h->nelem, h->ldg, update, prm, ke, te);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
