// fl_implicit_mma.cuh -- element stiffness of high-order hexahedra (p >= 2: hex27, hex64) on the fp64 tensor-core path.
//
// Same quantities as implicit_elements_kernel (fl_implicit.cuh; reference _LowLevelAssemblyDF_.h:69-131,
// _LowLevelAssemblyDPF_.h:73-150, _ConstitutiveStiffnessDF_.h:83-156, _GeometricStiffness_.h:61-144), written in the
// parent-element form that turns the local B^T H B into dense GEMMs with a SHARED operand:
//
//   grad_x N_a = J_x^-1 Jm_g[:,a]   =>   K[(a,i),(b,j)] = sum_g sum_pq Jm[p][a][g] * Chat_g[i,p,j,q] * Jm[q][b][g]
//   Chat_g[i,p,j,q] = detJ sum_kl J_x^-1[k][p] ( C_g[i,k,j,l] + delta_ij sigma_g[k][l] ) J_x^-1[l][q]
//   C[i,k,j,l] = H_voigt[VI(i,k)][VI(j,l)]  (VI = Voigt index of the pair, or HS+k for the potential dof)
//
// so that per dof pair (i,j):   K^{ij} (npe x npe) = JmT (npe x 3ng)  *  W^{ij} (3ng x npe),
//                               W^{ij}[(g,p)][b]   = sum_q Chat_g[i,p,j,q] Jm[q][b][g]
// JmT is the same for every element and every (i,j): its DMMA A-fragments are streamed from L1; W^{ij} is produced chunk by
// chunk in shared memory by scalar FMAs (3 per entry, 2-10 % of the work) and consumed as B-fragments.  The geometric
// stiffness rides in Chat for free, and no spatial-gradient array is ever stored.
#pragma once
#include "fl_explicit_mma.cuh"

namespace fl {

#ifndef FL_IMMA_THREADS
#define FL_IMMA_THREADS 256
#endif
constexpr int IMMA_THREADS = FL_IMMA_THREADS;  // warps 0-3 issue the DMMAs, the remaining warps produce the next W chunk (double buffer)

template <int NPE, int NG, int NV, int KC>
struct imma_shape {
    static constexpr int C3 = 3 * NV;                       // rows/cols of Chat_g
    static constexpr int MT = (NPE + 7) / 8, NT = (NPE + 7) / 8, KS = (3 * NG + 3) / 4;
    static constexpr int NCH = (KS + KC - 1) / KC;          // K chunks
    static constexpr int LDW = NT * 8 + 4;                  // == 4 (mod 16): conflict-free 64-bit B-fragment loads (see fl_explicit_mma.cuh)
    static constexpr int CH_SZ = NG * C3 * C3, W_SZ = KC * 4 * LDW, XX_SZ = NPE * 7, P_SZ = NG * C3;
    static constexpr size_t SMEM = sizeof(double) * (size_t)(CH_SZ + 2 * W_SZ + XX_SZ + P_SZ);
};

template <int NV>
__device__ __forceinline__ int voigt_index(int i, int k) {
    // mechanics rows: Voigt index of the unordered pair (i,k); potential dof (i == 3): HS + k
    if (i == 3) return 6 + k;
    if (i == k) return i;
    const int s = i + k;  // (0,1)->1, (0,2)->2, (1,2)->3
    return s + 2;         // -> 3, 4, 5
}

// Two blocks per SM: while one block evaluates the per-element prologue (kinematics, Hessian, Chat: 64 of its 256 threads), the
// other one keeps the tensor pipe busy.  That needs K chunks of 6 k-steps (107 KB of shared memory per block) and a 128-register
// cap -- the prologue spills ~1.5 KB per thread to local memory, still a net win: 11.7 -> 10.9 ms on 13 824 hex64 elements
// (1 block/SM: KC = 12 11.7 ms, KC = 6 12.5 ms; 192 threads x 2 blocks: 14.6 ms, the two producer warps cannot keep up).
#ifndef FL_IMMA_MINB
#define FL_IMMA_MINB 2
#endif
#ifndef FL_IMMA_KC64
#define FL_IMMA_KC64 6
#endif
template <int MAT, int NPE, int NG, int KC>
__global__ void __launch_bounds__(IMMA_THREADS, FL_IMMA_MINB)
implicit_elements_mma_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                             const double* __restrict__ phi, const double* __restrict__ jm, const double* __restrict__ jmT,
                             const double* __restrict__ gw, int64_t nelem, int ldg, int update, MatParams prm,
                             double* __restrict__ ke, double* __restrict__ te, int plane_major) {
    constexpr int D = 3;
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr bool GEO = mat_traits<MAT>::geometric;
    constexpr int NV = D + (EL ? 1 : 0), HS = 6, HT = HS + (EL ? D : 0);
    using S = imma_shape<NPE, NG, NV, KC>;
    constexpr int C3 = S::C3, MTW = (S::MT + 3) / 4;
    constexpr int ndof = NPE * NV;
    extern __shared__ double smem[];
    double* CH = smem;                 // [g][(i,p)][(j,q)]
    double* Ws = CH + S::CH_SZ;        // 2 x [KC*4][LDW] K chunks of W^{ij} (double buffer)
    double* XXs = Ws + 2 * S::W_SZ;        // [a][7]        X, x, phi
    double* Pt = XXs + S::XX_SZ;       // [g][p][i]     traction operand  detJ J_x^-T [sigma | D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;

    for (int64_t e = blockIdx.x; e < nelem; e += gridDim.x) {
        __syncthreads();
        for (int a = threadIdx.x; a < NPE; a += IMMA_THREADS) {
            const int64_t n = conn[e * NPE + a];
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                XXs[a * 7 + l] = X[n * 3 + l];
                XXs[a * 7 + 3 + l] = x[n * 3 + l];
            }
            XXs[a * 7 + 6] = EL ? phi[n] : 0.0;
        }
        __syncthreads();
        // ---- per Gauss point: kinematics, material, Chat_g and the traction operand
        for (int g = threadIdx.x; g < NG; g += IMMA_THREADS) {
            double JX[9], Jx[9], gp[3] = {0, 0, 0};
#pragma unroll
            for (int i = 0; i < 9; ++i) JX[i] = Jx[i] = 0.0;
            for (int a = 0; a < NPE; ++a) {
                double j[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) j[k] = jmT[(g * 3 + k) * NPE + a];
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const double Xa = XXs[a * 7 + l], xa = XXs[a * 7 + 3 + l];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        JX[k * 3 + l] = fma(j[k], Xa, JX[k * 3 + l]);
                        Jx[k * 3 + l] = fma(j[k], xa, Jx[k * 3 + l]);
                    }
                }
                if (EL) {
                    const double p = XXs[a * 7 + 6];
#pragma unroll
                    for (int k = 0; k < 3; ++k) gp[k] = fma(j[k], p, gp[k]);
                }
            }
            double iJX[9], iJx[9];
            const double detX = invdet(JX, iJX);
            const double detx = invdet(Jx, iJx);
            const double detJ = gw[g] * fabs(update == 1 ? detx : detX);   // _KinematicMeasures_.h:94-99
            double F[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    double v = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) v = fma(iJX[l * 3 + k], Jx[k * 3 + i], v);
                    F[i * 3 + l] = v;
                }
            double E[3], Dv[3], sig[9], hess[HT * HT];
            if (EL) {
#pragma unroll
                for (int k = 0; k < 3; ++k) E[k] = -(iJx[k * 3] * gp[0] + iJx[k * 3 + 1] * gp[1] + iJx[k * 3 + 2] * gp[2]);
            }
            kinetic_measures<D, MAT, true>(F, E, prm, sig, Dv, hess);
            // Chat[i,p,j,q] = detJ sum_kl iJx[k][p] (C[i,k,j,l] + delta_ij sigma_sym[k][l]) iJx[l][q]
            double* Cg = CH + g * C3 * C3;
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = i; j < NV; ++j) {   // only the dof pairs i <= j are contracted below (K^{ji} is the mirror image)
                    // tmp[k][q] = sum_l (C[i,k,j,l] + geo) iJx[l][q]
                    double tmp[3][3];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            double v = 0;
#pragma unroll
                            for (int l = 0; l < 3; ++l) {
                                double c = hess[voigt_index<NV>(i, k) * HT + voigt_index<NV>(j, l)];
                                if (GEO && i == j && i < 3) c += (k <= l ? sig[k * 3 + l] : sig[l * 3 + k]);
                                v = fma(c, iJx[l * 3 + q], v);
                            }
                            tmp[k][q] = v;
                        }
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            double v = 0;
#pragma unroll
                            for (int k = 0; k < 3; ++k) v = fma(iJx[k * 3 + p], tmp[k][q], v);
                            Cg[(i * 3 + p) * C3 + j * 3 + q] = v * detJ;
                        }
                }
            // traction operand P[p][i] = detJ sum_k iJx[k][p] [sigma_sym | D][k][i]
#pragma unroll
            for (int p = 0; p < 3; ++p) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double v = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) v = fma(iJx[k * 3 + p], (k <= i ? sig[k * 3 + i] : sig[i * 3 + k]), v);
                    Pt[g * C3 + p * NV + i] = v * detJ;
                }
                if (EL) Pt[g * C3 + p * NV + 3] = (iJx[p] * Dv[0] + iJx[3 + p] * Dv[1] + iJx[6 + p] * Dv[2]) * detJ;
            }
        }
        __syncthreads();
        // ---- traction t_a = sum_g Jm_g[:,a]^T P_g (only when the geometry is updated)
        for (int a = threadIdx.x; a < NPE; a += IMMA_THREADS) {
            double t[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] = 0.0;
            if (update == 1) {
                for (int g = 0; g < NG; ++g) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const double j = jmT[(g * 3 + p) * NPE + a];
#pragma unroll
                        for (int i = 0; i < NV; ++i) t[i] = fma(j, Pt[g * C3 + p * NV + i], t[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) te[(e * NPE + a) * NV + i] = t[i];
        }
        // ---- K^{ij} = JmT * W^{ij}: steps s = (ij, chunk).  Warps 4-7 produce W(s+1) into the other buffer while warps 0-3
        //      run the DMMAs of step s; one block barrier per step.
        double* Ke = ke + (size_t)e * ndof * ndof;
        // K is symmetric (H and the geometric term are), so only the dof pairs i <= j are contracted; K^{ji} = (K^{ij})^T is
        // written as the mirror image.
        constexpr int NPAIR = NV * (NV + 1) / 2;
        constexpr int NS = NPAIR * S::NCH;
        auto pair_ij = [](int pr, int& i, int& j) {
            i = 0;
            int rem = pr;
            while (rem >= NV - i) { rem -= NV - i; ++i; }
            j = i + rem;
        };
        auto produce = [&](int s, int t0, int nthr) {
            const int ij = s / S::NCH, ch = s - ij * S::NCH;
            int i, j;
            pair_ij(ij, i, j);
            double* Wb = Ws + (s & 1) * S::W_SZ;
            constexpr int NB = S::NT * 8;          // padded column count
            constexpr int GPC = (KC * 4) / 3;      // Gauss points per chunk (KC*4 is a multiple of 3)
            static_assert((KC * 4) % 3 == 0, "a K chunk must hold whole Gauss points");
            // item = (Gauss point of the chunk, column node b): three rows (p = 0..2) share the three jmT loads.
            // b is the fastest index: coalesced jmT reads, Chat reads are warp-wide broadcasts.
            const int total = GPC * NB;
#pragma unroll 4
            for (int it = t0; it < total; it += nthr) {
                const int gl = it / NB, b = it - gl * NB;
                const int g = ch * GPC + gl;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0;
                if (g < NG && b < NPE) {
                    const double* jg = jmT + (g * 3) * NPE + b;
                    const double j0 = jg[0], j1 = jg[NPE], j2 = jg[2 * NPE];
                    const double* cg = CH + g * C3 * C3 + (i * 3) * C3 + j * 3;
                    w0 = fma(cg[0], j0, fma(cg[1], j1, cg[2] * j2));
                    w1 = fma(cg[C3], j0, fma(cg[C3 + 1], j1, cg[C3 + 2] * j2));
                    w2 = fma(cg[2 * C3], j0, fma(cg[2 * C3 + 1], j1, cg[2 * C3 + 2] * j2));
                }
                double* wo = Wb + (gl * 3) * S::LDW + b;
                wo[0] = w0;
                wo[S::LDW] = w1;
                wo[2 * S::LDW] = w2;
            }
        };
        __syncthreads();
        produce(0, threadIdx.x, IMMA_THREADS);
        __syncthreads();
        double c[MTW][S::NT][2];
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
            const int ij = s / S::NCH, ch = s - ij * S::NCH;
            if (warp >= 4) {
                if (s + 1 < NS) produce(s + 1, threadIdx.x - 128, IMMA_THREADS - 128);
            } else {
                if (ch == 0) {
#pragma unroll
                    for (int m = 0; m < MTW; ++m)
#pragma unroll
                        for (int q = 0; q < S::NT; ++q) c[m][q][0] = c[m][q][1] = 0.0;
                }
                const double* Wb = Ws + (s & 1) * S::W_SZ;
#pragma unroll
                for (int ks = 0; ks < KC; ++ks) {
                    const int r = (ch * KC + ks) * 4 + lc;        // global k index of this lane's A/B element
                    const int g = r / 3, p = r - 3 * g;
                    double afr[MTW];
#pragma unroll
                    for (int m = 0; m < MTW; ++m) {
                        const int a = 8 * (warp + 4 * m) + lr;
                        afr[m] = (a < NPE && r < 3 * NG) ? jmT[(g * 3 + p) * NPE + a] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < S::NT; ++q) {
                        const double bfr = Wb[(4 * ks + lc) * S::LDW + 8 * q + lr];
#pragma unroll
                        for (int m = 0; m < MTW; ++m) dmma884(c[m][q][0], c[m][q][1], afr[m], bfr);
                    }
                }
                if (ch == S::NCH - 1) {
                    int i, j;
                    pair_ij(ij, i, j);
#pragma unroll
                    for (int m = 0; m < MTW; ++m) {
                        const int a = 8 * (warp + 4 * m) + lr;
                        if (a < NPE) {
#pragma unroll
                            for (int q = 0; q < S::NT; ++q)
#pragma unroll
                                for (int z = 0; z < 2; ++z) {
                                    const int b = 8 * q + 2 * lc + z;
                                    if (b < NPE) {
                                        if (plane_major) {
                                            // K_e scratch as dof-pair planes [(i,j)][a][b]: the 4 lanes of a fragment row write 64
                                            // contiguous bytes, and so do the 8 lanes of a fragment column in the mirror plane
                                            Ke[((size_t)(i * NV + j) * NPE + a) * NPE + b] = c[m][q][z];
                                            if (i != j) Ke[((size_t)(j * NV + i) * NPE + b) * NPE + a] = c[m][q][z];
                                        } else {
                                            Ke[(size_t)(a * NV + i) * ndof + b * NV + j] = c[m][q][z];
                                            if (i != j) Ke[(size_t)(b * NV + j) * ndof + a * NV + i] = c[m][q][z];
                                        }
                                    }
                                }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

template <int MAT, int NPE, int NG, int KC>
int launch_impl_mma(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                    cudaStream_t st) {
    constexpr int NV = 3 + (mat_traits<MAT>::electro ? 1 : 0);
    using S = imma_shape<NPE, NG, NV, KC>;
    auto kern = implicit_elements_mma_kernel<MAT, NPE, NG, KC>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IMMA_THREADS, S::SMEM));
    if (occ < 1) occ = 1;
    const int grid = (int)(h->nelem < (int64_t)occ * h->sm_count ? h->nelem : (int64_t)occ * h->sm_count);
    if (grid == 0) return FL_OK;
    kern<<<grid, IMMA_THREADS, S::SMEM, st>>>(h->conn, h->points, Eulerx, Eulerp, h->jm, h->jmT, h->gw, h->nelem, h->ldg, update, prm, ke, te, h->ke_plane_major);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
