// fl_implicit_mma.cuh -- element stiffness of high-order elements (hex27, hex64, hex125 Poisson, tet20) on the fp64 tensor-core path.
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

template <int NPE, int NG, int NV, int KC, int NTB_ = 0>
struct imma_shape {
    static constexpr int C3 = 3 * NV;                       // rows/cols of Chat_g
    static constexpr int NPAIR = NV * (NV + 1) / 2;         // dof pairs i <= j that are contracted (K^{ji} is the mirror image)
    // m8n8 tiles over the padded node count; an even number of them, so that every DMMA warp carries two m-tiles (tet20: 3 -> 4;
    // the hexahedral shapes 27 / 64 / 125 already give 4 / 8 / 16)
    static constexpr int MT = (((NPE + 7) / 8) + 1) & ~1, NT = MT, KS = (3 * NG + 3) / 4;
    static constexpr int NTB = NTB_ > 0 ? NTB_ : NT;        // n-tiles per column block: bounds the accumulators (MT/4 x NTB x 2 doubles
    static constexpr int NCB = NT / NTB;                    // per lane) when the element is wide (hex125: 16 n-tiles in 4 blocks)
    static_assert(NT % NTB == 0, "column blocks must tile the n-tiles");
    static constexpr int NCH = (KS + KC - 1) / KC;          // K chunks
    static constexpr int LDW = NTB * 8 + 4;                 // == 4 (mod 16): conflict-free 64-bit B-fragment loads (see fl_explicit_mma.cuh)
    static constexpr int CH_SZ = (NG * NPAIR * 9 + 1) & ~1, W_SZ = KC * 4 * LDW;   // per-element Chat stride, even: 16-byte copies
    static constexpr size_t SMEM = sizeof(double) * (size_t)(CH_SZ + 2 * W_SZ);
    // prologue kernel: one thread per Gauss point, rounded up to whole warps
    static constexpr int PRO_THREADS = ((NG + 31) / 32) * 32;
    static constexpr size_t PRO_SMEM = sizeof(double) * (size_t)(NPE * 7 + NG * C3 + CH_SZ);
};

template <int NV>
__device__ __forceinline__ int voigt_index(int i, int k) {
    // mechanics rows: Voigt index of the unordered pair (i,k); potential dof (i == 3): HS + k
    if (i == 3) return 6 + k;
    if (i == k) return i;
    const int s = i + k;  // (0,1)->1, (0,2)->2, (1,2)->3
    return s + 2;         // -> 3, 4, 5
}

// ---------------------------------------------------------------------------------------------------------------------------
// Kernel 1 of 2: the per-element prologue.  One block per element, one thread per Gauss point: kinematics, the material law
// (stress + Hessian, Legendre transform for the electro-mechanical models), Chat_g for the dof pairs i <= j, and the traction.
// In round 1 this ran inside the tensor-core kernel on 64 of its 256 threads under a 128-register cap (744 B of spills per
// thread) while the DMMA warps of the block waited: ~45 % of that kernel's time.  As a kernel of its own it runs at full
// occupancy with its own register budget, and the tensor-core kernel below has nothing but GEMM work left.
// Output: chg[e][g][pair][3x3] (the only Chat blocks the GEMM reads, 10 of 16 for nvar = 4) and te[e][a][i].
template <int MAT, int NPE, int NG>
__global__ void __launch_bounds__(imma_shape<NPE, NG, 3 + (mat_traits<MAT>::electro ? 1 : 0), 6>::PRO_THREADS)
implicit_mma_prologue_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                             const double* __restrict__ phi, const double* __restrict__ jm, int ldg, const double* __restrict__ jmT,
                             const double* __restrict__ gw, int64_t nelem, int update, MatParams prm, double* __restrict__ chg,
                             double* __restrict__ te) {
    constexpr int D = 3;
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr bool GEO = mat_traits<MAT>::geometric;
    constexpr int NV = D + (EL ? 1 : 0), HS = 6, HT = HS + (EL ? D : 0);
    using S = imma_shape<NPE, NG, NV, 6>;
    constexpr int C3 = S::C3, NT = S::PRO_THREADS;
    extern __shared__ double smem[];
    double* XXs = smem;                // [a][7]        X, x, phi
    double* Pt = XXs + NPE * 7;        // [g][p][i]     traction operand  detJ J_x^-T [sigma | D]
    double* CHs = Pt + NG * C3;        // [g][pair][9]  staged so that the block writes one contiguous, coalesced run
    for (int64_t e = blockIdx.x; e < nelem; e += gridDim.x) {
        __syncthreads();
        for (int a = threadIdx.x; a < NPE; a += NT) {
            const int64_t n = conn[e * NPE + a];
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                XXs[a * 7 + l] = X[n * 3 + l];
                XXs[a * 7 + 3 + l] = x[n * 3 + l];
            }
            XXs[a * 7 + 6] = EL ? phi[n] : 0.0;
        }
        __syncthreads();
        const int g = threadIdx.x;
        if (g < NG) {
            double JX[9], Jx[9], gp[3] = {0, 0, 0};
#pragma unroll
            for (int i = 0; i < 9; ++i) JX[i] = Jx[i] = 0.0;
            for (int a = 0; a < NPE; ++a) {
                double j[3];
                // jm is [k][a][g] with g fastest: the lanes (consecutive Gauss points) read consecutive doubles; the node-fastest
                // twin jmT would make every lane touch its own sector (ncu: l1tex at 90 % with the pipes idle)
#pragma unroll
                for (int k = 0; k < 3; ++k) j[k] = jm[(k * NPE + a) * ldg + g];
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
            double* Cg = CHs + g * (S::NPAIR * 9);
            int pr = 0;
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = i; j < NV; ++j, ++pr) {   // only the dof pairs i <= j are contracted (K^{ji} is the mirror image)
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
                            Cg[pr * 9 + p * 3 + q] = v * detJ;
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
        // ---- Chat of the element: one contiguous run
        {
            double* dst = chg + (size_t)e * S::CH_SZ;
            for (int t = threadIdx.x; t < S::CH_SZ; t += NT) dst[t] = CHs[t];
        }
        // ---- traction t_a = sum_g Jm_g[:,a]^T P_g (only when the geometry is updated)
        for (int a = threadIdx.x; a < NPE; a += NT) {
            double t[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] = 0.0;
            if (update == 1) {
                for (int gg = 0; gg < NG; ++gg) {
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const double j = jmT[(gg * 3 + p) * NPE + a];
#pragma unroll
                        for (int i = 0; i < NV; ++i) t[i] = fma(j, Pt[gg * C3 + p * NV + i], t[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) te[(e * NPE + a) * NV + i] = t[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Kernel 2 of 2: K^{ij} = JmT * W^{ij} on the fp64 tensor cores, for the dof pairs i <= j.  Material-independent (only nvar
// matters).  Two blocks per SM: one block's Chat load / first W chunk overlaps the other block's DMMA loop.
#ifndef FL_IMMA_MINB
#define FL_IMMA_MINB 2
#endif
#ifndef FL_IMMA_KC64
#define FL_IMMA_KC64 12   // K chunk of 12 k-steps: 52 KB of W double buffer + 46 KB of Chat per block, two blocks per SM, 40 barriers per element
#endif
// small elements (tet20, hex27: 4 x 4 tiles, 34-60 KB of shared memory per block) hold few accumulators: three blocks per SM
template <int NV, int NPE, int NG, int KC, int NTB>
__global__ void __launch_bounds__(IMMA_THREADS, (NPE <= 32 ? 3 : FL_IMMA_MINB))
implicit_mma_gemm_kernel(const double* __restrict__ jmT, const double* __restrict__ chg, int64_t nelem, double* __restrict__ ke,
                         int plane_major, int sym_diag) {
    using S = imma_shape<NPE, NG, NV, KC, NTB>;
    // All eight warps issue DMMAs AND produce the next W chunk, a few items per k-step.  Round 2's first version kept round 1's
    // specialisation (warps 0-3 DMMA, warps 4-7 producers): ncu showed the producers' DMUL/DFMA stalled on `math pipe throttle`
    // behind the DMMAs of the other warps (fp64 FMAs and DMMA share the fp64 pipe) until a W chunk took as long as the DMMAs that
    // consume it, and the DMMA warps spent 28 % of their time at the step barrier (tensor pipe 67 % active).  With one role the
    // production is in every warp's own instruction stream and the barrier only closes a step all warps finish together.
#ifndef FL_IMMA_SPEC
#define FL_IMMA_SPEC 1    // measured on 24^3 hex64 EM_108 (profiles/r2_hex64_variants.md): specialised 9.4 ms, one role for all warps 10.2 ms
#endif
    constexpr bool SPEC = FL_IMMA_SPEC != 0;   // 1: warps 0-3 issue the DMMAs, warps 4-7 produce (A/B timing of the two schedules)
    constexpr int NWARP = SPEC ? 4 : IMMA_THREADS / 32;
    // two m-tiles per warp where the element allows it (hex64: 4 x 2 warps, hex125: 8 x 1): one B-fragment load feeds two DMMAs, and
    // the upper triangle of a symmetric K^{ii} can be dealt out evenly (m-tiles {wm, MT-1-wm}, n-tiles interleaved over wn)
    constexpr int WM = (S::MT / 2 >= NWARP) ? NWARP : (S::MT / 2 >= 1 ? S::MT / 2 : 1), WN = NWARP / WM;
    constexpr int MTW = S::MT / WM, NTW = S::NTB / WN;
    static_assert(S::MT % WM == 0 && S::NTB % WN == 0 && NWARP % WM == 0, "tiles must split evenly over the warps");
    constexpr int ndof = NPE * NV;
    constexpr int NB = S::NTB * 8;             // padded column count of a block
    constexpr int GPC = (KC * 4) / 3;          // Gauss points per chunk (KC*4 is a multiple of 3)
    static_assert((KC * 4) % 3 == 0, "a K chunk must hold whole Gauss points");
    constexpr int ITEMS = GPC * NB, U = (ITEMS + IMMA_THREADS - 1) / IMMA_THREADS;
    static_assert(SPEC || U <= KC, "one W item per thread and k-step must cover the chunk");
    extern __shared__ double smem[];
    double* CH = smem;                 // [g][pair][9]
    double* Ws = CH + S::CH_SZ;        // 2 x [KC*4][LDW] K chunks of one column block of W^{ij} (double buffer)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    // m-tiles of this warp: {wm, 2 WM - 1 - wm, 2 WM + wm, ...}: the upper-triangular tile counts of a symmetric K^{ii} then
    // add up to the same number for every warp
    auto mtile = [&](int m) { return (m & 1) ? WM * m + (WM - 1 - wm) : WM * m + wm; };

    for (int64_t e = blockIdx.x; e < nelem; e += gridDim.x) {
        __syncthreads();
        {
            const double2* src = reinterpret_cast<const double2*>(chg + (size_t)e * S::CH_SZ);
            double2* dst = reinterpret_cast<double2*>(CH);
            for (int t = threadIdx.x; t < S::CH_SZ / 2; t += IMMA_THREADS) dst[t] = src[t];
        }
        // ---- K^{ij}[:, column block] = JmT * W^{ij}[:, column block]: steps s = (ij, column block, K chunk)
        double* Ke = ke + (size_t)e * ndof * ndof;
        constexpr int NPAIR = S::NPAIR;
        constexpr int NS = NPAIR * S::NCB * S::NCH;
        auto pair_ij = [](int pr, int& i, int& j) {
            i = 0;
            int rem = pr;
            while (rem >= NV - i) { rem -= NV - i; ++i; }
            j = i + rem;
        };
        // one item = (Gauss point of the chunk, column node b): three rows (p = 0..2) share the three jmT loads.
        // b is the fastest index: coalesced jmT reads, Chat reads are warp-wide broadcasts.
        auto produce_item = [&](int s, int it) {
            if (it >= ITEMS) return;
            const int ch = s % S::NCH, cbij = s / S::NCH;
            const int cb = cbij % S::NCB, ij = cbij / S::NCB;
            double* Wb = Ws + (s & 1) * S::W_SZ;
            const int gl = it / NB, bl = it - gl * NB;
            const int g = ch * GPC + gl, b = cb * NB + bl;
            double w0 = 0.0, w1 = 0.0, w2 = 0.0;
            if (g < NG && b < NPE) {
                const double* jg = jmT + (g * 3) * NPE + b;
                const double j0 = jg[0], j1 = jg[NPE], j2 = jg[2 * NPE];
                const double* cg = CH + (g * NPAIR + ij) * 9;
                w0 = fma(cg[0], j0, fma(cg[1], j1, cg[2] * j2));
                w1 = fma(cg[3], j0, fma(cg[4], j1, cg[5] * j2));
                w2 = fma(cg[6], j0, fma(cg[7], j1, cg[8] * j2));
            }
            double* wo = Wb + (gl * 3) * S::LDW + bl;
            wo[0] = w0;
            wo[S::LDW] = w1;
            wo[2 * S::LDW] = w2;
        };
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; ++u) produce_item(0, threadIdx.x + u * IMMA_THREADS);
        __syncthreads();
        double c[MTW][NTW][2];
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
            const int ch = s % S::NCH, cbij = s / S::NCH;
            const int cb = cbij % S::NCB, ij = cbij / S::NCB;
            int i, j;
            pair_ij(ij, i, j);
            // K^{ii} is a symmetric matrix (Chat^{ii} has the major symmetry): with sym_diag, tiles strictly below the diagonal are
            // not computed and the upper triangle is mirrored -- exactly symmetric like the reference's fill (…Laplacian_.h:338-381)
            const bool symp = sym_diag && i == j;
            if (ch == 0) {
#pragma unroll
                for (int m = 0; m < MTW; ++m)
#pragma unroll
                    for (int q = 0; q < NTW; ++q) c[m][q][0] = c[m][q][1] = 0.0;
            }
            const double* Wb = Ws + (s & 1) * S::W_SZ;
            if (SPEC && warp >= 4) {
                if (s + 1 < NS) {
                    // producer warps: a lane takes column b = lane (+32, ...) of one Gauss point, so the nine Chat entries of the
                    // point are loaded once per lane and reused for NB/32 columns
                    const int s1 = s + 1;
                    const int ch1 = s1 % S::NCH, cbij1 = s1 / S::NCH;
                    const int cb1 = cbij1 % S::NCB, ij1 = cbij1 / S::NCB;
                    double* Wn = Ws + (s1 & 1) * S::W_SZ;
                    for (int gl = warp - 4; gl < GPC; gl += IMMA_THREADS / 32 - 4) {
                        const int g = ch1 * GPC + gl;
                        double cg[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) cg[k] = g < NG ? CH[(g * NPAIR + ij1) * 9 + k] : 0.0;
#pragma unroll
                        for (int bl = lane; bl < NB; bl += 32) {
                            const int b = cb1 * NB + bl;
                            double j0 = 0.0, j1 = 0.0, j2 = 0.0;
                            if (g < NG && b < NPE) {
                                const double* jg = jmT + (g * 3) * NPE + b;
                                j0 = jg[0]; j1 = jg[NPE]; j2 = jg[2 * NPE];
                            }
                            double* wo = Wn + (gl * 3) * S::LDW + bl;
#ifdef FL_IMMA_FAKEPROD   // timing experiment only: how fast is the step when W costs no fp64 work? (results are wrong)
                            wo[0] = j0; wo[S::LDW] = j1; wo[2 * S::LDW] = j2 + cg[0];
#else
                            wo[0] = fma(cg[0], j0, fma(cg[1], j1, cg[2] * j2));
                            wo[S::LDW] = fma(cg[3], j0, fma(cg[4], j1, cg[5] * j2));
                            wo[2 * S::LDW] = fma(cg[6], j0, fma(cg[7], j1, cg[8] * j2));
#endif
                        }
                    }
                }
                __syncthreads();
                continue;
            }
            // A fragments (slices of the shared JmT table, read through L1/L2) are fetched one k-step ahead: with ~200 KB of the SM's
            // 256 KB configured as shared memory little L1 is left, and an L2 round trip per k-step stalled the in-order DMMA stream
            auto load_a = [&](int ks, double (&af)[MTW]) {
                const int r = (ch * KC + ks) * 4 + lc;        // global k index of this lane's A/B element
                const int g = r / 3, p = r - 3 * g;
#pragma unroll
                for (int m = 0; m < MTW; ++m) {
                    const int a = 8 * mtile(m) + lr;
                    af[m] = (a < NPE && r < 3 * NG) ? jmT[(g * 3 + p) * NPE + a] : 0.0;
                }
            };
            double afr[2][MTW];
            load_a(0, afr[0]);
#pragma unroll
            for (int ks = 0; ks < KC; ++ks) {
                if (!SPEC && ks < U && s + 1 < NS) produce_item(s + 1, threadIdx.x + ks * IMMA_THREADS);
                if (ks + 1 < KC) load_a(ks + 1, afr[(ks + 1) & 1]);
#pragma unroll
                for (int q = 0; q < NTW; ++q) {
                    const double bfr = Wb[(4 * ks + lc) * S::LDW + 8 * (q * WN + wn) + lr];
#pragma unroll
                    for (int m = 0; m < MTW; ++m)
                        if (!(symp && mtile(m) > cb * S::NTB + q * WN + wn)) dmma884(c[m][q][0], c[m][q][1], afr[ks & 1][m], bfr);
                }
            }
            if (ch == S::NCH - 1) {
#pragma unroll
                for (int m = 0; m < MTW; ++m) {
                    const int mt = mtile(m);
                    const int a = 8 * mt + lr;
                    if (a < NPE) {
#pragma unroll
                        for (int q = 0; q < NTW; ++q) {
                            const int nt = cb * S::NTB + q * WN + wn;
                            if (symp && mt > nt) continue;
#pragma unroll
                            for (int z = 0; z < 2; ++z) {
                                const int b = 8 * nt + 2 * lc + z;
                                if (b < NPE && !(symp && a > b)) {
                                    const bool mirror = (i != j) || (symp && a != b);
                                    if (plane_major == 2) {
                                        // K_e scratch as per-row-node planes [a][(i,j)][b]: the same 64-byte fragment rows and
                                        // columns as below, and the NV*NV plane rows of one row node -- what one visit of the CSR
                                        // reduction reads -- are one contiguous run of NV*NV*NPE doubles
                                        Ke[((size_t)a * (NV * NV) + (i * NV + j)) * NPE + b] = c[m][q][z];
                                        if (mirror) Ke[((size_t)b * (NV * NV) + (j * NV + i)) * NPE + a] = c[m][q][z];
                                    } else if (plane_major) {
                                        // K_e scratch as dof-pair planes [(i,j)][a][b]: the 4 lanes of a fragment row write 64
                                        // contiguous bytes, and so do the 8 lanes of a fragment column in the mirror plane
                                        Ke[((size_t)(i * NV + j) * NPE + a) * NPE + b] = c[m][q][z];
                                        if (mirror) Ke[((size_t)(j * NV + i) * NPE + b) * NPE + a] = c[m][q][z];
                                    } else {
                                        Ke[(size_t)(a * NV + i) * ndof + b * NV + j] = c[m][q][z];
                                        if (mirror) Ke[(size_t)(b * NV + j) * ndof + a * NV + i] = c[m][q][z];
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

// Poisson (nvar = 1): Chat_g is Q_g = w |det J_X| J_X^-T e J_X^-1 (3 x 3), the only element-specific datum of
// K_ab = sum_g Jm_g[:,a]^T Q_g Jm_g[:,b] (_LowLevelAssemblyPerfectLaplacian_.h:246-420); the GEMM kernel above does the rest.
template <int NPE, int NG>
__global__ void __launch_bounds__(((NG + 31) / 32) * 32)
laplacian_mma_prologue_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ jm, int ldg,
                              const double* __restrict__ gw, int64_t nelem, const double* __restrict__ e_dev, double* __restrict__ chg) {
    constexpr int NT = ((NG + 31) / 32) * 32;
    __shared__ double Xs[NPE * 3];
    double et[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) et[i] = e_dev[i];
    for (int64_t e = blockIdx.x; e < nelem; e += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < NPE * 3; t += NT) {
            const int a = t / 3, l = t - 3 * a;
            Xs[t] = X[(int64_t)conn[e * NPE + a] * 3 + l];
        }
        __syncthreads();
        const int g = threadIdx.x;
        if (g < NG) {
            double JX[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) JX[i] = 0.0;
            for (int a = 0; a < NPE; ++a) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double j = jm[(k * NPE + a) * ldg + g];
#pragma unroll
                    for (int l = 0; l < 3; ++l) JX[k * 3 + l] += j * Xs[a * 3 + l];
                }
            }
            double iJX[9];
            const double detJ = gw[g] * fabs(invdet(JX, iJX));
            // Q[p][q] = detJ sum_kl iJX[k][p] e[k][l] iJX[l][q]   (same operation order as laplacian_elements_kernel)
            double tmp[9];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    double v = 0;
#pragma unroll
                    for (int l = 0; l < 3; ++l) v += et[k * 3 + l] * iJX[l * 3 + q];
                    tmp[k * 3 + q] = v;
                }
            double* Qo = chg + (size_t)e * imma_shape<NPE, NG, 1, 6>::CH_SZ + g * 9;
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    double v = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) v += iJX[k * 3 + p] * tmp[k * 3 + q];
                    Qo[p * 3 + q] = v * detJ;
                }
        }
    }
}

template <int NV, int NPE, int NG, int KC, int NTB>
int launch_mma_gemm(fl_handle* h, double* ke, int plane_major, int sym_diag, cudaStream_t st) {
    using S = imma_shape<NPE, NG, NV, KC, NTB>;
    auto kern = implicit_mma_gemm_kernel<NV, NPE, NG, KC, NTB>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IMMA_THREADS, S::SMEM));
    if (occ < 1) occ = 1;
    const int grid = (int)(h->nelem < (int64_t)occ * h->sm_count ? h->nelem : (int64_t)occ * h->sm_count);
    kern<<<grid, IMMA_THREADS, S::SMEM, st>>>(h->jmT, h->ch, h->nelem, ke, plane_major, sym_diag);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

template <int NPE, int NG, int KC, int NTB>
int launch_laplacian_mma(fl_handle* h, const double* e_dev, int symmetric, double* ke, cudaStream_t st) {
    using S = imma_shape<NPE, NG, 1, KC, NTB>;
    if (h->nelem == 0) return FL_OK;
    int rc = ensure_scratch(&h->ch, &h->ch_bytes, sizeof(double) * (size_t)h->nelem * S::CH_SZ);
    if (rc) return rc;
    constexpr int NT = ((NG + 31) / 32) * 32;
    const int64_t cap = (int64_t)h->sm_count * 32;
    laplacian_mma_prologue_kernel<NPE, NG><<<(int)(h->nelem < cap ? h->nelem : cap), NT, 0, st>>>(h->conn, h->points, h->jm, h->ldg, h->gw,
                                                                                                   h->nelem, e_dev, h->ch);
    FL_CUDA_CHECK(cudaGetLastError());
    return launch_mma_gemm<1, NPE, NG, KC, NTB>(h, ke, 0, symmetric ? 1 : 0, st);
}

template <int MAT, int NPE, int NG, int KC>
int launch_impl_mma(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                    cudaStream_t st) {
    constexpr int NV = 3 + (mat_traits<MAT>::electro ? 1 : 0);
    using S = imma_shape<NPE, NG, NV, KC>;
    if (h->nelem == 0) return FL_OK;
    int rc = ensure_scratch(&h->ch, &h->ch_bytes, sizeof(double) * (size_t)h->nelem * S::CH_SZ);
    if (rc) return rc;
    {
        auto pro = implicit_mma_prologue_kernel<MAT, NPE, NG>;
        FL_CUDA_CHECK(cudaFuncSetAttribute(pro, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::PRO_SMEM));
        int occ = 1;
        FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pro, S::PRO_THREADS, S::PRO_SMEM));
        if (occ < 1) occ = 1;
        const int64_t cap = (int64_t)occ * h->sm_count * 4;
        const int grid = (int)(h->nelem < cap ? h->nelem : cap);
        pro<<<grid, S::PRO_THREADS, S::PRO_SMEM, st>>>(h->conn, h->points, Eulerx, Eulerp, h->jm, h->ldg, h->jmT, h->gw, h->nelem, update, prm,
                                                       h->ch, te);
        FL_CUDA_CHECK(cudaGetLastError());
    }
    return launch_mma_gemm<NV, NPE, NG, KC, 0>(h, ke, h->ke_plane_major, 1, st);
}

}  // namespace fl
