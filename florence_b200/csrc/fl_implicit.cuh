// fl_implicit.cuh -- element stiffness + traction kernel (implicit path), templated on <ndim, material, A>.
//
// Device replacement of the element loop body of _GlobalAssemblyDF_<Material> / _GlobalAssemblyDPF_<Material>
// (Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyDF_.h:69-131, _LowLevelAssemblyDPF_.h:73-150):
//   KinematicMeasures        (Florence/FiniteElements/LocalAssembly/_KinematicMeasures_/_KinematicMeasures_.h:64-119)
//   material _KineticMeasures_ (fl_math.cuh)
//   _ConstitutiveStiffnessIntegrandDF_Filler_ / ...DPF_  (Florence/VariationalPrinciple/_ConstitutiveStiffness_/
//                              _ConstitutiveStiffnessDF_.h:83-156, _ConstitutiveStiffnessDPF_.h:97-196)
//   _GeometricStiffnessFiller_ (Florence/VariationalPrinciple/_GeometricStiffness_/_GeometricStiffness_.h:61-144)
//
// The reference forms B (ndof x H) densely and calls dgemm twice per Gauss point (2H^2 ndof + 2H ndof^2 flops).  Here
// B's structure (3 non-zeros per column) is used directly: per Gauss point and column dof (b,j) a thread forms
// G = H_g B_b[:,j] (3H FMAs) once and then, for each row node a, the nvar entries B_a^T G (3 nvar FMAs), plus the
// geometric term on the diagonal block.  One thread owns column (b,j) of K_e for a chunk of A row nodes, so K_e rows are
// written as contiguous (coalesced) segments and nothing is recomputed.  fp64 accumulators live in registers.
#pragma once
#include "fl_implicit_mma.cuh"
#include "fl_implicit_warp.cuh"

namespace fl {

constexpr int IMPL_THREADS = 128;
#ifndef FL_AI
#define FL_AI 3
#endif
#ifndef FL_EB_DIV
#define FL_EB_DIV 5
#endif
#ifndef FL_MINB_ISO
#define FL_MINB_ISO 4
#endif
#ifndef FL_MINB_SPEC
#define FL_MINB_SPEC 4
#endif

template <int D, bool EL>
struct impl_dims {
    static constexpr int HS = voigt_map<D>::HS;
    static constexpr int HT = HS + (EL ? D : 0);
    static constexpr int NV = D + (EL ? 1 : 0);
    static constexpr int SSZ = D * D + (EL ? D : 0);  // sigma*detJ (+ D*detJ)
};

__host__ __device__ inline int odd_stride(int n) { return n | 1; }  // odd per-element strides: elements sharing a warp hit distinct banks

// Shared-memory carve-up, identical on host (sizing) and device.
template <int D, int MAT>
struct impl_layout {
    static constexpr bool EL = mat_traits<MAT>::electro;
    static constexpr bool CONST_H = (MAT == MAT_LINEAR_ELASTIC);  // tangent independent of F: one copy per block
    // isotropic constant tangent C = mu (I_ikjl + I_iljk) + lamb I_ijkl (_LinearElastic_.h:46-49): K_ab = lamb S + mu S^T + mu tr(S) I
    // with S_ab = sum_g detJ grad N_a (x) grad N_b, i.e. 9 FMAs per Gauss point and node pair instead of 27 + the G product
    static constexpr bool ISO = (MAT == MAT_LINEAR_ELASTIC);
    static constexpr int AI = FL_AI;  // row nodes per thread on the ISO path
    using dims = impl_dims<D, EL>;
    int xstride, ijs, sgs, hss, sss, djs;
    size_t jm_off, X_off, x_off, ph_off, iJ_off, SG_off, H_off, S_off, dJ_off, K_off, total;
    __host__ __device__ impl_layout(int npe, int ng, int ldg, int EB, bool jm_in_smem, bool stage = false) {
        xstride = odd_stride(npe * D);
        ijs = odd_stride(ng * D * D);
        sgs = odd_stride(ng * npe * D);
        hss = CONST_H ? 0 : odd_stride(ng * dims::HT * dims::HT);
        sss = odd_stride(ng * dims::SSZ);
        djs = odd_stride(ng);
        size_t o = 0;
        jm_off = o; o += jm_in_smem ? (size_t)D * npe * ldg : 0;
        X_off = o; o += (size_t)2 * EB * xstride;          // coordinates are double-buffered (cp.async prefetch of the next batch)
        x_off = o; o += (size_t)2 * EB * xstride;
        ph_off = o; o += EL ? (size_t)2 * EB * npe : 0;
        iJ_off = o; o += (size_t)EB * ijs;
        SG_off = o; o += (size_t)EB * sgs;
        H_off = o; o += CONST_H ? (size_t)dims::HT * dims::HT : (size_t)EB * hss;
        S_off = o; o += (size_t)EB * sss;
        dJ_off = o; o += (size_t)EB * djs;
        // K_e staging: the band mapping scatters a thread's blocks over rows, so K_e is assembled in shared memory and
        // streamed out as one contiguous run per batch
        o += (o & 1);  // 16-byte alignment of the staging tile (vectorised write-out)
        K_off = o; o += stage ? (size_t)EB * (npe * dims::NV) * (npe * dims::NV) : 0;
        total = o;
    }
};

// SYM = 1: K_e = K_e^T (the Voigt tangent is symmetric by construction, Numeric.pyx:181-229, and so are B^T H B and the
// geometric term), so a thread owning column (b,j) computes only the cyclic band of row nodes a = b, b+1, ..., b+npe/2 (mod npe)
// and writes each block and its mirror image: half the fp64 work, same K_e layout.  SYM = 0 computes all rows.
// NPE_T / NG_T > 0 fix nodes-per-element and Gauss-point counts at compile time for the hot element types (tet10, hex8,
// hex27): the Gauss loop unrolls and every shared-memory address becomes base + immediate.  MINB is the occupancy target.
template <int D, int MAT, int A, int SYM, int JM_SMEM, int NPE_T, int NG_T, int MINB, int STAGE>
__global__ void __launch_bounds__(IMPL_THREADS, MINB)
implicit_elements_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                         const double* __restrict__ phi, const double* __restrict__ jm_g, const double* __restrict__ gw,
                         int64_t nelem, int npe_rt, int ng_rt, int ldg_rt, int EB, int update, MatParams prm,
                         double* __restrict__ ke, double* __restrict__ te) {
    const int npe = NPE_T ? NPE_T : npe_rt;
    const int ng = NG_T ? NG_T : ng_rt;
    const int ldg = NG_T ? (NG_T | 1) : ldg_rt;
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr bool GEO = mat_traits<MAT>::geometric;
    using L = impl_layout<D, MAT>;
    constexpr bool CONST_H = L::CONST_H;
    using dims = impl_dims<D, EL>;
    constexpr int HT = dims::HT, NV = dims::NV, SSZ = dims::SSZ;
    extern __shared__ double smem[];
    const L lay(npe, ng, ldg, EB, JM_SMEM, STAGE);
    double* Kst = smem + lay.K_off;
    double* jm_s = smem + lay.jm_off;
    double* Xs0 = smem + lay.X_off;
    double* xs0 = smem + lay.x_off;
    double* ph0 = smem + lay.ph_off;
    double* iJ = smem + lay.iJ_off;   // [el][g][D*D]  J_x^-1
    double* SG = smem + lay.SG_off;   // [el][g][a][D] spatial gradients
    double* Hs = smem + lay.H_off;    // [el][g][HT*HT] hessian * detJ   (CONST_H: one unscaled copy)
    double* Ss = smem + lay.S_off;    // [el][g][SSZ]   sigma * detJ, D * detJ
    double* dJ = smem + lay.dJ_off;   // [el][g]        detJ
    const double* jm = JM_SMEM ? jm_s : jm_g;
    const int xstride = lay.xstride;
    const int ndof = npe * NV;
    const int half = npe / 2;
    const int nrows = SYM ? half + 1 : npe;   // row nodes per column
    constexpr bool ISO = L::ISO;
    constexpr int AI = L::AI;
    const int nch = ISO ? (nrows + AI - 1) / AI : (nrows + A - 1) / A;
    const int tpe = ISO ? nch * npe : nch * ndof;

    if (JM_SMEM)
        for (int i = threadIdx.x; i < D * npe * ldg; i += blockDim.x) jm_s[i] = jm_g[i];
    if (CONST_H) {
        // F-independent tangent (_LinearElastic_.h:46-51): evaluate once per block
        if (threadIdx.x == 0) {
            double F[D * D], sig[D * D], hess[HT * HT];
#pragma unroll
            for (int i = 0; i < D * D; ++i) F[i] = (i % (D + 1) == 0) ? 1.0 : 0.0;
            kinetic_measures<D, MAT, true>(F, nullptr, prm, sig, nullptr, hess);
#pragma unroll
            for (int i = 0; i < HT * HT; ++i) Hs[i] = hess[i];
        }
    }

    // asynchronous gather (cp.async, 8 bytes per value) of the nodal coordinates / potentials of one batch into buffer `buf`
    auto gather = [&](int64_t b0, int buf) {
        const int nb = (int)min((int64_t)EB, nelem - b0);
        double* Xb = Xs0 + buf * EB * xstride;
        double* xb = xs0 + buf * EB * xstride;
        double* pb = ph0 + buf * EB * npe;
        for (int it = threadIdx.x; it < nb * npe; it += blockDim.x) {
            const int el = it / npe, a = it - el * npe;
            const int64_t n = conn[b0 * npe + it];
            const unsigned sX = (unsigned)__cvta_generic_to_shared(Xb + el * xstride + a * D);
            const unsigned sx = (unsigned)__cvta_generic_to_shared(xb + el * xstride + a * D);
#pragma unroll
            for (int l = 0; l < D; ++l) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sX + 8 * l), "l"(X + n * D + l));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sx + 8 * l), "l"(x + n * D + l));
            }
            if (EL) {
                const unsigned sp = (unsigned)__cvta_generic_to_shared(pb + el * npe + a);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sp), "l"(phi + n));
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    const int64_t nbatch = (nelem + EB - 1) / EB;
    __syncthreads();
    if ((int64_t)blockIdx.x < nbatch) gather((int64_t)blockIdx.x * EB, 0);
    int buf = 0;
    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x, buf ^= 1) {
        const int64_t e0 = batch * EB;
        const int ne = (int)min((int64_t)EB, nelem - e0);
        double* Xs = Xs0 + buf * EB * xstride;
        double* xs = xs0 + buf * EB * xstride;
        double* ph = ph0 + buf * EB * npe;
        asm volatile("cp.async.wait_all;");
        __syncthreads();
        if (batch + gridDim.x < nbatch) gather((batch + gridDim.x) * EB, buf ^ 1);
        // ---- phase 1: kinematics and kinetics at (element, Gauss point)
        for (int it = threadIdx.x; it < ne * ng; it += blockDim.x) {
            const int el = it / ng, g = it - el * ng;
            double JX[D * D], Jx[D * D], gp[D];
#pragma unroll
            for (int i = 0; i < D * D; ++i) JX[i] = Jx[i] = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) gp[i] = 0.0;
            const double* Xe = Xs + el * xstride;
            const double* xe = xs + el * xstride;
            for (int a = 0; a < npe; ++a) {
                double j[D];
#pragma unroll
                for (int k = 0; k < D; ++k) j[k] = jm[(k * npe + a) * ldg + g];
#pragma unroll
                for (int l = 0; l < D; ++l) {
                    const double Xa = Xe[a * D + l], xa = xe[a * D + l];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        JX[k * D + l] = fma(j[k], Xa, JX[k * D + l]);
                        Jx[k * D + l] = fma(j[k], xa, Jx[k * D + l]);
                    }
                }
                if (EL) {
                    const double p = ph[el * npe + a];
#pragma unroll
                    for (int k = 0; k < D; ++k) gp[k] = fma(j[k], p, gp[k]);
                }
            }
            double iJX[D * D], iJx[D * D];
            const double detX = invdet(JX, iJX);
            const double detx = invdet(Jx, iJx);
            // _KinematicMeasures_.h:94-99
            const double detJ = gw[g] * fabs(update == 1 ? detx : detX);
            double F[D * D];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int l = 0; l < D; ++l) {
                    double v = 0;
#pragma unroll
                    for (int k = 0; k < D; ++k) v = fma(iJX[l * D + k], Jx[k * D + i], v);
                    F[i * D + l] = v;
                }
            double E[D], Dv[D], sig[D * D], hess[HT * HT];
            if (EL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double v = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) v = fma(iJx[k * D + jj], gp[jj], v);
                    E[k] = -v;
                }
            }
            kinetic_measures<D, MAT, !CONST_H>(F, E, prm, sig, Dv, hess);
            double* iJo = iJ + el * lay.ijs + g * D * D;
#pragma unroll
            for (int i = 0; i < D * D; ++i) iJo[i] = iJx[i];
            if (!CONST_H) {
                double* Ho = Hs + el * lay.hss + g * HT * HT;
#pragma unroll
                for (int i = 0; i < HT * HT; ++i) Ho[i] = hess[i] * detJ;
            }
            double* So = Ss + el * lay.sss + g * SSZ;
#pragma unroll
            for (int i = 0; i < D * D; ++i) So[i] = sig[i] * detJ;
            if (EL) {
#pragma unroll
                for (int i = 0; i < D; ++i) So[D * D + i] = Dv[i] * detJ;
            }
            dJ[el * lay.djs + g] = detJ;
        }
        __syncthreads();
        // ---- phase 2: spatial gradients grad_x N_a = J_x^-1 Jm_g[:,a]
        for (int it = threadIdx.x; it < ne * ng * npe; it += blockDim.x) {
            const int a = it % npe, eg = it / npe, g = eg % ng, el = eg / ng;
            const double* iJo = iJ + el * lay.ijs + g * D * D;
            double j[D];
#pragma unroll
            for (int k = 0; k < D; ++k) j[k] = jm[(k * npe + a) * ldg + g];
            double* sgo = SG + el * lay.sgs + (g * npe + a) * D;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double v = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) v = fma(iJo[k * D + jj], j[jj], v);
                sgo[k] = v;
            }
        }
        // the previous batch's bulk copy must have read the staging tile before phase 3 overwrites it
        if (STAGE && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        // ---- phase 3 (isotropic constant tangent): S_ab = sum_g detJ grad N_a (x) grad N_b for column node b and AI row nodes
        if constexpr (ISO) {
            for (int it = threadIdx.x; it < ne * tpe; it += blockDim.x) {
                const int el = it / tpe, r = it - el * tpe;
                const int chunk = r / npe, b = r - chunk * npe;
                const int r0 = chunk * AI;
                int arow[AI];
#pragma unroll
                for (int aa = 0; aa < AI; ++aa) {
                    int a = min(r0 + aa, nrows - 1) + (SYM ? b : 0);
                    if (a >= npe) a -= npe;
                    arow[aa] = a * D;
                }
                double S[AI][D][D];
#pragma unroll
                for (int aa = 0; aa < AI; ++aa)
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = 0; j < D; ++j) S[aa][i][j] = 0.0;
                const double* sge = SG + el * lay.sgs;
#pragma unroll(NG_T > 0 && NG_T <= 8 ? NG_T : 1)
                for (int g = 0; g < ng; ++g) {
                    const double* sg = sge + g * npe * D;
                    const double d = dJ[el * lay.djs + g];
                    double bg[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) bg[k] = sg[b * D + k] * d;
#pragma unroll
                    for (int aa = 0; aa < AI; ++aa) {
                        const double* ap = sg + arow[aa];
#pragma unroll
                        for (int i = 0; i < D; ++i) {
                            const double ai = ap[i];
#pragma unroll
                            for (int j = 0; j < D; ++j) S[aa][i][j] = fma(ai, bg[j], S[aa][i][j]);
                        }
                    }
                }
                double* Ke = STAGE ? Kst + (size_t)el * ndof * ndof : ke + (size_t)(e0 + el) * ndof * ndof;
#pragma unroll
                for (int aa = 0; aa < AI; ++aa) {
                    const bool dup = SYM && ((npe & 1) == 0) && (r0 + aa == half) && (b >= half);
                    if (r0 + aa < nrows && !dup) {
                        const int a = arow[aa] / D;
                        double tr = 0;
#pragma unroll
                        for (int i = 0; i < D; ++i) tr += S[aa][i][i];
                        double Kb[D][D];
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = 0; j < D; ++j)
                                Kb[i][j] = __dadd_rn(fma(prm.lamb, S[aa][i][j], __dmul_rn(prm.mu, S[aa][j][i])), i == j ? __dmul_rn(prm.mu, tr) : 0.0);
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = 0; j < D; ++j) Ke[(size_t)(a * NV + i) * ndof + b * NV + j] = Kb[i][j];
                        if (SYM && a != b) {
#pragma unroll
                            for (int j = 0; j < D; ++j)
#pragma unroll
                                for (int i = 0; i < D; ++i) Ke[(size_t)(b * NV + j) * ndof + a * NV + i] = Kb[i][j];
                        }
                    }
                }
            }
        } else
        // ---- phase 3: column (b,j) of K_e for a chunk of A row nodes
        for (int it = threadIdx.x; it < ne * tpe; it += blockDim.x) {
            const int el = it / tpe, r = it - el * tpe;
            const int chunk = r / ndof, col = r - chunk * ndof;
            const int b = col / NV, j = col - b * NV;
            const int r0 = chunk * A;
            // non-zero rows of column j of B_b and which gradient component sits there
            // (FillConstitutiveB_, _ConstitutiveStiffnessDF_.h:32-77, ...DPF_.h:42-96)
            int k0, k1, k2, c0, c1, c2;
            if (D == 3) {
                if (j == 0) { k0 = 0; k1 = 3; k2 = 4; c0 = 0; c1 = 1; c2 = 2; }
                else if (j == 1) { k0 = 1; k1 = 3; k2 = 5; c0 = 1; c1 = 0; c2 = 2; }
                else if (j == 2) { k0 = 2; k1 = 4; k2 = 5; c0 = 2; c1 = 0; c2 = 1; }
                else { k0 = 6; k1 = 7; k2 = 8; c0 = 0; c1 = 1; c2 = 2; }
            } else {
                if (j == 0) { k0 = 0; k1 = 2; c0 = 0; c1 = 1; }
                else if (j == 1) { k0 = 1; k1 = 2; c0 = 1; c1 = 0; }
                else { k0 = 3; k1 = 4; c0 = 0; c1 = 1; }
                k2 = 0; c2 = 0;
            }
            // row nodes of this thread: a = b + r0 + aa (cyclic) when SYM, else a = r0 + aa
            int arow[A];
#pragma unroll
            for (int aa = 0; aa < A; ++aa) {
                int a = min(r0 + aa, nrows - 1) + (SYM ? b : 0);
                if (a >= npe) a -= npe;
                arow[aa] = a * D;
            }
            double acc[A][NV];
#pragma unroll
            for (int aa = 0; aa < A; ++aa)
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[aa][i] = 0.0;
            const double* sge = SG + el * lay.sgs;
#pragma unroll(NG_T > 0 && NG_T <= 8 ? NG_T : 1)
            for (int g = 0; g < ng; ++g) {
                const double* sg = sge + g * npe * D;
                const double* Hg = CONST_H ? Hs : Hs + el * lay.hss + g * HT * HT;
                double w0 = sg[b * D + c0], w1 = sg[b * D + c1], w2 = (D == 3) ? sg[b * D + c2] : 0.0;
                if (CONST_H) {
                    const double d = dJ[el * lay.djs + g];
                    w0 *= d; w1 *= d; w2 *= d;
                }
                double G[HT];
#pragma unroll
                for (int v = 0; v < HT; ++v) {
                    double t = Hg[v * HT + k0] * w0;
                    t = fma(Hg[v * HT + k1], w1, t);
                    if (D == 3) t = fma(Hg[v * HT + k2], w2, t);
                    G[v] = t;
                }
                double sb[D];
                if (GEO) {
                    // sigma*detJ applied to grad N_b, upper triangle of sigma as in _GeometricStiffness_.h:81-86;
                    // only the diagonal component i == j receives it, so fold the selection into sb
                    const double* So = Ss + el * lay.sss + g * SSZ;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        double t = 0;
#pragma unroll
                        for (int l = 0; l < D; ++l) t = fma(k <= l ? So[k * D + l] : So[l * D + k], sg[b * D + l], t);
                        sb[k] = t;
                    }
                }
#pragma unroll
                for (int aa = 0; aa < A; ++aa) {
                    const double* ap = sg + arow[aa];
                    double ag[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) ag[k] = ap[k];
                    if (D == 3) {
                        acc[aa][0] = fma(ag[0], G[0], fma(ag[1], G[3], fma(ag[2], G[4], acc[aa][0])));
                        acc[aa][1] = fma(ag[1], G[1], fma(ag[0], G[3], fma(ag[2], G[5], acc[aa][1])));
                        acc[aa][2] = fma(ag[2], G[2], fma(ag[0], G[4], fma(ag[1], G[5], acc[aa][2])));
                        if (EL) acc[aa][NV - 1] = fma(ag[0], G[6 % HT], fma(ag[1], G[7 % HT], fma(ag[2], G[8 % HT], acc[aa][NV - 1])));
                    } else {
                        acc[aa][0] = fma(ag[0], G[0], fma(ag[1], G[2], acc[aa][0]));
                        acc[aa][1] = fma(ag[1], G[1], fma(ag[0], G[2], acc[aa][1]));
                        if (EL) acc[aa][NV - 1] = fma(ag[0], G[3 % HT], fma(ag[1], G[4 % HT], acc[aa][NV - 1]));
                    }
                    if (GEO) {
                        double dum = 0;
#pragma unroll
                        for (int k = 0; k < D; ++k) dum = fma(ag[k], sb[k], dum);
#pragma unroll
                        for (int i = 0; i < D; ++i) acc[aa][i] += (i == j) ? dum : 0.0;
                    }
                }
            }
            double* Ke = STAGE ? Kst + (size_t)el * ndof * ndof : ke + (size_t)(e0 + el) * ndof * ndof;
#pragma unroll
            for (int aa = 0; aa < A; ++aa) {
                // even npe: the last band row pairs b with b+npe/2, which both columns reach; only the lower one writes it
                const bool dup = SYM && ((npe & 1) == 0) && (r0 + aa == half) && (b >= half);
                if (r0 + aa < nrows && !dup) {
                    const int a = arow[aa] / D;
#pragma unroll
                    for (int i = 0; i < NV; ++i) Ke[(size_t)(a * NV + i) * ndof + col] = acc[aa][i];
                    if (SYM && a != b) {
                        // mirror image K[(b,j),(a,i)] = K[(a,i),(b,j)]
#pragma unroll
                        for (int i = 0; i < NV; ++i) Ke[(size_t)col * ndof + a * NV + i] = acc[aa][i];
                    }
                }
            }
        }
        // ---- phase 4: traction t_a = sum_g B_a^T [sigma; D] detJ  (only when the geometry is updated, :136-148)
        for (int it = threadIdx.x; it < ne * npe; it += blockDim.x) {
            const int el = it / npe, a = it - el * npe;
            double t[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] = 0.0;
            if (update == 1) {
                for (int g = 0; g < ng; ++g) {
                    const double* ag = SG + el * lay.sgs + (g * npe + a) * D;
                    const double* So = Ss + el * lay.sss + g * SSZ;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
#pragma unroll
                        for (int l = 0; l < D; ++l) t[i] = fma(ag[l], l <= i ? So[l * D + i] : So[i * D + l], t[i]);
                    }
                    if (EL) {
#pragma unroll
                        for (int l = 0; l < D; ++l) t[NV - 1] = fma(ag[l], So[D * D + l], t[NV - 1]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) te[(e0 * npe + it) * NV + i] = t[i];
        }
        if (STAGE) {
            // ---- phase 5: the staged element matrices of this batch are one contiguous run in K_e: hand them to the TMA engine
            //      (one bulk shared -> global copy issued by thread 0) so that neither the tile read nor the global stores pass
            //      through the LSU data pipe, which bounds this kernel.  Odd-sized / misaligned runs take the 16-byte store loop.
            double* dst = ke + (size_t)e0 * ndof * ndof;
            const int tot = ne * ndof * ndof;
            const bool bulk = ((((size_t)dst) & 15) == 0) && ((tot & 1) == 0);
            if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (bulk) {
                if (threadIdx.x == 0) {
                    const unsigned src = (unsigned)__cvta_generic_to_shared(Kst);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(tot * 8) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if ((((size_t)dst) & 15) == 0) {
                double2* d2 = reinterpret_cast<double2*>(dst);
                const double2* s2 = reinterpret_cast<const double2*>(Kst);
                for (int t = threadIdx.x; t < tot / 2; t += blockDim.x) d2[t] = s2[t];
                if ((tot & 1) && threadIdx.x == 0) dst[tot - 1] = Kst[tot - 1];
            } else {
                for (int t = threadIdx.x; t < tot; t += blockDim.x) dst[t] = Kst[t];
            }
        }
    }
    if (STAGE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int D, int MAT, int A, int SYM, int JM_SMEM, int NPE_T, int NG_T, int MINB, int STAGE>
int launch_impl_cfg(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                    cudaStream_t st, int EB, size_t smem) {
    auto kern = implicit_elements_kernel<D, MAT, A, SYM, JM_SMEM, NPE_T, NG_T, MINB, STAGE>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IMPL_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t nbatch = (h->nelem + EB - 1) / EB;
    const int grid = (int)(nbatch < (int64_t)occ * h->sm_count ? nbatch : (int64_t)occ * h->sm_count);
    if (grid == 0) return FL_OK;
    kern<<<grid, IMPL_THREADS, smem, st>>>(h->conn, h->points, Eulerx, Eulerp, h->jm, h->gw, h->nelem, h->npe, h->ng, h->ldg, EB, update,
                                           prm, ke, te);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

template <int D, int MAT, int A>
int launch_impl_A(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                  cudaStream_t st) {
    constexpr bool EL = mat_traits<MAT>::electro;
    using dims = impl_dims<D, EL>;
    using L = impl_layout<D, MAT>;
    const int npe = h->npe, ng = h->ng, ldg = h->ldg;
    const int ndof = npe * dims::NV;
    const int nrows = npe / 2 + 1;
    const int tpe = L::ISO ? ((nrows + L::AI - 1) / L::AI) * npe : ((nrows + A - 1) / A) * ndof;
    const size_t limit = (size_t)h->max_smem_optin;
    const size_t jm_bytes = sizeof(double) * D * npe * ldg;
    bool jm_in_smem = jm_bytes <= 64 * 1024 && sizeof(double) * L(npe, ng, ldg, 1, true).total <= limit;
    if (sizeof(double) * L(npe, ng, ldg, 1, jm_in_smem).total > limit) {
        set_error("implicit kernel: one %d-node element needs %zu bytes of shared memory", npe, sizeof(double) * L(npe, ng, ldg, 1, false).total);
        return FL_ERR_UNSUPPORTED;
    }
    // stage K_e in shared memory when one element's matrix plus its Gauss data fits in half of the SM's shared memory
    const bool stage = jm_in_smem && sizeof(double) * L(npe, ng, ldg, 1, true, true).total <= limit / 2;
    // two phase-3 rounds worth of elements per batch (keeps phases 1-2 populated), within a fifth of the SM's shared memory
    int EB = (2 * IMPL_THREADS) / tpe;
    if (EB < 1) EB = 1;
    while (EB > 1 && sizeof(double) * L(npe, ng, ldg, EB, jm_in_smem, stage).total > limit / FL_EB_DIV) --EB;
    const size_t smem = sizeof(double) * L(npe, ng, ldg, EB, jm_in_smem, stage).total;
    // compile-time element shapes of the benchmark configs (mechanics, 3-D): tet10 (8 gp), hex8, hex27
    if constexpr (D == 3 && !EL && A == 6) {
        if (stage && npe == 10 && ng == 8) return launch_impl_cfg<D, MAT, A, 1, 1, 10, 8, (L::ISO ? FL_MINB_ISO : FL_MINB_SPEC), 1>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem);
        if (stage && npe == 8 && ng == 8) return launch_impl_cfg<D, MAT, A, 1, 1, 8, 8, FL_MINB_SPEC, 1>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem);
    }
    if constexpr (D == 3 && !EL && A == 7) {
        if (stage && npe == 27 && ng == 27) return launch_impl_cfg<D, MAT, A, 1, 1, 27, 27, 4, 1>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem);
    }
    if (stage) return launch_impl_cfg<D, MAT, A, 1, 1, 0, 0, 4, 1>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem);
    return jm_in_smem ? launch_impl_cfg<D, MAT, A, 1, 1, 0, 0, 4, 0>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem)
                      : launch_impl_cfg<D, MAT, A, 1, 0, 0, 0, 4, 0>(h, Eulerx, Eulerp, prm, update, ke, te, st, EB, smem);
}

// rows-per-thread chunk A over the band of npe/2+1 row nodes: 6 covers tet10 / tri6 / quad9 / hex8 in one pass, 7 splits hex27's
// 14 rows in two, 11 splits hex64's 33 rows in three
template <int D, int MAT>
int launch_impl_mat(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                    cudaStream_t st) {
    if constexpr (D == 3) {
        // p = 3 hexahedra: dense parent-space contraction on the fp64 tensor cores (fl_implicit_mma.cuh); hex27 only when
        // forced (option value 2): its 27 -> 32 tile padding makes the generic kernel the faster one
        if (h->use_mma_implicit == 2 && h->npe == 27 && h->ng == 27) return launch_impl_mma<MAT, 27, 27, 21>(h, Eulerx, Eulerp, prm, update, ke, te, st);
        if (h->use_mma_implicit && h->npe == 64 && h->ng == 64) return launch_impl_mma<MAT, 64, 64, FL_IMMA_KC64>(h, Eulerx, Eulerp, prm, update, ke, te, st);
        // p = 3 tetrahedra (tet20, the reference's 14-point rule): K = 42 -> one chunk of 12 k-steps, 20 -> 32 node padding.
        // Default for the electro-mechanical models (82 944 elements, EM_108: 5.08 ms against 6.76 ms for the generic kernel); for
        // mechanics the padding costs more than the tensor pipe gains (2.85 against 2.47 ms, NeoHookean): option value 2 only
        if (h->use_mma_implicit && (mat_traits<MAT>::electro || h->use_mma_implicit == 2) && h->npe == 20 && h->ng == 14)
            return launch_impl_mma<MAT, 20, 14, 12>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    }
    if constexpr (D == 3 && MAT == MAT_LINEAR_ELASTIC) {
        // tet10 / hex8 (8 Gauss points): warp-autonomous kernel, no block barriers, no staging tile (fl_implicit_warp.cuh)
        const bool warp_ok = h->use_warp_iso && h->ng == 8 && (((size_t)ke) & 31) == 0;   // 32-byte stores
        if (warp_ok && h->npe == 10) return launch_impl_iso_warp<10, 8>(h, Eulerx, prm, update, ke, te, st);
        // hex8: measured slightly slower than the block-wide kernel with its TMA write-out (1.80 vs 1.68 ms per 1 M elements): opt-in
        if (warp_ok && h->use_warp_iso == 2 && h->npe == 8) return launch_impl_iso_warp<8, 8>(h, Eulerx, prm, update, ke, te, st);
    }
    const int rows = h->npe / 2 + 1;
    if (rows <= 3) return launch_impl_A<D, MAT, 3>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    if (rows <= 6) return launch_impl_A<D, MAT, 6>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    if (rows % 7 == 0 || rows > 44) return launch_impl_A<D, MAT, 7>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    return launch_impl_A<D, MAT, 11>(h, Eulerx, Eulerp, prm, update, ke, te, st);
}

template <int MAT>
int launch_implicit_T(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                      cudaStream_t st) {
    return h->ndim == 3 ? launch_impl_mat<3, MAT>(h, Eulerx, Eulerp, prm, update, ke, te, st)
                        : launch_impl_mat<2, MAT>(h, Eulerx, Eulerp, prm, update, ke, te, st);
}

}  // namespace fl
