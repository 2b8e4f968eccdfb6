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
#include "fl_internal.cuh"

namespace fl {

constexpr int IMPL_THREADS = 128;

template <int D, bool EL>
struct impl_dims {
    static constexpr int HS = voigt_map<D>::HS;
    static constexpr int HT = HS + (EL ? D : 0);
    static constexpr int NV = D + (EL ? 1 : 0);
    static constexpr int SSZ = D * D + (EL ? D : 0);  // sigma*detJ (+ D*detJ)
};

template <int D, int MAT, int A>
__global__ void __launch_bounds__(IMPL_THREADS)
implicit_elements_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                         const double* __restrict__ phi, const double* __restrict__ jm_g, const double* __restrict__ gw,
                         int64_t nelem, int npe, int ng, int ldg, int EB, int jm_in_smem, int update, MatParams prm,
                         double* __restrict__ ke, double* __restrict__ te) {
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr bool GEO = mat_traits<MAT>::geometric;
    using dims = impl_dims<D, EL>;
    constexpr int HT = dims::HT, NV = dims::NV, SSZ = dims::SSZ;
    extern __shared__ double smem[];
    const int xstride = (npe * D) | 1;
    double* jm_s = smem;
    double* Xs = jm_s + (jm_in_smem ? D * npe * ldg : 0);
    double* xs = Xs + EB * xstride;
    double* ph = xs + EB * xstride;
    double* iJ = ph + (EL ? EB * npe : 0);   // [el][g][D*D]  J_x^-1
    double* SG = iJ + EB * ng * D * D;       // [el][g][a][D] spatial gradients
    double* Hs = SG + EB * ng * npe * D;     // [el][g][HT*HT] hessian * detJ
    double* Ss = Hs + EB * ng * HT * HT;     // [el][g][SSZ]   sigma * detJ, D * detJ
    const double* jm = jm_in_smem ? jm_s : jm_g;
    const int ndof = npe * NV;
    const int nch = (npe + A - 1) / A;
    const int tpe = nch * ndof;

    if (jm_in_smem)
        for (int i = threadIdx.x; i < D * npe * ldg; i += blockDim.x) jm_s[i] = jm_g[i];

    const int64_t nbatch = (nelem + EB - 1) / EB;
    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int64_t e0 = batch * EB;
        const int ne = (int)min((int64_t)EB, nelem - e0);
        __syncthreads();
        for (int it = threadIdx.x; it < ne * npe; it += blockDim.x) {
            const int el = it / npe, a = it - el * npe;
            const int64_t n = conn[e0 * npe + it];
#pragma unroll
            for (int l = 0; l < D; ++l) {
                Xs[el * xstride + a * D + l] = X[n * D + l];
                xs[el * xstride + a * D + l] = x[n * D + l];
            }
            if (EL) ph[el * npe + a] = phi[n];
        }
        __syncthreads();
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
                        JX[k * D + l] += j[k] * Xa;
                        Jx[k * D + l] += j[k] * xa;
                    }
                }
                if (EL) {
                    const double p = ph[el * npe + a];
#pragma unroll
                    for (int k = 0; k < D; ++k) gp[k] += j[k] * p;
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
                    for (int k = 0; k < D; ++k) v += iJX[l * D + k] * Jx[k * D + i];
                    F[i * D + l] = v;
                }
            double E[D], Dv[D], sig[D * D], hess[HT * HT];
            if (EL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double v = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) v += iJx[k * D + jj] * gp[jj];
                    E[k] = -v;
                }
            }
            kinetic_measures<D, MAT, true>(F, E, prm, sig, Dv, hess);
            double* iJo = iJ + (el * ng + g) * D * D;
#pragma unroll
            for (int i = 0; i < D * D; ++i) iJo[i] = iJx[i];
            double* Ho = Hs + (el * ng + g) * HT * HT;
#pragma unroll
            for (int i = 0; i < HT * HT; ++i) Ho[i] = hess[i] * detJ;
            double* So = Ss + (el * ng + g) * SSZ;
#pragma unroll
            for (int i = 0; i < D * D; ++i) So[i] = sig[i] * detJ;
            if (EL) {
#pragma unroll
                for (int i = 0; i < D; ++i) So[D * D + i] = Dv[i] * detJ;
            }
        }
        __syncthreads();
        // ---- phase 2: spatial gradients grad_x N_a = J_x^-1 Jm_g[:,a]
        for (int it = threadIdx.x; it < ne * ng * npe; it += blockDim.x) {
            const int a = it % npe, eg = it / npe, g = eg % ng;
            const double* iJo = iJ + eg * D * D;
            double j[D];
#pragma unroll
            for (int k = 0; k < D; ++k) j[k] = jm[(k * npe + a) * ldg + g];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double v = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) v += iJo[k * D + jj] * j[jj];
                SG[it * D + k] = v;
            }
        }
        __syncthreads();
        // ---- phase 3: column (b,j) of K_e for a chunk of A row nodes
        for (int it = threadIdx.x; it < ne * tpe; it += blockDim.x) {
            const int el = it / tpe, r = it - el * tpe;
            const int chunk = r / ndof, col = r - chunk * ndof;
            const int b = col / NV, j = col - b * NV;
            const int a0 = chunk * A;
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
            double acc[A][NV];
#pragma unroll
            for (int aa = 0; aa < A; ++aa)
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[aa][i] = 0.0;
            for (int g = 0; g < ng; ++g) {
                const double* sg = SG + (size_t)(el * ng + g) * npe * D;
                const double* Hg = Hs + (el * ng + g) * HT * HT;
                const double w0 = sg[b * D + c0], w1 = sg[b * D + c1], w2 = (D == 3) ? sg[b * D + c2] : 0.0;
                double G[HT];
#pragma unroll
                for (int v = 0; v < HT; ++v) {
                    double t = Hg[v * HT + k0] * w0 + Hg[v * HT + k1] * w1;
                    if (D == 3) t += Hg[v * HT + k2] * w2;
                    G[v] = t;
                }
                double sb[D];
                if (GEO) {
                    // sigma*detJ applied to grad N_b, upper triangle of sigma as in _GeometricStiffness_.h:81-86
                    const double* So = Ss + (el * ng + g) * SSZ;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        double t = 0;
#pragma unroll
                        for (int l = 0; l < D; ++l) t += (k <= l ? So[k * D + l] : So[l * D + k]) * sg[b * D + l];
                        sb[k] = t;
                    }
                }
#pragma unroll
                for (int aa = 0; aa < A; ++aa) {
                    const int a = min(a0 + aa, npe - 1);
                    double ag[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) ag[k] = sg[a * D + k];
                    if (D == 3) {
                        acc[aa][0] += ag[0] * G[0] + ag[1] * G[3] + ag[2] * G[4];
                        acc[aa][1] += ag[1] * G[1] + ag[0] * G[3] + ag[2] * G[5];
                        acc[aa][2] += ag[2] * G[2] + ag[0] * G[4] + ag[1] * G[5];
                        if (EL) acc[aa][3] += ag[0] * G[6 % HT] + ag[1] * G[7 % HT] + ag[2] * G[8 % HT];
                    } else {
                        acc[aa][0] += ag[0] * G[0] + ag[1] * G[2];
                        acc[aa][1] += ag[1] * G[1] + ag[0] * G[2];
                        if (EL) acc[aa][2] += ag[0] * G[3 % HT] + ag[1] * G[4 % HT];
                    }
                    if (GEO) {
                        double dum = 0;
#pragma unroll
                        for (int k = 0; k < D; ++k) dum += ag[k] * sb[k];
#pragma unroll
                        for (int i = 0; i < D; ++i) acc[aa][i] += (i == j) ? dum : 0.0;
                    }
                }
            }
            double* Ke = ke + (size_t)(e0 + el) * ndof * ndof;
#pragma unroll
            for (int aa = 0; aa < A; ++aa) {
                const int a = a0 + aa;
                if (a < npe) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) Ke[(size_t)(a * NV + i) * ndof + col] = acc[aa][i];
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
                    const double* ag = SG + ((size_t)(el * ng + g) * npe + a) * D;
                    const double* So = Ss + (el * ng + g) * SSZ;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double v = 0;
#pragma unroll
                        for (int l = 0; l < D; ++l) v += ag[l] * (l <= i ? So[l * D + i] : So[i * D + l]);
                        t[i] += v;
                    }
                    if (EL) {
                        double v = 0;
#pragma unroll
                        for (int l = 0; l < D; ++l) v += ag[l] * So[D * D + l];
                        t[NV - 1] += v;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) te[(e0 * npe + it) * NV + i] = t[i];
        }
    }
}

template <int D, int MAT, int A>
int launch_impl_A(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                  cudaStream_t st) {
    constexpr bool EL = mat_traits<MAT>::electro;
    using dims = impl_dims<D, EL>;
    const int npe = h->npe, ng = h->ng, ldg = h->ldg;
    const int ndof = npe * dims::NV;
    const int nch = (npe + A - 1) / A;
    const int tpe = nch * ndof;
    const int xstride = (npe * D) | 1;
    const size_t jm_bytes = sizeof(double) * D * npe * ldg;
    const size_t per_elem = sizeof(double) * ((size_t)2 * xstride + (EL ? npe : 0) + (size_t)ng * D * D + (size_t)ng * npe * D +
                                              (size_t)ng * dims::HT * dims::HT + (size_t)ng * dims::SSZ);
    const size_t limit = (size_t)h->max_smem_optin;
    bool jm_in_smem = jm_bytes + per_elem <= limit && jm_bytes <= 64 * 1024;
    if ((jm_in_smem ? jm_bytes : 0) + per_elem > limit) {
        set_error("implicit kernel: one %d-node element needs %zu bytes of shared memory", npe, per_elem);
        return FL_ERR_UNSUPPORTED;
    }
    // enough elements per batch to occupy the block in phase 3, within half the shared memory so two blocks fit per SM
    int EB = (IMPL_THREADS + tpe - 1) / tpe;
    if (EB < 1) EB = 1;
    const size_t budget = limit / 2 > (jm_in_smem ? jm_bytes : 0) + per_elem ? limit / 2 : limit;
    while (EB > 1 && (jm_in_smem ? jm_bytes : 0) + per_elem * EB > budget) --EB;
    const size_t smem = (jm_in_smem ? jm_bytes : 0) + per_elem * EB;
    auto kern = implicit_elements_kernel<D, MAT, A>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, IMPL_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t nbatch = (h->nelem + EB - 1) / EB;
    const int grid = (int)(nbatch < (int64_t)occ * h->sm_count ? nbatch : (int64_t)occ * h->sm_count);
    if (grid == 0) return FL_OK;
    kern<<<grid, IMPL_THREADS, smem, st>>>(h->conn, h->points, Eulerx, Eulerp, h->jm, h->gw, h->nelem, npe, ng, ldg, EB,
                                           jm_in_smem ? 1 : 0, update, prm, ke, te);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// rows-per-thread chunk: whole element for tet10, 9 for hex27/quad9, 8 otherwise; electro uses 4-wide columns already
template <int D, int MAT>
int launch_impl_mat(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                    cudaStream_t st) {
    const int npe = h->npe;
    if (npe <= 4) return launch_impl_A<D, MAT, 4>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    if (npe == 10) return launch_impl_A<D, MAT, 10>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    if (npe % 9 == 0) return launch_impl_A<D, MAT, 9>(h, Eulerx, Eulerp, prm, update, ke, te, st);
    return launch_impl_A<D, MAT, 8>(h, Eulerx, Eulerp, prm, update, ke, te, st);
}

template <int MAT>
int launch_implicit_T(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, int update, double* ke, double* te,
                      cudaStream_t st) {
    return h->ndim == 3 ? launch_impl_mat<3, MAT>(h, Eulerx, Eulerp, prm, update, ke, te, st)
                        : launch_impl_mat<2, MAT>(h, Eulerx, Eulerp, prm, update, ke, te, st);
}

}  // namespace fl
