// fl_implicit.cu -- dispatch of the implicit element kernels, Poisson stiffness and mass kernels.
//
// Compiled several times: -DFL_IMPL_PART=<k> selects which material instantiations of fl_implicit.cuh this object holds
// (they are heavy fp64 kernels; the parts build in parallel).  Part 0 also holds the dispatcher, the Laplacian kernel
// (_GlobalAssemblyPerfectLaplacian_, Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyPerfectLaplacian_.h:246-420)
// and the mass kernels (_GenericConstantMassIntegrand_, Florence/VariationalPrinciple/_Mass_/_MassIntegrand_.h:249-395).
#include "fl_implicit.cuh"

#ifndef FL_IMPL_PART
#define FL_IMPL_PART 0
#endif

namespace fl {

#define FL_INST(MATID)                                                                                                               \
    template int launch_implicit_T<MATID>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);

#if FL_IMPL_PART == 1
FL_INST(MAT_LINEAR_ELASTIC)
FL_INST(MAT_NEOHOOKEAN)
#elif FL_IMPL_PART == 2
FL_INST(MAT_MOONEY_RIVLIN)
#elif FL_IMPL_PART == 3
FL_INST(MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN)
#elif FL_IMPL_PART == 4
FL_INST(MAT_ELECTRO_101)
#elif FL_IMPL_PART == 5
FL_INST(MAT_ELECTRO_105)
#elif FL_IMPL_PART == 6
FL_INST(MAT_ELECTRO_108)
#endif

#if FL_IMPL_PART == 0

extern template int launch_implicit_T<MAT_LINEAR_ELASTIC>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_NEOHOOKEAN>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_MOONEY_RIVLIN>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_ELECTRO_101>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_ELECTRO_105>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);
extern template int launch_implicit_T<MAT_ELECTRO_108>(fl_handle*, const double*, const double*, const MatParams&, int, double*, double*, cudaStream_t);

int launch_implicit_elements(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation,
                             int update, double* ke, double* te, cudaStream_t st) {
    MatParams p;
    p.mu = mat->mu; p.mu1 = mat->mu1; p.mu2 = mat->mu2; p.mu3 = mat->mu3; p.mue = mat->mue; p.lamb = mat->lamb;
    p.eps_1 = mat->eps_1; p.eps_2 = mat->eps_2; p.eps_3 = mat->eps_3; p.eps_e = mat->eps_e;
    const int m = mat->material_number;
    const bool electro = (m == MAT_ELECTRO_101 || m == MAT_ELECTRO_105 || m == MAT_ELECTRO_108);
    if (electro != (formulation == 1)) {
        set_error("material %d does not match formulation_number %d", m, formulation);
        return FL_ERR_INVALID;
    }
    if (electro && !Eulerp) {
        set_error("Eulerp is required for electro-mechanical materials");
        return FL_ERR_INVALID;
    }
    switch (m) {
        case MAT_LINEAR_ELASTIC: return launch_implicit_T<MAT_LINEAR_ELASTIC>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_NEOHOOKEAN: return launch_implicit_T<MAT_NEOHOOKEAN>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_MOONEY_RIVLIN: return launch_implicit_T<MAT_MOONEY_RIVLIN>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN:
            return launch_implicit_T<MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_ELECTRO_101: return launch_implicit_T<MAT_ELECTRO_101>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_ELECTRO_105: return launch_implicit_T<MAT_ELECTRO_105>(h, Eulerx, Eulerp, p, update, ke, te, st);
        case MAT_ELECTRO_108: return launch_implicit_T<MAT_ELECTRO_108>(h, Eulerx, Eulerp, p, update, ke, te, st);
        default:
            // _LowLevelAssembly_.py:58-60: no "_LowLevelAssemblyD(P)F__<Material>_" for this material
            set_error("Turning optimise option on for material number %d is not supported yet", m);
            return FL_ERR_UNSUPPORTED;
    }
}

// ------------------------------------------------------------------------------------------------ Laplacian
// K_ab = sum_g w|det J_X| grad_X N_a . e . grad_X N_b  =  sum_g Jm_g[:,a]^T Q_g Jm_g[:,b],  Q_g = w|det J_X| J_X^-T e J_X^-1:
// the element-specific data shrink to one d x d matrix per Gauss point; both operands of the contraction are the shared table.
constexpr int LAP_THREADS = 128;
template <int D, int A>
__global__ void __launch_bounds__(LAP_THREADS)
laplacian_elements_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ jm_g,
                          const double* __restrict__ gw, int64_t nelem, int npe, int ng, int ldg, int EB, int jm_in_smem,
                          const double* __restrict__ e_dev, int symmetric, double* __restrict__ ke) {
    extern __shared__ double smem[];
    const int xstride = (npe * D) | 1;
    double* jm_s = smem;
    double* Xs = jm_s + (jm_in_smem ? D * npe * ldg : 0);
    double* Q = Xs + EB * xstride;  // [el][g][D*D]
    const double* jm = jm_in_smem ? jm_s : jm_g;
    const int nch = (npe + A - 1) / A;
    const int tpe = nch * npe;
    double et[D * D];
#pragma unroll
    for (int i = 0; i < D * D; ++i) et[i] = e_dev[i];
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
            for (int l = 0; l < D; ++l) Xs[el * xstride + a * D + l] = X[n * D + l];
        }
        __syncthreads();
        for (int it = threadIdx.x; it < ne * ng; it += blockDim.x) {
            const int el = it / ng, g = it - el * ng;
            double JX[D * D];
#pragma unroll
            for (int i = 0; i < D * D; ++i) JX[i] = 0.0;
            const double* Xe = Xs + el * xstride;
            for (int a = 0; a < npe; ++a) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double j = jm[(k * npe + a) * ldg + g];
#pragma unroll
                    for (int l = 0; l < D; ++l) JX[k * D + l] += j * Xe[a * D + l];
                }
            }
            double iJX[D * D];
            const double detJ = gw[g] * fabs(invdet(JX, iJX));
            // Q[p][q] = detJ sum_kl iJX[k][p] e[k][l] iJX[l][q]
            double tmp[D * D];
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int q = 0; q < D; ++q) {
                    double v = 0;
#pragma unroll
                    for (int l = 0; l < D; ++l) v += et[k * D + l] * iJX[l * D + q];
                    tmp[k * D + q] = v;
                }
            double* Qo = Q + (el * ng + g) * D * D;
#pragma unroll
            for (int p = 0; p < D; ++p)
#pragma unroll
                for (int q = 0; q < D; ++q) {
                    double v = 0;
#pragma unroll
                    for (int k = 0; k < D; ++k) v += iJX[k * D + p] * tmp[k * D + q];
                    Qo[p * D + q] = v * detJ;
                }
        }
        __syncthreads();
        for (int it = threadIdx.x; it < ne * tpe; it += blockDim.x) {
            const int el = it / tpe, r = it - el * tpe;
            const int chunk = r / npe, b = r - chunk * npe;
            const int a0 = chunk * A;
            // symmetric Hessian, high order: only row chunks on or above the diagonal are needed (the mirror image is written
            // below); for small elements the idle lanes cost more than the skipped work (measured: hex125 14.3 -> 9.0 ms, hex27 slower)
            if (symmetric && npe >= 64 && a0 > b) continue;
            double acc[A];
#pragma unroll
            for (int aa = 0; aa < A; ++aa) acc[aa] = 0.0;
            for (int g = 0; g < ng; ++g) {
                const double* Qg = Q + (el * ng + g) * D * D;
                double jb[D], qb[D];
#pragma unroll
                for (int q = 0; q < D; ++q) jb[q] = jm[(q * npe + b) * ldg + g];
#pragma unroll
                for (int p = 0; p < D; ++p) {
                    double v = 0;
#pragma unroll
                    for (int q = 0; q < D; ++q) v += Qg[p * D + q] * jb[q];
                    qb[p] = v;
                }
#pragma unroll
                for (int aa = 0; aa < A; ++aa) {
                    const int a = min(a0 + aa, npe - 1);
#pragma unroll
                    for (int p = 0; p < D; ++p) acc[aa] += jm[(p * npe + a) * ldg + g] * qb[p];
                }
            }
            double* Ke = ke + (size_t)(e0 + el) * npe * npe;
#pragma unroll
            for (int aa = 0; aa < A; ++aa) {
                const int a = a0 + aa;
                // symmetric Hessian: the reference fills the upper triangle and mirrors it (:338-381)
                if (a < npe && (!symmetric || a <= b)) {
                    Ke[(size_t)a * npe + b] = acc[aa];
                    if (symmetric && a < b) Ke[(size_t)b * npe + a] = acc[aa];
                }
            }
        }
    }
}

int launch_laplacian_elements(fl_handle* h, const double* e_dev, int symmetric, double* ke, cudaStream_t st) {
    const int npe = h->npe, ng = h->ng, ldg = h->ldg, D = h->ndim;
    // p >= 3 hexahedra: K = JmT (Q_g Jm) is the shared-operand GEMM of the mechanics kernels with nvar = 1 -> fp64 tensor cores
    if (D == 3 && h->use_mma_implicit && npe == 125 && ng == 125) return launch_laplacian_mma<125, 125, 12, 4>(h, e_dev, symmetric, ke, st);
    if (D == 3 && h->use_mma_implicit && npe == 64 && ng == 64) return launch_laplacian_mma<64, 64, 12, 0>(h, e_dev, symmetric, ke, st);
    constexpr int A = 8;
    const int xstride = (npe * D) | 1;
    const size_t jm_bytes = sizeof(double) * D * npe * ldg;
    const size_t per_elem = sizeof(double) * ((size_t)xstride + (size_t)ng * D * D);
    const size_t limit = (size_t)h->max_smem_optin;
    const bool jm_in_smem = jm_bytes <= 96 * 1024 && jm_bytes + per_elem <= limit;
    const int tpe = ((npe + A - 1) / A) * npe;
    int EB = LAP_THREADS / tpe;
    if (EB < 1) EB = 1;
    while (EB > 1 && (jm_in_smem ? jm_bytes : 0) + per_elem * EB > limit / 2) --EB;
    const size_t smem = (jm_in_smem ? jm_bytes : 0) + per_elem * EB;
    if (smem > limit) { set_error("laplacian kernel needs %zu bytes of shared memory", smem); return FL_ERR_UNSUPPORTED; }
    const int64_t nbatch = (h->nelem + EB - 1) / EB;
    if (nbatch == 0) return FL_OK;
#define FL_LAP(D_)                                                                                                         \
    {                                                                                                                      \
        auto kern = laplacian_elements_kernel<D_, A>;                                                                      \
        FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        int occ = 1;                                                                                                       \
        FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LAP_THREADS, smem));                       \
        if (occ < 1) occ = 1;                                                                                              \
        const int grid = (int)(nbatch < (int64_t)occ * h->sm_count ? nbatch : (int64_t)occ * h->sm_count);                 \
        kern<<<grid, LAP_THREADS, smem, st>>>(h->conn, h->points, h->jm, h->gw, h->nelem, npe, ng, ldg, EB, jm_in_smem ? 1 : 0, \
                                              e_dev, symmetric, ke);                                                       \
    }
    if (D == 3) FL_LAP(3) else FL_LAP(2)
#undef FL_LAP
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// ------------------------------------------------------------------------------------------------ mass
// lumped: m_(a,i) = rho sum_g N_a(g) (sum_b N_b(g)) w|det J_X|, i < ndim (row sums of rho N N^T, _MassIntegrand_.h:352-366)
// consistent: M_(a,i),(b,i) = rho sum_g N_a N_b w|det J_X|
constexpr int MASS_THREADS = 128;
template <int D>
__global__ void __launch_bounds__(MASS_THREADS)
mass_elements_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ jm,
                     const double* __restrict__ bases, const double* __restrict__ gw, int64_t nelem, int npe, int ng, int ldg, int nvar,
                     double rho, int lumped, double* __restrict__ out) {
    extern __shared__ double smem[];
    double* Xs = smem;          // npe*D
    double* dJ = Xs + npe * D;  // ng
    double* Sg = dJ + ng;       // ng: sum_b N_b(g)
    const int ndof = npe * nvar;
    for (int64_t e = blockIdx.x; e < nelem; e += gridDim.x) {
        __syncthreads();
        for (int it = threadIdx.x; it < npe; it += blockDim.x) {
            const int64_t n = conn[e * npe + it];
#pragma unroll
            for (int l = 0; l < D; ++l) Xs[it * D + l] = X[n * D + l];
        }
        __syncthreads();
        for (int g = threadIdx.x; g < ng; g += blockDim.x) {
            double JX[D * D];
#pragma unroll
            for (int i = 0; i < D * D; ++i) JX[i] = 0.0;
            double s = 0;
            for (int a = 0; a < npe; ++a) {
                s += bases[a * ng + g];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double j = jm[(k * npe + a) * ldg + g];
#pragma unroll
                    for (int l = 0; l < D; ++l) JX[k * D + l] += j * Xs[a * D + l];
                }
            }
            dJ[g] = gw[g] * fabs(det_of(JX));
            Sg[g] = s;
        }
        __syncthreads();
        if (lumped) {
            for (int a = threadIdx.x; a < npe; a += blockDim.x) {
                double m = 0;
                for (int g = 0; g < ng; ++g) m += bases[a * ng + g] * Sg[g] * dJ[g];
                m *= rho;
                for (int i = 0; i < nvar; ++i) out[(e * npe + a) * nvar + i] = (i < D) ? m : 0.0;
            }
        } else {
            for (int it = threadIdx.x; it < npe * npe; it += blockDim.x) {
                const int a = it / npe, b = it - a * npe;
                double m = 0;
                for (int g = 0; g < ng; ++g) m += bases[a * ng + g] * bases[b * ng + g] * dJ[g];
                m *= rho;
                double* Ke = out + (size_t)e * ndof * ndof;
                for (int i = 0; i < nvar; ++i)
                    for (int j = 0; j < nvar; ++j) Ke[(size_t)(a * nvar + i) * ndof + b * nvar + j] = (i == j && i < D) ? m : 0.0;
            }
        }
    }
}

int launch_mass_elements(fl_handle* h, double rho, int nvar, int lumped, double* out, cudaStream_t st) {
    if (h->nelem == 0) return FL_OK;
    const size_t smem = sizeof(double) * ((size_t)h->npe * h->ndim + 2 * (size_t)h->ng);
    const int64_t cap = (int64_t)h->sm_count * 16;
    const int grid = (int)(h->nelem < cap ? h->nelem : cap);
    if (h->ndim == 3)
        mass_elements_kernel<3><<<grid, MASS_THREADS, smem, st>>>(h->conn, h->points, h->jm, h->bases, h->gw, h->nelem, h->npe, h->ng, h->ldg,
                                                                  nvar, rho, lumped, out);
    else
        mass_elements_kernel<2><<<grid, MASS_THREADS, smem, st>>>(h->conn, h->points, h->jm, h->bases, h->gw, h->nelem, h->npe, h->ng, h->ldg,
                                                                  nvar, rho, lumped, out);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

#endif  // FL_IMPL_PART == 0

}  // namespace fl
