// fl_explicit.cu -- matrix-free internal force (B^T sigma) and the fused node-gather / central-difference update.
//
// Device replacement of _GlobalAssemblyExplicit_DF_DPF_<2>/<3>
// (Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.h:215-454, :458-717) and of the
// per-step vector algebra of ExplicitStructuralDynamicIntegrator.Solver
// (Florence/TimeIntegrators/ExplicitStructuralDynamicIntegrator.py:117-184).
//
// Algebra per element (same quantities as the reference, fewer flops -- SURVEY.md H1c):
//   J_X[g] = Jm_g X,  J_x[g] = Jm_g x                (two d x d contractions over the element's nodes)
//   F = (J_X^-1 J_x)^T,  grad_x N_a = J_x^-1 Jm_g[:,a],  detJ = w_g |det J_x|   (AVX rule, :96)
//   sigma(F[,E]) by the compile-time material functor (replaces the runtime material_number chain, :589-621)
//   P_g = detJ J_x^-T sigma  (d x nvar),  t_a = sum_g Jm_g[:,a]^T P_g
// so the five npe-sized GEMMs per Gauss point of the reference become two contractions with the shared Jm table.
// Thread mapping: phase 1 one thread per (element, Gauss point); phase 2 one thread per (element, node); elements are
// packed back to back across the block so no lane idles for npe, ngauss < 32.  Per-element tractions are written to a
// buffer and reduced per node in ascending element order (deterministic; the reference's summation order).
#include "fl_explicit_mma.cuh"

// hex8 batch of the DMMA force kernel: elements per batch and warp split of the traction GEMM.  16 elements per batch need 42 kB of
// shared memory per block, so four blocks (16 warps) fit an SM instead of two: 3.08 -> 2.79 ms per 8 M elements (the kernel is bound
// by its own instruction latencies: issue 29 %, `wait` 2.6 per issue at 8 warps per SM).  Five or six blocks (96 / 80 registers) are
// slower again (2.81 / 3.02 ms): profiles/hex8_explicit_bench.py.
#ifndef FL_HEX8_NE
#define FL_HEX8_NE 16
#define FL_HEX8_WM3 2
#endif

namespace fl {

constexpr int EXPL_THREADS = 256;

template <int D, int MAT, int JM_SMEM>
__global__ void __launch_bounds__(EXPL_THREADS)
explicit_elements_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                         const double* __restrict__ phi, const double* __restrict__ jm_g, const double* __restrict__ gw,
                         int64_t nelem, int npe, int ng, int ldg, int EB, MatParams prm, double* __restrict__ te) {
    constexpr int jm_in_smem = JM_SMEM;
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr int NV = D + (EL ? 1 : 0);
    constexpr int PS = D * NV;
    extern __shared__ double smem[];
    const int xstride = (npe * D) | 1;  // odd element stride: distinct banks for the elements sharing a warp
    double* jm_s = smem;
    double* Xs = jm_s + (jm_in_smem ? D * npe * ldg : 0);
    double* xs = Xs + EB * xstride;
    double* ph = xs + EB * xstride;
    double* Ps = ph + (EL ? EB * npe : 0);
    const double* jm = jm_in_smem ? jm_s : jm_g;

    if (jm_in_smem)
        for (int i = threadIdx.x; i < D * npe * ldg; i += blockDim.x) jm_s[i] = jm_g[i];

    const int64_t nbatch = (nelem + EB - 1) / EB;
    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int64_t e0 = batch * EB;
        const int ne = (int)min((int64_t)EB, nelem - e0);
        __syncthreads();
        // gather nodal coordinates / potentials (connectivity read is fully coalesced)
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
        // phase 1: kinematics + constitutive law at one Gauss point
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
            invdet(JX, iJX);
            const double detx = invdet(Jx, iJx);
            const double detJ = gw[g] * fabs(detx);
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
            double E[D], Dv[D], sig[D * D];
            if (EL) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double v = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) v += iJx[k * D + jj] * gp[jj];
                    E[k] = -v;
                }
            }
            kinetic_measures<D, MAT, false>(F, E, prm, sig, Dv, nullptr);
            // P[k][i] = detJ sum_j iJx[j][k] sigma_sym[j][i]; the reference reads the upper triangle of sigma (:624-640)
            double* P = Ps + (el * ng + g) * PS;
#pragma unroll
            for (int k = 0; k < D; ++k) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    double v = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) v += iJx[jj * D + k] * (jj <= i ? sig[jj * D + i] : sig[i * D + jj]);
                    P[k * NV + i] = v * detJ;
                }
                if (EL) {
                    double v = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) v += iJx[jj * D + k] * Dv[jj];
                    P[k * NV + D] = v * detJ;
                }
            }
        }
        __syncthreads();
        // phase 2: t_a = sum_g Jm_g[:,a]^T P_g
        for (int it = threadIdx.x; it < ne * npe; it += blockDim.x) {
            const int el = it / npe, a = it - el * npe;
            // two interleaved accumulator sets (even / odd Gauss points) shorten the dependent DFMA chains
            double t[NV], u[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] = u[i] = 0.0;
            const double* Pe = Ps + el * ng * PS;
            int g = 0;
            for (; g + 1 < ng; g += 2) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double j0 = jm[(k * npe + a) * ldg + g], j1 = jm[(k * npe + a) * ldg + g + 1];
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        t[i] = fma(j0, Pe[g * PS + k * NV + i], t[i]);
                        u[i] = fma(j1, Pe[(g + 1) * PS + k * NV + i], u[i]);
                    }
                }
            }
            if (g < ng) {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double j0 = jm[(k * npe + a) * ldg + g];
#pragma unroll
                    for (int i = 0; i < NV; ++i) t[i] = fma(j0, Pe[g * PS + k * NV + i], t[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) t[i] += u[i];
#pragma unroll
            for (int i = 0; i < NV; ++i) te[(e0 * npe + it) * NV + i] = t[i];
        }
    }
}

// T[n] = sum over the (element, local node) pairs of node n, ascending element order (RHSAssemblyNative.pyx:30-39 order)
template <int NV>
__global__ void gather_nodes_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                                    const double* __restrict__ te, int64_t nnode, double* __restrict__ T) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    const int64_t k1 = adj_ptr[n + 1];
    for (int64_t k = adj_ptr[n]; k < k1; ++k) {
        const int64_t idx = adj_idx[k];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += te[idx * NV + i];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) T[n * NV + i] = acc[i];
}

// order-preserving map double -> int64 (signed compare), so that the running maximum of a vector can be kept with atomicMax
__device__ __forceinline__ long long ordered_key(double v) {
    const long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7FFFFFFFFFFFFFFFLL);
}
__device__ __forceinline__ double key_to_double(long long k) {
    return __longlong_as_double(k >= 0 ? k : (k ^ 0x7FFFFFFFFFFFFFFFLL));
}

// Central-difference update of one node (mechanics, lumped mass).  Written with explicit rounding so that the result is
// the one numpy produces for ExplicitStructuralDynamicIntegrator.py:131-157 given the same T:
//   R = (fs*F - T) + (((2/dt^2) M) U0 - ((1/dt^2) M) U00);  U = ((dt^2) (1/M)) R;  Eulerx = (X + U) + IncDirichlet
// FUSED: T of the node is reduced here from the per-element tractions (ascending element order); for a node on a partition
// interface (iface_slot[n] >= 0) the reduced force is instead read from T_iface, where the partial sums of all ranks sharing the
// node have been added in ascending rank order (Assembly.py:1352-1354 `T_all[pnodes] += T_p`, made order-deterministic).
// growth (2 x int64, optional): running signed maxima of the new U and of the previous U0 -- the operands of the reference's
// blow-up test `abs(U.max() / (U0.max() + 1e-14)) > tol` (:175-180), evaluated by growth_check_kernel once the grid has finished.
template <int D, bool FUSED>
__global__ void explicit_update_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                                       const double* __restrict__ te, int64_t nnode, double dt, double fs, double ds,
                                       const double* __restrict__ M, const double* __restrict__ fext, const uint8_t* __restrict__ fixed,
                                       const double* __restrict__ inc_dir, const double* __restrict__ X, double* __restrict__ T,
                                       int write_T, const int32_t* __restrict__ iface_slot, const double* __restrict__ T_iface,
                                       double* __restrict__ U0, double* __restrict__ U00, double* __restrict__ Eulerx,
                                       int32_t* __restrict__ nan_flag, long long* __restrict__ growth, const Contact contact) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = n < nnode;
    double umax = -INFINITY, u0max = -INFINITY;
    if (live) {
        double t[D];
        if (FUSED) {
            const int slot = iface_slot ? iface_slot[n] : -1;
            if (slot >= 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) t[i] = T_iface[(int64_t)slot * D + i];
            } else {
#pragma unroll
                for (int i = 0; i < D; ++i) t[i] = 0.0;
                const int64_t k1 = adj_ptr[n + 1];
                for (int64_t k = adj_ptr[n]; k < k1; ++k) {
                    const int64_t idx = adj_idx[k];
#pragma unroll
                    for (int i = 0; i < D; ++i) t[i] += te[idx * D + i];
                }
            }
            // TractionForces += contact tractions evaluated at the geometry the element forces were computed on
            // (ExplicitStructuralDynamicIntegrator.py:190-197); a caller-supplied T (FUSED = false) already contains them
            if (contact.surf && contact.surf[n]) {
                double xo[D], fc[D];
#pragma unroll
                for (int i = 0; i < D; ++i) xo[i] = Eulerx[n * D + i];
                if (contact_force<D>(contact, xo, fc)) {
#pragma unroll
                    for (int i = 0; i < D; ++i) t[i] = __dadd_rn(t[i], fc[i]);
                }
            }
            if (write_T) {
#pragma unroll
                for (int i = 0; i < D; ++i) T[n * D + i] = t[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) t[i] = T[n * D + i];
        }
        const double dt2 = __dmul_rn(dt, dt);
        const double c2 = 2. / dt2, c1 = 1. / dt2;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const int64_t q = n * D + i;
            const double m = M[q], u0 = U0[q], u00 = U00[q];
            const double f = fext ? __dmul_rn(fext[q], fs) : 0.0;
            double R = __dadd_rn(f, -t[i]);
            const double inert = __dadd_rn(__dmul_rn(__dmul_rn(c2, m), u0), -__dmul_rn(__dmul_rn(c1, m), u00));
            R = __dadd_rn(R, inert);
            double U = __dmul_rn(__dmul_rn(dt2, 1.0 / m), R);
            const bool fx = fixed && fixed[q];
            if (fx) U = 0.0;
            const double incd = (fx && inc_dir) ? __dmul_rn(inc_dir[q], ds) : 0.0;
            Eulerx[q] = __dadd_rn(__dadd_rn(X[q], U), incd);
            U00[q] = u0;
            U0[q] = U;
            bad |= (U != U);
            umax = fmax(umax, U);
            u0max = fmax(u0max, u0);
        }
        if (bad) atomicOr(nan_flag, 1);
    }
    if (growth) {
        // block maximum first, and the global word is only touched when it would change: two same-address atomics per warp
        // (4 M of them at 64 M nodes) serialise in L2 and cost more than the update itself
        __shared__ double s_u[8], s_u0[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            umax = fmax(umax, __shfl_xor_sync(0xffffffffu, umax, o));
            u0max = fmax(u0max, __shfl_xor_sync(0xffffffffu, u0max, o));
        }
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { s_u[w] = umax; s_u0[w] = u0max; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            for (int k = 1; k < nw; ++k) { umax = fmax(umax, s_u[k]); u0max = fmax(u0max, s_u0[k]); }
            if (umax > -INFINITY) {
                const long long ku = ordered_key(umax), k0 = ordered_key(u0max);
                if (ku > *reinterpret_cast<volatile long long*>(growth)) atomicMax(growth, ku);
                if (k0 > *reinterpret_cast<volatile long long*>(growth + 1)) atomicMax(growth + 1, k0);
            }
        }
    }
}

// The reference's blow-up test (ExplicitStructuralDynamicIntegrator.py:175-180) on the maxima collected by the update kernel:
// tol = 1e200 before increment 5, 10 afterwards.  status[0] |= 2 and status[1] = increment at the first detection (a NaN sets
// bit 0 in the update kernel; its increment is recorded here as well).  The maxima are reset for the next step.
__global__ void growth_check_kernel(long long* __restrict__ growth, long long increment, int32_t* __restrict__ status) {
    const double umax = key_to_double(growth[0]), u0max = key_to_double(growth[1]);
    const double tol = increment < 5 ? 1e200 : 10.0;
    int32_t s = status[0];
    if (fabs(umax / (u0max + 1e-14)) > tol) s |= 2;
    if (s != 0 && status[1] == 0) status[1] = (int32_t)increment;
    status[0] = s;
    growth[0] = growth[1] = (long long)0x8000000000000000ULL;
}

// buf[k] = sum over the (element, local node) visits of node ids[k] (ascending element order) of the per-element tractions:
// gather_nodes_kernel restricted to a node list, written densely -- the partial interface forces a rank sends to its neighbours
template <int NV>
__global__ void gather_pack_kernel(const int64_t* __restrict__ adj_ptr, const int32_t* __restrict__ adj_idx,
                                   const double* __restrict__ te, const int32_t* __restrict__ ids, int64_t n, double* __restrict__ buf) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t node = ids[k];
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    const int64_t k1 = adj_ptr[node + 1];
    for (int64_t q = adj_ptr[node]; q < k1; ++q) {
        const int64_t idx = adj_idx[q];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += te[idx * NV + i];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[k * NV + i] = acc[i];
}

int launch_gather_pack(fl_handle* h, int nvar, const double* te, const int32_t* ids, int64_t n, double* buf, cudaStream_t st) {
    if (n == 0) return FL_OK;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    switch (nvar) {
        case 1: gather_pack_kernel<1><<<blocks, 128, 0, st>>>(h->adj_ptr, h->adj_idx, te, ids, n, buf); break;
        case 2: gather_pack_kernel<2><<<blocks, 128, 0, st>>>(h->adj_ptr, h->adj_idx, te, ids, n, buf); break;
        case 3: gather_pack_kernel<3><<<blocks, 128, 0, st>>>(h->adj_ptr, h->adj_idx, te, ids, n, buf); break;
        case 4: gather_pack_kernel<4><<<blocks, 128, 0, st>>>(h->adj_ptr, h->adj_idx, te, ids, n, buf); break;
        default: set_error("nvar=%d unsupported", nvar); return FL_ERR_INVALID;
    }
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int launch_growth_check(int64_t* growth, int64_t increment, int32_t* status, cudaStream_t st) {
    growth_check_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(growth), (long long)increment, status);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// ---------------------------------------------------------------------------------------------------------
template <int D, int MAT>
static int launch_expl(fl_handle* h, const double* Eulerx, const double* Eulerp, const MatParams& prm, double* te, cudaStream_t st) {
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr int NV = D + (EL ? 1 : 0);
    const int npe = h->npe, ng = h->ng, ldg = h->ldg;
    if constexpr (D == 3 && !EL) {
        // fixed shapes of the benchmark configs run the tensor-core (DMMA) formulation
        if (h->use_mma && npe == 27 && ng == 27) return launch_expl_mma<MAT, 27, 27, 8, 4, 4>(h, Eulerx, prm, te, st);
        if (h->use_mma && npe == 8 && ng == 8) return launch_expl_mma<MAT, 8, 8, FL_HEX8_NE, 1, FL_HEX8_WM3>(h, Eulerx, prm, te, st);
    }
    const int per = npe > ng ? npe : ng;
    if (per > EXPL_THREADS) {
        set_error("element with %d nodes / %d gauss points exceeds the explicit kernel's block", npe, ng);
        return FL_ERR_UNSUPPORTED;
    }
    int EB = EXPL_THREADS / per;
    const int xstride = (npe * D) | 1;
    const size_t jm_bytes = sizeof(double) * D * npe * ldg;
    auto smem_for = [&](int eb, bool jm_s) {
        return (jm_s ? jm_bytes : 0) + sizeof(double) * ((size_t)2 * eb * xstride + (EL ? eb * npe : 0) + (size_t)eb * ng * D * NV);
    };
    const bool jm_in_smem = jm_bytes <= 100 * 1024;
    while (EB > 1 && smem_for(EB, jm_in_smem) > (size_t)h->max_smem_optin) --EB;
    const size_t smem = smem_for(EB, jm_in_smem);
    if (smem > (size_t)h->max_smem_optin) {
        set_error("explicit kernel needs %zu bytes of shared memory", smem);
        return FL_ERR_UNSUPPORTED;
    }
    auto kern = jm_in_smem ? explicit_elements_kernel<D, MAT, 1> : explicit_elements_kernel<D, MAT, 0>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, EXPL_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t e0 = h->el0, ne = (h->el1 >= 0 ? h->el1 : h->nelem) - e0;   // element range of the call (fl_explicit_forces)
    const int64_t nbatch = (ne + EB - 1) / EB;
    const int grid = (int)(nbatch < (int64_t)occ * h->sm_count ? nbatch : (int64_t)occ * h->sm_count);
    if (grid <= 0) return FL_OK;
    kern<<<grid, EXPL_THREADS, smem, st>>>(h->conn + e0 * npe, h->points, Eulerx, Eulerp, h->jm, h->gw, ne, npe, ng, ldg, EB, prm,
                                           te + e0 * npe * NV);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

static MatParams to_params(const fl_material* m) {
    MatParams p;
    p.mu = m->mu; p.mu1 = m->mu1; p.mu2 = m->mu2; p.mu3 = m->mu3; p.mue = m->mue; p.lamb = m->lamb;
    p.eps_1 = m->eps_1; p.eps_2 = m->eps_2; p.eps_3 = m->eps_3; p.eps_e = m->eps_e;
    return p;
}

int launch_explicit_elements(fl_handle* h, const double* Eulerx, const double* Eulerp, const fl_material* mat, int formulation,
                             double* te, cudaStream_t st) {
    const MatParams prm = to_params(mat);
    const int m = mat->material_number;
    const bool electro = (m == MAT_ELECTRO_101 || m == MAT_ELECTRO_105 || m == MAT_ELECTRO_108 || m == MAT_EXPLICIT_ELECTRO_108);
    if (electro != (formulation == 1)) {
        set_error("material %d does not match formulation_number %d", m, formulation);
        return FL_ERR_INVALID;
    }
    if (electro && !Eulerp) {
        set_error("Eulerp is required for electro-mechanical materials");
        return FL_ERR_INVALID;
    }
#define FL_CASE(MATID)                                                                  \
    case MATID:                                                                         \
        return h->ndim == 3 ? launch_expl<3, MATID>(h, Eulerx, Eulerp, prm, te, st)     \
                            : launch_expl<2, MATID>(h, Eulerx, Eulerp, prm, te, st);
    switch (m) {
        FL_CASE(MAT_EXPLICIT_MOONEY_RIVLIN)
        FL_CASE(MAT_NEOHOOKEAN)
        FL_CASE(MAT_MOONEY_RIVLIN)
        FL_CASE(MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN)
        FL_CASE(MAT_ELECTRO_101)
        FL_CASE(MAT_ELECTRO_105)
        FL_CASE(MAT_ELECTRO_108)
        FL_CASE(MAT_EXPLICIT_ELECTRO_108)
        FL_CASE(MAT_LINEAR_ELASTIC)
        default:
            // same behaviour as _LowLevelAssemblyExplicit_DF_DPF_.pyx:107-109 (NotImplementedError)
            set_error("Low level assembly for material number %d not available for explicit analysis", m);
            return FL_ERR_UNSUPPORTED;
    }
#undef FL_CASE
}

int launch_gather_nodes(fl_handle* h, int nvar, const double* te, double* T, cudaStream_t st) {
    if (h->nnode == 0) return FL_OK;
    const int threads = 256;
    const int64_t blocks = (h->nnode + threads - 1) / threads;
    switch (nvar) {
        case 1: gather_nodes_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(h->adj_ptr, h->adj_idx, te, h->nnode, T); break;
        case 2: gather_nodes_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(h->adj_ptr, h->adj_idx, te, h->nnode, T); break;
        case 3: gather_nodes_kernel<3><<<(unsigned)blocks, threads, 0, st>>>(h->adj_ptr, h->adj_idx, te, h->nnode, T); break;
        case 4: gather_nodes_kernel<4><<<(unsigned)blocks, threads, 0, st>>>(h->adj_ptr, h->adj_idx, te, h->nnode, T); break;
        default: set_error("nvar=%d unsupported", nvar); return FL_ERR_INVALID;
    }
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

int launch_explicit_update(fl_handle* h, int fused_gather, const double* te, double dt, double fext_scale, double incd_scale,
                           const double* M, const double* fext, const uint8_t* fixed, const double* inc_dir, double* T,
                           const int32_t* iface_slot, const double* T_iface, double* U0, double* U00, double* Eulerx, int32_t* nan_flag,
                           int64_t* growth, cudaStream_t st) {
    if (h->nnode == 0) return FL_OK;
    const int threads = 256;
    const unsigned blocks = (unsigned)((h->nnode + threads - 1) / threads);
    const int write_T = (fused_gather == 2);
#define FL_UPD(D_, F_)                                                                                                          \
    explicit_update_kernel<D_, F_><<<blocks, threads, 0, st>>>(h->adj_ptr, h->adj_idx, te, h->nnode, dt, fext_scale, incd_scale, M, fext, \
                                                               fixed, inc_dir, h->points, T, write_T, iface_slot, T_iface, U0, U00, Eulerx,  \
                                                               nan_flag, reinterpret_cast<long long*>(growth), h->contact)
    if (h->ndim == 3) {
        if (fused_gather) FL_UPD(3, true); else FL_UPD(3, false);
    } else {
        if (fused_gather) FL_UPD(2, true); else FL_UPD(2, false);
    }
#undef FL_UPD
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

// AssembleTractions of the contact formulation as a stand-alone pass: T (+)= kappa * gap * n on the contact nodes
template <int D>
__global__ void contact_kernel(const Contact c, const double* __restrict__ x, int64_t nnode, double* __restrict__ T, int accumulate) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= nnode) return;
    double xo[D], fc[D];
    bool hit = false;
    if (c.surf[n]) {
#pragma unroll
        for (int i = 0; i < D; ++i) xo[i] = x[n * D + i];
        hit = contact_force<D>(c, xo, fc);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        if (accumulate) {
            if (hit) T[n * D + i] = __dadd_rn(T[n * D + i], fc[i]);
        } else {
            T[n * D + i] = hit ? fc[i] : 0.0;
        }
    }
}

int launch_contact(fl_handle* h, const double* Eulerx, double* T, int accumulate, cudaStream_t st) {
    if (!h->contact.surf) { set_error("fl_set_contact has not been called"); return FL_ERR_STATE; }
    if (h->nnode == 0) return FL_OK;
    const unsigned blocks = (unsigned)((h->nnode + 255) / 256);
    if (h->ndim == 3) contact_kernel<3><<<blocks, 256, 0, st>>>(h->contact, Eulerx, h->nnode, T, accumulate);
    else contact_kernel<2><<<blocks, 256, 0, st>>>(h->contact, Eulerx, h->nnode, T, accumulate);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
