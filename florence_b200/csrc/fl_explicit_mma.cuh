// fl_explicit_mma.cuh -- matrix-free internal force for 3-D mechanics on fixed element shapes, with both contractions
// against the shared Jm table done on the fp64 tensor-core path (mma.sync m8n8k4 f64, DMMA).
//
// Same algebra as explicit_elements_kernel (fl_explicit.cu; reference _LowLevelAssemblyExplicit_DF_DPF_.h:458-717), regrouped
// as two small GEMMs per batch of NE elements whose A operand is the SAME for every element:
//   (1) Jac[(g,k), (e,c)]  = sum_a  Jm[k][a][g] * XX[a, (e,c)]        XX = [X | x] nodal coordinates, c = 0..5
//   (2) t  [a,     (e,i)]  = sum_gk Jm[k][a][g] * P [(g,k), (e,i)]    P_g = w|J_x| J_x^-T sigma
// Each warp keeps its A fragments (slices of the Jm table) in registers for the whole kernel; B fragments stream from
// shared memory once per 8 output rows, so one shared-memory load feeds 256 FMAs instead of 1-2 in the scalar kernel.
// Between the GEMMs one thread per (element, Gauss point) inverts the Jacobians and evaluates the material law.
#pragma once
#include "fl_internal.cuh"

namespace fl {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int MMA_THREADS = 128;

template <int NPE, int NG, int NE>
struct mma_shape {
    static constexpr int M1 = 3 * NG, K1 = NPE, N1 = 6 * NE;
    static constexpr int M3 = NPE, K3 = 3 * NG, N3 = 3 * NE;
    static constexpr int MT1 = (M1 + 7) / 8, KS1 = (K1 + 3) / 4, NT1 = N1 / 8;
    static constexpr int MT3 = (M3 + 7) / 8, KS3 = (K3 + 3) / 4, NT3 = N3 / 8;
    // B-fragment tiles: lane 4 n + k reads [k][n], and a 64-bit shared load is served per half warp (n = 0..3 | 4..7, k = 0..3),
    // so the leading dimension must be == 4 or 12 (mod 16) doubles for the 16 addresses to fall into 16 distinct banks
    // (== 8, the rule for 32-bit fragments, is a 2-way conflict here: ncu showed 4 wavefronts per LDS.64 instead of 2)
    static constexpr int pad_b(int n) { return ((4 - n % 16) + 16) % 16 < ((12 - n % 16) + 16) % 16 ? ((4 - n % 16) + 16) % 16 : ((12 - n % 16) + 16) % 16; }
    static constexpr int LDX = N1 + pad_b(N1), LDP = N3 + pad_b(N3), LDJ = NPE > 8 ? N1 + 1 : N1 + 2;  // hex27: 3 blocks/SM need the smaller tile; hex8 keeps 16-byte aligned rows
    static constexpr int XX_SZ = KS1 * 4 * LDX, JAC_SZ = MT1 * 8 * LDJ, P_SZ = KS3 * 4 * LDP, TE_SZ = NE * NPE * 3;
    static constexpr int SCR_SZ = JAC_SZ > TE_SZ ? JAC_SZ : TE_SZ;  // te staging aliases the Jacobian tile
    static constexpr size_t SMEM = sizeof(double) * (size_t)(2 * XX_SZ + SCR_SZ + P_SZ);  // coordinates are double-buffered
};

#ifndef FL_HEX8_MINB
#define FL_HEX8_MINB 4
#endif
template <int MAT, int NPE, int NG, int NE, int WM1, int WM3>
__global__ void __launch_bounds__(MMA_THREADS, (NPE > 8 ? 3 : FL_HEX8_MINB))
explicit_elements_mma_kernel(const int32_t* __restrict__ conn, const double* __restrict__ X, const double* __restrict__ x,
                             const double* __restrict__ jm, const double* __restrict__ gw, int64_t nelem, int ldg, MatParams prm,
                             double* __restrict__ te) {
    using S = mma_shape<NPE, NG, NE>;
    constexpr int D = 3;
    constexpr int WN1 = 4 / WM1, WN3 = 4 / WM3;
    constexpr int MTW1 = (S::MT1 + WM1 - 1) / WM1, NTW1 = S::NT1 / WN1;
    constexpr int MTW3 = (S::MT3 + WM3 - 1) / WM3, NTW3 = S::NT3 / WN3;
    static_assert(S::NT1 % WN1 == 0 && S::NT3 % WN3 == 0, "n-tiles must split evenly over the warps");
    constexpr int NB1 = (NTW1 % 3 == 0) ? 3 : (NTW1 % 2 == 0 ? 2 : 1);
    constexpr int GI = (NE * NPE + MMA_THREADS - 1) / MMA_THREADS;  // gather items per thread
    extern __shared__ double smem[];
    double* XXs = smem;                 // 2 x [KS1*4][LDX] nodal coordinates (double buffer), rows >= NPE stay zero
    double* Scr = XXs + 2 * S::XX_SZ;   // [MT1*8][LDJ]   Jacobians, later [NE][NPE][3] tractions
    double* Ps = Scr + S::SCR_SZ;       // [KS3*4][LDP]   P, rows >= 3 NG stay zero
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lc = lane & 3;  // fragment row group / thread-in-group
    const int wm1 = warp % WM1, wn1 = warp / WM1, wm3 = warp % WM3, wn3 = warp / WM3;

    // A fragments (row-major 8x4 tiles): a[r][c] lives in lane 4 r + c
    double A1[MTW1][S::KS1], A3[MTW3][S::KS3];
#pragma unroll
    for (int j = 0; j < MTW1; ++j)
#pragma unroll
        for (int ks = 0; ks < S::KS1; ++ks) {
            const int r = 8 * (wm1 + WM1 * j) + lr, a = 4 * ks + lc;
            const int g = r / 3, k = r - 3 * g;
            A1[j][ks] = (r < S::M1 && a < NPE) ? jm[(k * NPE + a) * ldg + g] : 0.0;
        }
#pragma unroll
    for (int j = 0; j < MTW3; ++j)
#pragma unroll
        for (int ks = 0; ks < S::KS3; ++ks) {
            const int a = 8 * (wm3 + WM3 * j) + lr, r = 4 * ks + lc;
            const int g = r / 3, k = r - 3 * g;
            A3[j][ks] = (a < NPE && r < S::K3) ? jm[(k * NPE + a) * ldg + g] : 0.0;
        }
    for (int i = threadIdx.x; i < 2 * S::XX_SZ; i += MMA_THREADS) XXs[i] = 0.0;
    for (int i = threadIdx.x; i < S::P_SZ; i += MMA_THREADS) Ps[i] = 0.0;

    // Asynchronous gather (cp.async, 8 bytes per coordinate) of the [X | x] tile of a batch into buffer `buf`:
    // XX[a][e*6 + c].  Elements beyond the end of the mesh are zero-filled.  The connectivity entries of the batch were loaded
    // into registers one iteration earlier (load_conn), so the cp.async addresses never wait on a global load.
    int32_t cn[GI];
    auto load_conn = [&](int64_t b0) {
        const int nb = (int)min((int64_t)NE, nelem - b0);
#pragma unroll
        for (int q = 0; q < GI; ++q) {
            const int it = threadIdx.x + q * MMA_THREADS;
            const int a = it / NE, el = it - a * NE;
            cn[q] = (it < NE * NPE && el < nb) ? conn[(b0 + el) * NPE + a] : -1;
        }
    };
    auto gather = [&](int buf) {
        double* dstb = XXs + buf * S::XX_SZ;
#pragma unroll
        for (int q = 0; q < GI; ++q) {
            const int it = threadIdx.x + q * MMA_THREADS;
            if (it < NE * NPE) {
                // element index fastest: the lanes of a warp write one tile row (stride 6 doubles), not one tile column
                // (stride LDX, a 16-way bank conflict on every cp.async)
                const int a = it / NE, el = it - a * NE;
                double* d = dstb + a * S::LDX + el * 6;
                if (cn[q] >= 0) {
                    const int64_t n = cn[q];
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + 8 * l), "l"(X + n * 3 + l));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa + 8 * (3 + l)), "l"(x + n * 3 + l));
                    }
                } else {
#pragma unroll
                    for (int l = 0; l < 6; ++l) d[l] = 0.0;
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    };

    const int64_t nbatch = (nelem + NE - 1) / NE;
    __syncthreads();
    if ((int64_t)blockIdx.x < nbatch) {
        load_conn((int64_t)blockIdx.x * NE);
        gather(0);
        if ((int64_t)blockIdx.x + gridDim.x < nbatch) load_conn(((int64_t)blockIdx.x + gridDim.x) * NE);
    }
    int buf = 0;
    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x, buf ^= 1) {
        const int64_t e0 = batch * NE;
        const int ne = (int)min((int64_t)NE, nelem - e0);
        const double* XXc = XXs + buf * S::XX_SZ;
        asm volatile("cp.async.wait_all;");
        __syncthreads();   // coordinates of this batch have landed; previous batch's write-out is finished
        // prefetch the next batch's coordinates while this one is computed, and the connectivity of the one after
        if (batch + gridDim.x < nbatch) {
            gather(buf ^ 1);
            if (batch + 2 * (int64_t)gridDim.x < nbatch) load_conn((batch + 2 * (int64_t)gridDim.x) * NE);
        }
        // ---- GEMM 1: Jacobians.  NB1 n-tiles are accumulated together so MTW1*NB1 independent DMMA chains are in flight.
#pragma unroll 1
        for (int jn = 0; jn < NTW1; jn += NB1) {
            double c[MTW1][NB1][2];
#pragma unroll
            for (int j = 0; j < MTW1; ++j)
#pragma unroll
                for (int q = 0; q < NB1; ++q) c[j][q][0] = c[j][q][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < S::KS1; ++ks) {
                double b[NB1];
#pragma unroll
                for (int q = 0; q < NB1; ++q) b[q] = XXc[(4 * ks + lc) * S::LDX + 8 * (wn1 + WN1 * (jn + q)) + lr];
#pragma unroll
                for (int j = 0; j < MTW1; ++j)
#pragma unroll
                    for (int q = 0; q < NB1; ++q) dmma884(c[j][q][0], c[j][q][1], A1[j][ks], b[q]);
            }
#pragma unroll
            for (int j = 0; j < MTW1; ++j) {
                const int mt = wm1 + WM1 * j;
                if (mt < S::MT1) {
#pragma unroll
                    for (int q = 0; q < NB1; ++q) {
                        double* o = Scr + (8 * mt + lr) * S::LDJ + 8 * (wn1 + WN1 * (jn + q)) + 2 * lc;
                        o[0] = c[j][q][0];
                        o[1] = c[j][q][1];
                    }
                }
            }
        }
        __syncthreads();
        // ---- kinematics + constitutive law at (element, Gauss point)
        for (int it = threadIdx.x; it < NE * NG; it += MMA_THREADS) {
            // element index fastest: a warp reads / writes along tile rows (the P stores were a 16-way conflict column-wise)
            const int g = it / NE, el = it - g * NE;
            double Pv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (el < ne) {
                double JX[9], Jx[9];
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        JX[k * 3 + l] = Scr[(g * 3 + k) * S::LDJ + el * 6 + l];
                        Jx[k * 3 + l] = Scr[(g * 3 + k) * S::LDJ + el * 6 + 3 + l];
                    }
                double iJX[9], iJx[9];
                invdet(JX, iJX);
                const double detx = invdet(Jx, iJx);
                const double detJ = gw[g] * fabs(detx);
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
                double sig[9];
                kinetic_measures<D, MAT, false>(F, nullptr, prm, sig, nullptr, nullptr);
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        double v = 0;
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) v = fma(iJx[jj * 3 + k], (jj <= i ? sig[jj * 3 + i] : sig[i * 3 + jj]), v);
                        Pv[k * 3 + i] = v * detJ;
                    }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int i = 0; i < 3; ++i) Ps[(g * 3 + k) * S::LDP + el * 3 + i] = Pv[k * 3 + i];
        }
        __syncthreads();
        // ---- GEMM 2: nodal tractions, staged in shared memory in te layout [e][a][i].  All n-tiles of the warp and two
        // K halves are accumulated together (MTW3*NTW3*2 independent DMMA chains).
        {
            double c[MTW3][NTW3][2][2];
#pragma unroll
            for (int j = 0; j < MTW3; ++j)
#pragma unroll
                for (int q = 0; q < NTW3; ++q) c[j][q][0][0] = c[j][q][0][1] = c[j][q][1][0] = c[j][q][1][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < S::KS3; ++ks) {
                double b[NTW3];
#pragma unroll
                for (int q = 0; q < NTW3; ++q) b[q] = Ps[(4 * ks + lc) * S::LDP + 8 * (wn3 + WN3 * q) + lr];
#pragma unroll
                for (int j = 0; j < MTW3; ++j)
#pragma unroll
                    for (int q = 0; q < NTW3; ++q) dmma884(c[j][q][ks & 1][0], c[j][q][ks & 1][1], A3[j][ks], b[q]);
            }
#pragma unroll
            for (int j = 0; j < MTW3; ++j) {
                const int a = 8 * (wm3 + WM3 * j) + lr;
                if (a < NPE) {
#pragma unroll
                    for (int q = 0; q < NTW3; ++q)
#pragma unroll
                        for (int z = 0; z < 2; ++z) {
                            const int n = 8 * (wn3 + WN3 * q) + 2 * lc + z;
                            const int el = n / 3, i = n - 3 * el;
                            Scr[el * (NPE * 3) + a * 3 + i] = c[j][q][0][z] + c[j][q][1][z];
                        }
                }
            }
        }
        __syncthreads();
        double* dst = te + (size_t)e0 * NPE * 3;
        for (int t = threadIdx.x; t < ne * NPE * 3; t += MMA_THREADS) dst[t] = Scr[t];
    }
}

template <int MAT, int NPE, int NG, int NE, int WM1, int WM3>
int launch_expl_mma(fl_handle* h, const double* Eulerx, const MatParams& prm, double* te, cudaStream_t st) {
    using S = mma_shape<NPE, NG, NE>;
    auto kern = explicit_elements_mma_kernel<MAT, NPE, NG, NE, WM1, WM3>;
    FL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    int occ = 1;
    FL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MMA_THREADS, S::SMEM));
    if (occ < 1) occ = 1;
    const int64_t e0 = h->el0, ne = (h->el1 >= 0 ? h->el1 : h->nelem) - e0;   // element range of the call (fl_explicit_forces)
    const int64_t nbatch = (ne + NE - 1) / NE;
    const int grid = (int)(nbatch < (int64_t)occ * h->sm_count ? nbatch : (int64_t)occ * h->sm_count);
    if (grid <= 0) return FL_OK;
    kern<<<grid, MMA_THREADS, S::SMEM, st>>>(h->conn + e0 * NPE, h->points, Eulerx, h->jm, h->gw, ne, h->ldg, prm, te + e0 * NPE * 3);
    FL_CUDA_CHECK(cudaGetLastError());
    return FL_OK;
}

}  // namespace fl
