// fl_math.cuh -- device-side small-matrix algebra and material-point kernels (fp64, registers).
//
// Replaces, on device, the reference's
//   Florence/Tensor/_det_inv_.h:32-63,158-199            (adjugate inverse / determinant)
//   Florence/MaterialLibrary/LLDispatch/CythonSource/_<Material>_.h::_KineticMeasures_
//   Florence/MaterialLibrary/LLDispatch/CythonSource/_LegendreTransform_.h:67-82, _helper_.h:69-113
//   Fastor::voigt (Python twin Florence/Tensor/Numeric.pyx:181-278)
// Every loop below has compile-time bounds so the tensors stay in registers.
#pragma once
#include <cstdint>

namespace fl {

// material numbers: _LowLevelAssemblyExplicit_DF_DPF_.pyx:72-109
enum : int {
    MAT_EXPLICIT_MOONEY_RIVLIN = 0,
    MAT_NEOHOOKEAN = 1,
    MAT_MOONEY_RIVLIN = 2,
    MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN = 3,
    MAT_ELECTRO_101 = 4,
    MAT_ELECTRO_105 = 5,
    MAT_ELECTRO_106 = 6,
    MAT_ELECTRO_107 = 7,
    MAT_ELECTRO_108 = 8,
    MAT_EXPLICIT_ELECTRO_108 = 9,
    MAT_LINEAR_ELASTIC = 10
};

// constants in the order of the reference C signature (_LowLevelAssemblyDF_.pyx:38-48)
struct MatParams {
    double mu, mu1, mu2, mu3, mue, lamb, eps_1, eps_2, eps_3, eps_e;
};

template <int MAT>
struct mat_traits {
    static constexpr bool electro = (MAT == MAT_ELECTRO_101 || MAT == MAT_ELECTRO_105 || MAT == MAT_ELECTRO_108 || MAT == MAT_EXPLICIT_ELECTRO_108);
    static constexpr bool has_tangent = !(MAT == MAT_EXPLICIT_MOONEY_RIVLIN || MAT == MAT_EXPLICIT_ELECTRO_108);
    static constexpr bool geometric = (MAT != MAT_LINEAR_ELASTIC);  // AOT_Assembler.py:79-86
};

template <int D>
struct voigt_map;
template <>
struct voigt_map<3> {
    static constexpr int HS = 6;
    __host__ __device__ static constexpr int i(int I) { return I < 3 ? I : (I == 5 ? 1 : 0); }
    __host__ __device__ static constexpr int j(int I) { return I < 3 ? I : (I == 3 ? 1 : 2); }
};
template <>
struct voigt_map<2> {
    static constexpr int HS = 3;
    __host__ __device__ static constexpr int i(int I) { return I < 2 ? I : 0; }
    __host__ __device__ static constexpr int j(int I) { return I < 2 ? I : 1; }
};

__device__ __forceinline__ double kd(int a, int b) { return a == b ? 1.0 : 0.0; }

// adjugate inverse, returns det (same expression order as _det_inv_.h)
__device__ __forceinline__ double invdet(const double (&s)[9], double (&d)[9]) {
    d[0] = +s[4] * s[8] - s[5] * s[7];
    d[1] = -s[1] * s[8] + s[2] * s[7];
    d[2] = +s[1] * s[5] - s[2] * s[4];
    d[3] = -s[3] * s[8] + s[5] * s[6];
    d[4] = +s[0] * s[8] - s[2] * s[6];
    d[5] = -s[0] * s[5] + s[2] * s[3];
    d[6] = +s[3] * s[7] - s[4] * s[6];
    d[7] = -s[0] * s[7] + s[1] * s[6];
    d[8] = +s[0] * s[4] - s[1] * s[3];
    const double det = s[0] * d[0] + s[1] * d[3] + s[2] * d[6];
    const double r = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] *= r;
    return det;
}
__device__ __forceinline__ double invdet(const double (&s)[4], double (&d)[4]) {
    d[0] = +s[3];
    d[1] = -s[1];
    d[2] = -s[2];
    d[3] = +s[0];
    const double det = s[0] * d[0] + s[1] * d[2];
    const double r = 1.0 / det;
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] *= r;
    return det;
}
__device__ __forceinline__ double det_of(const double (&s)[9]) {
    return s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
}
__device__ __forceinline__ double det_of(const double (&s)[4]) { return s[0] * s[3] - s[1] * s[2]; }

// ---------------------------------------------------------------------------------------------------------
// Elasticity tensors as functions of (i,j,k,l): evaluated only at the 21 (6) Voigt pairs, fully unrolled.
// ---------------------------------------------------------------------------------------------------------
template <int D>
struct PointState {
    double F[D * D], b[D * D], J;
    double Dv[D];     // electric displacement
    double g[D * D];  // H H^T (nearly-incompressible model only)
    double c[8];      // per-material scalar coefficients
};

template <int D, int MAT>
__device__ __forceinline__ double elasticity_ijkl(const PointState<D>& s, const MatParams& p, int i, int j, int k, int l) {
    const double II_ijkl = kd(i, j) * kd(k, l), II_s = kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k);
    if (MAT == MAT_LINEAR_ELASTIC) {
        return p.mu * II_s + p.lamb * II_ijkl;  // _LinearElastic_.h:49
    } else if (MAT == MAT_NEOHOOKEAN || MAT == MAT_ELECTRO_101) {
        // _NeoHookean_.h:45, _IsotropicElectroMechanics_101_.h:57
        return (p.mu / s.J - p.lamb * (s.J - 1.)) * II_s + p.lamb * (2. * s.J - 1.) * II_ijkl;
    } else if (MAT == MAT_MOONEY_RIVLIN || MAT == MAT_ELECTRO_105 || MAT == MAT_ELECTRO_108) {
        // _MooneyRivlin_.h:57-58
        const double* b = s.b;
        double v = 2.0 * p.mu2 / s.J * (2.0 * b[i * D + j] * b[k * D + l] - b[i * D + k] * b[j * D + l] - b[i * D + l] * b[j * D + k]) +
                   (2. * (p.mu1 + 2 * p.mu2) / s.J - p.lamb * (s.J - 1.)) * II_s + p.lamb * (2. * s.J - 1.) * II_ijkl;
        if (MAT == MAT_ELECTRO_108) {
            // C_elect, _IsotropicElectroMechanics_108_.h:87-88
            const double DD = s.c[0];
            v += 1. / p.eps_2 * (0.5 * DD * (II_ijkl + II_s) - kd(i, j) * s.Dv[k] * s.Dv[l] - s.Dv[i] * s.Dv[j] * kd(k, l));
        }
        return v;
    } else if (MAT == MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN) {
        // _NearlyIncompressibleMooneyRivlin_.h:62-83; (alpha,beta,kappa) = (mu1,mu2,mu3)
        const double alpha = p.mu1, beta = p.mu2, kappa = p.mu3;
        const double c0 = s.c[0], c1 = s.c[1], c2 = s.c[2], c3 = s.c[3], c4 = s.c[4], trb = s.c[5];
        const double *b = s.b, *g = s.g;
        const double II_ikjl = kd(i, k) * kd(j, l), II_iljk = kd(i, l) * kd(j, k);
        const double bI = b[i * D + j] * kd(k, l), Ib = kd(i, j) * b[k * D + l];
        const double gI = g[i * D + j] * kd(k, l), Ig = kd(i, j) * g[k * D + l];
        const double gI_ikjl = g[i * D + k] * kd(j, l), gI_iljk = g[i * D + l] * kd(j, k);
        const double Ig_ikjl = kd(i, k) * g[j * D + l], Ig_iljk = kd(i, l) * g[j * D + k];
        return -4 / 3. * alpha * c0 * (bI + Ib) + 4. * alpha / 9. * c0 * trb * II_ijkl + 2 / 3. * alpha * c0 * trb * (II_ikjl + II_iljk) +
               beta * c4 * c3 * (II_ijkl - II_ikjl - II_iljk) - 3. * beta * c4 * c1 * (gI + Ig) +
               3. * beta * c4 * c1 * (gI_ikjl + gI_iljk + Ig_ikjl + Ig_iljk) + 3. * beta * c4 * c2 * g[i * D + j] * g[k * D + l] +
               kappa * (2.0 * s.J - 1) * II_ijkl - kappa * (s.J - 1) * (II_ikjl + II_iljk);
    }
    return 0.0;
}

// W_coupling_ijk (internal-energy form) of the electro-mechanical models
template <int D, int MAT>
__device__ __forceinline__ double coupling_ijk(const PointState<D>& s, const MatParams& p, int i, int j, int k) {
    if (MAT == MAT_ELECTRO_101) return s.J / p.eps_1 * (kd(i, k) * s.Dv[j] + s.Dv[i] * kd(j, k));  // _101_.h:62-64
    if (MAT == MAT_ELECTRO_105) return s.J / p.eps_2 * (kd(i, k) * s.Dv[j] + s.Dv[i] * kd(j, k));  // _105_.h:86-90
    if (MAT == MAT_ELECTRO_108) return 1. / p.eps_2 * (kd(i, k) * s.Dv[j] + s.Dv[i] * kd(j, k) - kd(i, j) * s.Dv[k]);  // _108_.h:92-96
    return 0.0;
}

// MooneyRivlin Cauchy stress, shared by materials 0, 2, 5, 8, 9 (_MooneyRivlin_.h:41-51)
template <int D>
__device__ __forceinline__ void mooney_stress(const PointState<D>& s, const MatParams& p, double (&sig)[D * D]) {
    double trb = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) trb += s.b[i * D + i];
    if (D == 2) trb += 1.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double bb = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) bb += s.b[i * D + k] * s.b[k * D + j];
            sig[i * D + j] = 2. * p.mu1 / s.J * s.b[i * D + j] + 2. * p.mu2 / s.J * (trb * s.b[i * D + j] - bb) -
                             2. * (p.mu1 + 2 * p.mu2) / s.J * kd(i, j) + p.lamb * (s.J - 1) * kd(i, j);
        }
}

/*
 * Kinetic measures at one material point.
 *   F (D x D row-major), E (D; electro only).
 *   out: sig (D x D), Dv (D; electro), and, when WANT_H, hess (HT x HT row-major, HT = HS (+D electro)).
 */
template <int D, int MAT, bool WANT_H>
__device__ __forceinline__ void kinetic_measures(const double (&F)[D * D], const double* E, const MatParams& p, double (&sig)[D * D],
                                                 double* Dv, double* hess) {
    using VM = voigt_map<D>;
    constexpr int HS = VM::HS;
    constexpr bool EL = mat_traits<MAT>::electro;
    constexpr int HT = HS + (EL ? D : 0);
    PointState<D> s;
#pragma unroll
    for (int i = 0; i < D * D; ++i) s.F[i] = F[i];
    s.J = det_of(s.F);
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double v = 0;
#pragma unroll
            for (int k = 0; k < D; ++k) v += F[i * D + k] * F[j * D + k];
            s.b[i * D + j] = v;
        }
#pragma unroll
    for (int i = 0; i < D; ++i) s.Dv[i] = 0;
    double Wd[D * D];  // W_dielectric (electro)

    if (MAT == MAT_LINEAR_ELASTIC) {
        // _LinearElastic_.h:34-44
        double tre = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) tre += 0.5 * ((F[i * D + i] - 1.0) + (F[i * D + i] - 1.0));
        if (D == 2) tre += 1.;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double e = 0.5 * ((F[i * D + j] - kd(i, j)) + (F[j * D + i] - kd(j, i)));
                sig[i * D + j] = 2 * p.mu * e + p.lamb * tre * kd(i, j);
            }
    } else if (MAT == MAT_NEOHOOKEAN) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) sig[i * D + j] = p.mu / s.J * (s.b[i * D + j] - kd(i, j)) + p.lamb * (s.J - 1) * kd(i, j);
    } else if (MAT == MAT_MOONEY_RIVLIN || MAT == MAT_EXPLICIT_MOONEY_RIVLIN) {
        mooney_stress<D>(s, p, sig);
    } else if (MAT == MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN) {
        // _NearlyIncompressibleMooneyRivlin_.h:36-60
        const double alpha = p.mu1, beta = p.mu2, kappa = p.mu3;
        double Hc[D * D];
        if constexpr (D == 3) {
            double inv[D * D];
            invdet(s.F, inv);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j < D; ++j) Hc[i * D + j] = s.J * inv[j * D + i];
        } else {
            Hc[0] = F[3];
            Hc[1] = -F[2];
            Hc[2] = -F[1];
            Hc[3] = F[0];
        }
        double trb = 0, trg = 0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double v = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) v += Hc[i * D + k] * Hc[j * D + k];
                s.g[i * D + j] = v;
            }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            trb += s.b[i * D + i];
            trg += s.g[i * D + i];
        }
        if (D == 2) {
            trb += 1.;
            trg += s.J * s.J;
        }
        const double c0 = pow(s.J, -5. / 3.), c1 = sqrt(trg), c2 = 1. / c1, c3 = trg * c1, c4 = 1. / (s.J * s.J * s.J);
        s.c[0] = c0; s.c[1] = c1; s.c[2] = c2; s.c[3] = c3; s.c[4] = c4; s.c[5] = trb;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j)
                sig[i * D + j] = 2. * alpha * c0 * s.b[i * D + j] - 2. / 3. * alpha * c0 * trb * kd(i, j) + beta * c4 * c3 * kd(i, j) -
                                 3 * beta * c4 * c1 * s.g[i * D + j] + kappa * (s.J - 1.0) * kd(i, j);
    } else if (MAT == MAT_ELECTRO_101) {
        // _IsotropicElectroMechanics_101_.h:47-52
#pragma unroll
        for (int i = 0; i < D; ++i) s.Dv[i] = (p.eps_1 / s.J) * E[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                sig[i * D + j] = p.mu / s.J * (s.b[i * D + j] - kd(i, j)) + p.lamb * (s.J - 1) * kd(i, j) + s.J / p.eps_1 * s.Dv[i] * s.Dv[j];
                Wd[i * D + j] = s.J / p.eps_1 * kd(i, j);
            }
    } else if (MAT == MAT_ELECTRO_105) {
        // _IsotropicElectroMechanics_105_.h:51-72
        double binv[D * D], Wdinv[D * D];
        invdet(s.b, binv);
#pragma unroll
        for (int i = 0; i < D * D; ++i) Wd[i] = s.J / p.eps_1 * binv[i] + s.J / p.eps_2 * kd(i / D, i % D);
        invdet(Wd, Wdinv);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double v = 0;
#pragma unroll
            for (int j = 0; j < D; ++j) v += Wdinv[i * D + j] * E[j];
            s.Dv[i] = v;
        }
        mooney_stress<D>(s, p, sig);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) sig[i * D + j] += s.J / p.eps_2 * s.Dv[i] * s.Dv[j];
    } else if (MAT == MAT_ELECTRO_108 || MAT == MAT_EXPLICIT_ELECTRO_108) {
        // _IsotropicElectroMechanics_108_.h:50-72
        double DD = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            s.Dv[i] = p.eps_2 * E[i];
            DD += s.Dv[i] * s.Dv[i];
        }
        s.c[0] = DD;
        mooney_stress<D>(s, p, sig);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                sig[i * D + j] += 1. / p.eps_2 * (s.Dv[i] * s.Dv[j] - 0.5 * DD * kd(i, j));
                Wd[i * D + j] = 1. / p.eps_2 * kd(i, j);
            }
    }
    if (EL) {
#pragma unroll
        for (int i = 0; i < D; ++i) Dv[i] = s.Dv[i];
    }

    if (!WANT_H || !mat_traits<MAT>::has_tangent) return;

    if (!EL) {
        // hessian = voigt(elasticity): upper triangle, mirrored (Numeric.pyx:181-229)
#pragma unroll
        for (int I = 0; I < HS; ++I)
#pragma unroll
            for (int Jv = I; Jv < HS; ++Jv) {
                const int i = VM::i(I), j = VM::j(I), k = VM::i(Jv), l = VM::j(Jv);
                const double v = (k == l) ? elasticity_ijkl<D, MAT>(s, p, i, j, k, l)
                                          : 0.5 * (elasticity_ijkl<D, MAT>(s, p, i, j, k, l) + elasticity_ijkl<D, MAT>(s, p, i, j, l, k));
                hess[I * HT + Jv] = v;
                hess[Jv * HT + I] = v;
            }
    } else {
        // Legendre transform: _LegendreTransform_.h:67-82; LegendreTransform.py:22-41
        double Hd[D * D];
        invdet(Wd, Hd);
#pragma unroll
        for (int i = 0; i < D * D; ++i) Hd[i] = -Hd[i];
        // H_coupling_klj = - W_coupling_kli H_dielectric_ij
        double Hc[D * D * D];
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int l = 0; l < D; ++l)
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double v = 0;
#pragma unroll
                    for (int i = 0; i < D; ++i) v += coupling_ijk<D, MAT>(s, p, k, l, i) * Hd[i * D + j];
                    Hc[(k * D + l) * D + j] = -v;
                }
        // elastic block: voigt(W_elasticity_ijlm - W_coupling_ijk H_coupling_mlk)
#pragma unroll
        for (int I = 0; I < HS; ++I)
#pragma unroll
            for (int Jv = I; Jv < HS; ++Jv) {
                const int i = VM::i(I), j = VM::j(I), l = VM::i(Jv), m = VM::j(Jv);
                double v1 = elasticity_ijkl<D, MAT>(s, p, i, j, l, m), v2 = elasticity_ijkl<D, MAT>(s, p, i, j, m, l);
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double w = coupling_ijk<D, MAT>(s, p, i, j, k);
                    v1 -= w * Hc[(m * D + l) * D + k];
                    v2 -= w * Hc[(l * D + m) * D + k];
                }
                const double v = (l == m) ? v1 : 0.5 * (v1 + v2);
                hess[I * HT + Jv] = v;
                hess[Jv * HT + I] = v;
            }
        // coupling block: -voigt3(H_coupling) (HS x D) and its transpose (_helper_.h:84-94)
#pragma unroll
        for (int I = 0; I < HS; ++I)
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int i = VM::i(I), j = VM::j(I);
                const double v = (i == j) ? Hc[(i * D + i) * D + k] : 0.5 * (Hc[(i * D + j) * D + k] + Hc[(j * D + i) * D + k]);
                hess[I * HT + HS + k] = -v;
                hess[(HS + k) * HT + I] = -v;
            }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) hess[(HS + i) * HT + HS + j] = Hd[i * D + j];
    }
}

}  // namespace fl
