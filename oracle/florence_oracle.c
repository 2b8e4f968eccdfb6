/*
 * florence_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the element-assembly hot path of
 * romeric/florence (the Cython/Fastor "low level" assemblers).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * can check (and time) the CUDA path against the reference ALGORITHM.  Nothing under
 * florence_b200/ may import, link or call it.
 *
 * Parity status: the reference's own LL assembler cannot be compiled here (needs the
 * un-vendored Fastor headers and cblas.h), so this file restates it loop
 * by loop, and is pinned by tests/test_oracle_golden.py against
 *   (i)  outputs of the reference's own Fastor-free native code run in the build
 *        container (ComputeSparsityPattern, SparseAssemblyNative, RHSAssemblyNative), and
 *   (ii) outputs of the reference's own pure-numpy twins of the same formulas
 *        (MaterialLibrary/<Material>.py CauchyStress/Hessian, DisplacementFormulation.
 *        GetLocalStiffness, DisplacementPotentialFormulation.GetLocalStiffness, ...),
 * both committed as fixtures under tests/golden/ together with the generating script, and
 *   (iii) the reference's own native sparsity-pattern / CSR-scatter code compiled from /root/reference into
 *        oracle/_ref/libflorence_ref.so (oracle/ref_shim.cpp, tests/test_reference_natives.py).
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Third-party arithmetic that is absent from /root/reference: romeric/Fastor (unpinned,
 * CI clones master) -- its `voigt`, `determinant`, `inverse`, `matmul`, `einsum`,
 * `permutation`, `outer` are restated from the published definitions, cross-read with
 * the reference's Python twins (Florence/Tensor/Numeric.pyx:181-278 for Voigt); and CBLAS
 * dgemm (restated as a naive triple loop).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double Real;
typedef int64_t Integer;
typedef uint64_t UInteger;

#define FLO_API __attribute__((visibility("default")))

/* material numbers: Florence/FiniteElements/Assembly/_Assembly_/_LowLevelAssemblyExplicit_DF_DPF_.pyx:72-109 */
enum {
    MAT_EXPLICIT_MOONEY_RIVLIN = 0,
    MAT_NEOHOOKEAN = 1,
    MAT_MOONEY_RIVLIN = 2,
    MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN = 3,
    MAT_ELECTRO_101 = 4,
    MAT_ELECTRO_105 = 5,
    MAT_ELECTRO_106 = 6, /* not restated (out of scope) */
    MAT_ELECTRO_107 = 7, /* not restated */
    MAT_ELECTRO_108 = 8,
    MAT_EXPLICIT_ELECTRO_108 = 9,
    MAT_LINEAR_ELASTIC = 10
};

/* params layout (10 doubles), same order as the reference C signature
 * (_LowLevelAssemblyDF_.pyx:38-48): mu, mu1, mu2, mu3, mue, lamb, eps_1, eps_2, eps_3, eps_e */
enum { P_MU = 0, P_MU1, P_MU2, P_MU3, P_MUE, P_LAMB, P_EPS1, P_EPS2, P_EPS3, P_EPSE };

/* ------------------------------------------------------------------------------------------ */
/* small matrices: Florence/Tensor/_det_inv_.h:6-63 (inv2x2, inv3x3), :158-199 (invdet3x3)      */
/* ------------------------------------------------------------------------------------------ */
static Real invdet2(const Real *s, Real *d) {
    d[0] = +s[3]; d[1] = -s[1]; d[2] = -s[2]; d[3] = +s[0];
    Real det = s[0] * d[0] + s[1] * d[2];
    Real r = 1.0 / det;
    d[0] *= r; d[1] *= r; d[2] *= r; d[3] *= r;
    return det;
}

static Real invdet3(const Real *s, Real *d) {
    d[0] = +s[4] * s[8] - s[5] * s[7];
    d[1] = -s[1] * s[8] + s[2] * s[7];
    d[2] = +s[1] * s[5] - s[2] * s[4];
    d[3] = -s[3] * s[8] + s[5] * s[6];
    d[4] = +s[0] * s[8] - s[2] * s[6];
    d[5] = -s[0] * s[5] + s[2] * s[3];
    d[6] = +s[3] * s[7] - s[4] * s[6];
    d[7] = -s[0] * s[7] + s[1] * s[6];
    d[8] = +s[0] * s[4] - s[1] * s[3];
    Real det = s[0] * d[0] + s[1] * d[3] + s[2] * d[6];
    Real r = 1.0 / det;
    for (int i = 0; i < 9; ++i) d[i] *= r;
    return det;
}

static Real invdet(int d, const Real *s, Real *dst) { return d == 3 ? invdet3(s, dst) : invdet2(s, dst); }

static Real detN(int d, const Real *s) {
    if (d == 2) return s[0] * s[3] - s[1] * s[2];
    return s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
}

/* C(m x n) = A(m x k) * B(k x n), row-major: stands in for Florence/Tensor/_matmul_.h:240-289 and cblas_dgemm */
static void matmul(int m, int n, int k, const Real *A, const Real *B, Real *C) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            Real s = 0;
            for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
            C[i * n + j] = s;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Voigt maps: Florence/Tensor/Numeric.pyx:181-229 (rank 4), :251-278 (rank 3)                 */
/* ------------------------------------------------------------------------------------------ */
static const int VP3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {0, 2}, {1, 2}};
static const int VP2[3][2] = {{0, 0}, {1, 1}, {0, 1}};

static int voigt_size(int d) { return d == 3 ? 6 : 3; }
static const int (*voigt_pairs(int d))[2] { return d == 3 ? VP3 : VP2; }

#define T4(C, d, i, j, k, l) (C)[(((i) * (d) + (j)) * (d) + (k)) * (d) + (l)]
#define T3(C, d, i, j, k) (C)[((i) * (d) + (j)) * (d) + (k)]

/* H (hs x hs, leading dimension ld): upper triangle 0.5*(C_ijkl + C_ijlk), lower mirrored */
static void voigt4(int d, const Real *C, Real *H, int ld) {
    const int hs = voigt_size(d);
    const int(*vp)[2] = voigt_pairs(d);
    for (int I = 0; I < hs; ++I)
        for (int J = I; J < hs; ++J) {
            int i = vp[I][0], j = vp[I][1], k = vp[J][0], l = vp[J][1];
            Real v = (k == l) ? T4(C, d, i, j, k, l) : 0.5 * (T4(C, d, i, j, k, l) + T4(C, d, i, j, l, k));
            H[I * ld + J] = v;
            H[J * ld + I] = v;
        }
}

/* P (hs x d): P[I][k] = 0.5*(e_ijk + e_jik) */
static void voigt3(int d, const Real *e, Real *P) {
    const int hs = voigt_size(d);
    const int(*vp)[2] = voigt_pairs(d);
    for (int I = 0; I < hs; ++I) {
        int i = vp[I][0], j = vp[I][1];
        for (int k = 0; k < d; ++k)
            P[I * d + k] = (i == j) ? T3(e, d, i, i, k) : 0.5 * (T3(e, d, i, j, k) + T3(e, d, j, i, k));
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Legendre transform of (W_elasticity, W_coupling, W_dielectric) to the enthalpy Hessian:
 * Florence/MaterialLibrary/LLDispatch/CythonSource/_LegendreTransform_.h:67-82,
 * block layout _helper_.h:69-113; Python twin Florence/LegendreTransform/LegendreTransform.py:22-41 */
/* ------------------------------------------------------------------------------------------ */
static void legendre_hessian(int d, const Real *We, const Real *Wc, const Real *Wd, Real *H) {
    const int hs = voigt_size(d), n = hs + d;
    Real Hd[9], Hc[27], He[81], P[18];
    invdet(d, Wd, Hd);
    for (int i = 0; i < d * d; ++i) Hd[i] = -Hd[i];
    /* H_coupling_klj = - W_coupling_kli H_dielectric_ij */
    for (int k = 0; k < d; ++k)
        for (int l = 0; l < d; ++l)
            for (int j = 0; j < d; ++j) {
                Real s = 0;
                for (int i = 0; i < d; ++i) s += T3(Wc, d, k, l, i) * Hd[i * d + j];
                T3(Hc, d, k, l, j) = -s;
            }
    /* H_elasticity_ijlm = W_elasticity_ijlm - W_coupling_ijk H_coupling_mlk */
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j)
            for (int l = 0; l < d; ++l)
                for (int m = 0; m < d; ++m) {
                    Real s = 0;
                    for (int k = 0; k < d; ++k) s += T3(Wc, d, i, j, k) * T3(Hc, d, m, l, k);
                    T4(He, d, i, j, l, m) = T4(We, d, i, j, l, m) - s;
                }
    voigt4(d, He, H, n);
    voigt3(d, Hc, P);
    for (int I = 0; I < hs; ++I)
        for (int k = 0; k < d; ++k) {
            H[I * n + hs + k] = -P[I * d + k];
            H[(hs + k) * n + I] = -P[I * d + k];
        }
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) H[(hs + i) * n + hs + j] = Hd[i * d + j];
}

/* ------------------------------------------------------------------------------------------ */
/* material point kernels (a9)                                                                 */
/* ------------------------------------------------------------------------------------------ */
static inline Real kd(int i, int j) { return i == j ? 1.0 : 0.0; }

/* MooneyRivlin stress, shared by materials 0, 2, 5, 8, 9: _MooneyRivlin_.h:41-51 */
static void mooney_stress(int d, Real mu1, Real mu2, Real lamb, Real J, const Real *b, Real *s) {
    Real trb = 0, bb[9];
    for (int i = 0; i < d; ++i) trb += b[i * d + i];
    if (d == 2) trb += 1.0;
    matmul(d, d, d, b, b, bb);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j)
            s[i * d + j] = 2. * mu1 / J * b[i * d + j] + 2. * mu2 / J * (trb * b[i * d + j] - bb[i * d + j]) -
                           2. * (mu1 + 2 * mu2) / J * kd(i, j) + lamb * (J - 1) * kd(i, j);
}

/* MooneyRivlin elasticity tensor: _MooneyRivlin_.h:53-63 */
static void mooney_elasticity(int d, Real mu1, Real mu2, Real lamb, Real J, const Real *b, Real *C) {
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j)
            for (int k = 0; k < d; ++k)
                for (int l = 0; l < d; ++l)
                    T4(C, d, i, j, k, l) =
                        2.0 * mu2 / J * (2.0 * b[i * d + j] * b[k * d + l] - b[i * d + k] * b[j * d + l] - b[i * d + l] * b[j * d + k]) +
                        (2. * (mu1 + 2 * mu2) / J - lamb * (J - 1.)) * (kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k)) +
                        lamb * (2. * J - 1.) * kd(i, j) * kd(k, l);
}

/*
 * One material point.  F (d x d row-major), E (d) [electro only]; out: D (d) [electro], stress (d x d),
 * hessian (H x H row-major, H = 3/6 mechanics, 5/9 electro) -- hessian may be NULL (explicit path discards it).
 * Returns 0, or -1 for a material number that is not restated.
 */
FLO_API int flo_material_point(int material_number, int ndim, const Real *F, const Real *E, const Real *prm,
                               Real *D, Real *stress, Real *hessian) {
    const int d = ndim;
    Real b[9], C[81];
    const Real J = detN(d, F);
    /* b = F F^T */
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            Real s = 0;
            for (int k = 0; k < d; ++k) s += F[i * d + k] * F[j * d + k];
            b[i * d + j] = s;
        }

    switch (material_number) {
    case MAT_LINEAR_ELASTIC: {
        /* _LinearElastic_.h:24-57 */
        const Real mu = prm[P_MU], lamb = prm[P_LAMB];
        Real strain[9], tre = 0;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                strain[i * d + j] = 0.5 * ((F[i * d + j] - kd(i, j)) + (F[j * d + i] - kd(j, i)));
        for (int i = 0; i < d; ++i) tre += strain[i * d + i];
        if (d == 2) tre += 1.;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) stress[i * d + j] = 2 * mu * strain[i * d + j] + lamb * tre * kd(i, j);
        if (hessian) {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k)
                        for (int l = 0; l < d; ++l)
                            T4(C, d, i, j, k, l) = mu * (kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k)) + lamb * kd(i, j) * kd(k, l);
            voigt4(d, C, hessian, voigt_size(d));
        }
        return 0;
    }
    case MAT_NEOHOOKEAN: {
        /* _NeoHookean_.h:24-52 */
        const Real mu = prm[P_MU], lamb = prm[P_LAMB];
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) stress[i * d + j] = mu / J * (b[i * d + j] - kd(i, j)) + lamb * (J - 1) * kd(i, j);
        if (hessian) {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k)
                        for (int l = 0; l < d; ++l)
                            T4(C, d, i, j, k, l) = (mu / J - lamb * (J - 1.)) * (kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k)) +
                                                   lamb * (2. * J - 1.) * kd(i, j) * kd(k, l);
            voigt4(d, C, hessian, voigt_size(d));
        }
        return 0;
    }
    case MAT_EXPLICIT_MOONEY_RIVLIN: /* _ExplicitMooneyRivlin_.h:27-55 (stress only) */
    case MAT_MOONEY_RIVLIN: {        /* _MooneyRivlin_.h:27-70 */
        const Real mu1 = prm[P_MU1], mu2 = prm[P_MU2], lamb = prm[P_LAMB];
        mooney_stress(d, mu1, mu2, lamb, J, b, stress);
        if (hessian && material_number == MAT_MOONEY_RIVLIN) {
            mooney_elasticity(d, mu1, mu2, lamb, J, b, C);
            voigt4(d, C, hessian, voigt_size(d));
        }
        return 0;
    }
    case MAT_NEARLY_INCOMPRESSIBLE_MOONEY_RIVLIN: {
        /* _NearlyIncompressibleMooneyRivlin_.h:25-95; (alpha,beta,kappa) arrive as (mu1,mu2,mu3) in the
         * explicit wrapper (_LowLevelAssemblyExplicit_DF_DPF_.pyx:82-84) and as (mu1,mu2,lamb) in the implicit
         * one (AOT_Assembler.py:71-72): the caller passes them in prm[P_MU1], prm[P_MU2], prm[P_MU3]. */
        const Real alpha = prm[P_MU1], beta = prm[P_MU2], kappa = prm[P_MU3];
        Real Hc[9], g[9];
        /* cofactor H = J F^{-T} */
        if (d == 3) {
            Real inv[9];
            invdet3(F, inv);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Hc[i * 3 + j] = J * inv[j * 3 + i];
        } else {
            Hc[0] = F[3]; Hc[1] = -F[2]; Hc[2] = -F[1]; Hc[3] = F[0];
        }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                Real s = 0;
                for (int k = 0; k < d; ++k) s += Hc[i * d + k] * Hc[j * d + k];
                g[i * d + j] = s;
            }
        Real trb = 0, trg = 0;
        for (int i = 0; i < d; ++i) { trb += b[i * d + i]; trg += g[i * d + i]; }
        if (d == 2) { trb += 1.; trg += J * J; }
        const Real c0 = pow(J, -5. / 3.), c1 = sqrt(trg), c2 = 1. / c1, c3 = trg * c1, c4 = 1. / (J * J * J);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                stress[i * d + j] = 2. * alpha * c0 * b[i * d + j] - 2. / 3. * alpha * c0 * trb * kd(i, j) + beta * c4 * c3 * kd(i, j) -
                                    3 * beta * c4 * c1 * g[i * d + j] + kappa * (J - 1.0) * kd(i, j);
        if (hessian) {
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k)
                        for (int l = 0; l < d; ++l) {
                            const Real II_ijkl = kd(i, j) * kd(k, l), II_ikjl = kd(i, k) * kd(j, l), II_iljk = kd(i, l) * kd(j, k);
                            const Real bI = b[i * d + j] * kd(k, l), Ib = kd(i, j) * b[k * d + l];
                            const Real gI = g[i * d + j] * kd(k, l), Ig = kd(i, j) * g[k * d + l];
                            /* permutation<Index<i,k,j,l>>(A_ijkl): out(i,j,k,l) = A(i,k,j,l) i.e. g_ik I_jl etc. */
                            const Real gI_ikjl = g[i * d + k] * kd(j, l), gI_iljk = g[i * d + l] * kd(j, k);
                            const Real Ig_ikjl = kd(i, k) * g[j * d + l], Ig_iljk = kd(i, l) * g[j * d + k];
                            const Real gg = g[i * d + j] * g[k * d + l];
                            T4(C, d, i, j, k, l) = -4 / 3. * alpha * c0 * (bI + Ib) + 4. * alpha / 9. * c0 * trb * II_ijkl +
                                                   2 / 3. * alpha * c0 * trb * (II_ikjl + II_iljk) +
                                                   beta * c4 * c3 * (II_ijkl - II_ikjl - II_iljk) - 3. * beta * c4 * c1 * (gI + Ig) +
                                                   3. * beta * c4 * c1 * (gI_ikjl + gI_iljk + Ig_ikjl + Ig_iljk) + 3. * beta * c4 * c2 * gg +
                                                   kappa * (2.0 * J - 1) * II_ijkl - kappa * (J - 1) * (II_ikjl + II_iljk);
                        }
            voigt4(d, C, hessian, voigt_size(d));
        }
        return 0;
    }
    case MAT_ELECTRO_101: {
        /* _IsotropicElectroMechanics_101_.h:28-76; Python twin IsotropicElectroMechanics_101.py:38-85 */
        const Real mu = prm[P_MU], lamb = prm[P_LAMB], eps_1 = prm[P_EPS1];
        for (int i = 0; i < d; ++i) D[i] = (eps_1 / J) * E[i];
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                stress[i * d + j] = mu / J * (b[i * d + j] - kd(i, j)) + lamb * (J - 1) * kd(i, j) + J / eps_1 * D[i] * D[j];
        if (hessian) {
            Real Wc[27], Wd[9];
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k) {
                        for (int l = 0; l < d; ++l)
                            T4(C, d, i, j, k, l) = lamb * (2. * J - 1.) * kd(i, j) * kd(k, l) +
                                                   (mu / J - lamb * (J - 1)) * (kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k));
                        T3(Wc, d, i, j, k) = J / eps_1 * (kd(i, k) * D[j] + D[i] * kd(j, k));
                    }
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) Wd[i * d + j] = J / eps_1 * kd(i, j);
            legendre_hessian(d, C, Wc, Wd, hessian);
        }
        return 0;
    }
    case MAT_ELECTRO_105: {
        /* _IsotropicElectroMechanics_105_.h:28-100; Python twin IsotropicElectroMechanics_105.py:42-102 */
        const Real mu1 = prm[P_MU1], mu2 = prm[P_MU2], lamb = prm[P_LAMB], eps_1 = prm[P_EPS1], eps_2 = prm[P_EPS2];
        Real binv[9], Wd[9], Wdinv[9];
        invdet(d, b, binv);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) Wd[i * d + j] = J / eps_1 * binv[i * d + j] + J / eps_2 * kd(i, j);
        invdet(d, Wd, Wdinv);
        for (int i = 0; i < d; ++i) {
            Real s = 0;
            for (int j = 0; j < d; ++j) s += Wdinv[i * d + j] * E[j];
            D[i] = s;
        }
        mooney_stress(d, mu1, mu2, lamb, J, b, stress);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) stress[i * d + j] += J / eps_2 * D[i] * D[j];
        if (hessian) {
            Real Wc[27];
            mooney_elasticity(d, mu1, mu2, lamb, J, b, C);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k) T3(Wc, d, i, j, k) = J / eps_2 * (kd(i, k) * D[j] + D[i] * kd(j, k));
            legendre_hessian(d, C, Wc, Wd, hessian);
        }
        return 0;
    }
    case MAT_EXPLICIT_ELECTRO_108: /* _ExplicitIsotropicElectroMechanics_108_.h:31-72 (D, stress only) */
    case MAT_ELECTRO_108: {        /* _IsotropicElectroMechanics_108_.h:31-105; Python twin IsotropicElectroMechanics_108.py:39-105 */
        const Real mu1 = prm[P_MU1], mu2 = prm[P_MU2], lamb = prm[P_LAMB], eps_2 = prm[P_EPS2];
        Real DD = 0;
        for (int i = 0; i < d; ++i) { D[i] = eps_2 * E[i]; }
        for (int i = 0; i < d; ++i) DD += D[i] * D[i];
        mooney_stress(d, mu1, mu2, lamb, J, b, stress);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) stress[i * d + j] += 1. / eps_2 * (D[i] * D[j] - 0.5 * DD * kd(i, j));
        if (hessian && material_number == MAT_ELECTRO_108) {
            Real Wc[27], Wd[9];
            mooney_elasticity(d, mu1, mu2, lamb, J, b, C);
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j)
                    for (int k = 0; k < d; ++k) {
                        for (int l = 0; l < d; ++l)
                            T4(C, d, i, j, k, l) += 1. / eps_2 * (0.5 * DD * (kd(i, j) * kd(k, l) + kd(i, k) * kd(j, l) + kd(i, l) * kd(j, k)) -
                                                                  kd(i, j) * D[k] * D[l] - D[i] * D[j] * kd(k, l));
                        T3(Wc, d, i, j, k) = 1. / eps_2 * (kd(i, k) * D[j] + D[i] * kd(j, k) - kd(i, j) * D[k]);
                    }
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) Wd[i * d + j] = 1. / eps_2 * kd(i, j);
            legendre_hessian(d, C, Wc, Wd, hessian);
        }
        return 0;
    }
    default:
        return -1;
    }
}

/* Does the material produce a tangent?  (ExplicitMooneyRivlin / Explicit_108 are stress-only.) */
static int material_is_electro(int m) { return m == MAT_ELECTRO_101 || m == MAT_ELECTRO_105 || m == MAT_ELECTRO_108 || m == MAT_EXPLICIT_ELECTRO_108; }

/* ------------------------------------------------------------------------------------------ */
/* kinematics                                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* Jm is the reference table Jm[k][a][g] (d x npe x ng).  current_Jm[k][a] for one g:
 * _KinematicMeasures_.h:82-86 / _LowLevelAssemblyExplicit_DF_DPF_.h:521-531 */
static void gather_Jm(int d, int npe, int ng, const Real *Jm, int g, Real *cur) {
    for (int k = 0; k < d; ++k)
        for (int a = 0; a < npe; ++a) cur[k * npe + a] = Jm[(k * npe + a) * ng + g];
}

/*
 * One Gauss point of KinematicMeasures (implicit, _KinematicMeasures_.h:64-119) or KinematicMeasures__
 * (explicit, _LowLevelAssemblyExplicit_DF_DPF_.h:66-104).  X, x are (npe x d).  Outputs: sp (d x npe) spatial
 * gradient, F (d x d), returns J_x determinant and J_X determinant through pointers.
 */
static void kinematics_gauss(int d, int npe, const Real *cur_Jm, const Real *X, const Real *x, Real *matgrad /* d x npe */,
                             Real *sp /* d x npe */, Real *F, Real *detX, Real *detx) {
    Real PX[9], Px[9], iPX[9], iPx[9], Ft[9];
    matmul(d, d, npe, cur_Jm, X, PX);
    matmul(d, d, npe, cur_Jm, x, Px);
    *detX = invdet(d, PX, iPX);
    *detx = invdet(d, Px, iPx);
    matmul(d, npe, d, iPX, cur_Jm, matgrad);
    matmul(d, npe, d, iPx, cur_Jm, sp);
    matmul(d, d, npe, matgrad, x, Ft);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) F[i * d + j] = Ft[j * d + i];
}

/* ------------------------------------------------------------------------------------------ */
/* explicit (matrix-free) internal force: _LowLevelAssemblyExplicit_DF_DPF_.h:215-454 (2-D), :458-717 (3-D) */
/* ------------------------------------------------------------------------------------------ */
FLO_API int flo_assemble_explicit(const Real *points, const UInteger *elements, const Real *Eulerx, const Real *Eulerp,
                                  const Real *Jm, const Real *AllGauss, Integer ndim, Integer nvar, Integer ngauss,
                                  Integer elem_begin, Integer elem_end, Integer nodeperelem, Real *T, const Real *prm,
                                  int material_number, int formulation_number) {
    const int d = (int)ndim, npe = (int)nodeperelem, ng = (int)ngauss, nv = (int)nvar;
    const int ndof = nv * npe;
    Real *X = malloc(sizeof(Real) * npe * d), *x = malloc(sizeof(Real) * npe * d), *phi = malloc(sizeof(Real) * npe);
    Real *cur = malloc(sizeof(Real) * d * npe * ng), *mg = malloc(sizeof(Real) * d * npe), *sp = malloc(sizeof(Real) * d * npe);
    Real *loc = malloc(sizeof(Real) * ndof), *trac = malloc(sizeof(Real) * ndof);
    int rc = 0;
    for (int g = 0; g < ng; ++g) gather_Jm(d, npe, ng, Jm, g, cur + (size_t)g * d * npe);

    for (Integer e = elem_begin; e < elem_end && rc == 0; ++e) {
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int j = 0; j < d; ++j) { X[a * d + j] = points[n * d + j]; x[a * d + j] = Eulerx[n * d + j]; }
            phi[a] = Eulerp[n];
        }
        memset(trac, 0, sizeof(Real) * ndof);
        for (int g = 0; g < ng; ++g) {
            Real F[9], Ef[3] = {0, 0, 0}, D[3] = {0, 0, 0}, s[9], dX, dx;
            kinematics_gauss(d, npe, cur + (size_t)g * d * npe, X, x, mg, sp, F, &dX, &dx);
            /* AVX build: detJ = w |det J_x| regardless of `update` (:96) */
            const Real detJ = AllGauss[g] * fabs(dx);
            if (formulation_number == 1) {
                for (int k = 0; k < d; ++k) {
                    Real s_ = 0;
                    for (int a = 0; a < npe; ++a) s_ += sp[k * npe + a] * phi[a];
                    Ef[k] = -s_;
                }
            }
            rc = flo_material_point(material_number, d, F, Ef, prm, D, s, NULL);
            if (rc) break;
            for (int a = 0; a < npe; ++a) {
                for (int i = 0; i < d; ++i) {
                    Real t = 0;
                    for (int j = 0; j < d; ++j) {
                        /* the reference uses the upper triangle of sigma (s12 for both (0,1) and (1,0)), :624-640 */
                        const Real sij = (j <= i) ? s[j * d + i] : s[i * d + j];
                        t += sp[j * npe + a] * sij;
                    }
                    loc[a * nv + i] = t;
                }
                if (formulation_number == 1) {
                    Real t = 0;
                    for (int j = 0; j < d; ++j) t += sp[j * npe + a] * D[j];
                    loc[a * nv + d] = t;
                }
            }
            for (int i = 0; i < ndof; ++i) trac[i] += loc[i] * detJ;
        }
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int i = 0; i < nv; ++i) T[n * nv + i] += trac[a * nv + i];
        }
    }
    free(X); free(x); free(phi); free(cur); free(mg); free(sp); free(loc); free(trac);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* implicit integrands                                                                         */
/* ------------------------------------------------------------------------------------------ */
/* B (ndof x H): _ConstitutiveStiffnessDF_.h:32-77, _ConstitutiveStiffnessDPF_.h:42-96.  sg is (npe x d). */
static void fill_B(Real *B, const Real *sg, int d, int nv, int npe, int H) {
    for (int a = 0; a < npe; ++a) {
        Real *Ba = B + (size_t)a * H * nv;
        if (d == 3) {
            const Real a0 = sg[a * 3], a1 = sg[a * 3 + 1], a2 = sg[a * 3 + 2];
            Ba[0] = a0; Ba[H + 1] = a1; Ba[2 * (H + 1)] = a2;
            Ba[H + 5] = a2; Ba[2 * H + 5] = a1;
            Ba[4] = a2; Ba[2 * H + 4] = a0;
            Ba[3] = a1; Ba[H + 3] = a0;
            if (nv == 4) { Ba[3 * H + 6] = a0; Ba[3 * H + 7] = a1; Ba[3 * H + 8] = a2; }
        } else {
            const Real a0 = sg[a * 2], a1 = sg[a * 2 + 1];
            Ba[0] = a0; Ba[H + 1] = a1;
            Ba[2] = a1; Ba[H + 2] = a0;
            if (nv == 3) { Ba[2 * H + 3] = a0; Ba[2 * H + 4] = a1; }
        }
    }
}

/* total traction vector: _ConstitutiveStiffnessDF_.h:15-29, ...DPF_.h:17-39 */
static void total_traction(Real *t, const Real *s, const Real *D, int d, int electro) {
    if (d == 3) {
        t[0] = s[0]; t[1] = s[4]; t[2] = s[8]; t[3] = s[1]; t[4] = s[2]; t[5] = s[5];
        if (electro) { t[6] = D[0]; t[7] = D[1]; t[8] = D[2]; }
    } else {
        t[0] = s[0]; t[1] = s[3]; t[2] = s[1];
        if (electro) { t[3] = D[0]; t[4] = D[1]; }
    }
}

/* geometric stiffness: _GeometricStiffness_.h:61-144.  sg (npe x d) */
static void geometric_stiffness(Real *Kg, const Real *sg, const Real *s, Real detJ, int d, int nv, int npe) {
    const int ndof = nv * npe;
    for (int a = 0; a < npe; ++a)
        for (int b = 0; b < npe; ++b) {
            Real dum;
            if (d == 3) {
                const Real a0 = sg[a * 3], a1 = sg[a * 3 + 1], a2 = sg[a * 3 + 2];
                const Real b0 = sg[b * 3], b1 = sg[b * 3 + 1], b2 = sg[b * 3 + 2];
                const Real s00 = s[0], s01 = s[1], s02 = s[2], s11 = s[4], s12 = s[5], s22 = s[8];
                dum = a0 * (s00 * b0 + s01 * b1 + s02 * b2) + a1 * (s01 * b0 + s11 * b1 + s12 * b2) + a2 * (s02 * b0 + s12 * b1 + s22 * b2);
            } else {
                const Real a0 = sg[a * 2], a1 = sg[a * 2 + 1], b0 = sg[b * 2], b1 = sg[b * 2 + 1];
                const Real s00 = s[0], s01 = s[1], s11 = s[3];
                dum = a0 * (s00 * b0 + s01 * b1) + a1 * (s01 * b0 + s11 * b1);
            }
            for (int i = 0; i < d; ++i) Kg[(size_t)(a * nv + i) * ndof + (b * nv + i)] += dum * detJ;
        }
}

/* binary_locate: SparseAssemblyNative.h:18-26 */
static int binary_locate(const int *first, int n, int val) {
    int lo = 0;
    while (1) {
        if (val == first[lo]) return lo;
        if (n == 1) return lo;
        int half = n / 2;
        if (val < first[lo + half]) n = half;
        else { lo += half; n -= half; }
    }
}

/*
 * Element stiffness/traction of one element, the body of the loops
 * _LowLevelAssemblyDF_.h:69-131 and _LowLevelAssemblyDPF_.h:73-150.
 * geometric: 0 for LinearElastic (AOT_Assembler.py:79-86 strips it), 1 otherwise.
 */
static int element_implicit(int d, int nv, int npe, int ng, int H, const Real *curJm /* ng x d x npe */, const Real *AllGauss,
                            const Real *X, const Real *x, const Real *phi, const Real *prm, int material_number, int update,
                            int geometric, Real *K /* ndof^2 */, Real *Tr /* ndof */, Real *work) {
    const int ndof = nv * npe, electro = (nv == d + 1);
    Real *B = work, *HBT = B + (size_t)H * ndof, *BDB = HBT + (size_t)H * ndof, *mg = BDB + (size_t)ndof * ndof;
    Real *sp = mg + d * npe, *sg = sp + d * npe, *Kg = sg + d * npe;
    memset(K, 0, sizeof(Real) * ndof * ndof);
    memset(Tr, 0, sizeof(Real) * ndof);
    memset(B, 0, sizeof(Real) * H * ndof);
    if (geometric) memset(Kg, 0, sizeof(Real) * ndof * ndof);
    for (int g = 0; g < ng; ++g) {
        Real F[9], Ef[3] = {0, 0, 0}, D[3] = {0, 0, 0}, s[9], hess[81], t[9], dX, dx;
        kinematics_gauss(d, npe, curJm + (size_t)g * d * npe, X, x, mg, sp, F, &dX, &dx);
        /* _KinematicMeasures_.h:94-99 */
        const Real detJ = AllGauss[g] * fabs(update == 1 ? dx : dX);
        for (int a = 0; a < npe; ++a)
            for (int k = 0; k < d; ++k) sg[a * d + k] = sp[k * npe + a];
        if (electro) {
            /* _LowLevelAssemblyDPF_.h:103-112 */
            for (int k = 0; k < d; ++k) {
                Real s_ = 0;
                for (int a = 0; a < npe; ++a) s_ += sg[a * d + k] * phi[a];
                Ef[k] = -s_;
            }
        }
        if (flo_material_point(material_number, d, F, Ef, prm, D, s, hess)) return -1;
        fill_B(B, sg, d, nv, npe, H);
        /* HBT = H * B^T (H x ndof);  BDB = B * HBT: _ConstitutiveStiffnessDF_.h:113-117 */
        for (int i = 0; i < H; ++i)
            for (int j = 0; j < ndof; ++j) {
                Real v = 0;
                for (int l = 0; l < H; ++l) v += hess[i * H + l] * B[j * H + l];
                HBT[i * ndof + j] = v;
            }
        matmul(ndof, ndof, H, B, HBT, BDB);
        for (int i = 0; i < ndof * ndof; ++i) K[i] += BDB[i] * detJ;
        if (update == 1) {
            total_traction(t, s, D, d, electro);
            for (int i = 0; i < ndof; ++i) {
                Real tmp = 0;
                for (int j = 0; j < H; ++j) tmp += B[i * H + j] * t[j];
                Tr[i] += tmp * detJ;
            }
        }
        if (geometric) geometric_stiffness(Kg, sg, s, detJ, d, nv, npe);
    }
    if (geometric)
        for (int i = 0; i < ndof * ndof; ++i) K[i] += Kg[i];
    return 0;
}

/*
 * Implicit global assembly: _LowLevelAssemblyDF_.h:8-176 / _LowLevelAssemblyDPF_.h:45-198, scatter
 * _MassIntegrand_.h:69-166 + SparseAssemblyNative.h:32-111.
 * mode 0 = COO triplets (recompute_sparsity_pattern); I, J, V have ndof^2*nelem entries (whole-mesh indexing).
 * mode 1 = CSR through slot maps (data_local_indices / data_global_indices); I=indptr, J=indices.
 * mode 2 = CSR by binary search (squeeze_sparsity_pattern), using sorted_elements / sorter.
 */
FLO_API int flo_assemble_implicit(const Real *points, const UInteger *elements, const Real *Eulerx, const Real *Eulerp,
                                  const Real *Jm, const Real *AllGauss, Integer ndim, Integer nvar, Integer ngauss,
                                  Integer elem_begin, Integer elem_end, Integer nodeperelem, Integer H_VoigtSize,
                                  Integer requires_geometry_update, int *I, int *J, Real *V, Real *T, int mode,
                                  const int *data_local_indices, const int *data_global_indices,
                                  const UInteger *sorted_elements, const Integer *sorter, const Real *prm, int material_number) {
    const int d = (int)ndim, npe = (int)nodeperelem, ng = (int)ngauss, nv = (int)nvar, H = (int)H_VoigtSize;
    const int ndof = nv * npe;
    const size_t cap = (size_t)ndof * ndof;
    const int geometric = (material_number != MAT_LINEAR_ELASTIC);
    Real *X = malloc(sizeof(Real) * npe * d), *x = malloc(sizeof(Real) * npe * d), *phi = malloc(sizeof(Real) * npe);
    Real *cur = malloc(sizeof(Real) * d * npe * ng);
    Real *K = malloc(sizeof(Real) * cap), *Tr = malloc(sizeof(Real) * ndof);
    Real *work = malloc(sizeof(Real) * (2 * (size_t)H * ndof + 2 * cap + 3 * (size_t)d * npe));
    int *rc_glob = malloc(sizeof(int) * ndof), *rc_loc = malloc(sizeof(int) * ndof);
    int rc = 0;
    for (int g = 0; g < ng; ++g) gather_Jm(d, npe, ng, Jm, g, cur + (size_t)g * d * npe);

    for (Integer e = elem_begin; e < elem_end; ++e) {
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int j = 0; j < d; ++j) { X[a * d + j] = points[n * d + j]; x[a * d + j] = Eulerx[n * d + j]; }
            phi[a] = Eulerp[n];
        }
        rc = element_implicit(d, nv, npe, ng, H, cur, AllGauss, X, x, phi, prm, material_number, (int)requires_geometry_update,
                              geometric, K, Tr, work);
        if (rc) break;
        if (mode == 0) {
            /* fill_triplet: _MassIntegrand_.h:69-110 */
            size_t nc = cap * (size_t)e;
            for (int r = 0; r < ndof; ++r) {
                const int gr = (int)(nv * elements[e * npe + r / nv] + r % nv);
                for (int c = 0; c < ndof; ++c) {
                    I[nc] = gr;
                    J[nc] = (int)(nv * elements[e * npe + c / nv] + c % nv);
                    V[nc] = K[(size_t)r * ndof + c];
                    ++nc;
                }
            }
        } else if (mode == 1) {
            /* SparseAssemblyNativeCSR_: SparseAssemblyNative.h:32-45 */
            for (size_t i = 0; i < cap; ++i) V[data_global_indices[cap * e + i]] += K[data_local_indices[cap * e + i]];
        } else {
            /* SparseAssemblyNativeCSR_RecomputeDataIndex_: SparseAssemblyNative.h:49-111 (I=indptr, J=indices) */
            for (int c = 0; c < npe; ++c)
                for (int n = 0; n < nv; ++n) {
                    rc_glob[nv * c + n] = (int)(nv * sorted_elements[e * npe + c]) + n;
                    rc_loc[nv * c + n] = (int)sorter[e * npe + c] * nv + n;
                }
            for (int i = 0; i < ndof; ++i) {
                const int row0 = I[rc_glob[i]], nnz = I[rc_glob[i] + 1] - row0;
                for (int j = 0; j < ndof; ++j) {
                    const int it = binary_locate(J + row0, nnz, rc_glob[j]);
                    V[row0 + it] += K[(size_t)rc_loc[i] * ndof + rc_loc[j]];
                }
            }
        }
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int i = 0; i < nv; ++i) T[n * nv + i] += Tr[a * nv + i];
        }
    }
    free(X); free(x); free(phi); free(cur); free(K); free(Tr); free(work); free(rc_glob); free(rc_loc);
    return rc;
}

/* element-level entry (for golden checks against DisplacementFormulation.GetLocalStiffness) */
FLO_API int flo_element_implicit(const Real *X, const Real *x, const Real *phi, const Real *Jm, const Real *AllGauss, int ndim,
                                 int nvar, int ngauss, int nodeperelem, int H, int update, int geometric, const Real *prm,
                                 int material_number, Real *K, Real *Tr) {
    const int ndof = nvar * nodeperelem;
    const size_t cap = (size_t)ndof * ndof;
    Real *cur = malloc(sizeof(Real) * ndim * nodeperelem * ngauss);
    Real *work = malloc(sizeof(Real) * (2 * (size_t)H * ndof + 2 * cap + 3 * (size_t)ndim * nodeperelem));
    for (int g = 0; g < ngauss; ++g) gather_Jm(ndim, nodeperelem, ngauss, Jm, g, cur + (size_t)g * ndim * nodeperelem);
    int rc = element_implicit(ndim, nvar, nodeperelem, ngauss, H, cur, AllGauss, X, x, phi, prm, material_number, update, geometric, K, Tr, work);
    free(cur); free(work);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Laplacian: _LowLevelAssemblyPerfectLaplacian_.h:40-74 (kinematics), :246-420 (3-D), 2-D analogue; e = -material.e */
/* ------------------------------------------------------------------------------------------ */
FLO_API int flo_assemble_laplacian(const Real *points, const UInteger *elements, const Real *Jm, const Real *AllGauss, Integer ndim,
                                   Integer ngauss, Integer elem_begin, Integer elem_end, Integer nodeperelem, int *I, int *J, Real *V,
                                   const Real *e_tensor, int is_hessian_symmetric, int mode, const int *data_local_indices,
                                   const int *data_global_indices, const UInteger *sorted_elements, const Integer *sorter) {
    const int d = (int)ndim, npe = (int)nodeperelem, ng = (int)ngauss;
    const size_t cap = (size_t)npe * npe;
    Real *X = malloc(sizeof(Real) * npe * d), *cur = malloc(sizeof(Real) * d * npe * ng);
    Real *mg = malloc(sizeof(Real) * d * npe), *eM = malloc(sizeof(Real) * d * npe);
    Real *K = malloc(sizeof(Real) * cap);
    for (int g = 0; g < ng; ++g) gather_Jm(d, npe, ng, Jm, g, cur + (size_t)g * d * npe);
    for (Integer e = elem_begin; e < elem_end; ++e) {
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int j = 0; j < d; ++j) X[a * d + j] = points[n * d + j];
        }
        memset(K, 0, sizeof(Real) * cap);
        for (int g = 0; g < ng; ++g) {
            Real PX[9], iPX[9];
            const Real *cj = cur + (size_t)g * d * npe;
            matmul(d, d, npe, cj, X, PX);
            const Real detJ = AllGauss[g] * fabs(invdet(d, PX, iPX));
            matmul(d, npe, d, iPX, cj, mg);
            matmul(d, npe, d, e_tensor, mg, eM);
            /* symmetric Hessian: upper triangle only, mirrored after the Gauss loop (:338-381) */
            for (int a = 0; a < npe; ++a)
                for (int b = (is_hessian_symmetric ? a : 0); b < npe; ++b) {
                    Real v = 0;
                    for (int k = 0; k < d; ++k) v += mg[k * npe + a] * eM[k * npe + b];
                    K[(size_t)a * npe + b] += v * detJ;
                }
        }
        if (is_hessian_symmetric)
            for (int a = 0; a < npe; ++a)
                for (int b = a; b < npe; ++b) K[(size_t)b * npe + a] = K[(size_t)a * npe + b];
        if (mode == 0) {
            size_t nc = cap * (size_t)e;
            for (int r = 0; r < npe; ++r)
                for (int c = 0; c < npe; ++c) {
                    I[nc] = (int)elements[e * npe + r];
                    J[nc] = (int)elements[e * npe + c];
                    V[nc] = K[(size_t)r * npe + c];
                    ++nc;
                }
        } else if (mode == 1) {
            for (size_t i = 0; i < cap; ++i) V[data_global_indices[cap * e + i]] += K[data_local_indices[cap * e + i]];
        } else {
            for (int i = 0; i < npe; ++i) {
                const int gi = (int)sorted_elements[e * npe + i], li = (int)sorter[e * npe + i];
                const int row0 = I[gi], nnz = I[gi + 1] - row0;
                for (int j = 0; j < npe; ++j) {
                    const int gj = (int)sorted_elements[e * npe + j], lj = (int)sorter[e * npe + j];
                    V[row0 + binary_locate(J + row0, nnz, gj)] += K[(size_t)li * npe + lj];
                }
            }
        }
    }
    free(X); free(cur); free(mg); free(eM); free(K);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* mass: _MassIntegrand_.h:249-395 with constant_mass_integrand = rho N N^T (VariationalPrinciple.py:254-290)   */
/* lumped (mass_type 0): row sums scattered like T.  consistent (1): COO triplets.             */
/* ------------------------------------------------------------------------------------------ */
FLO_API int flo_assemble_mass(const Real *points, const UInteger *elements, const Real *bases /* npe x ng */, const Real *Jm,
                              const Real *AllGauss, Integer ndim, Integer nvar, Integer ngauss, Integer elem_begin, Integer elem_end,
                              Integer nodeperelem, Real rho, int mass_type, Real *mass /* nnode*nvar */, int *I, int *J, Real *V) {
    const int d = (int)ndim, npe = (int)nodeperelem, ng = (int)ngauss, nv = (int)nvar, ndof = nv * npe;
    const size_t cap = (size_t)ndof * ndof;
    Real *X = malloc(sizeof(Real) * npe * d), *cur = malloc(sizeof(Real) * d * npe * ng);
    Real *cmi = calloc(cap * ng, sizeof(Real)), *M = malloc(sizeof(Real) * cap);
    for (int g = 0; g < ng; ++g) {
        gather_Jm(d, npe, ng, Jm, g, cur + (size_t)g * d * npe);
        for (int a = 0; a < npe; ++a)
            for (int b = 0; b < npe; ++b)
                for (int i = 0; i < d; ++i)
                    cmi[(size_t)g * cap + (size_t)(a * nv + i) * ndof + (b * nv + i)] = rho * bases[a * ng + g] * bases[b * ng + g];
    }
    for (Integer e = elem_begin; e < elem_end; ++e) {
        for (int a = 0; a < npe; ++a) {
            const UInteger n = elements[e * npe + a];
            for (int j = 0; j < d; ++j) X[a * d + j] = points[n * d + j];
        }
        memset(M, 0, sizeof(Real) * cap);
        for (int g = 0; g < ng; ++g) {
            Real PX[9];
            matmul(d, d, npe, cur + (size_t)g * d * npe, X, PX);
            const Real detJ = AllGauss[g] * fabs(detN(d, PX));
            for (size_t i = 0; i < cap; ++i) M[i] += cmi[(size_t)g * cap + i] * detJ;
        }
        if (mass_type == 0) {
            for (int a = 0; a < npe; ++a)
                for (int i = 0; i < nv; ++i) {
                    Real s = 0;
                    for (int c = 0; c < ndof; ++c) s += M[(size_t)(a * nv + i) * ndof + c];
                    mass[elements[e * npe + a] * nv + i] += s;
                }
        } else {
            size_t nc = cap * (size_t)e;
            for (int r = 0; r < ndof; ++r)
                for (int c = 0; c < ndof; ++c) {
                    I[nc] = (int)(nv * elements[e * npe + r / nv] + r % nv);
                    J[nc] = (int)(nv * elements[e * npe + c / nv] + c % nv);
                    V[nc] = M[(size_t)r * ndof + c];
                    ++nc;
                }
        }
    }
    free(X); free(cur); free(cmi); free(M);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sparsity pattern: ComputeSparsityPattern.h:22-62 (+ the numpy glue of ComputeSparsityPattern.pyx:44-112),
 * data indices ComputeSparsityPattern.h:67-122.  elements here are uint64 (nelem x npe) in mesh order.         */
/* ------------------------------------------------------------------------------------------ */
static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

/* pass 1: indptr (nvar*nnode+1).  Returns nnz.  pass 2 (indices != NULL) fills indices. */
FLO_API int64_t flo_sparsity_pattern(const UInteger *elements, Integer nelem, Integer nodeperelem, Integer nnode, Integer nvar,
                                     int *indptr, int *indices) {
    const int npe = (int)nodeperelem, nv = (int)nvar;
    /* node -> elements inversion (idx_start / elem_container of the .pyx) */
    int64_t *start = calloc(nnode + 1, sizeof(int64_t));
    for (Integer i = 0; i < nelem * npe; ++i) start[elements[i] + 1]++;
    for (Integer n = 0; n < nnode; ++n) start[n + 1] += start[n];
    int *cont = malloc(sizeof(int) * (size_t)(nelem * npe));
    int64_t *fill = malloc(sizeof(int64_t) * nnode);
    memcpy(fill, start, sizeof(int64_t) * nnode);
    for (Integer e = 0; e < nelem; ++e)
        for (int a = 0; a < npe; ++a) cont[fill[elements[e * npe + a]]++] = (int)e;
    int64_t nnz = 0;
    size_t capn = 64;
    int *loc = malloc(sizeof(int) * capn);
    indptr[0] = 0;
    for (Integer n = 0; n < nnode; ++n) {
        const size_t cnt = (size_t)(start[n + 1] - start[n]) * npe;
        if (cnt > capn) { capn = cnt * 2; loc = realloc(loc, sizeof(int) * capn); }
        size_t c = 0;
        for (int64_t j = start[n]; j < start[n + 1]; ++j)
            for (int k = 0; k < npe; ++k) loc[c++] = (int)elements[(Integer)cont[j] * npe + k];
        qsort(loc, c, sizeof(int), cmp_int);
        size_t u = 0;
        for (size_t k = 0; k < c; ++k)
            if (k == 0 || loc[k] != loc[k - 1]) loc[u++] = loc[k];
        for (int j = 0; j < nv; ++j) {
            if (indices)
                for (size_t k = 0; k < u; ++k)
                    for (int l = 0; l < nv; ++l) indices[nnz + (int64_t)k * nv + l] = nv * loc[k] + l;
            nnz += (int64_t)u * nv;
            indptr[n * nv + j + 1] = (int)nnz;
        }
    }
    free(start); free(cont); free(fill); free(loc);
    return nnz;
}

/* data_local_indices / data_global_indices: ComputeSparsityPattern.h:67-122 with sorter = argsort(elements, axis=1)
 * (ComputeSparsityPattern.pyx:51-52; numpy's default quicksort is not stable but nodes within an element are unique). */
FLO_API int flo_data_indices(const UInteger *elements, Integer nelem, Integer nodeperelem, Integer nvar, const int *indptr,
                             const int *indices, int *data_local_indices, int *data_global_indices, Integer *sorter_out,
                             UInteger *sorted_elements_out) {
    const int npe = (int)nodeperelem, nv = (int)nvar, ndof = nv * npe;
    const size_t cap = (size_t)ndof * ndof;
    int *srt = malloc(sizeof(int) * npe), *rg = malloc(sizeof(int) * ndof), *rl = malloc(sizeof(int) * ndof);
    for (Integer e = 0; e < nelem; ++e) {
        for (int a = 0; a < npe; ++a) srt[a] = a;
        /* insertion argsort by node number */
        for (int a = 1; a < npe; ++a) {
            int v = srt[a], p = a - 1;
            while (p >= 0 && elements[e * npe + srt[p]] > elements[e * npe + v]) { srt[p + 1] = srt[p]; --p; }
            srt[p + 1] = v;
        }
        for (int c = 0; c < npe; ++c) {
            if (sorter_out) sorter_out[e * npe + c] = srt[c];
            if (sorted_elements_out) sorted_elements_out[e * npe + c] = elements[e * npe + srt[c]];
            for (int n = 0; n < nv; ++n) {
                rg[nv * c + n] = (int)(nv * elements[e * npe + srt[c]]) + n;
                rl[nv * c + n] = srt[c] * nv + n;
            }
        }
        if (data_local_indices)
            for (int i = 0; i < ndof; ++i) {
                const int row0 = indptr[rg[i]], nnz = indptr[rg[i] + 1] - row0;
                for (int j = 0; j < ndof; ++j) {
                    data_global_indices[cap * e + (size_t)i * ndof + j] = row0 + binary_locate(indices + row0, nnz, rg[j]);
                    data_local_indices[cap * e + (size_t)i * ndof + j] = rl[i] * ndof + rl[j];
                }
            }
    }
    free(srt); free(rg); free(rl);
    return 0;
}

/* RHSAssemblyNative_: RHSAssemblyNative.pyx:30-39 is folded into the loops above (T[conn*nvar+i] += t_e). */
