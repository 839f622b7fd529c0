// physics.cuh -- CUDA-side implementations of the reference's shipped physics.
//
// The reference's Physics interface (src/Physics.jl:1-18) is a set of Julia closures called once
// per quadrature point:  residual / stiffness / stiffness_action(physics, interps, x_el, t, dt,
// u_el, u_el_old, [v_el,] state_old_q, state_new_q, props_el)  (Assemblers.jl:427, MatrixAction.jl:72).
// Closures cannot cross the C ABI, so each shipped physics is a device functor with three hooks:
//
//   flux   (grad u)          -> P[d][j]   (+ body term b[d]):  R[a,d]  = JxW (sum_j dN_X[a,j] P[d][j] + N[a] b[d])
//   dflux  (grad u, grad v)  -> dP = A : grad v                (matrix-free stiffness_action)
//   tangent(grad u)          -> A[(d1,j1)][(d2,j2)] = dP[d1][j1]/d(grad u)[d2][j2]
//
// which is exactly the structure of scatter_with_gradients! / scatter_with_gradients_and_gradients!
// (src/Formulations.jl:27-49, 89-126).  Mechanics physics work on 3x3 tensors; in 2-D the gradient
// is padded with zeros (PlaneStrain, Formulations.jl:421-427) and the in-plane block is kept.
#pragma once
#include <cuda_runtime.h>

namespace fec {

#define FEC_DEV __device__ __forceinline__

FEC_DEV double det3(const double (&F)[3][3]) {
  return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) -
         F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
         F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
}

// H = F^{-T}  (H[i][J] = Finv[J][i]) and J = det F
FEC_DEV double inv_transpose3(const double (&F)[3][3], double (&H)[3][3]) {
  const double c00 = F[1][1] * F[2][2] - F[1][2] * F[2][1];
  const double c01 = F[1][2] * F[2][0] - F[1][0] * F[2][2];
  const double c02 = F[1][0] * F[2][1] - F[1][1] * F[2][0];
  const double J = F[0][0] * c00 + F[0][1] * c01 + F[0][2] * c02;
  const double iJ = 1.0 / J;
  // cofactor matrix C[i][j]; F^{-T} = C / J
  H[0][0] = c00 * iJ; H[0][1] = c01 * iJ; H[0][2] = c02 * iJ;
  H[1][0] = (F[0][2] * F[2][1] - F[0][1] * F[2][2]) * iJ;
  H[1][1] = (F[0][0] * F[2][2] - F[0][2] * F[2][0]) * iJ;
  H[1][2] = (F[0][1] * F[2][0] - F[0][0] * F[2][1]) * iJ;
  H[2][0] = (F[0][1] * F[1][2] - F[0][2] * F[1][1]) * iJ;
  H[2][1] = (F[0][2] * F[1][0] - F[0][0] * F[1][2]) * iJ;
  H[2][2] = (F[0][0] * F[1][1] - F[0][1] * F[1][0]) * iJ;
  return J;
}

template <int NF, int ND>
FEC_DEV void pad3(const double (&g)[NF][ND], double (&G)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G[i][j] = (i < NF && j < ND) ? g[i < NF ? i : 0][j < ND ? j : 0] : 0.0;
}

template <int NF, int ND>
FEC_DEV void restrict3(const double (&P3)[3][3], double (&P)[NF][ND]) {
#pragma unroll
  for (int i = 0; i < NF; ++i)
#pragma unroll
    for (int j = 0; j < ND; ++j) P[i][j] = P3[i][j];
}

// -------------------------------------------------------------------------------------------
// Poisson  (test/poisson/TestPoissonCommon.jl:4-139), AbstractPhysics{1,0,0}
//   residual  JxW (grad_u . dN_X^T - N f)   :75-83      stiffness JxW dN_X dN_X^T   :100-107
//   action    JxW dN_X (dN_X^T v)           :121-127    mass      JxW N N^T         :18-41
// -------------------------------------------------------------------------------------------
template <int ND>
struct PhysPoisson {
  static constexpr int NF = 1, NS = 0;
  static constexpr bool kHasSource = true;
  FEC_DEV static double density(const double*) { return 1.0; }
  FEC_DEV static void flux(const double (&gu)[1][ND], double fq, const double*, const double*, double*,
                           double (&P)[1][ND], double (&b)[1]) {
#pragma unroll
    for (int j = 0; j < ND; ++j) P[0][j] = gu[0][j];
    b[0] = -fq;
  }
  FEC_DEV static void dflux(const double (&)[1][ND], const double (&gv)[1][ND], const double*, const double*,
                            double (&dP)[1][ND]) {
#pragma unroll
    for (int j = 0; j < ND; ++j) dP[0][j] = gv[0][j];
  }
  FEC_DEV static void tangent(const double (&)[1][ND], const double*, const double*, double (&A)[ND][ND]) {
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j) A[i][j] = (i == j) ? 1.0 : 0.0;
  }
  // energy density e_q = 1/2 grad u . grad u - u_q f(X_q)   (test/poisson/TestPoissonCommon.jl:8-16)
  static constexpr bool kHasEnergy = true;
  FEC_DEV static double energy(const double (&gu)[1][ND], const double (&uq)[1], double fq, const double*, const double*) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < ND; ++j) s = fma(gu[0][j], gu[0][j], s);
    return 0.5 * s - uq[0] * fq;
  }
};

// Shared plumbing of the mechanics physics: Impl works on 3x3 tensors.
template <int ND, class Impl>
struct PhysMech3 {
  static constexpr int NF = ND, NS = Impl::NS;
  static constexpr bool kHasSource = false;
  FEC_DEV static double density(const double* props) { return props[0]; }
  FEC_DEV static void flux(const double (&gu)[ND][ND], double, const double* props, const double* so, double* sn,
                           double (&P)[ND][ND], double (&b)[ND]) {
    double G3[3][3], P3[3][3];
    pad3<ND, ND>(gu, G3);
    Impl::stress(G3, props, so, sn, P3);
    restrict3<ND, ND>(P3, P);
#pragma unroll
    for (int d = 0; d < ND; ++d) b[d] = 0.0;
  }
  FEC_DEV static void dflux(const double (&gu)[ND][ND], const double (&gv)[ND][ND], const double* props,
                            const double* so, double (&dP)[ND][ND]) {
    double G3[3][3], V3[3][3], D3[3][3];
    pad3<ND, ND>(gu, G3);
    pad3<ND, ND>(gv, V3);
    Impl::dstress(G3, V3, props, so, D3);
    restrict3<ND, ND>(D3, dP);
  }
  // strain energy density psi(grad u)  (energy(physics::Mechanics, ...), test/mechanics/TestMechanicsCommon.jl:38-50)
  static constexpr bool kHasEnergy = Impl::kHasEnergy;
  FEC_DEV static double energy(const double (&gu)[ND][ND], const double (&)[ND], double, const double* props, const double* so) {
    double G3[3][3];
    pad3<ND, ND>(gu, G3);
    if constexpr (Impl::kHasEnergy) return Impl::energy(G3, props, so);
    else return 0.0;
  }
  // A[(d1*ND+j1)][(d2*ND+j2)]
  FEC_DEV static void tangent(const double (&gu)[ND][ND], const double* props, const double* so,
                              double (&A)[ND * ND][ND * ND]) {
    double G3[3][3];
    pad3<ND, ND>(gu, G3);
    typename Impl::Pre pre;
    Impl::prepare(G3, props, so, pre);
    if constexpr (Impl::kScalesTangent) Impl::scale_tangent(pre, 1.0);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int k = 0; k < ND; ++k)
#pragma unroll
          for (int l = 0; l < ND; ++l) A[i * ND + j][k * ND + l] = Impl::A(pre, i, j, k, l);
  }
  // scale * A (the matrix kernels need JxW * A): constitutive laws whose tangent is a sum of coefficient-scaled
  // products fold the factor into the coefficients (Impl::scale_tangent), the others multiply the entries
  FEC_DEV static void tangent_scaled(const double (&gu)[ND][ND], const double* props, const double* so, const double scale,
                                     double (&A)[ND * ND][ND * ND]) {
    double G3[3][3];
    pad3<ND, ND>(gu, G3);
    typename Impl::Pre pre;
    Impl::prepare(G3, props, so, pre);
    if constexpr (Impl::kScalesTangent) Impl::scale_tangent(pre, scale);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int k = 0; k < ND; ++k)
#pragma unroll
          for (int l = 0; l < ND; ++l) {
            const double a = Impl::A(pre, i, j, k, l);
            A[i * ND + j][k * ND + l] = Impl::kScalesTangent ? a : a * scale;
          }
  }
  // The same tangent pulled back to reference gradients, Ahat[(i,k1)][(k,k2)] = sum_{J,L} Ji[k1][J] A_iJkL Ji[k2][L]
  // (Ji = inverse element Jacobian), for laws that can build it at the cost of A itself (Impl::kRefTangent): the Walsh
  // form of k_mat2 consumes it directly and skips its per-pair transform.
  static constexpr bool kRefTangent = (ND == 3) && Impl::kRefTangent;
  FEC_DEV static void tangent_scaled_ref(const double (&gu)[ND][ND], const double (&Ji)[ND][ND], const double* props,
                                         const double* so, const double scale, double (&A)[ND * ND][ND * ND]) {
    if constexpr (kRefTangent) {
      typename Impl::Pre pre;
      Impl::prepare(gu, props, so, pre);
      Impl::to_reference(pre, Ji);
      if constexpr (Impl::kScalesTangent) Impl::scale_tangent(pre, scale);
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int k = 0; k < ND; ++k)
#pragma unroll
            for (int l = 0; l < ND; ++l) {
              const double a = Impl::A_ref(pre, i, j, k, l);
              A[i * ND + j][k * ND + l] = Impl::kScalesTangent ? a : a * scale;
            }
    }
  }
};

// -------------------------------------------------------------------------------------------
// Linear elasticity (test/mechanics/TestMechanicsCommon.jl:14-20):
//   psi = 1/2 K tr(eps)^2 + G dev(eps):dev(eps),  eps = sym(grad u), props = (rho, K, G)
// -------------------------------------------------------------------------------------------
struct LinearElasticImpl {
  static constexpr int NS = 0;
  static constexpr bool kScalesTangent = true;    // A is linear in (K, G): a scale factor goes into the two moduli
  static constexpr bool kHasEnergy = true;
  static constexpr bool kRefTangent = true;
  struct Pre { double K, G, Ji[3][3], C[3][3]; };
  // psi = 1/2 K tr(eps)^2 + G dev(eps):dev(eps)   (TestMechanicsCommon.jl:14-20)
  FEC_DEV static double energy(const double (&g)[3][3], const double* props, const double*) {
    const double K = props[1], G = props[2];
    const double tr = g[0][0] + g[1][1] + g[2][2];
    double ee = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { const double e = 0.5 * (g[i][j] + g[j][i]); ee = fma(e, e, ee); }
    return 0.5 * K * tr * tr + G * (ee - tr * tr * (1.0 / 3.0));
  }
  FEC_DEV static void stress(const double (&g)[3][3], const double* props, const double*, double*, double (&P)[3][3]) {
    const double K = props[1], G = props[2];
    const double tr = g[0][0] + g[1][1] + g[2][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double e = 0.5 * (g[i][j] + g[j][i]);
        P[i][j] = 2.0 * G * e + ((i == j) ? (K - 2.0 * G / 3.0) * tr : 0.0);
      }
  }
  FEC_DEV static void dstress(const double (&)[3][3], const double (&v)[3][3], const double* props, const double*,
                              double (&D)[3][3]) {
    stress(v, props, nullptr, nullptr, D);
  }
  FEC_DEV static void prepare(const double (&)[3][3], const double* props, const double*, Pre& p) {
    p.K = props[1]; p.G = props[2];
  }
  FEC_DEV static void scale_tangent(Pre& p, const double s) { p.K *= s; p.G *= s; }
  FEC_DEV static double A(const Pre& p, int i, int j, int k, int l) {
    const double dij = (i == j), dkl = (k == l), dik = (i == k), djl = (j == l), dil = (i == l), djk = (j == k);
    return (p.K - 2.0 * p.G / 3.0) * dij * dkl + p.G * (dik * djl + dil * djk);
  }
  // pulled back: lam Ji[k1][i] Ji[k2][k] + G (d_ik (Ji Ji^T)[k1][k2] + Ji[k2][i] Ji[k1][k])
  FEC_DEV static void to_reference(Pre& p, const double (&Ji)[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        p.Ji[a][b] = Ji[a][b];
        double s = Ji[a][0] * Ji[b][0];
        s = fma(Ji[a][1], Ji[b][1], s);
        p.C[a][b] = fma(Ji[a][2], Ji[b][2], s);
      }
  }
  FEC_DEV static double A_ref(const Pre& p, int i, int k1, int k, int k2) {
    double a = (p.K - 2.0 * p.G / 3.0) * p.Ji[k1][i] * p.Ji[k2][k];
    a = fma(p.G * p.Ji[k2][i], p.Ji[k1][k], a);
    return (i == k) ? fma(p.G, p.C[k1][k2], a) : a;
  }
};

// -------------------------------------------------------------------------------------------
// Compressible neo-Hookean (test/mechanics/TestMechanicsLargeDeformation.jl:17-27; stale script,
// parity unpinned): psi = 1/2 K U(J) + 1/2 G (J^-2/3 tr(F F^T) - 3),  F = I + grad u.
//   AS_WRITTEN = false: U = 1/2 (J^2 - 1) - ln J   (stress free at F = I)
//   AS_WRITTEN = true : U = 1/2 (J - 1)^2 - ln J   (the script verbatim)
// P = c H + G m (F - I1/3 H),   H = F^-T, m = J^-2/3, c = 1/2 K (J^2-1) | 1/2 K (J^2-J-1)
// A_iJkL = c' J H_iJ H_kL - c H_iL H_kJ
//        + G [ -(2/3) m (F - I1/3 H)_iJ H_kL + m (d_ik d_JL - (2/3) H_iJ F_kL + (I1/3) H_iL H_kJ) ]
// (analytic form of Tensors.gradient / Tensors.hessian of psi).
// -------------------------------------------------------------------------------------------
template <bool AS_WRITTEN>
struct NeoHookeanImpl {
  static constexpr int NS = 0;
  static constexpr bool kHasEnergy = true;
  // Hs, Hg, Hb: the three coefficient-scaled copies of H / F the tangent is built from (see A below)
  static constexpr bool kRefTangent = true;
  struct Pre { double F[3][3], H[3][3], Hs[3][3], Hg[3][3], Hb[3][3], C[3][3], J, I1, m, c, cpJ, G, gm; };
  // psi = 1/2 K U(J) + 1/2 G (J^-2/3 tr(F F^T) - 3)   (TestMechanicsLargeDeformation.jl:17-27; U: see prepare)
  FEC_DEV static double energy(const double (&g)[3][3], const double* props, const double*) {
    Pre p;
    prepare(g, props, nullptr, p);
    const double K = props[1];
    const double U = AS_WRITTEN ? 0.5 * (p.J - 1.0) * (p.J - 1.0) - log(p.J) : 0.5 * (p.J * p.J - 1.0) - log(p.J);
    return 0.5 * K * U + 0.5 * p.G * (p.m * p.I1 - 3.0);
  }
  FEC_DEV static void prepare(const double (&g)[3][3], const double* props, const double*, Pre& p) {
    const double K = props[1];
    p.G = props[2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) p.F[i][j] = g[i][j] + ((i == j) ? 1.0 : 0.0);
    p.J = inv_transpose3(p.F, p.H);
    p.I1 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) p.I1 = fma(p.F[i][j], p.F[i][j], p.I1);
    p.m = rcbrt(p.J * p.J);   // J^(-2/3): one routine instead of cbrt + a division on the dependent chain
    if (AS_WRITTEN) {
      p.c = 0.5 * K * (p.J * p.J - p.J - 1.0);
      p.cpJ = 0.5 * K * (2.0 * p.J - 1.0) * p.J;
    } else {
      p.c = 0.5 * K * (p.J * p.J - 1.0);
      p.cpJ = K * p.J * p.J;
    }
  }
  FEC_DEV static void stress(const double (&g)[3][3], const double* props, const double*, double*, double (&P)[3][3]) {
    Pre p;
    prepare(g, props, nullptr, p);
    const double gm = p.G * p.m, a = p.c - gm * p.I1 * (1.0 / 3.0);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) P[i][j] = a * p.H[i][j] + gm * p.F[i][j];
  }
  FEC_DEV static void dstress(const double (&g)[3][3], const double (&dF)[3][3], const double* props, const double*,
                              double (&D)[3][3]) {
    Pre p;
    prepare(g, props, nullptr, p);
    double HdF = 0.0, FdF = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) { HdF = fma(p.H[i][j], dF[i][j], HdF); FdF = fma(p.F[i][j], dF[i][j], FdF); }
    // T = H dF^T H
    double W[3][3], T[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s = fma(dF[k][i], p.H[k][j], s);  // (dF^T H)[i][j]
        W[i][j] = s;
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s = fma(p.H[i][k], W[k][j], s);
        T[i][j] = s;
      }
    const double gm = p.G * p.m, third = 1.0 / 3.0;
    // coefficients: D = aH*H + aF*F + aT*T + gm*dF
    const double aH = p.cpJ * HdF + gm * ((2.0 * third) * HdF * p.I1 * third - (2.0 * third) * FdF);
    const double aF = -(2.0 * third) * gm * HdF;
    const double aT = -p.c + gm * p.I1 * third;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        D[i][j] = aH * p.H[i][j] + aF * p.F[i][j] + aT * T[i][j] + gm * dF[i][j];
  }
  static constexpr bool kScalesTangent = true;
  // A_iJkL = Hs_iJ H_kL + Hg_iJ F_kL + Hb_iL H_kJ + gm d_ik d_JL   with
  //   Hs = (c'J + 2/9 gm I1) H - 2/3 gm F,  Hg = -2/3 gm H,  Hb = (gm I1/3 - c) H
  // (the formula in the header comment with the common factors collected: 3 FP64 instructions per entry).  `s` scales the
  // whole tangent (the matrix kernels need JxW * A): it is folded into the four coefficients.
  FEC_DEV static void scale_tangent(Pre& p, const double s) {
    const double third = 1.0 / 3.0;
    const double gm = p.G * p.m;
    p.gm = gm * s;
    const double as = (p.cpJ + (2.0 * third * third) * gm * p.I1) * s, ag = -(2.0 * third) * p.gm, ab = (gm * p.I1 * third - p.c) * s;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        p.Hs[i][j] = fma(as, p.H[i][j], ag * p.F[i][j]);
        p.Hg[i][j] = ag * p.H[i][j];
        p.Hb[i][j] = ab * p.H[i][j];
      }
  }
  FEC_DEV static double A(const Pre& p, int i, int j, int k, int l) {
    double a = p.Hs[i][j] * p.H[k][l];
    a = fma(p.Hg[i][j], p.F[k][l], a);
    a = fma(p.Hb[i][l], p.H[k][j], a);
    return (i == k && j == l) ? a + p.gm : a;
  }
  // pulled back to reference gradients: every factor X_iJ of A becomes (X Ji^T)[i][k]; d_JL becomes (Ji Ji^T)[k1][k2].
  // Call between prepare and scale_tangent (Hs, Hg, Hb are linear in H and F).
  FEC_DEV static void to_reference(Pre& p, const double (&Ji)[3][3]) {
    double Ht[3][3], Ft[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double h = p.H[i][0] * Ji[k][0], f = p.F[i][0] * Ji[k][0];
#pragma unroll
        for (int j = 1; j < 3; ++j) { h = fma(p.H[i][j], Ji[k][j], h); f = fma(p.F[i][j], Ji[k][j], f); }
        Ht[i][k] = h; Ft[i][k] = f;
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        p.H[i][k] = Ht[i][k]; p.F[i][k] = Ft[i][k];
        double s = Ji[i][0] * Ji[k][0];
        s = fma(Ji[i][1], Ji[k][1], s);
        p.C[i][k] = fma(Ji[i][2], Ji[k][2], s);
      }
  }
  FEC_DEV static double A_ref(const Pre& p, int i, int k1, int k, int k2) {
    double a = p.Hs[i][k1] * p.H[k][k2];
    a = fma(p.Hg[i][k1], p.F[k][k2], a);
    a = fma(p.Hb[i][k2], p.H[k][k1], a);
    return (i == k) ? fma(p.gm, p.C[k1][k2], a) : a;
  }
};

// -------------------------------------------------------------------------------------------
// Small-strain J2 plasticity, linear isotropic hardening, radial return.  The reference ships only
// the hooks (AbstractPhysics{.,.,7}, state_old/state_new views: Assemblers.jl:248-251,
// test/mechanics_with_state/TestMechanicsWithState.jl:15-67); the law itself is defined by the
// oracle (parity unpinned).  state = [ep_xx, ep_yy, ep_zz, ep_yz, ep_xz, ep_xy, eqps],
// props = (rho, K, G, sigma_y, H).
// -------------------------------------------------------------------------------------------
struct J2Impl {
  static constexpr int NS = 7;
  static constexpr bool kScalesTangent = false;
  static constexpr bool kRefTangent = false;  // general A: the Walsh kernel pulls the pair blocks back itself
  static constexpr bool kHasEnergy = false;  // no energy is defined for the (oracle-defined) J2 law
  struct Pre { double K, G, theta, thbar, n[3][3]; };
  struct RM { double tr, s[3][3], n[3][3], dg, q; bool yld; };
  FEC_DEV static void return_map(const double (&g)[3][3], const double* props, const double* so, RM& r) {
    const double G = props[2], sy = props[3], Hh = props[4];
    r.tr = g[0][0] + g[1][1] + g[2][2];
    double ep[3][3];
    ep[0][0] = so[0]; ep[1][1] = so[1]; ep[2][2] = so[2];
    ep[1][2] = ep[2][1] = so[3]; ep[0][2] = ep[2][0] = so[4]; ep[0][1] = ep[1][0] = so[5];
    double nrm2 = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double e = 0.5 * (g[i][j] + g[j][i]) - ((i == j) ? r.tr * (1.0 / 3.0) : 0.0) - ep[i][j];
        r.s[i][j] = 2.0 * G * e;
        nrm2 = fma(r.s[i][j], r.s[i][j], nrm2);
      }
    const double nrm = sqrt(nrm2);
    r.q = sqrt(1.5) * nrm;
    const double f = r.q - (sy + Hh * so[6]);
    r.yld = f > 0.0;
    r.dg = r.yld ? f / (3.0 * G + Hh) : 0.0;
    const double inv = nrm > 0.0 ? 1.0 / nrm : 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) r.n[i][j] = r.s[i][j] * inv;
  }
  FEC_DEV static void stress(const double (&g)[3][3], const double* props, const double* so, double* sn,
                             double (&P)[3][3]) {
    RM r;
    return_map(g, props, so, r);
    const double K = props[1], G = props[2];
    const double fac = 2.0 * G * sqrt(1.5) * r.dg;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) P[i][j] = r.s[i][j] - fac * r.n[i][j] + ((i == j) ? K * r.tr : 0.0);
    if (sn) {
      const double de = sqrt(1.5) * r.dg;
      sn[0] = so[0] + de * r.n[0][0]; sn[1] = so[1] + de * r.n[1][1]; sn[2] = so[2] + de * r.n[2][2];
      sn[3] = so[3] + de * r.n[1][2]; sn[4] = so[4] + de * r.n[0][2]; sn[5] = so[5] + de * r.n[0][1];
      sn[6] = so[6] + r.dg;
    }
  }
  FEC_DEV static void prepare(const double (&g)[3][3], const double* props, const double* so, Pre& p) {
    RM r;
    return_map(g, props, so, r);
    p.K = props[1]; p.G = props[2];
    const double Hh = props[4];
    const double qs = r.q > 0.0 ? r.q : 1.0;
    p.theta = r.yld ? 1.0 - 3.0 * p.G * r.dg / qs : 1.0;
    p.thbar = r.yld ? 1.0 / (1.0 + Hh / (3.0 * p.G)) - (1.0 - p.theta) : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) p.n[i][j] = r.n[i][j];
  }
  FEC_DEV static void dstress(const double (&g)[3][3], const double (&v)[3][3], const double* props, const double* so,
                              double (&D)[3][3]) {
    Pre p;
    prepare(g, props, so, p);
    const double tr = v[0][0] + v[1][1] + v[2][2];
    double nde = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) nde = fma(p.n[i][j], 0.5 * (v[i][j] + v[j][i]), nde);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double e = 0.5 * (v[i][j] + v[j][i]);
        D[i][j] = 2.0 * p.G * p.theta * (e - ((i == j) ? tr * (1.0 / 3.0) : 0.0)) + ((i == j) ? p.K * tr : 0.0) -
                  2.0 * p.G * p.thbar * nde * p.n[i][j];
      }
  }
  FEC_DEV static double A(const Pre& p, int i, int j, int k, int l) {
    const double dij = (i == j), dkl = (k == l), dik = (i == k), djl = (j == l), dil = (i == l), djk = (j == k);
    return p.K * dij * dkl + 2.0 * p.G * p.theta * (0.5 * (dik * djl + dil * djk) - dij * dkl * (1.0 / 3.0)) -
           2.0 * p.G * p.thbar * p.n[i][j] * p.n[k][l];
  }
};

// -------------------------------------------------------------------------------------------
// TEST physics with a NON-symmetric tangent (A_ijkl != A_klij): linear elasticity plus beta * delta_ij T_kl with a
// fixed non-symmetric T.  Not a material model: it exists so that the transposed COO labelling of the reference's
// pattern (Assemblers.jl:109-124 vs SparsityPatterns.jl:76-83) is observable in a parity test -- every
// shipped law has a symmetric tangent, for which the convention is invisible.  props = (rho, K, G, beta).
// -------------------------------------------------------------------------------------------
struct NonSymmetricTestImpl {
  static constexpr int NS = 0;
  static constexpr bool kRefTangent = false;
  static constexpr bool kScalesTangent = true;
  static constexpr bool kHasEnergy = false;
  struct Pre { double K, G, beta; };
  FEC_DEV static double T(int k, int l) {
    return (k == 0 && l == 1) ? 1.0 : (k == 1 && l == 2) ? 2.0 : (k == 2 && l == 0) ? 3.0 : (k == 1 && l == 0) ? -0.5 : (k == l ? 0.25 * (k + 1) : 0.0);
  }
  FEC_DEV static void stress(const double (&g)[3][3], const double* props, const double*, double*, double (&P)[3][3]) {
    LinearElasticImpl::stress(g, props, nullptr, nullptr, P);
    double tg = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) tg = fma(T(k, l), g[k][l], tg);
#pragma unroll
    for (int i = 0; i < 3; ++i) P[i][i] = fma(props[3], tg, P[i][i]);
  }
  FEC_DEV static void dstress(const double (&)[3][3], const double (&v)[3][3], const double* props, const double*,
                              double (&D)[3][3]) {
    stress(v, props, nullptr, nullptr, D);
  }
  FEC_DEV static void prepare(const double (&)[3][3], const double* props, const double*, Pre& p) {
    p.K = props[1]; p.G = props[2]; p.beta = props[3];
  }
  FEC_DEV static void scale_tangent(Pre& p, const double s) { p.K *= s; p.G *= s; p.beta *= s; }
  FEC_DEV static double A(const Pre& p, int i, int j, int k, int l) {
    const double dij = (i == j), dkl = (k == l), dik = (i == k), djl = (j == l), dil = (i == l), djk = (j == k);
    return (p.K - 2.0 * p.G / 3.0) * dij * dkl + p.G * (dik * djl + dil * djk) + p.beta * dij * T(k, l);
  }
  FEC_DEV static double energy(const double (&)[3][3], const double*, const double*) { return 0.0; }
};

template <int ND> using PhysLinearElastic = PhysMech3<ND, LinearElasticImpl>;
template <int ND> using PhysNonSymmetricTest = PhysMech3<ND, NonSymmetricTestImpl>;
template <int ND> using PhysNeoHookean = PhysMech3<ND, NeoHookeanImpl<false>>;
template <int ND> using PhysNeoHookeanAsWritten = PhysMech3<ND, NeoHookeanImpl<true>>;
template <int ND> using PhysJ2 = PhysMech3<ND, J2Impl>;

}  // namespace fec
