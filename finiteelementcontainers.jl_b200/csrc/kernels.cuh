// kernels.cuh -- element kernels of the assembly hot path (sm_100a, FP64).
//
// Replaces the reference's one-thread-per-element KernelAbstractions loops
//   _assemble_block!                      src/assemblers/Assemblers.jl:402-464
//   _assemble_block_matrix_free_action!   src/assemblers/MatrixAction.jl:49-77
//   _assemble_block_matrix_action!        src/assemblers/MatrixAction.jl:208-270
// with two hand-written kernel families:
//
//  k_vec  (residual / matrix-free action): one CTA per locality tile of TE elements.
//     1. the tile's unique nodes are gathered ONCE into shared memory (X, U, V) -- each node is
//        read once per tile instead of once per incident element,
//     2. one thread per element runs the quadrature loop in registers; reference tables
//        (N, dN/dxi, w) live in the kernel-parameter constant bank and feed DFMA as uniform operands,
//     3. element vectors are staged in shared memory and reduced per NODE in a fixed order by the
//        tile's node-owner threads (deterministic segmented reduction, no shared-memory atomics:
//        FP64 shared atomics are CAS loops on sm_100a),
//     4. one red.global.add.f64 per (node, dof) per tile -- only tile-boundary nodes ever see
//        more than one.
//
//  k_mat  (stiffness / mass): NNPE threads per element.  Thread q computes the geometry and the
//     material tangent of quadrature point q once and publishes it in shared memory; thread b then
//     owns block-column b of K_el in registers and writes it straight into the CSR/CSC values
//     through the precomputed element -> CSR slot map (node-pair position bytes + per-row offsets).
//     No COO is ever materialised (the reference writes 576 doubles/element of COO and runs
//     SparseArrays.sparse! every Newton iteration, SparsityPatterns.jl:301-308).
#pragma once
#include "common.cuh"
#include "physics.cuh"
#include <memory>
#include <type_traits>
#include <algorithm>
#include <cmath>
#include <cstdlib>

// experiment switch: the Walsh residual kernel with its X / U coefficients in the shared-memory stash and 3 CTAs per SM
// (154 registers, no spills): 1.88 ms against 1.84 ms with everything in registers at 2 CTAs per SM -- off
#ifndef FEC_VEC_STASH_R
#define FEC_VEC_STASH_R 0
#endif

namespace fec {

template <int ND, int NNPE, int NQT>
struct Tables {
  static constexpr int NQ = (NQT > 0) ? NQT : kMaxNQ;
  double N[NQ][NNPE];
  double dN[NQ][NNPE][ND];
  double w[NQ];
};

template <int ND, int NNPE, int NQT>
struct VecParams {
  const double* X;
  const double* U;
  const double* V;
  double* out;
  const int32_t* tile_node_ptr;
  const int32_t* tile_nodes;
  const uint16_t* lconn;
  const int32_t* inc_ptr;
  const uint16_t* inc;
  const double* state_old;
  double* state_new;
  const double* source;
  PeerScatter peer;
  int32_t ne, nq;
  int32_t body_doubles, max_nodes;  // shared-memory layout: [node data | element-vector stage][node ids][inc_ptr][inc]
  double wr[4];                     // Walsh path: c^n / 8, n = 0..3 (c = |xi| of the 2-point rule per axis)
  int32_t nos[8], pos[8];           //             sign index -> local node / quadrature point (identity otherwise)
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};

template <int ND>
FEC_DEV double invert(const double (&J)[ND][ND], double (&Ji)[ND][ND]) {
  if constexpr (ND == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    Ji[0][0] = J[1][1] * id; Ji[0][1] = -J[0][1] * id;
    Ji[1][0] = -J[1][0] * id; Ji[1][1] = J[0][0] * id;
    return det;
  } else {
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    Ji[0][0] = c00 * id; Ji[1][0] = c01 * id; Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    return det;
  }
}

// One quadrature point of the vector kernels.  MappedH1OrL2Interpolants (src/Physics.jl:86-92):
//   J[i][j] = sum_a x[a][i] dN[a][j],  dN_X = dN J^-1,  JxW = det J * w.
// dN_X is never formed: grad u = (sum_a u_a (x) dN_a) J^-1 and the scatter uses P J^-T, which is the
// same arithmetic with 2*NNPE*ND*ND fewer FMAs per point.
// where the element fields of the quadrature loop live: registers, or the thread's column of a shared-memory stash
// (entry (a, c) at base[(a * C + c) * STRIDE], STRIDE = threads per CTA: conflict-free).  The stash is for elements whose
// fields do not fit next to the constitutive temporaries (TET10 mechanics: 3 x 30 doubles): see vec_qstash below.
template <int N, int C>
struct FieldReg {
  const double (&f)[N][C];
  FEC_DEV double operator()(int a, int c) const { return f[a][c]; }
};
template <int C, int STRIDE>
struct FieldStash {
  const double* base;
  FEC_DEV double operator()(int a, int c) const { return base[(a * C + c) * STRIDE]; }
};
#ifndef FEC_VEC_QSTASH
#define FEC_VEC_QSTASH 1
#endif
#ifndef FEC_VEC_QSTASH_MINB
#define FEC_VEC_QSTASH_MINB 1   // extra CTAs per SM of the stashed kernels (register cap 168 instead of 255)
#endif
template <int NNPE, int NF, bool WALSH>
__host__ __device__ constexpr bool vec_qstash() { return FEC_VEC_QSTASH && !WALSH && NNPE >= 10 && NF == 3; }

template <int ND, int NNPE, int NF, class Phys, int MODE, class Tab, class AX, class AU, class AV>
FEC_DEV void vec_qp(const Tab& tab, const int q, const AX& x, const AU& u,
                    const AV& v, const double* props, const double fq, const double* so, double* sn,
                    double (&r)[NNPE][NF]) {
  double J[ND][ND];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < NNPE; ++a) s = fma(x(a, i), tab.dN[q][a][j], s);
      J[i][j] = s;
    }
  double Ji[ND][ND];
  const double JxW = invert<ND>(J, Ji) * tab.w[q];

  if constexpr (MODE == MODE_LUMPED_MASS || MODE == MODE_DIAG_MASS) {
    // lumped_mass: rho JxW N[a] in every direction (row sum of the consistent element mass, partition of unity;
    // LumpedMass.jl:1-30, TestMechanicsCommon.jl:98-125).  Diagonal of the consistent mass: rho JxW N[a]^2
    // (assemble_diagonal!(asm, mass, ...), Diagonal.jl:1-14 + Assemblers.jl:42-45).
    const double rho = Phys::density(props) * JxW;
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
      const double m = rho * tab.N[q][a] * (MODE == MODE_DIAG_MASS ? tab.N[q][a] : 1.0);
#pragma unroll
      for (int d = 0; d < NF; ++d) r[a][d] += m;
    }
    return;
  } else if constexpr (MODE == MODE_ACTION_MASS) {
    // mass_action: JxW rho N (N . v)   (TestPoissonCommon.jl:66-72, Formulations.jl:264-288)
    const double rho = Phys::density(props) * JxW;
#pragma unroll
    for (int d = 0; d < NF; ++d) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < NNPE; ++a) s = fma(tab.N[q][a], v(a, d), s);
      s *= rho;
#pragma unroll
      for (int a = 0; a < NNPE; ++a) r[a][d] = fma(tab.N[q][a], s, r[a][d]);
    }
    return;
  } else {
    // grad u in reference coordinates, then push to physical
    double gx[NF][ND], gu[NF][ND];
#pragma unroll
    for (int d = 0; d < NF; ++d)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; ++a) s = fma(u(a, d), tab.dN[q][a][j], s);
        gx[d][j] = s;
      }
#pragma unroll
    for (int d = 0; d < NF; ++d)
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j) s = fma(gx[d][j], Ji[j][k], s);
        gu[d][k] = s;
      }
    if constexpr (MODE == MODE_DIAG_STIFFNESS) {
      // diagonal of the element stiffness (Assemblers.jl:42-45 on K_q = G^T A G):
      //   K_el[(a,d),(a,d)] += JxW sum_{j1,j2} dN_X[a][j1] A[(d,j1)][(d,j2)] dN_X[a][j2]
      double A[NF * ND][NF * ND];
      Phys::tangent(gu, props, so, A);
#pragma unroll
      for (int a = 0; a < NNPE; ++a) {
        double g[ND];
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < ND; ++j) s = fma(tab.dN[q][a][j], Ji[j][k], s);
          g[k] = s;
        }
#pragma unroll
        for (int d = 0; d < NF; ++d) {
          double s = 0.0;
#pragma unroll
          for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
            for (int j2 = 0; j2 < ND; ++j2) s = fma(g[j1] * A[d * ND + j1][d * ND + j2], g[j2], s);
          r[a][d] = fma(s, JxW, r[a][d]);
        }
      }
      return;
    }
    double P[NF][ND], b[NF];
    if constexpr (MODE == MODE_RESIDUAL) {
      Phys::flux(gu, fq, props, so, sn, P, b);
    } else {
      double gvx[NF][ND], gv[NF][ND];
#pragma unroll
      for (int d = 0; d < NF; ++d)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < NNPE; ++a) s = fma(v(a, d), tab.dN[q][a][j], s);
          gvx[d][j] = s;
        }
#pragma unroll
      for (int d = 0; d < NF; ++d)
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < ND; ++j) s = fma(gvx[d][j], Ji[j][k], s);
          gv[d][k] = s;
        }
      Phys::dflux(gu, gv, props, so, P);
#pragma unroll
      for (int d = 0; d < NF; ++d) b[d] = 0.0;
    }
    // Pxi[d][j] = JxW sum_k P[d][k] Ji[j][k]
    double Px[NF][ND];
#pragma unroll
    for (int d = 0; d < NF; ++d)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < ND; ++k) s = fma(P[d][k], Ji[j][k], s);
        Px[d][j] = s * JxW;
      }
#pragma unroll
    for (int a = 0; a < NNPE; ++a)
#pragma unroll
      for (int d = 0; d < NF; ++d) {
        double s = r[a][d];
#pragma unroll
        for (int j = 0; j < ND; ++j) s = fma(tab.dN[q][a][j], Px[d][j], s);
        if constexpr (MODE == MODE_RESIDUAL && Phys::kHasSource) s = fma(tab.N[q][a] * JxW, b[d], s);
        r[a][d] = s;
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Walsh form of the HEX8 / 2x2x2 vector kernels (same identities as kernel_mat2.cuh).  With nodes and
// points labelled by sign triples, dN_a/dxi_k (q) = 1/8 sum_{S subset of the other axes} c^|S| s_a^({k} u S) sigma_q^S, so
//   J_q[i][k]     = sum_S sigma_q^S Xh[{k} u S][i],     Xh[alpha][i] = c^(|alpha|-1)/8 sum_a s_a^alpha x_a[i]   (once per element)
//   (grad_xi u)_q = the same with Uh,
//   r[a][d]       = sum_alpha s_a^alpha rh[alpha][d],   rh[alpha][d] = c^(|alpha|-1)/8 sum_q sum_{k in alpha} sigma_q^(alpha\k) Px_q[d][k]:
// 3 add/sub per entry and point instead of 8 FMA for J and grad u, 36 add/sub instead of 72 FMA per point for the
// scatter.  Sign index i: bit k set <=> +1 on axis k.  The kernel gathers the element fields in SIGN order (p.nos) and walks
// the points in sign order (p.pos), so any node / point numbering of the table works (detect_walsh on the host).
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int vw_popc3(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }
__host__ __device__ constexpr bool vw_sign(int q, int S) { return (vw_popc3((~q) & S) & 1) != 0; }   // sigma_q^S == -1 ?

// f[i][c] (sign order) -> pre-scaled monomial coefficients fh[alpha][c], alpha = 1..7 (alpha = 0 is never needed)
template <int NC>
FEC_DEV void vw_analyse(const double (&f)[8][NC], const double* wr, double (&fh)[8][NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f[i][c];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (!(i & (1 << k))) {
          const double lo = v[i], hi = v[i | (1 << k)];
          v[i] = hi + lo;
          v[i | (1 << k)] = hi - lo;
        }
#pragma unroll
    for (int al = 1; al < 8; ++al) fh[al][c] = v[al] * wr[vw_popc3(al) - 1];
  }
}
// where the coefficients of a field live: registers, or the thread's column of a shared-memory stash
// (entry (alpha, c) at base[((alpha - 1) * NC + c) * STRIDE], STRIDE = threads per CTA: conflict-free)
template <int NC>
struct VwReg {
  const double (&fh)[8][NC];
  FEC_DEV double operator()(int al, int c) const { return fh[al][c]; }
};
template <int NC, int STRIDE>
struct VwStash {
  const double* base;
  FEC_DEV double operator()(int al, int c) const { return base[((al - 1) * NC + c) * STRIDE]; }
};
template <int NC, int STRIDE>
FEC_DEV void vw_stash(const double (&fh)[8][NC], double* base) {
#pragma unroll
  for (int al = 1; al < 8; ++al)
#pragma unroll
    for (int c = 0; c < NC; ++c) base[((al - 1) * NC + c) * STRIDE] = fh[al][c];
}
// g[c][k] = sum_a f[a][c] dN[q][a][k] from the coefficients: 4 signed terms
template <int NC, int Q, class Acc>
FEC_DEV void vw_gradient(const Acc& fh, double (&g)[NC][3]) {
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double s = fh(1 << k, c);
#pragma unroll
      for (int S = 1; S < 8; ++S)
        if (!(S & (1 << k))) s = vw_sign(Q, S) ? s - fh(S | (1 << k), c) : s + fh(S | (1 << k), c);
      g[c][k] = s;
    }
}

template <int NF, class Phys, int MODE, int Q, class AccX, class AccU, class AccV>
FEC_DEV void vec_qp_walsh(const double wq, const AccX& Xh, const AccU& Uh, const AccV& Vh,
                          const double* props, const double fq, const double* so, double* sn, double (&rh)[8][NF],
                          double (&sh)[8][NF]) {
  constexpr int ND = 3;
  double Jt[ND][ND], J[ND][ND], Ji[ND][ND];
  vw_gradient<ND, Q>(Xh, Jt);   // Jt[i][k] = dx_i / dxi_k
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int k = 0; k < ND; ++k) J[i][k] = Jt[i][k];
  const double JxW = invert<ND>(J, Ji) * wq;
  double gx[NF][ND], gu[NF][ND];
  vw_gradient<NF, Q>(Uh, gx);
#pragma unroll
  for (int d = 0; d < NF; ++d)
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < ND; ++j) s = fma(gx[d][j], Ji[j][k], s);
      gu[d][k] = s;
    }
  double P[NF][ND], b[NF];
  if constexpr (MODE == MODE_RESIDUAL) {
    Phys::flux(gu, fq, props, so, sn, P, b);
  } else {
    double gvx[NF][ND], gv[NF][ND];
    vw_gradient<NF, Q>(Vh, gvx);
#pragma unroll
    for (int d = 0; d < NF; ++d)
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j) s = fma(gvx[d][j], Ji[j][k], s);
        gv[d][k] = s;
      }
    Phys::dflux(gu, gv, props, so, P);
  }
#pragma unroll
  for (int d = 0; d < NF; ++d) {
    double Px[ND];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < ND; ++k) s = fma(P[d][k], Ji[j][k], s);
      Px[j] = s * JxW;
    }
#pragma unroll
    for (int al = 1; al < 8; ++al)
#pragma unroll
      for (int k = 0; k < ND; ++k)
        if (al & (1 << k)) rh[al][d] = vw_sign(Q, al & ~(1 << k)) ? rh[al][d] - Px[k] : rh[al][d] + Px[k];
    if constexpr (MODE == MODE_RESIDUAL && Phys::kHasSource) {   // N[q][a] JxW b[d]: N_a = 1/8 sum_S c^|S| s_a^S sigma_q^S
      const double sb = JxW * b[d];
#pragma unroll
      for (int al = 0; al < 8; ++al) sh[al][d] = vw_sign(Q, al) ? sh[al][d] - sb : sh[al][d] + sb;
    }
  }
}

template <int NF, class Phys, int MODE, int Q, class Tab, class Params, class AccX, class AccU, class AccV>
FEC_DEV void vec_walsh_points(const Tab& tab, const Params& p, const int e, const AccX& Xh, const AccU& Uh,
                              const AccV& Vh, double (&rh)[8][NF], double (&sh)[8][NF]) {
  if constexpr (Q < 8) {
    if constexpr (!std::is_same<AccX, VwReg<3>>::value)
      asm volatile("" ::: "memory");   // keep the stash loads of later points from being hoisted (register pressure)
    constexpr int NS = Phys::NS;
    double so[NS > 0 ? NS : 1], sn[NS > 0 ? NS : 1];
    if constexpr (NS > 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + p.pos[Q]) * p.ne + e];
    }
    double fq = 0.0;
    if constexpr (Phys::kHasSource && MODE == MODE_RESIDUAL) {
      if (p.source) fq = p.source[(size_t)p.pos[Q] * p.ne + e];
    }
    vec_qp_walsh<NF, Phys, MODE, Q>(tab.w[p.pos[Q]], Xh, Uh, Vh, p.props, fq, so, (NS > 0 && MODE == MODE_RESIDUAL) ? sn : nullptr, rh, sh);
    if constexpr (NS > 0 && MODE == MODE_RESIDUAL) {
#pragma unroll
      for (int s = 0; s < NS; ++s) p.state_new[((size_t)s * p.nq + p.pos[Q]) * p.ne + e] = sn[s];
    }
    vec_walsh_points<NF, Phys, MODE, Q + 1>(tab, p, e, Xh, Uh, Vh, rh, sh);
  }
}

// the whole element, fields and result in SIGN order, for MODE_RESIDUAL / MODE_ACTION_STIFFNESS.  STASH: the
// coefficients of X and V wait in the thread's column of shared memory (the action kernel of NF = 3 holds four 21-entry
// fields plus the constitutive temporaries: 432 B of spills otherwise); `stash` = 2 * 21 * TE doubles, this thread's column.
template <int NF, class Phys, int MODE, bool STASH, int TE, class Params>
FEC_DEV void vec_element_walsh(const Params& p, const int e, const double (&x)[8][3], const double (&u)[8][NF],
                               const double (&v)[8][NF], double (&r)[8][NF], double* stash) {
  double rh[8][NF], sh[8][NF];
#pragma unroll
  for (int al = 0; al < 8; ++al)
#pragma unroll
    for (int d = 0; d < NF; ++d) { rh[al][d] = 0.0; sh[al][d] = 0.0; }
  if constexpr (STASH) {
    {
      double Xh[8][3];
      vw_analyse<3>(x, p.wr, Xh);
      vw_stash<3, TE>(Xh, stash);
    }
    double* st2 = stash + 7 * 3 * TE;
    if constexpr (MODE == MODE_ACTION_STIFFNESS) {   // X and V in the stash, U in registers
      double Uh[8][NF];
      vw_analyse<NF>(u, p.wr, Uh);
      {
        double Vh[8][NF];
        vw_analyse<NF>(v, p.wr, Vh);
        vw_stash<NF, TE>(Vh, st2);
      }
      vec_walsh_points<NF, Phys, MODE, 0>(p.tab, p, e, VwStash<3, TE>{stash}, VwReg<NF>{Uh}, VwStash<NF, TE>{st2}, rh, sh);
    } else {                                         // residual: X and U in the stash
      {
        double Uh[8][NF];
        vw_analyse<NF>(u, p.wr, Uh);
        vw_stash<NF, TE>(Uh, st2);
      }
      vec_walsh_points<NF, Phys, MODE, 0>(p.tab, p, e, VwStash<3, TE>{stash}, VwStash<NF, TE>{st2}, VwStash<NF, TE>{st2}, rh, sh);
    }
  } else {
    double Xh[8][3], Uh[8][NF], Vh[8][NF];
    vw_analyse<3>(x, p.wr, Xh);
    vw_analyse<NF>(u, p.wr, Uh);
    if constexpr (MODE == MODE_ACTION_STIFFNESS) vw_analyse<NF>(v, p.wr, Vh);
    vec_walsh_points<NF, Phys, MODE, 0>(p.tab, p, e, VwReg<3>{Xh}, VwReg<NF>{Uh}, VwReg<NF>{Vh}, rh, sh);
  }
#pragma unroll
  for (int d = 0; d < NF; ++d) {
    double w[8];
    w[0] = 0.0;
#pragma unroll
    for (int al = 1; al < 8; ++al) w[al] = rh[al][d] * p.wr[vw_popc3(al) - 1];
    if constexpr (MODE == MODE_RESIDUAL && Phys::kHasSource) {
#pragma unroll
      for (int al = 0; al < 8; ++al) w[al] = fma(sh[al][d], p.wr[vw_popc3(al)], w[al]);
    }
    // synthesis over the sign bits: out[i] = sum_alpha s_i^alpha w[alpha]
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (!(i & (1 << k))) {
          const double lo = w[i], hi = w[i | (1 << k)];
          w[i] = lo - hi;
          w[i | (1 << k)] = lo + hi;
        }
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i][d] = w[i];
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int MODE, int TE, int MINB, bool WALSH = false>
__global__ void __launch_bounds__(TE, MINB) k_vec(const __grid_constant__ VecParams<ND, NNPE, NQT> p) {
  extern __shared__ double smem[];
  constexpr bool kNeedV = (MODE == MODE_ACTION_STIFFNESS || MODE == MODE_ACTION_MASS);
  constexpr int NS = Phys::NS;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int nb = p.tile_node_ptr[tile];
  const int nn = p.tile_node_ptr[tile + 1] - nb;

  // ---- 1. gather the tile's nodes once; the scatter metadata of phase 4 (node ids, incidence lists) is
  // fetched now as well, so that phase has no global-load latency chain left (it was ~25 % of the kernel)
  double* sX = smem;
  double* sU = sX + nn * ND;
  double* sV = sU + nn * NF;
  int32_t* sNode = reinterpret_cast<int32_t*>(smem + p.body_doubles);
  int32_t* sIncPtr = sNode + p.max_nodes;
  uint16_t* sInc = reinterpret_cast<uint16_t*>(sIncPtr + p.max_nodes + 1);
  const int k_base = p.inc_ptr[nb];
  for (int i = tid; i < nn; i += TE) {
    const int n = p.tile_nodes[nb + i];
    sNode[i] = n;
    sIncPtr[i] = p.inc_ptr[nb + i] - k_base;
#pragma unroll
    for (int j = 0; j < ND; ++j) sX[i * ND + j] = p.X[(size_t)n * ND + j];
#pragma unroll
    for (int d = 0; d < NF; ++d) sU[i * NF + d] = p.U[(size_t)n * NF + d];
    if constexpr (kNeedV) {
#pragma unroll
      for (int d = 0; d < NF; ++d) sV[i * NF + d] = p.V[(size_t)n * NF + d];
    }
  }
  const int n_inc = p.inc_ptr[nb + nn] - k_base;
  if (tid == 0) sIncPtr[nn] = n_inc;
  for (int k = tid; k < n_inc; k += TE) sInc[k] = p.inc[k_base + k];
  __syncthreads();

  // ---- 2. element-level fields into registers (_element_level_fields_flat, Assemblers.jl:161-188)
  const int e = tile * TE + tid;
  const bool active = e < p.ne;
  double x[NNPE][ND], u[NNPE][NF], v[NNPE][NF];
  if (active) {
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
      const int l = p.lconn[((size_t)tile * NNPE + (WALSH ? p.nos[a] : a)) * TE + tid];   // Walsh form: fields in sign order
#pragma unroll
      for (int j = 0; j < ND; ++j) x[a][j] = sX[l * ND + j];
#pragma unroll
      for (int d = 0; d < NF; ++d) u[a][d] = sU[l * NF + d];
#pragma unroll
      for (int d = 0; d < NF; ++d) v[a][d] = kNeedV ? sV[l * NF + d] : 0.0;
    }
  }
  __syncthreads();  // node data consumed; shared memory is re-used as the element-vector stage

  // ---- 3. quadrature loop in registers
  double r[NNPE][NF];
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int d = 0; d < NF; ++d) r[a][d] = 0.0;
  if constexpr (WALSH) {
    constexpr bool kStash = (NF == 3) && (MODE == MODE_ACTION_STIFFNESS || FEC_VEC_STASH_R);
    if (active) vec_element_walsh<NF, Phys, MODE, kStash, TE>(p, e, x, u, v, r, smem + tid);
  } else if (active) {
    // TET10 mechanics: the element fields (3 x 30 doubles) wait in the thread's column of shared memory during the
    // quadrature loop instead of in registers (the element-vector stage is idle until phase 4)
    constexpr bool kQS = vec_qstash<NNPE, NF, WALSH>();
    double* const st = smem + tid;
    if constexpr (kQS) {
#pragma unroll
      for (int a = 0; a < NNPE; ++a) {
#pragma unroll
        for (int j = 0; j < ND; ++j) st[(a * ND + j) * TE] = x[a][j];
#pragma unroll
        for (int d = 0; d < NF; ++d) st[(NNPE * ND + a * NF + d) * TE] = u[a][d];
        if constexpr (kNeedV) {
#pragma unroll
          for (int d = 0; d < NF; ++d) st[(NNPE * (ND + NF) + a * NF + d) * TE] = v[a][d];
        }
      }
    }
    auto body = [&](const int q) {
      if constexpr (kQS) asm volatile("" ::: "memory");   // keep the stash loads of later points from being hoisted
      double so[NS > 0 ? NS : 1], sn[NS > 0 ? NS : 1];
      if constexpr (NS > 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
      }
      double fq = 0.0;
      if constexpr (Phys::kHasSource && MODE == MODE_RESIDUAL) {
        if (p.source) fq = p.source[(size_t)q * p.ne + e];
      }
      if constexpr (kQS)
        vec_qp<ND, NNPE, NF, Phys, MODE>(p.tab, q, FieldStash<ND, TE>{st}, FieldStash<NF, TE>{st + NNPE * ND * TE},
                                         FieldStash<NF, TE>{st + NNPE * (ND + NF) * TE}, p.props, fq, so,
                                         (NS > 0 && MODE == MODE_RESIDUAL) ? sn : nullptr, r);
      else
      vec_qp<ND, NNPE, NF, Phys, MODE>(p.tab, q, FieldReg<NNPE, ND>{x}, FieldReg<NNPE, NF>{u}, FieldReg<NNPE, NF>{v}, p.props, fq, so,
                                       (NS > 0 && MODE == MODE_RESIDUAL) ? sn : nullptr, r);
      if constexpr (NS > 0 && MODE == MODE_RESIDUAL) {
#pragma unroll
        for (int s = 0; s < NS; ++s) p.state_new[((size_t)s * p.nq + q) * p.ne + e] = sn[s];
      }
    };
    if constexpr (NQT > 0) {
#pragma unroll
      for (int q = 0; q < NQT; ++q) body(q);
    } else {
      for (int q = 0; q < p.nq; ++q) body(q);
    }
  }

  // ---- 4. stage element vectors, reduce per node in fixed order, one red per (node, dof)
  double* sR = smem;
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int d = 0; d < NF; ++d) sR[((WALSH ? p.nos[a] : a) * NF + d) * TE + tid] = r[a][d];
  __syncthreads();
  for (int i = tid; i < nn; i += TE) {
    double acc[NF];
#pragma unroll
    for (int d = 0; d < NF; ++d) acc[d] = 0.0;
    const int k0 = sIncPtr[i], k1 = sIncPtr[i + 1];
    for (int k = k0; k < k1; ++k) {
      const int s = sInc[k];  // = a*NF*TE + t
#pragma unroll
      for (int d = 0; d < NF; ++d) acc[d] += sR[s + d * TE];
    }
    const int n = sNode[i];
#pragma unroll
    for (int d = 0; d < NF; ++d) scatter_add(p.peer, p.out, n, NF, d, acc[d]);  // RED.E.ADD.F64 (local) / RED.SYS (ghost)
  }
}

// ------------------------------------------------------------------------------------------------
// Quadrature-point scalars: assemble_scalar!(asm, energy, Uu, p)  (src/assemblers/QuadratureQuantity.jl:4-45):
//   storage[1, q, e] = JxW * e_q   (_accumulate_q_value(::AssembledScalar, ...), Assemblers.jl:47-51) -- no scatter.
// One thread per element, direct gathers; out is [q * ne + e] in tile order (coalesced), un-permuted at the ABI.
// ------------------------------------------------------------------------------------------------
template <int ND, int NNPE, int NQT>
struct ScalarParams {
  const double* X;
  const double* U;
  const int32_t* conn;       // [ne*NNPE] tile-ordered global node ids
  const double* state_old;
  const double* source;      // [q*ne + e] or nullptr
  double* out;               // [q*ne + e]
  int32_t ne, nq;
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};

template <int ND, int NNPE, int NF, int NQT, class Phys>
__global__ void __launch_bounds__(128) k_energy(const __grid_constant__ ScalarParams<ND, NNPE, NQT> p) {
  constexpr int NS = Phys::NS;
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= p.ne) return;
  double x[NNPE][ND], u[NNPE][NF];
#pragma unroll
  for (int a = 0; a < NNPE; ++a) {
    const int n = p.conn[(size_t)e * NNPE + a];
#pragma unroll
    for (int j = 0; j < ND; ++j) x[a][j] = p.X[(size_t)n * ND + j];
#pragma unroll
    for (int d = 0; d < NF; ++d) u[a][d] = p.U[(size_t)n * NF + d];
  }
  const int nq = (NQT > 0) ? NQT : p.nq;
#pragma unroll 1
  for (int q = 0; q < nq; ++q) {
    double J[ND][ND];
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; ++a) s = fma(x[a][i], p.tab.dN[q][a][j], s);
        J[i][j] = s;
      }
    double Ji[ND][ND];
    const double JxW = invert<ND>(J, Ji) * p.tab.w[q];
    double gx[NF][ND], gu[NF][ND], uq[NF];
#pragma unroll
    for (int d = 0; d < NF; ++d) {
      double s0 = 0.0;
#pragma unroll
      for (int a = 0; a < NNPE; ++a) s0 = fma(p.tab.N[q][a], u[a][d], s0);
      uq[d] = s0;
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; ++a) s = fma(u[a][d], p.tab.dN[q][a][j], s);
        gx[d][j] = s;
      }
    }
#pragma unroll
    for (int d = 0; d < NF; ++d)
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j) s = fma(gx[d][j], Ji[j][k], s);
        gu[d][k] = s;
      }
    double so[NS > 0 ? NS : 1];
    if constexpr (NS > 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
    }
    const double fq = (Phys::kHasSource && p.source) ? p.source[(size_t)q * p.ne + e] : 0.0;
    p.out[(size_t)q * p.ne + e] = JxW * Phys::energy(gu, uq, fq, p.props, so);
  }
}

// ------------------------------------------------------------------------------------------------
// Matrix kernel
// ------------------------------------------------------------------------------------------------
template <int ND, int NNPE, int NQT>
struct MatParams {
  const double* X;
  const double* U;
  double* nz;
  const int32_t* conn;      // [ne*NNPE] tile-ordered global node ids (gathers)
  const int32_t* sconn;     // [ne*NNPE] the same with periodic side-b nodes replaced by their side-a node (scatter)
  const uint8_t* epos;      // [ne*NNPE*NNPE]  epos[(e*NNPE + b)*NNPE + a] = position of node a in adj row of node b
  const int32_t* adjptr;
  const uint16_t* coloff;
  const uint8_t* freemask;
  const int64_t* rowstart;
  const double* state_old;
  int32_t ne, nq;
  ZeroFill zf;
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};

// KIND: FECB200_STIFFNESS / FECB200_MASS.  TRANS selects which triangle convention is produced:
// the reference labels COO slot (i,j) with (row=dof_conn[i], col=dof_conn[j]) but stores K_el.data
// column-major, i.e. global K[dof_i, dof_j] += K_el[j, i]  (Assemblers.jl:109-124 vs
// SparsityPatterns.jl:76-83).  Thread b owns block-column b of K_el and therefore block-ROW
// conn[b] of the global matrix: contiguous CSR rows.  For CSC output the same addressing is used on the
// transposed element matrix (TRANS = true).
template <int ND, int NNPE, int NF, int NQT, class Phys, int KIND, int EPB, bool TRANS>
__global__ void __launch_bounds__(EPB * NNPE) k_mat(const __grid_constant__ MatParams<ND, NNPE, NQT> p) {
  extern __shared__ double smem[];
  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];
  zero_fill_begin(p.zf, zero_page);
  constexpr int NDF = NF * ND;
  constexpr int NS = Phys::NS;
  constexpr int SLOT = (KIND == FECB200_MASS) ? (NNPE + 1) : (NNPE * ND + NDF * NDF);
  const int tid = threadIdx.x;
  const int el = tid / NNPE, r = tid % NNPE;
  const int e = blockIdx.x * EPB + el;
  const bool active = e < p.ne;
  double* myslots = smem + (size_t)el * NNPE * SLOT;

  int conn[NNPE];
  double x[NNPE][ND], u[NNPE][NF];
  if (active) {
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
      const int n = p.conn[(size_t)e * NNPE + a];
      conn[a] = n;
#pragma unroll
      for (int j = 0; j < ND; ++j) x[a][j] = p.X[(size_t)n * ND + j];
#pragma unroll
      for (int d = 0; d < NF; ++d) u[a][d] = p.U[(size_t)n * NF + d];
    }
  }

  double acc[NNPE][NF][NF];  // K_el[(a,d1),(b=r,d2)]  (or its transpose for TRANS)
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) acc[a][i][j] = 0.0;

  const int nq = (NQT > 0) ? NQT : p.nq;
  for (int q0 = 0; q0 < nq; q0 += NNPE) {
    const int q = q0 + r;
    if (active && q < nq) {
      // geometry of quadrature point q (thread r of the element)
      double J[ND][ND];
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < NNPE; ++a) s = fma(x[a][i], p.tab.dN[q][a][j], s);
          J[i][j] = s;
        }
      double Ji[ND][ND];
      const double JxW = invert<ND>(J, Ji) * p.tab.w[q];
      double* slot = myslots + (size_t)r * SLOT;
      if constexpr (KIND == FECB200_MASS) {
#pragma unroll
        for (int a = 0; a < NNPE; ++a) slot[a] = p.tab.N[q][a];
        slot[NNPE] = JxW * Phys::density(p.props);
      } else {
        double gu[NF][ND];
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int k = 0; k < ND; ++k) gu[d][k] = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
          double g[ND];
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < ND; ++j) s = fma(p.tab.dN[q][a][j], Ji[j][k], s);
            g[k] = s;
            slot[a * ND + k] = s;
          }
#pragma unroll
          for (int d = 0; d < NF; ++d)
#pragma unroll
            for (int k = 0; k < ND; ++k) gu[d][k] = fma(u[a][d], g[k], gu[d][k]);
        }
        double so[NS > 0 ? NS : 1];
        if constexpr (NS > 0) {
#pragma unroll
          for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
        }
        double A[NDF][NDF];
        Phys::tangent(gu, p.props, so, A);
#pragma unroll
        for (int i = 0; i < NDF; ++i)
#pragma unroll
          for (int j = 0; j < NDF; ++j) slot[NNPE * ND + i * NDF + j] = A[i][j] * JxW;
      }
    }
    __syncthreads();
    if (active) {
      const int nc = (nq - q0) < NNPE ? (nq - q0) : NNPE;
      for (int c = 0; c < nc; ++c) {
        const double* slot = myslots + (size_t)c * SLOT;
        if constexpr (KIND == FECB200_MASS) {
          const double f = slot[NNPE] * slot[r];
#pragma unroll
          for (int a = 0; a < NNPE; ++a) {
            const double m = f * slot[a];
#pragma unroll
            for (int d = 0; d < NF; ++d) acc[a][d][d] += m;
          }
        } else {
          double gb[ND];
#pragma unroll
          for (int k = 0; k < ND; ++k) gb[k] = slot[r * ND + k];
          const double* A = slot + NNPE * ND;
          // s[(d1,j1)][d2] = sum_j2 A[(d1,j1)][(d2,j2)] gb[j2]      (TRANS: A[(d2,j2)][(d1,j1)])
          double s[NDF][NF];
#pragma unroll
          for (int i = 0; i < NDF; ++i)
#pragma unroll
            for (int d2 = 0; d2 < NF; ++d2) {
              double t = 0.0;
#pragma unroll
              for (int j2 = 0; j2 < ND; ++j2)
                t = fma(TRANS ? A[(d2 * ND + j2) * NDF + i] : A[i * NDF + d2 * ND + j2], gb[j2], t);
              s[i][d2] = t;
            }
#pragma unroll
          for (int a = 0; a < NNPE; ++a) {
            double ga[ND];
#pragma unroll
            for (int k = 0; k < ND; ++k) ga[k] = slot[a * ND + k];
#pragma unroll
            for (int d1 = 0; d1 < NF; ++d1)
#pragma unroll
              for (int d2 = 0; d2 < NF; ++d2) {
                double t = acc[a][d1][d2];
#pragma unroll
                for (int j1 = 0; j1 < ND; ++j1) t = fma(ga[j1], s[d1 * ND + j1][d2], t);
                acc[a][d1][d2] = t;
              }
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- scatter block-row conn[r]: slot = rowstart[row dof] + coloff[adj entry] + rank of d1 among kept dofs
  if (active) {
    const int32_t* sc = p.sconn + (size_t)e * NNPE;  // == conn unless periodic BCs fold side b into side a
    const int nb = sc[r];
    const int abase = p.adjptr[nb];
    const uint8_t* ep = p.epos + ((size_t)e * NNPE + r) * NNPE;
    int64_t rs[NF];
#pragma unroll
    for (int d2 = 0; d2 < NF; ++d2) rs[d2] = p.rowstart[(size_t)nb * NF + d2];
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
      const int off = p.coloff[abase + ep[a]];
      const unsigned mask = p.freemask[sc[a]];
#pragma unroll
      for (int d2 = 0; d2 < NF; ++d2) {
        if (rs[d2] < 0) continue;
#pragma unroll
        for (int d1 = 0; d1 < NF; ++d1) {
          if (mask & (1u << d1)) {
            const int rank = __popc(mask & ((1u << d1) - 1u));
            atomicAdd(&p.nz[rs[d2] + off + rank], acc[a][d1][d2]);
          }
        }
      }
    }
  }
  zero_fill_end(p.zf);
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
template <int ND, int NNPE, int NQT>
void fill_tables(const BlockPlan& b, Tables<ND, NNPE, NQT>& t) {
  constexpr int NQ = Tables<ND, NNPE, NQT>::NQ;
  FEC_REQUIRE(b.nq <= NQ, "too many quadrature points for this kernel instantiation");
  memset(&t, 0, sizeof(t));
  for (int q = 0; q < b.nq; ++q) {
    for (int a = 0; a < NNPE; ++a) {
      t.N[q][a] = b.N[(size_t)q * NNPE + a];
      for (int j = 0; j < ND; ++j) t.dN[q][a][j] = b.dN[((size_t)q * NNPE + a) * ND + j];
    }
    t.w[q] = b.w[q];
  }
}

inline void timing_begin(fecb200_handle* h) {
  if (h->timing) FEC_CUDA(cudaEventRecord(h->ev0, h->stream));
}
inline void timing_end(fecb200_handle* h) {
  if (h->timing) {
    FEC_CUDA(cudaEventRecord(h->ev1, h->stream));
    FEC_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    FEC_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms += ms;
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int MODE, int TE, int MINB, bool WALSH>
void run_vec_t(fecb200_handle* h, BlockPlan& b, const VecLaunch& a, const double walsh_c) {
  FEC_REQUIRE(b.te == TE, "tile size does not match the compiled kernel");
  auto pp = std::make_unique<VecParams<ND, NNPE, NQT>>();  // large: keep off the stack
  auto& p = *pp;
  p.X = h->d_X.p; p.U = a.U; p.V = a.V; p.out = a.out;
  p.tile_node_ptr = b.d_tile_node_ptr.p; p.tile_nodes = b.d_tile_nodes.p; p.lconn = b.d_lconn.p;
  p.inc_ptr = b.d_inc_ptr.p; p.inc = b.d_inc.p;
  p.state_old = b.d_state_old.p; p.state_new = b.d_state_new.p; p.source = b.d_source.p;
  p.peer = h->peer;
  if (!h->peer_enabled || a.out != (h->peer_field == FECB200_FIELD_RESIDUAL ? h->d_R.p : h->d_Av.p)) p.peer.n_owned = -1;
  p.ne = (int32_t)b.ne; p.nq = b.nq;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  for (int n = 0; n < 4; ++n) p.wr[n] = std::pow(walsh_c, n) / 8.0;
  for (int i = 0; i < 8; ++i) { p.nos[i] = WALSH ? b.node_of_sign[i] : i; p.pos[i] = WALSH ? b.point_of_sign[i] : i; }
  const int nfields = (MODE == MODE_ACTION_STIFFNESS || MODE == MODE_ACTION_MASS) ? 2 : 1;
  size_t sm_nodes = (size_t)b.max_tile_nodes * (ND + nfields * NF) * sizeof(double);
  size_t sm_stage = (size_t)NNPE * NF * TE * sizeof(double);
  size_t body = sm_nodes > sm_stage ? sm_nodes : sm_stage;
  if (WALSH && NF == 3) body = std::max(body, (size_t)(7 * 3 + 7 * NF) * TE * sizeof(double));   // the coefficient stash
  if (vec_qstash<NNPE, NF, WALSH>()) body = std::max(body, (size_t)NNPE * (ND + nfields * NF) * TE * sizeof(double));   // the field stash
  p.body_doubles = (int32_t)(body / sizeof(double));
  p.max_nodes = b.max_tile_nodes;
  size_t smem = body + (size_t)(2 * b.max_tile_nodes + 1) * sizeof(int32_t) + (size_t)NNPE * TE * sizeof(uint16_t) + 8;
  auto kern = k_vec<ND, NNPE, NF, NQT, Phys, MODE, TE, MINB, WALSH>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  timing_begin(h);
  kern<<<b.ntiles, TE, smem, h->stream>>>(p);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int MODE, int TE, int MINB>
void run_vec(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  // Walsh form where it pays (B200, 192^3 / 128^3): mechanics residual 2.23 -> 1.82 ms; scalar residual 0.339 -> 0.324 ms and
  // action 0.285 -> 0.257 ms once the kernel is capped at 168 registers (3 CTAs per SM; at 184 registers it LOSES, 0.45 ms).
  // The mechanics action is register bound either way (3.16 ms both, more spills in the Walsh form) and keeps the quadrature
  // loop; FECB200_VEC_WALSH_ALL=1 is the A/B switch.
  constexpr bool kPays = NF == 3 || NF == 1;
  constexpr int WMINB = (NF == 1 || (FEC_VEC_STASH_R && NF == 3 && MODE == MODE_RESIDUAL)) ? MINB + 1 : MINB;
  if constexpr (ND == 3 && NNPE == 8 && NQT == 8 && (MODE == MODE_RESIDUAL || MODE == MODE_ACTION_STIFFNESS)) {
    if (!kPays && !getenv("FECB200_VEC_WALSH_ALL")) { run_vec_t<ND, NNPE, NF, NQT, Phys, MODE, TE, MINB, false>(h, b, a, 0.0); return; }
    if (!getenv("FECB200_VEC_CLASSIC") && b.walsh) {
      run_vec_t<ND, NNPE, NF, NQT, Phys, MODE, TE, WMINB, true>(h, b, a, b.walsh_c);
      return;
    }
  }
  // field stash (TET10 mechanics, J2 64^3 x 6): the residual fits 168 registers with 72 B of spills -> a third CTA per SM,
  // 0.601 -> 0.495 ms; the action spills 336 B there (0.749 ms) and is better off at 2 CTAs per SM (0.757 -> 0.697 ms)
  constexpr int QMINB = (vec_qstash<NNPE, NF, false>() && MODE == MODE_RESIDUAL) ? MINB + FEC_VEC_QSTASH_MINB : MINB;
  run_vec_t<ND, NNPE, NF, NQT, Phys, MODE, TE, QMINB, false>(h, b, a, 0.0);
}

template <int ND, int NNPE, int NF, int NQT, class Phys>
void run_energy(fecb200_handle* h, BlockPlan& b, const double* U) {
  if constexpr (!Phys::kHasEnergy) {
    throw Error("fecb200: no energy is defined for this physics");
  } else {
    auto pp = std::make_unique<ScalarParams<ND, NNPE, NQT>>();
    auto& p = *pp;
    if (b.d_scalar.n != (size_t)b.nq * b.ne) b.d_scalar.alloc((size_t)b.nq * b.ne);
    p.X = h->d_X.p; p.U = U; p.conn = b.d_conn_perm.p; p.state_old = b.d_state_old.p; p.source = b.d_source.p;
    p.out = b.d_scalar.p; p.ne = (int32_t)b.ne; p.nq = b.nq;
    for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
    fill_tables<ND, NNPE, NQT>(b, p.tab);
    timing_begin(h);
    k_energy<ND, NNPE, NF, NQT, Phys><<<(int)((b.ne + 127) / 128), 128, 0, h->stream>>>(p);
    FEC_CUDA(cudaGetLastError());
    timing_end(h);
    h->launches++;
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int KIND, int EPB, bool TRANS>
void run_mat_t(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  auto pp = std::make_unique<MatParams<ND, NNPE, NQT>>();
  auto& p = *pp;
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz;
  p.conn = b.d_conn_perm.p; p.epos = b.d_epos.p;
  p.sconn = b.d_sconn_perm.p ? b.d_sconn_perm.p : b.d_conn_perm.p;
  p.adjptr = h->d_adjptr.p; p.coloff = h->d_coloff.p; p.freemask = h->d_freemask.p; p.rowstart = h->d_rowstart.p;
  p.state_old = b.d_state_old.p;
  p.ne = (int32_t)b.ne; p.nq = b.nq;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  constexpr int NDF = NF * ND;
  constexpr int SLOT = (KIND == FECB200_MASS) ? (NNPE + 1) : (NNPE * ND + NDF * NDF);
  size_t smem = (size_t)EPB * NNPE * SLOT * sizeof(double);
  auto kern = k_mat<ND, NNPE, NF, NQT, Phys, KIND, EPB, TRANS>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)((b.ne + EPB - 1) / EPB);
  p.zf = make_zero_fill(a, grid);
  timing_begin(h);
  kern<<<grid, EPB * NNPE, smem, h->stream>>>(p);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int EPB>
void run_mat(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  const bool trans = (h->opts.matrix_type == FECB200_CSC);
  if (a.kind == FECB200_MASS) {
    // the mass matrix is symmetric: one instantiation serves CSR and CSC
    run_mat_t<ND, NNPE, NF, NQT, Phys, FECB200_MASS, EPB, false>(h, b, a);
  } else if (trans) {
    run_mat_t<ND, NNPE, NF, NQT, Phys, FECB200_STIFFNESS, EPB, true>(h, b, a);
  } else {
    run_mat_t<ND, NNPE, NF, NQT, Phys, FECB200_STIFFNESS, EPB, false>(h, b, a);
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int TE, int MINB>
void run_vec_modes(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  switch (a.mode) {
    case MODE_RESIDUAL: run_vec<ND, NNPE, NF, NQT, Phys, MODE_RESIDUAL, TE, MINB>(h, b, a); break;
    case MODE_ACTION_STIFFNESS: run_vec<ND, NNPE, NF, NQT, Phys, MODE_ACTION_STIFFNESS, TE, MINB>(h, b, a); break;
    case MODE_ACTION_MASS: run_vec<ND, NNPE, NF, NQT, Phys, MODE_ACTION_MASS, TE, MINB>(h, b, a); break;
    case MODE_LUMPED_MASS: run_vec<ND, NNPE, NF, NQT, Phys, MODE_LUMPED_MASS, TE, MINB>(h, b, a); break;
    case MODE_DIAG_MASS: run_vec<ND, NNPE, NF, NQT, Phys, MODE_DIAG_MASS, TE, MINB>(h, b, a); break;
    case MODE_DIAG_STIFFNESS: run_vec<ND, NNPE, NF, NQT, Phys, MODE_DIAG_STIFFNESS, TE, MINB>(h, b, a); break;
    default: throw Error("fecb200: bad vector mode");
  }
}


}  // namespace fec
