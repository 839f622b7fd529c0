// kernel_mat2.cuh -- second-generation tangent kernel ("pair owner") for NF = ND mechanics.
//
// Why (ncu, profiles/r01a_kmat_details.txt): the first kernel (k_mat, one thread per block-column)
// spends its time in the L1/TEX pipe -- 3.8e9 RED sectors (every lane of every RED its own 32 B
// sector) plus 3.5e9 shared-memory wavefronts (108 LDS per 297 DFMA) -- with the FP64 pipe 37 % busy.
//
// Design:
//   * NP = NF(NF+1)/2 threads per element, one per component pair (d1 <= d2); the tangents of all shipped
//     physics have major symmetry (A_iJkL = A_kLiJ, checked in tests/test_oracle_pins.py), so pair (d1,d2)
//     also yields its mirror.  Thread (d1,d2) owns M[a][b] = K_el[(a,d1),(b,d2)] for all NNPE^2 node
//     pairs in registers and needs only dN_X (NNPE*ND) + its ND x ND block of A per quadrature point:
//     33 LDS per 264 DFMA instead of 108 per 297.
//   * 32/NP elements per warp; an element never leaves its warp, so the kernel needs __syncwarp only.
//   * phase G: the element's threads split its quadrature points, compute dN_X and JxW*A once and park
//     them in shared memory (A packed symmetric).
//   * phase K: register accumulation over the quadrature points.
//   * phase S: K_el is staged in shared memory in GLOBAL order (rows = dofs of the row node, columns
//     sorted by global node id), then the warp issues REDs over flat (row, col) indices: consecutive
//     lanes hit consecutive CSR slots (runs of NF * #adjacent nodes doubles), 4-5x fewer sectors per RED.
//   * the element -> CSR slot map is one contiguous 240-byte record per element (row offsets, column offsets,
//     masks, node ranks) built on the device at update_dofs (k_build_emeta) and fetched with cp.async.
#pragma once
#include "kernels.cuh"
#include <cstdlib>

namespace fec {

template <int ND, int NNPE, int NQT>
struct Mat2Params {
  const double* X;
  const double* U;
  double* nz;
  const int32_t* conn;       // [ne*NNPE] tile-ordered global node ids
  const unsigned char* emeta;  // [ne * REC] per-element scatter record, see Mat2Layout (built by k_build_emeta)
  const double* state_old;
  double* state_new;         // written when the residual is fused (stateful physics)
  double* R;                 // fused residual target (full-length field) or nullptr
  int64_t nnz;               // the trash region of the branch-free RED stream starts at nz[nnz]
  PeerScatter peer;          // ghost rows of the fused residual go to their owner over NVLink
  int32_t ne, nq;
  int32_t ko;                // knock-out mask of the phase-cost experiment (only read when built with -DFEC_MAT2_KO)
  ZeroFill zf;               // in-kernel clear of the idle CSR value buffer (common.cuh)
  double wc[5];              // Walsh path: c^n / 64, n = 0..4 (c = |xi| of the 2-point rule per axis, read off the tables)
  double wr[3];              //             c^n / 8,  n = 0..2 (fused residual)
  int32_t qslot[8];          //             quadrature point -> sign index (the slot its thread publishes into)
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};
// Phase-cost experiment (tools/ko_sweep.py): -DFEC_MAT2_KO compiles run-time predicates into k_mat2 that skip one
// phase at a time (FECB200_KO = 1 REDs, 2 scatter read-back + REDs, 4 staging + scatter, 8 phase K, 16 phase G, 32
// zero-fill), so the cost of each phase in the overlapped steady state can be measured.  Results are wrong by design.
#ifdef FEC_MAT2_KO
#define FEC_KO(bit) ((p.ko & (bit)) != 0)
#else
#define FEC_KO(bit) false
#endif
template <int N>
__host__ __device__ constexpr int sym_index(int i, int j) {  // packed upper triangle, i <= j
  return i * N - (i * (i - 1)) / 2 + (j - i);
}

template <int ND, int NNPE, int NF, int NQ, bool WITH_R, bool WALSH = false, bool REF = false>
struct Mat2Layout {
  static constexpr int NP = NF * (NF + 1) / 2;
  // threads per element.  Classic: one per component pair.  Walsh (FEC_MAT2_TPE, default 8): phase K is cheap enough
  // that phase G dominates, so an element gets one thread per QUADRATURE POINT (phase G is one full round of 32
  // (element, point) tasks instead of 30 + 10) and phase K / S1 run on NP of them.
#ifndef FEC_MAT2_TPE
#define FEC_MAT2_TPE 8
#endif
  static constexpr int TPE = (WALSH && FEC_MAT2_TPE > NP) ? FEC_MAT2_TPE : NP;
  static constexpr int EPW = 32 / TPE;
  static constexpr int NDF = NF * ND;
  static constexpr int ASZ = NDF * (NDF + 1) / 2;
  // classic slot: dN_X [NNPE*ND] + packed JxW*A [ASZ] [+ JxW*P].  Walsh slot: J^-1 [ND*ND] + JxW*A as NP pair blocks of
  // ND*ND (block t = (d1,d2) at OFF_A + t*ND*ND, overwritten in place by the pair thread with J^-1 A J^-T) [+ JxW*P];
  // pair stride 9 and element stride == NP (mod 16) keep the lanes (element, pair) on distinct 8-byte banks.
  // REF (laws that deliver the tangent in reference coordinates, Phys::kRefTangent): no J^-1, the blocks and P arrive
  // pulled back.
  static constexpr int OFF_A = WALSH ? (REF ? 0 : ND * ND) : NNPE * ND;
  static constexpr int OFF_P = OFF_A + (WALSH ? NP * ND * ND : ASZ);   // JxW * P (only when the residual is fused)
  static constexpr int SLOT_RAW = OFF_P + (WITH_R ? NDF : 0);
#ifndef FEC_MAT2_NOPAD
  // bank spreading (8-byte banks, 16 per wavefront): slot stride == 1 and element stride == NP (mod 16) put the
  // EPW*NP lanes that store / load "the same field of different (element, quadrature point)" on distinct banks
  static constexpr int SLOT = SLOT_RAW + (17 - SLOT_RAW % 16) % 16;
#else
  static constexpr int SLOT = SLOT_RAW;
#endif
  static constexpr int NROW = NNPE * NF;
  // row stride of the staged K_el: a bank simulation of the S1 stores (lane = (element, pair), 8-byte banks) gives
  // 384 wavefronts per warp for stride 24 against 512 for 25 with NROW = 24; the S2 loads are conflict-free either way
#ifndef FEC_MAT2_RS
#define FEC_MAT2_RS 34
#endif
  // Walsh form, 8 threads per element: lanes (element, pair (d1,d2)) store K_el[(a,d1)][(b,d2)] at row (a,d1), column
  // (b,d2); with row stride == 2 and element stride == 8 (mod 16) the 12 pair lanes of a half-warp hit the banks
  // {0,1,2,3,4,6} + 8 * (element & 1) and the mirror stores {2,4,5} + 8 * (element & 1): conflict-free, 2 wavefronts per
  // STS instead of 6 with stride 24 (the slack is free: the element's shared memory is bounded by the point slots)
  static constexpr int RSTRIDE = (WALSH && TPE == 8) ? FEC_MAT2_RS : ((NROW % 16 == 8) ? NROW : NROW + 1);
  // staging passes: the Walsh form stages and scatters K_el in two halves of rows (nodes 0..3, then 4..7) so that the
  // element's shared memory is bounded by the quadrature-point slots (520 doubles) instead of the 600 of a full K_el +
  // residual row: 568 doubles per element = 12 warps per SM with 4 elements per warp
#ifndef FEC_MAT2_HALF
#define FEC_MAT2_HALF 1
#endif
  static constexpr int NH = (WALSH && FEC_MAT2_HALF && NNPE % 2 == 0) ? 2 : 1;
  static constexpr int HROWS = NROW / NH;
  static constexpr int KSZ = HROWS * RSTRIDE;
  static constexpr int R_OFF = KSZ;                     // staged fused-residual row (NROW doubles) behind K_el
  static constexpr int BODY = (NQ * SLOT > KSZ + NROW) ? NQ * SLOT : KSZ + NROW;
  // per-element scatter record (global, contiguous; copied verbatim into shared memory with cp.async):
  //   uint32 rowstart[NROW]  (0xFFFFFFFF = row eliminated)   CSR offset of the row of dof (b, d)
  //   uint16 ecol[NNPE][NNPE]                                ecol[b][k]: column offset of local node k in row node b
  //   uint8  mask[NNPE]                                      kept-dof mask of local node k
  //   uint8  rank[NNPE]                                      (unused: identity)
  static constexpr int OFF_EC = NROW * 4;
  static constexpr int OFF_MK = OFF_EC + NNPE * NNPE * 2;
  static constexpr int OFF_RK = OFF_MK + NNPE;
  static constexpr int OFF_ND = OFF_RK + NNPE;          //   uint32 node[NNPE]: global node id of local node k
  static constexpr int REC = ((OFF_ND + 4 * NNPE + 15) / 16) * 16;
  static constexpr int META = REC / 8;
  static constexpr int BODY16 = ((BODY + 1) / 2) * 2;   // keep the record 16-byte aligned in shared memory
#ifndef FEC_MAT2_NOPAD
  static constexpr int ELSM = BODY16 + META + ((TPE + 16 - (BODY16 + META) % 16) % 16);
#else
  static constexpr int ELSM = BODY16 + META;
#endif
  static_assert(ELSM % 2 == 0, "element stride must keep 16-byte alignment");
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void red_add_f64_pred(double* addr, double v, bool ok) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p red.global.add.f64 [%0], %1; }" ::"l"(addr), "d"(v),
               "r"((int)ok)
               : "memory");
}

// ---- Walsh form of phase K for trilinear hexahedra with a symmetric 2-point rule per axis (tools/walsh/derive.py).
// With nodes a and points q labelled by their sign triples s_a, sigma_q in {+-1}^3,
//   dN_a/dxi_k (q) = 1/8 sum_{S subset of the other two axes} c^|S| s_a^({k} u S) sigma_q^S        (monomials of signs),
// so the pair block M[a][b] = sum_q sum_{k1,k2} dN[q][a][k1] B_q[k1][k2] dN[q][b][k2], B_q = J^-1 (JxW A9) J^-T, is
//   M[a][b] = sum_{alpha,beta} s_a^alpha s_b^beta Mh[alpha][beta],
//   Mh[alpha][beta] = c^(|alpha|+|beta|-2)/64 sum_{k1 in alpha, k2 in beta} Bh_{(alpha\k1) xor (beta\k2)}[k1][k2],
//   Bh_m = sum_q sigma_q^m B_q                                                     (Walsh transform over the 8 points).
// Per pair thread: 432 (B) + 216 (transform over q) + 144 (Mh) + 330 (synthesis; Mh[0][.] = Mh[.][0] = 0) = 1122 FP64
// instructions instead of the 2112 of the quadrature loop, and 49 instead of 64 accumulators.  Sign index i: bit k set
// <=> +1 on axis k.  Any node / point numbering works: the host reads the sign triples off the table (detect_walsh),
// point q publishes into slot qslot[q], K_el is staged in SIGN order and the scatter records are written in that order
// (k_build_emeta); tables of another kind take the classic loop.
// in-place 8-point transform over sign bits: (lo, hi) -> (lo + hi, hi - lo): v[m] = sum_i s_i^m v[i]
FEC_DEV void walsh_fwd8(double (&v)[8]) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(i & (1 << k))) {
        const double lo = v[i], hi = v[i | (1 << k)];
        v[i] = hi + lo;
        v[i | (1 << k)] = hi - lo;
      }
}
// in-place synthesis v[i] = sum_alpha s_i^alpha v[alpha] with v[0] == 0 on entry (its value is ignored)
FEC_DEV void walsh_syn8_z(double (&v)[8]) {
  v[0] = -v[1];                                   // (0 - v1, 0 + v1)
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    const double lo = v[i], hi = v[i + 1];
    v[i] = lo - hi;
    v[i + 1] = lo + hi;
  }
#pragma unroll
  for (int k = 1; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(i & (1 << k))) {
        const double lo = v[i], hi = v[i | (1 << k)];
        v[i] = lo - hi;
        v[i | (1 << k)] = lo + hi;
      }
}
__host__ __device__ constexpr int popc3(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }

// CTAs per SM the Walsh form is compiled for (register cap 65536 / (FEC_MAT2_MINB * WARPS * 32)); the classic loop needs
// all 255
#ifndef FEC_MAT2_MINB
#define FEC_MAT2_MINB 6
#endif
template <int ND, int NNPE, int NF, int NQT, class Phys, int WARPS, bool WITH_R, bool WALSH = false>
__global__ void __launch_bounds__(WARPS * 32, WALSH ? FEC_MAT2_MINB : 1) k_mat2(const __grid_constant__ Mat2Params<ND, NNPE, NQT> p) {
  static_assert(NQT > 0, "k_mat2 is compiled for fixed quadrature rules");
  static_assert(!WALSH || (ND == 3 && NNPE == 8 && NF == 3 && NQT == 8), "the Walsh form is the HEX8 / 2x2x2 case");
  constexpr bool REF = WALSH && Phys::kRefTangent;
  using L = Mat2Layout<ND, NNPE, NF, NQT, WITH_R, WALSH, REF>;
  constexpr int NP = L::NP, EPW = L::EPW, NDF = L::NDF, SLOT = L::SLOT, NROW = L::NROW, RS = L::RSTRIDE;
  constexpr int NS = Phys::NS;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TPE = L::TPE;
  const int elw = lane / TPE, t = lane % TPE;
  const bool lane_valid = elw < EPW;
  const int e0 = (blockIdx.x * WARPS + warp) * EPW;      // first element of this warp
  const int e = e0 + elw;
  const bool active = lane_valid && e < p.ne;
  const bool pair_active = active && t < NP;   // threads that own a component pair in phases K and S1
  double* wsm = smem + (size_t)warp * EPW * L::ELSM;
  double* esm = wsm + (size_t)(lane_valid ? elw : 0) * L::ELSM;
  // component pair of this thread: enumerate d1 <= d2
  int d1 = 0, d2 = 0;
  {
    int k = t;
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
      for (int j = i; j < NF; ++j) { if (k == 0) { d1 = i; d2 = j; } --k; }
  }

  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];

  // ---- meta: the warp's per-element scatter records are one contiguous run in global memory; fetch them
  // asynchronously (LDGSTS) so the copy overlaps phase G.  Needed from phase S1 on.
  {
    const int nel = (p.ne - e0) < EPW ? (p.ne - e0) : EPW;
    const unsigned char* g = p.emeta + (size_t)e0 * L::REC;
    constexpr int CH = L::REC / 16;
    for (int i = lane; i < nel * CH; i += 32) {
      const int el = i / CH, r = i - el * CH;
      cp_async16(reinterpret_cast<unsigned char*>(wsm + (size_t)el * L::ELSM + L::BODY16) + r * 16, g + (size_t)i * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- phase G: geometry + material tangent of the element's quadrature points, split over its threads
  double x[NNPE][ND], u[NNPE][NF];
  if (active && !FEC_KO(16)) {
    int nid[NNPE];
    if constexpr (NNPE % 4 == 0) {   // the element's connectivity row is 16-byte aligned: NNPE / 4 vector loads
      const int4* c4 = reinterpret_cast<const int4*>(p.conn + (size_t)e * NNPE);
#pragma unroll
      for (int i = 0; i < NNPE / 4; ++i) {
        const int4 v = c4[i];
        nid[4 * i] = v.x; nid[4 * i + 1] = v.y; nid[4 * i + 2] = v.z; nid[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int a = 0; a < NNPE; ++a) nid[a] = p.conn[(size_t)e * NNPE + a];
    }
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
      const int n = nid[a];
#pragma unroll
      for (int j = 0; j < ND; ++j) x[a][j] = p.X[(size_t)n * ND + j];
#pragma unroll
      for (int d = 0; d < NF; ++d) u[a][d] = p.U[(size_t)n * NF + d];
    }
  }
  if (!FEC_KO(32)) zero_fill_begin(p.zf, zero_page);   // queued while the gathers above are in flight (0.13 ms better than up front)
  if (active && !FEC_KO(16)) {
    for (int q = t; q < NQT; q += TPE) {
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
      double* slot = esm + (size_t)(WALSH ? p.qslot[q] : q) * SLOT;
      double gu[NF][ND];
      if constexpr (WALSH) {
        // publish J^-1 only; grad u = (sum_a u_a (x) dN_a/dxi) J^-1
        double H[NF][ND];
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < NNPE; ++a) s = fma(u[a][d], p.tab.dN[q][a][k], s);
            H[d][k] = s;
          }
        if constexpr (!REF) {
#pragma unroll
          for (int k = 0; k < ND; ++k)
#pragma unroll
            for (int j = 0; j < ND; ++j) slot[k * ND + j] = Ji[k][j];
        }
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int j = 0; j < ND; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < ND; ++k) s = fma(H[d][k], Ji[k][j], s);
            gu[d][j] = s;
          }
      } else {
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int k = 0; k < ND; ++k) gu[d][k] = 0.0;
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < ND; ++j) s = fma(p.tab.dN[q][a][j], Ji[j][k], s);
            slot[a * ND + k] = s;
#pragma unroll
            for (int d = 0; d < NF; ++d) gu[d][k] = fma(u[a][d], s, gu[d][k]);
          }
        }
      }
      double so[NS > 0 ? NS : 1];
      if constexpr (NS > 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
      }
      double A[NDF][NDF];
      if constexpr (REF) Phys::tangent_scaled_ref(gu, Ji, p.props, so, JxW, A);   // JxW * J^-1 A J^-T
      else Phys::tangent_scaled(gu, p.props, so, JxW, A);   // JxW * A (folded into the law's coefficients where possible)
      if constexpr (WALSH) {
        int tb = 0;
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
          for (int j = i; j < NF; ++j) {
#pragma unroll
            for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
              for (int j2 = 0; j2 < ND; ++j2) slot[L::OFF_A + tb * ND * ND + j1 * ND + j2] = A[i * ND + j1][j * ND + j2];
            ++tb;
          }
      } else {
#pragma unroll
        for (int i = 0; i < NDF; ++i)
#pragma unroll
          for (int j = i; j < NDF; ++j) slot[NNPE * ND + sym_index<NDF>(i, j)] = A[i][j];
      }
      if constexpr (WITH_R) {
        // fused residual: P at the same state (the compiler shares the kinematics with the tangent above)
        double P[NF][ND], bsrc[NF], sn[NS > 0 ? NS : 1];
        Phys::flux(gu, 0.0, p.props, so, NS > 0 ? sn : nullptr, P, bsrc);
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            if constexpr (REF) {   // pulled back: sum_j J^-1[k][j] P[d][j]
              double s = Ji[k][0] * P[d][0];
#pragma unroll
              for (int j = 1; j < ND; ++j) s = fma(Ji[k][j], P[d][j], s);
              slot[L::OFF_P + d * ND + k] = s * JxW;
            } else {
              slot[L::OFF_P + d * ND + k] = P[d][k] * JxW;
            }
          }
        if constexpr (NS > 0) {
#pragma unroll
          for (int s = 0; s < NS; ++s) p.state_new[((size_t)s * p.nq + q) * p.ne + e] = sn[s];
        }
      }
    }
  }
  __syncwarp();

  // ---- phase K: M[a][b] = sum_q sum_{j1,j2} dN_X[a][j1] A[(d1,j1)][(d2,j2)] dN_X[b][j2]
  double M[NNPE][NNPE];
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int b = 0; b < NNPE; ++b) M[a][b] = 0.0;
  double rr[WITH_R ? NNPE : 1];  // fused residual rows (a, d1) of the diagonal-pair threads (d1 == d2)
#pragma unroll
  for (int a = 0; a < (WITH_R ? NNPE : 1); ++a) rr[a] = 0.0;
  if constexpr (WALSH) {
    if (pair_active && !FEC_KO(8)) {
      const int blk = L::OFF_A + t * ND * ND;
      // pass 1: this pair's block of every point goes to reference coordinates, in place: B = J^-1 A9 J^-T
      double Ph[(WITH_R && !REF) ? NQT : 1][ND];   // !REF: pulled-back flux rows of the diagonal-pair threads
#pragma unroll
      for (int q = 0; q < (REF ? 0 : NQT); ++q) {
        double* slot = esm + (size_t)q * SLOT;
        double Ji[ND][ND], A9[ND][ND], T[ND][ND];
#pragma unroll
        for (int k = 0; k < ND; ++k)
#pragma unroll
          for (int j = 0; j < ND; ++j) Ji[k][j] = slot[k * ND + j];
#pragma unroll
        for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
          for (int j2 = 0; j2 < ND; ++j2) A9[j1][j2] = slot[blk + j1 * ND + j2];
#pragma unroll
        for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
          for (int k2 = 0; k2 < ND; ++k2) {
            double s = A9[j1][0] * Ji[k2][0];
#pragma unroll
            for (int j2 = 1; j2 < ND; ++j2) s = fma(A9[j1][j2], Ji[k2][j2], s);
            T[j1][k2] = s;
          }
#pragma unroll
        for (int k1 = 0; k1 < ND; ++k1)
#pragma unroll
          for (int k2 = 0; k2 < ND; ++k2) {
            double s = Ji[k1][0] * T[0][k2];
#pragma unroll
            for (int j1 = 1; j1 < ND; ++j1) s = fma(Ji[k1][j1], T[j1][k2], s);
            slot[blk + k1 * ND + k2] = s;
          }
        if constexpr (WITH_R) {
          if (d1 == d2) {  // (JxW P)[d1][.] pulled back the same way: Ph[k] = sum_j J^-1[k][j] P[d1][j]
            double Pd[ND];
#pragma unroll
            for (int k = 0; k < ND; ++k) Pd[k] = slot[L::OFF_P + d1 * ND + k];
#pragma unroll
            for (int k = 0; k < ND; ++k) {
              double s = Ji[k][0] * Pd[0];
#pragma unroll
              for (int j = 1; j < ND; ++j) s = fma(Ji[k][j], Pd[j], s);
              Ph[q][k] = s;
            }
          }
        }
      }
      // fused residual first (before the spectrum is live): one flux component at a time
      if constexpr (WITH_R) {
        if (d1 == d2) {  // rr[a] = sum_q sum_k dN[q][a][k] Ph_q[k] = sum_alpha s_a^alpha rh[alpha]
          double rh[8];
#pragma unroll
          for (int al = 0; al < 8; ++al) rh[al] = 0.0;
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double pq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if constexpr (REF) pq[q] = esm[(size_t)q * SLOT + L::OFF_P + d1 * ND + k];   // published pulled back
              else pq[q] = Ph[q][k];
            }
            walsh_fwd8(pq);
#pragma unroll
            for (int s1 = 0; s1 < 8; ++s1)
              if (!(s1 & (1 << k))) rh[s1 | (1 << k)] = fma(p.wr[popc3(s1)], pq[s1], rh[s1 | (1 << k)]);
          }
          walsh_syn8_z(rh);
#pragma unroll
          for (int ia = 0; ia < 8; ++ia) rr[ia] = rh[ia];
        }
      }
      // pass 2: Walsh transform over the points, entry by entry, and accumulation of the 7 x 7 spectrum Mh
#pragma unroll
      for (int k1 = 0; k1 < ND; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < ND; ++k2) {
          double bq[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) bq[q] = esm[(size_t)q * SLOT + blk + k1 * ND + k2];
          walsh_fwd8(bq);
#pragma unroll
          for (int s1 = 0; s1 < 8; ++s1)
#pragma unroll
            for (int s2 = 0; s2 < 8; ++s2)
              if (!(s1 & (1 << k1)) && !(s2 & (1 << k2))) {
                const int al = s1 | (1 << k1), be = s2 | (1 << k2);
                M[al][be] = fma(p.wc[popc3(s1) + popc3(s2)], bq[s1 ^ s2], M[al][be]);
              }
        }
      // synthesis over beta (rows alpha = 1..7), then over alpha (all columns): M[ia][ib] in sign indices
#pragma unroll
      for (int al = 1; al < 8; ++al) walsh_syn8_z(M[al]);
#pragma unroll
      for (int ib = 0; ib < 8; ++ib) {
        double col[8];
#pragma unroll
        for (int al = 0; al < 8; ++al) col[al] = M[al][ib];
        walsh_syn8_z(col);
#pragma unroll
        for (int ia = 0; ia < 8; ++ia) M[ia][ib] = col[ia];
      }
    }
  }
  if (!WALSH && active && !FEC_KO(8)) {
    // packed indices of this thread's ND x ND block (d1 <= d2 so (d1,j1) <= (d2,j2) unless d1 == d2 and j1 > j2)
    int aidx[ND][ND];
#pragma unroll
    for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
      for (int j2 = 0; j2 < ND; ++j2) {
        const int i = d1 * ND + j1, j = d2 * ND + j2;
        aidx[j1][j2] = NNPE * ND + (i <= j ? i * NDF - (i * (i - 1)) / 2 + (j - i) : j * NDF - (j * (j - 1)) / 2 + (i - j));
      }
#pragma unroll 1
    for (int q = 0; q < NQT; ++q) {
      const double* slot = esm + (size_t)q * SLOT;
      double A9[ND][ND];
#pragma unroll
      for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
        for (int j2 = 0; j2 < ND; ++j2) A9[j1][j2] = slot[aidx[j1][j2]];
      double g[NNPE][ND];
#pragma unroll
      for (int a = 0; a < NNPE; ++a)
#pragma unroll
        for (int k = 0; k < ND; ++k) g[a][k] = slot[a * ND + k];
      if constexpr (WITH_R) {
        if (d1 == d2) {  // R[a, d] += sum_j dN_X[a][j] (JxW P)[d][j]   (Formulations.jl:27-49)
          double Pd[ND];
#pragma unroll
          for (int k = 0; k < ND; ++k) Pd[k] = slot[L::OFF_P + d1 * ND + k];
#pragma unroll
          for (int a = 0; a < NNPE; ++a)
#pragma unroll
            for (int k = 0; k < ND; ++k) rr[a] = fma(g[a][k], Pd[k], rr[a]);
        }
      }
#pragma unroll
      for (int b = 0; b < NNPE; ++b) {
        double tb[ND];  // tb[j1] = sum_j2 A9[j1][j2] g[b][j2]
#pragma unroll
        for (int j1 = 0; j1 < ND; ++j1) {
          double s = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < ND; ++j2) s = fma(A9[j1][j2], g[b][j2], s);
          tb[j1] = s;
        }
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
          double s = M[a][b];
#pragma unroll
          for (int j1 = 0; j1 < ND; ++j1) s = fma(g[a][j1], tb[j1], s);
          M[a][b] = s;
        }
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp();  // every thread of the warp is done reading the slots (re-used as the K_el stage); records landed

  // ---- phases S1 + S2, once per staging pass (L::NH = 1: whole K_el; 2: rows of nodes 0..3, then of nodes 4..7).
  // S1: stage K_el.  Storage row = dof of the ROW node, storage column = (local column node, dof).
  // K_el is symmetric, so the reference's transposed COO convention and the CSR/CSC distinction do
  // not change the values.  (Sorting the columns by global node id was measured to make no difference: the RED
  // coalescer merges a warp's lanes into sectors whatever their order, so all offsets here are static.)
  // S2: REDs.  Lane = one storage column (local node k, dof dc) of the element; the warp walks the rows, so one RED
  // instruction covers one CSR row segment of the element: NROW consecutive-ish slots.  All shared-memory reads of an
  // element are issued before its REDs so their latencies overlap.  The RED stream has no per-entry tests: rows that
  // are not stored (Dirichlet dofs, ghost rows) carry a row offset inside a 4096-slot hashed trash region behind the
  // matrix, written by k_build_emeta.
  constexpr int NH = L::NH, HROWS = L::HROWS, HNODES = NNPE / NH;
#pragma unroll
  for (int hp = 0; hp < NH; ++hp) {
    if (hp > 0) __syncwarp();   // the previous pass has been read back
    if (pair_active && !FEC_KO(4)) {
#pragma unroll
      for (int a = 0; a < NNPE; ++a) {
#pragma unroll
        for (int b = 0; b < NNPE; ++b) {
          // entry (row dof (a,d1), col dof (b,d2)) and its mirror (row (b,d2), col (a,d1)); the Walsh form holds M, stages
          // K_el and reads its scatter records in sign order
          const int na = a, nb = b;
          if (na / HNODES == hp) esm[((na - hp * HNODES) * NF + d1) * RS + nb * NF + d2] = M[a][b];
          if (nb / HNODES == hp && d1 != d2) esm[((nb - hp * HNODES) * NF + d2) * RS + na * NF + d1] = M[a][b];
        }
      }
      if constexpr (WITH_R) {
        if (hp == 0 && d1 == d2) {
#pragma unroll
          for (int a = 0; a < NNPE; ++a) esm[L::R_OFF + a * NF + d1] = rr[a];  // residual row, laid out like the columns
        }
      }
    }
    __syncwarp();

    if (lane < NROW && !FEC_KO(2 | 4)) {
      const int nel = (p.ne - e0) < EPW ? (p.ne - e0) : EPW;
      const int k = lane / NF, dc = lane - k * NF;
      for (int el = 0; el < nel; ++el) {
        const double* ks = wsm + (size_t)el * L::ELSM;
        const unsigned char* rec = reinterpret_cast<const unsigned char*>(ks + L::BODY16);
        const uint32_t* rs = reinterpret_cast<const uint32_t*>(rec) + hp * HROWS;
        const uint16_t* ec = reinterpret_cast<const uint16_t*>(rec + L::OFF_EC) + hp * HNODES * NNPE;
        const unsigned mask = rec[L::OFF_MK + k];
        if (mask & (1u << dc)) {  // eliminated column (Dirichlet dof, rare): the lane sits this element out
          const int rank = __popc(mask & ((1u << dc) - 1u));
          uint32_t r0[HROWS];
          double val[HROWS];
          uint32_t off[HNODES];
#pragma unroll
          for (int b = 0; b < HNODES; ++b) off[b] = ec[b * NNPE + k] + rank;
          if constexpr (HROWS % 4 == 0) {  // the row offsets are 16-byte aligned in the record: broadcast LDS.128
            const uint4* rs4 = reinterpret_cast<const uint4*>(rs);
#pragma unroll
            for (int i = 0; i < HROWS / 4; ++i) {
              const uint4 v = rs4[i];
              r0[4 * i] = v.x; r0[4 * i + 1] = v.y; r0[4 * i + 2] = v.z; r0[4 * i + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int row = 0; row < HROWS; ++row) r0[row] = rs[row];
          }
#pragma unroll
          for (int row = 0; row < HROWS; ++row) val[row] = ks[row * RS + lane];
          if (!FEC_KO(1)) {
#pragma unroll
            for (int row = 0; row < HROWS; ++row) {  // rows that are not stored point into the trash region (k_build_emeta)
#ifdef FEC_MAT2_KO
              if (FEC_KO(64) && row % 3 == 2) continue;   // a third of the RED sectors gone: what would merging buy?
#endif
              asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p.nz + (r0[row] + off[row / NF])), "d"(val[row]));
            }
          }
#ifdef FEC_MAT2_KO
          else {  // keep the loads alive
            double sacc = 0.0;
#pragma unroll
            for (int row = 0; row < HROWS; ++row) sacc += val[row] + (double)(r0[row] + off[row / NF]);
            if (sacc == 1.234567e300) p.nz[0] = sacc;
          }
#endif
        }
        if constexpr (WITH_R) {
          // fused residual: lane (k, dc) adds the staged entry into R[node_k, dc] -- 3 consecutive doubles per node
          // (ghost nodes go to their owner over NVLink, see scatter_add)
          if (hp == 0) {
            const uint32_t n = reinterpret_cast<const uint32_t*>(rec + L::OFF_ND)[k];
            scatter_add(p.peer, p.R, (int64_t)n, NF, dc, ks[L::R_OFF + lane]);
          }
        }
      }
    }
  }
  if (!FEC_KO(32)) zero_fill_end(p.zf);
#ifdef FEC_MAT2_KO
  if (FEC_KO(4)) {  // staging knocked out: keep phase K alive
    double sacc = 0.0;
#pragma unroll
    for (int a = 0; a < NNPE; ++a)
#pragma unroll
      for (int b = 0; b < NNPE; ++b) sacc += M[a][b];
    if (sacc == 1.234567e300) p.nz[0] = sacc;
  }
#endif
}

// element -> CSR column-offset table, rebuilt whenever update_dofs changes the kept-dof masks
__global__ void k_build_emeta(const int32_t* conn, const int32_t* gconn, const uint8_t* epos, const int32_t* adjptr, const uint16_t* coloff,
                              const uint8_t* freemask, const int64_t* rowstart, unsigned char* emeta, int nnpe, int nf,
                              int rec, int64_t ne, int64_t nnz, int trash_rows, EmetaOrder ord);

template <int ND, int NNPE, int NF, int NQT, class Phys, int WARPS, bool WITH_R, bool WALSH = false>
void run_mat2_t(fecb200_handle* h, BlockPlan& b, const MatLaunch& a, double walsh_c = 0.0) {
  using L = Mat2Layout<ND, NNPE, NF, NQT, WITH_R, WALSH, WALSH && Phys::kRefTangent>;
  auto pp = std::make_unique<Mat2Params<ND, NNPE, NQT>>();
  auto& p = *pp;
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz;
  FEC_REQUIRE((int64_t)nz_alloc_len(h) < (int64_t)0xFFFFFFFFll, "k_mat2 needs nnz < 2^32 (32-bit row offsets in the scatter records)");
  FEC_REQUIRE((int)b.emeta_rec == L::REC, "scatter record size mismatch");
  p.conn = b.d_conn_perm.p; p.emeta = b.d_emeta.p;
  p.R = a.R; p.state_new = b.d_state_new.p;
  p.peer = h->peer;
  if (!h->peer_enabled || h->peer_field != FECB200_FIELD_RESIDUAL) p.peer.n_owned = -1;
  p.state_old = b.d_state_old.p;
  p.ne = (int32_t)b.ne; p.nq = b.nq; p.nnz = h->nnz;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  for (int n = 0; n < 5; ++n) p.wc[n] = std::pow(walsh_c, n) / 64.0;
  for (int n = 0; n < 3; ++n) p.wr[n] = std::pow(walsh_c, n) / 8.0;
  for (int q = 0; q < 8; ++q) p.qslot[q] = WALSH ? b.sign_of_point[q] : q;
  FEC_REQUIRE(b.emeta_sign_order == WALSH, "scatter records are not in the order this kernel stages K_el in");
#ifdef FEC_MAT2_KO
  p.ko = getenv("FECB200_KO") ? atoi(getenv("FECB200_KO")) : 0;
#endif
  size_t smem = (size_t)WARPS * L::EPW * L::ELSM * sizeof(double);
  const int epc = WARPS * L::EPW;
  const int grid = (int)((b.ne + epc - 1) / epc);
  p.zf = make_zero_fill(a, grid);
  timing_begin(h);
  // K_el is symmetric here, so CSR and CSC storage receive the same values through the same addressing
  auto kern = k_mat2<ND, NNPE, NF, NQT, Phys, WARPS, WITH_R, WALSH>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, WARPS * 32, smem, h->stream>>>(p);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int WARPS>
void run_mat2(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if constexpr (ND == 3 && NNPE == 8 && NF == 3 && NQT == 8) {
    if (b.emeta_sign_order) {   // decided with the records (build_ecol): Walsh tables and not FECB200_MAT2_CLASSIC
      if (a.R) run_mat2_t<ND, NNPE, NF, NQT, Phys, WARPS, true, true>(h, b, a, b.walsh_c);
      else run_mat2_t<ND, NNPE, NF, NQT, Phys, WARPS, false, true>(h, b, a, b.walsh_c);
      return;
    }
  }
  if (a.R) run_mat2_t<ND, NNPE, NF, NQT, Phys, WARPS, true>(h, b, a);
  else run_mat2_t<ND, NNPE, NF, NQT, Phys, WARPS, false>(h, b, a);
}

}  // namespace fec
