// kernel_mat2c.cuh -- "column-split pair owner": k_mat2 with every component pair's node-pair block M[a][b] split
// over HS = 2 threads by column half, so a thread carries 32 instead of 64 FP64 accumulators.
//
// Why (profiles/r01p_fused_kmat2_details.txt + warp-state samples): k_mat2 needs 252-255
// registers, i.e. 8 warps per SM = 2 per scheduler.  Phase K alone runs at 78 % of the DGEMM FP64 rate, but a warp is
// in phase K only ~36 % of its life (phase G 30 %, staging 8 %, REDs 26 %), and with two warps per scheduler the
// chance that at least one of them is feeding the FP64 pipe is 1 - 0.64^2 = 59 % -- the measured FP64 pipe busy is
// 54.7 %.  Halving the accumulators lets 12-15 warps share an SM (3-4 per scheduler) at the same number of resident
// elements (shared memory per element is unchanged), so other warps' phase K hides a warp's gather / scatter phases.
//
// Differences from k_mat2:
//   * NP * HS = 12 threads per element; elements span warps, CTA = EPC elements, phases separated by CTA barriers
//     (3-warp CTAs keep the barrier cheap and the CTAs of an SM out of lockstep).
//   * phase G: one thread per (element, quadrature point) packed into whole warps (EPC * NQ tasks; the remaining
//     warp idles at the barrier instead of occupying FP64 issue slots with mostly-predicated-off instructions).
//     The reference tables are read with a per-lane quadrature index, which the constant bank would serialise, so
//     this kernel keeps dN (q fastest) and w in shared memory (north_star: "reference shape-function gradients and
//     quadrature weights sit in shared memory").
//   * dN_X is stored k-major (g[k][a]) so phase K fetches node pairs with 128-bit broadcast loads:
//     27 LDS (18 of them LDS.128) per 132 DFMA.
//   * no FMA is duplicated by the split: tb[b] = A9 g[b] is computed for the thread's own 4 columns only.
#pragma once
#include "../../finiteelementcontainers.jl_b200/csrc/kernel_mat2.cuh"

namespace fec {

template <int ND, int NNPE, int NF, int NQ, bool WITH_R>
struct Mat2cLayout {
  using L2 = Mat2Layout<ND, NNPE, NF, NQ, WITH_R>;   // the scatter record (REC, OFF_*) is shared with k_mat2
  static constexpr int NP = NF * (NF + 1) / 2;
  static constexpr int HS = 2;
  static constexpr int TPE = NP * HS;
  static constexpr int NB = NNPE / HS;                  // columns (nodes b) per thread
  static constexpr int NDF = NF * ND;
  static constexpr int ASZ = NDF * (NDF + 1) / 2;
  static constexpr int OFF_A = NNPE * ND;               // packed JxW*A behind g[k][a]
  static constexpr int OFF_P = OFF_A + ASZ;             // JxW*P directly behind A (stored as one run with it)
  static constexpr int RUN = ASZ + (WITH_R ? NDF : 0);
  static constexpr int SLOT = ((OFF_A + RUN + 1) / 2) * 2;   // even: 16-byte aligned slots
  static constexpr int NROW = NNPE * NF;
  // staged K_el row stride: bank simulation of the S1 stores (lane = (element, pair, half)) gives 576 wavefronts per
  // CTA for stride 25 / element stride = 2 (mod 16) against 896 for stride 24
  static constexpr int RSTRIDE = NROW + 1;
  static constexpr int KSZ = NROW * RSTRIDE;
  static constexpr int R_OFF = KSZ;
  static constexpr int BODY = (NQ * SLOT > KSZ + NROW) ? NQ * SLOT : KSZ + NROW;
  static constexpr int BODY16 = ((BODY + 1) / 2) * 2;
  static constexpr int META = L2::REC / 8;
  static constexpr int ELSM = BODY16 + META + ((2 + 16 - (BODY16 + META) % 16) % 16);   // == 2 (mod 16)
  static constexpr int TAB = NNPE * ND * NQ + NQ;       // dN[(a*ND+j)*NQ + q], w[q]
  static_assert(NNPE % (2 * HS) == 0 && ELSM % 2 == 0 && OFF_A % 2 == 0, "128-bit shared-memory accesses need even offsets");
};

template <int ND, int NNPE, int NF, int NQT, class Phys, int EPC, int MINB, bool WITH_R>
__global__ void __launch_bounds__(EPC * Mat2cLayout<ND, NNPE, NF, NQT, WITH_R>::TPE)
__maxnreg__(MINB >= 5 ? 136 : (MINB == 4 ? 168 : 255))   // 15 / 12 warps per SM with 3-warp CTAs
k_mat2c(const __grid_constant__ Mat2Params<ND, NNPE, NQT> p) {
  using L = Mat2cLayout<ND, NNPE, NF, NQT, WITH_R>;
  using L2 = typename L::L2;
  constexpr int NP = L::NP, HS = L::HS, TPE = L::TPE, NB = L::NB, NDF = L::NDF, SLOT = L::SLOT, NROW = L::NROW, RS = L::RSTRIDE;
  constexpr int NS = Phys::NS;
  constexpr int THREADS = EPC * TPE, WARPS = THREADS / 32;
  static_assert(THREADS % 32 == 0 && EPC * NQT <= THREADS, "CTA shape");
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e0 = blockIdx.x * EPC;
  const int nel = (p.ne - e0) < EPC ? (p.ne - e0) : EPC;
  double* tabs = smem + (size_t)EPC * L::ELSM;           // dN (q fastest), then w

  // ---- scatter records of the CTA's elements: one contiguous run in global memory, fetched with LDGSTS
  {
    const unsigned char* g = p.emeta + (size_t)e0 * L2::REC;
    constexpr int CH = L2::REC / 16;
    for (int i = tid; i < nel * CH; i += THREADS) {
      const int el = i / CH, r = i - el * CH;
      cp_async16(reinterpret_cast<unsigned char*>(smem + (size_t)el * L::ELSM + L::BODY16) + r * 16, g + (size_t)i * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int i = tid; i < NNPE * ND * NQT; i += THREADS) {
    const int q = i % NQT, aj = i / NQT;
    tabs[i] = p.tab.dN[q][aj / ND][aj % ND];
  }
  if (tid < NQT) tabs[NNPE * ND * NQT + tid] = p.tab.w[tid];
  __syncthreads();

  // ---- phase G: task = (element, quadrature point)
  {
    const int el = tid / NQT, q = tid - el * NQT;
    const int e = e0 + el;
    const bool task = tid < EPC * NQT && el < nel;
    double x[NNPE][ND];
    int nd[NNPE];
    if (task) {
#pragma unroll
      for (int a = 0; a < NNPE; ++a) {
        nd[a] = p.conn[(size_t)e * NNPE + a];
#pragma unroll
        for (int j = 0; j < ND; ++j) x[a][j] = p.X[(size_t)nd[a] * ND + j];
      }
    }
    zero_fill_begin(p.zf, zero_page);   // queued while the gathers above are in flight
    if (task) {
      double* slot = smem + (size_t)el * L::ELSM + (size_t)q * SLOT;
      const double* dNq = tabs + q;
      double J[ND][ND];
#pragma unroll
      for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j) J[i][j] = 0.0;
#pragma unroll
      for (int a = 0; a < NNPE; ++a)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          const double dn = dNq[(a * ND + j) * NQT];
#pragma unroll
          for (int i = 0; i < ND; ++i) J[i][j] = fma(x[a][i], dn, J[i][j]);
        }
      double Ji[ND][ND];
      const double JxW = invert<ND>(J, Ji) * tabs[NNPE * ND * NQT + q];
      double gu[NF][ND];
#pragma unroll
      for (int d = 0; d < NF; ++d)
#pragma unroll
        for (int k = 0; k < ND; ++k) gu[d][k] = 0.0;
#pragma unroll
      for (int a2 = 0; a2 < NNPE / 2; ++a2) {
        double s[2][ND];
#pragma unroll
        for (int aa = 0; aa < 2; ++aa) {
          const int a = 2 * a2 + aa;
          double dn[ND], ua[NF];
#pragma unroll
          for (int j = 0; j < ND; ++j) dn[j] = dNq[(a * ND + j) * NQT];
#pragma unroll
          for (int d = 0; d < NF; ++d) ua[d] = p.U[(size_t)nd[a] * NF + d];
#pragma unroll
          for (int k = 0; k < ND; ++k) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < ND; ++j) t = fma(dn[j], Ji[j][k], t);
            s[aa][k] = t;
#pragma unroll
            for (int d = 0; d < NF; ++d) gu[d][k] = fma(ua[d], t, gu[d][k]);
          }
        }
#pragma unroll
        for (int k = 0; k < ND; ++k) *reinterpret_cast<double2*>(slot + k * NNPE + 2 * a2) = make_double2(s[0][k], s[1][k]);
      }
      double so[NS > 0 ? NS : 1];
      if constexpr (NS > 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
      }
      double A[NDF][NDF];
      Phys::tangent(gu, p.props, so, A);
      double P[NF][ND];
      if constexpr (WITH_R) {
        double bsrc[NF], sn[NS > 0 ? NS : 1];
        Phys::flux(gu, 0.0, p.props, so, NS > 0 ? sn : nullptr, P, bsrc);
        if constexpr (NS > 0) {
#pragma unroll
          for (int s = 0; s < NS; ++s) p.state_new[((size_t)s * p.nq + q) * p.ne + e] = sn[s];
        }
      }
      // one run of ASZ (+ NDF) doubles behind g, stored pairwise (128-bit stores)
      int cnt = 0;
      double pend = 0.0;
      auto put = [&](double v) {
        if (cnt & 1) *reinterpret_cast<double2*>(slot + L::OFF_A + cnt - 1) = make_double2(pend, v);
        else pend = v;
        ++cnt;
      };
#pragma unroll
      for (int i = 0; i < NDF; ++i)
#pragma unroll
        for (int j = i; j < NDF; ++j) put(A[i][j] * JxW);
      if constexpr (WITH_R) {
#pragma unroll
        for (int d = 0; d < NF; ++d)
#pragma unroll
          for (int k = 0; k < ND; ++k) put(P[d][k] * JxW);
      }
      if (cnt & 1) slot[L::OFF_A + cnt - 1] = pend;
    }
  }
  __syncthreads();

  // ---- phase K: thread (element, pair (d1,d2), half h) owns M[a][b], b in [h*NB, (h+1)*NB)
  const int el = tid / TPE, r = tid - el * TPE;
  const int t = r / HS, h = r - t * HS;
  const bool active = el < nel;
  double* esm = smem + (size_t)(active ? el : 0) * L::ELSM;
  int d1 = 0, d2 = 0;
  {
    int k = t;
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
      for (int j = i; j < NF; ++j) { if (k == 0) { d1 = i; d2 = j; } --k; }
  }
  double M[NNPE][NB];
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) M[a][b] = 0.0;
  double rr[WITH_R ? NB : 1];   // fused residual rows (a = h*NB + bl, d1) of the diagonal-pair threads
#pragma unroll
  for (int b = 0; b < (WITH_R ? NB : 1); ++b) rr[b] = 0.0;
  if (active) {
    // packed-symmetric index of A[(d1,j1)][(d2,j2)] = rb[j1] + j2 whenever (d1,j1) <= (d2,j2); the diagonal pairs
    // (d1 == d2) read their lower triangle through the mirror entry.  3 registers instead of a 9-entry table.
    int rb[ND];
#pragma unroll
    for (int j1 = 0; j1 < ND; ++j1) {
      const int i = d1 * ND + j1;
      rb[j1] = L::OFF_A + i * NDF - (i * (i - 1)) / 2 + (d2 * ND - i);
    }
    const bool diag = d1 == d2;
#pragma unroll 1
    for (int q = 0; q < NQT; ++q) {
      const double* slot = esm + (size_t)q * SLOT;
      double A9[ND][ND];
#pragma unroll
      for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
        for (int j2 = 0; j2 < ND; ++j2) A9[j1][j2] = slot[(j2 < j1 && diag) ? rb[j2] + j1 : rb[j1] + j2];
      double tb[NB][ND];
      double Pd[ND];
      if constexpr (WITH_R) {
        if (diag) {
#pragma unroll
          for (int k = 0; k < ND; ++k) Pd[k] = slot[L::OFF_P + d1 * ND + k];
        }
      }
#pragma unroll
      for (int b2 = 0; b2 < NB / 2; ++b2) {   // two own columns at a time (keeps the live set small)
        double gb[ND][2];
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const double2 v = *reinterpret_cast<const double2*>(slot + k * NNPE + h * NB + 2 * b2);
          gb[k][0] = v.x; gb[k][1] = v.y;
        }
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
#pragma unroll
          for (int j1 = 0; j1 < ND; ++j1) {
            double s = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < ND; ++j2) s = fma(A9[j1][j2], gb[j2][bb], s);
            tb[2 * b2 + bb][j1] = s;
          }
          if constexpr (WITH_R) {
            if (diag) {  // R[a, d] += sum_j dN_X[a][j] (JxW P)[d][j]   (Formulations.jl:27-49), rows a of this half
#pragma unroll
              for (int k = 0; k < ND; ++k) rr[2 * b2 + bb] = fma(gb[k][bb], Pd[k], rr[2 * b2 + bb]);
            }
          }
        }
      }
#pragma unroll
      for (int a2 = 0; a2 < NNPE / 2; ++a2) {
        double ga[ND][2];
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const double2 v = *reinterpret_cast<const double2*>(slot + k * NNPE + 2 * a2);
          ga[k][0] = v.x; ga[k][1] = v.y;
        }
#pragma unroll
        for (int aa = 0; aa < 2; ++aa)
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            double s = M[2 * a2 + aa][b];
#pragma unroll
            for (int j1 = 0; j1 < ND; ++j1) s = fma(ga[j1][aa], tb[b][j1], s);
            M[2 * a2 + aa][b] = s;
          }
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();  // every thread is done reading the slots (re-used as the K_el stage); records landed

  // ---- phase S1: stage K_el (row = dof of the row node, column = (local column node, dof)), as in k_mat2
  if (active) {
#pragma unroll
    for (int a = 0; a < NNPE; ++a) {
#pragma unroll
      for (int bl = 0; bl < NB; ++bl) {
        const int b = h * NB + bl;
        esm[(a * NF + d1) * RS + b * NF + d2] = M[a][bl];
        if (d1 != d2) esm[(b * NF + d2) * RS + a * NF + d1] = M[a][bl];
      }
    }
    if constexpr (WITH_R) {
      if (d1 == d2) {
#pragma unroll
        for (int bl = 0; bl < NB; ++bl) esm[L::R_OFF + (h * NB + bl) * NF + d1] = rr[bl];
      }
    }
  }
  __syncthreads();

  // ---- phase S2: REDs, one warp per element at a time (lane = storage column, the warp walks the rows)
  if (lane < NROW) {
    const int k = lane / NF, dc = lane - k * NF;
    for (int el2 = warp; el2 < nel; el2 += WARPS) {
      const double* ks = smem + (size_t)el2 * L::ELSM;
      const unsigned char* rec = reinterpret_cast<const unsigned char*>(ks + L::BODY16);
      const uint16_t* ec = reinterpret_cast<const uint16_t*>(rec + L2::OFF_EC);
      const unsigned mask = rec[L2::OFF_MK + k];
      if (mask & (1u << dc)) {  // eliminated column (Dirichlet dof, rare): the lane sits this element out
        const int rank = __popc(mask & ((1u << dc) - 1u));
        uint32_t r0[NROW];
        double val[NROW];
        uint32_t off[NNPE];
#pragma unroll
        for (int b = 0; b < NNPE; ++b) off[b] = ec[b * NNPE + k] + rank;
        static_assert(NROW % 4 == 0, "row offsets are fetched with broadcast LDS.128");
        const uint4* rs4 = reinterpret_cast<const uint4*>(rec);
#pragma unroll
        for (int i = 0; i < NROW / 4; ++i) {
          const uint4 v = rs4[i];
          r0[4 * i] = v.x; r0[4 * i + 1] = v.y; r0[4 * i + 2] = v.z; r0[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int row = 0; row < NROW; ++row) val[row] = ks[row * RS + lane];
#pragma unroll
        for (int row = 0; row < NROW; ++row)  // rows that are not stored point into the trash region (k_build_emeta)
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p.nz + (r0[row] + off[row / NF])), "d"(val[row]));
      }
      if constexpr (WITH_R) {
        const uint32_t n = reinterpret_cast<const uint32_t*>(rec + L2::OFF_ND)[k];
        scatter_add(p.peer, p.R, (int64_t)n, NF, dc, ks[L::R_OFF + lane]);
      }
    }
  }
  zero_fill_end(p.zf);
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int EPC, int MINB, bool WITH_R>
void run_mat2c_t(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  using L = Mat2cLayout<ND, NNPE, NF, NQT, WITH_R>;
  auto pp = std::make_unique<Mat2Params<ND, NNPE, NQT>>();
  auto& p = *pp;
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz;
  FEC_REQUIRE((int64_t)nz_alloc_len(h) < (int64_t)0xFFFFFFFFll, "k_mat2c needs nnz < 2^32 (32-bit row offsets in the scatter records)");
  FEC_REQUIRE((int)b.emeta_rec == L::L2::REC, "scatter record size mismatch");
  p.conn = b.d_conn_perm.p; p.emeta = b.d_emeta.p;
  p.R = a.R; p.state_new = b.d_state_new.p;
  p.peer = h->peer;
  if (!h->peer_enabled || h->peer_field != FECB200_FIELD_RESIDUAL) p.peer.n_owned = -1;
  p.state_old = b.d_state_old.p;
  p.ne = (int32_t)b.ne; p.nq = b.nq; p.nnz = h->nnz;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  const size_t smem = ((size_t)EPC * L::ELSM + L::TAB) * sizeof(double);
  const int grid = (int)((b.ne + EPC - 1) / EPC);
  p.zf = make_zero_fill(a, grid);
  timing_begin(h);
  auto kern = k_mat2c<ND, NNPE, NF, NQT, Phys, EPC, MINB, WITH_R>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, EPC * L::TPE, smem, h->stream>>>(p);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

template <int ND, int NNPE, int NF, int NQT, class Phys, int EPC, int MINB>
void run_mat2c(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if (a.R) run_mat2c_t<ND, NNPE, NF, NQT, Phys, EPC, MINB, true>(h, b, a);
  else run_mat2c_t<ND, NNPE, NF, NQT, Phys, EPC, MINB, false>(h, b, a);
}

}  // namespace fec
