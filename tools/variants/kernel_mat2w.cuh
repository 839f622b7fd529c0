// kernel_mat2w.cuh -- warp-specialised, persistent form of the pair-owner tangent kernel.
//
// Why (knock-out sweep profiles/r01u_ko_sweep.txt): in k_mat2 a warp walks its phases one after the
// other -- phase G (gathers, geometry, constitutive tangent: latency-bound, FP64 pipe 44 % busy when run alone), phase K
// (register accumulation: 69-85 % of the FP64 peak), staging + REDs (memory-bound) -- and at 8 warps per SM only the
// memory side overlaps with compute: 16.3 ms against floors of 8.7 (FP64 pipe), 11.9 (LSU wavefronts), 12.3 ms (DRAM side).
//
// Here the phases run in different warps of one persistent CTA per SM:
//   * PT producer teams of 2 warps: one thread per (element, quadrature point) of an 8-element batch -> geometry,
//     tangent and flux into a shared-memory stage (tables in shared memory, q fastest);
//   * CT consumer teams of 3 warps: the column-split pair owners of kernel_mat2c.cuh (12 threads per element, 32
//     accumulators) -> phase K, staging, REDs;
//   * a ring of NST stages (8 element slots each) with one produced / one consumed use counter per stage; the teams of a role take
//     the CTA's batches round-robin, so up to PT batches are being produced while CT are being consumed;
//   * the idle CSR value buffer is cleared batch by batch by one producer lane (TMA bulk stores), as in k_mat2.
#pragma once
#include "kernel_mat2c.cuh"

namespace fec {

// Stage hand-over: monotonic use counters in shared memory instead of mbarrier phase parities.  The teams of a role
// take the CTA's batches round-robin, so two different teams may be one and two uses ahead on the same stage; a parity
// wait cannot tell "two phases ahead" from "done" (it hung at 192^3), a counter can.  Lane 0 of a warp polls, the warp
// converges on __syncwarp.
__device__ __forceinline__ void seq_wait(const int* c, int want) {
  if ((threadIdx.x & 31) == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(c);
    int v;
    for (;;) {
      asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      if (v >= want) break;
      __nanosleep(200);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void seq_post(int* c, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(c)), "r"(v) : "memory");
}
__device__ __forceinline__ void team_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int PT_, int CT_, int NST_>
struct Mat2wShape {
  static constexpr int EPB = 8;                       // elements per batch (= per stage)
  static constexpr int PT = PT_, CT = CT_, NST = NST_;
  static constexpr int PWARPS = 2, CWARPS = 3;        // 64 tasks = 8 elements x 8 points; 96 = 8 elements x 12 threads
  static constexpr int THREADS = 32 * (PT * PWARPS + CT * CWARPS);
};

template <int ND, int NNPE, int NF, int NQT, class Phys, class SH, int MAXREG, bool WITH_R>
__global__ void __launch_bounds__(SH::THREADS) __maxnreg__(MAXREG)
k_mat2w(const __grid_constant__ Mat2Params<ND, NNPE, NQT> p) {
  using L = Mat2cLayout<ND, NNPE, NF, NQT, WITH_R>;
  using L2 = typename L::L2;
  constexpr int NP = L::NP, HS = L::HS, TPE = L::TPE, NB = L::NB, NDF = L::NDF, SLOT = L::SLOT, NROW = L::NROW, RS = L::RSTRIDE;
  constexpr int NS = Phys::NS;
  constexpr int EPB = SH::EPB, PT = SH::PT, CT = SH::CT, NST = SH::NST, PWARPS = SH::PWARPS, CWARPS = SH::CWARPS;
  static_assert(EPB * NQT == 32 * PWARPS && EPB * TPE == 32 * CWARPS, "team shapes");
  constexpr int STAGE = EPB * L::ELSM;                 // doubles per stage
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];
  __shared__ int full_seq[NST], empty_seq[NST];   // completed uses of every stage (produced / consumed)
  double* tabs = smem + (size_t)NST * STAGE;           // dN (q fastest), then w
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nbatch = (p.ne + EPB - 1) / EPB;

  if (tid < NST) { full_seq[tid] = 0; empty_seq[tid] = 0; }
  for (int i = tid; i < NNPE * ND * NQT; i += SH::THREADS) {
    const int q = i % NQT, aj = i / NQT;
    tabs[i] = p.tab.dN[q][aj / ND][aj % ND];
  }
  if (tid < NQT) tabs[NNPE * ND * NQT + tid] = p.tab.w[tid];
  if (p.zf.p != nullptr) {
    for (int i = tid; i < kZeroPageBytes / 8; i += SH::THREADS) zero_page[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (wid < PT * PWARPS) {
    // =============================== producer team ===============================
    const int team = wid / PWARPS, t = tid - team * (32 * PWARPS);
    const int el = t / NQT, q = t - el * NQT;
    const bool zlane = (t == 0) && p.zf.p != nullptr;
    for (int k = team; ; k += PT) {                    // k-th batch of this CTA
      const int64_t gb = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      if (gb >= nbatch) break;
      const int stage = k % NST, use = k / NST;
      const int e = (int)gb * EPB + el;
      const bool task = e < p.ne;
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
      if (zlane) {   // this batch's share of the idle value buffer, queued while the gathers are in flight
        const int64_t per = (p.zf.total16 + nbatch - 1) / nbatch;
        const int64_t beg = gb * per;
        int64_t rem = (p.zf.total16 - beg < per ? p.zf.total16 - beg : per) * 16;
        char* g = reinterpret_cast<char*>(p.zf.p) + beg * 16;
        const unsigned zs = (unsigned)__cvta_generic_to_shared(zero_page);
        for (; rem > 0; rem -= kZeroPageBytes, g += kZeroPageBytes) {
          const unsigned nbytes = rem < kZeroPageBytes ? (unsigned)rem : (unsigned)kZeroPageBytes;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(zs), "r"(nbytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      seq_wait(&empty_seq[stage], use);                // the consumers are done with the previous batch in this stage
      double* sbase = smem + (size_t)stage * STAGE;
      if (task) {
        double* slot = sbase + (size_t)el * L::ELSM + (size_t)q * SLOT;
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
      team_sync(1 + team, 32 * PWARPS);                // all 64 tasks of the batch are in shared memory
      if (t == 0) seq_post(&full_seq[stage], use + 1);
    }
    if (zlane) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else {
    // =============================== consumer team ===============================
    const int cw = wid - PT * PWARPS;
    const int team = cw / CWARPS, cwarp = cw - team * CWARPS;
    const int t96 = tid - 32 * (PT * PWARPS + team * CWARPS);
    const int el = t96 / TPE, r = t96 - el * TPE;
    const int t = r / HS, h = r - t * HS;
    int d1 = 0, d2 = 0;
    {
      int kk = t;
#pragma unroll
      for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = i; j < NF; ++j) { if (kk == 0) { d1 = i; d2 = j; } --kk; }
    }
    for (int k = team; ; k += CT) {
      const int64_t gb = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      if (gb >= nbatch) break;
      const int stage = k % NST, use = k / NST;
      const int e0 = (int)gb * EPB;
      const int nel = (p.ne - e0) < EPB ? (p.ne - e0) : EPB;
      const bool active = el < nel;
      double* sbase = smem + (size_t)stage * STAGE;
      double* esm = sbase + (size_t)(active ? el : 0) * L::ELSM;
      seq_wait(&full_seq[stage], use + 1);             // the producers have published this batch
      {  // scatter records of the batch (one contiguous run in global memory), needed from S2 on
        const unsigned char* g = p.emeta + (size_t)e0 * L2::REC;
        constexpr int CH = L2::REC / 16;
        for (int i = t96; i < nel * CH; i += 32 * CWARPS) {
          const int el_ = i / CH, r_ = i - el_ * CH;
          cp_async16(reinterpret_cast<unsigned char*>(sbase + (size_t)el_ * L::ELSM + L::BODY16) + r_ * 16, g + (size_t)i * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
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
      team_sync(1 + PT + team, 32 * CWARPS);  // every thread of the team is done reading the slots; records landed
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

      team_sync(1 + PT + team, 32 * CWARPS);
      if (lane < NROW) {
        const int k = lane / NF, dc = lane - k * NF;
        for (int el2 = cwarp; el2 < nel; el2 += CWARPS) {
          const double* ks = sbase + (size_t)el2 * L::ELSM;
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

      team_sync(1 + PT + team, 32 * CWARPS);           // the whole team is done with the stage
      if (t96 == 0) seq_post(&empty_seq[stage], use + 1);
    }
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, class SH, int MAXREG, bool WITH_R>
void run_mat2w_t(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  using L = Mat2cLayout<ND, NNPE, NF, NQT, WITH_R>;
  auto pp = std::make_unique<Mat2Params<ND, NNPE, NQT>>();
  auto& p = *pp;
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz;
  FEC_REQUIRE((int64_t)nz_alloc_len(h) < (int64_t)0xFFFFFFFFll, "k_mat2w needs nnz < 2^32 (32-bit row offsets in the scatter records)");
  FEC_REQUIRE((int)b.emeta_rec == L::L2::REC, "scatter record size mismatch");
  p.conn = b.d_conn_perm.p; p.emeta = b.d_emeta.p;
  p.R = a.R; p.state_new = b.d_state_new.p;
  p.peer = h->peer;
  if (!h->peer_enabled || h->peer_field != FECB200_FIELD_RESIDUAL) p.peer.n_owned = -1;
  p.state_old = b.d_state_old.p;
  p.ne = (int32_t)b.ne; p.nq = b.nq; p.nnz = h->nnz;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  const size_t smem = ((size_t)SH::NST * SH::EPB * L::ELSM + L::TAB) * sizeof(double);
  int sms = 0;
  FEC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const int64_t nbatch = (b.ne + SH::EPB - 1) / SH::EPB;
  const int grid = (int)(nbatch < sms ? nbatch : sms);
  p.zf = make_zero_fill(a, grid);   // only p / total16 are used: the share is per batch, not per CTA
  timing_begin(h);
  auto kern = k_mat2w<ND, NNPE, NF, NQT, Phys, SH, MAXREG, WITH_R>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, SH::THREADS, smem, h->stream>>>(p);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

template <int ND, int NNPE, int NF, int NQT, class Phys, class SH, int MAXREG>
void run_mat2w(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if (a.R) run_mat2w_t<ND, NNPE, NF, NQT, Phys, SH, MAXREG, true>(h, b, a);
  else run_mat2w_t<ND, NNPE, NF, NQT, Phys, SH, MAXREG, false>(h, b, a);
}

}  // namespace fec
