// kernel_mat3.cuh -- warp-specialised, persistent tangent kernel for NF = ND mechanics on hex8 (round 2).
//
// Why (profiles/r01v_fused_sass_profile.txt): in k_mat2 every warp walks three phases one after
// the other -- G (gathers, geometry, constitutive tangent: a long dependent chain that issues at a third of the FP64
// rate and runs its 8 quadrature points on 6 threads in two rounds, 62 % lane efficiency), K (register accumulation,
// saturates the FP64 pipe of its scheduler on its own) and S (staging + REDs, no FP64 at all) -- with two warps per
// scheduler at 254 registers.  The FP64 pipe is busy only while one of the two is in K: 58 %.
//
// Here one persistent CTA per SM runs 4 PRODUCER warps and 4 CONSUMER warps (one of each per scheduler) around a ring
// of 40 element slots in shared memory:
//   producers  thread = (element, quadrature point): 4 elements x 8 points per pass, every lane busy, one round.
//              The 8 threads of an element fetch one node each (connectivity read coalesced), exchange X / U through
//              the element's slot, compute dN_X, JxW * A (packed symmetric) [and JxW * P for the fused residual]
//              and publish them; the scatter record of the element arrives with cp.async.  Reference tables are
//              read from a q-fastest shared-memory copy (one wavefront per load, no per-lane-indexed LDC).
//   consumers  5 elements x 6 component pairs per batch, exactly k_mat2's phases K, S1, S2 -- but they never gather,
//              never wait for a constitutive chain, and their K phases are fed back to back.
// Hand-over: mbarriers in shared memory, one "full" barrier per consumer batch slot (5 arrivals, one per element)
// and one "empty" barrier per producer batch slot (4 arrivals).  Every slot use has exactly one producer warp and
// one consumer warp, and a slot is refilled only after its consumer released it, so plain phase parities are safe.
// A bounded spin turns a protocol error into a trap instead of a hang.
//
// MEASURED (round 2, B200, 192^3, profiles/r02_kmat3_prof.txt): correct (all parity tests), but 22.2 ms against 16.0 ms for
// k_mat2 (31.7 ms with the in-kernel zero-fill, whose bulk stores stall the issuing producer lane).  Per consumer warp
// phase K takes 9.9 k cycles per 5-element batch -- the same as in k_mat2: ONE warp in K reaches only ~55 % of its
// scheduler's FP64 pipe (3 dependent DFMAs per accumulator, ptxas has no registers left to interleave chains), and with
// one 255-register consumer per scheduler nobody fills the gaps; in k_mat2 the second general-purpose warp does.  The
// register file (64 K) holds 8 such warps and no more, so specialisation cannot add K-capable warps.  Kept here
// (compiled only with -DFEC_MAT3=1, see tools/build_variants.sh) as the record of that experiment.
//
// The CTA also clears its share of the idle CSR value buffer (TMA bulk stores from a zero page), a few pages per
// producer pass, so the stores drain under the whole kernel instead of in one burst.
#pragma once
#include "../../finiteelementcontainers.jl_b200/csrc/kernel_mat2.cuh"

namespace fec {

namespace mat3 {
constexpr int kProducers = 4, kConsumers = 4, kWarps = kProducers + kConsumers;
constexpr int kRing = 40;             // element slots: lcm(4 producer elements, 5 consumer elements) x 2
constexpr int kPB = 4, kCB = 5;       // elements per producer pass / per consumer batch
constexpr int kNPB = kRing / kPB, kNCB = kRing / kCB;

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// returns the cycles spent waiting (only meaningful with -DFEC_MAT3_PROF; the compiler drops it otherwise)
__device__ __forceinline__ long long mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 20000000000ll) __trap();   // ~10 s: a protocol error must not hang the device
  }
  return clock64() - t0;
}
#ifdef FEC_MAT3_PROF
__device__ unsigned long long g_prof[8];   // [0] producer wait, [1] producer total, [2] consumer wait, [3] consumer total, [4] K, [5] S
#define FEC_PROF(...) __VA_ARGS__
#else
#define FEC_PROF(...)
#endif
}  // namespace mat3

template <int ND, int NNPE, int NQT>
struct Mat3Params {
  Mat2Params<ND, NNPE, NQT> m;   // same fields as k_mat2 (tables are copied to shared memory at start)
  int32_t groups_per_cta;        // groups of 20 elements per CTA (contiguous chunk of the tile-ordered elements)
  int32_t zf_pages_per_pass;     // zero pages each producer warp queues per pass
};

template <int ND, int NNPE, int NF, int NQT, class Phys, bool WITH_R>
__global__ void __launch_bounds__(mat3::kWarps * 32, 1) k_mat3(const __grid_constant__ Mat3Params<ND, NNPE, NQT> pp) {
  using namespace mat3;
  static_assert(ND == 3 && NNPE == 8 && NF == 3 && NQT == 8, "k_mat3 is written for hex8, NF = 3, 8-point rules");
  using L = Mat2Layout<ND, NNPE, NF, NQT, WITH_R>;
  constexpr int NP = L::NP, NDF = L::NDF, SLOT = L::SLOT, NROW = L::NROW, RS = L::RSTRIDE;
  constexpr int NS = Phys::NS;
  const Mat2Params<ND, NNPE, NQT>& p = pp.m;
  extern __shared__ __align__(16) double smem[];
  double* ring = smem;                                            // kRing * ELSM
  double* tabs = ring + (size_t)kRing * L::ELSM;                  // dN q-fastest [(a*ND + j)*NQT + q], then w[q]
  double* zero_page = tabs + (NNPE * ND + 1) * NQT;               // kZeroPageBytes
  uint64_t* full_cb = reinterpret_cast<uint64_t*>(zero_page + kZeroPageBytes / 8);
  uint64_t* empty_pb = full_cb + kNCB;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- CTA set-up
  for (int i = threadIdx.x; i < NNPE * ND * NQT; i += blockDim.x) {
    const int q = i % NQT, aj = i / NQT;
    tabs[i] = p.tab.dN[q][aj / ND][aj % ND];
  }
  if (threadIdx.x < NQT) tabs[NNPE * ND * NQT + threadIdx.x] = p.tab.w[threadIdx.x];
  for (int i = threadIdx.x; i < kZeroPageBytes / 8; i += blockDim.x) zero_page[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNCB; ++i) mbar_init(&full_cb[i], kCB);
    for (int i = 0; i < kNPB; ++i) mbar_init(&empty_pb[i], kPB);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  const int64_t e_begin = (int64_t)blockIdx.x * pp.groups_per_cta * 20;
  int64_t n_el = (int64_t)p.ne - e_begin;
  if (n_el > (int64_t)pp.groups_per_cta * 20) n_el = (int64_t)pp.groups_per_cta * 20;
  if (n_el < 0) n_el = 0;
  const int n_cb = (int)((n_el + kCB - 1) / kCB);
  const int n_pb = (n_cb * kCB + kPB - 1) / kPB;      // producers cover every slot a consumer batch waits for

  if (warp < kProducers) {
    // =========================== producer: thread = (element, quadrature point) ===========================
    const int elw = lane >> 3, q = lane & 7;
    // zero-fill share of this CTA (16-byte units), dealt to the 4 producer lanes 0 in pages
    int64_t zf_pos = 0, zf_end = 0;
    if (p.zf.p != nullptr && lane == 0) {
      const int64_t beg = (int64_t)blockIdx.x * p.zf.chunk16;
      const int64_t end = beg + p.zf.chunk16 < p.zf.total16 ? beg + p.zf.chunk16 : p.zf.total16;
      const int64_t share = (end - beg + kProducers - 1) / kProducers;
      zf_pos = beg + warp * share;
      zf_end = zf_pos + share < end ? zf_pos + share : end;
    }
    FEC_PROF(long long t_wait = 0; const long long t_start = clock64();)
    for (int pb = warp; pb < n_pb; pb += kProducers) {
      const int sp = pb % kNPB, use = pb / kNPB;
      if (use > 0) { FEC_PROF(t_wait +=) mbar_wait(&empty_pb[sp], (use - 1) & 1); }
      const int64_t i = (int64_t)pb * kPB + elw;        // element index inside the CTA's chunk
      const int64_t e = e_begin + i;
      const bool active = i < n_el;
      double* esm = ring + (size_t)(i % kRing) * L::ELSM;
      // scatter records of the pass (contiguous in global memory): LDGSTS under the compute below
      {
        const int64_t i0 = (int64_t)pb * kPB;
        const int nel = (int)((n_el - i0) < kPB ? (n_el - i0 > 0 ? n_el - i0 : 0) : kPB);
        const unsigned char* g = p.emeta + (size_t)(e_begin + i0) * L::REC;
        constexpr int CH = L::REC / 16;
        for (int c = lane; c < nel * CH; c += 32) {
          const int el = c / CH, r = c - el * CH;
          cp_async16(reinterpret_cast<unsigned char*>(ring + (size_t)((i0 + el) % kRing) * L::ELSM + L::BODY16) + r * 16, g + (size_t)c * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      // zero pages of this pass (lane 0 only)
      if (lane == 0 && zf_pos < zf_end) {
        const unsigned zs = (unsigned)__cvta_generic_to_shared(zero_page);
        for (int k = 0; k < pp.zf_pages_per_pass && zf_pos < zf_end; ++k) {
          int64_t rem16 = zf_end - zf_pos;
          const unsigned nb = rem16 * 16 < kZeroPageBytes ? (unsigned)(rem16 * 16) : (unsigned)kZeroPageBytes;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char*>(p.zf.p) + zf_pos * 16), "r"(zs), "r"(nb) : "memory");
          zf_pos += nb / 16;
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      // gather: thread q of the element fetches node q; X / U travel through the slot (scratch behind slot 7's dN_X)
      double* scr = esm + (size_t)(NQT - 1) * SLOT + NNPE * ND;   // 48 doubles: x[8][3], u[8][3]
      if (active) {
        const int n = p.conn[(size_t)e * NNPE + q];
#pragma unroll
        for (int j = 0; j < ND; ++j) scr[q * ND + j] = p.X[(size_t)n * ND + j];
#pragma unroll
        for (int d = 0; d < NF; ++d) scr[NNPE * ND + q * NF + d] = p.U[(size_t)n * NF + d];
      }
      __syncwarp();
      double x[NNPE][ND], u[NNPE][NF];
      if (active) {
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
#pragma unroll
          for (int j = 0; j < ND; ++j) x[a][j] = scr[a * ND + j];
#pragma unroll
          for (int d = 0; d < NF; ++d) u[a][d] = scr[NNPE * ND + a * NF + d];
        }
      }
      __syncwarp();   // scratch is dead: thread 7 may overwrite it with its tangent
      if (active) {
        double J[ND][ND];
#pragma unroll
        for (int i2 = 0; i2 < ND; ++i2)
#pragma unroll
          for (int j = 0; j < ND; ++j) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < NNPE; ++a) s = fma(x[a][i2], tabs[(a * ND + j) * NQT + q], s);
            J[i2][j] = s;
          }
        double Ji[ND][ND];
        const double JxW = invert<ND>(J, Ji) * tabs[NNPE * ND * NQT + q];
        double* slot = esm + (size_t)q * SLOT;
        double gu[NF][ND];
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
            for (int j = 0; j < ND; ++j) s = fma(tabs[(a * ND + j) * NQT + q], Ji[j][k], s);
            slot[a * ND + k] = s;
#pragma unroll
            for (int d = 0; d < NF; ++d) gu[d][k] = fma(u[a][d], s, gu[d][k]);
          }
        }
        double so[NS > 0 ? NS : 1];
        if constexpr (NS > 0) {
#pragma unroll
          for (int s = 0; s < NS; ++s) so[s] = p.state_old[((size_t)s * p.nq + q) * p.ne + e];
        }
        double A[NDF][NDF];
        Phys::tangent_scaled(gu, p.props, so, JxW, A);
#pragma unroll
        for (int i2 = 0; i2 < NDF; ++i2)
#pragma unroll
          for (int j = i2; j < NDF; ++j) slot[NNPE * ND + sym_index<NDF>(i2, j)] = A[i2][j];
        if constexpr (WITH_R) {
          double P[NF][ND], bsrc[NF], sn[NS > 0 ? NS : 1];
          Phys::flux(gu, 0.0, p.props, so, NS > 0 ? sn : nullptr, P, bsrc);
#pragma unroll
          for (int d = 0; d < NF; ++d)
#pragma unroll
            for (int k = 0; k < ND; ++k) slot[L::OFF_P + d * ND + k] = P[d][k] * JxW;
          if constexpr (NS > 0) {
#pragma unroll
            for (int s = 0; s < NS; ++s) p.state_new[((size_t)s * p.nq + q) * p.ne + e] = sn[s];
          }
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kPB; ++k) {
          const int64_t ik = (int64_t)pb * kPB + k;
          if (ik < (int64_t)n_cb * kCB) mbar_arrive(&full_cb[(ik / kCB) % kNCB]);
        }
      }
    }
    FEC_PROF(if (lane == 0) { atomicAdd(&g_prof[0], (unsigned long long)t_wait); atomicAdd(&g_prof[1], (unsigned long long)(clock64() - t_start)); })
    if (lane == 0 && p.zf.p != nullptr) {
      // whatever is left of the share (short chunks), then keep the zero page alive until the stores have read it
      const unsigned zs = (unsigned)__cvta_generic_to_shared(zero_page);
      while (zf_pos < zf_end) {
        int64_t rem16 = zf_end - zf_pos;
        const unsigned nb = rem16 * 16 < kZeroPageBytes ? (unsigned)(rem16 * 16) : (unsigned)kZeroPageBytes;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char*>(p.zf.p) + zf_pos * 16), "r"(zs), "r"(nb) : "memory");
        zf_pos += nb / 16;
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    // =========================== consumer: k_mat2's phases K, S1, S2 on 5 elements x 6 pairs ===========================
    const int cw = warp - kProducers;
    const int elw = lane / NP, t = lane % NP;
    const bool lane_valid = elw < kCB;
    int d1 = 0, d2 = 0;
    {
      int k = t;
#pragma unroll
      for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = i; j < NF; ++j) { if (k == 0) { d1 = i; d2 = j; } --k; }
    }
    int aidx[ND][ND];
#pragma unroll
    for (int j1 = 0; j1 < ND; ++j1)
#pragma unroll
      for (int j2 = 0; j2 < ND; ++j2) {
        const int i = d1 * ND + j1, j = d2 * ND + j2;
        aidx[j1][j2] = NNPE * ND + (i <= j ? i * NDF - (i * (i - 1)) / 2 + (j - i) : j * NDF - (j * (j - 1)) / 2 + (i - j));
      }
    FEC_PROF(long long t_wait = 0, t_k = 0, t_s = 0; const long long t_start = clock64();)
#pragma unroll 1
    for (int cb = cw; cb < n_cb; cb += kConsumers) {
      const int sc = cb % kNCB, use = cb / kNCB;
      FEC_PROF(t_wait +=) mbar_wait(&full_cb[sc], use & 1);
      FEC_PROF(const long long tk0 = clock64();)
      const int64_t i0 = (int64_t)cb * kCB;
      const int nel = (int)((n_el - i0) < kCB ? (n_el - i0) : kCB);
      double* wsm = ring + (size_t)(i0 % kRing) * L::ELSM;
      double* esm = wsm + (size_t)(lane_valid ? elw : 0) * L::ELSM;
      const bool active = lane_valid && elw < nel;
      // ---- phase K
      double M[NNPE][NNPE];
#pragma unroll
      for (int a = 0; a < NNPE; ++a)
#pragma unroll
        for (int b = 0; b < NNPE; ++b) M[a][b] = 0.0;
      double rr[WITH_R ? NNPE : 1];
#pragma unroll
      for (int a = 0; a < (WITH_R ? NNPE : 1); ++a) rr[a] = 0.0;
      if (active) {
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
            if (d1 == d2) {
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
            double tb[ND];
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
      __syncwarp();   // every thread of the warp is done reading the slots (re-used as the K_el stage)
      FEC_PROF(const long long ts0 = clock64(); t_k += ts0 - tk0;)
      // ---- phase S1: stage K_el (row = dof of the row node, column = (local node, dof)); see k_mat2
      if (active) {
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
#pragma unroll
          for (int b = 0; b < NNPE; ++b) {
            esm[(a * NF + d1) * RS + b * NF + d2] = M[a][b];
            if (d1 != d2) esm[(b * NF + d2) * RS + a * NF + d1] = M[a][b];
          }
        }
        if constexpr (WITH_R) {
          if (d1 == d2) {
#pragma unroll
            for (int a = 0; a < NNPE; ++a) esm[L::R_OFF + a * NF + d1] = rr[a];
          }
        }
      }
      __syncwarp();
      // ---- phase S2: REDs, lane = storage column (local node k, dof dc), the warp walks the rows
      if (lane < NROW) {
        const int k = lane / NF, dc = lane - k * NF;
        for (int el = 0; el < nel; ++el) {
          const double* ks = wsm + (size_t)el * L::ELSM;
          const unsigned char* rec = reinterpret_cast<const unsigned char*>(ks + L::BODY16);
          const uint16_t* ec = reinterpret_cast<const uint16_t*>(rec + L::OFF_EC);
          const unsigned mask = rec[L::OFF_MK + k];
          if (mask & (1u << dc)) {
            const int rank = __popc(mask & ((1u << dc) - 1u));
            uint32_t r0[NROW];
            double val[NROW];
            uint32_t off[NNPE];
#pragma unroll
            for (int b = 0; b < NNPE; ++b) off[b] = ec[b * NNPE + k] + rank;
            const uint4* rs4 = reinterpret_cast<const uint4*>(rec);
#pragma unroll
            for (int i = 0; i < NROW / 4; ++i) {
              const uint4 v = rs4[i];
              r0[4 * i] = v.x; r0[4 * i + 1] = v.y; r0[4 * i + 2] = v.z; r0[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int row = 0; row < NROW; ++row) val[row] = ks[row * RS + lane];
#pragma unroll
            for (int row = 0; row < NROW; ++row)
              asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p.nz + (r0[row] + off[row / NF])), "d"(val[row]));
          }
          if constexpr (WITH_R) {
            const uint32_t n = reinterpret_cast<const uint32_t*>(rec + L::OFF_ND)[k];
            scatter_add(p.peer, p.R, (int64_t)n, NF, dc, ks[L::R_OFF + lane]);
          }
        }
      }
      __syncwarp();   // all reads of the batch's slots are done: hand them back
      FEC_PROF(t_s += clock64() - ts0;)
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kCB; ++k) mbar_arrive(&empty_pb[((i0 + k) / kPB) % kNPB]);
      }
    }
    FEC_PROF(if (lane == 0) { atomicAdd(&g_prof[2], (unsigned long long)t_wait); atomicAdd(&g_prof[3], (unsigned long long)(clock64() - t_start)); atomicAdd(&g_prof[4], (unsigned long long)t_k); atomicAdd(&g_prof[5], (unsigned long long)t_s); })
  }
}

template <int ND, int NNPE, int NF, int NQT, class Phys, bool WITH_R>
void run_mat3_t(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  using L = Mat2Layout<ND, NNPE, NF, NQT, WITH_R>;
  auto ppp = std::make_unique<Mat3Params<ND, NNPE, NQT>>();
  auto& pp = *ppp;
  auto& p = pp.m;
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz;
  FEC_REQUIRE((int64_t)nz_alloc_len(h) < (int64_t)0xFFFFFFFFll, "k_mat3 needs nnz < 2^32 (32-bit row offsets in the scatter records)");
  FEC_REQUIRE((int)b.emeta_rec == L::REC, "scatter record size mismatch");
  p.conn = b.d_conn_perm.p; p.emeta = b.d_emeta.p;
  p.R = a.R; p.state_new = b.d_state_new.p;
  p.peer = h->peer;
  if (!h->peer_enabled || h->peer_field != FECB200_FIELD_RESIDUAL) p.peer.n_owned = -1;
  p.state_old = b.d_state_old.p;
  p.ne = (int32_t)b.ne; p.nq = b.nq; p.nnz = h->nnz;
  p.ko = 0;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  int sms = 0;
  FEC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const int64_t groups = (b.ne + 19) / 20;
  const int grid = (int)std::min<int64_t>(sms, groups);
  pp.groups_per_cta = (int32_t)((groups + grid - 1) / grid);
  p.zf = make_zero_fill(a, grid);
  {
    const int64_t passes = ((int64_t)pp.groups_per_cta * 20 / mat3::kPB + mat3::kProducers - 1) / mat3::kProducers;   // per producer warp
    const int64_t pages = ((int64_t)p.zf.chunk16 * 16 / mat3::kProducers + kZeroPageBytes - 1) / kZeroPageBytes;      // per producer warp
    pp.zf_pages_per_pass = (int32_t)std::max<int64_t>(1, (pages + passes - 1) / std::max<int64_t>(1, passes));
  }
  const size_t smem = ((size_t)mat3::kRing * L::ELSM + (NNPE * ND + 1) * NQT) * sizeof(double) + kZeroPageBytes +
                      (mat3::kNCB + mat3::kNPB) * sizeof(uint64_t);
  timing_begin(h);
  auto kern = k_mat3<ND, NNPE, NF, NQT, Phys, WITH_R>;
  FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, mat3::kWarps * 32, smem, h->stream>>>(pp);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
#ifdef FEC_MAT3_PROF
  {
    unsigned long long v[8], z[8] = {0};
    FEC_CUDA(cudaStreamSynchronize(h->stream));
    FEC_CUDA(cudaMemcpyFromSymbol(v, mat3::g_prof, sizeof v));
    FEC_CUDA(cudaMemcpyToSymbol(mat3::g_prof, z, sizeof z));
    const double np = (double)grid * mat3::kProducers, nc = (double)grid * mat3::kConsumers;
    fprintf(stderr, "[k_mat3 prof] per warp, Mcycles: producer wait %.2f of %.2f | consumer wait %.2f of %.2f (K %.2f, S %.2f)\n",
            v[0] / np / 1e6, v[1] / np / 1e6, v[2] / nc / 1e6, v[3] / nc / 1e6, v[4] / nc / 1e6, v[5] / nc / 1e6);
  }
#endif
}

template <int ND, int NNPE, int NF, int NQT, class Phys>
void run_mat3(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  if (a.R) run_mat3_t<ND, NNPE, NF, NQT, Phys, true>(h, b, a);
  else run_mat3_t<ND, NNPE, NF, NQT, Phys, false>(h, b, a);
}

}  // namespace fec
