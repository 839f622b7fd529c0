// kernel_mat_scalar.cuh -- stiffness / mass -> CSR for scalar problems (NF = 1, e.g. Poisson, BASELINE config 2).
//
// One thread per element, everything in registers: no shared memory, no barriers.  The NNPE x NNPE element
// matrix is accumulated over the quadrature points and written straight into the CSR values through the
// per-element scatter record (row offsets u32, column offsets u16 indexed by LOCAL node -- see k_build_emeta with
// trash_rows = 0).  REDs are branch-free (eliminated rows / columns go to the hashed trash region behind nz).
// Honours the reference's transposed COO convention (K[dof_b, dof_a] += K_el[a, b]) through TRANS.
#pragma once
#include "kernels.cuh"

namespace fec {

template <int ND, int NNPE, int NQT>
struct MatSParams {
  const double* X;
  const double* U;
  double* nz;
  const int32_t* conn;
  const unsigned char* emeta;
  int64_t nnz;
  int32_t ne, nq, rec;
  ZeroFill zf;
  double wc[5], wr[4];           // Walsh form: c^n / 64 (n = 0..4), c^n / 8 (n = 0..3)
  int32_t nos[8], pos[8];        //             sign index -> local node / quadrature point
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};

// WALSH (HEX8 / 2-point rule per axis, stiffness only; same identities as kernel_mat2.cuh / DESIGN.md 3.2b): the element
// fields are gathered in sign order and analysed once, J_q and grad_xi u are 3 add/sub per entry, the pulled-back tangent
// B_q = JxW J^-1 A J^-T of every point is parked in the thread's column of shared memory (the scatter stage is idle), its 9
// entries are transformed over the points one at a time into the 7 x 7 spectrum, and one two-sided synthesis yields K_el:
// ~2000 instead of ~4000 FP64 instructions per element (ncu r02t: the quadrature loop keeps the FP64 pipe 58 % busy at 8
// warps per SM, the REDs alone would need 0.47 of the 0.75 ms at 128^3).
template <int ND, int NNPE, int NQT, class Phys, int KIND, bool TRANS, bool WALSH = false>
__global__ void __launch_bounds__(128) k_mat_scalar(const __grid_constant__ MatSParams<ND, NNPE, NQT> p) {
  static_assert(Phys::NF == 1 && Phys::NS == 0, "scalar, stateless physics");
  static_assert(!WALSH || (ND == 3 && NNPE == 8 && NQT == 8 && KIND == FECB200_STIFFNESS), "Walsh form: HEX8 stiffness");
  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];
  extern __shared__ __align__(16) unsigned char sm_raw[];
  zero_fill_begin(p.zf, zero_page);
  const int e = blockIdx.x * 128 + threadIdx.x;
  const bool live = e < p.ne;   // the whole warp takes part in the staged scatter
  double x[NNPE][ND], u[NNPE][1];
#pragma unroll
  for (int a = 0; a < NNPE; ++a) {
    const int n = live ? p.conn[(size_t)e * NNPE + (WALSH ? p.nos[a] : a)] : 0;
#pragma unroll
    for (int j = 0; j < ND; ++j) x[a][j] = p.X[(size_t)n * ND + j];
    u[a][0] = p.U[n];
  }
  double K[NNPE][NNPE];
#pragma unroll
  for (int a = 0; a < NNPE; ++a)
#pragma unroll
    for (int b = 0; b < NNPE; ++b) K[a][b] = 0.0;
  const int nq = (NQT > 0) ? NQT : p.nq;
  if constexpr (WALSH) {
    double* const st = reinterpret_cast<double*>(sm_raw) + threadIdx.x;   // this thread's stash column, stride 128
    {
      double Xh[8][3], Uh[8][1];
      vw_analyse<3>(x, p.wr, Xh);
      vw_analyse<1>(u, p.wr, Uh);
      auto point = [&](auto QC) {
        constexpr int Q = decltype(QC)::value;
        double Jt[3][3], J[3][3], Ji[3][3], gx[1][3], gu[1][3];
        vw_gradient<3, Q>(VwReg<3>{Xh}, Jt);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) J[i][k] = Jt[i][k];
        const double JxW = invert<3>(J, Ji) * p.tab.w[p.pos[Q]];
        vw_gradient<1, Q>(VwReg<1>{Uh}, gx);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j) s = fma(gx[0][j], Ji[j][k], s);
          gu[0][k] = s;
        }
        double A[3][3], T[3][3];
        Phys::tangent(gu, p.props, nullptr, A);
#pragma unroll
        for (int j1 = 0; j1 < 3; ++j1)
#pragma unroll
          for (int k2 = 0; k2 < 3; ++k2) {
            double s = A[j1][0] * Ji[k2][0];
#pragma unroll
            for (int j2 = 1; j2 < 3; ++j2) s = fma(A[j1][j2], Ji[k2][j2], s);
            T[j1][k2] = s * JxW;
          }
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
          for (int k2 = 0; k2 < 3; ++k2) {
            double s = Ji[k1][0] * T[0][k2];
#pragma unroll
            for (int j1 = 1; j1 < 3; ++j1) s = fma(Ji[k1][j1], T[j1][k2], s);
            st[(Q * 9 + k1 * 3 + k2) * 128] = s;   // B_q[k1][k2]
          }
      };
      point(std::integral_constant<int, 0>{}); point(std::integral_constant<int, 1>{});
      point(std::integral_constant<int, 2>{}); point(std::integral_constant<int, 3>{});
      point(std::integral_constant<int, 4>{}); point(std::integral_constant<int, 5>{});
      point(std::integral_constant<int, 6>{}); point(std::integral_constant<int, 7>{});
    }
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 3; ++k2) {
        double bq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) bq[q] = st[(q * 9 + k1 * 3 + k2) * 128];
        walsh_fwd8(bq);
#pragma unroll
        for (int s1 = 0; s1 < 8; ++s1)
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2)
            if (!(s1 & (1 << k1)) && !(s2 & (1 << k2))) {
              const int al = s1 | (1 << k1), be = s2 | (1 << k2);
              K[al][be] = fma(p.wc[popc3(s1) + popc3(s2)], bq[s1 ^ s2], K[al][be]);
            }
      }
#pragma unroll
    for (int al = 1; al < 8; ++al) walsh_syn8_z(K[al]);
#pragma unroll
    for (int ib = 0; ib < 8; ++ib) {
      double col[8];
#pragma unroll
      for (int al = 0; al < 8; ++al) col[al] = K[al][ib];
      walsh_syn8_z(col);
#pragma unroll
      for (int ia = 0; ia < 8; ++ia) K[ia][ib] = col[ia];
    }
    __syncthreads();   // every thread of the CTA is done with its stash column: the region becomes the scatter stage
  }
#pragma unroll 1
  for (int q = 0; q < (WALSH ? 0 : nq); ++q) {
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
    if constexpr (KIND == FECB200_MASS) {
      const double rho = JxW * Phys::density(p.props);
#pragma unroll
      for (int a = 0; a < NNPE; ++a)
#pragma unroll
        for (int b = 0; b < NNPE; ++b) K[a][b] = fma(rho * p.tab.N[q][a], p.tab.N[q][b], K[a][b]);
    } else {
      double g[NNPE][ND], gu[1][ND];
#pragma unroll
      for (int k = 0; k < ND; ++k) gu[0][k] = 0.0;
#pragma unroll
      for (int a = 0; a < NNPE; ++a)
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < ND; ++j) s = fma(p.tab.dN[q][a][j], Ji[j][k], s);
          g[a][k] = s;
          gu[0][k] = fma(u[a][0], s, gu[0][k]);
        }
      double A[ND][ND];
      Phys::tangent(gu, p.props, nullptr, A);
#pragma unroll
      for (int b = 0; b < NNPE; ++b) {
        double tb[ND];
#pragma unroll
        for (int j1 = 0; j1 < ND; ++j1) {
          double s = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < ND; ++j2) s = fma(A[j1][j2], g[b][j2], s);
          tb[j1] = s * JxW;
        }
#pragma unroll
        for (int a = 0; a < NNPE; ++a) {
          double s = K[a][b];
#pragma unroll
          for (int j1 = 0; j1 < ND; ++j1) s = fma(g[a][j1], tb[j1], s);
          K[a][b] = s;
        }
      }
    }
  }
  // ---- scatter: record = u32 rowstart[NNPE] | u16 ecol[NNPE][NNPE] (by local node) | u8 mask[NNPE] | u8 rank[NNPE]
  // One RED per lane straight from the registers puts every lane of an instruction into its own 32-byte sector (64
  // sectors per element; the L2 atomic units were the limiter, 131 G sectors/s at 128^3).  Instead every warp stages its 32
  // element matrices and records in shared memory and walks them element by element with lane = (row r of 4, column c):
  // the 8 columns of a row are 4 runs of two x-adjacent nodes, i.e. 16 instead of 32 sectors per RED instruction.
  constexpr int KST = NNPE * NNPE + 1;                       // odd stride: conflict-free staging stores
  constexpr int RECB = NNPE * 4 + NNPE * NNPE * 2 + 2 * NNPE;  // bytes of a record (rowstart | ecol | mask | rank)
  constexpr int RECS = ((RECB + 15) / 16) * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* kst = reinterpret_cast<double*>(sm_raw) + (size_t)warp * 32 * KST;
  unsigned char* rst = sm_raw + (size_t)(blockDim.x >> 5) * 32 * KST * sizeof(double) + (size_t)warp * 32 * RECS;
  if (live) {
#pragma unroll
    for (int r = 0; r < NNPE; ++r)
#pragma unroll
      for (int c = 0; c < NNPE; ++c) {   // storage (row r, col c); the Walsh form holds K in sign order
        if constexpr (WALSH) kst[lane * KST + p.nos[r] * NNPE + p.nos[c]] = TRANS ? K[r][c] : K[c][r];
        else kst[lane * KST + r * NNPE + c] = TRANS ? K[r][c] : K[c][r];
      }
    const uint4* src = reinterpret_cast<const uint4*>(p.emeta + (size_t)e * p.rec);
    uint4* dst = reinterpret_cast<uint4*>(rst + (size_t)lane * RECS);
#pragma unroll
    for (int i = 0; i < RECS / 16; ++i) dst[i] = src[i];
  }
  __syncwarp();
  {
    const int e0 = blockIdx.x * 128 + warp * 32;
    const int nel = (p.ne - e0) < 32 ? (p.ne - e0) : 32;
    static_assert(NNPE == 8 || NNPE == 4, "lane = (row of 32 / NNPE, column)");
    constexpr int RPI = 32 / NNPE;                           // rows per RED instruction
    const int c = lane % NNPE, rl = lane / NNPE;
    for (int el = 0; el < nel; ++el) {
      const unsigned char* rec = rst + (size_t)el * RECS;
      const uint32_t* rs = reinterpret_cast<const uint32_t*>(rec);
      const uint16_t* ec = reinterpret_cast<const uint16_t*>(rec + NNPE * 4);
      const bool colok = rec[NNPE * 4 + NNPE * NNPE * 2 + c] & 1u;
      const uint32_t trash = (uint32_t)p.nnz + (((uint32_t)(e0 + el) * 613u) & 4095u);
#pragma unroll
      for (int it = 0; it < NNPE / RPI; ++it) {
        const int r = it * RPI + rl;
        const uint32_t rsr = rs[r];
        const bool ok = rsr != 0xFFFFFFFFu && colok;
        const uint32_t idx = ok ? rsr + ec[r * NNPE + c] : (uint32_t)p.nnz + ((trash + r * 8u + c) & 4095u);
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p.nz + idx), "d"(kst[el * KST + r * NNPE + c]));
      }
    }
  }
  zero_fill_end(p.zf);
}

template <int ND, int NNPE, int NQT, class Phys>
void run_mat_scalar(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  auto pp = std::make_unique<MatSParams<ND, NNPE, NQT>>();
  auto& p = *pp;
  FEC_REQUIRE(h->nnz + 4096 < (int64_t)0xFFFFFFFFll && b.d_emeta.p && !b.emeta_trash_rows, "scalar matrix kernel: bad scatter records");
  p.X = h->d_X.p; p.U = a.U; p.nz = a.nz; p.conn = b.d_conn_perm.p; p.emeta = b.d_emeta.p; p.nnz = h->nnz;
  p.ne = (int32_t)b.ne; p.nq = b.nq; p.rec = (int32_t)b.emeta_rec;
  for (int i = 0; i < kMaxProps; ++i) p.props[i] = i < (int)b.props.size() ? b.props[i] : 0.0;
  fill_tables<ND, NNPE, NQT>(b, p.tab);
  for (int n = 0; n < 5; ++n) p.wc[n] = std::pow(b.walsh_c, n) / 64.0;
  for (int n = 0; n < 4; ++n) p.wr[n] = std::pow(b.walsh_c, n) / 8.0;
  for (int i = 0; i < 8; ++i) { p.nos[i] = b.walsh ? b.node_of_sign[i] : i; p.pos[i] = b.walsh ? b.point_of_sign[i] : i; }
  const int grid = (int)((b.ne + 127) / 128);
  p.zf = make_zero_fill(a, grid);
  const bool trans = (h->opts.matrix_type == FECB200_CSC);
  constexpr int RECS = ((NNPE * 4 + NNPE * NNPE * 2 + 2 * NNPE + 15) / 16) * 16;
  FEC_REQUIRE((int)b.emeta_rec >= RECS && b.emeta_rec % 16 == 0, "scalar matrix kernel: scatter record size mismatch");
  size_t smem = 128 * ((size_t)(NNPE * NNPE + 1) * sizeof(double) + RECS);
  bool walsh = false;
  if constexpr (ND == 3 && NNPE == 8 && NQT == 8) walsh = b.walsh && a.kind == FECB200_STIFFNESS && !getenv("FECB200_MAT2_CLASSIC");
  if (walsh) smem = std::max(smem, (size_t)128 * 72 * sizeof(double));   // the B_q stash
  auto launch = [&](auto kern) {
    FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, h->stream>>>(p);
  };
  timing_begin(h);
  if constexpr (ND == 3 && NNPE == 8 && NQT == 8) {
    if (walsh) {
      if (trans) launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, true, true>);
      else launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, false, true>);
      FEC_CUDA(cudaGetLastError());
      timing_end(h);
      h->launches++;
      return;
    }
  }
  if (a.kind == FECB200_MASS) launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_MASS, false>);
  else if (trans) launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, true>);
  else launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, false>);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

}  // namespace fec
