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
  double props[kMaxProps];
  Tables<ND, NNPE, NQT> tab;
};

template <int ND, int NNPE, int NQT, class Phys, int KIND, bool TRANS>
__global__ void __launch_bounds__(128) k_mat_scalar(const __grid_constant__ MatSParams<ND, NNPE, NQT> p) {
  static_assert(Phys::NF == 1 && Phys::NS == 0, "scalar, stateless physics");
  __shared__ __align__(16) double zero_page[kZeroPageBytes / 8];
  zero_fill_begin(p.zf, zero_page);
  const int e = blockIdx.x * 128 + threadIdx.x;
  const bool live = e < p.ne;   // the whole warp takes part in the staged scatter
  double x[NNPE][ND], u[NNPE][1];
#pragma unroll
  for (int a = 0; a < NNPE; ++a) {
    const int n = live ? p.conn[(size_t)e * NNPE + a] : 0;
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
  extern __shared__ __align__(16) unsigned char sm_raw[];
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
      for (int c = 0; c < NNPE; ++c) kst[lane * KST + r * NNPE + c] = TRANS ? K[r][c] : K[c][r];   // storage (row r, col c)
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
  const int grid = (int)((b.ne + 127) / 128);
  p.zf = make_zero_fill(a, grid);
  const bool trans = (h->opts.matrix_type == FECB200_CSC);
  constexpr int RECS = ((NNPE * 4 + NNPE * NNPE * 2 + 2 * NNPE + 15) / 16) * 16;
  FEC_REQUIRE((int)b.emeta_rec >= RECS && b.emeta_rec % 16 == 0, "scalar matrix kernel: scatter record size mismatch");
  const size_t smem = 128 * ((size_t)(NNPE * NNPE + 1) * sizeof(double) + RECS);
  auto launch = [&](auto kern) {
    FEC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, h->stream>>>(p);
  };
  timing_begin(h);
  if (a.kind == FECB200_MASS) launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_MASS, false>);
  else if (trans) launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, true>);
  else launch(k_mat_scalar<ND, NNPE, NQT, Phys, FECB200_STIFFNESS, false>);
  FEC_CUDA(cudaGetLastError());
  timing_end(h);
  h->launches++;
}

}  // namespace fec
