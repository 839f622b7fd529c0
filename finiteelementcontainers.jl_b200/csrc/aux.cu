// aux.cu -- the small kernels that bracket every hot-path call in the reference, plus the
// device-resident Krylov pieces.
//
//   _update_for_assembly!             src/Parameters.jl:404-425
//   update_field_dirichlet_bcs!       src/bcs/DirichletBCs.jl:411-418
//   update_field_unknowns!            src/DofManagers.jl:349-411
//   update_field_periodic_bcs!        src/bcs/PeriodicBCs.jl:253-262
//   extract_field_unknowns!           src/DofManagers.jl:203-213
//   _adjust_*_for_constraints!        src/assemblers/Utils.jl:53-167
#include "common.cuh"

namespace fec {

static inline int grid_for(int64_t n, int bs = 256) { return (int)((n + bs - 1) / bs); }

__global__ void k_fill_indexed(double* f, const int32_t* idx, double v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) f[idx[i]] = v;
}
__global__ void k_set_indexed(double* f, const int32_t* idx, const double* vals, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) f[idx[i]] = vals[i];
}
__global__ void k_scatter_unknowns(double* f, const int32_t* ud, const double* Uu, int64_t n, int condensed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { const int32_t g = ud[i]; f[g] = condensed ? Uu[g] : Uu[i]; }
}
__global__ void k_periodic(double* f, const int32_t* pa, const int32_t* pb, const double* vals, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) f[pb[i]] = f[pa[i]] + vals[i];
}
__global__ void k_gather_unknowns(const double* f, const int32_t* ud, double* out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = f[ud[i]];
}
__global__ void k_periodic_fold(double* f, const int32_t* pa, const int32_t* pb, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&f[pa[i]], f[pb[i]]);
}
__global__ void k_constrain_vec(double* out, const double* f, const double* c, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (1.0 - c[i]) * f[i];
}
__global__ void k_constrain_action(double* out, const double* f, const double* c, const double* v, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (1.0 - c[i]) * f[i] + c[i] * v[i];
}
__global__ void k_zero_indexed(double* f, const int32_t* idx, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) f[idx[i]] = 0.0;
}

// _update_for_assembly! (Parameters.jl:404-425) as ONE launch: thread i handles Dirichlet entry i,
// unknown i - n_bc or periodic pair i - n_bc - n_u.  The three index sets are disjoint (DofManagers.jl:261-284) and a
// periodic side-a dof is always an unknown (checked in build_dof_structures), so U[b] = Uu[unknown(a)] + val needs no
// ordering against the other two segments.
__global__ void k_update_field_fused(double* f, const double* Uu, const int32_t* bc_dofs, const double* bc_vals, int64_t n_bc,
                                     const int32_t* ud, int64_t n_u, int condensed, const int32_t* pa, const int32_t* pb,
                                     const double* pvals, const int32_t* d2u, int64_t n_per) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_bc) { f[bc_dofs[i]] = bc_vals[i]; return; }
  i -= n_bc;
  if (i < n_u) { const int32_t g = ud[i]; f[g] = condensed ? Uu[g] : Uu[i]; return; }
  i -= n_u;
  if (i < n_per) f[pb[i]] = Uu[d2u[pa[i]]] + pvals[i];
}

void k_update_field(fecb200_handle* h, double* field, const double* Uu, bool with_bcs) {
  const int64_t n_bc = with_bcs ? h->n_bc : 0, n_per = with_bcs ? h->n_per : 0;
  const int64_t n = n_bc + h->n_unknowns + n_per;
  if (!n) return;
  k_update_field_fused<<<grid_for(n), 256, 0, h->stream>>>(field, Uu, h->d_bc_dofs.p, h->d_bc_vals.p, n_bc, h->d_unknown_dofs.p,
                                                          h->n_unknowns, h->opts.condensed, h->d_per_a.p, h->d_per_b.p,
                                                          h->d_per_vals.p, h->d_d2u.p, n_per);
  h->launches++;
  FEC_CUDA(cudaGetLastError());
}

void k_extract_unknowns(fecb200_handle* h, const double* field, double* out) {
  if (!h->n_unknowns) return;
  k_gather_unknowns<<<grid_for(h->n_unknowns), 256, 0, h->stream>>>(field, h->d_unknown_dofs.p, out, h->n_unknowns);
  h->launches++;
  FEC_CUDA(cudaGetLastError());
}

// residual(asm)  (Assemblers.jl:347-371)
void k_residual_accessor(fecb200_handle* h, double* out) {
  if (h->opts.condensed) {
    k_constrain_vec<<<grid_for(h->ndof), 256, 0, h->stream>>>(h->d_R.p, h->d_R.p, h->d_constraint.p, h->ndof);
    h->launches++;
    FEC_CUDA(cudaMemcpyAsync(out, h->d_R.p, h->ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  } else {
    if (h->n_per) {
      k_periodic_fold<<<grid_for(h->n_per), 256, 0, h->stream>>>(h->d_R.p, h->d_per_a.p, h->d_per_b.p, h->n_per);
      h->launches++;
    }
    k_extract_unknowns(h, h->d_R.p, out);
  }
  FEC_CUDA(cudaGetLastError());
}

// hvp(asm, v)  (Assemblers.jl:310-324)
void k_hvp_accessor(fecb200_handle* h, const double* v, double* out) {
  if (h->opts.condensed) {
    k_constrain_action<<<grid_for(h->ndof), 256, 0, h->stream>>>(h->d_Av.p, h->d_Av.p, h->d_constraint.p, v, h->ndof);
    h->launches++;
    FEC_CUDA(cudaMemcpyAsync(out, h->d_Av.p, h->ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  } else {
    k_extract_unknowns(h, h->d_Av.p, out);
  }
  FEC_CUDA(cudaGetLastError());
}

void k_zero_bc_slots(fecb200_handle* h, double* field) {
  if (!h->n_bc) return;
  k_zero_indexed<<<grid_for(h->n_bc), 256, 0, h->stream>>>(field, h->d_bc_dofs.p, h->n_bc);
  h->launches++;
}

// ---- per-element scatter records (layout: Mat2Layout in kernel_mat2.cuh); one thread per element.
//   conn  = scatter connectivity (periodic side-b nodes folded into side a), gconn = the nodes the element gathers from
//   trash_rows: rows that are not stored get an offset into the hashed trash region behind the values (k_mat2 adds
//   rows without testing); otherwise the 0xFFFFFFFF sentinel the scalar kernel tests for
__global__ void k_build_emeta(const int32_t* conn, const int32_t* gconn, const uint8_t* epos, const int32_t* adjptr, const uint16_t* coloff,
                              const uint8_t* freemask, const int64_t* rowstart, unsigned char* emeta, int nnpe, int nf,
                              int rec, int64_t ne, int64_t nnz, int trash_rows, EmetaOrder ord) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  // record entry i describes local node ord.node[i]: identity, or the SIGN order of the Walsh form of k_mat2
  const int32_t* c = conn + e * nnpe;
  unsigned char* r = emeta + e * rec;
  uint32_t* rs = reinterpret_cast<uint32_t*>(r);
  uint16_t* ec = reinterpret_cast<uint16_t*>(r + nnpe * nf * 4);
  uint8_t* mk = r + nnpe * nf * 4 + nnpe * nnpe * 2;
  uint8_t* rk = mk + nnpe;   // reserved (identity): columns are indexed by local node
  uint32_t* nd = reinterpret_cast<uint32_t*>(rk + nnpe);
  for (int i = 0; i < nnpe; ++i) {
    const int a = ord.node[i];
    rk[i] = (uint8_t)a;
    mk[i] = freemask[c[a]];
    nd[i] = (uint32_t)gconn[e * nnpe + a];   // fused residual: the node the element really gathers from
  }
  for (int i = 0; i < nnpe; ++i) {
    const int b = ord.node[i];
    for (int d = 0; d < nf; ++d) {
      const int64_t v = rowstart[(int64_t)c[b] * nf + d];
      const uint32_t dead = trash_rows ? (uint32_t)nnz + (uint32_t)(((uint32_t)e * 613u + (uint32_t)(i * nf + d) * 97u) & 4095u) : 0xFFFFFFFFu;
      rs[i * nf + d] = v < 0 ? dead : (uint32_t)v;
    }
    const int base = adjptr[c[b]];
    for (int j = 0; j < nnpe; ++j) ec[i * nnpe + j] = coloff[base + epos[(e * nnpe + b) * nnpe + ord.node[j]]];
  }
}

void build_ecol(fecb200_handle* h) {
  PhaseTimer _pt("build_ecol");
  h->adjx_ok = false;   // the dof maps changed: the SpMV's packed column table is rebuilt on its next use
  if (!h->matrix_ready || (int64_t)nz_alloc_len(h) >= (int64_t)0xFFFFFFFFll) return;
  for (auto& b : h->blocks) {
    if (b.nnpe > 16) continue;
    const size_t rec = (((size_t)b.nnpe * h->nf * 4 + (size_t)b.nnpe * b.nnpe * 2 + 2 * b.nnpe + 4 * b.nnpe + 15) / 16) * 16;
    if (b.d_emeta.n != rec * b.ne) b.d_emeta.alloc(rec * b.ne);
    b.emeta_rec = rec;
    b.emeta_trash_rows = h->nf > 1;  // k_mat2 records (dead rows point into the trash region); else scalar-kernel records
    // the Walsh form of k_mat2 keeps K_el in sign order: its records are written in that order (free at run time)
    b.emeta_sign_order = b.walsh && b.elem_type == FECB200_HEX8 && h->nf == 3 && !getenv("FECB200_MAT2_CLASSIC");
    EmetaOrder ord;
    for (int i = 0; i < 16; ++i) ord.node[i] = (b.emeta_sign_order && i < 8) ? b.node_of_sign[i] : i;
    k_build_emeta<<<grid_for(b.ne), 256, 0, h->stream>>>(b.d_sconn_perm.p ? b.d_sconn_perm.p : b.d_conn_perm.p, b.d_conn_perm.p, b.d_epos.p, h->d_adjptr.p, h->d_coloff.p,
                                                         h->d_freemask.p, h->d_rowstart.p, b.d_emeta.p, b.nnpe, h->nf,
                                                         (int)rec, b.ne, h->nnz, b.emeta_trash_rows ? 1 : 0, ord);
    h->launches++;
  }
  FEC_CUDA(cudaGetLastError());
}

// ---- reductions ---------------------------------------------------------------------------
__global__ void k_dot(const double* a, const double* b, int64_t n, double* out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fma(a[i], b[i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}
__global__ void k_gather_sum(const double* a, const int64_t* idx, int64_t n, double* out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (idx[i] >= 0) s += a[idx[i]];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

static double read_scalar(fecb200_handle* h) {
  double v = 0.0;
  FEC_CUDA(cudaMemcpyAsync(&v, h->d_red.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  return v;
}

double dot(fecb200_handle* h, const double* a, const double* b, int64_t n) {
  if (!h->d_red.p) h->d_red.alloc(4);
  FEC_CUDA(cudaMemsetAsync(h->d_red.p, 0, sizeof(double), h->stream));
  const int grid = (int)std::min<int64_t>(148 * 8, (n + 255) / 256);
  if (n) { k_dot<<<grid, 256, 0, h->stream>>>(a, b, n, h->d_red.p); h->launches++; }
  if (comm_active(h)) comm_allreduce_sum(h, h->d_red.p, 1);   // partitioned handles pass their OWNED length (owned_len)
  return read_scalar(h);
}

__global__ void k_axpy(double alpha, const double* x, double* y, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(alpha, x[i], y[i]);
}
__global__ void k_xpay(const double* x, double beta, double* y, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(beta, y[i], x[i]);
}
void axpy(fecb200_handle* h, double alpha, const double* x, double* y, int64_t n) {
  if (n) { k_axpy<<<grid_for(n), 256, 0, h->stream>>>(alpha, x, y, n); h->launches++; }
}
void xpay(fecb200_handle* h, const double* x, double beta, double* y, int64_t n) {
  if (n) { k_xpay<<<grid_for(n), 256, 0, h->stream>>>(x, beta, y, n); h->launches++; }
}

// ---- condensed-mode matrix adjustment (assemblers/Utils.jl:53-148): penalty = 1e6 tr(K)/n;
// every stored entry of row (CSR) / column (CSC) i is scaled by (1 - c_i), diagonal += penalty c_i.
// With the structurally symmetric block storage both formats are the same sweep over the storage.
template <int NF>
__global__ void k_adjust_rows(double* nz, const int32_t* adjptr, const int32_t* adj, const uint8_t* freemask,
                              const uint16_t* coloff, const int64_t* rowstart, const int64_t* diagslot,
                              const double* c, double penalty, int64_t nn) {
  // one warp per node
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= nn) return;
  const int k0 = adjptr[n], k1 = adjptr[n + 1];
  int rowlen = 0;
  if (k1 > k0) rowlen = coloff[k1 - 1] + __popc(freemask[adj[k1 - 1]]);
#pragma unroll
  for (int d = 0; d < NF; ++d) {
    const int64_t rs = rowstart[n * NF + d];
    if (rs < 0) continue;
    const double ci = c[n * NF + d];
    if (ci == 0.0) continue;
    for (int j = lane; j < rowlen; j += 32) nz[rs + j] *= (1.0 - ci);
    __syncwarp();
    if (lane == 0) nz[diagslot[n * NF + d]] += penalty * ci;
  }
}

void k_adjust_matrix(fecb200_handle* h, double* nz) {
  if (!h->d_red.p) h->d_red.alloc(4);
  FEC_CUDA(cudaMemsetAsync(h->d_red.p, 0, sizeof(double), h->stream));
  k_gather_sum<<<(int)std::min<int64_t>(148 * 8, (h->ndof + 255) / 256), 256, 0, h->stream>>>(
      nz, h->d_diagslot.p, h->ndof, h->d_red.p);
  h->launches++;
  const double tr = read_scalar(h);
  const double penalty = 1.0e6 * tr / (double)h->nmat;
  const int64_t threads = h->nn * 32;
  const int grid = grid_for(threads);
#define ADJ(NF_)                                                                                         \
  k_adjust_rows<NF_><<<grid, 256, 0, h->stream>>>(nz, h->d_adjptr.p, h->d_adj.p, h->d_freemask.p,        \
                                                  h->d_coloff.p, h->d_rowstart.p, h->d_diagslot.p,      \
                                                  h->d_constraint.p, penalty, h->nn)
  switch (h->nf) {
    case 1: ADJ(1); break;
    case 2: ADJ(2); break;
    case 3: ADJ(3); break;
    default: throw Error("fecb200: NF > 3 not supported");
  }
#undef ADJ
  h->launches++;
  FEC_CUDA(cudaGetLastError());
}

// ---- SpMV on the block-compressed CSR: y = K x, one warp per node (NF rows) -------------------
// x, y are indexed like Uu (unknown ids when not condensed).  `d2u` = dof -> Uu index or -1.
template <int NF, bool TRANSPOSED>
__global__ void k_spmv(const double* __restrict__ nz, const double* __restrict__ x, double* __restrict__ y,
                       const int32_t* __restrict__ adjptr, const int32_t* __restrict__ adj,
                       const uint8_t* __restrict__ freemask, const uint16_t* __restrict__ coloff,
                       const int64_t* __restrict__ rowstart, const int32_t* __restrict__ d2u, int64_t nn) {
  const int64_t n = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= nn) return;
  const int k0 = adjptr[n], k1 = adjptr[n + 1];
  int64_t rs[NF];
  double acc[NF];
#pragma unroll
  for (int d = 0; d < NF; ++d) { rs[d] = rowstart[n * NF + d]; acc[d] = 0.0; }
  const int npairs = (k1 - k0) * NF;
  for (int i = lane; i < npairs; i += 32) {
    const int k = k0 + i / NF, d2 = i % NF;
    const int m = adj[k];
    const unsigned mask = freemask[m];
    if (!(mask & (1u << d2))) continue;
    const int pos = coloff[k] + __popc(mask & ((1u << d2) - 1u));
    const double xv = x[d2u[m * NF + d2]];
#pragma unroll
    for (int d = 0; d < NF; ++d)
      if (rs[d] >= 0) acc[d] = fma(nz[rs[d] + pos], xv, acc[d]);
  }
#pragma unroll
  for (int d = 0; d < NF; ++d) {
    double s = acc[d];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && rs[d] >= 0) y[d2u[n * NF + d]] = s;
  }
}

// SpMV, second version.  The first one chains four dependent loads per entry (adjptr -> adj -> freemask / d2u -> x) and
// walks a row in a run-time loop: 5.09 ms for the 13.8 GB of values at 192^3 (2.7 TB/s).  Here the column node's kept-dof
// mask and the Uu index of its first kept dof (the others follow: unknowns are numbered in dof order) are packed into one
// word per adjacency entry (d_adjx, rebuilt when the dof maps change), and the first three 32-lane steps of a row -- every
// row of a trilinear mesh -- are unrolled so that all their loads are in flight together.
__global__ void k_build_adjx(const int32_t* adj, const uint8_t* freemask, const int32_t* d2u, uint32_t* adjx, int nf, int64_t nadj) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nadj) return;
  const int m = adj[k];
  const unsigned mask = freemask[m];
  const int first = mask ? __ffs((int)mask) - 1 : 0;
  const int32_t ub = mask ? d2u[(int64_t)m * nf + first] : 0;
  adjx[k] = (mask << 29) | (uint32_t)(ub < 0 ? 0 : ub);
}
template <int NF, int NPW>
__global__ void __launch_bounds__(256) k_spmv2(const double* __restrict__ nz, const double* __restrict__ x, double* __restrict__ y,
                                               const int32_t* __restrict__ adjptr, const uint32_t* __restrict__ adjx,
                                               const uint16_t* __restrict__ coloff, const int64_t* __restrict__ rowstart,
                                               const int32_t* __restrict__ d2u, int64_t nn) {
  // NPW nodes per warp, walked together: every level of the dependent chain (adjptr -> adjx / coloff -> x, values) is
  // issued for all of them before the next one, which is what hides the DRAM latency
  const int64_t n0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5) * NPW;
  const int lane = threadIdx.x & 31;
  if (n0 >= nn) return;
  constexpr int UN = 3;
  int k0[NPW], npairs[NPW];
  int64_t rs[NPW][NF];
  double acc[NPW][NF];
#pragma unroll
  for (int u = 0; u < NPW; ++u) {
    const bool on = n0 + u < nn;
    const int64_t n = on ? n0 + u : n0;
    k0[u] = adjptr[n];
    npairs[u] = on ? (adjptr[n + 1] - k0[u]) * NF : 0;
#pragma unroll
    for (int d = 0; d < NF; ++d) { rs[u][d] = on ? rowstart[n * NF + d] : -1; acc[u][d] = 0.0; }
  }
  uint32_t ax[NPW][UN];
  int co[NPW][UN], dd[UN];
#pragma unroll
  for (int it = 0; it < UN; ++it) dd[it] = (lane + 32 * it) % NF;
#pragma unroll
  for (int u = 0; u < NPW; ++u)
#pragma unroll
    for (int it = 0; it < UN; ++it) {
      const int i = lane + 32 * it;
      const bool valid = i < npairs[u];
      const int k = k0[u] + (valid ? i / NF : 0);
      ax[u][it] = valid ? adjx[k] : 0u;
      co[u][it] = valid ? (int)coloff[k] : 0;
    }
  double xv[NPW][UN];
  int pos[NPW][UN];
#pragma unroll
  for (int u = 0; u < NPW; ++u)
#pragma unroll
    for (int it = 0; it < UN; ++it) {
      const unsigned mask = ax[u][it] >> 29;
      const bool keep = (mask >> dd[it]) & 1u;
      const int r = __popc(mask & ((1u << dd[it]) - 1u));
      pos[u][it] = keep ? co[u][it] + r : -1;
      xv[u][it] = keep ? __ldg(x + (ax[u][it] & 0x1FFFFFFFu) + r) : 0.0;
    }
  // all value loads are issued before the first FMA (predicated, independent of each other)
  double v[NPW][UN][NF];
#pragma unroll
  for (int u = 0; u < NPW; ++u)
#pragma unroll
    for (int it = 0; it < UN; ++it)
#pragma unroll
      for (int d = 0; d < NF; ++d) v[u][it][d] = (pos[u][it] >= 0 && rs[u][d] >= 0) ? __ldcs(nz + rs[u][d] + pos[u][it]) : 0.0;
#pragma unroll
  for (int u = 0; u < NPW; ++u)
#pragma unroll
    for (int it = 0; it < UN; ++it)
#pragma unroll
      for (int d = 0; d < NF; ++d) acc[u][d] = fma(v[u][it][d], xv[u][it], acc[u][d]);
#pragma unroll
  for (int u = 0; u < NPW; ++u) {
    for (int i = lane + 32 * UN; i < npairs[u]; i += 32) {   // longer rows (unstructured meshes, higher valence)
      const int k = k0[u] + i / NF, d2 = i % NF;
      const uint32_t a = adjx[k];
      const unsigned mask = a >> 29;
      if (!((mask >> d2) & 1u)) continue;
      const int r = __popc(mask & ((1u << d2) - 1u));
      const int ps = coloff[k] + r;
      const double xvv = x[(a & 0x1FFFFFFFu) + r];
#pragma unroll
      for (int d = 0; d < NF; ++d)
        if (rs[u][d] >= 0) acc[u][d] = fma(nz[rs[u][d] + ps], xvv, acc[u][d]);
    }
#pragma unroll
    for (int d = 0; d < NF; ++d) {
      double s = acc[u][d];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && rs[u][d] >= 0) y[d2u[(n0 + u) * NF + d]] = s;
    }
  }
}

// For CSC storage the same arrays hold K^T row-wise, i.e. y = K x needs the transposed product; the
// operators assembled here are symmetric in structure, and CG is only used for symmetric K, so
// the row-wise product is used for both (documented in DESIGN.md).
void spmv(fecb200_handle* h, const double* nz, const double* x, double* y) {
  const int64_t threads = h->nn * 32;
  const int grid = grid_for(threads);
  const int32_t* d2u = h->d_d2u.p;
  if (h->nf <= 3 && h->n_unknowns < (int64_t)0x1FFFFFFF && !getenv("FECB200_SPMV1")) {
    const int64_t nadj = (int64_t)h->d_adj.n;
    if (!h->adjx_ok) {
      if (h->d_adjx.n != (size_t)nadj) h->d_adjx.alloc((size_t)nadj);
      k_build_adjx<<<grid_for(nadj), 256, 0, h->stream>>>(h->d_adj.p, h->d_freemask.p, d2u, h->d_adjx.p, h->nf, nadj);
      h->adjx_ok = true;
      h->launches++;
    }
    constexpr int NPW = 1;   // 2 nodes per warp measured slower (4.06 vs 3.36 ms at 192^3: registers, occupancy)
    const int grid2 = grid_for((h->nn + NPW - 1) / NPW * 32);
#define SPMV2(NF_) k_spmv2<NF_, NPW><<<grid2, 256, 0, h->stream>>>(nz, x, y, h->d_adjptr.p, h->d_adjx.p, h->d_coloff.p, h->d_rowstart.p, d2u, h->nn)
    switch (h->nf) {
      case 1: SPMV2(1); break;
      case 2: SPMV2(2); break;
      default: SPMV2(3); break;
    }
#undef SPMV2
    h->launches++;
    FEC_CUDA(cudaGetLastError());
    return;
  }
#define SPMV(NF_)                                                                                       \
  k_spmv<NF_, false><<<grid, 256, 0, h->stream>>>(nz, x, y, h->d_adjptr.p, h->d_adj.p, h->d_freemask.p, \
                                                  h->d_coloff.p, h->d_rowstart.p, d2u, h->nn)
  switch (h->nf) {
    case 1: SPMV(1); break;
    case 2: SPMV(2); break;
    case 3: SPMV(3); break;
    default: throw Error("fecb200: NF > 3 not supported");
  }
#undef SPMV
  h->launches++;
  FEC_CUDA(cudaGetLastError());
}

// ---- state / source layout permutation: reference [NS,NQ,NE] <-> kernel [(s*NQ+q)*NE + e_tile] ----
__global__ void k_state_in(const double* src, double* dst, const int32_t* perm, int ns, int nq, int64_t ne) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ne * nq * ns) return;
  const int64_t e = i % ne;
  const int sq = (int)(i / ne);  // s*nq + q
  const int s = sq / nq, q = sq % nq;
  dst[i] = src[s + (int64_t)ns * (q + (int64_t)nq * perm[e])];
}
__global__ void k_state_out(const double* src, double* dst, const int32_t* perm, int ns, int nq, int64_t ne) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ne * nq * ns) return;
  const int64_t e = i % ne;
  const int sq = (int)(i / ne);
  const int s = sq / nq, q = sq % nq;
  dst[s + (int64_t)ns * (q + (int64_t)nq * perm[e])] = src[i];
}
void k_permute_state_in(fecb200_handle* h, BlockPlan& b, const double* src_dev, double* dst) {
  const int64_t n = b.ne * b.nq * b.nstate;
  if (n) { k_state_in<<<grid_for(n), 256, 0, h->stream>>>(src_dev, dst, b.d_perm.p, b.nstate, b.nq, b.ne); h->launches++; }
}
void k_permute_state_out(fecb200_handle* h, BlockPlan& b, const double* src, double* dst_dev) {
  const int64_t n = b.ne * b.nq * b.nstate;
  if (n) { k_state_out<<<grid_for(n), 256, 0, h->stream>>>(src, dst_dev, b.d_perm.p, b.nstate, b.nq, b.ne); h->launches++; }
}
void k_permute_scalar_out(fecb200_handle* h, BlockPlan& b, const double* src, double* dst_dev) {
  const int64_t n = b.ne * b.nq;
  if (n) { k_state_out<<<grid_for(n), 256, 0, h->stream>>>(src, dst_dev, b.d_perm.p, 1, b.nq, b.ne); h->launches++; }
}
void k_permute_source_in(fecb200_handle* h, BlockPlan& b, const double* src_dev, double* dst) {
  const int64_t n = b.ne * b.nq;
  if (n) { k_state_in<<<grid_for(n), 256, 0, h->stream>>>(src_dev, dst, b.d_perm.p, 1, b.nq, b.ne); h->launches++; }
}

// ---- halo pack / unpack-add (ghost -> owner accumulation, ext/PartitionedArraysExt.jl:469-481) ----
__global__ void k_pack(const double* f, const int32_t* nodes, double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n * nf) buf[i] = f[(int64_t)nodes[i / nf] * nf + i % nf];
}
__global__ void k_unpack_add(double* f, const int32_t* nodes, const double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n * nf) atomicAdd(&f[(int64_t)nodes[i / nf] * nf + i % nf], buf[i]);
}
void fill_indexed(fecb200_handle* h, double* field, const int32_t* idx, double v, int64_t n) {
  if (n) { k_fill_indexed<<<grid_for(n), 256, 0, h->stream>>>(field, idx, v, n); h->launches++; }
}
void halo_pack(fecb200_handle* h, const double* field, double* buf) {
  const int64_t n = (int64_t)h->d_send_nodes.n;
  if (n) { k_pack<<<grid_for(n * h->nf), 256, 0, h->stream>>>(field, h->d_send_nodes.p, buf, n, h->nf); h->launches++; }
}
void halo_unpack_add(fecb200_handle* h, double* field, const double* buf) {
  const int64_t n = (int64_t)h->d_recv_nodes.n;
  if (n) { k_unpack_add<<<grid_for(n * h->nf), 256, 0, h->stream>>>(field, h->d_recv_nodes.p, buf, n, h->nf); h->launches++; }
}

}  // namespace fec
