// loads.cu -- external-load vectors of the residual: the two calls solve! makes right after assemble_vector!
// (src/Solvers.jl:66-69, 133-137):
//   assemble_vector_source!      R[(n,d)] += - sum_q JxW N_n(q) b_d(q,e)     (src/assemblers/Source.jl:10-64)
//   assemble_vector_neumann_bc!  R[(n,d)] += + sum_q JxW_s N_n(q) g_d(q,e)   (src/assemblers/WeaklyEnforcedBCs.jl:4-15,61-83)
// with b / g pre-evaluated at the (cell / surface) quadrature points by the host, exactly like the reference's
// SourceContainer.vals / NeumannBCContainer.vals (Sources.jl:55-66, NeumannBCs.jl:60-71): the closures cannot cross
// the ABI.  Neither term depends on U, while the reference re-integrates both in every Newton iteration.  Here they
// are integrated once per value update into a cached nodal vector (run-time-shaped kernels, one thread per element /
// side) and each assemble call is one streaming add of that vector (HBM-bound, NDOF * 24 B).
#include "common.cuh"

namespace fec {

static inline int grid_for(int64_t n, int block = 128) { return (int)((n + block - 1) / block); }
constexpr int kMaxLoadNodes = 10;  // TET10 cells; QUAD4 / TRI6 faces

// one thread per element (tile order; vals are addressed through perm = tile order -> caller's order)
__global__ void __launch_bounds__(128) k_body_force(int64_t ne, int nnpe, int nq, int nd, int nf, const int32_t* conn,
                                                    const int32_t* perm, const double* tab, const double* vals,
                                                    const double* X, double* out, PeerScatter peer) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const double* N = tab;
  const double* dN = tab + (size_t)nq * nnpe;
  const double* w = dN + (size_t)nq * nnpe * nd;
  double x[kMaxLoadNodes][3], r[kMaxLoadNodes][3];
  int32_t node[kMaxLoadNodes];
  for (int a = 0; a < nnpe; ++a) {
    node[a] = conn[e * nnpe + a];
    for (int j = 0; j < 3; ++j) { x[a][j] = j < nd ? X[(size_t)node[a] * nd + j] : 0.0; r[a][j] = 0.0; }
  }
  const double* b = vals + (size_t)perm[e] * nq * nf;
  for (int q = 0; q < nq; ++q) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < nnpe; ++a)
      for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) J[i][j] = fma(x[a][i], dN[((size_t)q * nnpe + a) * nd + j], J[i][j]);
    const double det = nd == 2 ? J[0][0] * J[1][1] - J[0][1] * J[1][0]
                               : J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) + J[0][1] * (J[1][2] * J[2][0] - J[1][0] * J[2][2]) +
                                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    const double JxW = det * w[q];
    for (int a = 0; a < nnpe; ++a) {
      const double s = -JxW * N[(size_t)q * nnpe + a];   // scatter_with_values!(..., N, -JxW * b_val), Source.jl:60
      for (int d = 0; d < nf; ++d) r[a][d] = fma(s, b[(size_t)q * nf + d], r[a][d]);
    }
  }
  for (int a = 0; a < nnpe; ++a)
    for (int d = 0; d < nf; ++d) scatter_add(peer, out, (int64_t)node[a], nf, d, r[a][d]);
}

// one thread per side.  Surface map: t_k = sum_a x_a dNs[a][k]; JxW = |t_0| w (edges), |t_0 x t_1| w (faces)
__global__ void __launch_bounds__(128) k_surface_load(int64_t nsides, int nnps, int nqs, int nd, int nf, const int32_t* nodes,
                                                      const double* tab, const double* vals, const double* X, double* out,
                                                      PeerScatter peer) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nsides) return;
  const int ns = nd - 1;
  const double* N = tab;
  const double* dN = tab + (size_t)nqs * nnps;
  const double* w = dN + (size_t)nqs * nnps * ns;
  double x[kMaxLoadNodes][3], r[kMaxLoadNodes][3];
  int32_t node[kMaxLoadNodes];
  for (int a = 0; a < nnps; ++a) {
    node[a] = nodes[e * nnps + a];
    for (int j = 0; j < 3; ++j) { x[a][j] = j < nd ? X[(size_t)node[a] * nd + j] : 0.0; r[a][j] = 0.0; }
  }
  const double* g = vals + (size_t)e * nqs * nf;
  for (int q = 0; q < nqs; ++q) {
    double t[2][3] = {{0, 0, 0}, {0, 0, 0}};
    for (int a = 0; a < nnps; ++a)
      for (int k = 0; k < ns; ++k)
        for (int i = 0; i < nd; ++i) t[k][i] = fma(x[a][i], dN[((size_t)q * nnps + a) * ns + k], t[k][i]);
    double jac;
    if (nd == 2) {
      jac = sqrt(t[0][0] * t[0][0] + t[0][1] * t[0][1]);
    } else {
      const double c0 = t[0][1] * t[1][2] - t[0][2] * t[1][1], c1 = t[0][2] * t[1][0] - t[0][0] * t[1][2],
                   c2 = t[0][0] * t[1][1] - t[0][1] * t[1][0];
      jac = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    }
    const double JxW = jac * w[q];
    for (int a = 0; a < nnps; ++a) {
      const double s = JxW * N[(size_t)q * nnps + a];    // scatter_with_values!(..., Nvec, JxW * f_val), WeaklyEnforcedBCs.jl:80
      for (int d = 0; d < nf; ++d) r[a][d] = fma(s, g[(size_t)q * nf + d], r[a][d]);
    }
  }
  for (int a = 0; a < nnps; ++a)
    for (int d = 0; d < nf; ++d) scatter_add(peer, out, (int64_t)node[a], nf, d, r[a][d]);
}

// ---- Robin BCs (src/assemblers/WeaklyEnforcedBCs.jl:17-32, 85-180; src/bcs/RobinBCs.jl:72-86).  The flux law
// func(X_q, t, u_q) is a user closure in the reference (differentiated with ForwardDiff); it cannot cross the ABI, so
// the host hands over its affine form at the surface quadrature points: vals = g0 + D u_q, dvalsdu = D.
// One thread per side; JxW as in k_surface_load.
__device__ __forceinline__ double surface_jxw(const double (*x)[3], const double* dN, const double* w, int q, int nnps, int nd) {
  const int ns = nd - 1;
  double t[2][3] = {{0, 0, 0}, {0, 0, 0}};
  for (int a = 0; a < nnps; ++a)
    for (int k = 0; k < ns; ++k)
      for (int i = 0; i < nd; ++i) t[k][i] = fma(x[a][i], dN[((size_t)q * nnps + a) * ns + k], t[k][i]);
  double jac;
  if (nd == 2) {
    jac = sqrt(t[0][0] * t[0][0] + t[0][1] * t[0][1]);
  } else {
    const double c0 = t[0][1] * t[1][2] - t[0][2] * t[1][1], c1 = t[0][2] * t[1][0] - t[0][0] * t[1][2],
                 c2 = t[0][0] * t[1][1] - t[0][1] * t[1][0];
    jac = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
  }
  return jac * w[q];
}

// assemble_vector_robin_bc!: R[(n,d)] += sum_q JxW N_n (g0_d + sum_c D_dc u_c(q)),  u(q) = sum_a N_a U_a  (RobinBCs.jl:80-83)
__global__ void __launch_bounds__(128) k_robin_vector(int64_t nsides, int nnps, int nqs, int nd, int nf, const int32_t* nodes,
                                                      const double* tab, const double* g0, const double* D, const double* X,
                                                      const double* U, double* out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nsides) return;
  const double* N = tab;
  const double* dN = tab + (size_t)nqs * nnps;
  const double* w = dN + (size_t)nqs * nnps * (nd - 1);
  double x[kMaxLoadNodes][3], u[kMaxLoadNodes][3], r[kMaxLoadNodes][3];
  int32_t node[kMaxLoadNodes];
  for (int a = 0; a < nnps; ++a) {
    node[a] = nodes[e * nnps + a];
    for (int j = 0; j < 3; ++j) {
      x[a][j] = j < nd ? X[(size_t)node[a] * nd + j] : 0.0;
      u[a][j] = j < nf ? U[(size_t)node[a] * nf + j] : 0.0;
      r[a][j] = 0.0;
    }
  }
  for (int q = 0; q < nqs; ++q) {
    const double JxW = surface_jxw(x, dN, w, q, nnps, nd);
    double uq[3] = {0, 0, 0}, g[3];
    for (int a = 0; a < nnps; ++a)
      for (int c = 0; c < nf; ++c) uq[c] = fma(N[(size_t)q * nnps + a], u[a][c], uq[c]);
    const double* g0q = g0 + ((size_t)e * nqs + q) * nf;
    const double* Dq = D + ((size_t)e * nqs + q) * nf * nf;
    for (int d = 0; d < nf; ++d) {
      g[d] = g0q[d];
      for (int c = 0; c < nf; ++c) g[d] = fma(Dq[d + nf * c], uq[c], g[d]);
    }
    for (int a = 0; a < nnps; ++a) {
      const double s = JxW * N[(size_t)q * nnps + a];
      for (int d = 0; d < nf; ++d) r[a][d] = fma(s, g[d], r[a][d]);
    }
  }
  for (int a = 0; a < nnps; ++a)
    for (int d = 0; d < nf; ++d) atomicAdd(&out[(size_t)node[a] * nf + d], r[a][d]);
}

// assemble_matrix_robin_bc!: K_el[(i,di),(j,dj)] = sum_q JxW N_i N_j D[di,dj] over the side's nodes, added into the
// element's COO slots (_assemble_element_add!, :155-165), i.e. with the transposed labelling of the pattern
// (SparsityPatterns.jl:76-83): stored entry (row dof(j,dj), col dof(i,di)) += K_el[(i,di),(j,dj)].  Here: straight into
// the CSR / CSC values.  `csc` swaps the roles once more (values of K^T are stored row-wise).
__global__ void __launch_bounds__(128) k_robin_matrix(int64_t nsides, int nnps, int nqs, int nd, int nf, const int32_t* nodes,
                                                      const double* tab, const double* D, const double* X, double* nz,
                                                      const int32_t* adjptr, const int32_t* adj, const uint16_t* coloff,
                                                      const uint8_t* freemask, const int64_t* rowstart, int csc) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nsides) return;
  const double* N = tab;
  const double* dN = tab + (size_t)nqs * nnps;
  const double* w = dN + (size_t)nqs * nnps * (nd - 1);
  double x[kMaxLoadNodes][3];
  int32_t node[kMaxLoadNodes];
  for (int a = 0; a < nnps; ++a) {
    node[a] = nodes[e * nnps + a];
    for (int j = 0; j < 3; ++j) x[a][j] = j < nd ? X[(size_t)node[a] * nd + j] : 0.0;
  }
  for (int i = 0; i < nnps; ++i)
    for (int j = 0; j < nnps; ++j) {
      double kij[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // [di + nf*dj]
      for (int q = 0; q < nqs; ++q) {
        const double s = surface_jxw(x, dN, w, q, nnps, nd) * N[(size_t)q * nnps + i] * N[(size_t)q * nnps + j];
        const double* Dq = D + ((size_t)e * nqs + q) * nf * nf;
        for (int k = 0; k < nf * nf; ++k) kij[k] = fma(s, Dq[k], kij[k]);
      }
      // stored (row, col) = (dof(j,dj), dof(i,di)) in CSR; CSC stores the transpose row-wise -> (dof(i,di), dof(j,dj))
      const int rn = csc ? node[i] : node[j], cn = csc ? node[j] : node[i];
      const int32_t* row = adj + adjptr[rn];
      const int len = adjptr[rn + 1] - adjptr[rn];
      int pos = 0;
      while (pos < len && row[pos] != cn) ++pos;
      if (pos == len) continue;
      const unsigned cmask = freemask[cn];
      for (int di = 0; di < nf; ++di)
        for (int dj = 0; dj < nf; ++dj) {
          const int rd = csc ? di : dj, cd = csc ? dj : di;
          const int64_t rs = rowstart[(int64_t)rn * nf + rd];
          if (rs < 0 || !(cmask & (1u << cd))) continue;
          atomicAdd(&nz[rs + coloff[adjptr[rn] + pos] + __popc(cmask & ((1u << cd) - 1u))], kij[di + nf * dj]);
        }
    }
}

__global__ void k_add_field(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
// fused peer halo: entries of ghost nodes go straight into the owner's field over NVLink, like the element kernels do
__global__ void k_add_field_peer(double* dst, const double* __restrict__ src, int64_t ndof, int nf, PeerScatter peer) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ndof) return;
  const double v = src[i];
  if (v != 0.0) scatter_add(peer, dst, i / nf, nf, (int)(i % nf), v);
}

static PeerScatter local_only(const fecb200_handle* h) {
  PeerScatter ps = h->peer;
  ps.n_owned = -1;   // the cached vector is rank-local: ghost entries are forwarded when it is added (add_cached)
  return ps;
}

static void add_cached(fecb200_handle* h, double* field, const double* cache) {
  // ghost-node entries: with the NCCL halo they stay in the local field and the pack / send-recv / unpack-add sum carries
  // them to the owner; with the fused peer halo the field's ghost slots are never exchanged, so they are RED-added into
  // the owner's field here (between the same two stream-ordered barriers as the element kernels' ghost REDs)
  if (h->peer_enabled && h->peer_field == FECB200_FIELD_RESIDUAL && field == h->d_R.p && h->n_owned_nodes < h->nn)
    k_add_field_peer<<<grid_for(h->ndof, 256), 256, 0, h->stream>>>(field, cache, h->ndof, h->nf, h->peer);
  else
    k_add_field<<<grid_for(h->ndof, 256), 256, 0, h->stream>>>(field, cache, h->ndof);
  FEC_CUDA(cudaGetLastError());
  h->launches++;
}

void add_neumann_loads(fecb200_handle* h, double* field) {
  if (!h->has_neumann()) return;
  if (h->neumann_dirty || (int64_t)h->d_F_neumann.n != h->ndof) {
    if ((int64_t)h->d_F_neumann.n != h->ndof) h->d_F_neumann.alloc(h->ndof);
    h->d_F_neumann.zero(h->stream);
    for (auto& s : h->surface_loads) {
      if (!s.nsides) continue;
      FEC_REQUIRE(s.vals.p, "Neumann BC values were never set (fecb200_set_neumann_values)");
      k_surface_load<<<grid_for(s.nsides), 128, 0, h->stream>>>(s.nsides, s.nnps, s.nqs, h->nd, h->nf, s.nodes.p, s.tab.p, s.vals.p,
                                                               h->d_X.p, h->d_F_neumann.p, local_only(h));
      FEC_CUDA(cudaGetLastError());
      h->launches++;
    }
    h->neumann_dirty = false;
  }
  add_cached(h, field, h->d_F_neumann.p);
}

void add_source_loads(fecb200_handle* h, double* field) {
  if (!h->has_source()) return;
  if (h->source_dirty || (int64_t)h->d_F_source.n != h->ndof) {
    if ((int64_t)h->d_F_source.n != h->ndof) h->d_F_source.alloc(h->ndof);
    h->d_F_source.zero(h->stream);
    for (auto& b : h->blocks) {
      if (!b.d_body_force.p || b.halo || !b.ne) continue;
      if (!b.d_tab.p) {
        std::vector<double> t;
        t.insert(t.end(), b.N.begin(), b.N.end());
        t.insert(t.end(), b.dN.begin(), b.dN.end());
        t.insert(t.end(), b.w.begin(), b.w.end());
        b.d_tab.upload(t, h->stream);
      }
      k_body_force<<<grid_for(b.ne), 128, 0, h->stream>>>(b.ne, b.nnpe, b.nq, b.nd, h->nf, b.d_conn_perm.p, b.d_perm.p, b.d_tab.p,
                                                         b.d_body_force.p, h->d_X.p, h->d_F_source.p, local_only(h));
      FEC_CUDA(cudaGetLastError());
      h->launches++;
    }
    h->source_dirty = false;
  }
  add_cached(h, field, h->d_F_source.p);
}

}  // namespace fec

using namespace fec;

#define FEC_API_BEGIN try {
#define FEC_API_END                                     \
  return 0;                                             \
  }                                                     \
  catch (const std::exception& e) {                     \
    fec::g_last_error = e.what();                       \
    return 1;                                           \
  }                                                     \
  catch (...) {                                         \
    fec::g_last_error = "fecb200: unknown exception";   \
    return 1;                                           \
  }

static void upload_doubles(fecb200_handle* h, DevBuf<double>& dst, const double* src, size_t n) {
  if (dst.n != n) dst.alloc(n);
  if (!n) return;
  FEC_CUDA(cudaMemcpyAsync(dst.p, src, n * sizeof(double), is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                           h->stream));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
}

extern "C" {

int fecb200_set_neumann_bc(fecb200_handle* h, int32_t id, int64_t nsides, int32_t nnps, int32_t nqs, const int64_t* side_nodes,
                           const double* Ns, const double* dNs, const double* ws) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && id >= 0 && id <= (int)h->surface_loads.size(), "Neumann BC ids are 0, 1, 2, ... in order");
  FEC_REQUIRE(nsides >= 0 && nnps >= 1 && nnps <= kMaxLoadNodes && nqs >= 1, "bad side-set shape");
  FEC_REQUIRE(nsides == 0 || (side_nodes && Ns && dNs && ws), "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  if (id == (int)h->surface_loads.size()) h->surface_loads.emplace_back();
  SurfaceLoad& s = h->surface_loads[id];
  s.nsides = nsides; s.nnps = nnps; s.nqs = nqs;
  std::vector<int32_t> nodes((size_t)nsides * nnps);
  for (size_t i = 0; i < nodes.size(); ++i) {
    FEC_REQUIRE(side_nodes[i] >= 1 && side_nodes[i] <= h->nn, "side node id out of range");
    nodes[i] = (int32_t)(side_nodes[i] - 1);
  }
  s.nodes.upload(nodes, h->stream);
  const int ns = h->nd - 1;
  std::vector<double> t;
  if (nsides) {
    t.insert(t.end(), Ns, Ns + (size_t)nqs * nnps);
    t.insert(t.end(), dNs, dNs + (size_t)nqs * nnps * ns);
    t.insert(t.end(), ws, ws + nqs);
  }
  s.tab.upload(t, h->stream);
  s.vals.release();
  h->neumann_dirty = true;
  FEC_API_END
}

int fecb200_set_neumann_values(fecb200_handle* h, int32_t id, const double* vals) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && id >= 0 && id < (int)h->surface_loads.size(), "unknown Neumann BC id");
  FEC_CUDA(cudaSetDevice(h->device));
  SurfaceLoad& s = h->surface_loads[id];
  FEC_REQUIRE(vals || !s.nsides, "null argument");
  upload_doubles(h, s.vals, vals, (size_t)s.nsides * s.nqs * h->nf);
  h->neumann_dirty = true;
  FEC_API_END
}

int fecb200_clear_neumann_bcs(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  h->surface_loads.clear();
  h->neumann_dirty = true;
  FEC_API_END
}

int fecb200_set_source_values(fecb200_handle* h, int32_t block, const double* vals) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && block >= 0 && block < (int)h->blocks.size(), "bad block index");
  FEC_CUDA(cudaSetDevice(h->device));
  BlockPlan& b = h->blocks[block];
  FEC_REQUIRE(b.nnpe <= kMaxLoadNodes, "element type not supported by the load kernels");
  if (!vals) { FEC_CUDA(cudaStreamSynchronize(h->stream)); b.d_body_force.release(); }
  else upload_doubles(h, b.d_body_force, vals, (size_t)b.ne * b.nq * h->nf);
  h->source_dirty = true;
  FEC_API_END
}

int fecb200_set_robin_bc(fecb200_handle* h, int32_t id, int64_t nsides, int32_t nnps, int32_t nqs, const int64_t* side_nodes,
                         const double* Ns, const double* dNs, const double* ws) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && id >= 0 && id <= (int)h->robin_loads.size(), "Robin BC ids are 0, 1, 2, ... in order");
  FEC_REQUIRE(nsides >= 0 && nnps >= 1 && nnps <= kMaxLoadNodes && nqs >= 1, "bad side-set shape");
  FEC_REQUIRE(nsides == 0 || (side_nodes && Ns && dNs && ws), "null argument");
  FEC_REQUIRE(h->n_owned_nodes == h->nn, "Robin BCs are not supported on a partitioned handle");
  FEC_CUDA(cudaSetDevice(h->device));
  if (id == (int)h->robin_loads.size()) h->robin_loads.emplace_back();
  SurfaceLoad& s = h->robin_loads[id].geo;
  s.nsides = nsides; s.nnps = nnps; s.nqs = nqs;
  std::vector<int32_t> nodes((size_t)nsides * nnps);
  for (size_t i = 0; i < nodes.size(); ++i) {
    FEC_REQUIRE(side_nodes[i] >= 1 && side_nodes[i] <= h->nn, "side node id out of range");
    nodes[i] = (int32_t)(side_nodes[i] - 1);
  }
  s.nodes.upload(nodes, h->stream);
  std::vector<double> t;
  if (nsides) {
    t.insert(t.end(), Ns, Ns + (size_t)nqs * nnps);
    t.insert(t.end(), dNs, dNs + (size_t)nqs * nnps * (h->nd - 1));
    t.insert(t.end(), ws, ws + nqs);
  }
  s.tab.upload(t, h->stream);
  h->robin_loads[id].g0.release();
  h->robin_loads[id].D.release();
  FEC_API_END
}

int fecb200_set_robin_values(fecb200_handle* h, int32_t id, const double* g0, const double* dvalsdu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && id >= 0 && id < (int)h->robin_loads.size(), "unknown Robin BC id");
  FEC_CUDA(cudaSetDevice(h->device));
  RobinLoad& r = h->robin_loads[id];
  FEC_REQUIRE((g0 && dvalsdu) || !r.geo.nsides, "null argument");
  upload_doubles(h, r.g0, g0, (size_t)r.geo.nsides * r.geo.nqs * h->nf);
  upload_doubles(h, r.D, dvalsdu, (size_t)r.geo.nsides * r.geo.nqs * h->nf * h->nf);
  FEC_API_END
}

int fecb200_clear_robin_bcs(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  h->robin_loads.clear();
  FEC_API_END
}

int fecb200_assemble_vector_robin_bc(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  for (auto& r : h->robin_loads) {
    const SurfaceLoad& s = r.geo;
    if (!s.nsides) continue;
    FEC_REQUIRE(r.g0.p && r.D.p, "Robin BC values were never set (fecb200_set_robin_values)");
    k_robin_vector<<<grid_for(s.nsides), 128, 0, h->stream>>>(s.nsides, s.nnps, s.nqs, h->nd, h->nf, s.nodes.p, s.tab.p, r.g0.p, r.D.p,
                                                             h->d_X.p, h->d_U.p, h->d_R.p);
    FEC_CUDA(cudaGetLastError());
    h->launches++;
  }
  FEC_API_END
}

int fecb200_assemble_matrix_robin_bc(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_REQUIRE(!h->opts.matrix_free, "assemble_matrix_robin_bc! called on a matrix-free SparseMatrixAssembler");
  ensure_matrix_structure(h);
  FEC_REQUIRE(h->matrix_ready && h->d_nz_stiff.p, "assemble_stiffness! must run before assemble_matrix_robin_bc!");
  FEC_REQUIRE(!h->stiff_adjusted, "assemble_matrix_robin_bc! must run before stiffness(asm) applies the constraint adjustment");
  FEC_REQUIRE(h->per_b.empty(), "Robin BCs together with periodic BCs are not supported in matrix assembly");
  for (auto& r : h->robin_loads) {
    const SurfaceLoad& s = r.geo;
    if (!s.nsides) continue;
    FEC_REQUIRE(r.D.p, "Robin BC values were never set (fecb200_set_robin_values)");
    k_robin_matrix<<<grid_for(s.nsides), 128, 0, h->stream>>>(s.nsides, s.nnps, s.nqs, h->nd, h->nf, s.nodes.p, s.tab.p, r.D.p, h->d_X.p,
                                                             h->d_nz_stiff.p, h->d_adjptr.p, h->d_adj.p, h->d_coloff.p,
                                                             h->d_freemask.p, h->d_rowstart.p, h->opts.matrix_type == FECB200_CSC);
    FEC_CUDA(cudaGetLastError());
    h->launches++;
  }
  FEC_API_END
}

int fecb200_assemble_vector_neumann_bc(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  add_neumann_loads(h, h->d_R.p);
  FEC_API_END
}

int fecb200_assemble_vector_source(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  add_source_loads(h, h->d_R.p);
  FEC_API_END
}

}  // extern "C"
