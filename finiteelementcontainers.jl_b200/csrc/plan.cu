// plan.cu -- host-side planning: locality tiles, node adjacency, DOF maps, CSR structure.
//
// Restates, as data-structure builders for the CUDA kernels:
//   SparseMatrixPattern(dof)            src/assemblers/SparsityPatterns.jl:53-117
//   _update_dofs!(pattern, dof, ...)    src/assemblers/SparsityPatterns.jl:160-231
//   update_dofs!(dof, ...)              src/DofManagers.jl:227-298
// The reference stores one (I, J) pair per COO entry (NE * NDOF^2 Int64 each) plus a sort
// permutation; here the same pattern is held as a NODE adjacency (NN * ~27 int32) plus one
// position byte per element node pair, from which every CSR slot is computed in the kernel.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <omp.h>
#include <parallel/algorithm>

// METIS (libmetis_static.a shipped with the CUDA toolkit; 64-bit idx_t, 32-bit real_t; no metis.h in the image)
extern "C" {
int METIS_SetDefaultOptions(int64_t* options);
int METIS_PartMeshDual(int64_t* ne, int64_t* nn, int64_t* eptr, int64_t* eind, int64_t* vwgt, int64_t* vsize,
                       int64_t* ncommon, int64_t* nparts, float* tpwgts, int64_t* options, int64_t* objval,
                       int64_t* epart, int64_t* npart);
int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                        int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options,
                        int64_t* edgecut, int64_t* part);
}

namespace fec {

int metis_mesh_dual(int64_t ne, int64_t nn, const int64_t* eptr, const int64_t* eind, int64_t ncommon, int64_t nparts,
                    int64_t* epart, int64_t* npart) {
  int64_t options[40];
  METIS_SetDefaultOptions(options);
  options[17] = 0;  // METIS_OPTION_NUMBERING = C-style
  int64_t objval = 0;
  if (nparts == 1) {
    std::fill(epart, epart + ne, 0);
    std::fill(npart, npart + nn, 0);
    return 1;
  }
  return METIS_PartMeshDual(&ne, &nn, const_cast<int64_t*>(eptr), const_cast<int64_t*>(eind), nullptr, nullptr, &ncommon,
                            &nparts, nullptr, options, &objval, epart, npart);
}
int metis_graph(int64_t nv, const int64_t* xadj, const int64_t* adjncy, int64_t nparts, int64_t* part) {
  int64_t options[40];
  METIS_SetDefaultOptions(options);
  options[17] = 0;
  int64_t ncon = 1, cut = 0;
  if (nparts == 1) { std::fill(part, part + nv, 0); return 1; }
  return METIS_PartGraphKway(&nv, &ncon, const_cast<int64_t*>(xadj), const_cast<int64_t*>(adjncy), nullptr, nullptr, nullptr,
                             &nparts, nullptr, nullptr, options, &cut, part);
}

static inline uint64_t spread3(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
static inline uint64_t spread2(uint32_t v) {
  uint64_t x = v;
  x = (x | x << 16) & 0x0000ffff0000ffffULL;
  x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
  x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | x << 2) & 0x3333333333333333ULL;
  x = (x | x << 1) & 0x5555555555555555ULL;
  return x;
}

// Elements are sorted along a Morton curve of their centroids (quantised on a grid with
// ~NE^(1/ND) cells per axis, so structured meshes -- StructuredMesh.jl enumerates ez fastest,
// nodes x fastest -- fall into exact bricks: 256 elements = 8x8x4) and cut into tiles of `te`.
void build_block_tiles(fecb200_handle* h, BlockPlan& b, const double* coords) {
  if (use_gpu_plan()) { build_block_tiles_gpu(h, b); return; }
  PhaseTimer _pt("build_block_tiles");
  const int nd = h->nd, nnpe = b.nnpe, nf = h->nf;
  const int64_t ne = b.ne;
  const int te = b.te;
  FEC_REQUIRE((int64_t)nnpe * nf * te <= 65536, "tile too large for 16-bit incidence slots");
  // bounding box of the block's nodes
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma omp parallel
  {
    double tlo[3] = {1e300, 1e300, 1e300}, thi[3] = {-1e300, -1e300, -1e300};
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < ne * nnpe; ++i) {
      const int n = b.conn0[i];
      for (int j = 0; j < nd; ++j) {
        const double c = coords[(size_t)n * nd + j];
        tlo[j] = std::min(tlo[j], c);
        thi[j] = std::max(thi[j], c);
      }
    }
#pragma omp critical
    for (int j = 0; j < nd; ++j) { lo[j] = std::min(lo[j], tlo[j]); hi[j] = std::max(hi[j], thi[j]); }
  }
  int nbins = (int)std::llround(std::pow((double)ne, 1.0 / nd));
  const int maxbins = (nd == 3) ? (1 << 20) : (1 << 30);
  nbins = std::max(1, std::min(nbins, maxbins));
  std::vector<uint64_t> key(ne);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    uint32_t ic[3] = {0, 0, 0};
    for (int j = 0; j < nd; ++j) {
      double c = 0.0;
      for (int a = 0; a < nnpe; ++a) c += coords[(size_t)b.conn0[e * nnpe + a] * nd + j];
      c /= nnpe;
      const double ext = hi[j] - lo[j];
      double t = ext > 0 ? (c - lo[j]) / ext * nbins : 0.0;
      int it = (int)std::floor(t);
      ic[j] = (uint32_t)std::max(0, std::min(it, nbins - 1));
    }
    key[e] = (nd == 3) ? (spread3(ic[0]) | spread3(ic[1]) << 1 | spread3(ic[2]) << 2)
                       : (spread2(ic[0]) | spread2(ic[1]) << 1);
  }
  b.perm.resize(ne);
  {
    // (key, original index) pairs: a plain lexicographic sort is the stable sort by key, and runs multi-threaded
    std::vector<std::pair<uint64_t, int32_t>> ki(ne);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; ++e) ki[e] = {key[e], (int32_t)e};
    __gnu_parallel::sort(ki.begin(), ki.end());
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; ++e) b.perm[e] = ki[e].second;
  }

  b.ntiles = (int)((ne + te - 1) / te);
  std::vector<int32_t> tile_node_ptr(b.ntiles + 1, 0);
  std::vector<std::vector<int32_t>> tnodes(b.ntiles);
  std::vector<uint16_t> lconn((size_t)b.ntiles * nnpe * te, 0);
  std::vector<int32_t> conn_perm((size_t)ne * nnpe);
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < b.ntiles; ++t) {
    const int64_t e0 = (int64_t)t * te, e1 = std::min<int64_t>(ne, e0 + te);
    std::vector<int32_t>& nodes = tnodes[t];
    nodes.reserve((e1 - e0) * nnpe);
    for (int64_t e = e0; e < e1; ++e)
      for (int a = 0; a < nnpe; ++a) {
        const int32_t n = b.conn0[(size_t)b.perm[e] * nnpe + a];
        nodes.push_back(n);
        conn_perm[(size_t)e * nnpe + a] = n;
      }
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
    for (int64_t e = e0; e < e1; ++e)
      for (int a = 0; a < nnpe; ++a) {
        const int32_t n = b.conn0[(size_t)b.perm[e] * nnpe + a];
        const int l = (int)(std::lower_bound(nodes.begin(), nodes.end(), n) - nodes.begin());
        lconn[((size_t)t * nnpe + a) * te + (e - e0)] = (uint16_t)l;
      }
  }
  b.max_tile_nodes = 0;
  for (int t = 0; t < b.ntiles; ++t) {
    tile_node_ptr[t + 1] = tile_node_ptr[t] + (int32_t)tnodes[t].size();
    b.max_tile_nodes = std::max<int>(b.max_tile_nodes, (int)tnodes[t].size());
  }
  FEC_REQUIRE(b.max_tile_nodes <= 65535, "tile has too many nodes");
  const size_t tot_nodes = tile_node_ptr[b.ntiles];
  std::vector<int32_t> tile_nodes(tot_nodes);
  std::vector<int32_t> inc_ptr(tot_nodes + 1, 0);
  // incidence lists: for every tile-local node the (a, t) slots of the element vectors it sums
  std::vector<int32_t> counts(tot_nodes, 0);
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < b.ntiles; ++t) {
    std::copy(tnodes[t].begin(), tnodes[t].end(), tile_nodes.begin() + tile_node_ptr[t]);
    const int64_t e0 = (int64_t)t * te, e1 = std::min<int64_t>(ne, e0 + te);
    for (int64_t e = e0; e < e1; ++e)
      for (int a = 0; a < nnpe; ++a) counts[tile_node_ptr[t] + lconn[((size_t)t * nnpe + a) * te + (e - e0)]]++;
  }
  for (size_t i = 0; i < tot_nodes; ++i) inc_ptr[i + 1] = inc_ptr[i] + counts[i];
  std::vector<uint16_t> inc((size_t)inc_ptr[tot_nodes]);
  std::fill(counts.begin(), counts.end(), 0);
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < b.ntiles; ++t) {
    const int64_t e0 = (int64_t)t * te, e1 = std::min<int64_t>(ne, e0 + te);
    // fixed summation order: element slot ascending, then local node index
    for (int64_t e = e0; e < e1; ++e)
      for (int a = 0; a < nnpe; ++a) {
        const size_t g = tile_node_ptr[t] + lconn[((size_t)t * nnpe + a) * te + (e - e0)];
        inc[inc_ptr[g] + counts[g]++] = (uint16_t)(a * nf * te + (e - e0));
      }
  }
  b.d_perm.upload(b.perm, h->stream);
  b.d_tile_node_ptr.upload(tile_node_ptr, h->stream);
  b.d_tile_nodes.upload(tile_nodes, h->stream);
  b.d_lconn.upload(lconn, h->stream);
  b.d_inc_ptr.upload(inc_ptr, h->stream);
  b.d_inc.upload(inc, h->stream);
  b.d_conn_perm.upload(conn_perm, h->stream);
}

static inline const std::vector<int32_t>& scatter_conn(const BlockPlan& b) { return b.sconn0.empty() ? b.conn0 : b.sconn0; }

// Node adjacency = sparsity pattern of the condensed operator at node granularity.
// Row n lists, ascending, every node sharing an element with n (all blocks).  With periodic BCs the scatter
// connectivity has side-b nodes replaced by their side-a node (dof_to_unknown_index, DofManagers.jl:188-201).
void build_adjacency(fecb200_handle* h) {
  if (use_gpu_plan()) { build_adjacency_gpu(h); return; }
  PhaseTimer _pt("build_adjacency");
  const int64_t nn = h->nn;
  // node -> (block, element) incidence via counting sort
  std::vector<int64_t> nptr(nn + 1, 0);
  for (auto& b : h->blocks) {
    const std::vector<int32_t>& c = scatter_conn(b);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)c.size(); ++i) {
#pragma omp atomic
      nptr[c[i] + 1]++;
    }
  }
  for (int64_t n = 0; n < nn; ++n) nptr[n + 1] += nptr[n];
  std::vector<int64_t> fill(nptr.begin(), nptr.end() - 1);
  struct Ref { int32_t blk; int32_t el; };
  std::vector<Ref> refs(nptr[nn]);   // order inside a node's list is irrelevant: the neighbour sets are sorted below
  for (size_t bi = 0; bi < h->blocks.size(); ++bi) {
    auto& b = h->blocks[bi];
    const std::vector<int32_t>& c = scatter_conn(b);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < b.ne; ++e)
      for (int a = 0; a < b.nnpe; ++a) {
        int64_t pos;
#pragma omp atomic capture
        pos = fill[c[e * b.nnpe + a]]++;
        refs[pos] = {(int32_t)bi, (int32_t)e};
      }
  }
  h->adjptr.assign(nn + 1, 0);
  std::vector<int32_t> cnt(nn);
  // one pass: schedule(static) hands every thread ONE contiguous node range, so each thread appends its rows to a
  // private buffer in node order; the buffers are then copied behind each other
  const int nt = omp_get_max_threads();
  std::vector<std::vector<int32_t>> tout(nt);
  std::vector<int64_t> tfirst(nt, -1);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    std::vector<int32_t>& out = tout[t];
    std::vector<int32_t> tmp;
#pragma omp for schedule(static)
    for (int64_t n = 0; n < nn; ++n) {
      if (tfirst[t] < 0) tfirst[t] = n;
      tmp.clear();
      for (int64_t k = nptr[n]; k < nptr[n + 1]; ++k) {
        const auto& b = h->blocks[refs[k].blk];
        const int32_t* c = &scatter_conn(b)[(size_t)refs[k].el * b.nnpe];
        tmp.insert(tmp.end(), c, c + b.nnpe);
      }
      std::sort(tmp.begin(), tmp.end());
      const auto end = std::unique(tmp.begin(), tmp.end());
      cnt[n] = (int32_t)(end - tmp.begin());
      out.insert(out.end(), tmp.begin(), end);
    }
  }
  int64_t tot = 0;
  for (int64_t n = 0; n < nn; ++n) { h->adjptr[n] = (int32_t)tot; tot += cnt[n]; }
  FEC_REQUIRE(tot < (int64_t)INT32_MAX, "node adjacency exceeds int32 range");
  h->adjptr[nn] = (int32_t)tot;
  h->adj.resize(tot);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int t = 0; t < nt; ++t)
    if (tfirst[t] >= 0) std::copy(tout[t].begin(), tout[t].end(), h->adj.begin() + h->adjptr[tfirst[t]]);
  tout.clear();
  h->d_adjptr.upload(h->adjptr, h->stream);
  h->d_adj.upload(h->adj, h->stream);
  // element -> adjacency-position bytes (the element -> CSR slot map)
  for (auto& b : h->blocks) {
    const int nnpe = b.nnpe;
    std::vector<uint8_t> epos((size_t)b.ne * nnpe * nnpe);
    bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
    for (int64_t e = 0; e < b.ne; ++e) {
      const int32_t* c = &scatter_conn(b)[(size_t)b.perm[e] * nnpe];
      for (int r = 0; r < nnpe; ++r) {
        const int32_t* row = &h->adj[h->adjptr[c[r]]];
        const int len = h->adjptr[c[r] + 1] - h->adjptr[c[r]];
        if (len > 256) ok = false;
        for (int a = 0; a < nnpe; ++a) {
          const int pos = (int)(std::lower_bound(row, row + len, c[a]) - row);
          epos[((size_t)e * nnpe + r) * nnpe + a] = (uint8_t)pos;
        }
      }
    }
    FEC_REQUIRE(ok, "a node has more than 256 neighbours");
    b.d_epos.upload(epos, h->stream);
  }
}

// DofManager maps + dof-level CSR offsets after update_dofs!.
void build_dof_structures(fecb200_handle* h) {
  PhaseTimer _pt("build_dof_structures");
  const int nf = h->nf;
  const int64_t nn = h->nn, ndof = h->ndof;
  // ---- update_dofs!(dof, dirichlet, per_a, per_b)  (DofManagers.jl:227-298)
  std::vector<int64_t>& dd = h->dirichlet_dofs;
  std::sort(dd.begin(), dd.end());
  dd.erase(std::unique(dd.begin(), dd.end()), dd.end());
  for (int64_t d : dd) FEC_REQUIRE(d >= 1 && d <= ndof, "dirichlet dof out of range");
  // resolve periodic chains (:300-325) and drop duplicate pairs
  {
    std::vector<int64_t> a = h->per_a, bb = h->per_b;
    std::vector<std::pair<int64_t, int64_t>> map;  // b -> a
    for (size_t i = 0; i < bb.size(); ++i) map.push_back({bb[i], a[i]});
    std::sort(map.begin(), map.end());
    auto find = [&](int64_t d) -> const std::pair<int64_t, int64_t>* {
      auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(d, (int64_t)INT64_MIN));
      // last entry for a repeated key wins (Dict insertion semantics)
      const std::pair<int64_t, int64_t>* hit = nullptr;
      while (it != map.end() && it->first == d) { hit = &*it; ++it; }
      return hit;
    };
    std::vector<int64_t> ra, rb;
    std::vector<std::pair<int64_t, int64_t>> seen;
    for (size_t i = 0; i < bb.size(); ++i) {
      int64_t d = find(bb[i])->second;
      int guard = 0;
      while (auto* p = find(d)) { d = p->second; FEC_REQUIRE(++guard < 1000000, "periodic chain cycle"); }
      std::pair<int64_t, int64_t> pr{d, bb[i]};
      if (std::find(seen.begin(), seen.end(), pr) == seen.end()) {
        seen.push_back(pr); ra.push_back(d); rb.push_back(bb[i]);
      }
    }
    for (size_t i = 0; i < ra.size(); ++i)
      FEC_REQUIRE(ra[i] >= 1 && ra[i] <= ndof && rb[i] >= 1 && rb[i] <= ndof, "periodic dof out of range");
    h->per_a = ra; h->per_b = rb;
  }
  std::vector<int64_t>& d2u = h->dof_to_unknown;
  d2u.assign(ndof, 0);
  for (int64_t d : dd) d2u[d - 1] = -1;
  for (int64_t d : h->per_b) d2u[d - 1] = -2;
  // unknown ids in ascending dof order: chunked count / prefix / fill (21.6 M dofs at 192^3)
  {
    const int nth = omp_get_max_threads();
    std::vector<int64_t> cnt(nth + 1, 0);
    const int64_t chunk = (ndof + nth - 1) / nth;
#pragma omp parallel num_threads(nth)
    {
      const int t = omp_get_thread_num();
      const int64_t lo = std::min<int64_t>(ndof, t * chunk), hi = std::min<int64_t>(ndof, lo + chunk);
      int64_t c = 0;
      for (int64_t g = lo; g < hi; ++g) c += d2u[g] == 0;
      cnt[t + 1] = c;
#pragma omp barrier
#pragma omp single
      {
        for (int i = 0; i < nth; ++i) cnt[i + 1] += cnt[i];
        h->unknown_dofs.resize(cnt[nth]);
      }
      int64_t k = cnt[t];
      for (int64_t g = lo; g < hi; ++g)
        if (d2u[g] == 0) { h->unknown_dofs[k] = g + 1; d2u[g] = ++k; }
    }
  }
  h->n_unknowns = (int64_t)h->unknown_dofs.size();
  h->b2a_unknown.assign(ndof, 0);
  for (size_t i = 0; i < h->per_a.size(); ++i) {
    FEC_REQUIRE(d2u[h->per_a[i] - 1] > 0, "periodic side-a dof is constrained");
    h->b2a_unknown[h->per_b[i] - 1] = d2u[h->per_a[i] - 1];
  }
  // device copies
  {
    std::vector<int32_t> ud(h->n_unknowns);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < h->n_unknowns; ++k) ud[k] = (int32_t)(h->unknown_dofs[k] - 1);
    h->d_unknown_dofs.upload(ud, h->stream);
    std::vector<int32_t> d2ui(ndof);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < ndof; ++g)
      d2ui[g] = h->opts.condensed ? (int32_t)g : (d2u[g] > 0 ? (int32_t)(d2u[g] - 1) : -1);
    h->d_d2u.upload(d2ui, h->stream);
    // constraint_storage: 1.0 at the Dirichlet dofs (SparseMatrixAssembler.jl:93-94): cleared on the device, the few
    // non-zeros set by index
    if ((int64_t)h->d_constraint.n != ndof) h->d_constraint.alloc(ndof);
    h->d_constraint.zero(h->stream);
    if (!dd.empty()) {
      std::vector<int32_t> di(dd.size());
      for (size_t i = 0; i < dd.size(); ++i) di[i] = (int32_t)(dd[i] - 1);
      DevBuf<int32_t> ddev;
      ddev.upload(di, h->stream);
      fill_indexed(h, h->d_constraint.p, ddev.p, 1.0, (int64_t)di.size());
      FEC_CUDA(cudaStreamSynchronize(h->stream));
    }
    std::vector<int32_t> pa(h->per_a.size()), pb(h->per_b.size());
    for (size_t i = 0; i < pa.size(); ++i) { pa[i] = (int32_t)(h->per_a[i] - 1); pb[i] = (int32_t)(h->per_b[i] - 1); }
    h->n_per = (int64_t)pa.size();
    h->d_per_a.upload(pa, h->stream);
    h->d_per_b.upload(pb, h->stream);
    std::vector<double> pv(pa.size(), 0.0);
    h->d_per_vals.upload(pv, h->stream);
  }
  // default Dirichlet values: zeros at the (sorted, unique) Dirichlet dofs
  {
    std::vector<int32_t> bd(dd.size());
    for (size_t i = 0; i < dd.size(); ++i) bd[i] = (int32_t)(dd[i] - 1);
    h->n_bc = (int64_t)bd.size();
    h->d_bc_dofs.upload(bd, h->stream);
    std::vector<double> bv(dd.size(), 0.0);
    h->d_bc_vals.upload(bv, h->stream);
  }
  // staging vectors: kept across update_dofs calls (cudaFree / cudaMalloc of 3 x 173 MB at 192^3 is not free)
  if ((int64_t)h->d_Uu.n != ndof) h->d_Uu.alloc(ndof);
  if ((int64_t)h->d_Vu.n != ndof) h->d_Vu.alloc(ndof);
  if ((int64_t)h->d_out.n != ndof) h->d_out.alloc(ndof);

  // the CSR structure depends on the kept dofs: rebuilt on demand (ensure_matrix_structure)
  h->matrix_ready = false;
  h->matrix_dirty = !h->opts.matrix_free;
}

void ensure_matrix_structure(fecb200_handle* h) {
  if (!h->matrix_dirty) return;
  h->matrix_dirty = false;
  build_matrix_structure(h);
}

// Periodic BCs in the assembled matrix: _update_dofs! maps a side-b dof to the unknown of its side-a dof
// (SparsityPatterns.jl:160-231 with dof_to_unknown_index, DofManagers.jl:188-201), i.e. rows / columns of b are
// summed into those of a.  Here that is a node-level fold of the SCATTER connectivity: the adjacency, the position
// bytes and the scatter records are rebuilt on it, the gathers keep the original connectivity.
static void fold_periodic_nodes(fecb200_handle* h) {
  const bool need = !h->per_b.empty();
  if (!need && !h->adj_folded) return;
  const int nf = h->nf;
  std::vector<int32_t> rep;
  if (need) {
    FEC_REQUIRE(h->n_owned_nodes == h->nn, "periodic BCs are not supported on a partitioned handle");
    rep.resize(h->nn);
    std::iota(rep.begin(), rep.end(), 0);
    std::vector<int32_t> cnt(h->nn, 0);
    for (size_t i = 0; i < h->per_b.size(); ++i) {
      const int64_t gb = h->per_b[i] - 1, ga = h->per_a[i] - 1;
      const int64_t nb = gb / nf, na = ga / nf;
      FEC_REQUIRE(gb % nf == ga % nf, "periodic pair couples different field components: not supported in matrix assembly");
      FEC_REQUIRE(cnt[nb] == 0 || rep[nb] == (int32_t)na, "the dofs of a periodic side-b node map to different side-a nodes");
      rep[nb] = (int32_t)na;
      cnt[nb]++;
    }
    for (int64_t n = 0; n < h->nn; ++n)
      FEC_REQUIRE(cnt[n] == 0 || cnt[n] == nf,
                  "a node with only some components periodic is not supported in matrix assembly (use the matrix-free path)");
  }
  for (auto& b : h->blocks) {
    if (need) {
      b.sconn0.resize(b.conn0.size());
      for (size_t i = 0; i < b.conn0.size(); ++i) b.sconn0[i] = rep[b.conn0[i]];
      std::vector<int32_t> sp((size_t)b.ne * b.nnpe);
      for (int64_t e = 0; e < b.ne; ++e)
        for (int a = 0; a < b.nnpe; ++a) sp[(size_t)e * b.nnpe + a] = b.sconn0[(size_t)b.perm[e] * b.nnpe + a];
      b.d_sconn_perm.upload(sp, h->stream);
    } else {
      b.sconn0.clear();
      b.d_sconn_perm.release();
    }
  }
  build_adjacency(h);
  h->adj_folded = need;
}

// dof-level CSR offsets on top of the node adjacency (rebuilt by update_dofs and partition_setup)
void build_matrix_structure(fecb200_handle* h) {
  PhaseTimer _pt("build_matrix_structure");
  h->matrix_dirty = false;
  const int nf = h->nf;
  const int64_t nn = h->nn, ndof = h->ndof;
  const std::vector<int64_t>& d2u = h->dof_to_unknown;
  // ---- CSR structure (matrix_free assemblers carry none: SparseMatrixAssembler.jl:82-88)
  h->matrix_ready = false;
  if (h->opts.matrix_free) return;
  FEC_REQUIRE(h->per_b.empty() || h->opts.condensed == 0,
              "Currently not supported periodic bcs in condensed mode");  // SparseMatrixAssembler.jl:251
  fold_periodic_nodes(h);
  if (use_gpu_plan()) {
    build_matrix_offsets_gpu(h);
    h->d_nz_stiff.alloc_compressible(nz_alloc_len(h), h->device);
    h->d_nz_stiff.zero(h->stream);
    h->d_nz_stiff_alt.release();
    h->alt_clean = false;
    if (h->double_buffer) { h->d_nz_stiff_alt.alloc_compressible(nz_alloc_len(h), h->device); h->d_nz_stiff_alt.zero(h->stream); h->alt_clean = true; }
    h->d_nz_mass.release();
    h->matrix_ready = true;
    h->stiff_adjusted = h->mass_adjusted = false;
    build_ecol(h);
    return;
  }
  const bool condensed = h->opts.condensed != 0;
  h->freemask_h.assign(nn, 0);
  std::vector<uint8_t> nfree(nn);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < nn; ++n) {
    unsigned m = 0;
    for (int d = 0; d < nf; ++d)
      if (condensed || d2u[n * nf + d] > 0) m |= 1u << d;
    h->freemask_h[n] = (uint8_t)m;
    nfree[n] = (uint8_t)__builtin_popcount(m);
  }
  const int64_t nadj = h->adjptr[nn];
  std::vector<uint16_t> coloff(nadj);
  std::vector<int32_t> rowlen(nn);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < nn; ++n) {
    int off = 0;
    for (int32_t k = h->adjptr[n]; k < h->adjptr[n + 1]; ++k) {
      coloff[k] = (uint16_t)off;
      off += nfree[h->adj[k]];
    }
    rowlen[n] = off;
  }
  h->max_rowlen = 0;
  for (int64_t n = 0; n < nn; ++n) {
    FEC_REQUIRE(rowlen[n] < 65536, "row too long for 16-bit column offsets");
    h->max_rowlen = std::max(h->max_rowlen, rowlen[n]);
  }
  h->rowstart_h.assign(ndof, -1);
  std::vector<int64_t> diag(ndof, -1);
  // first value slot of every owned node (ghost rows are not stored): serial prefix sum, the rest in parallel
  std::vector<int64_t> nodebase(h->n_owned_nodes + 1, 0);
  int64_t nmat = 0;
  for (int64_t n = 0; n < h->n_owned_nodes; ++n) {
    nodebase[n + 1] = nodebase[n] + (int64_t)nfree[n] * rowlen[n];
    nmat += nfree[n];
  }
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < h->n_owned_nodes; ++n) {
    const unsigned m = h->freemask_h[n];
    // self position
    const int32_t* row = &h->adj[h->adjptr[n]];
    const int len = h->adjptr[n + 1] - h->adjptr[n];
    const int ks = (int)(std::lower_bound(row, row + len, (int32_t)n) - row);
    int64_t pos = nodebase[n];
    for (int d = 0; d < nf; ++d) {
      if (!(m & (1u << d))) continue;
      h->rowstart_h[n * nf + d] = pos;
      diag[n * nf + d] = pos + coloff[h->adjptr[n] + ks] + __builtin_popcount(m & ((1u << d) - 1u));
      pos += rowlen[n];
    }
  }
  h->nnz = nodebase[h->n_owned_nodes];
  h->nmat = nmat;
  h->d_coloff.upload(coloff, h->stream);
  h->d_freemask.upload(h->freemask_h, h->stream);
  h->d_rowstart.upload(h->rowstart_h, h->stream);
  h->d_diagslot.upload(diag, h->stream);
  h->d_nz_stiff.alloc_compressible(nz_alloc_len(h), h->device);
  h->d_nz_stiff.zero(h->stream);
  h->d_nz_stiff_alt.release();
  h->alt_clean = false;
  if (h->double_buffer) { h->d_nz_stiff_alt.alloc_compressible(nz_alloc_len(h), h->device); h->d_nz_stiff_alt.zero(h->stream); h->alt_clean = true; }
  h->d_nz_mass.release();
  h->matrix_ready = true;
  h->stiff_adjusted = h->mass_adjusted = false;
  build_ecol(h);
}

// rowptr/colval (CSR) or colptr/rowval (CSC), Int64 1-based, as SparseArrays.sparse! +
// SparseMatrixCSR(csc) produce them (SparsityPatterns.jl:301-329): indices ascending in every
// row/column, duplicates merged, explicit zeros kept.  The pattern is structurally symmetric, so
// both formats share the arrays.
void export_pattern(fecb200_handle* h, int64_t* ptr, int64_t* idx) {
  ensure_matrix_structure(h);
  FEC_REQUIRE(h->matrix_ready, "no matrix pattern (matrix_free assembler or update_dofs not called)");
  ensure_host_structure(h);
  const int nf = h->nf;
  const int64_t nn = h->nn;
  const bool condensed = h->opts.condensed != 0;
  int64_t row = 0;
  for (int64_t n = 0; n < std::min(nn, h->n_owned_nodes); ++n) {
    const unsigned m = h->freemask_h[n];
    for (int d = 0; d < nf; ++d) {
      if (!(m & (1u << d))) continue;
      int64_t p = h->rowstart_h[n * nf + d];
      if (ptr) ptr[row] = p + 1;
      if (idx) {
        for (int32_t k = h->adjptr[n]; k < h->adjptr[n + 1]; ++k) {
          const int64_t nb = h->adj[k];
          const unsigned mb = h->freemask_h[nb];
          for (int d2 = 0; d2 < nf; ++d2)
            if (mb & (1u << d2)) idx[p++] = condensed ? (nb * nf + d2 + 1) : h->dof_to_unknown[nb * nf + d2];
        }
      }
      ++row;
    }
  }
  if (ptr) ptr[row] = h->nnz + 1;
}

}  // namespace fec
