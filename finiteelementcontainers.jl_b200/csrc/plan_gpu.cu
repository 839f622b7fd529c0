// plan_gpu.cu -- the set-up of plan.cu on the device: locality tiles, node adjacency (= the sparsity
// pattern at node granularity), element -> adjacency-position bytes and the dof-level CSR offsets.
//
// Restates, as sort / scan / per-node kernels, what the reference builds with O(COO) host loops and a sortperm:
//   SparseMatrixPattern(dof)            src/assemblers/SparsityPatterns.jl:53-117
//   _update_dofs!(pattern, dof, ...)    src/assemblers/SparsityPatterns.jl:160-231
// Results are the same arrays plan.cu's OpenMP builders produce (checked bit for bit by every pattern test);
// FECB200_HOST_PLAN=1 selects the host builders instead.  Host copies of the adjacency / row starts are fetched
// lazily (export_pattern is the only host consumer).
#include "common.cuh"
#include <thrust/binary_search.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/transform_reduce.h>
#include <algorithm>
#include <cmath>

namespace fec {

namespace {
inline int grid_for(int64_t n, int bs = 256) { return (int)((n + bs - 1) / bs); }

__host__ __device__ inline uint64_t spread3(uint32_t v) {
  uint64_t x = v & 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
__host__ __device__ inline uint64_t spread2(uint32_t v) {
  uint64_t x = v;
  x = (x | x << 16) & 0x0000ffff0000ffffULL;
  x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
  x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | x << 2) & 0x3333333333333333ULL;
  x = (x | x << 1) & 0x5555555555555555ULL;
  return x;
}

struct Box { double lo[3], hi[3]; };
struct BoxOf {
  const int32_t* conn; const double* X; int nd;
  __device__ Box operator()(int64_t i) const {
    Box b;
    const int n = conn[i];
    for (int j = 0; j < 3; ++j) { const double c = j < nd ? X[(size_t)n * nd + j] : 0.0; b.lo[j] = c; b.hi[j] = c; }
    return b;
  }
};
struct BoxMerge {
  __host__ __device__ Box operator()(const Box& a, const Box& b) const {
    Box r;
    for (int j = 0; j < 3; ++j) { r.lo[j] = a.lo[j] < b.lo[j] ? a.lo[j] : b.lo[j]; r.hi[j] = a.hi[j] > b.hi[j] ? a.hi[j] : b.hi[j]; }
    return r;
  }
};

__global__ void k_morton(const int32_t* conn, const double* X, int nd, int nnpe, int64_t ne, Box box, int nbins, uint64_t* key) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  uint32_t ic[3] = {0, 0, 0};
  for (int j = 0; j < nd; ++j) {
    double c = 0.0;
    for (int a = 0; a < nnpe; ++a) c += X[(size_t)conn[e * nnpe + a] * nd + j];
    c /= nnpe;
    const double ext = box.hi[j] - box.lo[j];
    const double t = ext > 0 ? (c - box.lo[j]) / ext * nbins : 0.0;
    int it = (int)floor(t);
    it = it < 0 ? 0 : (it > nbins - 1 ? nbins - 1 : it);
    ic[j] = (uint32_t)it;
  }
  key[e] = (nd == 3) ? (spread3(ic[0]) | spread3(ic[1]) << 1 | spread3(ic[2]) << 2) : (spread2(ic[0]) | spread2(ic[1]) << 1);
}

__global__ void k_permute_conn(const int32_t* conn, const int32_t* perm, int nnpe, int64_t ne, int32_t* out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ne * nnpe) return;
  const int64_t e = i / nnpe;
  out[i] = conn[(int64_t)perm[e] * nnpe + (i - e * nnpe)];
}

// One CTA per tile: unique nodes (sorted), tile-local connectivity, incidence lists.  CAP = power of two >= te * nnpe.
template <int CAP>
__global__ void __launch_bounds__(128) k_tile_plan(const int32_t* conn_perm, int nnpe, int nf, int te, int64_t ne,
                                                   int32_t* tile_nodes_tmp, int32_t* inc_ptr_tmp, int32_t* counts,
                                                   uint16_t* lconn, uint16_t* inc) {
  __shared__ int32_t ids[CAP];
  __shared__ int32_t uniq[CAP];
  __shared__ int32_t cnt[CAP];     // occurrences per unique node, then running fill positions
  __shared__ int32_t ptr[CAP + 1];
  __shared__ int32_t nu_s;
  const int t = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t e0 = (int64_t)t * te;
  const int nel = (int)((ne - e0) < te ? (ne - e0) : te);
  const int n = nel * nnpe;
  for (int i = tid; i < CAP; i += nt) ids[i] = i < n ? conn_perm[e0 * nnpe + i] : 0x7fffffff;
  __syncthreads();
  // bitonic sort of CAP keys
  for (int k = 2; k <= CAP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < CAP; i += nt) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const int32_t a = ids[i], b = ids[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { ids[i] = b; ids[ixj] = a; }
        }
      }
      __syncthreads();
    }
  // unique: serial prefix by warp 0 over chunks would be slow; a simple two-level scan with the block
  if (tid == 0) nu_s = 0;
  for (int i = tid; i < CAP; i += nt) cnt[i] = (i < n && (i == 0 || ids[i] != ids[i - 1])) ? 1 : 0;
  __syncthreads();
  // exclusive scan of cnt (CAP entries) by thread 0..nt-1 in contiguous chunks
  {
    const int chunk = CAP / nt;   // CAP and nt are powers of two, CAP >= nt
    int s = 0;
    for (int i = 0; i < chunk; ++i) s += cnt[tid * chunk + i];
    ptr[tid] = s;
    __syncthreads();
    if (tid == 0) {
      int acc = 0;
      for (int i = 0; i < nt; ++i) { const int v = ptr[i]; ptr[i] = acc; acc += v; }
      nu_s = acc;
    }
    __syncthreads();
    int base = ptr[tid];
    __syncthreads();
    for (int i = 0; i < chunk; ++i) {
      const int k = tid * chunk + i;
      if (cnt[k]) uniq[base++] = ids[k];
    }
  }
  __syncthreads();
  const int nu = nu_s;
  for (int i = tid; i < nu; i += nt) { tile_nodes_tmp[(size_t)t * CAP + i] = uniq[i]; cnt[i] = 0; }
  if (tid == 0) counts[t] = nu;
  __syncthreads();
  // tile-local connectivity + occurrence counts
  for (int i = tid; i < n; i += nt) {
    const int el = i / nnpe, a = i - el * nnpe;
    const int32_t node = conn_perm[e0 * nnpe + i];
    int lo = 0, hi = nu;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (uniq[mid] < node) lo = mid + 1; else hi = mid; }
    lconn[((size_t)t * nnpe + a) * te + el] = (uint16_t)lo;
    ids[i] = lo;   // re-used: local node of entry i
    atomicAdd(&cnt[lo], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < nu; ++i) { ptr[i] = acc; acc += cnt[i]; }
    ptr[nu] = acc;
  }
  __syncthreads();
  for (int i = tid; i <= nu; i += nt) inc_ptr_tmp[(size_t)t * (CAP + 1) + i] = ptr[i];
  // incidence lists in a FIXED order: thread per local node walks the tile's entries element-major (slot ascending,
  // then local node index) -- the same order as plan.cu, so the per-tile sums are reproducible
  uint16_t* inc_t = inc + e0 * nnpe;
  for (int l = tid; l < nu; l += nt) {
    int pos = ptr[l];
    for (int i = 0; i < n; ++i)
      if (ids[i] == l) {
        const int el = i / nnpe, a = i - el * nnpe;
        inc_t[pos++] = (uint16_t)(a * nf * te + el);
      }
  }
}

__global__ void k_compact_tiles(const int32_t* tile_nodes_tmp, const int32_t* inc_ptr_tmp, const int32_t* tile_node_ptr, int cap,
                                int nnpe, int te, int ntiles, int32_t* tile_nodes, int32_t* inc_ptr) {
  const int t = blockIdx.x;
  const int base = tile_node_ptr[t], nu = tile_node_ptr[t + 1] - base;
  const int64_t e0 = (int64_t)t * te;
  for (int i = threadIdx.x; i < nu; i += blockDim.x) {
    tile_nodes[base + i] = tile_nodes_tmp[(size_t)t * cap + i];
    inc_ptr[base + i] = (int32_t)(e0 * nnpe) + inc_ptr_tmp[(size_t)t * (cap + 1) + i];
  }
}

// ---- adjacency
constexpr int kMaxBlocks = 16;
struct BlockRefs {
  int nblocks;
  const int32_t* conn[kMaxBlocks];   // scatter connectivity, tile order
  int64_t first[kMaxBlocks + 1];     // global element id range of each block
  int nnpe[kMaxBlocks];
};

__global__ void k_incidence_pairs(BlockRefs R, int b, int32_t* key, int32_t* val, int64_t offset) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (R.first[b + 1] - R.first[b]) * R.nnpe[b];
  if (i >= n) return;
  key[offset + i] = R.conn[b][i];
  val[offset + i] = (int32_t)(R.first[b] + i / R.nnpe[b]);
}

constexpr int kMaxAdj = 256;   // position bytes are uint8
template <bool FILL>
__global__ void __launch_bounds__(128) k_neighbours(BlockRefs R, const int32_t* nptr, const int32_t* inc_el, int64_t nn, int32_t* cnt,
                                                    const int32_t* adjptr, int32_t* adj, int* overflow) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  int32_t list[kMaxAdj];
  int len = 0;
  for (int k = nptr[n]; k < nptr[n + 1]; ++k) {
    const int32_t ge = inc_el[k];
    int b = 0;
    while (b + 1 < R.nblocks && ge >= R.first[b + 1]) ++b;
    const int nnpe = R.nnpe[b];
    const int32_t* c = R.conn[b] + (ge - R.first[b]) * nnpe;
    for (int a = 0; a < nnpe; ++a) {
      const int32_t m = c[a];
      int lo = 0, hi = len;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (list[mid] < m) lo = mid + 1; else hi = mid; }
      if (lo < len && list[lo] == m) continue;
      if (len >= kMaxAdj) { *overflow = 1; continue; }
      for (int i = len; i > lo; --i) list[i] = list[i - 1];
      list[lo] = m;
      ++len;
    }
  }
  if (!FILL) cnt[n] = len;
  else {
    int32_t* out = adj + adjptr[n];
    for (int i = 0; i < len; ++i) out[i] = list[i];
  }
}

__global__ void k_epos(const int32_t* conn, int nnpe, int64_t ne, const int32_t* adjptr, const int32_t* adj, uint8_t* epos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // (element, row node r)
  if (i >= ne * nnpe) return;
  const int64_t e = i / nnpe;
  const int32_t* c = conn + e * nnpe;
  const int32_t r = c[i - e * nnpe];
  const int32_t* row = adj + adjptr[r];
  const int len = adjptr[r + 1] - adjptr[r];
  for (int a = 0; a < nnpe; ++a) {
    const int32_t m = c[a];
    int lo = 0, hi = len;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] < m) lo = mid + 1; else hi = mid; }
    epos[i * nnpe + a] = (uint8_t)lo;
  }
}

struct PopMask { const uint8_t* m; __device__ int64_t operator()(int64_t n) const { return (int64_t)__popc((unsigned)m[n]); } };

// ---- dof-level CSR offsets
__global__ void k_freemask(const int32_t* d2u, int nf, int condensed, int64_t nn, uint8_t* freemask) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  unsigned m = 0;
  for (int d = 0; d < nf; ++d)
    if (condensed || d2u[n * nf + d] >= 0) m |= 1u << d;
  freemask[n] = (uint8_t)m;
}
__global__ void k_coloff(const int32_t* adjptr, const int32_t* adj, const uint8_t* freemask, int64_t nn, int64_t n_owned, uint16_t* coloff,
                         int32_t* rowlen, int64_t* node_vals, int* too_long) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  int off = 0;
  for (int k = adjptr[n]; k < adjptr[n + 1]; ++k) {
    coloff[k] = (uint16_t)off;
    off += __popc(freemask[adj[k]]);
  }
  if (off >= 65536) *too_long = 1;
  rowlen[n] = off;
  if (n < n_owned) node_vals[n] = (int64_t)__popc(freemask[n]) * off;
}
__global__ void k_rowstart(const int32_t* adjptr, const int32_t* adj, const uint8_t* freemask, const uint16_t* coloff, const int32_t* rowlen,
                           const int64_t* nodebase, int nf, int64_t nn, int64_t n_owned, int64_t* rowstart, int64_t* diag) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= nn) return;
  if (n >= n_owned) {
    for (int d = 0; d < nf; ++d) { rowstart[n * nf + d] = -1; diag[n * nf + d] = -1; }
    return;
  }
  const unsigned m = freemask[n];
  const int32_t* row = adj + adjptr[n];
  const int len = adjptr[n + 1] - adjptr[n];
  int lo = 0, hi = len;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (row[mid] < (int32_t)n) lo = mid + 1; else hi = mid; }
  int64_t pos = nodebase[n];
  for (int d = 0; d < nf; ++d) {
    if (!(m & (1u << d))) { rowstart[n * nf + d] = -1; diag[n * nf + d] = -1; continue; }
    rowstart[n * nf + d] = pos;
    diag[n * nf + d] = pos + coloff[adjptr[n] + lo] + __popc(m & ((1u << d) - 1u));
    pos += rowlen[n];
  }
}
}  // namespace

bool use_gpu_plan() {
  static const bool host = [] { const char* e = getenv("FECB200_HOST_PLAN"); return e && e[0] == '1'; }();
  return !host;
}

// build_block_tiles on the device: b.conn0 (host, caller's order) -> perm, tiles, tile-local connectivity, incidences
void build_block_tiles_gpu(fecb200_handle* h, BlockPlan& b) {
  PhaseTimer _pt("build_block_tiles (gpu)");
  const int nd = h->nd, nnpe = b.nnpe, nf = h->nf, te = b.te;
  const int64_t ne = b.ne;
  cudaStream_t s = h->stream;
  auto pol = thrust::cuda::par.on(s);
  FEC_REQUIRE((int64_t)nnpe * nf * te <= 65536, "tile too large for 16-bit incidence slots");
  DevBuf<int32_t> d_conn;
  d_conn.upload(b.conn0, s);
  // bounding box of the block's nodes, Morton keys of the element centroids, stable sort
  Box init;
  for (int j = 0; j < 3; ++j) { init.lo[j] = 1e300; init.hi[j] = -1e300; }
  const Box box = thrust::transform_reduce(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(ne * nnpe),
                                           BoxOf{d_conn.p, h->d_X.p, nd}, init, BoxMerge());
  int nbins = (int)std::llround(std::pow((double)ne, 1.0 / nd));
  const int maxbins = (nd == 3) ? (1 << 20) : (1 << 30);
  nbins = std::max(1, std::min(nbins, maxbins));
  DevBuf<uint64_t> d_key;
  d_key.alloc(ne);
  k_morton<<<grid_for(ne), 256, 0, s>>>(d_conn.p, h->d_X.p, nd, nnpe, ne, box, nbins, d_key.p);
  b.d_perm.alloc(ne);
  thrust::sequence(pol, thrust::device_pointer_cast(b.d_perm.p), thrust::device_pointer_cast(b.d_perm.p) + ne);
  thrust::stable_sort_by_key(pol, thrust::device_pointer_cast(d_key.p), thrust::device_pointer_cast(d_key.p) + ne,
                             thrust::device_pointer_cast(b.d_perm.p));
  b.d_conn_perm.alloc((size_t)ne * nnpe);
  k_permute_conn<<<grid_for(ne * nnpe), 256, 0, s>>>(d_conn.p, b.d_perm.p, nnpe, ne, b.d_conn_perm.p);
  b.perm.resize(ne);
  FEC_CUDA(cudaMemcpyAsync(b.perm.data(), b.d_perm.p, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  // tiles
  b.ntiles = (int)((ne + te - 1) / te);
  int cap = 128;
  while (cap < te * nnpe) cap <<= 1;
  FEC_REQUIRE(cap <= 2048, "tile plan: te * nnpe exceeds 2048");
  DevBuf<int32_t> d_tn_tmp, d_ip_tmp, d_counts;
  d_tn_tmp.alloc((size_t)b.ntiles * cap);
  d_ip_tmp.alloc((size_t)b.ntiles * (cap + 1));
  d_counts.alloc(b.ntiles + 1);
  b.d_lconn.alloc((size_t)b.ntiles * nnpe * te);
  FEC_CUDA(cudaMemsetAsync(b.d_lconn.p, 0, b.d_lconn.n * sizeof(uint16_t), s));
  b.d_inc.alloc((size_t)ne * nnpe);
#define FEC_TILE(CAP_) k_tile_plan<CAP_><<<b.ntiles, 128, 0, s>>>(b.d_conn_perm.p, nnpe, nf, te, ne, d_tn_tmp.p, d_ip_tmp.p, d_counts.p, b.d_lconn.p, b.d_inc.p)
  switch (cap) {
    case 128: FEC_TILE(128); break;
    case 256: FEC_TILE(256); break;
    case 512: FEC_TILE(512); break;
    case 1024: FEC_TILE(1024); break;
    default: FEC_TILE(2048); break;
  }
#undef FEC_TILE
  FEC_CUDA(cudaGetLastError());
  b.max_tile_nodes = thrust::reduce(pol, thrust::device_pointer_cast(d_counts.p), thrust::device_pointer_cast(d_counts.p) + b.ntiles, 0,
                                    thrust::maximum<int32_t>());
  FEC_REQUIRE(b.max_tile_nodes <= 65535, "tile has too many nodes");
  b.d_tile_node_ptr.alloc(b.ntiles + 1);
  FEC_CUDA(cudaMemsetAsync(d_counts.p + b.ntiles, 0, sizeof(int32_t), s));
  thrust::exclusive_scan(pol, thrust::device_pointer_cast(d_counts.p), thrust::device_pointer_cast(d_counts.p) + b.ntiles + 1,
                         thrust::device_pointer_cast(b.d_tile_node_ptr.p));
  int32_t tot = 0;
  FEC_CUDA(cudaMemcpyAsync(&tot, b.d_tile_node_ptr.p + b.ntiles, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FEC_CUDA(cudaStreamSynchronize(s));
  b.d_tile_nodes.alloc(tot);
  b.d_inc_ptr.alloc((size_t)tot + 1);
  k_compact_tiles<<<b.ntiles, 128, 0, s>>>(d_tn_tmp.p, d_ip_tmp.p, b.d_tile_node_ptr.p, cap, nnpe, te, b.ntiles, b.d_tile_nodes.p, b.d_inc_ptr.p);
  const int32_t last = (int32_t)(ne * nnpe);
  FEC_CUDA(cudaMemcpyAsync(b.d_inc_ptr.p + tot, &last, sizeof(int32_t), cudaMemcpyHostToDevice, s));
  FEC_CUDA(cudaGetLastError());
  FEC_CUDA(cudaStreamSynchronize(s));
  h->launches += 6;
}

// build_adjacency on the device (node adjacency of the scatter connectivity + the position bytes)
void build_adjacency_gpu(fecb200_handle* h) {
  PhaseTimer _pt("build_adjacency (gpu)");
  const int64_t nn = h->nn;
  cudaStream_t s = h->stream;
  auto pol = thrust::cuda::par.on(s);
  FEC_REQUIRE((int)h->blocks.size() <= kMaxBlocks, "too many element blocks");
  BlockRefs R{};
  R.nblocks = (int)h->blocks.size();
  int64_t npairs = 0, nel = 0;
  for (int bi = 0; bi < R.nblocks; ++bi) {
    BlockPlan& b = h->blocks[bi];
    R.conn[bi] = b.d_sconn_perm.p ? b.d_sconn_perm.p : b.d_conn_perm.p;
    R.first[bi] = nel;
    R.nnpe[bi] = b.nnpe;
    nel += b.ne;
    npairs += b.ne * b.nnpe;
  }
  R.first[R.nblocks] = nel;
  FEC_REQUIRE(nel < (int64_t)INT32_MAX && npairs < (int64_t)INT32_MAX, "mesh too large for 32-bit incidence ids");
  DevBuf<int32_t> d_key, d_val, d_nptr, d_cnt;
  d_key.alloc(npairs); d_val.alloc(npairs);
  int64_t off = 0;
  for (int bi = 0; bi < R.nblocks; ++bi) {
    const int64_t n = h->blocks[bi].ne * h->blocks[bi].nnpe;
    if (n) k_incidence_pairs<<<grid_for(n), 256, 0, s>>>(R, bi, d_key.p, d_val.p, off);
    off += n;
  }
  thrust::sort_by_key(pol, thrust::device_pointer_cast(d_key.p), thrust::device_pointer_cast(d_key.p) + npairs,
                      thrust::device_pointer_cast(d_val.p));
  d_nptr.alloc(nn + 1);
  thrust::lower_bound(pol, thrust::device_pointer_cast(d_key.p), thrust::device_pointer_cast(d_key.p) + npairs,
                      thrust::counting_iterator<int32_t>(0), thrust::counting_iterator<int32_t>((int32_t)nn + 1),
                      thrust::device_pointer_cast(d_nptr.p));
  d_cnt.alloc(nn + 1);
  DevBuf<int> d_flag;
  d_flag.alloc(1);
  d_flag.zero(s);
  FEC_CUDA(cudaMemsetAsync(d_cnt.p + nn, 0, sizeof(int32_t), s));
  k_neighbours<false><<<grid_for(nn, 128), 128, 0, s>>>(R, d_nptr.p, d_val.p, nn, d_cnt.p, nullptr, nullptr, d_flag.p);
  h->d_adjptr.alloc(nn + 1);
  thrust::exclusive_scan(pol, thrust::device_pointer_cast(d_cnt.p), thrust::device_pointer_cast(d_cnt.p) + nn + 1,
                         thrust::device_pointer_cast(h->d_adjptr.p));
  // int32 overflow check of the total in 64 bits
  const int64_t tot64 = thrust::reduce(pol, thrust::device_pointer_cast(d_cnt.p), thrust::device_pointer_cast(d_cnt.p) + nn, (int64_t)0);
  FEC_REQUIRE(tot64 < (int64_t)INT32_MAX, "node adjacency exceeds int32 range");
  int flag = 0;
  FEC_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEC_CUDA(cudaStreamSynchronize(s));
  FEC_REQUIRE(!flag, "a node has more than 256 neighbours");
  h->d_adj.alloc((size_t)tot64);
  k_neighbours<true><<<grid_for(nn, 128), 128, 0, s>>>(R, d_nptr.p, d_val.p, nn, nullptr, h->d_adjptr.p, h->d_adj.p, d_flag.p);
  for (auto& b : h->blocks) {
    b.d_epos.alloc((size_t)b.ne * b.nnpe * b.nnpe);
    if (b.ne)
      k_epos<<<grid_for(b.ne * b.nnpe), 256, 0, s>>>(b.d_sconn_perm.p ? b.d_sconn_perm.p : b.d_conn_perm.p, b.nnpe, b.ne, h->d_adjptr.p,
                                                    h->d_adj.p, b.d_epos.p);
  }
  FEC_CUDA(cudaGetLastError());
  FEC_CUDA(cudaStreamSynchronize(s));
  h->adjptr.clear();   // host copies are fetched on demand (ensure_host_structure)
  h->adj.clear();
  h->host_structure_valid = false;
  h->launches += 5;
}

// the dof-level part of build_matrix_structure on the device; returns false when the device path does not apply
void build_matrix_offsets_gpu(fecb200_handle* h) {
  PhaseTimer _pt("build_matrix_offsets (gpu)");
  const int nf = h->nf;
  const int64_t nn = h->nn, ndof = h->ndof, n_owned = h->n_owned_nodes;
  cudaStream_t s = h->stream;
  auto pol = thrust::cuda::par.on(s);
  int64_t nadj = 0;
  {
    int32_t last = 0;
    FEC_CUDA(cudaMemcpyAsync(&last, h->d_adjptr.p + nn, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FEC_CUDA(cudaStreamSynchronize(s));
    nadj = last;
  }
  if ((int64_t)h->d_freemask.n != nn) h->d_freemask.alloc(nn);
  if ((int64_t)h->d_coloff.n != nadj) h->d_coloff.alloc(nadj);
  if ((int64_t)h->d_rowstart.n != ndof) h->d_rowstart.alloc(ndof);
  if ((int64_t)h->d_diagslot.n != ndof) h->d_diagslot.alloc(ndof);
  DevBuf<int32_t> d_rowlen;
  DevBuf<int64_t> d_vals, d_base;
  DevBuf<int> d_flag;
  d_rowlen.alloc(nn); d_vals.alloc(n_owned + 1); d_base.alloc(n_owned + 1); d_flag.alloc(1);
  d_flag.zero(s);
  FEC_CUDA(cudaMemsetAsync(d_vals.p, 0, (n_owned + 1) * sizeof(int64_t), s));
  k_freemask<<<grid_for(nn), 256, 0, s>>>(h->d_d2u.p, nf, h->opts.condensed != 0, nn, h->d_freemask.p);
  k_coloff<<<grid_for(nn), 256, 0, s>>>(h->d_adjptr.p, h->d_adj.p, h->d_freemask.p, nn, n_owned, h->d_coloff.p, d_rowlen.p, d_vals.p, d_flag.p);
  thrust::exclusive_scan(pol, thrust::device_pointer_cast(d_vals.p), thrust::device_pointer_cast(d_vals.p) + n_owned + 1,
                         thrust::device_pointer_cast(d_base.p));
  k_rowstart<<<grid_for(nn), 256, 0, s>>>(h->d_adjptr.p, h->d_adj.p, h->d_freemask.p, h->d_coloff.p, d_rowlen.p, d_base.p, nf, nn, n_owned,
                                         h->d_rowstart.p, h->d_diagslot.p);
  h->max_rowlen = thrust::reduce(pol, thrust::device_pointer_cast(d_rowlen.p), thrust::device_pointer_cast(d_rowlen.p) + nn, 0,
                                 thrust::maximum<int32_t>());
  // nmat = number of kept dofs of owned nodes
  h->nmat = thrust::transform_reduce(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(n_owned),
                                     PopMask{h->d_freemask.p}, (int64_t)0, thrust::plus<int64_t>());
  int64_t nnz = 0;
  int flag = 0;
  FEC_CUDA(cudaMemcpyAsync(&nnz, d_base.p + n_owned, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FEC_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEC_CUDA(cudaStreamSynchronize(s));
  FEC_CUDA(cudaGetLastError());
  FEC_REQUIRE(!flag, "row too long for 16-bit column offsets");
  h->nnz = nnz;
  h->rowstart_h.clear();
  h->freemask_h.clear();
  h->host_structure_valid = false;
  h->launches += 3;
}

// host copies for export_pattern (and anything else that walks the pattern on the host)
void ensure_host_structure(fecb200_handle* h) {
  if (h->host_structure_valid) return;
  PhaseTimer _pt("host copies of the pattern");
  cudaStream_t s = h->stream;
  if (h->adjptr.empty() && h->d_adjptr.p) {
    h->adjptr.resize(h->nn + 1);
    FEC_CUDA(cudaMemcpyAsync(h->adjptr.data(), h->d_adjptr.p, (h->nn + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FEC_CUDA(cudaStreamSynchronize(s));
    h->adj.resize(h->adjptr[h->nn]);
    FEC_CUDA(cudaMemcpyAsync(h->adj.data(), h->d_adj.p, h->adj.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  if (h->rowstart_h.empty() && h->d_rowstart.p && h->matrix_ready) {
    h->rowstart_h.resize(h->ndof);
    h->freemask_h.resize(h->nn);
    FEC_CUDA(cudaMemcpyAsync(h->rowstart_h.data(), h->d_rowstart.p, h->ndof * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    FEC_CUDA(cudaMemcpyAsync(h->freemask_h.data(), h->d_freemask.p, h->nn, cudaMemcpyDeviceToHost, s));
  }
  FEC_CUDA(cudaStreamSynchronize(s));
  h->host_structure_valid = true;
}

}  // namespace fec
