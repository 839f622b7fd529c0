// common.cuh -- internal declarations of libfecb200 (not part of the C ABI).
#pragma once
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include "../../include/fecb200.h"

namespace fec {

extern thread_local std::string g_last_error;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define FEC_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      char _buf[512];                                                                      \
      snprintf(_buf, sizeof _buf, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),      \
               __FILE__, __LINE__, cudaGetErrorString(_e));                                \
      throw fec::Error(_buf);                                                              \
    }                                                                                      \
  } while (0)

#define FEC_REQUIRE(cond, msg)                                                             \
  do {                                                                                     \
    if (!(cond)) throw fec::Error(std::string("fecb200: ") + (msg));                       \
  } while (0)

// vmm.cu: compressible allocations through the driver's virtual-memory API (nullptr = not available)
void* vmm_alloc_compressible(int device, size_t bytes, size_t* mapped, unsigned long long* handle);
void vmm_free(void* p, size_t mapped, unsigned long long handle);

template <class T>
struct DevBuf {  // owning device array
  T* p = nullptr;
  size_t n = 0;
  size_t vmm_bytes = 0;              // != 0: p came from vmm_alloc_compressible
  unsigned long long vmm_handle = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), vmm_bytes(o.vmm_bytes), vmm_handle(o.vmm_handle) { o.p = nullptr; o.n = 0; o.vmm_bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; vmm_bytes = o.vmm_bytes; vmm_handle = o.vmm_handle;
      o.p = nullptr; o.n = 0; o.vmm_bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) { if (vmm_bytes) vmm_free(p, vmm_bytes, vmm_handle); else cudaFree(p); }
    p = nullptr; n = 0; vmm_bytes = 0;
  }
  // a compressible allocation when the driver offers one (see vmm.cu; FECB200_COMPRESS=0 opts out), cudaMalloc otherwise
  void alloc_compressible(size_t count, int device) {
    release();
    const char* env = getenv("FECB200_COMPRESS");
    if (count && !(env && env[0] == '0')) {
      void* q = vmm_alloc_compressible(device, count * sizeof(T), &vmm_bytes, &vmm_handle);
      if (q) { p = static_cast<T*>(q); n = count; return; }
      vmm_bytes = 0;
    }
    alloc(count);
  }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) FEC_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void upload(const std::vector<T>& v, cudaStream_t s) {
    alloc(v.size());
    if (!v.empty()) {
      FEC_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
      FEC_CUDA(cudaStreamSynchronize(s));
    }
  }
  void zero(cudaStream_t s) { if (n) FEC_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// FECB200_VERBOSE=1: wall time of the set-up phases on stderr
struct PhaseTimer {
  const char* name;
  std::chrono::steady_clock::time_point t0;
  bool on;
  explicit PhaseTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("FECB200_VERBOSE") != nullptr) {}
  ~PhaseTimer() {
    if (on) fprintf(stderr, "[fecb200] %-28s %.3f s\n", name, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};

// tuning knobs of the vector kernels (overridable at build time for sweeps)
#ifndef FEC_TE
#define FEC_TE 128
#endif
#ifndef FEC_MINB3
#define FEC_MINB3 2
#endif
#ifndef FEC_MINB1
#define FEC_MINB1 2
#endif
constexpr int kTE = FEC_TE;        // elements per tile = threads per CTA of the vector kernels
constexpr int kMinB3 = FEC_MINB3;  // __launch_bounds__ min CTAs/SM for NF = 3 vector kernels
constexpr int kMinB1 = FEC_MINB1;  // ... for NF <= 2
constexpr int kMaxProps = 8;
constexpr int kMaxPeers = 8;

// Ghost-node scatter over NVLink peer memory (fecb200_peer_attach): contributions to a node this rank does not
// own are added straight into the OWNER's field with system-scope REDs, instead of pack -> NCCL -> unpack.
struct PeerScatter {
  int64_t n_owned;              // local nodes >= n_owned are ghosts; < 0 disables the remote path
  const int32_t* ghost_peer;    // [nn - n_owned] index into base[] (or -1: ghost never touched by owned elements)
  const int32_t* ghost_node;    // [nn - n_owned] 0-based local node id on the owner
  double* base[kMaxPeers];      // peer-mapped residual fields (cudaIpcOpenMemHandle)
};

__device__ __forceinline__ void scatter_add(const PeerScatter& ps, double* local, int64_t n, int nf, int d, double v) {
  if (ps.n_owned >= 0 && n >= ps.n_owned) {
    const int64_t g = n - ps.n_owned;
    const int pr = ps.ghost_peer[g];
    if (pr >= 0) {
      double* dst = ps.base[pr] + (int64_t)ps.ghost_node[g] * nf + d;
      asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
    }
  } else {
    atomicAdd(&local[n * nf + d], v);
  }
}
// In-kernel zero-fill of the idle CSR value buffer (fecb200_set_matrix_double_buffer).  CTA i of a matrix kernel
// clears 16-byte units [i*chunk16, min((i+1)*chunk16, total16)) of p: one thread queues bulk stores (TMA, async
// proxy) from a shared-memory zero page; they drain to HBM underneath the element kernel without touching the
// LSU / L1 data pipe its REDs are bound by.
struct ZeroFill { double* p; int64_t total16; int32_t chunk16; };
struct EmetaOrder { int node[16]; };   // order of the per-element scatter record entries (k_build_emeta)
#ifndef FEC_ZPAGE
#define FEC_ZPAGE 1024
#endif
constexpr int kZeroPageBytes = FEC_ZPAGE;

// every thread of the CTA calls this; zp = kZeroPageBytes of shared memory.  Only warp 0 works: it clears the page,
// then its lane 0 queues the CTA's stores (issuing from every warp measured slower: 19.13 vs 18.80 ms at 192^3).
__device__ __forceinline__ void zero_fill_begin(const ZeroFill& z, double* zp) {
  if (z.p == nullptr || threadIdx.x >= 32) return;
  for (int i = threadIdx.x; i < kZeroPageBytes / 8; i += 32) zp[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    const int64_t beg = (int64_t)blockIdx.x * z.chunk16;
    int64_t rem = (z.total16 - beg < z.chunk16 ? z.total16 - beg : (int64_t)z.chunk16) * 16;
    char* g = reinterpret_cast<char*>(z.p) + beg * 16;
    const unsigned zs = (unsigned)__cvta_generic_to_shared(zp);
    for (; rem > 0; rem -= kZeroPageBytes, g += kZeroPageBytes) {
      const unsigned nb = rem < kZeroPageBytes ? (unsigned)rem : (unsigned)kZeroPageBytes;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(zs), "r"(nb) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// thread 0 calls this before it exits: the zero page must outlive the bulk stores' shared-memory reads
__device__ __forceinline__ void zero_fill_end(const ZeroFill& z) {
  if (z.p != nullptr && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

constexpr int kMaxNQ = 27;  // runtime-NQ kernels (e.g. 3-point GLL on hex8)

// One element block: FunctionSpace block + ReferenceFE tables + physics (host side of the plan)
struct BlockPlan {
  int elem_type = 0, nnpe = 0, nd = 0, nq = 0, physics = 0, nprops = 0, nstate = 0;
  int64_t ne = 0;
  std::vector<double> N, dN, w, props;
  // Walsh form of the HEX8 kernels: set by detect_walsh() when the dN table is the trilinear one on a
  // symmetric 2-point rule per axis, in ANY node / point numbering.  Sign index i: bit k set <=> +1 on axis k.
  bool walsh = false;
  double walsh_c = 0.0;                 // |xi| of the rule
  int node_of_sign[8] = {0, 1, 2, 3, 4, 5, 6, 7};    // sign index -> local node
  int point_of_sign[8] = {0, 1, 2, 3, 4, 5, 6, 7};   // sign index -> quadrature point
  int sign_of_point[8] = {0, 1, 2, 3, 4, 5, 6, 7};   // quadrature point -> sign index
  std::vector<int32_t> conn0;  // [ne*nnpe] 0-based node ids, caller's element order
  std::vector<int32_t> sconn0; // scatter connectivity: periodic side-b nodes folded into side a (empty = conn0)
  // tiling (vector kernels): elements permuted into locality-ordered tiles of `te` elements
  bool halo = false;  // neighbour-owned elements (partitioned runs): matrix assembly only
  int te = 0, ntiles = 0, max_tile_nodes = 0;
  std::vector<int32_t> perm;   // tile order -> original element index
  DevBuf<int32_t> d_perm;
  DevBuf<int32_t> d_tile_node_ptr, d_tile_nodes, d_inc_ptr;
  DevBuf<uint16_t> d_lconn, d_inc;
  DevBuf<int32_t> d_conn_perm;  // [ne*nnpe] global node ids in tile order (matrix kernels)
  DevBuf<int32_t> d_sconn_perm; // folded twin of d_conn_perm (only with periodic BCs)
  DevBuf<uint8_t> d_epos;       // [ne*nnpe*nnpe] position of node a in the adjacency row of node b
  DevBuf<unsigned char> d_emeta;  // [ne * emeta_rec] per-element scatter records of k_mat2 (kernel_mat2.cuh)
  size_t emeta_rec = 0;
  bool emeta_trash_rows = true;  // which flavour of scatter record d_emeta holds (k_build_emeta)
  bool emeta_sign_order = false; // records written in the sign order of the Walsh form (node_of_sign) instead of local order
  DevBuf<double> d_state_old, d_state_new;  // [(s*nq+q)*ne + e_tile_order]
  DevBuf<double> d_source;                  // [q*ne + e_tile_order]
  DevBuf<double> d_scalar;                  // [q*ne + e_tile_order] assemble_scalar! storage (allocated on first use)
  DevBuf<double> d_body_force;              // [NF, NQ, NE] body-force values, caller's element order (Sources.jl:38-46)
  DevBuf<double> d_tab;                     // N [nq*nnpe], dN [nq*nnpe*nd], w [nq] for the run-time-shaped load kernel
};

// One NeumannBCContainer (src/bcs/NeumannBCs.jl:29-48): the sides of a side set with their surface tables.
struct SurfaceLoad {
  int64_t nsides = 0;
  int nnps = 0, nqs = 0;
  DevBuf<int32_t> nodes;  // [nsides*nnps] 0-based node ids of each side (surface_connectivity)
  DevBuf<double> tab;     // Ns [nqs*nnps], dNs [nqs*nnps*(nd-1)], ws [nqs]
  DevBuf<double> vals;    // [NF, nqs, nsides] = Matrix{SVector{NF}}(nqs, nsides)
};

// One RobinBCContainer (src/bcs/RobinBCs.jl:29-48): side-set geometry like a Neumann BC plus the flux law at the
// surface quadrature points in affine form  vals(q, e) = g0(q, e) + D(q, e) u_q  (dvalsdu = D).
struct RobinLoad {
  SurfaceLoad geo;        // nodes, tables (vals unused)
  DevBuf<double> g0;      // [NF, nqs, nsides]
  DevBuf<double> D;       // [NF, NF, nqs, nsides]: D[di + NF*dj] = d vals_di / d u_dj  (column-major SMatrix{NF,NF})
};

}  // namespace fec

struct fecb200_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  fecb200_opts opts{};
  int nd = 0, nf = 0;
  int64_t nn = 0, ndof = 0;
  int64_t n_owned_nodes = 0;  // == nn unless fecb200_partition_setup was called
  std::vector<fec::BlockPlan> blocks;
  double t = 0.0, dt = 0.0;

  // full-length nodal fields (H1Field data: [(n)*NF + d])
  fec::DevBuf<double> d_X, d_U, d_V, d_R, d_Av;
  // unknown-length / staging vectors
  fec::DevBuf<double> d_Uu, d_Vu, d_out;

  // DofManager (host copies are 1-based Int64 like the reference)
  std::vector<int64_t> dirichlet_dofs, unknown_dofs, dof_to_unknown, per_a, per_b, b2a_unknown;
  int64_t n_unknowns = 0;
  fec::DevBuf<int32_t> d_unknown_dofs;  // 0-based dof ids
  fec::DevBuf<int32_t> d_d2u;           // dof -> index into Uu (or -1)
  fec::DevBuf<uint32_t> d_adjx;         // per adjacency entry: kept-dof mask << 29 | Uu index of the column node's first kept dof (SpMV)
  bool adjx_ok = false;                 // d_adjx matches the current dof maps (reset by build_ecol)
  fec::DevBuf<double> d_constraint;     // 1.0 at Dirichlet dofs
  // Dirichlet / periodic values
  int64_t n_bc = 0, n_per = 0;
  fec::DevBuf<int32_t> d_bc_dofs, d_per_a, d_per_b;
  fec::DevBuf<double> d_bc_vals, d_per_vals;

  // node adjacency and the (block-compressed) CSR structure
  std::vector<int32_t> adjptr, adj;  // host
  fec::DevBuf<int32_t> d_adjptr, d_adj;
  bool host_structure_valid = true;  // host copies of adjacency / row starts match the device (plan_gpu.cu fetches them lazily)
  bool matrix_ready = false;
  bool adj_folded = false;           // the node adjacency was built on the periodic-folded connectivity
  bool matrix_dirty = false;         // DOF maps changed since the CSR structure was built (built lazily after create)
  int64_t nmat = 0, nnz = 0;
  int32_t max_rowlen = 0;            // longest node row (kept dofs), bounds the column offsets
  std::vector<int64_t> rowstart_h;   // per dof, -1 if the row is eliminated
  std::vector<uint8_t> freemask_h;   // per node
  fec::DevBuf<uint16_t> d_coloff;    // per adjacency entry: kept dofs before this neighbour in the row
  fec::DevBuf<uint8_t> d_freemask;
  fec::DevBuf<int64_t> d_rowstart, d_diagslot;
  fec::DevBuf<double> d_nz_stiff, d_nz_mass, d_scratch;
  fec::DevBuf<double> d_nz_stiff_alt;  // fecb200_set_matrix_double_buffer: cleared by the kernel that fills the other one
  bool double_buffer = false, alt_clean = false;
  bool stiff_adjusted = false, mass_adjusted = false;

  // external loads (loads.cu): U-independent, so they are integrated once per value update into a cached nodal
  // vector and every assemble_vector_neumann_bc! / assemble_vector_source! is one streaming add
  std::vector<fec::SurfaceLoad> surface_loads;
  std::vector<fec::RobinLoad> robin_loads;
  fec::DevBuf<double> d_F_neumann, d_F_source;
  bool neumann_dirty = false, source_dirty = false;
  bool has_neumann() const { return !surface_loads.empty(); }
  bool has_source() const { for (auto& b : blocks) if (b.d_body_force.p) return true; return false; }

  // peer-memory halo (fecb200_peer_attach)
  bool peer_enabled = false;
  int peer_field = 0;
  std::vector<void*> peer_opened;          // pointers returned by cudaIpcOpenMemHandle (closed in destroy)
  fec::DevBuf<int32_t> d_ghost_peer, d_ghost_node;
  fec::PeerScatter peer{-1, nullptr, nullptr, {nullptr}};

  // halo exchange
  int n_neighbors = 0;
  std::vector<int32_t> neighbor_ranks;
  std::vector<int64_t> send_ptr, recv_ptr;
  fec::DevBuf<int32_t> d_send_nodes, d_recv_nodes;
  fec::DevBuf<double> d_sendbuf, d_recvbuf;
  // owner -> ghost update lists (fecb200_ghost_setup): EVERY ghost node, including the far nodes of halo elements that
  // no owned element touches (they are columns of owned Jacobian rows, but never receive residual contributions)
  bool ghost_lists = false;
  std::vector<int32_t> g_ranks;
  std::vector<int64_t> g_own_ptr, g_ghost_ptr;
  fec::DevBuf<int32_t> d_g_own_nodes, d_g_ghost_nodes;

  // collective plane (comm.cu): NCCL communicator of this rank, created by fecb200_comm_init
  void* comm = nullptr;
  int comm_rank = 0, comm_nranks = 1;
  fec::DevBuf<float> d_bar;

  // opt-in asynchronous host copies (fecb200_set_async): H2D / D2H run on their own streams, ordered with events
  bool async_copies = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_in_consumed = nullptr, ev_prod = nullptr, ev_d2h = nullptr;
  bool h2d_pending = false, d2h_pending = false, in_consumed_valid = false;

  // instrumentation
  int64_t launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;

  // CG workspace
  fec::DevBuf<double> d_cg_r, d_cg_p, d_cg_Ap, d_cg_x, d_red;
};

namespace fec {

// CSR value buffers: nnz values + the trash region of the branch-free RED streams (4096 hashed row starts, each
// followed by up to one row of column offsets), even length (16-byte units)
inline size_t nz_alloc_len(const fecb200_handle* h) { return ((size_t)h->nnz + 4096 + (size_t)h->max_rowlen + 8 + 1) & ~(size_t)1; }

// plan_gpu.cu: the same builders on the device (default; FECB200_HOST_PLAN=1 keeps the OpenMP versions of plan.cu)
bool use_gpu_plan();
void build_block_tiles_gpu(fecb200_handle* h, BlockPlan& b);
void build_adjacency_gpu(fecb200_handle* h);
void build_matrix_offsets_gpu(fecb200_handle* h);
void ensure_host_structure(fecb200_handle* h);

// Recognises the trilinear HEX8 table on a symmetric 2-point rule per axis in any node / point numbering and fills the
// block's sign maps: s_a^k = sign(sum_q dN[q][a][k]) (the sum is exactly s_a^k), c sigma_q^k' = sum_a s_a^k s_a^k' dN[q][a][k]
// for k' != k.  Every entry is then checked against dN = s_a^k/8 prod_{k' != k} (1 + c s_a^k' sigma_q^k'); weights are free
// (folded into JxW).  Anything else keeps the plain quadrature loop.
inline void detect_walsh(BlockPlan& b) {
  b.walsh = false;
  if (b.elem_type != FECB200_HEX8 || b.nq != 8 || b.nnpe != 8 || b.dN.size() != 8u * 8u * 3u) return;
  auto dn = [&](int q, int a, int k) { return b.dN[((size_t)q * 8 + a) * 3 + k]; };
  int sa[8][3], sq[8][3];
  for (int a = 0; a < 8; ++a)
    for (int k = 0; k < 3; ++k) {
      double s = 0.0;
      for (int q = 0; q < 8; ++q) s += dn(q, a, k);
      if (std::fabs(std::fabs(s) - 1.0) > 1e-12) return;
      sa[a][k] = s > 0 ? 1 : -1;
    }
  double c = -1.0;
  for (int q = 0; q < 8; ++q)
    for (int kp = 0; kp < 3; ++kp) {
      const int k = (kp + 1) % 3;
      double s = 0.0;
      for (int a = 0; a < 8; ++a) s += sa[a][k] * sa[a][kp] * dn(q, a, k);
      if (c < 0.0) c = std::fabs(s);
      if (std::fabs(std::fabs(s) - c) > 1e-12 || !(c > 1e-3 && c <= 1.0 + 1e-12)) return;
      sq[q][kp] = s > 0 ? 1 : -1;
    }
  int nos[8], pos[8], seen_n = 0, seen_q = 0;
  for (int a = 0; a < 8; ++a) { const int i = (sa[a][0] > 0) | ((sa[a][1] > 0) << 1) | ((sa[a][2] > 0) << 2); nos[i] = a; seen_n |= 1 << i; }
  for (int q = 0; q < 8; ++q) { const int i = (sq[q][0] > 0) | ((sq[q][1] > 0) << 1) | ((sq[q][2] > 0) << 2); pos[i] = q; seen_q |= 1 << i; }
  if (seen_n != 0xFF || seen_q != 0xFF) return;
  for (int q = 0; q < 8; ++q)
    for (int a = 0; a < 8; ++a)
      for (int k = 0; k < 3; ++k) {
        double v = sa[a][k] / 8.0;
        for (int kp = 0; kp < 3; ++kp)
          if (kp != k) v *= 1.0 + c * sa[a][kp] * sq[q][kp];
        if (std::fabs(v - dn(q, a, k)) > 1e-13) return;
      }
  b.walsh = true;
  b.walsh_c = c;
  for (int i = 0; i < 8; ++i) { b.node_of_sign[i] = nos[i]; b.point_of_sign[i] = pos[i]; b.sign_of_point[pos[i]] = i; }
}

// plan.cu
void build_block_tiles(fecb200_handle* h, BlockPlan& b, const double* coords);
void build_adjacency(fecb200_handle* h);
void build_dof_structures(fecb200_handle* h);
void build_matrix_structure(fecb200_handle* h);
void ensure_matrix_structure(fecb200_handle* h);
void export_pattern(fecb200_handle* h, int64_t* ptr, int64_t* idx);

// dispatch (one translation unit per element family)
enum { MODE_RESIDUAL = 0, MODE_ACTION_STIFFNESS = 1, MODE_ACTION_MASS = 2, MODE_LUMPED_MASS = 3, MODE_DIAG_MASS = 4, MODE_DIAG_STIFFNESS = 5 };
struct VecLaunch { const double* U; const double* V; double* out; int mode; };
struct MatLaunch {
  const double* U; double* nz; int kind;
  double* R = nullptr;                     // != null: fused residual
  double* zf = nullptr; int64_t zf_n = 0;  // != null: the kernel also clears zf[0 .. zf_n) (double-buffered CSR values)
};
inline ZeroFill make_zero_fill(const MatLaunch& a, int grid) {
  ZeroFill z{nullptr, 0, 0};
  if (a.zf && a.zf_n > 0 && grid > 0) {
    if (a.zf_n % 2 != 0 || (reinterpret_cast<uintptr_t>(a.zf) & 15) != 0) throw Error("fecb200: zero-fill range must be 16-byte aligned");
    z.p = a.zf;
    z.total16 = a.zf_n / 2;
    z.chunk16 = (int32_t)((z.total16 + grid - 1) / grid);
  }
  return z;
}
bool matrix_kernel_fuses_residual(fecb200_handle* h, const BlockPlan& b);
void launch_vector(fecb200_handle* h, BlockPlan& b, const VecLaunch& a);
void launch_matrix(fecb200_handle* h, BlockPlan& b, const MatLaunch& a);
void launch_scalar(fecb200_handle* h, BlockPlan& b, const double* U);
void launch_scalar_quad_tri(fecb200_handle* h, BlockPlan& b, const double* U);
void launch_scalar_hex8(fecb200_handle* h, BlockPlan& b, const double* U);
void launch_scalar_tet(fecb200_handle* h, BlockPlan& b, const double* U);
void launch_vector_quad_tri(fecb200_handle* h, BlockPlan& b, const VecLaunch& a);
void launch_matrix_quad_tri(fecb200_handle* h, BlockPlan& b, const MatLaunch& a);
void launch_vector_hex8(fecb200_handle* h, BlockPlan& b, const VecLaunch& a);
void launch_matrix_hex8(fecb200_handle* h, BlockPlan& b, const MatLaunch& a);
void launch_vector_tet(fecb200_handle* h, BlockPlan& b, const VecLaunch& a);
void launch_matrix_tet(fecb200_handle* h, BlockPlan& b, const MatLaunch& a);

// aux.cu
void k_update_field(fecb200_handle* h, double* field, const double* Uu, bool with_bcs);
void k_extract_unknowns(fecb200_handle* h, const double* field, double* out);
void k_residual_accessor(fecb200_handle* h, double* out);
void k_hvp_accessor(fecb200_handle* h, const double* v, double* out);
void k_adjust_matrix(fecb200_handle* h, double* nz);
void k_permute_state_in(fecb200_handle* h, BlockPlan& b, const double* src_dev, double* dst);
void k_permute_state_out(fecb200_handle* h, BlockPlan& b, const double* src, double* dst_dev);
void k_permute_source_in(fecb200_handle* h, BlockPlan& b, const double* src_dev, double* dst);
void k_permute_scalar_out(fecb200_handle* h, BlockPlan& b, const double* src, double* dst_dev);
void k_zero_bc_slots(fecb200_handle* h, double* field);
void build_ecol(fecb200_handle* h);
void spmv(fecb200_handle* h, const double* nz, const double* x, double* y);
double dot(fecb200_handle* h, const double* a, const double* b, int64_t n);
void axpy(fecb200_handle* h, double alpha, const double* x, double* y, int64_t n);
void xpay(fecb200_handle* h, const double* x, double beta, double* y, int64_t n);  // y = x + beta*y
void fill_indexed(fecb200_handle* h, double* field, const int32_t* idx, double v, int64_t n);  // field[idx[i]] = v
void halo_pack(fecb200_handle* h, const double* field, double* buf);
void halo_unpack_add(fecb200_handle* h, double* field, const double* buf);

// comm.cu
bool comm_active(const fecb200_handle* h);
void comm_allreduce_sum(fecb200_handle* h, double* dev, int n);
void comm_barrier(fecb200_handle* h);
void comm_halo_sum_field(fecb200_handle* h, double* field);
void comm_halo_update_field(fecb200_handle* h, double* field);
void comm_halo_update_unknowns(fecb200_handle* h, double* v);
int64_t owned_len(const fecb200_handle* h);
void comm_release(fecb200_handle* h);

// loads.cu
void add_neumann_loads(fecb200_handle* h, double* field);   // field += int_Gamma N g      (WeaklyEnforcedBCs.jl:61-83)
void add_source_loads(fecb200_handle* h, double* field);    // field += -int_Omega N b    (Source.jl:44-63)

inline bool is_device_ptr(const void* p) {
  cudaPointerAttributes a{};
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline bool is_pinned_host_ptr(const void* p) {
  cudaPointerAttributes a{};
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace fec
