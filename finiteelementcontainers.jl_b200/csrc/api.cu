// api.cu -- the extern "C" entry points declared in include/fecb200.h.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <memory>

namespace fec {
thread_local std::string g_last_error;
int metis_mesh_dual(int64_t, int64_t, const int64_t*, const int64_t*, int64_t, int64_t, int64_t*, int64_t*);
int metis_graph(int64_t, const int64_t*, const int64_t*, int64_t, int64_t*);

void launch_vector(fecb200_handle* h, BlockPlan& b, const VecLaunch& a) {
  switch (b.elem_type) {
    case FECB200_QUAD4: case FECB200_TRI3: launch_vector_quad_tri(h, b, a); break;
    case FECB200_HEX8: launch_vector_hex8(h, b, a); break;
    case FECB200_TET4: case FECB200_TET10: launch_vector_tet(h, b, a); break;
    default: throw Error("fecb200: unknown element type");
  }
}
void launch_matrix(fecb200_handle* h, BlockPlan& b, const MatLaunch& a) {
  switch (b.elem_type) {
    case FECB200_QUAD4: case FECB200_TRI3: launch_matrix_quad_tri(h, b, a); break;
    case FECB200_HEX8: launch_matrix_hex8(h, b, a); break;
    case FECB200_TET4: case FECB200_TET10: launch_matrix_tet(h, b, a); break;
    default: throw Error("fecb200: unknown element type");
  }
}

void launch_scalar(fecb200_handle* h, BlockPlan& b, const double* U) {
  switch (b.elem_type) {
    case FECB200_QUAD4: case FECB200_TRI3: launch_scalar_quad_tri(h, b, U); break;
    case FECB200_HEX8: launch_scalar_hex8(h, b, U); break;
    case FECB200_TET4: case FECB200_TET10: launch_scalar_tet(h, b, U); break;
    default: throw Error("fecb200: unknown element type");
  }
}

static int elem_nnpe(int t) {
  switch (t) {
    case FECB200_QUAD4: return 4; case FECB200_TRI3: return 3; case FECB200_HEX8: return 8;
    case FECB200_TET4: return 4; case FECB200_TET10: return 10; default: return -1;
  }
}
static int elem_nd(int t) { return (t == FECB200_QUAD4 || t == FECB200_TRI3) ? 2 : 3; }
static int phys_nstate(int p) { return p == FECB200_PHYS_J2_PLASTICITY ? 7 : 0; }

// copy `n` doubles from a host-or-device pointer into device staging
static const double* stage_in(fecb200_handle* h, const double* src, double* staging, int64_t n) {
  if (is_device_ptr(src)) return src;
  if (h->async_copies) {
    // H2D on its own stream: it overlaps whatever the main stream does until join_inputs() (e.g. the CSR memset)
    if (h->in_consumed_valid) FEC_CUDA(cudaStreamWaitEvent(h->s_h2d, h->ev_in_consumed, 0));  // staging is free again
    FEC_CUDA(cudaMemcpyAsync(staging, src, n * sizeof(double), cudaMemcpyHostToDevice, h->s_h2d));
    FEC_CUDA(cudaEventRecord(h->ev_h2d, h->s_h2d));
    h->h2d_pending = true;
    return staging;
  }
  FEC_CUDA(cudaMemcpyAsync(staging, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // "every call is complete on return" (fecb200.h): a pageable source is staged by the driver before the call
  // returns, a page-locked one is read by DMA later -- the caller may update Uu in place right after this call
  if (is_pinned_host_ptr(src)) FEC_CUDA(cudaStreamSynchronize(h->stream));
  return staging;
}
// main stream waits for pending H2D copies right before their first consumer
static void join_inputs(fecb200_handle* h) {
  if (h->h2d_pending) { FEC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_h2d, 0)); h->h2d_pending = false; }
}
// record that the staged inputs have been consumed (the next H2D may overwrite the staging buffers)
static void inputs_consumed(fecb200_handle* h) {
  if (h->async_copies) { FEC_CUDA(cudaEventRecord(h->ev_in_consumed, h->stream)); h->in_consumed_valid = true; }
}
// before the main stream overwrites d_out: a D2H copy of its previous content may still be in flight
static void wait_out_free(fecb200_handle* h) {
  if (h->d2h_pending) { FEC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_d2h, 0)); h->d2h_pending = false; }
}
static void copy_out(fecb200_handle* h, double* dst, const double* src_dev, int64_t n) {
  if (dst == src_dev) return;
  const bool dev = is_device_ptr(dst);
  if (!dev && h->async_copies) {
    FEC_CUDA(cudaEventRecord(h->ev_prod, h->stream));
    FEC_CUDA(cudaStreamWaitEvent(h->s_d2h, h->ev_prod, 0));
    FEC_CUDA(cudaMemcpyAsync(dst, src_dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->s_d2h));
    FEC_CUDA(cudaEventRecord(h->ev_d2h, h->s_d2h));
    h->d2h_pending = true;  // caller reads dst after fecb200_synchronize
    return;
  }
  FEC_CUDA(cudaMemcpyAsync(dst, src_dev, n * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                           h->stream));
  if (!dev) FEC_CUDA(cudaStreamSynchronize(h->stream));
}
static int64_t len_Uu(fecb200_handle* h) { return h->opts.condensed ? h->ndof : h->n_unknowns; }
static double* field_ptr(fecb200_handle* h, int which) {
  switch (which) {
    case FECB200_FIELD_U: return h->d_U.p;
    case FECB200_FIELD_RESIDUAL: return h->d_R.p;
    case FECB200_FIELD_ACTION: return h->d_Av.p;
    case FECB200_FIELD_V: return h->d_V.p;
    default: throw Error("fecb200: bad field selector");
  }
}

static void assemble_vector_impl(fecb200_handle* h, int mode, const double* Uu_dev, const double* Vu_dev,
                                 double* out_field) {
  // fill!(storage, 0) (Vector.jl:30).  In peer-scatter mode the field is zeroed right after it is read instead
  // (fecb200_residual): neighbours may already be adding into it when this call starts.
  if (!(h->peer_enabled && out_field == field_ptr(h, h->peer_field)))
    FEC_CUDA(cudaMemsetAsync(out_field, 0, h->ndof * sizeof(double), h->stream));
  join_inputs(h);
  k_update_field(h, h->d_U.p, Uu_dev, true);
  if (Vu_dev) k_update_field(h, h->d_V.p, Vu_dev, false);  // V's BC slots stay 0 (Parameters.jl:415-425)
  inputs_consumed(h);
  h->last_ms = 0.f;
  for (auto& b : h->blocks) {
    if (b.halo) continue;  // neighbour-owned elements: their contribution arrives through the halo exchange
    VecLaunch a{h->d_U.p, Vu_dev ? h->d_V.p : nullptr, out_field, mode};
    launch_vector(h, b, a);
  }
}

static double* nz_for_kind(fecb200_handle* h, int kind, bool alloc) {
  FEC_REQUIRE(!h->opts.matrix_free,
              "assemble_matrix! called on a matrix-free SparseMatrixAssembler.  Re-create the assembler with "
              "matrix_free=false to enable matrix assembly.");  // Matrix.jl:23-28
  ensure_matrix_structure(h);
  FEC_REQUIRE(h->matrix_ready, "matrix pattern not built");
  if (kind == FECB200_STIFFNESS) return h->d_nz_stiff.p;
  if (kind == FECB200_MASS) {
    if (!h->d_nz_mass.p && alloc) { h->d_nz_mass.alloc(nz_alloc_len(h)); h->d_nz_mass.zero(h->stream); }
    return h->d_nz_mass.p;
  }
  throw Error("fecb200: matrix kind must be FECB200_STIFFNESS or FECB200_MASS");
}

static void ensure_cg(fecb200_handle* h) {
  const int64_t n = len_Uu(h);
  if ((int64_t)h->d_cg_r.n < n) {
    h->d_cg_r.alloc(n); h->d_cg_p.alloc(n); h->d_cg_Ap.alloc(n); h->d_cg_x.alloc(n);
    h->d_cg_Ap.zero(h->stream);  // ghost rows of K p are never written on a partitioned handle
  }
}

// operator application for CG: y = K x (assembled) or the matrix-free action at the current U
// Partitioned handles (fecb200_comm_init): x is made consistent first (owner -> ghost copies, PVector consistent!,
// ext/PartitionedArraysExt.jl:449-459); the assembled operator then needs no further exchange (every owned row is
// complete), the matrix-free one sums the ghost contributions of the action into their owners (:469-481).
static void apply_operator(fecb200_handle* h, bool matrix_free, double* x, double* y) {
  const bool dist = comm_active(h) && h->n_neighbors > 0;
  if (dist) comm_halo_update_unknowns(h, x);
  if (!matrix_free) {
    spmv(h, h->d_nz_stiff.p, x, y);
  } else {
    FEC_REQUIRE(!(dist && h->peer_enabled && h->peer_field == FECB200_FIELD_ACTION) , "matrix-free CG over the peer-memory action halo is not wired; use the NCCL halo");
    FEC_CUDA(cudaMemsetAsync(h->d_Av.p, 0, h->ndof * sizeof(double), h->stream));
    k_update_field(h, h->d_V.p, x, false);
    for (auto& b : h->blocks) {
      if (b.halo) continue;
      VecLaunch a{h->d_U.p, h->d_V.p, h->d_Av.p, MODE_ACTION_STIFFNESS};
      launch_vector(h, b, a);
    }
    if (dist) comm_halo_sum_field(h, h->d_Av.p);
    k_hvp_accessor(h, x, y);
  }
}

static void cg_impl(fecb200_handle* h, const double* b_dev, double* x_dev, double atol, double rtol, int64_t itmax,
                    bool matrix_free, int64_t* iters_out, double* rnorm_out) {
  // Krylov.jl cg with x0 = 0: r = b, p = r; stop when ||r|| <= atol + rtol ||r0||
  // Partitioned handles: vectors keep the rank-local Uu layout (owned entries first, then ghosts); dots run over the
  // OWNED prefix and are summed over the ranks (dot() all-reduces when a communicator is attached), so every rank
  // takes the same branches and the iteration count equals the serial solve's.
  const int64_t n = len_Uu(h);
  const int64_t nred = owned_len(h);
  ensure_cg(h);
  double* r = h->d_cg_r.p; double* p = h->d_cg_p.p; double* Ap = h->d_cg_Ap.p;
  FEC_CUDA(cudaMemsetAsync(x_dev, 0, n * sizeof(double), h->stream));
  FEC_CUDA(cudaMemcpyAsync(r, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  FEC_CUDA(cudaMemcpyAsync(p, b_dev, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  inputs_consumed(h);  // b may live in the H2D staging buffer (fecb200_cg_solve): free for the next async copy
  double gamma = dot(h, r, r, nred);
  const double rn0 = std::sqrt(gamma);
  const double tol = atol + rtol * rn0;
  int64_t it = 0;
  while (rn0 > 0 && std::sqrt(gamma) > tol && it < itmax) {
    apply_operator(h, matrix_free, p, Ap);
    const double pAp = dot(h, p, Ap, nred);
    const double alpha = gamma / pAp;
    axpy(h, alpha, p, x_dev, n);
    axpy(h, -alpha, Ap, r, n);
    const double gnew = dot(h, r, r, nred);
    xpay(h, r, gnew / gamma, p, n);  // p = r + beta p
    gamma = gnew;
    ++it;
  }
  if (comm_active(h) && h->n_neighbors > 0) comm_halo_update_unknowns(h, x_dev);  // ghost entries of the solution
  if (iters_out) *iters_out = it;
  if (rnorm_out) *rnorm_out = std::sqrt(gamma);
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

// Double-buffered CSR values (fecb200_set_matrix_double_buffer): the assembly writes into the buffer the PREVIOUS
// assembly's kernels cleared and clears the other one on the way (k_mat2, TMA bulk stores), so no stand-alone
// fill!(storage, 0) pass (Matrix.jl:39) runs in steady state.  Each launch clears a share proportional to its elements.
struct ZeroFillPlan {
  bool on = false;
  double* base = nullptr;
  int64_t total16 = 0, ne_total = 0, ne_done = 0;
  void next(const fec::BlockPlan& b, fec::MatLaunch& a) {
    if (!on) return;
    const int64_t beg = (int64_t)((__int128)total16 * ne_done / ne_total);
    ne_done += b.ne;
    const int64_t end = (int64_t)((__int128)total16 * ne_done / ne_total);
    if (end > beg) { a.zf = base + 2 * beg; a.zf_n = 2 * (end - beg); }
  }
};

static ZeroFillPlan begin_stiffness_fill(fecb200_handle* h) {
  ZeroFillPlan z;
  if (!h->double_buffer || !h->d_nz_stiff_alt.p) return z;
  for (auto& b : h->blocks) z.ne_total += b.ne;  // every matrix kernel clears its share (zero_fill_begin, common.cuh)
  if (z.ne_total == 0) return z;
  if (!h->alt_clean) { h->d_nz_stiff_alt.zero(h->stream); h->alt_clean = true; }
  std::swap(h->d_nz_stiff, h->d_nz_stiff_alt);  // d_nz_stiff is now the clean buffer; the kernels clear the other
  z.on = true;
  z.base = h->d_nz_stiff_alt.p;
  z.total16 = (int64_t)(nz_alloc_len(h) / 2);
  return z;
}

extern "C" {

static void assemble_vector_and_matrix_impl(fecb200_handle* h, const double* u);
static double* adjusted_values(fecb200_handle* h, int kind);

const char* fecb200_last_error(void) { return fec::g_last_error.c_str(); }
int fecb200_version(void) { return 100; }

int fecb200_create(const fecb200_mesh_desc* mesh, const fecb200_opts* opts, fecb200_handle** out) {
  FEC_API_BEGIN
  PhaseTimer _pt("fecb200_create (total)");
  FEC_REQUIRE(mesh && opts && out, "null argument");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw Error("fecb200: no CUDA device available -- libfecb200 has no CPU fallback");
  }
  FEC_REQUIRE(opts->device >= 0 && opts->device < ndev, "bad device ordinal");
  FEC_REQUIRE(mesh->nf >= 1 && mesh->nf <= 3, "NF must be 1..3");
  FEC_REQUIRE(mesh->ndim == 2 || mesh->ndim == 3, "ndim must be 2 or 3");
  FEC_REQUIRE(mesh->nblocks >= 1 && mesh->nblocks <= 16, "1..16 element blocks (MAX_BLOCKS, FunctionSpaces.jl:63)");
  FEC_REQUIRE(mesh->nnodes > 0 && mesh->nnodes * (int64_t)mesh->nf < (int64_t)INT32_MAX, "bad node count");
  FEC_REQUIRE(opts->matrix_type == FECB200_CSC || opts->matrix_type == FECB200_CSR,
              "Unsupported sparse matrix type. Only csc and csr are supported.");
  { PhaseTimer _t("cuda context"); FEC_CUDA(cudaSetDevice(opts->device)); FEC_CUDA(cudaFree(nullptr)); }
  std::unique_ptr<fecb200_handle> h(new fecb200_handle());
  h->device = opts->device;
  h->opts = *opts;
  h->nd = mesh->ndim; h->nf = mesh->nf; h->nn = mesh->nnodes; h->ndof = h->nn * h->nf;
  h->n_owned_nodes = h->nn;
  FEC_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  FEC_CUDA(cudaEventCreate(&h->ev0));
  FEC_CUDA(cudaEventCreate(&h->ev1));
  const int te = opts->tile_elems > 0 ? opts->tile_elems : kTE;
  FEC_REQUIRE(te == kTE, "tile_elems does not match the tile size this build was compiled for");
  {
    PhaseTimer _t("coords upload");   // first: the device tile builder bins element centroids
    h->d_X.alloc((size_t)h->nn * h->nd);
    FEC_CUDA(cudaMemcpyAsync(h->d_X.p, mesh->coords, (size_t)h->nn * h->nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FEC_CUDA(cudaStreamSynchronize(h->stream));
  }
  h->blocks.resize(mesh->nblocks);
  for (int bi = 0; bi < mesh->nblocks; ++bi) {
    const fecb200_block_desc& d = mesh->blocks[bi];
    BlockPlan& b = h->blocks[bi];
    FEC_REQUIRE(elem_nnpe(d.elem_type) == d.nnpe, "nnpe does not match the element type");
    FEC_REQUIRE(elem_nd(d.elem_type) == mesh->ndim, "element dimension does not match ndim");
    FEC_REQUIRE(d.nelem > 0 && d.nelem < (int64_t)INT32_MAX / 16, "bad element count");
    FEC_REQUIRE(d.nq >= 1 && d.nq <= kMaxNQ, "unsupported number of quadrature points");
    FEC_REQUIRE(d.nprops >= 0 && d.nprops <= kMaxProps, "too many properties");
    FEC_REQUIRE(d.nstate == phys_nstate(d.physics_id), "nstate does not match the physics");
    b.elem_type = d.elem_type; b.nnpe = d.nnpe; b.nd = mesh->ndim; b.nq = d.nq; b.physics = d.physics_id;
    b.nprops = d.nprops; b.nstate = d.nstate; b.ne = d.nelem; b.te = te;
    b.N.assign(d.N, d.N + (size_t)d.nq * d.nnpe);
    b.dN.assign(d.dN, d.dN + (size_t)d.nq * d.nnpe * mesh->ndim);
    b.w.assign(d.w, d.w + d.nq);
    detect_walsh(b);   // HEX8 tables of the trilinear / 2-point-rule kind take the Walsh form of the element kernels
    if (d.nprops) b.props.assign(d.props, d.props + d.nprops);
    {
      PhaseTimer _t("connectivity narrow + check");
      b.conn0.resize((size_t)d.nelem * d.nnpe);
      bool ok = true;
      const int64_t nc = (int64_t)b.conn0.size(), nnodes = mesh->nnodes;
#pragma omp parallel for schedule(static) reduction(&& : ok)
      for (int64_t i = 0; i < nc; ++i) {
        const int64_t n = d.conn[i];
        ok = ok && n >= 1 && n <= nnodes;
        b.conn0[i] = (int32_t)(n - 1);
      }
      FEC_REQUIRE(ok, "connectivity entry out of range (expects 1-based node ids)");
    }
    build_block_tiles(h.get(), b, mesh->coords);
    if (b.nstate) {
      b.d_state_old.alloc((size_t)b.nstate * b.nq * b.ne); b.d_state_old.zero(h->stream);
      b.d_state_new.alloc((size_t)b.nstate * b.nq * b.ne); b.d_state_new.zero(h->stream);
    }
  }
  h->d_U.alloc(h->ndof); h->d_U.zero(h->stream);
  h->d_V.alloc(h->ndof); h->d_V.zero(h->stream);
  h->d_R.alloc(h->ndof); h->d_R.zero(h->stream);
  h->d_Av.alloc(h->ndof); h->d_Av.zero(h->stream);
  if (!opts->matrix_free) build_adjacency(h.get());
  build_dof_structures(h.get());
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  *out = h.release();
  FEC_API_END
}

int fecb200_destroy(fecb200_handle* h) {
  FEC_API_BEGIN
  if (h) {
    cudaSetDevice(h->device);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    cudaStream_t s = h->own_stream ? h->stream : nullptr;
    cudaStreamSynchronize(h->stream);
    for (void* pp : h->peer_opened) cudaIpcCloseMemHandle(pp);
    comm_release(h);
    if (h->s_h2d) { cudaStreamSynchronize(h->s_h2d); cudaStreamDestroy(h->s_h2d); }
    if (h->s_d2h) { cudaStreamSynchronize(h->s_d2h); cudaStreamDestroy(h->s_d2h); }
    for (cudaEvent_t ev : {h->ev_h2d, h->ev_in_consumed, h->ev_prod, h->ev_d2h}) if (ev) cudaEventDestroy(ev);
    delete h;
    if (s) cudaStreamDestroy(s);
  }
  FEC_API_END
}

int fecb200_set_stream(fecb200_handle* h, void* cuda_stream) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (cuda_stream) { h->stream = (cudaStream_t)cuda_stream; h->own_stream = false; }
  else { FEC_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
  FEC_API_END
}

int fecb200_synchronize(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  if (h->s_h2d) FEC_CUDA(cudaStreamSynchronize(h->s_h2d));
  if (h->s_d2h) FEC_CUDA(cudaStreamSynchronize(h->s_d2h));
  h->d2h_pending = h->h2d_pending = false;
  FEC_API_END
}

int fecb200_set_async(fecb200_handle* h, int32_t on) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  if (on && !h->s_h2d) {
    FEC_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    FEC_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    FEC_CUDA(cudaEventCreateWithFlags(&h->ev_h2d, cudaEventDisableTiming));
    FEC_CUDA(cudaEventCreateWithFlags(&h->ev_in_consumed, cudaEventDisableTiming));
    FEC_CUDA(cudaEventCreateWithFlags(&h->ev_prod, cudaEventDisableTiming));
    FEC_CUDA(cudaEventCreateWithFlags(&h->ev_d2h, cudaEventDisableTiming));
  }
  if (!on && h->s_h2d) { FEC_CUDA(cudaStreamSynchronize(h->s_h2d)); FEC_CUDA(cudaStreamSynchronize(h->s_d2h)); }
  h->async_copies = on != 0;
  h->h2d_pending = h->d2h_pending = h->in_consumed_valid = false;
  FEC_API_END
}

int fecb200_update_dofs(fecb200_handle* h, const int64_t* dd, int64_t nd, const int64_t* pa, const int64_t* pb,
                        int64_t np) {
  FEC_API_BEGIN
  PhaseTimer _pt("fecb200_update_dofs (total)");
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  h->dirichlet_dofs.assign(dd, dd + nd);
  h->per_a.assign(pa, pa + np);
  h->per_b.assign(pb, pb + np);
  build_dof_structures(h);
  ensure_matrix_structure(h);  // eager here: configuration errors (e.g. periodic BCs + matrix assembly) surface at update_dofs!
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

int fecb200_sizes(fecb200_handle* h, int64_t* n_total, int64_t* n_unknowns, int64_t* lu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  if (n_total) *n_total = h->ndof;
  if (n_unknowns) *n_unknowns = h->n_unknowns;
  if (lu) *lu = len_Uu(h);
  FEC_API_END
}

int fecb200_dof_maps_copy(fecb200_handle* h, int64_t* unknown_dofs, int64_t* dof_to_unknown) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  if (unknown_dofs) std::copy(h->unknown_dofs.begin(), h->unknown_dofs.end(), unknown_dofs);
  if (dof_to_unknown) std::copy(h->dof_to_unknown.begin(), h->dof_to_unknown.end(), dof_to_unknown);
  FEC_API_END
}

int fecb200_pattern_sizes(fecb200_handle* h, int64_t* n, int64_t* nnz) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  ensure_matrix_structure(h);
  FEC_REQUIRE(h->matrix_ready, "no matrix pattern (matrix_free assembler)");
  if (n) *n = h->nmat;
  if (nnz) *nnz = h->nnz;
  FEC_API_END
}

int fecb200_pattern_copy(fecb200_handle* h, int64_t* ptr, int64_t* idx) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  export_pattern(h, ptr, idx);
  FEC_API_END
}

int fecb200_set_dirichlet_values(fecb200_handle* h, const int64_t* dofs, const double* vals, int64_t n) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  std::vector<int32_t> bd(n);
  for (int64_t i = 0; i < n; ++i) {
    FEC_REQUIRE(dofs[i] >= 1 && dofs[i] <= h->ndof, "dirichlet dof out of range");
    bd[i] = (int32_t)(dofs[i] - 1);
  }
  std::vector<double> bv(vals, vals + n);
  h->n_bc = n;
  h->d_bc_dofs.upload(bd, h->stream);
  h->d_bc_vals.upload(bv, h->stream);
  FEC_API_END
}

int fecb200_set_periodic_values(fecb200_handle* h, const double* vals, int64_t n) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_REQUIRE(!vals || n == h->n_per, "periodic value count does not match the resolved periodic pairs");
  std::vector<double> pv(h->n_per, 0.0);   // NULL = all jumps zero, whatever n says
  if (vals) pv.assign(vals, vals + n);
  h->d_per_vals.upload(pv, h->stream);
  FEC_API_END
}

int fecb200_set_time(fecb200_handle* h, double t, double dt) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  h->t = t; h->dt = dt;
  FEC_API_END
}

int fecb200_set_source_q(fecb200_handle* h, int32_t block, const double* fq) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && block >= 0 && block < (int)h->blocks.size(), "bad block index");
  FEC_CUDA(cudaSetDevice(h->device));
  BlockPlan& b = h->blocks[block];
  const int64_t n = b.ne * b.nq;
  if (!fq) { b.d_source.release(); return 0; }
  DevBuf<double> tmp;
  const double* src = fq;
  if (!is_device_ptr(fq)) {
    tmp.alloc(n);
    FEC_CUDA(cudaMemcpyAsync(tmp.p, fq, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    src = tmp.p;
  }
  if ((int64_t)b.d_source.n != n) b.d_source.alloc(n);
  k_permute_source_in(h, b, src, b.d_source.p);
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

int fecb200_state_set(fecb200_handle* h, int32_t block, int32_t which, const double* state) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && block >= 0 && block < (int)h->blocks.size(), "bad block index");
  FEC_CUDA(cudaSetDevice(h->device));
  BlockPlan& b = h->blocks[block];
  const int64_t n = b.ne * b.nq * b.nstate;
  if (!n) return 0;
  DevBuf<double> tmp;
  const double* src = state;
  if (!is_device_ptr(state)) {
    tmp.alloc(n);
    FEC_CUDA(cudaMemcpyAsync(tmp.p, state, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    src = tmp.p;
  }
  k_permute_state_in(h, b, src, which ? b.d_state_new.p : b.d_state_old.p);
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

int fecb200_state_get(fecb200_handle* h, int32_t block, int32_t which, double* state) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && block >= 0 && block < (int)h->blocks.size(), "bad block index");
  FEC_CUDA(cudaSetDevice(h->device));
  BlockPlan& b = h->blocks[block];
  const int64_t n = b.ne * b.nq * b.nstate;
  if (!n) return 0;
  const double* src = which ? b.d_state_new.p : b.d_state_old.p;
  if (is_device_ptr(state)) {
    k_permute_state_out(h, b, src, state);
    FEC_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    DevBuf<double> tmp;
    tmp.alloc(n);
    k_permute_state_out(h, b, src, tmp.p);
    FEC_CUDA(cudaMemcpyAsync(state, tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FEC_CUDA(cudaStreamSynchronize(h->stream));
  }
  FEC_API_END
}

int fecb200_state_swap(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  for (auto& b : h->blocks) {
    const size_t n = b.d_state_old.n;
    if (n) FEC_CUDA(cudaMemcpyAsync(b.d_state_old.p, b.d_state_new.p, n * sizeof(double), cudaMemcpyDeviceToDevice,
                                    h->stream));
  }
  FEC_API_END
}

int fecb200_assemble_vector(fecb200_handle* h, int32_t kind, const double* Uu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  int mode = -1;
  switch (kind) {
    case FECB200_RESIDUAL: mode = MODE_RESIDUAL; break;
    case FECB200_LUMPED_MASS: mode = MODE_LUMPED_MASS; break;
    case FECB200_DIAGONAL_STIFFNESS: mode = MODE_DIAG_STIFFNESS; break;
    case FECB200_DIAGONAL_MASS: mode = MODE_DIAG_MASS; break;
  }
  FEC_REQUIRE(mode >= 0, "assemble_vector: kind must be FECB200_RESIDUAL, FECB200_LUMPED_MASS or FECB200_DIAGONAL_*");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  assemble_vector_impl(h, mode, u, nullptr, h->d_R.p);
  FEC_API_END
}

int fecb200_assemble_scalar(fecb200_handle* h, const double* Uu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  join_inputs(h);
  k_update_field(h, h->d_U.p, u, true);
  inputs_consumed(h);
  h->last_ms = 0.f;
  for (auto& b : h->blocks) {
    if (b.halo) continue;
    launch_scalar(h, b, h->d_U.p);
  }
  FEC_API_END
}

int fecb200_scalar_values(fecb200_handle* h, int32_t block, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_REQUIRE(block >= 0 && block < (int32_t)h->blocks.size(), "bad block index");
  FEC_CUDA(cudaSetDevice(h->device));
  BlockPlan& b = h->blocks[block];
  FEC_REQUIRE(b.d_scalar.p, "assemble_scalar! has not been called");
  const size_t n = (size_t)b.nq * b.ne;
  const bool dev = is_device_ptr(out);
  if ((size_t)h->d_scratch.n < n) h->d_scratch.alloc(n);
  double* target = dev ? out : h->d_scratch.p;
  k_permute_scalar_out(h, b, b.d_scalar.p, target);  // [q*ne + e_tile] -> [q + NQ*e] in the caller's element order
  if (!dev) {
    FEC_CUDA(cudaMemcpyAsync(out, target, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    FEC_CUDA(cudaStreamSynchronize(h->stream));
  }
  FEC_API_END
}

int fecb200_vector_values(fecb200_handle* h, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const bool dev = is_device_ptr(out);
  double* target = dev ? out : h->d_out.p;
  if (!dev) wait_out_free(h);
  if (h->opts.condensed) FEC_CUDA(cudaMemcpyAsync(target, h->d_R.p, h->ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  else k_extract_unknowns(h, h->d_R.p, target);
  if (!dev) copy_out(h, out, target, len_Uu(h));
  FEC_API_END
}

int fecb200_residual(fecb200_handle* h, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const bool dev = is_device_ptr(out);
  double* target = dev ? out : h->d_out.p;
  if (!dev) wait_out_free(h);
  k_residual_accessor(h, target);
  if (h->peer_enabled && h->peer_field == FECB200_FIELD_RESIDUAL)  // ready for the neighbours' next scatter
    FEC_CUDA(cudaMemsetAsync(h->d_R.p, 0, h->ndof * sizeof(double), h->stream));
  if (!dev) copy_out(h, out, target, len_Uu(h));
  FEC_API_END
}

int fecb200_assemble_matrix(fecb200_handle* h, int32_t kind, const double* Uu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  nz_for_kind(h, kind, true);
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  ZeroFillPlan zf = kind == FECB200_STIFFNESS ? begin_stiffness_fill(h) : ZeroFillPlan{};
  double* nz = nz_for_kind(h, kind, true);
  if (!zf.on) FEC_CUDA(cudaMemsetAsync(nz, 0, h->nnz * sizeof(double), h->stream));  // fill!(storage, 0)  Matrix.jl:39
  join_inputs(h);
  k_update_field(h, h->d_U.p, u, true);
  inputs_consumed(h);
  h->last_ms = 0.f;
  for (auto& b : h->blocks) {
    MatLaunch a{h->d_U.p, nz, kind};
    zf.next(b, a);
    launch_matrix(h, b, a);
  }
  (kind == FECB200_STIFFNESS ? h->stiff_adjusted : h->mass_adjusted) = false;
  FEC_API_END
}

int fecb200_assemble_vector_and_matrix(fecb200_handle* h, const double* Uu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  assemble_vector_and_matrix_impl(h, u);
  FEC_API_END
}

int fecb200_set_matrix_double_buffer(fecb200_handle* h, int32_t enable) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  h->double_buffer = enable != 0;
  ensure_matrix_structure(h);
  if (!h->double_buffer) { FEC_CUDA(cudaStreamSynchronize(h->stream)); h->d_nz_stiff_alt.release(); h->alt_clean = false; }
  else if (h->matrix_ready && !h->opts.matrix_free && !h->d_nz_stiff_alt.p) {
    h->d_nz_stiff_alt.alloc_compressible(nz_alloc_len(h), h->device);
    h->d_nz_stiff_alt.zero(h->stream);
    h->alt_clean = true;
  }
  FEC_API_END
}

static void assemble_vector_and_matrix_impl(fecb200_handle* h, const double* u) {
  nz_for_kind(h, FECB200_STIFFNESS, true);
  ZeroFillPlan zf = begin_stiffness_fill(h);
  double* nz = h->d_nz_stiff.p;
  if (!(h->peer_enabled && h->peer_field == FECB200_FIELD_RESIDUAL))
    FEC_CUDA(cudaMemsetAsync(h->d_R.p, 0, h->ndof * sizeof(double), h->stream));
  if (!zf.on) FEC_CUDA(cudaMemsetAsync(nz, 0, h->nnz * sizeof(double), h->stream));   // the H2D of Uu (async mode) overlaps this
  join_inputs(h);
  k_update_field(h, h->d_U.p, u, true);
  inputs_consumed(h);
  h->last_ms = 0.f;
  for (auto& b : h->blocks) {
    if (!b.halo && matrix_kernel_fuses_residual(h, b)) {
      MatLaunch a{h->d_U.p, nz, FECB200_STIFFNESS, h->d_R.p};
      zf.next(b, a);
      launch_matrix(h, b, a);
    } else {
      if (!b.halo) { VecLaunch v{h->d_U.p, nullptr, h->d_R.p, MODE_RESIDUAL}; launch_vector(h, b, v); }
      MatLaunch a{h->d_U.p, nz, FECB200_STIFFNESS, nullptr};
      zf.next(b, a);
      launch_matrix(h, b, a);
    }
  }
  h->stiff_adjusted = false;
}

static double* adjusted_values(fecb200_handle* h, int kind) {
  double* nz = nz_for_kind(h, kind, false);
  FEC_REQUIRE(nz, "matrix of this kind has not been assembled");
  bool& adj = (kind == FECB200_STIFFNESS) ? h->stiff_adjusted : h->mass_adjusted;
  if (h->opts.condensed && !adj) { k_adjust_matrix(h, nz); adj = true; }
  return nz;
}

int fecb200_matrix_values(fecb200_handle* h, int32_t kind, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  double* nz = adjusted_values(h, kind);
  copy_out(h, out, nz, h->nnz);
  FEC_API_END
}

int fecb200_matrix_values_device(fecb200_handle* h, int32_t kind, double** nz_dev) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && nz_dev, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  *nz_dev = adjusted_values(h, kind);
  FEC_API_END
}

int fecb200_assemble_action(fecb200_handle* h, int32_t kind, const double* Uu, const double* Vu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu && Vu, "null argument");
  FEC_REQUIRE(kind == FECB200_STIFFNESS || kind == FECB200_MASS, "action kind must be STIFFNESS or MASS");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  const double* v = stage_in(h, Vu, h->d_Vu.p, len_Uu(h));
  assemble_vector_impl(h, kind == FECB200_STIFFNESS ? MODE_ACTION_STIFFNESS : MODE_ACTION_MASS, u, v, h->d_Av.p);
  FEC_API_END
}

int fecb200_assemble_action_full(fecb200_handle* h, int32_t kind, const double* U_full, const double* v_full) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && U_full && v_full, "null argument");
  FEC_REQUIRE(kind == FECB200_STIFFNESS || kind == FECB200_MASS, "action kind must be STIFFNESS or MASS");
  FEC_CUDA(cudaSetDevice(h->device));
  // _update_for_assembly_full! (Parameters.jl:432-442): plain copies, BC slots are NOT overwritten
  const auto kindU = is_device_ptr(U_full) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const auto kindV = is_device_ptr(v_full) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  FEC_CUDA(cudaMemcpyAsync(h->d_U.p, U_full, h->ndof * sizeof(double), kindU, h->stream));
  FEC_CUDA(cudaMemcpyAsync(h->d_V.p, v_full, h->ndof * sizeof(double), kindV, h->stream));
  FEC_CUDA(cudaMemsetAsync(h->d_Av.p, 0, h->ndof * sizeof(double), h->stream));
  h->last_ms = 0.f;
  for (auto& b : h->blocks) {
    if (b.halo) continue;
    VecLaunch a{h->d_U.p, h->d_V.p, h->d_Av.p, kind == FECB200_STIFFNESS ? MODE_ACTION_STIFFNESS : MODE_ACTION_MASS};
    launch_vector(h, b, a);
  }
  k_zero_bc_slots(h, h->d_V.p);  // restore the "BC slots of V are zero" invariant (MatrixAction.jl:137-148)
  FEC_API_END
}

int fecb200_hvp(fecb200_handle* h, const double* v, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* vd = nullptr;
  if (h->opts.condensed) {
    FEC_REQUIRE(v, "hvp needs v in condensed mode");
    vd = stage_in(h, v, h->d_Vu.p, h->ndof);
  }
  const bool dev = is_device_ptr(out);
  double* target = dev ? out : h->d_out.p;
  join_inputs(h);
  if (!dev) wait_out_free(h);
  k_hvp_accessor(h, vd, target);
  inputs_consumed(h);
  if (!dev) copy_out(h, out, target, len_Uu(h));
  FEC_API_END
}

int fecb200_update_field(fecb200_handle* h, const double* Uu) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* u = stage_in(h, Uu, h->d_Uu.p, len_Uu(h));
  join_inputs(h);
  k_update_field(h, h->d_U.p, u, true);
  inputs_consumed(h);
  FEC_API_END
}

int fecb200_matrix_multiply(fecb200_handle* h, int32_t kind, const double* x, double* y) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && x && y, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  double* nz = adjusted_values(h, kind);
  const int64_t n = len_Uu(h);
  ensure_cg(h);
  const double* xd = stage_in(h, x, h->d_Vu.p, n);
  join_inputs(h);
  const bool dev = is_device_ptr(y);
  double* yd = dev ? y : h->d_cg_Ap.p;
  if (comm_active(h) && h->n_neighbors > 0) {   // ghost entries of x from their owners (on a private copy: x is const)
    FEC_CUDA(cudaMemcpyAsync(h->d_cg_p.p, xd, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    comm_halo_update_unknowns(h, h->d_cg_p.p);
    xd = h->d_cg_p.p;
  }
  if (!dev || h->n_owned_nodes < h->nn) FEC_CUDA(cudaMemsetAsync(yd, 0, n * sizeof(double), h->stream));  // rows that are not stored
  spmv(h, nz, xd, yd);
  inputs_consumed(h);
  if (!dev) copy_out(h, y, yd, n);
  FEC_API_END
}

int fecb200_field_copy(fecb200_handle* h, int32_t which, double* out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && out, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const double* src = nullptr;
  switch (which) {
    case FECB200_FIELD_U: src = h->d_U.p; break;
    case FECB200_FIELD_RESIDUAL: src = h->d_R.p; break;
    case FECB200_FIELD_ACTION: src = h->d_Av.p; break;
    case FECB200_FIELD_V: src = h->d_V.p; break;
    default: throw Error("fecb200: bad field selector");
  }
  copy_out(h, out, src, h->ndof);
  if (is_device_ptr(out)) FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

int fecb200_cg_solve(fecb200_handle* h, const double* b, double* x, double atol, double rtol, int64_t itmax,
                     int32_t matrix_free, int64_t* iters_out, double* rnorm_out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && b && x, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const int64_t n = len_Uu(h);
  if (!matrix_free) { ensure_matrix_structure(h); FEC_REQUIRE(h->matrix_ready && !h->opts.matrix_free, "no assembled matrix for CG"); }
  ensure_cg(h);
  const double* bd = stage_in(h, b, h->d_Vu.p, n);
  const bool dev = is_device_ptr(x);
  double* xd = dev ? x : h->d_cg_x.p;
  if (atol < 0) atol = std::sqrt(2.220446049250313e-16);
  if (rtol < 0) rtol = std::sqrt(2.220446049250313e-16);
  if (itmax <= 0) itmax = 2 * n;
  if (!matrix_free) adjusted_values(h, FECB200_STIFFNESS);
  join_inputs(h);  // async mode: the H2D of b runs on the copy stream; the main stream must not read it early
  cg_impl(h, bd, xd, atol, rtol, itmax, matrix_free != 0, iters_out, rnorm_out);
  if (!dev) copy_out(h, x, xd, n);
  FEC_API_END
}

int fecb200_newton_solve(fecb200_handle* h, double* Uu, int32_t max_iters, double tol, int32_t matrix_free,
                         int32_t* newton_iters_out, int64_t* cg_iters_out, double* rnorm_out) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu, "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  const int64_t n = len_Uu(h);
  const int64_t nred = owned_len(h);
  const bool dist = comm_active(h) && h->n_neighbors > 0;
  ensure_cg(h);
  const bool dev = is_device_ptr(Uu);
  double* u = h->d_Uu.p;
  if (dev) FEC_CUDA(cudaMemcpyAsync(u, Uu, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  else FEC_CUDA(cudaMemcpyAsync(u, Uu, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (max_iters <= 0) max_iters = 10;   // NewtonSolverSettings() (Solvers.jl:174-176)
  if (tol <= 0) tol = 1e-12;
  const double eps = std::sqrt(2.220446049250313e-16);
  double R0 = 0.0, nR = 0.0;
  int64_t cg_total = 0;
  int it = 0;
  double* Rb = h->d_out.p;   // residual(asm)
  double* dU = h->d_cg_x.p;
  const bool peer_R = h->peer_enabled && h->peer_field == FECB200_FIELD_RESIDUAL;
  FEC_REQUIRE(!peer_R || comm_active(h), "Newton over the peer-memory halo needs the library communicator (fecb200_comm_init)");
  if (dist) comm_halo_update_unknowns(h, u);   // consistent ghost entries of the initial guess
  if (peer_R) {
    FEC_CUDA(cudaMemsetAsync(h->d_R.p, 0, h->ndof * sizeof(double), h->stream));
    comm_barrier(h);
  }
  for (it = 1; it <= max_iters; ++it) {
    // solve!(IterativeLinearSolver) (Solvers.jl:128-153)
    if (matrix_free) {
      assemble_vector_impl(h, MODE_RESIDUAL, u, nullptr, h->d_R.p);
    } else {
      // residual and tangent at the same state: fused pass (see fecb200_assemble_vector_and_matrix)
      assemble_vector_and_matrix_impl(h, u);
    }
    add_source_loads(h, h->d_R.p);    // assemble_vector_source!      (Solvers.jl:135)
    add_neumann_loads(h, h->d_R.p);   // assemble_vector_neumann_bc!  (Solvers.jl:136)
    // ghost -> owner sum of the residual (PartitionedArraysExt.jl:469-481)
    if (peer_R) comm_barrier(h);      // fused halo: the kernels already added the ghost rows; wait for every rank
    else if (dist) comm_halo_sum_field(h, h->d_R.p);
    k_residual_accessor(h, Rb);
    if (peer_R) {                     // zero-after-read protocol of the peer halo (assemble_* does not clear R then)
      FEC_CUDA(cudaMemsetAsync(h->d_R.p, 0, h->ndof * sizeof(double), h->stream));
      comm_barrier(h);                // every rank re-zeroed its residual before anyone scatters again
    }
    if (!matrix_free) adjusted_values(h, FECB200_STIFFNESS);
    int64_t cgit = 0;
    cg_impl(h, Rb, dU, eps, eps, 2 * n, matrix_free != 0, &cgit, nullptr);
    cg_total += cgit;
    axpy(h, -1.0, dU, u, n);  // Uu += -solution
    const double ndU = std::sqrt(dot(h, dU, dU, nred));
    nR = std::sqrt(dot(h, Rb, Rb, nred));
    if (it == 1) R0 = nR;
    const double rel = R0 > 0.0 ? nR / R0 : nR;
    if (ndU < tol || nR < tol || rel < tol) break;
  }
  if (it > max_iters) it = max_iters;
  copy_out(h, Uu, u, n);
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  if (newton_iters_out) *newton_iters_out = it;
  if (cg_iters_out) *cg_iters_out = cg_total;
  if (rnorm_out) *rnorm_out = nR;
  FEC_API_END
}

int fecb200_metis_part_mesh_dual(int64_t ne, int64_t nn, const int64_t* eptr, const int64_t* eind, int64_t ncommon,
                                 int64_t nparts, int64_t* epart, int64_t* npart) {
  FEC_API_BEGIN
  FEC_REQUIRE(ne > 0 && nn > 0 && eptr && eind && epart && npart && nparts >= 1, "bad METIS arguments");
  const int rc = metis_mesh_dual(ne, nn, eptr, eind, ncommon, nparts, epart, npart);
  FEC_REQUIRE(rc == 1, "METIS_PartMeshDual failed");
  FEC_API_END
}

int fecb200_metis_part_graph(int64_t nv, const int64_t* xadj, const int64_t* adjncy, int64_t nparts, int64_t* part) {
  FEC_API_BEGIN
  FEC_REQUIRE(nv > 0 && xadj && adjncy && part && nparts >= 1, "bad METIS arguments");
  const int rc = metis_graph(nv, xadj, adjncy, nparts, part);
  FEC_REQUIRE(rc == 1, "METIS_PartGraphKway failed");
  FEC_API_END
}

int fecb200_partition_setup(fecb200_handle* h, int64_t n_owned_nodes, const int32_t* block_is_halo) {
  FEC_API_BEGIN
  PhaseTimer _pt("fecb200_partition_setup (total)");
  FEC_REQUIRE(h, "null handle");
  FEC_REQUIRE(n_owned_nodes >= 0 && n_owned_nodes <= h->nn, "n_owned_nodes out of range");
  FEC_CUDA(cudaSetDevice(h->device));
  h->n_owned_nodes = n_owned_nodes;
  for (size_t b = 0; b < h->blocks.size(); ++b) h->blocks[b].halo = block_is_halo && block_is_halo[b] != 0;
  FEC_REQUIRE(h->opts.matrix_free || h->opts.matrix_type == FECB200_CSR, "partitioned assembly stores owned ROWS: use CSR");
  build_matrix_structure(h);  // DOF maps and Dirichlet values are untouched
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

int fecb200_halo_setup(fecb200_handle* h, int32_t n_neighbors, const int32_t* ranks, const int64_t* send_ptr,
                       const int64_t* send_nodes, const int64_t* recv_ptr, const int64_t* recv_nodes) {
  FEC_API_BEGIN
  PhaseTimer _pt("fecb200_halo_setup (total)");
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  h->n_neighbors = n_neighbors;
  h->neighbor_ranks.assign(ranks, ranks + n_neighbors);
  h->send_ptr.assign(send_ptr, send_ptr + n_neighbors + 1);
  h->recv_ptr.assign(recv_ptr, recv_ptr + n_neighbors + 1);
  std::vector<int32_t> sn(h->send_ptr.back()), rn(h->recv_ptr.back());
  for (size_t i = 0; i < sn.size(); ++i) {
    FEC_REQUIRE(send_nodes[i] >= 1 && send_nodes[i] <= h->nn, "halo send node out of range");
    sn[i] = (int32_t)(send_nodes[i] - 1);
  }
  for (size_t i = 0; i < rn.size(); ++i) {
    FEC_REQUIRE(recv_nodes[i] >= 1 && recv_nodes[i] <= h->nn, "halo recv node out of range");
    rn[i] = (int32_t)(recv_nodes[i] - 1);
  }
  h->d_send_nodes.upload(sn, h->stream);
  h->d_recv_nodes.upload(rn, h->stream);
  FEC_API_END
}


int fecb200_halo_pack(fecb200_handle* h, int32_t which, double* sendbuf_dev) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && (sendbuf_dev || h->d_send_nodes.n == 0), "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  halo_pack(h, field_ptr(h, which), sendbuf_dev);
  FEC_API_END
}

int fecb200_halo_send_size(fecb200_handle* h, int64_t* n_doubles) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && n_doubles, "null argument");
  *n_doubles = (int64_t)h->d_send_nodes.n * h->nf;
  FEC_API_END
}

int fecb200_halo_unpack_add(fecb200_handle* h, int32_t which, const double* recvbuf_dev) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && (recvbuf_dev || h->d_recv_nodes.n == 0), "null argument");
  FEC_CUDA(cudaSetDevice(h->device));
  halo_unpack_add(h, field_ptr(h, which), recvbuf_dev);
  FEC_API_END
}

int fecb200_halo_recv_size(fecb200_handle* h, int64_t* n_doubles) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && n_doubles, "null argument");
  *n_doubles = (int64_t)h->d_recv_nodes.n * h->nf;
  FEC_API_END
}

int fecb200_ipc_export(fecb200_handle* h, int32_t which, void* handle64) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && handle64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  FEC_CUDA(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  FEC_CUDA(cudaIpcGetMemHandle(&mh, field_ptr(h, which)));
  memcpy(handle64, &mh, 64);
  FEC_API_END
}

static void peer_detach_impl(fecb200_handle* h) {
  for (void* p : h->peer_opened) cudaIpcCloseMemHandle(p);
  h->peer_opened.clear();
  h->peer_enabled = false;
  h->peer = fec::PeerScatter{-1, nullptr, nullptr, {nullptr}};
}

int fecb200_peer_attach(fecb200_handle* h, int32_t which, int32_t n_peers, const void* handles64,
                        const int64_t* peer_n_nodes, const int32_t* ghost_peer, const int64_t* ghost_node, int64_t n_ghosts) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && (n_peers == 0 || (handles64 && peer_n_nodes)), "null argument");
  FEC_REQUIRE(n_peers >= 0 && n_peers <= kMaxPeers, "too many peers");
  // validate BEFORE any handle is opened: the kernels write to base[peer] + ghost_node * NF + d in another GPU's memory
  for (int64_t g = 0; g < n_ghosts; ++g) {
    FEC_REQUIRE(ghost_peer[g] >= -1 && ghost_peer[g] < n_peers, "ghost_peer out of range");
    if (ghost_peer[g] >= 0)
      FEC_REQUIRE(ghost_node[g] >= 0 && ghost_node[g] < peer_n_nodes[ghost_peer[g]] && ghost_node[g] < (int64_t)INT32_MAX,
                  "ghost_node out of range for its owner (stale or mismatched exchange list)");
  }
  FEC_REQUIRE(n_ghosts == h->nn - h->n_owned_nodes, "ghost arrays must cover every ghost node (call partition_setup first)");
  FEC_REQUIRE(which == FECB200_FIELD_RESIDUAL || which == FECB200_FIELD_ACTION, "peer scatter targets the residual or action field");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  peer_detach_impl(h);
  h->peer.n_owned = h->n_owned_nodes;
  for (int i = 0; i < n_peers; ++i) {
    cudaIpcMemHandle_t mh;
    memcpy(&mh, (const char*)handles64 + 64 * i, 64);
    void* ptr = nullptr;
    FEC_CUDA(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
    h->peer_opened.push_back(ptr);
    h->peer.base[i] = (double*)ptr;
  }
  std::vector<int32_t> gp(ghost_peer, ghost_peer + n_ghosts), gn(n_ghosts);
  for (int64_t g = 0; g < n_ghosts; ++g) {
    gn[g] = gp[g] >= 0 ? (int32_t)ghost_node[g] : 0;
  }
  h->d_ghost_peer.upload(gp, h->stream);
  h->d_ghost_node.upload(gn, h->stream);
  h->peer.ghost_peer = h->d_ghost_peer.p;
  h->peer.ghost_node = h->d_ghost_node.p;
  h->peer_field = which;
  h->peer_enabled = true;
  FEC_API_END
}

int fecb200_peer_detach(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  peer_detach_impl(h);
  FEC_API_END
}

int fecb200_launch_count(fecb200_handle* h, int64_t* n) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && n, "null argument");
  *n = h->launches;
  FEC_API_END
}

int fecb200_block_kernel_form(fecb200_handle* h, int32_t block, int32_t* form) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && form, "null argument");
  FEC_REQUIRE(block >= 0 && block < (int32_t)h->blocks.size(), "block index out of range");
  *form = h->blocks[block].walsh ? 1 : 0;
  FEC_API_END
}

int fecb200_enable_timing(fecb200_handle* h, int32_t on) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  h->timing = on != 0;
  FEC_API_END
}

int fecb200_last_kernel_ms(fecb200_handle* h, float* ms) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && ms, "null argument");
  *ms = h->last_ms;
  FEC_API_END
}

}  // extern "C"
