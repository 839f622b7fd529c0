// comm.cu -- the collective plane of libfecb200: one process per GPU, NCCL over NVLink.
//
// Replaces, for a Julia host that only has include/fecb200.h:
//   PVector assembly of ghost contributions ("assemble!")     ext/PartitionedArraysExt.jl:469-481
//   consistent!(::PVector) (owner -> ghost copies)            ext/PartitionedArraysExt.jl:449-459
//   distributed dots / norms of the Krylov solve              ext/PartitionedArraysExt.jl:522-540, src/Solvers.jl:128-153
//
// NCCL is resolved with dlopen (libnccl.so.2): the library still loads where NCCL is absent, and inside a process
// that already carries an NCCL (e.g. a PyTorch host) the SAME library instance is used.  The only thing the host
// moves out of band is the 128-byte ncclUniqueId of rank 0 (MPI_Bcast in the reference's setting).
#include "common.cuh"
#include <algorithm>
#include <dlfcn.h>

namespace fec {

namespace {
// the slice of nccl.h this file needs (stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;

struct Nccl {
  bool ok = false;
  std::string why;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  Nccl() {
    // prefer an instance that is already mapped into the process (RTLD_NOLOAD), then FECB200_NCCL_LIB, then the loader path
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) { const char* e = getenv("FECB200_NCCL_LIB"); if (e) lib = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { why = "libnccl.so.2 not found (set FECB200_NCCL_LIB)"; return; }
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
    Send = reinterpret_cast<decltype(Send)>(sym("ncclSend"));
    Recv = reinterpret_cast<decltype(Recv)>(sym("ncclRecv"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    ok = GetUniqueId && CommInitRank && CommDestroy && GroupStart && GroupEnd && Send && Recv && AllReduce && AllGather &&
         GetErrorString;
    if (!ok) why = "libnccl.so.2 lacks a required symbol";
  }
};
Nccl& nccl() { static Nccl n; return n; }

#define FEC_NCCL(call)                                                                                   \
  do {                                                                                                   \
    ncclResult_t _r = (call);                                                                            \
    if (_r != ncclSuccess) throw Error(std::string("NCCL error: ") + nccl().GetErrorString(_r) + " in " #call); \
  } while (0)

inline ncclComm_t comm_of(fecb200_handle* h) {
  FEC_REQUIRE(h->comm, "no communicator: call fecb200_comm_init first");
  return static_cast<ncclComm_t>(h->comm);
}
inline int grid_for(int64_t n, int bs = 256) { return (int)((n + bs - 1) / bs); }

// unknown-indexed vectors (Uu layout): entries of the nodes in `nodes`, NF per node; constrained dofs travel as 0
__global__ void k_pack_unknowns(const double* v, const int32_t* nodes, const int32_t* d2u, double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * nf) return;
  const int32_t u = d2u[(int64_t)nodes[i / nf] * nf + i % nf];
  buf[i] = u >= 0 ? v[u] : 0.0;
}
__global__ void k_unpack_set_unknowns(double* v, const int32_t* nodes, const int32_t* d2u, const double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * nf) return;
  const int32_t u = d2u[(int64_t)nodes[i / nf] * nf + i % nf];
  if (u >= 0) v[u] = buf[i];
}
__global__ void k_pack_nodes(const double* f, const int32_t* nodes, double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n * nf) buf[i] = f[(int64_t)nodes[i / nf] * nf + i % nf];
}
__global__ void k_unpack_set_nodes(double* f, const int32_t* nodes, const double* buf, int64_t n, int nf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n * nf) f[(int64_t)nodes[i / nf] * nf + i % nf] = buf[i];
}

void ensure_bufs(fecb200_handle* h) {
  const size_t ns = std::max<size_t>(1, h->d_send_nodes.n * h->nf), nr = std::max<size_t>(1, h->d_recv_nodes.n * h->nf);
  const size_t ng = std::max(h->d_g_own_nodes.n, h->d_g_ghost_nodes.n) * h->nf;
  const size_t n = std::max(std::max(ns, nr), ng);  // each buffer serves both directions (halo_sum and halo_update)
  if (h->d_sendbuf.n < n) h->d_sendbuf.alloc(n);
  if (h->d_recvbuf.n < n) h->d_recvbuf.alloc(n);
}

// one grouped point-to-point exchange with every neighbour: segment i of `out` goes to ranks[i], segment i of `in`
// comes from it (segments in nodes, NF doubles per node)
void exchange(fecb200_handle* h, const std::vector<int32_t>& ranks, const std::vector<int64_t>& optr,
              const std::vector<int64_t>& iptr, const double* out, double* in) {
  Nccl& N = nccl();
  ncclComm_t c = comm_of(h);
  FEC_NCCL(N.GroupStart());
  for (size_t i = 0; i < ranks.size(); ++i) {
    const int64_t no = (optr[i + 1] - optr[i]) * h->nf, ni = (iptr[i + 1] - iptr[i]) * h->nf;
    if (no) FEC_NCCL(N.Send(out + optr[i] * h->nf, (size_t)no, ncclFloat64, ranks[i], c, h->stream));
    if (ni) FEC_NCCL(N.Recv(in + iptr[i] * h->nf, (size_t)ni, ncclFloat64, ranks[i], c, h->stream));
  }
  FEC_NCCL(N.GroupEnd());
}
// the lists of the owner -> ghost update: fecb200_ghost_setup's (every ghost) or, without them, the residual halo
// lists reversed (only the ghosts that owned elements touch)
struct UpdateLists {
  const std::vector<int32_t>* ranks; const std::vector<int64_t>* own_ptr; const std::vector<int64_t>* ghost_ptr;
  const int32_t* own_nodes; const int32_t* ghost_nodes; int64_t n_own, n_ghost;
};
UpdateLists update_lists(fecb200_handle* h) {
  if (h->ghost_lists)
    return {&h->g_ranks, &h->g_own_ptr, &h->g_ghost_ptr, h->d_g_own_nodes.p, h->d_g_ghost_nodes.p,
            (int64_t)h->d_g_own_nodes.n, (int64_t)h->d_g_ghost_nodes.n};
  return {&h->neighbor_ranks, &h->recv_ptr, &h->send_ptr, h->d_recv_nodes.p, h->d_send_nodes.p,
          (int64_t)h->d_recv_nodes.n, (int64_t)h->d_send_nodes.n};
}
}  // namespace

bool comm_active(const fecb200_handle* h) { return h->comm != nullptr; }

// sum of one device scalar (or a few) over all ranks, in place, stream-ordered
void comm_allreduce_sum(fecb200_handle* h, double* dev, int n) {
  FEC_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, comm_of(h), h->stream));
}

void comm_barrier(fecb200_handle* h) {
  if (!h->d_bar.p) { h->d_bar.alloc(2); h->d_bar.zero(h->stream); }
  FEC_NCCL(nccl().AllReduce(h->d_bar.p, h->d_bar.p, 1, ncclFloat32, ncclSum, comm_of(h), h->stream));
}

// ghost -> owner sum of a full-length nodal field
void comm_halo_sum_field(fecb200_handle* h, double* field) {
  ensure_bufs(h);
  halo_pack(h, field, h->d_sendbuf.p);
  exchange(h, h->neighbor_ranks, h->send_ptr, h->recv_ptr, h->d_sendbuf.p, h->d_recvbuf.p);
  halo_unpack_add(h, field, h->d_recvbuf.p);
}

// owner -> ghost copy of a full-length nodal field
void comm_halo_update_field(fecb200_handle* h, double* field) {
  ensure_bufs(h);
  const UpdateLists L = update_lists(h);
  const int64_t no = L.n_own, ni = L.n_ghost;
  if (no) { k_pack_nodes<<<grid_for(no * h->nf), 256, 0, h->stream>>>(field, L.own_nodes, h->d_sendbuf.p, no, h->nf); h->launches++; }
  exchange(h, *L.ranks, *L.own_ptr, *L.ghost_ptr, h->d_sendbuf.p, h->d_recvbuf.p);
  if (ni) { k_unpack_set_nodes<<<grid_for(ni * h->nf), 256, 0, h->stream>>>(field, L.ghost_nodes, h->d_recvbuf.p, ni, h->nf); h->launches++; }
  FEC_CUDA(cudaGetLastError());
}

// owner -> ghost copy of an unknown-indexed vector (the Uu layout): what consistent!(x) does before K * x
void comm_halo_update_unknowns(fecb200_handle* h, double* v) {
  ensure_bufs(h);
  const UpdateLists L = update_lists(h);
  const int64_t no = L.n_own, ni = L.n_ghost;
  if (no) { k_pack_unknowns<<<grid_for(no * h->nf), 256, 0, h->stream>>>(v, L.own_nodes, h->d_d2u.p, h->d_sendbuf.p, no, h->nf); h->launches++; }
  exchange(h, *L.ranks, *L.own_ptr, *L.ghost_ptr, h->d_sendbuf.p, h->d_recvbuf.p);
  if (ni) { k_unpack_set_unknowns<<<grid_for(ni * h->nf), 256, 0, h->stream>>>(v, L.ghost_nodes, h->d_d2u.p, h->d_recvbuf.p, ni, h->nf); h->launches++; }
  FEC_CUDA(cudaGetLastError());
}

// number of leading entries of a Uu-shaped vector that belong to OWNED nodes (owned nodes come first, dofs ascending)
int64_t owned_len(const fecb200_handle* h) {
  if (h->n_owned_nodes >= h->nn) return h->opts.condensed ? h->ndof : h->n_unknowns;
  const int64_t lim = h->n_owned_nodes * h->nf;  // dofs 1..lim are owned
  if (h->opts.condensed) return lim;
  return (int64_t)(std::upper_bound(h->unknown_dofs.begin(), h->unknown_dofs.end(), lim) - h->unknown_dofs.begin());
}

void comm_release(fecb200_handle* h) {
  if (h->comm) { nccl().CommDestroy(static_cast<ncclComm_t>(h->comm)); h->comm = nullptr; }
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

extern "C" {

int fecb200_comm_unique_id(void* id128) {
  FEC_API_BEGIN
  FEC_REQUIRE(id128, "null argument");
  FEC_REQUIRE(nccl().ok, nccl().why);
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FEC_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, 128);
  FEC_API_END
}

int fecb200_comm_init(fecb200_handle* h, int32_t rank, int32_t nranks, const void* id128) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && id128, "null argument");
  FEC_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  FEC_REQUIRE(nccl().ok, nccl().why);
  FEC_CUDA(cudaSetDevice(h->device));
  comm_release(h);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  FEC_NCCL(nccl().CommInitRank(&c, nranks, id, rank));
  h->comm = c;
  h->comm_rank = rank;
  h->comm_nranks = nranks;
  FEC_API_END
}

int fecb200_comm_destroy(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  comm_release(h);
  FEC_API_END
}

int fecb200_comm_barrier(fecb200_handle* h) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  comm_barrier(h);
  FEC_API_END
}

int fecb200_comm_allreduce_sum(fecb200_handle* h, double* vals, int32_t n) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && vals && n >= 1 && n <= 4, "allreduce of 1..4 host scalars");
  FEC_CUDA(cudaSetDevice(h->device));
  if (!h->d_red.p) h->d_red.alloc(4);
  FEC_CUDA(cudaMemcpyAsync(h->d_red.p, vals, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  comm_allreduce_sum(h, h->d_red.p, n);
  FEC_CUDA(cudaMemcpyAsync(vals, h->d_red.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  FEC_API_END
}

static double* comm_field(fecb200_handle* h, int which) {
  switch (which) {
    case FECB200_FIELD_U: return h->d_U.p;
    case FECB200_FIELD_RESIDUAL: return h->d_R.p;
    case FECB200_FIELD_ACTION: return h->d_Av.p;
    case FECB200_FIELD_V: return h->d_V.p;
    default: throw Error("fecb200: bad field selector");
  }
}

int fecb200_ghost_setup(fecb200_handle* h, int32_t n_neighbors, const int32_t* ranks, const int64_t* own_ptr,
                        const int64_t* own_nodes, const int64_t* ghost_ptr, const int64_t* ghost_nodes) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && n_neighbors >= 0 && (n_neighbors == 0 || (ranks && own_ptr && ghost_ptr)), "bad argument");
  FEC_CUDA(cudaSetDevice(h->device));
  h->g_ranks.assign(ranks, ranks + n_neighbors);
  h->g_own_ptr.assign(own_ptr, own_ptr + n_neighbors + 1);
  h->g_ghost_ptr.assign(ghost_ptr, ghost_ptr + n_neighbors + 1);
  std::vector<int32_t> on(h->g_own_ptr.back()), gn(h->g_ghost_ptr.back());
  for (size_t i = 0; i < on.size(); ++i) {
    FEC_REQUIRE(own_nodes[i] >= 1 && own_nodes[i] <= h->n_owned_nodes, "ghost_setup: own list holds a node this rank does not own");
    on[i] = (int32_t)(own_nodes[i] - 1);
  }
  for (size_t i = 0; i < gn.size(); ++i) {
    FEC_REQUIRE(ghost_nodes[i] > h->n_owned_nodes && ghost_nodes[i] <= h->nn, "ghost_setup: ghost list holds an owned node");
    gn[i] = (int32_t)(ghost_nodes[i] - 1);
  }
  h->d_g_own_nodes.upload(on, h->stream);
  h->d_g_ghost_nodes.upload(gn, h->stream);
  h->ghost_lists = true;
  FEC_API_END
}

int fecb200_halo_sum(fecb200_handle* h, int32_t which) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  if (h->peer_enabled && h->peer_field == which) comm_barrier(h);  // the kernels already added the ghost rows into their owners
  else comm_halo_sum_field(h, comm_field(h, which));
  FEC_API_END
}

int fecb200_halo_update(fecb200_handle* h, int32_t which) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  comm_halo_update_field(h, comm_field(h, which));
  FEC_API_END
}

int fecb200_halo_update_unknowns(fecb200_handle* h, double* Uu_dev) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && Uu_dev, "null argument");
  FEC_REQUIRE(is_device_ptr(Uu_dev), "halo_update_unknowns takes a device vector");
  FEC_CUDA(cudaSetDevice(h->device));
  comm_halo_update_unknowns(h, Uu_dev);
  FEC_API_END
}

int fecb200_owned_length(fecb200_handle* h, int64_t* n) {
  FEC_API_BEGIN
  FEC_REQUIRE(h && n, "null argument");
  *n = owned_len(h);
  FEC_API_END
}

// Switch the ghost -> owner sum of `which` to the fused NVLink path: the owner-local ids of this rank's ghosts and the
// CUDA IPC handles of every rank's field are exchanged over NCCL inside the library, then fecb200_peer_attach.
int fecb200_comm_peer_enable(fecb200_handle* h, int32_t which) {
  FEC_API_BEGIN
  FEC_REQUIRE(h, "null handle");
  FEC_CUDA(cudaSetDevice(h->device));
  Nccl& N = nccl();
  ncclComm_t c = comm_of(h);
  const int nb = h->n_neighbors;
  const int64_t nsend = h->send_ptr.empty() ? 0 : h->send_ptr.back(), nrecv = h->recv_ptr.empty() ? 0 : h->recv_ptr.back();
  // (1) every neighbour tells me ITS local ids of the nodes I ghost (its recv list for me; both sides sorted by global id)
  DevBuf<int32_t> d_theirs;
  d_theirs.alloc(std::max<int64_t>(1, nsend));
  FEC_NCCL(N.GroupStart());
  for (int i = 0; i < nb; ++i) {
    const int64_t nr = h->recv_ptr[i + 1] - h->recv_ptr[i], ns = h->send_ptr[i + 1] - h->send_ptr[i];
    if (nr) FEC_NCCL(N.Send(h->d_recv_nodes.p + h->recv_ptr[i], (size_t)nr, ncclInt32, h->neighbor_ranks[i], c, h->stream));
    if (ns) FEC_NCCL(N.Recv(d_theirs.p + h->send_ptr[i], (size_t)ns, ncclInt32, h->neighbor_ranks[i], c, h->stream));
  }
  FEC_NCCL(N.GroupEnd());
  // (2) IPC handle + node count of every rank's field
  struct Rec { unsigned char handle[64]; int64_t nn; int64_t pad; };
  static_assert(sizeof(Rec) == 80, "record layout");
  Rec mine{};
  cudaIpcMemHandle_t mh;
  FEC_CUDA(cudaIpcGetMemHandle(&mh, comm_field(h, which)));
  memcpy(mine.handle, &mh, 64);
  mine.nn = h->nn;
  DevBuf<unsigned char> d_all;
  d_all.alloc(sizeof(Rec) * (size_t)h->comm_nranks);
  DevBuf<unsigned char> d_mine;
  d_mine.alloc(sizeof(Rec));
  FEC_CUDA(cudaMemcpyAsync(d_mine.p, &mine, sizeof(Rec), cudaMemcpyHostToDevice, h->stream));
  FEC_NCCL(N.AllGather(d_mine.p, d_all.p, sizeof(Rec), ncclUint8, c, h->stream));
  std::vector<Rec> all(h->comm_nranks);
  std::vector<int32_t> theirs(std::max<int64_t>(1, nsend));
  FEC_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(Rec) * all.size(), cudaMemcpyDeviceToHost, h->stream));
  FEC_CUDA(cudaMemcpyAsync(theirs.data(), d_theirs.p, sizeof(int32_t) * theirs.size(), cudaMemcpyDeviceToHost, h->stream));
  std::vector<int32_t> send_nodes(std::max<int64_t>(1, nsend));
  if (nsend) FEC_CUDA(cudaMemcpyAsync(send_nodes.data(), h->d_send_nodes.p, sizeof(int32_t) * nsend, cudaMemcpyDeviceToHost, h->stream));
  FEC_CUDA(cudaStreamSynchronize(h->stream));
  (void)nrecv;
  // (3) peers = neighbours that own some of my ghosts
  std::vector<unsigned char> handles;
  std::vector<int64_t> peer_nn;
  const int64_t n_ghost = h->nn - h->n_owned_nodes;
  std::vector<int32_t> gpeer(n_ghost, -1);
  std::vector<int64_t> gnode(n_ghost, 0);
  int np = 0;
  for (int i = 0; i < nb; ++i) {
    const int64_t ns = h->send_ptr[i + 1] - h->send_ptr[i];
    if (!ns) continue;
    const Rec& r = all[h->neighbor_ranks[i]];
    handles.insert(handles.end(), r.handle, r.handle + 64);
    peer_nn.push_back(r.nn);
    for (int64_t k = h->send_ptr[i]; k < h->send_ptr[i + 1]; ++k) {
      const int64_t g = (int64_t)send_nodes[k] - h->n_owned_nodes;
      FEC_REQUIRE(g >= 0 && g < n_ghost, "halo send list holds an owned node");
      gpeer[g] = np;
      gnode[g] = theirs[k];
    }
    ++np;
  }
  const int rc = fecb200_peer_attach(h, which, np, handles.data(), peer_nn.data(), gpeer.data(), gnode.data(), n_ghost);
  if (rc) throw Error(fec::g_last_error);
  comm_barrier(h);  // every rank has opened its peers before anyone scatters
  FEC_API_END
}

}  // extern "C"
