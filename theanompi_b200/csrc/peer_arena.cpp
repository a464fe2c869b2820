// Symmetric peer-memory registry (C++): allocation, inter-process handle exchange, peer mapping, NVLS multicast.
//
// Reference counterpart: the vendored, patched `cuda_ndarray.cu` (test/test-train-mode/test-as-buffer/theano/sandbox/
// cuda/cuda_ndarray.cu:650-683) exposed a device pointer as a Python buffer so CUDA-aware MPI could move it; the
// loader used cudaIpc handles over ZeroMQ (models/data/imagenet.py:302-311).  Here every rank
//   1. allocates its arena + signal pad with the CUDA VMM API (cuMemCreate, POSIX-fd shareable handle),
//   2. ships the fds to its node-local peers over abstract-namespace Unix sockets (SCM_RIGHTS),
//   3. imports and maps every peer's allocation (cuMemImportFromShareableHandle / cuMemMap / cuMemSetAccess),
//   4. optionally binds all arenas to one multicast object (cuMulticastCreate / AddDevice / BindMem) so kernels can
//      use multimem.ld_reduce / multimem.st (NVLS, reduction inside the NVSwitch),
// and hands the kernels a pointer table (tmpi::CommCtx).  Fallback when VMM export is unavailable: cudaMalloc +
// cudaIpcGetMemHandle/cudaIpcOpenMemHandle (handles travel through the Python control plane).
#include "peer_arena.h"

#include <cuda.h>
#include <errno.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <mutex>
#include <stdexcept>

namespace tmpi {

// ------------------------------------------------------------------ driver entry points (no link-time libcuda dependency)
namespace drv {
#define TMPI_DRV_FN(name) static decltype(&::name) p_##name = nullptr;
TMPI_DRV_FN(cuMemCreate) TMPI_DRV_FN(cuMemRelease) TMPI_DRV_FN(cuMemExportToShareableHandle) TMPI_DRV_FN(cuMemImportFromShareableHandle)
TMPI_DRV_FN(cuMemAddressReserve) TMPI_DRV_FN(cuMemAddressFree) TMPI_DRV_FN(cuMemMap) TMPI_DRV_FN(cuMemUnmap) TMPI_DRV_FN(cuMemSetAccess)
TMPI_DRV_FN(cuMemGetAllocationGranularity) TMPI_DRV_FN(cuMulticastCreate) TMPI_DRV_FN(cuMulticastAddDevice) TMPI_DRV_FN(cuMulticastBindMem)
TMPI_DRV_FN(cuMulticastGetGranularity) TMPI_DRV_FN(cuDeviceGetAttribute) TMPI_DRV_FN(cuGetErrorString)
#undef TMPI_DRV_FN

template <typename F> static void load(F& f, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    (void)cudaGetLastError();
    throw std::runtime_error(std::string("tmpi_native: driver entry point unavailable: ") + name);
  }
  f = reinterpret_cast<F>(p);
}
static void init() {
  static std::once_flag once;
  std::call_once(once, [] {
#define L(n) load(p_##n, #n)
    L(cuMemCreate); L(cuMemRelease); L(cuMemExportToShareableHandle); L(cuMemImportFromShareableHandle); L(cuMemAddressReserve);
    L(cuMemAddressFree); L(cuMemMap); L(cuMemUnmap); L(cuMemSetAccess); L(cuMemGetAllocationGranularity); L(cuMulticastCreate);
    L(cuMulticastAddDevice); L(cuMulticastBindMem); L(cuMulticastGetGranularity); L(cuDeviceGetAttribute); L(cuGetErrorString);
#undef L
  });
}
static void check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    if (p_cuGetErrorString) p_cuGetErrorString(r, &s);
    throw std::runtime_error(std::string("tmpi_native: ") + what + " failed: " + (s ? s : "?") + " (" + std::to_string((int)r) + ")");
  }
}
}  // namespace drv

static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

// ------------------------------------------------------------------ unix-socket fd passing
static sockaddr_un make_addr(const std::string& job, int rank, socklen_t* len) {
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  std::string name = "tmpi-" + job + "-" + std::to_string(rank);
  a.sun_path[0] = '\0';                                   // abstract namespace: no filesystem entry to clean up
  strncpy(a.sun_path + 1, name.c_str(), sizeof(a.sun_path) - 2);
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
  return a;
}

static void send_fds(int sock, const int* fds, int nfds, const int* payload, int npayload) {
  msghdr msg; memset(&msg, 0, sizeof(msg));
  iovec io; io.iov_base = (void*)payload; io.iov_len = sizeof(int) * npayload;
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int) * 8)]; memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl; msg.msg_controllen = CMSG_SPACE(sizeof(int) * nfds);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int) * nfds);
  memcpy(CMSG_DATA(c), fds, sizeof(int) * nfds);
  if (sendmsg(sock, &msg, 0) < 0) throw std::runtime_error(std::string("tmpi_native: sendmsg(SCM_RIGHTS): ") + strerror(errno));
}

static int recv_fds(int sock, int* fds, int max_fds, int* payload, int npayload) {
  msghdr msg; memset(&msg, 0, sizeof(msg));
  iovec io; io.iov_base = payload; io.iov_len = sizeof(int) * npayload;
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int) * 8)]; memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, 0) < 0) throw std::runtime_error(std::string("tmpi_native: recvmsg: ") + strerror(errno));
  int n = 0;
  for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c)) {
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      n = (int)((c->cmsg_len - CMSG_LEN(0)) / sizeof(int));
      if (n > max_fds) n = max_fds;
      memcpy(fds, CMSG_DATA(c), sizeof(int) * n);
    }
  }
  return n;
}

// ------------------------------------------------------------------ PeerArena
struct PeerArena::Impl {
  CUmemGenericAllocationHandle h_arena = 0, h_sig = 0, h_mc = 0;
  CUmemGenericAllocationHandle peer_arena[kMaxRanks] = {0}, peer_sig[kMaxRanks] = {0};
  int fd_arena = -1, fd_sig = -1, fd_mc = -1;
  int listen_sock = -1;
  bool ipc_mode = false;
};

PeerArena::PeerArena(int rank, int world, int device, size_t arena_bytes, const std::string& job, bool force_ipc)
    : rank_(rank), world_(world), device_(device), job_(job), impl_(new Impl) {
  if (world < 1 || world > kMaxRanks) throw std::runtime_error("tmpi_native: world size must be 1..8 per node");
  for (int i = 0; i < kMaxRanks; ++i) { arena_[i] = nullptr; sig_[i] = nullptr; }
  check_cuda(cudaSetDevice(device), "cudaSetDevice");
  check_cuda(cudaFree(0), "context init");
  sig_bytes_ = (size_t)kMaxCommBlocks * kMaxRanks * 4 + (size_t)kMaxCommBlocks * 4 + 4096;
  bool vmm_ok = !force_ipc;
  if (vmm_ok) {
    try {
      drv::init();
      CUmemAllocationProp prop; memset(&prop, 0, sizeof(prop));
      prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      prop.location.id = device;
      prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      size_t gran = 0;
      drv::check(drv::p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
      gran_ = gran;
      if (world > 1) {
        // the multicast object (NVLS) wants the bound range to be a multiple of ITS granularity as well
        CUmulticastObjectProp mp; memset(&mp, 0, sizeof(mp));
        mp.numDevices = (unsigned)world; mp.size = round_up(arena_bytes, gran); mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        size_t mg = 0;
        if (drv::p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran_) gran_ = mg;
      }
      arena_bytes_ = round_up(arena_bytes, gran_);
      sig_bytes_ = round_up(sig_bytes_, gran);
      drv::check(drv::p_cuMemCreate(&impl_->h_arena, arena_bytes_, &prop, 0), "cuMemCreate(arena)");
      drv::check(drv::p_cuMemCreate(&impl_->h_sig, sig_bytes_, &prop, 0), "cuMemCreate(signal pad)");
      arena_[rank] = map_handle(impl_->h_arena, arena_bytes_);
      sig_[rank] = map_handle(impl_->h_sig, sig_bytes_);
      if (world > 1) {
        drv::check(drv::p_cuMemExportToShareableHandle(&impl_->fd_arena, impl_->h_arena, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export(arena)");
        drv::check(drv::p_cuMemExportToShareableHandle(&impl_->fd_sig, impl_->h_sig, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export(signal)");
      }
      mode_ = "vmm";
    } catch (const std::exception& e) {
      vmm_error_ = e.what();
      vmm_ok = false;
      if (impl_->h_arena) { /* leave cleanup to the destructor */ }
    }
  }
  if (!vmm_ok) {
    impl_->ipc_mode = true;
    mode_ = "ipc";
    arena_bytes_ = round_up(arena_bytes, 2u << 20);
    sig_bytes_ = round_up(sig_bytes_, 2u << 20);
    check_cuda(cudaMalloc(&arena_[rank], arena_bytes_), "cudaMalloc(arena)");
    check_cuda(cudaMalloc(&sig_[rank], sig_bytes_), "cudaMalloc(signal)");
  }
  check_cuda(cudaMemset(arena_[rank], 0, arena_bytes_), "memset arena");
  check_cuda(cudaMemset(sig_[rank], 0, sig_bytes_), "memset signal");
  check_cuda(cudaDeviceSynchronize(), "sync");
  if (world > 1 && !impl_->ipc_mode) {
    impl_->listen_sock = socket(AF_UNIX, SOCK_STREAM, 0);
    if (impl_->listen_sock < 0) throw std::runtime_error("tmpi_native: socket() failed");
    socklen_t len; sockaddr_un a = make_addr(job_, rank_, &len);
    if (bind(impl_->listen_sock, (sockaddr*)&a, len) < 0) throw std::runtime_error(std::string("tmpi_native: bind: ") + strerror(errno));
    if (listen(impl_->listen_sock, 64) < 0) throw std::runtime_error("tmpi_native: listen failed");
  }
}

void* PeerArena::map_handle(unsigned long long handle, size_t bytes) {
  CUdeviceptr va = 0;
  drv::check(drv::p_cuMemAddressReserve(&va, bytes, gran_, 0, 0), "cuMemAddressReserve");
  drv::check(drv::p_cuMemMap(va, bytes, 0, (CUmemGenericAllocationHandle)handle, 0), "cuMemMap");
  CUmemAccessDesc acc; memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  drv::check(drv::p_cuMemSetAccess(va, bytes, &acc, 1), "cuMemSetAccess");
  return reinterpret_cast<void*>(va);
}

static int connect_to(const std::string& job, int peer) {
  int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) throw std::runtime_error("tmpi_native: socket() failed");
  socklen_t len; sockaddr_un a = make_addr(job, peer, &len);
  for (int attempt = 0; attempt < 600; ++attempt) {
    if (connect(s, (sockaddr*)&a, len) == 0) return s;
    usleep(50000);
  }
  close(s);
  throw std::runtime_error("tmpi_native: could not connect to peer " + std::to_string(peer) + ": " + strerror(errno));
}

// phase 1 (after every rank constructed its arena): push my fds to one peer
void PeerArena::send_handles_to(int peer) {
  if (impl_->ipc_mode || peer == rank_) return;
  int s = connect_to(job_, peer);
  int fds[2] = {impl_->fd_arena, impl_->fd_sig};
  int payload[2] = {rank_, 0};
  send_fds(s, fds, 2, payload, 2);
  close(s);
}

// phase 2: accept world-1 connections, import + map what arrives
void PeerArena::recv_handles() {
  if (impl_->ipc_mode) return;
  for (int i = 0; i < world_ - 1; ++i) {
    int c = accept(impl_->listen_sock, nullptr, nullptr);
    if (c < 0) throw std::runtime_error(std::string("tmpi_native: accept: ") + strerror(errno));
    int fds[2] = {-1, -1}; int payload[2] = {-1, -1};
    int n = recv_fds(c, fds, 2, payload, 2);
    close(c);
    const int src = payload[0];
    if (n != 2 || src < 0 || src >= world_ || src == rank_) throw std::runtime_error("tmpi_native: bad handle message");
    if (payload[1] == 1) {           // multicast handle message
      impl_->fd_mc = fds[0]; if (fds[1] >= 0) close(fds[1]);
      --i;                           // does not count as a peer arena
      continue;
    }
    drv::check(drv::p_cuMemImportFromShareableHandle(&impl_->peer_arena[src], (void*)(uintptr_t)fds[0], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import(arena)");
    drv::check(drv::p_cuMemImportFromShareableHandle(&impl_->peer_sig[src], (void*)(uintptr_t)fds[1], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import(signal)");
    close(fds[0]); close(fds[1]);
    arena_[src] = map_handle(impl_->peer_arena[src], arena_bytes_);
    sig_[src] = map_handle(impl_->peer_sig[src], sig_bytes_);
  }
}

// cudaIpc fallback: handles are exchanged by the Python control plane as raw bytes
std::string PeerArena::ipc_handles() const {
  cudaIpcMemHandle_t h[2];
  check_cuda(cudaIpcGetMemHandle(&h[0], arena_[rank_]), "cudaIpcGetMemHandle(arena)");
  check_cuda(cudaIpcGetMemHandle(&h[1], sig_[rank_]), "cudaIpcGetMemHandle(signal)");
  return std::string(reinterpret_cast<const char*>(h), sizeof(h));
}
void PeerArena::ipc_open(int peer, const std::string& handles) {
  if (peer == rank_) return;
  if (handles.size() != 2 * sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("tmpi_native: bad ipc handle blob");
  cudaIpcMemHandle_t h[2];
  memcpy(h, handles.data(), sizeof(h));
  check_cuda(cudaIpcOpenMemHandle(&arena_[peer], h[0], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(arena)");
  check_cuda(cudaIpcOpenMemHandle(&sig_[peer], h[1], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(signal)");
}

// ------------------------------------------------------------------ NVLS multicast
bool PeerArena::multicast_supported() const {
  if (impl_->ipc_mode) return false;
  try {
    drv::init();
    int v = 0;
    if (drv::p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device_) != CUDA_SUCCESS) return false;
    return v != 0;
  } catch (...) { return false; }
}

static CUmulticastObjectProp mc_prop(int world, size_t bytes) {
  CUmulticastObjectProp p; memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)world; p.size = bytes; p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; p.flags = 0;
  return p;
}

// rank 0: create the multicast object and ship its fd to every peer (peers pick it up in mc_recv)
void PeerArena::mc_create_and_send() {
  if (rank_ != 0) return;
  CUmulticastObjectProp p = mc_prop(world_, arena_bytes_);
  size_t g = 0;
  drv::check(drv::p_cuMulticastGetGranularity(&g, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  if (arena_bytes_ % g) throw std::runtime_error("tmpi_native: arena size not a multiple of the multicast granularity");
  drv::check(drv::p_cuMulticastCreate(&impl_->h_mc, &p), "cuMulticastCreate");
  drv::check(drv::p_cuMemExportToShareableHandle(&impl_->fd_mc, impl_->h_mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export(multicast)");
  for (int peer = 1; peer < world_; ++peer) {
    int s = connect_to(job_, peer);
    int fds[2] = {impl_->fd_mc, impl_->fd_mc};
    int payload[2] = {0, 1};
    send_fds(s, fds, 2, payload, 2);
    close(s);
  }
}
void PeerArena::mc_recv() {
  if (rank_ == 0) return;
  int c = accept(impl_->listen_sock, nullptr, nullptr);
  if (c < 0) throw std::runtime_error("tmpi_native: accept(mc) failed");
  int fds[2] = {-1, -1}; int payload[2] = {-1, -1};
  int n = recv_fds(c, fds, 2, payload, 2);
  close(c);
  if (n < 1 || payload[1] != 1) throw std::runtime_error("tmpi_native: bad multicast handle message");
  if (n == 2 && fds[1] >= 0) close(fds[1]);
  drv::check(drv::p_cuMemImportFromShareableHandle(&impl_->h_mc, (void*)(uintptr_t)fds[0], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import(multicast)");
  close(fds[0]);
}
void PeerArena::mc_add_device() { drv::check(drv::p_cuMulticastAddDevice(impl_->h_mc, device_), "cuMulticastAddDevice"); }
// after ALL ranks added their device
void PeerArena::mc_bind_and_map() {
  drv::check(drv::p_cuMulticastBindMem(impl_->h_mc, 0, impl_->h_arena, 0, arena_bytes_, 0), "cuMulticastBindMem");
  mc_ = map_handle(impl_->h_mc, arena_bytes_);
}

CommCtx PeerArena::ctx() const {
  CommCtx c; memset(&c, 0, sizeof(c));
  for (int p = 0; p < kMaxRanks; ++p) { c.arena[p] = arena_[p]; c.sig[p] = reinterpret_cast<uint32_t*>(sig_[p]); }
  c.epoch = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sig_[rank_]) + (size_t)kMaxCommBlocks * kMaxRanks * 4);
  c.mc_arena = mc_;
  c.rank = rank_; c.world = world_;
  // a rank stuck behind a slow disk / loader / debugger must not kill the job: the device-side flag barriers spin for
  // TMPI_BARRIER_TIMEOUT_S seconds (default 600; at ~2 GHz) before they trap
  static const long long limit = [] {
    const char* e = getenv("TMPI_BARRIER_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (!(s > 0)) s = 600.0;
    return (long long)(s * 2.0e9);
  }();
  c.spin_limit = limit;
  return c;
}

PeerArena::~PeerArena() {
  cudaDeviceSynchronize();
  if (impl_->listen_sock >= 0) close(impl_->listen_sock);
  if (impl_->fd_arena >= 0) close(impl_->fd_arena);
  if (impl_->fd_sig >= 0) close(impl_->fd_sig);
  if (impl_->fd_mc >= 0 && rank_ == 0) close(impl_->fd_mc);
  if (impl_->ipc_mode) {
    for (int p = 0; p < world_; ++p) {
      if (p == rank_) continue;
      if (arena_[p]) cudaIpcCloseMemHandle(arena_[p]);
      if (sig_[p]) cudaIpcCloseMemHandle(sig_[p]);
    }
    if (arena_[rank_]) cudaFree(arena_[rank_]);
    if (sig_[rank_]) cudaFree(sig_[rank_]);
  } else if (drv::p_cuMemUnmap) {
    auto unmap = [&](void* p, size_t n) { if (p) { drv::p_cuMemUnmap((CUdeviceptr)p, n); drv::p_cuMemAddressFree((CUdeviceptr)p, n); } };
    if (mc_) unmap(mc_, arena_bytes_);
    for (int p = 0; p < world_; ++p) { unmap(arena_[p], arena_bytes_); unmap(sig_[p], sig_bytes_); }
    for (int p = 0; p < world_; ++p) { if (impl_->peer_arena[p]) drv::p_cuMemRelease(impl_->peer_arena[p]); if (impl_->peer_sig[p]) drv::p_cuMemRelease(impl_->peer_sig[p]); }
    if (impl_->h_mc) drv::p_cuMemRelease(impl_->h_mc);
    if (impl_->h_arena) drv::p_cuMemRelease(impl_->h_arena);
    if (impl_->h_sig) drv::p_cuMemRelease(impl_->h_sig);
  }
  delete impl_;
}

}  // namespace tmpi
