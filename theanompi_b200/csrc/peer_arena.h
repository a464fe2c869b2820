#pragma once
#include <string>
#include "api.h"

namespace tmpi {

class PeerArena {
 public:
  PeerArena(int rank, int world, int device, size_t arena_bytes, const std::string& job, bool force_ipc);
  ~PeerArena();
  PeerArena(const PeerArena&) = delete;
  PeerArena& operator=(const PeerArena&) = delete;

  // VMM path: fd exchange over unix sockets
  void send_handles_to(int peer);
  void recv_handles();
  // cudaIpc fallback: opaque handle bytes moved by the Python control plane
  std::string ipc_handles() const;
  void ipc_open(int peer, const std::string& handles);
  // NVLS multicast (VMM path only)
  bool multicast_supported() const;
  void mc_create_and_send();
  void mc_recv();
  void mc_add_device();
  void mc_bind_and_map();

  CommCtx ctx() const;
  void* arena_ptr(int p) const { return arena_[p]; }
  void* sig_ptr(int p) const { return sig_[p]; }
  void* mc_ptr() const { return mc_; }
  size_t arena_bytes() const { return arena_bytes_; }
  size_t sig_bytes() const { return sig_bytes_; }
  const std::string& mode() const { return mode_; }
  const std::string& vmm_error() const { return vmm_error_; }
  int rank() const { return rank_; }
  int world() const { return world_; }

 private:
  void* map_handle(unsigned long long handle, size_t bytes);
  struct Impl;
  int rank_, world_, device_;
  std::string job_;
  Impl* impl_;
  void* arena_[kMaxRanks];
  void* sig_[kMaxRanks];
  void* mc_ = nullptr;
  size_t arena_bytes_ = 0, sig_bytes_ = 0, gran_ = 2u << 20;
  std::string mode_, vmm_error_;
};

}  // namespace tmpi
