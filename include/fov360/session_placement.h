// Placement of client sessions on the GPUs of one box.
//
// The reference creates one OpenCLManager per accepted connection and always binds device 0
// (video_server.cc:62-66, opencl_manager.cc:13-34).  On an 8-GPU box every connection's frames,
// SATs and tables live on exactly one GPU for the connection's whole life and nothing is shared
// between connections, so placement is the only multi-GPU decision: a new session goes to the
// least-loaded device (ties -> lowest index, which is s % G while nobody disconnects - the static
// rule the benchmark uses), and gives its slot back when the connection closes.
//
//   static SessionPlacement placement(fov_device_count());      // one per server process
//   OpenCLManager cl_manager;
//   cl_manager.device_index = placement.Acquire();              // video_server.cc:62
//   cl_manager.InitializeContext();
//   ...
//   placement.Release(cl_manager.device_index);                 // connection closed
//
// Thread-safe: the reference serves each connection on its own thread (video_server.cc:85).
#pragma once
#include <mutex>
#include <vector>

class SessionPlacement {
 public:
  explicit SessionPlacement(int device_count) : load_(device_count > 0 ? device_count : 1, 0) {}

  int Acquire() {
    std::lock_guard<std::mutex> lock(mu_);
    int best = 0;
    for (int d = 1; d < static_cast<int>(load_.size()); ++d)
      if (load_[d] < load_[best]) best = d;
    ++load_[best];
    return best;
  }

  void Release(int device) {
    std::lock_guard<std::mutex> lock(mu_);
    if (device >= 0 && device < static_cast<int>(load_.size()) && load_[device] > 0) --load_[device];
  }

  int Sessions(int device) const {
    std::lock_guard<std::mutex> lock(mu_);
    return device >= 0 && device < static_cast<int>(load_.size()) ? load_[device] : 0;
  }

  int DeviceCount() const { return static_cast<int>(load_.size()); }

 private:
  mutable std::mutex mu_;
  std::vector<int> load_;
};
