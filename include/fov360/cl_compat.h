// cl_compat.h - the few OpenCL C++-wrapper names the reference's call sites use, mapped onto the
// fov360 C ABI, so that video_server.cc / video_client.cc / run_satlogrectilinear.cc keep their
// buffer-management code unchanged when OpenCLManager is replaced:
//
//   cl::Buffer buf(cl_manager.context, CL_MEM_READ_WRITE, nbytes);          video_server.cc:224-232
//   cl::copy(cl_manager.command_queue, host_begin, host_end, buf);          video_server.cc:297-299
//   cl::copy(cl_manager.command_queue, buf, host_begin, host_end);          video_server.cc:342-345
//   clFlush(cl_manager.command_queue()); clFinish(cl_manager.command_queue());        :302-303
//   encoder.EncodeFrameGPU(sat_buf(), src_buf(), w, h, linesize);           video_server.cc:300
//
// Only what those call sites need exists here; this is not an OpenCL implementation.
#pragma once
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <iterator>
#include <memory>

#include "../fov360.h"

typedef void *cl_mem;  // a device pointer obtained from fov_malloc
typedef int cl_int;
typedef uint64_t cl_mem_flags;
#ifndef CL_SUCCESS
#define CL_SUCCESS 0
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#endif

namespace cl {

// Holds the fov_ctx; shared by Context / CommandQueue / Buffer like the OpenCL handles are.
struct ContextState {
  fov_ctx *ctx = nullptr;
  ~ContextState() { fov_ctx_destroy(ctx); }
};

class Context {
 public:
  Context() = default;
  explicit Context(std::shared_ptr<ContextState> s) : state_(std::move(s)) {}
  fov_ctx *operator()() const { return state_ ? state_->ctx : nullptr; }
  const std::shared_ptr<ContextState> &state() const { return state_; }

 private:
  std::shared_ptr<ContextState> state_;
};

// The in-order command queue: the context's CUDA stream.
class CommandQueue {
 public:
  CommandQueue() = default;
  explicit CommandQueue(const Context &c) : state_(c.state()) {}
  fov_ctx *operator()() const { return state_ ? state_->ctx : nullptr; }
  cl_int flush() const { return CL_SUCCESS; }  // launches are submitted eagerly
  cl_int finish() const { return fov_sync((*this)()); }

 private:
  std::shared_ptr<ContextState> state_;
};

class Buffer {
 public:
  Buffer() = default;
  Buffer(const Context &context, cl_mem_flags /*flags*/, size_t size, void * /*host_ptr*/ = nullptr,
         cl_int *err = nullptr)
      : block_(std::make_shared<Block>()) {
    block_->owner = context.state();
    block_->size = size;
    const int rc = fov_malloc(context(), &block_->ptr, size);
    if (rc != FOV_OK) {
      std::cerr << "cl::Buffer: " << fov_last_error_string(context()) << std::endl;
      block_->ptr = nullptr;
    }
#ifdef FOV360_ZERO_NEW_BUFFERS  // deterministic contents for hash-based checks (OpenCL leaves them undefined)
    if (rc == FOV_OK) fov_memset(context(), block_->ptr, 0, size);
#endif
    if (err) *err = rc;
  }
  cl_mem operator()() const { return block_ ? block_->ptr : nullptr; }
  size_t size() const { return block_ ? block_->size : 0; }

 private:
  struct Block {
    std::shared_ptr<ContextState> owner;
    void *ptr = nullptr;
    size_t size = 0;
    ~Block() {
      if (ptr && owner) fov_free(owner->ctx, ptr);
    }
  };
  std::shared_ptr<Block> block_;
};

// Blocking host -> device copy of a contiguous range (cl::copy(queue, first, last, buffer)).
template <class It>
inline cl_int copy(const CommandQueue &q, It first, It last, const Buffer &buffer) {
  typedef typename std::iterator_traits<It>::value_type T;
  const size_t n = static_cast<size_t>(std::distance(first, last)) * sizeof(T);
  return n ? fov_memcpy_h2d(q(), buffer(), &*first, n) : CL_SUCCESS;
}

// Blocking device -> host copy (cl::copy(queue, buffer, first, last)).
template <class It>
inline cl_int copy(const CommandQueue &q, const Buffer &buffer, It first, It last) {
  typedef typename std::iterator_traits<It>::value_type T;
  const size_t n = static_cast<size_t>(std::distance(first, last)) * sizeof(T);
  return n ? fov_memcpy_d2h(q(), &*first, buffer(), n) : CL_SUCCESS;
}

}  // namespace cl

inline cl_int clFlush(fov_ctx *) { return CL_SUCCESS; }
inline cl_int clFinish(fov_ctx *queue) { return fov_sync(queue); }
