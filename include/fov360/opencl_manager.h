// Drop-in for the reference's src/opencl_manager.h (opencl_manager.h:8-22, .cc:7-67): same class
// name and public members the foveation call sites touch (`context`, `command_queue`,
// `InitializeContext()`, `GetCLErrorString()`), backed by a fov360 CUDA context.  The reference
// always binds device 0 of the first NVIDIA platform (opencl_manager.cc:13-34); set `device_index`
// before InitializeContext() to place a connection on another GPU of the box.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>

#include "cl_compat.h"

class OpenCLManager {
 public:
  cl::Context context;
  cl::CommandQueue command_queue;
  int device_index = 0;
  // Kept so that client code that sets them compiles; CUDA/GL interop is out of scope here.
  long gl_context = -1;
  long gl_display = -1;

  OpenCLManager() = default;
  ~OpenCLManager() = default;

  // opencl_manager.cc:7-67: fatal when no device can be used (there is no CPU fallback).
  int InitializeContext() {
    int err = 0;
    auto state = std::make_shared<cl::ContextState>();
    state->ctx = fov_ctx_create(device_index, &err);
    if (!state->ctx) {
      std::cerr << "Failed to create a CUDA context: " << fov_last_error_string(nullptr)
                << std::endl;
      std::exit(EXIT_FAILURE);
    }
    context = cl::Context(state);
    command_queue = cl::CommandQueue(context);
    return 0;
  }

  fov_ctx *handle() const { return context(); }

  static std::string GetCLErrorString(cl_int error) {
    switch (error) {
      case FOV_OK: return "FOV_OK";
      case FOV_ERR_NO_CONTEXT: return "FOV_ERR_NO_CONTEXT";
      case FOV_ERR_INVALID: return "FOV_ERR_INVALID";
      case FOV_ERR_NO_DEVICE: return "FOV_ERR_NO_DEVICE";
      case FOV_ERR_GRID: return "FOV_ERR_GRID";
      case FOV_ERR_UNSUPPORTED: return "FOV_ERR_UNSUPPORTED";
      default: return "cudaError " + std::to_string(error);
    }
  }
};
