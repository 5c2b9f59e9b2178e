// common.cuh — error plumbing and launch accounting shared by the libpnr translation units.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/pnr.h"

namespace pnr {

int set_error(int code, const char* fmt, ...);   // stores a thread-local message, returns code
void count_launch(int n = 1);

#define PNR_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return ::pnr::set_error(PNR_ERR_ARG, __VA_ARGS__); } while (0)

#define PNR_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::pnr::set_error(PNR_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                    \
                              cudaGetErrorString(e__), __FILE__, __LINE__);                    \
  } while (0)

#define PNR_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess)                                                                    \
      return ::pnr::set_error(PNR_ERR_CUDA, "launch of %s failed: %s", name,                   \
                              cudaGetErrorString(e__));                                        \
    ::pnr::count_launch();                                                                     \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pnr
