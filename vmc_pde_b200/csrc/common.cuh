// Shared helpers: error reporting for the C-ABI, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <string>
#include "../../include/vmcpde.h"

namespace vmc {

std::string& last_error_ref();
int set_error(int code, const std::string& msg);

#define VMC_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::vmc::set_error(VMCPDE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define VMC_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::vmc::set_error(VMCPDE_ECUDA, std::string(name) + " launch: " + cudaGetErrorString(_e)); \
  } while (0)

#define VMC_REQUIRE(cond, msg)                                           \
  do {                                                                   \
    if (!(cond)) return ::vmc::set_error(VMCPDE_EINVAL, std::string(msg)); \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace vmc
