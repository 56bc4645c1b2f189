// Host-side helpers shared by all translation units of libvitk: error reporting (thread-local
// message behind vitk_last_error), launch checks, device properties.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitk.h"

namespace vitk {

int set_error(int code, const char* fmt, ...);   // stores the message, returns code
int cuda_error(cudaError_t e, const char* what); // returns (int)e after storing the message
int num_sms();                                   // SM count of the current device (cached)
void count_launch();                             // bumps the process-wide kernel launch counter

#define VITK_REQUIRE(cond, code, ...)                         \
  do {                                                        \
    if (!(cond)) return ::vitk::set_error((code), __VA_ARGS__); \
  } while (0)

#define VITK_CUDA(expr)                                             \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return ::vitk::cuda_error(_e, #expr);    \
  } while (0)

#define VITK_LAUNCH_CHECK(name)                                     \
  do {                                                              \
    ::vitk::count_launch();                                         \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::vitk::cuda_error(_e, name);     \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace vitk
