// Host-side helpers shared by every translation unit of libdexbotic_b200.so:
// thread-local error message, launch counter, CUDA error checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define B200_CHECK(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::b200::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

#define B200_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

// call after every kernel launch
#define B200_LAUNCH_OK()                                                                     \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::b200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                              \
    }                                                                                        \
    ::b200::count_launch();                                                                  \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace b200
