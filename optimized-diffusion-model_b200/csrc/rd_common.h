// rd_common.h -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "../../include/rdb200.h"

namespace rd {

// thread-local last-error buffer behind rd_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return RD_OK;
}

constexpr int kNumSMs = 148;  // B200; grids for grid-stride kernels are sized in multiples of this

}  // namespace rd

#define RD_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::rd::fail(RD_E_INVALID, __VA_ARGS__);  \
  } while (0)
