// ab_common.h — internal helpers of libaesara_b200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "../../include/aesara_b200.h"

namespace ab {

std::string& last_error();
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

struct Module {
  cudaLibrary_t lib = nullptr;
  // lazily resolved kernels (names fixed by the skeletons in csrc/*.cuh)
  enum { EW_FLAT_VEC, EW_FLAT, EW_ROWS_VEC, EW_ROWS, EW_ND, EW_TILE, RED_ROWS, RED_ROWS_P, RED_ROWS_F,
         RED_COLS, RED_COLS_P, RED_COLS_F, N_KERNELS };
  cudaKernel_t k[N_KERNELS] = {};
  bool tried[N_KERNELS] = {};
  std::map<std::string, cudaKernel_t> named;  // kernels launched by name (ab_kernel_launch)
  int get(int which, cudaKernel_t* out);
};

}  // namespace ab

#define AB_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return ab::fail(AB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                      __FILE__, __LINE__);                                              \
  } while (0)
