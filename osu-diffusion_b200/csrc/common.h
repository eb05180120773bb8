// Shared host-side helpers for the C-ABI library (error slot, device properties).
#pragma once
#include <cuda_runtime.h>

namespace osudit {

// Records a message in the thread-local error slot read by osudit_last_error(); returns `code`.
int set_error(int code, const char* msg);
int num_sms();

#define OSUDIT_CHECK_LAUNCH()                                         \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return set_error(-6, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace osudit
